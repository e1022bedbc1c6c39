cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c22.log
timeout 600 python -m pytest tests/test_game_recorder.py tests/test_selfplay_cli.py -q -x -m gpu -k "limits or cheap" 2>&1 | tail -30 > $L
echo "== full gpu suite" >> $L
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 >> $L
cat $L | cut -c1-4000
