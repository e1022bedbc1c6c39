cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c21.log
timeout 900 python -m pytest tests/test_match_and_gatekeeper.py -q -x -m gpu 2>&1 | tail -30 > $L
timeout 900 python -m pytest tests/test_game_recorder.py tests/test_selfplay_cli.py -q -m gpu 2>&1 | tail -5 >> $L
cat $L | cut -c1-3000
