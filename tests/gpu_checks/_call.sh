cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c27.log
timeout 600 python -m pytest tests/test_komi_search.py -q -x -m gpu -s 2>&1 | tail -25 > $L
timeout 600 python -m pytest tests/test_selfplay_cli.py -q -x -m gpu -k "fair_komi" 2>&1 | tail -30 >> $L
cat $L | cut -c1-6000
