cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c33.log
timeout 400 python -m pytest tests/test_selfplay_cli.py -q -m gpu 2>&1 | tail -30 > $L
cat $L | cut -c1-5000
