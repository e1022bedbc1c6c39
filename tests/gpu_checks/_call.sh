cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_game_recorder.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r02_c13.log
cat gpurun_out/r02_c13.log
