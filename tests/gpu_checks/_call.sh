cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_board_selfplay.py -q -k "trained_net or fake_net" 2>&1 | tail -40 > gpurun_out/r02_c16.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 >> gpurun_out/r02_c16.log
cat gpurun_out/r02_c16.log
