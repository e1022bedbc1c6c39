cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c20.log
timeout 600 python tests/gpu_checks/pingpong.py 60 2>&1 | tail -6 > $L
timeout 600 python tests/gpu_checks/trained_net_loop.py 256 1200 2>&1 | tail -2 >> $L
timeout 900 python -m pytest tests/test_gpu_board_selfplay.py tests/test_gpu_mixed_sizes.py -q -x 2>&1 | tail -4 >> $L
cat $L | cut -c1-1500
