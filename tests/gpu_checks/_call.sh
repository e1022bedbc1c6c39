cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c25.log
timeout 600 python -m pytest tests/test_gpu_mixed_sizes.py -q -x -k "policy_init" 2>&1 | tail -8 > $L
for m in 0 2 3 4 6; do timeout 300 python tests/gpu_checks/trained_net_loop.py 256 800 $m 2>&1 | tail -1 >> $L; done
cat $L | cut -c1-6000
