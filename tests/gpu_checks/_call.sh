cd $GRAFT_REPO_ROOT
timeout 900 python tests/gpu_checks/conv_variants.py - KGB_T3_E=3 KGB_CONV_DBG=4 KGB_CONV_DBG=2 KGB_CONV_DBG=6 > gpurun_out/r02_c4_variants.log 2>&1
for e in "" "KGB_T3_E=3"; do echo "== $e"; env $e timeout 300 python tests/gpu_checks/perf_nn.py b18c384nbt 256 30; done > gpurun_out/r02_c4_perf_nn.log 2>&1
(echo "== contiguous-row 1x1 64->64, batch 2048 (raw+act)"; timeout 120 python tests/gpu_checks/conv_one.py 1 64 64 2048 3 1; timeout 120 python tests/gpu_checks/conv_one.py 1 384 192 256 3 1; timeout 120 python tests/gpu_checks/conv_one.py 1 192 384 256 2 1) > gpurun_out/r02_c4_layout_probe.log 2>&1
timeout 600 python -m pytest tests/test_gpu_nn_parity.py -x -q 2>&1 | tail -5 > gpurun_out/r02_c4_tests_nn_all.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kgb_conv_tc3 -s 3 -c 2 -o gpurun_out/r02_prof_pre1x1 python tests/gpu_checks/conv_one.py 1 384 192 256 3 1 > gpurun_out/r02_c4_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kgb_conv_tc3 -s 3 -c 2 -o gpurun_out/r02_prof_3x3 python tests/gpu_checks/conv_one.py 3 192 192 256 1 1 >> gpurun_out/r02_c4_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kgb_conv_tc3 -s 3 -c 2 -o gpurun_out/r02_prof_post1x1 python tests/gpu_checks/conv_one.py 1 192 384 256 2 1 >> gpurun_out/r02_c4_ncu.log 2>&1
(cd oracle/_ref && timeout 200 ./katago_cuda runtinynntests 2>&1 | tail -5) > gpurun_out/r02_c4_cuda_tinynn.log 2>&1
cat gpurun_out/r02_c4_variants.log gpurun_out/r02_c4_perf_nn.log gpurun_out/r02_c4_layout_probe.log gpurun_out/r02_c4_tests_nn_all.log gpurun_out/r02_c4_cuda_tinynn.log; tail -3 gpurun_out/r02_c4_ncu.log
