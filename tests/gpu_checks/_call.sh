cd $GRAFT_REPO_ROOT
KGB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spSelectKernel -s 6 -c 1 -o gpurun_out/r02_prof_select python tests/gpu_checks/wave_for_ncu.py 8 > gpurun_out/r02_c8_ncu.log 2>&1
KGB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spBackupKernel -s 6 -c 1 -o gpurun_out/r02_prof_backup python tests/gpu_checks/wave_for_ncu.py 8 >> gpurun_out/r02_c8_ncu.log 2>&1
(cd oracle/_ref && timeout 300 ./katago_cuda runtinynntests /tmp 1.0 2>&1 | tail -8) > gpurun_out/r02_c8_cuda_tinynn.log 2>&1
tail -5 gpurun_out/r02_c8_ncu.log; cat gpurun_out/r02_c8_cuda_tinynn.log
