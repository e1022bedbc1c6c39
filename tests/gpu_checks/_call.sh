cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c28.log
timeout 900 python tests/gpu_checks/soak_cli.py 48 64 > gpurun_out/r02_soak_cli_stdout.log 2> gpurun_out/r02_soak_cli_stderr.log
tail -2 gpurun_out/r02_soak_cli_stdout.log > $L
tail -5 gpurun_out/r02_soak_cli_stderr.log | cut -c1-600 >> $L
cat $L | cut -c1-4000
