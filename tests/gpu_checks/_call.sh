cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c23.log
KGB_NO_GRAPH=1 timeout 900 ncu --kernel-name regex:"spSelectKernel|spBackupKernel" --launch-skip 500 --launch-count 12 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02_tree_kernels_ncu.csv python tests/gpu_checks/wave_for_ncu.py 262 > gpurun_out/r02_tree_ncu_stdout.log 2>&1
tail -3 gpurun_out/r02_tree_ncu_stdout.log > $L
wc -l gpurun_out/r02_tree_kernels_ncu.csv >> $L
timeout 600 python bench.py --mixed-sizes --steps 30 2>&1 | tail -1 > gpurun_out/bench_r02_config4_mixed_sizes.json
cut -c1-1200 gpurun_out/bench_r02_config4_mixed_sizes.json >> $L
timeout 300 python -m pytest tests/test_selfplay_cli.py -q -m gpu -k cheap 2>&1 | tail -3 >> $L
cat $L | cut -c1-3000
