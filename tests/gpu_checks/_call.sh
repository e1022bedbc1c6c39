cd $GRAFT_REPO_ROOT
echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null) nproc: $(nproc) affinity: $(python -c 'import os; print(len(os.sched_getaffinity(0)))') load: $(cat /proc/loadavg)" > gpurun_out/r02_c11.log
timeout 600 python -m pytest tests/test_gpu_board_selfplay.py -x -q -k "playouts_per_wave or nn_cache or ladder_node_budget" 2>&1 | tail -4 >> gpurun_out/r02_c11.log
for mp in 16 3 2; do for cap in 256 512; do timeout 600 python bench.py --steps 40 --warmup 3 --ladder-nodes-per-wave $cap --max-playouts-per-wave $mp 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('maxplayouts', $mp, 'cap', $cap, d['value'], d['ms_per_step'], d['config']['nn_only_ms_per_step'], d['roofline_tree']['ms_select'], d['config']['ladder']['game_waves_without_leaf'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"; done; done >> gpurun_out/r02_c11.log 2>&1
cat gpurun_out/r02_c11.log
