cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_game_recorder.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_c14.log
timeout 1500 python bench.py --model b28c512nbt --games 1024 --visits 1600 --steps 10 --warmup 3 --settle-waves 200 > gpurun_out/r02_c14_bench_b28.json 2> gpurun_out/r02_c14_bench_b28.err
python - <<'PY' >> gpurun_out/r02_c14.log
import json
try:
    d=json.loads(open('gpurun_out/r02_c14_bench_b28.json').read().strip().splitlines()[-1])
    print('b28 value', d['value'], 'ms/step', d['ms_per_step'], 'nn', d['config']['nn_only_ms_per_step'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_forward_tflops'])
    print('e2e', d['e2e']['value'], d['e2e']['frac_of_value'], 'fp32', d['value_fp32_equivalent'], 'cpu', d['cpu_baseline']['value'], 'tree', d['roofline_tree']['ms_select'], d['roofline_tree']['ms_backup'])
    print('competitor', d['config']['gpu_competitor'])
except Exception as e:
    print('bench failed', e)
PY
cat gpurun_out/r02_c14.log; tail -5 gpurun_out/r02_c14_bench_b28.err
