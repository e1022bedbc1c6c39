cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_c7_gpu_suite.log
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r02_c7_bench.json 2> gpurun_out/r02_c7_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c7_smoke.log 2>&1
cat gpurun_out/r02_c7_gpu_suite.log; cat gpurun_out/r02_c7_bench.json | head -c 3000; tail -3 gpurun_out/r02_c7_bench.err; tail -5 gpurun_out/r02_c7_smoke.log
