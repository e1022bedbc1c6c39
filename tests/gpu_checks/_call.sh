cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c30.log
timeout 900 python bench.py --steps 50 2>&1 | tail -1 > gpurun_out/bench_r02_final_n1.json
cut -c1-400 gpurun_out/bench_r02_final_n1.json > $L
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600 >> $L
cat $L
