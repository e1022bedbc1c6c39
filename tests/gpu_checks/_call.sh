cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c19.log
timeout 600 python -m pytest tests/test_gpu_mixed_sizes.py -q -x 2>&1 | tail -25 > $L
timeout 300 python -m pytest tests/test_selfplay_cli.py -q -m gpu -x -k mixed 2>&1 | tail -25 >> $L
echo "== full gpu suite" >> $L
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 >> $L
echo "== bench" >> $L
timeout 600 python bench.py 2>&1 | tail -3 >> $L
cat $L | cut -c1-6000
