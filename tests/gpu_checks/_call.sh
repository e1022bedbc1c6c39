cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c34.log
timeout 240 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > $L
cat $L
