cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c26.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 >> $L
cat $L | cut -c1-4000
