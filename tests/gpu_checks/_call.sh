cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c31.log
timeout 600 python -m pytest tests/test_match_and_gatekeeper.py -q -x -m gpu -s 2>&1 | grep -v "^Game \|^\[config\]\|^Loaded\|^Moving\|^Found" | tail -25 > $L
cat $L | cut -c1-4000
