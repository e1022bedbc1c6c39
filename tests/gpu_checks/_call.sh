cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c24.log
timeout 600 python -m pytest tests/test_gpu_mixed_sizes.py -q -x -k "policy_init" 2>&1 | tail -25 > $L
timeout 600 python -m pytest tests/test_selfplay_cli.py tests/test_game_recorder.py tests/test_match_and_gatekeeper.py -q -x -m gpu 2>&1 | tail -25 >> $L
cat $L | cut -c1-5000
