cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c29.log
timeout 600 python -m pytest tests/test_selfplay_cli.py -q -x -m gpu -k "forks" 2>&1 | tail -30 > $L
timeout 600 python -m pytest tests/test_gpu_board_selfplay.py -q -x -k "symmetry" 2>&1 | tail -8 >> $L
timeout 900 python -m pytest tests/test_selfplay_cli.py tests/test_game_recorder.py tests/test_komi_search.py tests/test_match_and_gatekeeper.py -q -m gpu 2>&1 | tail -8 >> $L
cat $L | cut -c1-6000
