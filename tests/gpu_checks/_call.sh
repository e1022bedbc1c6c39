cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_c18.log
timeout 900 python -m pytest tests/test_gpu_board_selfplay.py -q -k "komi or fake_net or trained_net" -x 2>&1 | tail -15 > $L
timeout 600 python -m pytest tests/test_gpu_weight_swap.py tests/test_selfplay_cli.py tests/test_game_recorder.py -q -m gpu 2>&1 | tail -15 >> $L
echo "== nccl swap b18" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/gpu_checks/weight_swap_nccl.py b18c384nbt 64 2>&1 | grep -v "^W0\|^\*\*\*" | tail -3 >> $L
cat $L | cut -c1-3000
