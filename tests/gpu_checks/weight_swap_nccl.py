"""N-GPU check of the weight hot-swap (run on the GPU box under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_checks/weight_swap_nccl.py [model=b18c384nbt] [games=64]

Every rank starts from a DIFFERENT net of the architecture (its "previous" net), with a self-play loop running on it.  Rank 0
then stages a new net; the packed arena travels by the library's own ncclBroadcast (kgb_handle_broadcast_staged_weights) into the
other ranks' device memory; every rank commits between two waves.  Checks: afterwards all ranks evaluate a fixed batch to the
bit-identical result, equal to a handle built from the new file on rank 0, and the games went on across the swap.
Prints one JSON line with the broadcast's device time (max over ranks; first call = connection set-up, then steady state)."""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from katago_b200 import NeuralNet, modelgen                      # noqa: E402
from katago_b200.dist_weights import WeightBroadcaster           # noqa: E402
from katago_b200.nn_backend import SelfPlay                      # noqa: E402


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "b18c384nbt"
    games = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    tmp = tempfile.mkdtemp(prefix=f"kgb_swap{rank}_")
    old = modelgen.write_model(os.path.join(tmp, "old.bin"), model, seed=100 + rank)
    lm_old = NeuralNet.loadModelFile(old)
    ctx = NeuralNet.createComputeContext([local], 19, 19, True, lm_old)
    h = NeuralNet.createComputeHandle(ctx, lm_old, games, False, True, local)
    sp = SelfPlay(h, games, 64, komi=7.5, seed=rank, nn_cache_size_power_of_two=16)
    sp.run(8)
    spn, gln = modelgen.synthetic_inputs(4, 19, 19, seed=77)
    before = NeuralNet.getOutput(h, spn.reshape(4, -1), gln)["policy"].copy()
    lm_new = None
    if rank == 0:
        lm_new = NeuralNet.loadModelFile(modelgen.write_model(os.path.join(tmp, "new.bin"), model, seed=7))
    wb = WeightBroadcaster(h, 0, device)
    times = []
    for i in range(6):
        dist.barrier(); torch.cuda.synchronize()
        ms = wb.update(lm_new, selfplay_loops=(sp,))
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
        sp.run(4)                                   # the games go on between swaps
    visits = sp.stats()["total_visits"]
    after = NeuralNet.getOutput(h, spn.reshape(4, -1), gln)
    digest = np.frombuffer(np.ascontiguousarray(after["policy"]).tobytes() + np.ascontiguousarray(after["value"]).tobytes(), np.uint8)
    mine = torch.from_numpy(np.frombuffer(__import__("hashlib").sha256(digest.tobytes()).digest(), np.uint8).copy()).to(device)
    all_ = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(all_, mine)
    same = all(bool((x == all_[0]).all()) for x in all_)
    fresh_ok = None
    if rank == 0:
        ctx2 = NeuralNet.createComputeContext([local], 19, 19, True, lm_new)
        h2 = NeuralNet.createComputeHandle(ctx2, lm_new, 4, False, True, local)
        ref = NeuralNet.getOutput(h2, spn.reshape(4, -1), gln)
        fresh_ok = bool(np.array_equal(ref["policy"], after["policy"]) and np.array_equal(ref["value"], after["value"]))
        print(json.dumps({"check": "weight_swap_nccl", "model": model, "n_gpus": world, "arena_bytes": h.weights_bytes,
                          "broadcast_ms_first": times[0], "broadcast_ms_steady": sorted(times[1:])[len(times[1:]) // 2],
                          "broadcast_ms_all": times, "steady_GBps_per_receiver": h.weights_bytes / (sorted(times[1:])[len(times[1:]) // 2] * 1e-3) / 1e9,
                          "all_ranks_bit_identical": same, "equals_fresh_handle_of_new_net": fresh_ok,
                          "changed_from_previous_net": bool(not np.array_equal(before, after["policy"])), "visits_after_swaps": int(visits)}))
    ok = same and (fresh_ok is not False) and visits > 0
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
