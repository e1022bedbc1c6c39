"""Device-resident self-play throughput probe: python tests/gpu_checks/perf_selfplay.py [config] [games] [visits] [steps]"""
import os, sys, time, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from katago_b200 import NeuralNet, SelfPlay, modelgen
cfg = sys.argv[1] if len(sys.argv) > 1 else "b18c384nbt"
games = int(sys.argv[2]) if len(sys.argv) > 2 else 256
visits = int(sys.argv[3]) if len(sys.argv) > 3 else 600
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 600
cap = int(sys.argv[5]) if len(sys.argv) > 5 else 0
p = modelgen.write_model(os.path.join(tempfile.mkdtemp(), cfg + ".bin"), cfg, seed=0)
lm = NeuralNet.loadModelFile(p)
ctx = NeuralNet.createComputeContext([0], 19, 19, True, lm)
h = NeuralNet.createComputeHandle(ctx, lm, games, False, True, 0)
sp = SelfPlay(h, games, visits, komi=7.5, seed=1, ladder_nodes_per_wave=cap)
stream = torch.cuda.ExternalStream(h.stream)
sp.run(20); h.sync()
v0 = sp.stats()['total_visits']
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream); sp.run(steps); e1.record(stream); h.sync()
ms = e0.elapsed_time(e1)
st = sp.stats()
print(f"{cfg} games={games} maxVisits={visits} ladderNodesPerWave={cap}: {steps} waves in {ms:.1f} ms -> {(st['total_visits']-v0)/ms*1e3:.0f} visits/s, {ms/steps:.3f} ms/wave; stats {st}; avg leaf depth {st['sum_leaf_depth']/max(1,st['total_visits']):.2f}")
colors, info = sp.game(0)
print(info); print("\n".join("".join(".XO"[c] for c in row) for row in colors))
print("tree kernels (ms select, ms backup):", sp.time_tree_kernels(50))
