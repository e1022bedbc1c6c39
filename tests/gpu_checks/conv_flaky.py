"""Repeat single-conv parity runs to catch intermittent failures: python tests/gpu_checks/conv_flaky.py [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kg_nn_oracle as orc
from katago_b200 import NeuralNet
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
cases = [(1, 192, 384, 4, 19, 19), (1, 384, 192, 9, 19, 19), (3, 192, 192, 5, 19, 19), (3, 40, 320, 2, 19, 19), (1, 64, 64, 2, 19, 19)]
for (ky, cin, cout, n, X, Y) in cases:
    rng = np.random.default_rng(ky * 1000 + cin)
    w = (rng.standard_normal((ky, ky, cin, cout)) * np.sqrt(1.0 / (ky * ky * cin))).astype(np.float16).astype(np.float32)
    x = rng.standard_normal((n, Y, X, cin)).astype(np.float16).astype(np.float32)
    ref = orc.conv2d(x, orc.Conv("t", ky, ky, cin, cout, w))
    bad = 0
    for r in range(reps):
        got = NeuralNet.testEvaluateConv(ky, ky, cin, cout, w, n, X, Y, True, x)
        err = np.abs(got - ref)
        if err.max() >= 1e-4:
            bad += 1
            idx = np.argwhere(err >= 1e-4)
            print(f"  case {(ky, cin, cout, n, X, Y)} rep {r}: max err {err.max():.4g}, {len(idx)} bad elements; images {sorted(set(idx[:,0]))} "
                  f"rows y {sorted(set(idx[:,1]))[:8]} x {sorted(set(idx[:,2]))[:8]} channels {idx[:,3].min()}..{idx[:,3].max()} first {idx[0]} got {got[tuple(idx[0])]:.5f} ref {ref[tuple(idx[0])]:.5f}")
    print(f"case {(ky, cin, cout, n, X, Y)}: {bad} / {reps} runs wrong")
