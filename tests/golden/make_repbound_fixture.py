"""tests/golden/repbound.npz: Board::simpleRepetitionBoundGt(move, 11) after every move of random streams (reference Board)."""
import os, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
store = {}
for i, (X, Y, n, seed, bound) in enumerate([(5, 5, 400, 1, 11), (9, 9, 500, 2, 11), (19, 19, 700, 3, 11), (4, 4, 300, 4, 5), (7, 3, 300, 5, 8)]):
    out = subprocess.run([DRIVER, "repbound", str(X), str(Y), str(n), str(seed), str(bound)], capture_output=True, text=True, check=True).stdout
    a = np.array([[int(t) for t in ln.split()] for ln in out.splitlines()], np.int32)
    store[f"c{i}_shape"] = np.array([X, Y, bound], np.int32); store[f"c{i}_moves"] = a[:, :3].astype(np.int8); store[f"c{i}_flag"] = a[:, 3].astype(np.uint8)
    print(X, Y, bound, "true fraction", float(a[:, 3].mean()))
store["num"] = 5
np.savez_compressed(os.path.join(HERE, "repbound.npz"), **store)
