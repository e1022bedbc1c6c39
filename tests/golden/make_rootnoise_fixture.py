"""tests/golden/rootnoise.npz: the reference's Search::addDirichletNoise (searchhelpers.cpp:78-147) with Rand(seed string) on
pseudo-random policies with illegal moves (oracle/_ref/kgref_driver rootnoise)."""
import os, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
cases = [("searchfake$searchThread$0", 362, 1, 10.83, 0.25), ("abc", 82, 2, 10.83, 0.25), ("kgb200", 362, 3, 3.0, 0.5), ("x", 26, 5, 10.83, 0.25),
         ("another seed", 170, 7, 30.0, 0.1), ("z9", 362, 11, 0.5, 0.25), ("q", 362, 13, 10.83, 1.0), ("lots of small alphas", 362, 17, 0.05, 0.25)]
store = {"num": len(cases)}
for i, (seed, n, pseed, conc, w) in enumerate(cases):
    out = subprocess.run([DRIVER, "rootnoise", seed, str(n), str(pseed), repr(conc), repr(w)], capture_output=True, text=True, check=True).stdout
    lines = {ln.split()[0]: np.array([float(t) for t in ln.split()[1:]], np.float32) for ln in out.splitlines()}
    store[f"c{i}_seed"] = np.array(seed); store[f"c{i}_params"] = np.array([conc, w]); store[f"c{i}_in"] = lines["in"]; store[f"c{i}_out"] = lines["out"]
    print(i, seed, n, float(lines["in"][lines["in"] >= 0].sum()), float(lines["out"][lines["out"] >= 0].sum()))
np.savez_compressed(os.path.join(HERE, "rootnoise.npz"), **store)

# Search::chooseIndexWithTemperature draws (oracle/_ref/kgref_driver chooseidx)
cases = [("s", 10, 3, 0.75, 1.0, 200), ("nonSearchRand", 80, 5, 0.15, 1.0, 200), ("t1", 30, 7, 1.0, 1.0, 200), ("below", 40, 9, 0.5, 0.1, 200),
         ("argmax", 25, 11, 0.0, 1.0, 5)]
store = {"num": len(cases)}
for i, (seed, n, pseed, temp, below, count) in enumerate(cases):
    out = subprocess.run([DRIVER, "chooseidx", seed, str(n), str(pseed), repr(temp), repr(below), str(count)], capture_output=True, text=True, check=True).stdout
    lines = {ln.split()[0]: ln.split()[1:] for ln in out.splitlines()}
    store[f"c{i}_seed"] = np.array(seed); store[f"c{i}_params"] = np.array([temp, below])
    store[f"c{i}_weights"] = np.array([float(t) for t in lines["weights"]], np.float64); store[f"c{i}_draws"] = np.array([int(t) for t in lines["draws"]], np.int32)
np.savez_compressed(os.path.join(HERE, "chooseidx.npz"), **store)
