"""tests/golden/searchtargets.npz: per-turn training targets the reference derives from a finished search - value targets
(getNodeValues), Q targets, policy target (Play::extractPolicyTarget), policy surprise and entropies (program/play.cpp:848-948)
- next to the root / child statistics they are derived from, for a few of the searchfake cases (oracle/ref_driver.cpp)."""
import os, subprocess
import numpy as np
import make_search_fixtures as M

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [
    (9, 9, 300, M.prefix_from_stream("boardstream_9x9_multisuicide.npz", 12), {}),
    (9, 9, 600, M.prefix_from_stream("boardstream_9x9_multisuicide.npz", 31), dict(M.SELFPLAY8B18, useGraphSearch=1, **M.BIAS, **M.LCB)),
    (19, 19, 600, M.prefix_from_stream("boardstream_19x19_multisuicide.npz", 40),
     dict(M.SELFPLAY8B18, useGraphSearch=1, rootNumSymmetriesToSample=4, rootPolicyTemperature=1.1, rootPolicyTemperatureEarly=1.5, **M.BIAS, **M.LCB)),
    (13, 7, 500, M.prefix_from_stream("boardstream_13x7_nosuicide.npz", 20),
     {"useLcbForSelection": 1, "lcbStdevs": 3.0, "minVisitPropForLCB": 0.05, "chosenMoveSubtract": 2.0, "chosenMovePrune": 3.0, "valueWeightExponent": 0.5}),
    (5, 5, 1000, M.prefix_from_stream("boardstream_5x5_multisuicide.npz", 20), {"fullHistoryRules": 1, "useGraphSearch": 1}),
    (5, 5, 800, M.prefix_from_stream("boardstream_5x5_multisuicide.npz", 9), dict(M.SELFPLAY8B18, useGraphSearch=1, **M.BIAS, **M.LCB)),
    (5, 5, 90000, M.prefix_from_stream("boardstream_5x5_multisuicide.npz", 9), {"cpuctExploration": 0.3}),   # > 30000 visits on one move: the int16 cap of the policy target
]
store = {"num_cases": len(CASES)}
for i, (X, Y, visits, moves, params) in enumerate(CASES):
    root, v, u, pol, center, psv, threadseed, cstats, rstats = M.run(X, Y, visits, moves, params)
    s = " ".join("pass" if m is None else f"{m[0]},{m[1]}" for m in moves)
    extra = [f"{k}={float(val)!r}" for k, val in params.items() if k != "fullHistoryRules"]
    out = subprocess.run([M.DRIVER, "searchfake", M.MODEL, str(X), str(Y), str(visits), s] + extra, capture_output=True, text=True, check=True).stdout
    P = X * Y + 1
    q = np.zeros((P, 3), np.float64); qmask = np.zeros(P, bool); pt = np.full(P, -1, np.int32)
    for ln in out.splitlines():
        f = ln.split()
        pos = lambda x, y: X * Y if x < 0 else y * X + x
        if f[0] == "valuetargets":
            store[f"c{i}_value_targets"] = np.array([float(t) for t in f[1:5]], np.float32)
        elif f[0] == "surprise":
            store[f"c{i}_surprise"] = np.array([float(t) for t in f[1:4]], np.float64)
        elif f[0] == "qtarget":
            k = pos(int(f[1]), int(f[2])); q[k] = [float(f[3]), float(f[4]), int(f[5])]; qmask[k] = True
        elif f[0] == "policytarget":
            for j in range(1, len(f), 3):
                pt[pos(int(f[j]), int(f[j + 1]))] = int(f[j + 2])
    store[f"c{i}_shape"] = np.array([X, Y, visits], np.int32)
    store[f"c{i}_root_stats"] = rstats; store[f"c{i}_child_stats"] = cstats; store[f"c{i}_edge_visits"] = v
    store[f"c{i}_policy"] = pol; store[f"c{i}_play_selection"] = psv
    store[f"c{i}_q"] = q; store[f"c{i}_q_mask"] = qmask; store[f"c{i}_policy_target"] = pt
    print(i, X, Y, visits, "q entries", int(qmask.sum()), "policy target max", int(pt.max()), store[f"c{i}_surprise"])
np.savez_compressed(os.path.join(HERE, "searchtargets.npz"), **store)
