"""Generates tests/golden/searchfake_mixed.npz and featstream_*_in_19.npz: BASELINE config 4, games of different board sizes (and ko
rules) side by side in one device loop whose evaluator frame is 19 x 19.

Reference side: oracle/_ref/kgref_driver with KGREF_NN_LEN=19 - the reference's NNEvaluator is created with nnXLen = nnYLen = 19 and
requireExactNNLen = false, so a 9x9 / 13x7 / 5x5 board is evaluated inside the 19x19 frame exactly like `katago selfplay` does with
bSizes = 9,13,19 (nneval.cpp:874-883): move positions and policies are indexed in the frame (NNPos::locToPos), plane 0 marks the board.
One Search per game (searchfake), one fillRowV7 stream per size (featstream)."""
import json, os, struct, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_search_fixtures as F

FRAME = 19
# the loop's search block: selfplay8mainb18.cfg as far as the loop implements it, both root options of the stock configs, root policy
# temperature (its early-game interpolation scales with each game's own board area), no symmetry sampling (one search Rand per case)
PARAMS = dict(F.SELFPLAY8B18, useGraphSearch=1, rootEndingBonusPoints=0.5, rootPruneUselessMoves=1, rootPolicyTemperature=1.1,
              rootPolicyTemperatureEarly=1.5, chosenMoveTemperatureHalflife=19, **F.BIAS, **F.LCB)
VISITS = 400
GAMES = [   # X, Y, ko rule, move prefix
    (19, 19, 0, F.prefix_from_stream("boardstream_19x19_multisuicide.npz", 40)),
    (9, 9, 1, F.prefix_from_stream("boardstream_9x9_multisuicide.npz", 31)),
    (13, 7, 0, F.prefix_from_stream("boardstream_13x7_nosuicide.npz", 20)),
    (5, 5, 2, F.prefix_from_stream("boardstream_5x5_multisuicide.npz", 20)),
    (7, 7, 0, F.PRUNE_7X7),
    (9, 9, 0, F.prefix_from_stream("boardstream_9x9_multisuicide.npz", 12)),
]


def feat(X, Y, multi, seed, nmoves, every):
    tmp = os.path.join(tempfile.mkdtemp(), "f.bin")
    subprocess.check_call([F.DRIVER, "featstream", str(X), str(Y), str(int(multi)), "7.5", f"random:{seed}:{nmoves}", str(every), tmp],
                          stderr=subprocess.DEVNULL, env=dict(os.environ, KGREF_NN_LEN=str(FRAME)))
    toks = open(tmp + ".moves").read().split()
    raw = open(tmp, "rb").read()
    rec = 4 + (22 * FRAME * FRAME + 19) * 4
    n = len(raw) // rec
    steps = np.zeros(n, np.int32); rows = np.zeros((n, FRAME * FRAME, 22), np.float32); glob = np.zeros((n, 19), np.float32)
    for i in range(n):
        b = raw[i * rec:(i + 1) * rec]
        steps[i] = struct.unpack_from("<i", b, 0)[0]
        rows[i] = np.frombuffer(b, np.float32, 22 * FRAME * FRAME, 4).reshape(FRAME * FRAME, 22)
        glob[i] = np.frombuffer(b, np.float32, 19, 4 + 22 * FRAME * FRAME * 4)
    moves = np.array([(-1, -1) if t == "pass" else tuple(int(v) for v in t.split(",")) for t in toks], np.int8).reshape(-1, 2)
    name = f"featstream_{X}x{Y}_in_{FRAME}.npz"
    np.savez_compressed(os.path.join(HERE, name), X=X, Y=Y, frame=FRAME, multi=int(multi), komi=7.5, moves=moves, steps=steps, rows=rows.astype(np.float16), glob=glob)
    on = rows[:, :, 0].reshape(n, FRAME, FRAME)
    assert on[:, :Y, :X].all() and on.sum() == n * X * Y, "plane 0 must mark exactly the board"
    print(name, n, "rows; ladder planes set:", int(rows[:, :, 14:18].sum()), "area planes set:", int(rows[:, :, 18:20].sum()), os.path.getsize(os.path.join(HERE, name)) // 1024, "KB")


if __name__ == "__main__":
    store = {"num_games": len(GAMES), "frame": FRAME, "visits": VISITS, "params": np.array(json.dumps(PARAMS))}
    for i, (X, Y, ko, moves) in enumerate(GAMES):
        over = dict(PARAMS, koRule=ko) if ko else dict(PARAMS)
        root, v, u, pol, center, psv, threadseed, cstats, rstats = F.run(X, Y, VISITS, moves, over, frame=FRAME)
        assert root[0] == VISITS and v.sum() == VISITS - 1 and len(pol) == FRAME * FRAME + 1
        on_board = np.zeros((FRAME, FRAME), bool); on_board[:Y, :X] = True
        assert (pol[:-1].reshape(FRAME, FRAME)[~on_board] < 0).all() and v[:-1].reshape(FRAME, FRAME)[~on_board].sum() == 0
        store[f"g{i}_setup"] = np.array([X, Y, ko, 1], np.int32)
        store[f"g{i}_moves"] = np.array([(-1, -1) if m is None else m for m in moves], np.int8).reshape(-1, 2)
        store[f"g{i}_visits"] = v; store[f"g{i}_util"] = u; store[f"g{i}_policy"] = pol; store[f"g{i}_play_selection"] = psv
        store[f"g{i}_child_stats"] = cstats; store[f"g{i}_root_stats"] = rstats
        print(i, f"{X}x{Y} ko rule {ko}", len(moves), "moves; children", int((v > 0).sum()), "max visits", int(v.max()), "root util", root[1])
    np.savez_compressed(os.path.join(HERE, "searchfake_mixed.npz"), **store)
    feat(9, 9, True, 31, 160, 16)
    feat(13, 7, False, 32, 120, 15)
    feat(13, 13, True, 33, 260, 26)
