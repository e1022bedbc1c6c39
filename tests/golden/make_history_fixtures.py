"""tests/golden/histstream_*.npz: random games through the REFERENCE BoardHistory (oracle/_ref/kgref_driver histstream): per move
the game-over / no-result / pass-would-end-phase flags, BoardHistory::isLegal of every point for the next player and
superKoBanned - for simple, positional and situational ko on small boards where repetitions are frequent."""
import os, subprocess, struct
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
cases = [("simple_3x3", 3, 3, 0, 1, 11, 60, 60), ("simple_4x4", 4, 4, 0, 0, 12, 60, 80), ("positional_3x3", 3, 3, 1, 1, 13, 60, 60),
         ("positional_5x4", 5, 4, 1, 0, 14, 40, 120), ("situational_3x3", 3, 3, 2, 1, 15, 60, 60), ("situational_5x5", 5, 5, 2, 1, 16, 30, 150),
         ("positional_9x9", 9, 9, 1, 1, 17, 6, 300), ("simple_2x3", 3, 2, 0, 1, 18, 80, 50),
         ("spight_3x3", 3, 3, 3, 1, 19, 60, 80), ("spight_4x4", 4, 4, 3, 0, 20, 40, 120)]
for name, X, Y, ko, multi, seed, games, maxmoves in cases:
    tmp = os.path.join("/tmp", f"hist_{name}.bin")
    subprocess.run([DRIVER, "histstream", str(X), str(Y), str(ko), str(multi), str(seed), str(games), str(maxmoves), tmp], check=True)
    raw = open(tmp, "rb").read()
    hdr = struct.unpack_from("<6i", raw, 0); off = 24
    moves = np.full((games, maxmoves, 2), -2, np.int8); flags = np.zeros((games, maxmoves), np.uint8)
    legal = np.zeros((games, maxmoves, X * Y), np.uint8); banned = np.zeros((games, maxmoves, X * Y), np.uint8); lens = np.zeros(games, np.int32)
    rec = 3 + 2 * X * Y
    for g in range(games):
        n = struct.unpack_from("<i", raw, off)[0]; off += 4
        a = np.frombuffer(raw, np.int8, n * rec, off).reshape(n, rec); off += n * rec
        lens[g] = n; moves[g, :n] = a[:, :2]; flags[g, :n] = a[:, 2].astype(np.uint8)
        legal[g, :n] = a[:, 3:3 + X * Y]; banned[g, :n] = a[:, 3 + X * Y:]
    np.savez_compressed(os.path.join(HERE, f"histstream_{name}.npz"), X=X, Y=Y, ko_rule=ko, multi=multi, lens=lens, moves=moves, flags=flags,
                        legal=legal, banned=banned)
    fin = sum(int(flags[g, lens[g] - 1] & 1) for g in range(games)); nores = sum(int(flags[g, lens[g] - 1] & 2) > 0 for g in range(games))
    print(name, "games", games, "finished", fin, "noResult", nores, "moves", int(lens.sum()), "banned points", int(banned.sum()),
          "spight-like endings", sum(1 for g in range(games) if flags[g, lens[g] - 1] & 1 and not (flags[g, lens[g] - 1] & 2) and not (lens[g] >= 2 and moves[g, lens[g] - 1, 0] == -1 and moves[g, lens[g] - 2, 0] == -1)))
