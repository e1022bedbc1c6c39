"""Generates tests/golden/featstream_*.npz: NNInputs::fillRowV7 rows from the REFERENCE along fixture games
(oracle/_ref/kgref_driver featstream ...), rule subset of the device loop (area scoring, simple ko, no tax, komi 7.5)."""
import os, struct, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")


def make(X, Y, multi, seed, nmoves, every, komi=7.5, ko_rule=0, tag=""):
    tmp = os.path.join(tempfile.mkdtemp(), "f.bin")
    subprocess.check_call([DRIVER, "featstream", str(X), str(Y), str(int(multi)), str(komi), f"random:{seed}:{nmoves}", str(every), tmp] + ([str(ko_rule)] if ko_rule else []),
                          stderr=subprocess.DEVNULL)
    toks = open(tmp + ".moves").read().split()
    raw = open(tmp, "rb").read()
    rec = 4 + (22 * X * Y + 19) * 4
    n = len(raw) // rec
    steps = np.zeros(n, np.int32); rows = np.zeros((n, Y * X, 22), np.float32); glob = np.zeros((n, 19), np.float32)
    for i in range(n):
        b = raw[i * rec:(i + 1) * rec]
        steps[i] = struct.unpack_from("<i", b, 0)[0]
        rows[i] = np.frombuffer(b, np.float32, 22 * X * Y, 4).reshape(Y * X, 22)
        glob[i] = np.frombuffer(b, np.float32, 19, 4 + 22 * X * Y * 4)
    moves = np.array([(-1, -1) if t == "pass" else tuple(int(v) for v in t.split(",")) for t in toks], np.int8).reshape(-1, 2)
    name = f"featstream_{X}x{Y}{tag}.npz"
    np.savez_compressed(os.path.join(HERE, name), X=X, Y=Y, multi=int(multi), komi=komi, ko_rule=ko_rule, moves=moves, steps=steps,
                        rows=rows.astype(np.float16), glob=glob)
    print(name, n, "rows; plane 6 set:", int(rows[:, :, 6].sum()), "ladder planes set:", int(rows[:, :, 14:18].sum()), "area planes set:", int(rows[:, :, 18:20].sum()),
          os.path.getsize(os.path.join(HERE, name)) // 1024, "KB")


if __name__ == "__main__":
    make(19, 19, True, 11, 420, 12)
    make(9, 9, True, 12, 200, 7)
    make(13, 7, False, 13, 150, 5)
    # superko rules on small boards, where bans actually occur (plane 6, globals 6/7)
    make(4, 4, True, 21, 150, 2, ko_rule=1, tag="_positional")
    make(5, 5, True, 22, 220, 3, ko_rule=2, tag="_situational")
    make(9, 9, False, 23, 260, 6, ko_rule=1, tag="_positional")
