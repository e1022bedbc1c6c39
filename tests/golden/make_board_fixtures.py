"""Generates tests/golden/boardstream_*.npz from the REFERENCE Board (oracle/_ref/kgref_driver boardstream ...).
Run here (needs oracle/_ref built from /root/reference).  Random legal move streams incl. passes, captures, kos, and
(multi-stone) suicides; after every move: stones, ko point, capture counters, Zobrist pos_hash, per-stone liberties,
legality of every point for the next player (Board::isLegal, game/board.cpp:441-465), and the pass-alive/territory area
(Board::calculateArea with all flags on, game/board.cpp:1853-2228)."""
import os, struct, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")


def make(X, Y, n, seed, multi):
    tmp = os.path.join(tempfile.mkdtemp(), "bs.bin")
    subprocess.check_call([DRIVER, "boardstream", str(X), str(Y), str(n), str(seed), str(int(multi)), tmp])
    raw = open(tmp, "rb").read()
    x, y, nm, ms = struct.unpack_from("<iiii", raw, 0)
    assert (x, y, nm, ms) == (X, Y, n, int(multi))
    off = 16
    rec = 3 + 2 + 4 + 16 + 4 * X * Y
    moves = np.zeros((n, 3), np.int8); ko = np.zeros((n, 2), np.int8); caps = np.zeros((n, 2), np.int16)
    hashes = np.zeros((n, 2), np.uint64)
    colors = np.zeros((n, Y, X), np.uint8); libs = np.zeros((n, Y, X), np.uint8); legal = np.zeros((n, Y, X), np.uint8); area = np.zeros((n, Y, X), np.uint8)
    for i in range(n):
        b = raw[off:off + rec]; off += rec
        moves[i] = np.frombuffer(b, np.int8, 3, 0)
        ko[i] = np.frombuffer(b, np.int8, 2, 3)
        caps[i] = np.frombuffer(b, np.int16, 2, 5)
        hashes[i] = np.frombuffer(b, np.uint64, 2, 9)
        colors[i] = np.frombuffer(b, np.uint8, X * Y, 25).reshape(Y, X)
        libs[i] = np.frombuffer(b, np.uint8, X * Y, 25 + X * Y).reshape(Y, X)
        legal[i] = np.frombuffer(b, np.uint8, X * Y, 25 + 2 * X * Y).reshape(Y, X)
        area[i] = np.frombuffer(b, np.uint8, X * Y, 25 + 3 * X * Y).reshape(Y, X)
    assert off == len(raw)
    name = f"boardstream_{X}x{Y}_{'multisuicide' if multi else 'nosuicide'}.npz"
    np.savez_compressed(os.path.join(HERE, name), X=X, Y=Y, multi=int(multi), moves=moves, ko=ko, caps=caps, pos_hash=hashes,
                        colors=colors, libs=libs, legal_next=legal, area=area)
    print(name, "captures", caps[-1], "kos", int((ko[:, 0] >= 0).sum()), "passes", int((moves[:, 0] < 0).sum()),
          os.path.getsize(os.path.join(HERE, name)) // 1024, "KB")


if __name__ == "__main__":
    make(19, 19, 900, 1, True)
    make(19, 19, 900, 2, False)
    make(9, 9, 500, 3, True)
    make(13, 7, 400, 4, False)
    make(5, 5, 400, 5, True)
