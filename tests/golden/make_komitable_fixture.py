"""Writes tests/golden/komitable.json.gz: the reference's PlayUtils::computeLead on host cores, with the function its komi bisection saw.

For every case (a position of a committed board stream or a hand-written opening, a komi, a visit count) the UNMODIFIED reference - its Search and
NNEvaluator on the trained g170-b6c96 net of its own test suite, evaluated by the restated CPU backend oracle/cpubackend.cpp (fp32) - gives
  * `lead`, `leads`: PlayUtils::computeLead (program/playutils.cpp:612-660) of the position at the case's komi and at ~26 more starting komis
    [kgref_driver_cpu computelead]
  * `table`: komi -> (lead, winLoss) of PlayUtils::getWhiteScoreValues (the search evalKomi runs) for EVERY half-integer komi roundAndClipKomi can
    return on that board                                                                   [kgref_driver_cpu komitable]
The searches are deterministic and independent (no noise, symmetry 0, cleared tree), so the table is the function computeLead's own searches
sampled.  tests/test_komi_search.py drives katago_b200/komi_search.py's generators with the table and must land on `lead` to the last bit:
that pins the control flow of getNaiveEvenKomiHelper / computeLead (which komis are asked, the bracketing, the interpolation) on CPU.

Run here (needs /root/reference built into oracle/_ref by __graft_entry__.build()):  python tests/golden/make_komitable_fixture.py"""
import gzip
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_search_fixtures import prefix_from_stream

DRIVER = os.path.join(ROOT, "oracle", "_ref", "kgref_driver_cpu")
MODEL = os.path.join(HERE, "models", "g170-b6c96-s175395328-d26788732.bin.gz")

OPENING_9 = [(2, 2), (6, 6), (2, 6), (6, 2), (4, 4)]
OPENING_19 = [(3, 3), (15, 15), (15, 3), (3, 15), (2, 5), (16, 13), (9, 3)]
# (name, x, y, moves, komi, visits)
CASES = [
    ("empty9", 9, 9, [], 7.5, 10), ("empty9_komi0", 9, 9, [], 0.0, 6), ("empty9_far", 9, 9, [], -30.0, 8), ("empty9_int", 9, 9, [], 6.0, 12),
    ("opening9", 9, 9, OPENING_9, 7.5, 10), ("opening9_far", 9, 9, OPENING_9, 45.5, 6), ("opening9_neg", 9, 9, OPENING_9, -12.0, 20),
    ("random9_12", 9, 9, prefix_from_stream("boardstream_9x9_multisuicide.npz", 12), 7.5, 6),
    ("random9_31", 9, 9, prefix_from_stream("boardstream_9x9_multisuicide.npz", 31), 0.5, 10),
    ("random9_50", 9, 9, prefix_from_stream("boardstream_9x9_multisuicide.npz", 50), 7.0, 8),
    ("random9_70", 9, 9, prefix_from_stream("boardstream_9x9_multisuicide.npz", 70), -3.5, 6),
    ("random5_9", 5, 5, prefix_from_stream("boardstream_5x5_multisuicide.npz", 9), 7.5, 16),
    ("random5_20", 5, 5, prefix_from_stream("boardstream_5x5_multisuicide.npz", 20), 24.0, 8),
    ("rect13x7_15", 13, 7, prefix_from_stream("boardstream_13x7_nosuicide.npz", 15), 5.5, 6),
    ("empty19", 19, 19, [], 7.5, 6), ("opening19", 19, 19, OPENING_19, 6.5, 8),
    ("random19_40", 19, 19, prefix_from_stream("boardstream_19x19_multisuicide.npz", 40), 7.5, 6),
    ("random19_150", 19, 19, prefix_from_stream("boardstream_19x19_multisuicide.npz", 150), 0.0, 6),
]


# more starting komis for the same position (the table does not depend on them): 25 spread over the whole range the board allows, and some
# that send getNaiveEvenKomiHelper through its rarer exits (a first shift that made things worse, a capped second shift)
EXTRA_KOMIS = {"empty9": [-40.5, -40.0, -39.5, -34.5, 51.0], "empty9_far": [-39.0], "empty9_int": [-39.5]}


def main():
    out = []
    for name, x, y, moves, komi, visits in CASES:
        s = " ".join("pass" if m is None else f"{m[0]},{m[1]}" for m in moves)
        rng = 20.0 + x * y
        komis = [komi] + [round(2 * (-rng + i * 2 * rng / 24)) / 2 for i in range(25)] + EXTRA_KOMIS.get(name, [])
        leads = {}
        for k in komis:
            r = subprocess.run([DRIVER, "computelead", MODEL, str(x), str(y), str(visits), repr(k), s], capture_output=True, text=True, check=True).stdout
            leads["%.1f" % k] = float([ln.split()[1] for ln in r.splitlines() if ln.startswith("lead ")][0])
        lead = leads["%.1f" % komi]
        r = subprocess.run([DRIVER, "komitable", MODEL, str(x), str(y), str(visits), repr(-rng), repr(rng), s], capture_output=True, text=True, check=True).stdout
        table = {ln.split()[0]: [float(ln.split()[1]), float(ln.split()[2])] for ln in r.splitlines()}
        assert len(table) == int(4 * rng) + 1, (name, len(table))
        out.append(dict(name=name, x=x, y=y, komi=komi, visits=visits, moves=[None if m is None else list(m) for m in moves], lead=float(lead), leads=leads, table=table))
        print(f"{name}: lead {lead} at komi {komi}, table of {len(table)} komis", flush=True)
    blob = json.dumps({"model": os.path.basename(MODEL), "cases": out}, separators=(",", ":")).encode()
    with open(os.path.join(HERE, "komitable.json.gz"), "wb") as f:
        f.write(gzip.compress(blob, 9, mtime=0))
    print("wrote komitable.json.gz", len(blob), "bytes raw")


if __name__ == "__main__":
    main()
