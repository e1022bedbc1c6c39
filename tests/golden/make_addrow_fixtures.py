"""tests/golden/addrow_*.json.gz: arguments and resulting buffers of the reference's own TrainingWriteBuffers::addRow
(dataio/trainingwrite.cpp:448-852) on synthetic finished games, dumped by `kgref_driver addrow` (oracle/ref_driver.cpp)."""
import gzip, os, subprocess, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
#        name              X   Y  dataLen turns seed noResult bonus  [alwaysComputePassAliveUnderSuicideRules]
CASES = [("9x9",           9,  9,  9,     24,   5,   0,       0),
         ("19x19",         19, 19, 19,    60,   11,  0,       0),
         ("13x7_in_19",    13, 7,  19,    41,   23,  0,       0),
         ("5x5_in_19",     5,  5,  19,    45,   7,   0,       1.0),
         ("9x9_noresult",  9,  9,  9,     21,   31,  1,       0),
         ("7x7_in_9",      7,  7,  9,     36,   44,  0,       -2.0),
         ("9x9_hugescore", 9,  9,  9,     16,   19,  0,       0),
         ("9x9_passalive", 9,  9,  9,     20,   57,  0,       0,     1)]
ONLY = os.environ.get("ADDROW_ONLY")     # regenerate a single case (the others are byte-stable)
for case in CASES:
    name, X, Y, D, turns, seed, nores, bonus = case[:8]
    if ONLY and name != ONLY:
        continue
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "a.json")
        subprocess.run([DRIVER, "addrow", str(X), str(Y), str(D), str(turns), str(seed), str(nores), str(bonus), path] + [str(c) for c in case[8:]], check=True)
        raw = open(path, "rb").read()
    with gzip.GzipFile(os.path.join(HERE, f"addrow_{name}.json.gz"), "wb", mtime=0) as f:
        f.write(raw)
    print(name, len(raw))
