"""tests/golden/writegame_*.json.gz: synthetic FinishedGameData and the rows the reference's own TrainingDataWriter::writeGame
(dataio/trainingwrite.cpp:1097-1325) produced from it through its text sink, dumped by `kgref_driver writegame`."""
import gzip, os, subprocess, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
#        name              X   Y  dataLen turns seed maxRows firstFileProp noResult
CASES = [("9x9",           9,  9,  9,     30,   7,   16,     0.5,          0),
         ("9x9_b",         9,  9,  9,     34,   8,   1000,   1.0,          0),
         ("7x7_in_9",      7,  7,  9,     40,   12,  25,     0.0,          0),
         ("9x9_noresult",  9,  9,  9,     20,   21,  12,     0.3,          1),
         ("13x13",         13, 13, 13,    36,   34,  20,     0.8,          0),
         # the game's first 7 / 12 moves are its start history (a policy-initialised opening, a forked game): KGREF_START_MOVES
         ("9x9_start7",    9,  9,  9,     28,   41,  1000,   1.0,          0),
         ("7x7_in_9_start12", 7, 7, 9,    16,   43,  18,     0.4,          0)]
START_MOVES = {"9x9_start7": 7, "7x7_in_9_start12": 12}
for name, X, Y, D, turns, seed, max_rows, prop, nores in CASES:
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "a.json")
        subprocess.run([DRIVER, "writegame", str(X), str(Y), str(D), str(turns), str(seed), str(max_rows), str(prop), str(nores), path], check=True,
                       env=dict(os.environ, KGREF_START_MOVES=str(START_MOVES.get(name, 0))))
        raw = open(path, "rb").read()
    with gzip.GzipFile(os.path.join(HERE, f"writegame_{name}.json.gz"), "wb", mtime=0) as f:
        f.write(raw)
    print(name, len(raw))
