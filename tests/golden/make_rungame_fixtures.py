"""tests/golden/rungame.json.gz: whole self-play games of the reference's own Play::runGame on the CPU (fake net, `kgref_driver rungame`):
per turn the value targets, the raw net's win / loss / noResult and the policy surprise, and what runGame derived from them - the
value surprise of each turn and the surprise-weighted target weights."""
import gzip, json, os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
MODEL = os.path.join(HERE, "models", "torchref_b2c16.bin.gz")
#        size visits maxMoves seed psw  vsw  useSearchValueSurprise
CASES = [(9,  60,    60,      3,   0.5, 0.1, 0),
         (9,  40,    200,     4,   0.5, 0.1, 1),
         (13, 30,    90,      5,   0.3, 0.2, 0),
         (7,  50,    120,     6,   0.0, 0.4, 0),
         (19, 20,    50,      7,   0.5, 0.0, 0),
         (5,  80,    100,     8,   0.5, 0.1, 0)]
games = []
for c in CASES:
    out = subprocess.run([DRIVER, "rungame", MODEL] + [str(x) for x in c], capture_output=True, text=True, check=True).stdout
    games.append(json.loads(out))
    print(c, games[-1]["turns"], "turns, hit limit", games[-1]["hitTurnLimit"])
with gzip.GzipFile(os.path.join(HERE, "rungame.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps(games).encode())
