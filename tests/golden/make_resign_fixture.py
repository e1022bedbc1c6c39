"""tests/golden/resign.json.gz: games of the reference's Play::runGame (fake net, `kgref_driver rungame`) with allowResignation on (KGREF_RESIGN =
threshold,consecTurns): the root win/loss value after every search and how the game ended (resigned or not, winner).  The resignation rule of
katago_b200/match_play.py (`should_resign`) must fire exactly at the game's last move and nowhere before - or never, for games that ended otherwise."""
import gzip, json, os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
MODEL = os.path.join(HERE, "models", "torchref_b2c16.bin.gz")
#        size visits maxMoves seed  resign
CASES = [(9, 40, 120, 21, "-0.10,2"), (9, 40, 120, 22, "-0.05,3"), (7, 40, 100, 23, "-0.15,2"), (13, 30, 150, 24, "-0.10,3"), (9, 40, 60, 25, "-0.60,3"),
         (5, 40, 60, 26, "-0.02,1"), (9, 30, 120, 27, "0.0,4")]
games = []
for size, visits, max_moves, seed, resign in CASES:
    out = subprocess.run([DRIVER, "rungame", MODEL, str(size), str(visits), str(max_moves), str(seed), "0", "0", "0"], capture_output=True, text=True, check=True,
                         env=dict(os.environ, KGREF_RESIGN=resign)).stdout
    g = json.loads(out)
    games.append(dict(size=size, resign=resign, turns=g["turns"], rootWinLoss=g["rootWinLoss"], resigned=g["resigned"], winner=g["winner"], hitTurnLimit=g["hitTurnLimit"]))
    print(size, resign, g["turns"], "turns, resigned", g["resigned"], "winner", g["winner"])
with gzip.GzipFile(os.path.join(HERE, "resign.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps(games).encode())
