"""tests/golden/searchlimits.json.gz: whole games of the reference's Play::runGame (fake net, `kgref_driver rungame`) with the per-move search limits
switched on - reduceVisits (KGREF_REDUCE = threshold,lookback,minVisits,weight) and recorded cheap searches (KGREF_CHEAP = prob,visits,weight):
per turn the root's visits (= the budget getSearchLimitsThisMove gave the search), the target weight it gave the turn, and the root win/loss
value that runGame appends to historicalMctsWinLossValues.  katago_b200/game_recorder.py search_limits_this_move must map the values so far
to the next turn's budget and weight exactly."""
import gzip, json, os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
MODEL = os.path.join(HERE, "models", "torchref_b2c16.bin.gz")
#        size visits maxMoves seed  reduce                cheap
CASES = [(9,  60,    80,      11,  "0.0,1,10,0.1",      None),
         (7,  48,    90,      12,  "0.02,3,12,0.5",     None),
         (9,  60,    80,      13,  "0.05,1,15,0.1",     "0.4,16,0.3"),
         (13, 40,    70,      14,  None,                 "0.6,10,0.5")]
games = []
for size, visits, max_moves, seed, reduce, cheap in CASES:
    env = dict(os.environ)
    if reduce: env["KGREF_REDUCE"] = reduce
    if cheap: env["KGREF_CHEAP"] = cheap
    out = subprocess.run([DRIVER, "rungame", MODEL, str(size), str(visits), str(max_moves), str(seed), "0", "0", "0"], capture_output=True, text=True, check=True, env=env).stdout
    g = json.loads(out)
    games.append(dict(size=size, maxVisits=visits, reduce=reduce, cheap=cheap, rootWinLoss=g["rootWinLoss"], rootVisits=g["rootVisits"], targetWeight=g["targetWeightUnrounded"]))
    print(size, visits, reduce, cheap, g["turns"], "turns; distinct budgets", sorted(set(g["rootVisits"]))[:12], "weights", sorted(set(round(w, 3) for w in g["targetWeightUnrounded"]))[:8])
with gzip.GzipFile(os.path.join(HERE, "searchlimits.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps(games).encode())
