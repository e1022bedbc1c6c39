"""Soak: the selfplay command with everything the stock configuration switches on that is built - mixed board sizes and rules, komi noise,
komiAuto, policy-initialised openings, cheap and reduced searches, lead estimation, surprise weighting, evaluation cache, 4 root symmetries -
on the trained g170-b6c96 net, per-game release.    python tests/gpu_checks/soak_cli.py [games] [visits]"""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from katago_b200 import selfplay_cli as C
from test_selfplay_cli import STOCK_B18_SETTINGS
games = int(sys.argv[1]) if len(sys.argv) > 1 else 48
visits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tmp = tempfile.mkdtemp(prefix="kgb_soak_")
models = os.path.join(tmp, "models"); os.makedirs(models)
import shutil
shutil.copy(os.path.join(ROOT, "tests", "golden", "models", "g170-b6c96-s175395328-d26788732.bin.gz"), os.path.join(models, "g170.bin.gz"))
settings = dict(STOCK_B18_SETTINGS, bSizes="9,13,19", bSizeRelProbs="1,1,1", allowRectangleProb="0.1", dataBoardLen="19", maxVisits=str(visits), cheapSearchVisits=str(max(8, visits // 4)),
                reducedVisitsMin=str(max(8, visits // 4)), reduceVisitsThreshold="0.9", reduceVisitsThresholdLookback="3", reducedVisitsWeight="0.1", maxMovesPerGame="160",
                nnCacheSizePowerOfTwo="16", estimateLeadProb="0.1", estimateLeadVisits="6", compensateKomiVisits="20", maxRowsPerTrainFile="5000")
cfg = os.path.join(tmp, "soak.cfg")
open(cfg, "w").write("".join(f"{k} = {v}\n" for k, v in settings.items()))
out = os.path.join(tmp, "out")
t0 = time.time()
rc = C.main(["-models-dir", models, "-output-dir", out, "-config", cfg, "-max-games-total", str(games), "-games-per-gpu", "32", "-per-game-release"])
dt = time.time() - t0
rows, lead, weights, sizes, komis, starts = 0, 0, [], {}, set(), []
for f in os.listdir(os.path.join(out, "g170", "tdata")):
    with np.load(os.path.join(out, "g170", "tdata", f)) as z:
        g, gi = z["globalTargetsNC"], z["globalInputNC"]
        rows += len(g); lead += int((g[:, 29] > 0).sum()); weights += list(g[:, 25])
        on = np.unpackbits(z["binaryInputNCHWPacked"][:, 0], axis=1)[:, :361].sum(1)
        for a in on:
            sizes[int(a)] = sizes.get(int(a), 0) + 1
        komis |= set(np.round(np.abs(gi[:, 5]) * 20, 1).tolist()); starts += list(g[:, 53])
print(json.dumps({"check": "soak_cli", "rc": rc, "seconds": dt, "games": games, "rows": rows, "rows_with_lead": lead, "row_board_areas": sizes, "distinct_komis": len(komis),
                  "mean_opening_moves": float(np.mean(starts)) if starts else None, "mean_row_weight": float(np.mean(weights)) if weights else None}))
