"""tests/golden/gameinit_hist.json: what the reference's GameInitializer::createGame (program/play.cpp:330-650) draws for 200 000 games of a
configuration with mixed board sizes (rectangles allowed), three ko rules, both suicide rules and komi noise - as histograms
(`kgref_driver gameinit`).  katago_b200/game_initializer.py must reproduce the distributions (its random numbers are its own)."""
import collections, json, os, subprocess, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
CFG = {"koRules": "SIMPLE,POSITIONAL,SITUATIONAL", "scoringRules": "AREA", "taxRules": "NONE", "multiStoneSuicideLegals": "false,true", "hasButtons": "false",
       "bSizes": "9,13,19", "bSizeRelProbs": "2,1,5", "allowRectangleProb": "0.2", "komiMean": "7.0", "komiStdev": "1.0", "komiBigStdevProb": "0.1",
       "komiBigStdev": "12.0", "komiAllowIntegerProb": "0.5"}
N = 200000
if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "gi.cfg")
        open(path, "w").write("".join(f"{k} = {v}\n" for k, v in CFG.items()))
        out = subprocess.run([DRIVER, "gameinit", path, str(N), "gameinit-fixture"], capture_output=True, text=True, check=True).stdout
    sizes, ko, suicide, komi = collections.Counter(), collections.Counter(), collections.Counter(), {}
    for ln in out.splitlines():
        x, y, k, s, km = ln.split()
        sizes[f"{x}x{y}"] += 1; ko[k] += 1; suicide[s] += 1
        komi.setdefault(f"{x}x{y}", collections.Counter())[km] += 1
    json.dump({"cfg": CFG, "n": N, "sizes": sizes, "ko_rules": ko, "multi_stone_suicide": suicide, "komi_by_size": {k: dict(v) for k, v in komi.items() if k in ("9x9", "19x19")}},
              open(os.path.join(HERE, "gameinit_hist.json"), "w"), indent=0, sort_keys=True)
    print(dict(sizes), dict(ko), dict(suicide))
