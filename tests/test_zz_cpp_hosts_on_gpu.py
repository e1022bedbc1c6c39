"""First runs of the C++-only hosts (katago_b200/b200_selfplay, b200_gatekeeper: integration/*_main.cpp) on a real GPU.

They were written after this round's GPU minutes were spent, so until now they are validated on the CPU only - against the Python hosts on a mock of
the C ABI (tests/test_cpp_host.py).  These tests run last (file name) and are marked xfail(strict=False): a pass shows up as XPASS in the GPU suite's
summary, a failure as xfailed - it cannot turn the suite red or stop it (-x) - and the processes run under a timeout of their own."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SELFPLAY = os.path.join(ROOT, "katago_b200", "b200_selfplay")
GATEKEEPER = os.path.join(ROOT, "katago_b200", "b200_gatekeeper")

def _env():
    """The loader must find libcudart for a stand-alone binary (a Python process gets it from torch's own wheels): the CUDA runtime directories of
    this image in front of whatever LD_LIBRARY_PATH holds."""
    dirs = ["/usr/local/cuda/lib64"]
    try:
        import nvidia.cuda_runtime
        dirs.insert(0, os.path.join(os.path.dirname(nvidia.cuda_runtime.__file__), "lib"))
    except Exception:
        pass
    return dict(os.environ, LD_LIBRARY_PATH=":".join(dirs + [os.environ.get("LD_LIBRARY_PATH", "")]))


CFG = """maxVisits = 24
numGameThreads = 16
bSizes = 7,9
bSizeRelProbs = 1,2
koRules = SIMPLE,POSITIONAL
multiStoneSuicideLegals = false,true
komiMean = 7.0
komiStdev = 1.0
maxMovesPerGame = 60
dataBoardLen = 9
nnCacheSizePowerOfTwo = 16
rootNoiseEnabled = true
rootNumSymmetriesToSample = 2
policySurpriseDataWeight = 0.5
valueSurpriseDataWeight = 0.1
cheapSearchProb = 0.25
cheapSearchVisits = 8
cheapSearchTargetWeight = 0.0
initGamesWithPolicy = true
policyInitAreaProp = 0.04
estimateLeadProb = 0.1
estimateLeadVisits = 6
earlyForkGameProb = 0.2
earlyForkGameExpectedMoveProp = 0.1
forkGameProb = 0.1
forkGameMinChoices = 2
earlyForkGameMaxChoices = 3
forkGameMaxChoices = 3
forkSidePositionProb = 0.05
maxRowsPerTrainFile = 400
b200WavesPerPoll = 8
"""


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first GPU run of the C++-only selfplay host (CPU-validated so far)")
@pytest.mark.skipif(not os.path.exists(SELFPLAY), reason="katago_b200/b200_selfplay not built (__graft_entry__.build())")
def test_cpp_selfplay_host_on_the_device(tmp_path, tmp_models):
    cfg = tmp_path / "c.cfg"
    cfg.write_text(CFG)
    out = tmp_path / "out"
    r = subprocess.run([SELFPLAY, "-model", tmp_models["tiny_reg"], "-config", str(cfg), "-output-dir", str(out), "-max-games-total", "24", "-seed", "3"],
                       capture_output=True, text=True, timeout=75, env=_env())
    print(r.stderr[-3000:])
    assert r.returncode == 0, r.stderr[-2000:]
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    print("C++ selfplay host on the device:", summary)
    assert summary["games_written"] == 24 and summary["rows"] > 100 and summary["visits"] > 1000
    files = sorted(os.listdir(out / "tdata"))
    assert files and all(f.endswith(".npz") and len(f) == 20 for f in files)
    rows = 0
    for f in files:
        z = np.load(out / "tdata" / f)
        n = z["globalTargetsNC"].shape[0]
        rows += n
        assert z["binaryInputNCHWPacked"].shape == (n, 22, 11) and z["policyTargetsNCMove"].shape == (n, 2, 82) and z["valueTargetsNCHW"].shape == (n, 5, 9, 9)
        planes = np.unpackbits(z["binaryInputNCHWPacked"], axis=2)[:, :, :81]
        assert (planes[:, 0].sum(axis=1) >= 49).all()                       # the board mask: 7x7 .. 9x9 points
        assert not (planes[:, 1] & planes[:, 2]).any()                      # own and opponent stones never overlap
        assert (z["policyTargetsNCMove"][:, 0].sum(axis=1) > 0).all() and np.isfinite(z["globalTargetsNC"]).all()
        assert set(np.unique(z["globalTargetsNC"][:, 63])) == {3.0}         # data format version
    assert rows == summary["rows"]
    sgfs = os.listdir(out / "sgfs")
    assert len(sgfs) == 1 and open(out / "sgfs" / sgfs[0]).read().count("\n") == 24


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first GPU run of the C++-only gatekeeper host (CPU-validated so far)")
@pytest.mark.skipif(not os.path.exists(GATEKEEPER), reason="katago_b200/b200_gatekeeper not built (__graft_entry__.build())")
def test_cpp_gatekeeper_host_on_the_device(tmp_path, tmp_models):
    import shutil
    cfg = tmp_path / "g.cfg"
    cfg.write_text("maxVisits = 16\nnumGameThreads = 8\nnumGamesPerGating = 10\nbSizes = 9\nkomiMean = 7.0\nmaxMovesPerGame = 60\nchosenMoveTemperatureEarly = 0.5\n")
    for d in ("test/cand-s2", "accepted/base-s1"):
        os.makedirs(tmp_path / d)
    shutil.copy(tmp_models["tiny_reg"], tmp_path / "accepted" / "base-s1" / "model.bin")
    os.utime(tmp_path / "accepted" / "base-s1", (1000, 1000))
    shutil.copy(tmp_models["tiny_nbt"], tmp_path / "test" / "cand-s2" / "model.bin")
    r = subprocess.run([GATEKEEPER, "-config", str(cfg), "-test-models-dir", str(tmp_path / "test"), "-sgf-output-dir", str(tmp_path / "sgfs"),
                        "-accepted-models-dir", str(tmp_path / "accepted"), "-rejected-models-dir", str(tmp_path / "rejected"), "-quit-if-no-nets-to-test", "-games-per-gpu", "8"],
                       capture_output=True, text=True, timeout=75, env=_env())
    print(r.stderr[-3000:])
    assert r.returncode == 0, r.stderr[-2000:]
    assert ("Candidate won match" in r.stderr) != ("Candidate lost match" in r.stderr)
    won = "Candidate won match" in r.stderr
    assert os.path.isdir(tmp_path / ("accepted" if won else "rejected") / "cand-s2") and not os.listdir(tmp_path / "test")
    records = os.listdir(tmp_path / "sgfs" / "cand-s2")
    assert len(records) == 1 and open(tmp_path / "sgfs" / "cand-s2" / records[0]).read().count("(;FF[4]") >= 5


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first GPU run of the C++-only selfplay host next to the Python host (CPU-validated so far)")
@pytest.mark.skipif(not os.path.exists(SELFPLAY), reason="katago_b200/b200_selfplay not built (__graft_entry__.build())")
def test_cpp_and_python_selfplay_hosts_write_the_same_rows_on_the_device(tmp_path, tmp_models):
    """Both hosts make the same ABI calls in the same order (that is what tests/test_cpp_host.py shows on the mock), the device loop is deterministic,
    so the same seed gives the same games: the C++ host's rows and records are the Python host's, bit for bit (the Python command does not cap the
    games of its last pump, so it may write a few more at the end: compared is the common prefix, which must be everything the C++ host wrote).
    The evaluation cache is off here: which of two colliding entries survives a wave is the one thing a slot's timing could change."""
    import shutil
    os.makedirs(tmp_path / "nets")
    shutil.copy(tmp_models["tiny_reg"], tmp_path / "nets" / "tiny-s1.bin")
    cfg = tmp_path / "c.cfg"
    cfg.write_text(CFG.replace("nnCacheSizePowerOfTwo = 16", "nnCacheSizePowerOfTwo = 0").replace("maxRowsPerTrainFile = 400", "maxRowsPerTrainFile = 100000"))
    runs = {}
    for name, cmd in (("cpp", [SELFPLAY]), ("py", [sys.executable, "-m", "katago_b200.selfplay_cli", "-per-game-release"])):
        r = subprocess.run(cmd + ["-models-dir", str(tmp_path / "nets"), "-config", str(cfg), "-output-dir", str(tmp_path / name), "-max-games-total", "20", "-seed", "7"],
                           capture_output=True, text=True, timeout=75 if name == "cpp" else 150, env=_env(), cwd=ROOT)
        print(name, r.stderr[-1500:])
        assert r.returncode == 0, (name, r.stderr[-2000:])
        d = tmp_path / name / "tiny-s1"
        files = sorted(os.listdir(d / "tdata"), key=lambda f: os.path.getmtime(d / "tdata" / f))
        z = [np.load(d / "tdata" / f) for f in files]
        runs[name] = ({k: np.concatenate([x[k] for x in z]) for k in z[0].files}, open(d / "sgfs" / os.listdir(d / "sgfs")[0]).read().splitlines())
    (cpp_rows, cpp_sgf), (py_rows, py_sgf) = runs["cpp"], runs["py"]
    n = cpp_rows["globalTargetsNC"].shape[0]
    print(f"C++ host: {len(cpp_sgf)} games, {n} rows; Python host: {len(py_sgf)} games, {py_rows['globalTargetsNC'].shape[0]} rows")
    assert len(cpp_sgf) == 20 and py_sgf[:20] == cpp_sgf
    for k in cpp_rows:
        assert py_rows[k].shape[0] >= n and cpp_rows[k].tobytes() == py_rows[k][:n].tobytes(), k
