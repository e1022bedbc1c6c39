"""tests/golden/scorevalue_samples.npz: 3000 argument tuples and the reference's ScoreValue::expectedWhiteScoreValue of each
(oracle/_ref/kgref_driver svsamples, i.e. neuralnet/nninputs.cpp:98-192 itself)."""
import os, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
out = subprocess.run([DRIVER, "svsamples", "3000", "7"], capture_output=True, text=True, check=True).stdout
a = np.array([[float(t) for t in ln.split()] for ln in out.splitlines()], np.float64)
np.savez_compressed(os.path.join(HERE, "scorevalue_samples.npz"), args=a[:, :5], value=a[:, 5])
print(a.shape, a[:, 5].min(), a[:, 5].max())

out = subprocess.run([DRIVER, "vwtable"], capture_output=True, text=True, check=True).stdout
t = np.array([float(x) for x in out.split()], np.float64)
assert t.shape == (2000,)
np.save(os.path.join(HERE, "value_weight_cdf_table.npy"), t)
