"""tests/golden/npyheaders.npz: the 256-byte .npy headers the reference's NumpyBuffer writes for the seven training arrays."""
import os, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
store = {}
for rows in (0, 1, 7, 1024):
    out = subprocess.run([DRIVER, "npyheader", str(rows)], capture_output=True, text=True, check=True).stdout
    for ln in out.splitlines():
        name, hx = ln.split()
        store[f"r{rows}_{name}"] = np.frombuffer(bytes.fromhex(hx), np.uint8)
np.savez_compressed(os.path.join(HERE, "npyheaders.npz"), **store)
print(len(store), "headers")
