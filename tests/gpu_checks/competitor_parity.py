"""Compares the dumps of tests/gpu_checks/competitor_nnloop.sh (reference CUDA backend vs libkgb200, same inputs) with each other
and - given the model file - with the numpy oracle on the first rows.  Runs anywhere (CPU):
    python tests/gpu_checks/competitor_parity.py PREFIX [MODEL [rows]]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def read(path):
    raw = open(path, "rb").read()
    rows, C, G = np.frombuffer(raw[:12], np.int32)
    off = 12
    out = []
    for _ in range(rows):
        sym = int(np.frombuffer(raw[off:off + 4], np.int32)[0]); off += 4
        def take(n):
            nonlocal off
            a = np.frombuffer(raw[off:off + 4 * n], np.float32).copy(); off += 4 * n
            return a
        out.append(dict(sym=sym, spatial=take(C * 361), glob=take(G), policy=take(362), value=take(3), score=take(6), own=take(361)))
    return out


def main():
    prefix = sys.argv[1]
    dumps = {}
    for b in ("cuda", "b200"):
        for fp in (1, 0):
            p = f"{prefix}_{b}_fp16{fp}.bin"
            if os.path.exists(p):
                dumps[(b, fp)] = read(p)
    keys = ("policy", "value", "score", "own")
    base = dumps.get(("cuda", 0)) or next(iter(dumps.values()))
    for k, d in dumps.items():
        same_in = all(np.array_equal(a["spatial"], b["spatial"]) and np.array_equal(a["glob"], b["glob"]) and a["sym"] == b["sym"] for a, b in zip(d, base))
        print(k, "inputs identical to the first dump:", same_in, " max |diff| vs reference cuda fp32:",
              {q: float(max(np.abs(a[q] - b[q]).max() for a, b in zip(d, base))) for q in keys})
    if len(sys.argv) > 2:
        import kg_nn_oracle as orc
        model = orc.load_model(sys.argv[2])
        rows = int(sys.argv[3]) if len(sys.argv) > 3 else 2
        C = len(base[0]["spatial"]) // 361
        sp = np.stack([r["spatial"].reshape(19, 19, C) for r in base[:rows]])   # the dump is NHWC, like the oracle's input
        gl = np.stack([r["glob"] for r in base[:rows]])
        sym = np.array([r["sym"] for r in base[:rows]], np.int32)
        ref = orc.get_output(model, sp, gl, sym, np.zeros(rows, np.float32))
        for k, d in dumps.items():
            print("oracle vs", k, {"policy": float(np.abs(ref["policy"] - np.stack([r["policy"] for r in d[:rows]])).max()),
                                   "value": float(np.abs(ref["value"] - np.stack([r["value"] for r in d[:rows]])).max()),
                                   "ownership": float(np.abs(ref["ownership"].reshape(rows, -1) - np.stack([r["own"] for r in d[:rows]])).max())})


if __name__ == "__main__":
    main()
