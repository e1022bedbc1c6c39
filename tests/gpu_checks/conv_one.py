"""One convolution shape, a few launches (for ncu): python tests/gpu_checks/conv_one.py K CIN COUT [batch] [epilogue kind] [rotate]"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from katago_b200 import load_library
lib = load_library()
k, cin, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 256
kind = int(sys.argv[5]) if len(sys.argv) > 5 else 1
rot = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ms = np.zeros(1, np.float32)
rc = lib.kgb_bench_conv_ex(k, k, cin, cout, n, 19, 19, 1, kind, rot, 2, 6, ms.ctypes.data_as(C.POINTER(C.c_float)))
assert rc == 0, lib.kgb_last_error()
fl = 2.0 * k * k * cin * cout * 361 * n
print(f"{k}x{k} {cin}->{cout} batch {n} kind {kind} rotate {rot}: {float(ms[0]) * 1e3:.1f} us  {fl / ms[0] / 1e9:.1f} TFLOP/s")
