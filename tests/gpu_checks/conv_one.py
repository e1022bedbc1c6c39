"""One convolution shape, a few launches (for ncu): python tests/gpu_checks/conv_one.py K CIN COUT [batch]"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from katago_b200 import load_library
lib = load_library()
k, cin, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 256
ms = np.zeros(1, np.float32)
rc = lib.kgb_bench_conv(k, k, cin, cout, n, 19, 19, 1, 2, 6, ms.ctypes.data_as(C.POINTER(C.c_float)))
assert rc == 0, lib.kgb_last_error()
print(f"{k}x{k} {cin}->{cout} batch {n}: {float(ms[0]) * 1e3:.1f} us")
