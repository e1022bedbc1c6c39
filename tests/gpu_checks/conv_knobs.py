"""Timing experiments on single conv layers (kgb_bench_conv) under the bring-up knobs KGB_CONV_DBG / KGB_CONV_STAGES."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from katago_b200 import load_library
    import numpy as np
    lib = load_library()
    for (k, ci, co) in ((3, 192, 192), (1, 384, 192), (1, 192, 384), (3, 256, 256)):
        ms = np.zeros(1, np.float32)
        rc = lib.kgb_bench_conv(k, k, ci, co, 256, 19, 19, 1, 5, 30, ms.ctypes.data_as(C.POINTER(C.c_float)))
        fl = 2.0 * k * k * ci * co * 361 * 256
        print(f"  {k}x{k} {ci}->{co}: {ms[0]*1e3:7.1f} us  {fl/ms[0]/1e9:7.1f} TFLOP/s" if rc == 0 else "  error " + lib.kgb_last_error().decode(), flush=True)
else:
    for env in [dict(kv.split("=") for kv in a.split(",")) if a != "-" else {} for a in (sys.argv[1:] or ["-"])]:
        print("knobs", env, flush=True)
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, "child"], env=e, timeout=120)
