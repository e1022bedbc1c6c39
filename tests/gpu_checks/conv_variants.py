"""Single-layer conv timings for the trunk shapes of b18c384nbt / b28c512nbt with each production epilogue (kgb_bench_conv_ex),
under the knob sets given on the command line ("K=V,K=V" per run, "-" = defaults).  Rotating buffers: working set > L2.
    python tests/gpu_checks/conv_variants.py - KGB_CONV_IMPL=tc KGB_T3_E=1 ..."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SHAPES = [(3, 192, 192, 1, "3x3 mid (act)"), (3, 192, 192, 2, "3x3 unit end (+res,raw,act)"), (1, 384, 192, 3, "pre 1x1 (raw,act)"),
          (1, 192, 384, 2, "post 1x1 (+res,raw,act)"), (3, 256, 256, 1, "b28 3x3 mid")]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from katago_b200 import load_library
    import numpy as np
    lib = load_library()
    batch = int(os.environ.get("BENCH_BATCH", "256"))
    for (k, ci, co, kind, name) in SHAPES:
        for rot in (1, 4):
            ms = np.zeros(1, np.float32)
            rc = lib.kgb_bench_conv_ex(k, k, ci, co, batch, 19, 19, 1, kind, rot, 5, 40, ms.ctypes.data_as(C.POINTER(C.c_float)))
            fl = 2.0 * k * k * ci * co * 361 * batch
            print(f"  {name:30s} {k}x{k} {ci:3d}->{co:3d} rotate {rot}: {ms[0]*1e3:7.1f} us  {fl/ms[0]/1e9:7.1f} TFLOP/s" if rc == 0
                  else "  error " + lib.kgb_last_error().decode(), flush=True)
else:
    for env in [dict(kv.split("=") for kv in a.split(",")) if a != "-" else {} for a in (sys.argv[1:] or ["-"])]:
        print("knobs", env, flush=True)
        e = dict(os.environ); e.update(env)
        try:
            subprocess.run([sys.executable, __file__, "child"], env=e, timeout=180)
        except subprocess.TimeoutExpired:
            print("  TIMEOUT", flush=True)
