"""Quick NN-evaluator throughput probe (GPU box): python tests/gpu_checks/perf_nn.py [config] [batch] [iters]"""
import os, sys, time, tempfile, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from katago_b200 import NeuralNet, modelgen, load_library
cfg = sys.argv[1] if len(sys.argv) > 1 else "b18c384nbt"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
fp16 = os.environ.get("KGB_FP16", "1") == "1"
tmp = tempfile.mkdtemp()
p = modelgen.write_model(os.path.join(tmp, cfg + ".bin"), cfg, seed=0)
lm = NeuralNet.loadModelFile(p)
ctx = NeuralNet.createComputeContext([0], 19, 19, fp16, lm)
h = NeuralNet.createComputeHandle(ctx, lm, n, False, True)
sp, gl = modelgen.synthetic_inputs(n, 19, 19, seed=5)
dev = torch.device("cuda:0")
dsp = torch.from_numpy(sp.reshape(n, -1)).to(dev); dgl = torch.from_numpy(gl).to(dev)
dpol = torch.empty((n, 362), device=dev); dval = torch.empty((n, 3), device=dev); dsc = torch.empty((n, 6), device=dev); down = torch.empty((n, 361), device=dev)
lib = load_library()
stream = torch.cuda.ExternalStream(h.stream)
def run():
    rc = lib.kgb_forward_device(h._p, n, dsp.data_ptr(), dgl.data_ptr(), None, None, dpol.data_ptr(), dval.data_ptr(), dsc.data_ptr(), down.data_ptr())
    assert rc == 0, lib.kgb_last_error()
torch.cuda.synchronize()
for _ in range(3): run()
h.sync()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record(stream)
    for _ in range(iters): run()
    e1.record(stream)
h.sync()
ms = e0.elapsed_time(e1) / iters
macs = lm.desc["conv_macs_per_position"]
flop = 2.0 * macs * 361 * n
print(f"{cfg} batch {n} fp16={fp16} streams={os.environ.get('KGB_STREAM_FP32','default')}: {ms:.3f} ms/forward, {n/ms*1e3:.0f} evals/s, {flop/ms/1e9:.1f} TFLOP/s (algorithmic), launches/forward={h.launches_per_forward}", flush=True)
print("policy finite:", bool(torch.isfinite(dpol).all()), "value sample", dval[0].tolist())
