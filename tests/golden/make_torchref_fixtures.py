"""Generates tests/golden/torchref_<cfg>.npz + tests/golden/models/torchref_<cfg>.bin.gz   (run HERE, needs /root/reference).

An independent fp32 check of the oracle and of the CUDA path (third oracle): build the reference's
PyTorch model (python/katago/train/model_pytorch.py) for a small config, randomise its parameters, export it with the
reference's own exporter (python/export_model_pytorch.py -checkpoint) and record torch's outputs on seeded inputs.
Output mapping follows the exporter (export_model_pytorch.py:557-690): policy channels 0 and 5 -> inference channels
0/1, value -> 3 logits, miscvalues[0:4] + moremiscvalues[0:2] -> 6 score-value outputs, ownership.
"""
import gzip, os, shutil, subprocess, sys, tempfile
import numpy as np
import torch

REF = "/root/reference/python"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from katago.train import modelconfigs  # noqa: E402
from katago.train.model_pytorch import Model  # noqa: E402
from katago_b200 import modelgen  # noqa: E402


def make(cfg_name: str, n: int, seed: int, sizes=None, base=None, override=None):
    torch.manual_seed(seed)
    cfg = dict(modelconfigs.config_of_name[base or cfg_name])
    cfg.update(override or {})
    model = Model(cfg, 19)
    model.initialize()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn_like(p) * (0.05 if p.dim() > 1 else 0.2))
    model.eval()
    tmp = tempfile.mkdtemp()
    ckpt = os.path.join(tmp, "ckpt.ckpt")
    torch.save({"model": model.state_dict(), "config": cfg}, ckpt)
    subprocess.check_call([sys.executable, os.path.join(REF, "export_model_pytorch.py"), "-checkpoint", ckpt, "-export-dir", tmp,
                           "-model-name", "torchref-" + cfg_name, "-filename-prefix", "m"], cwd=REF, stdout=subprocess.DEVNULL)
    os.makedirs(os.path.join(HERE, "models"), exist_ok=True)
    dst = os.path.join(HERE, "models", f"torchref_{cfg_name}.bin.gz")
    with open(os.path.join(tmp, "m.bin"), "rb") as f, gzip.open(dst, "wb", compresslevel=9) as g:
        g.write(f.read())
    sp, gl = modelgen.synthetic_inputs(n, 19, 19, seed=seed + 100, board_sizes=sizes)
    with torch.no_grad():
        outs = model(torch.from_numpy(np.ascontiguousarray(sp.transpose(0, 3, 1, 2))), torch.from_numpy(gl))
    o = outs[0] if isinstance(outs[0], (tuple, list)) else outs
    policy, value, misc, moremisc, ownership = o[0].numpy(), o[1].numpy(), o[2].numpy(), o[3].numpy(), o[4].numpy()
    np.savez_compressed(os.path.join(HERE, f"torchref_{cfg_name}.npz"), spatial_nhwc=sp.astype(np.float16), global_=gl,
                        policy0=policy[:, 0, :], policy_opt=policy[:, 5, :], value=value,
                        score_value=np.concatenate([misc[:, 0:4], moremisc[:, 0:2]], axis=1),
                        ownership=ownership.reshape(n, -1))
    shutil.rmtree(tmp)
    print(cfg_name, "policy range", policy[:, 0].min(), policy[:, 0].max(), "value", value[0])


if __name__ == "__main__":
    make("b2c16", 4, 1, sizes=[(19, 19), (9, 9), (19, 19), (13, 7)])
    # the stock b1c6nbt config has regularC = mid - gpool = 0, which the reference's own C++ loader rejects
    # (desc.cpp:1700-1704); use the same nested-bottleneck family with a gpool inner block instead
    make("b2c32nbt", 4, 2, sizes=[(19, 19), (19, 19), (11, 11), (19, 19)], base="b1c6nbt",
         override=dict(trunk_num_channels=32, mid_num_channels=16, gpool_num_channels=8,
                       block_kind=[["rconv1", "bottlenest2"], ["rconv2", "bottlenest2gpool"]],
                       p1_num_channels=8, g1_num_channels=8, v1_num_channels=12, v2_size=16))
    make("b4c32", 3, 3)
