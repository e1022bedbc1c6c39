"""CPU tests of the training-data container (container + schema layer)."""
import os, subprocess, sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katago_b200 import npz_writer as W

GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_npy_headers_equal_the_reference_numpybuffer():
    """The 256-byte headers equal, byte for byte, what the reference's NumpyBuffer writes for the seven arrays (fixture:
    oracle/_ref/kgref_driver npyheader, tests/golden/make_npz_fixture.py) for 0, 1, 7 and 1024 rows."""
    d = np.load(os.path.join(GOLDEN, "npyheaders.npz"))
    sch = W.schema(19)
    for rows in (0, 1, 7, 1024):
        for name, (descr, rest) in sch.items():
            ref = bytes(d[f"r{rows}_{name}"])
            assert W.npy_header(descr, (rows,) + tuple(rest)) == ref, (rows, name)


def test_written_file_has_the_reference_schema_and_round_trips(tmp_path):
    rng = np.random.default_rng(0)
    n = 96
    sp = (rng.random((n, 361, 22)) < 0.2).astype(np.float32)
    gl = rng.standard_normal((n, 19)).astype(np.float32)
    visits = rng.integers(0, 50, (n, 362)).astype(np.float64)
    visits[visits < 25] = -1.0                      # no child there
    visits[:, 5] = 49.0
    rows = W.rows_from_root_observations(sp, gl, visits, turn_idx=np.arange(n), num_visits=np.maximum(visits, 0).sum(1))
    path = str(tmp_path / "rows.npz")
    assert W.write_npz(path, rows) == n
    with np.load(path) as z:
        ref = np.load(os.path.join(os.environ.get("KATAGO_REF", "/root/reference"), "python/testdata/benchmark_data_1024.npz")) if os.path.exists(
            "/root/reference/python/testdata/benchmark_data_1024.npz") else None
        assert set(z.files) == set(W.schema())
        for k, (descr, rest) in W.schema().items():
            assert z[k].dtype == np.dtype(descr) and z[k].shape == (n,) + tuple(rest)
            if ref is not None and k != "globalTargetsNC":          # the sample file is format version 2 (64 global targets)
                assert ref[k].dtype == z[k].dtype and ref[k].shape[1:] == z[k].shape[1:], k
        # the reader's unpacking (python/katago/train/data_processing_pytorch.py:89-95) gives the planes back
        unpacked = np.unpackbits(z["binaryInputNCHWPacked"], axis=2)[:, :, :361]
        assert np.array_equal(unpacked, np.transpose(sp, (0, 2, 1)).astype(np.uint8))
        assert np.array_equal(z["policyTargetsNCMove"][:, 0], np.maximum(visits, 0).astype(np.int16))
        assert (z["policyTargetsNCMove"][:, 1] == 1).all()
        assert np.array_equal(z["globalInputNC"], gl)


@pytest.mark.skipif(not os.path.exists("/root/reference/python/katago/train/data_processing_pytorch.py"), reason="reference python not present")
def test_reference_training_reader_consumes_the_file(tmp_path):
    """Drop-in check: the reference's own training-data reader (python/katago/train/data_processing_pytorch.py) batches the file."""
    sys.path.insert(0, "/root/reference/python")
    import torch
    from katago.train import data_processing_pytorch as dp, modelconfigs
    rng = np.random.default_rng(1)
    n = 64
    sp = (rng.random((n, 361, 22)) < 0.2).astype(np.float32); sp[:, :, 0] = 1.0
    gl = np.zeros((n, 19), np.float32)
    visits = rng.integers(0, 30, (n, 362)).astype(np.float64)
    path = str(tmp_path / "rows.npz")
    W.write_npz(path, W.rows_from_root_observations(sp, gl, visits))
    cfg = modelconfigs.config_of_name["b2c16"]
    batches = list(dp.read_npz_training_data([path], batch_size=32, world_size=1, rank=0, pos_len=19, device=torch.device("cpu"),
                                             randomize_symmetries=False, include_meta=False, model_config=cfg))
    assert len(batches) == 2
    b = batches[0]
    assert tuple(b["binaryInputNCHW"].shape) == (32, 22, 19, 19) and tuple(b["policyTargetsNCMove"].shape)[0] == 32
    assert float(b["binaryInputNCHW"][:, 0].min()) == 1.0


def test_policy_target_follows_extract_policy_target():
    """Play::extractPolicyTarget: largest value scaled up to 10 when smaller, capped at 30000 when larger, rounded to int16."""
    t = W.policy_target_from_play_selection(np.array([[-1.0, 2.0, 0.5, -1.0], [-1.0, 60000.0, 15000.0, 1.0], [12.0, 3.4, -1.0, 7.5]]))
    assert t.tolist() == [[0, 10, 3, 0], [0, 30000, 7500, 1], [12, 3, 0, 8]]
