"""Training-data container of the reference's self-play output (container + schema layer).

The reference writes one .npz per batch of rows (dataio/trainingwrite.cpp:854-886): seven arrays, each a NumPy v1.0 .npy whose
header is exactly 256 bytes in the reference's own compact spelling (dataio/numpywrite.cpp:97-226), stored in the zip under
the bare array name.  This module writes that container byte-compatibly (headers equal the reference's, tests/test_npz_writer.py)
and packs the binary input planes like `packBits` (trainingwrite.cpp:314-334: 8 points per byte, first point in the high bit,
46 bytes per 19x19 plane).

What fills the arrays: `rows_from_root_observations` turns what the device loop exposes for a root position - the fillRowV7 row
and the root's play selection values - into rows with the input planes, the global inputs and policy target 0
(`Play::extractPolicyTarget`, program/play.cpp:810-846: the selection values scaled so that the largest is at least 10, capped at
30000, rounded to int16; `fillPolicyTarget`, trainingwrite.cpp:346-357).  The value / ownership / score / Q targets of `TrainingWriteBuffers::addRow` (:448-852) need the
finished game and are NOT built yet: their weights in globalTargetsNC (C27, C28, C29, C33, C34; C24 and C35 = 1) say so, which
is the reference's own way of marking a row's missing targets.
"""
import zipfile

import numpy as np

NUM_BIN, NUM_GLOBAL = 22, 19                       # NNInputs::NUM_FEATURES_SPATIAL_V7 / GLOBAL_V7
POLICY_TARGET_CHANNELS, GLOBAL_TARGET_CHANNELS, VALUE_SPATIAL_CHANNELS, QVALUE_CHANNELS = 2, 80, 5, 3    # trainingwrite.cpp:276-279
SCORE_DISTR_RADIUS = 60                            # NNPos::EXTRA_SCORE_DISTR_RADIUS
HEADER_BYTES = 256

# array name -> (descr, trailing shape for a data length L = dataXLen = dataYLen)
def schema(L=19):
    packed = (L * L + 7) // 8
    ps = L * L + 1
    return {
        "binaryInputNCHWPacked": ("|u1", (NUM_BIN, packed)),
        "globalInputNC": ("<f4", (NUM_GLOBAL,)),
        "policyTargetsNCMove": ("<i2", (POLICY_TARGET_CHANNELS, ps)),
        "globalTargetsNC": ("<f4", (GLOBAL_TARGET_CHANNELS,)),
        "scoreDistrN": ("|i1", (2 * (L * L + SCORE_DISTR_RADIUS),)),
        "valueTargetsNCHW": ("|i1", (VALUE_SPATIAL_CHANNELS, L, L)),
        "qValueTargetsNCMove": ("<i2", (QVALUE_CHANNELS, ps)),
    }


def npy_header(descr: str, shape) -> bytes:
    """NumpyBuffer's header: magic, version 1.0, length 246, the dict without spaces, space padding, newline at byte 255."""
    d = "{'descr':'%s','fortran_order':False,'shape':(%s)}" % (descr, ",".join(str(int(x)) for x in shape))
    body = d.encode("ascii")
    if 10 + len(body) >= HEADER_BYTES:
        raise ValueError("numpy header too long")
    return b"\x93NUMPY\x01\x00" + bytes([(HEADER_BYTES - 10) & 0xFF, (HEADER_BYTES - 10) >> 8]) + body + b" " * (HEADER_BYTES - 11 - len(body)) + b"\n"


def pack_bits(planes: np.ndarray) -> np.ndarray:
    """[N, C, L*L] 0/1 -> [N, C, ceil(L*L/8)] uint8, first point in the most significant bit (packBits)."""
    return np.packbits(np.asarray(planes) != 0, axis=2, bitorder="big")


def write_npz(path: str, arrays: dict, L: int = 19, compress: bool = True):
    """Write the seven arrays (all with the same number of rows) the way TrainingWriteBuffers::writeToZipFile does."""
    sch = schema(L)
    n = None
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED if compress else zipfile.ZIP_STORED) as z:
        for name, (descr, rest) in sch.items():
            a = np.ascontiguousarray(arrays[name], dtype=np.dtype(descr))
            if a.shape[1:] != tuple(rest):
                raise ValueError(f"{name}: shape {a.shape} does not end in {rest}")
            n = a.shape[0] if n is None else n
            if a.shape[0] != n:
                raise ValueError(f"{name}: {a.shape[0]} rows, expected {n}")
            z.writestr(name, npy_header(descr, a.shape) + a.tobytes())
    return n


def policy_target_from_play_selection(values):
    """Play::extractPolicyTarget (program/play.cpp:810-846) on Search::getPlaySelectionValues by move position (-1 = no child):
    scaleMaxToAtLeast = 10, cap at 30000, round to nearest (C `round`: halves away from zero), int16."""
    v = np.where(np.asarray(values, np.float64) > 0, np.asarray(values, np.float64), 0.0)
    mx = v.max(axis=-1, keepdims=True)
    v = np.where((mx > 0) & (mx < 10.0), v * (10.0 / np.maximum(mx, 1e-300)), v)
    mx = v.max(axis=-1, keepdims=True)
    v = np.where(mx > 30000.0, v * (30000.0 / np.maximum(mx, 1e-300)), v)
    return np.floor(v + 0.5).astype(np.int16)


def rows_from_root_observations(spatial_nhwc, global_in, play_selection_values, L: int = 19, target_weight: float = 1.0, turn_idx=None,
                                num_visits=None):
    """spatial_nhwc [N, L*L, 22] and global_in [N, 19] as kgb_selfplay_get_nn_row gives them for a root, play_selection_values
    [N, L*L+1] from kgb_selfplay_get_play_selection_values (-1 = no child).  Everything that needs the finished game carries zero weight."""
    sp = np.asarray(spatial_nhwc, np.float32)
    n = sp.shape[0]
    sch = schema(L)
    out = {k: np.zeros((n,) + tuple(rest), np.dtype(descr)) for k, (descr, rest) in sch.items()}
    out["binaryInputNCHWPacked"] = pack_bits(np.transpose(sp, (0, 2, 1)))
    out["globalInputNC"] = np.asarray(global_in, np.float32)
    out["policyTargetsNCMove"][:, 0, :] = policy_target_from_play_selection(play_selection_values)
    out["policyTargetsNCMove"][:, 1, :] = 1                     # uniformPolicyTarget: no next-move target (weight C28 = 0)
    gt = out["globalTargetsNC"]
    gt[:, 24] = 1.0; gt[:, 35] = 1.0          # 1 - weight of the td value targets / of the value targets
    gt[:, 25] = target_weight                 # weight of the row
    gt[:, 26] = 1.0                           # policy target present
    gt[:, 36:41] = 1.0                        # history masks: use all five previous moves
    gt[:, 48] = 1.0                           # area scoring
    gt[:, 63] = 3.0                           # data format version
    if turn_idx is not None:
        gt[:, 51] = np.asarray(turn_idx, np.float32)
    if num_visits is not None:
        gt[:, 60] = np.asarray(num_visits, np.float32)
    return out
