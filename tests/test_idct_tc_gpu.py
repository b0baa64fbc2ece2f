"""The tcgen05 IDCT experiment (north_star): the linearised 8x8 IDCT as a TF32 GEMM on the 5th-gen tensor
cores (csrc/ef_idct_tc.cu) against the reference's integer transform (oracle efo_idct = player.cpp:922-996), teacher
forced: every block is transformed from its TRUE coefficients, no error is carried from block to block. The bar the
north_star sets for this stage is +-1 LSB; the test also records how often the two differ, which is what decides
whether the tensor path could feed the bit-exact motion-compensation loop (it cannot: see DESIGN.md)."""
import json
import os

import numpy as np
import pytest

import espflix_b200
from espflix_b200 import capi
from tests import idct_linear

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _blocks(n, seed, dense):
    """prescaled coefficient blocks as idct() sees them: v * scale_dct_q[zz], v odd in [-2047, 2047] (dequantised levels)"""
    import re
    hdr = open(os.path.join(ROOT, "espflix_b200", "csrc", "ef_iso11172_tables.h")).read()
    i = hdr.index("ef_aan_prescale[64]")
    pre = np.array([int(x) for x in re.findall(r"\d+", hdr[hdr.index("{", i):hdr.index("};", i)])], dtype=np.int64)
    r = np.random.RandomState(seed)
    v = np.zeros((n, 64), dtype=np.int64)
    k = 64 if dense else 7
    for b in range(n):
        idx = r.choice(64, size=r.randint(2, k + 1), replace=False) if not dense else np.arange(64)
        mag = r.randint(0, 1024, size=idx.size) if dense else (r.geometric(0.02, size=idx.size) % 1024)
        v[b, idx] = (2 * mag + 1) * r.choice([-1, 1], size=idx.size)
    return (v * pre[None, :]).astype(np.int32)


@pytest.mark.parametrize("dense,tiles_per_sm", [(False, 4), (True, 4), (False, 64)], ids=["sparse", "dense", "sparse_1.2M_blocks"])
def test_tensor_core_idct_is_within_one_lsb(oracle, dense, tiles_per_sm):
    n = 128 * 148 * tiles_per_sm                          # 64 tiles per SM = 1.2 M blocks: a throughput-sized run (one K1b launch rebuilds ~5.9 M)
    if tiles_per_sm > 4:
        small = _blocks(128 * 148 * 4, 13, dense)
        coefs = np.ascontiguousarray(np.tile(small, (tiles_per_sm // 4, 1)))
    else:
        coefs = _blocks(n, 11 + dense, dense)
    L = idct_linear.matrix()
    out, prep_ms, mma_ms = capi.idct_tc_run(coefs, L, repeats=5)
    want = coefs.copy()
    for b in range(0, n, 997):                            # the oracle is a Python-called C routine: sample the blocks
        blk = coefs[b].copy()
        oracle.lib.efo_idct(blk.ctypes.data)
        want[b] = blk
    sel = np.arange(0, n, 997)
    diff = out[sel].astype(np.int64) - want[sel].astype(np.int64)
    exact = idct_linear.linear_idct(coefs[sel].astype(np.float64).reshape(-1, 8, 8)).reshape(-1, 64)
    assert np.abs(out[sel] - np.rint(exact)).max() <= 1      # the GEMM itself reproduces the linear map (ties aside)
    assert np.abs(diff).max() <= 1, "more than 1 LSB from the reference transform"
    frac = float((diff != 0).mean())
    rec = {"blocks": n, "dense": bool(dense), "samples_compared": int(diff.size), "fraction_off_by_one": frac,
           "blocks_with_any_difference": float((np.abs(diff).max(axis=1) > 0).mean()), "prep_ms": prep_ms, "mma_ms": mma_ms,
           "blocks_per_s_mma": n / (mma_ms / 1000.0), "blocks_per_s_total": n / ((mma_ms + prep_ms) / 1000.0)}
    print("tcgen05 IDCT experiment:", json.dumps(rec))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "idct_tc_%s%s.json" % ("dense" if dense else "sparse", "_large" if tiles_per_sm > 4 else "")), "w") as f:
        json.dump(rec, f)
