"""The linear map behind MpegDecoder::idct() (player.cpp:922-996): the same butterfly network with its constants
362/256, 473/256, 196/256 but WITHOUT the rounding shifts, columns then rows, final / 256. L[o][i] = response of output
sample o to input coefficient i. Used by the tcgen05 IDCT experiment (tests/test_idct_tc_gpu.py)."""
import numpy as np


def _pass(v):
    """one 8-point pass on the last axis (float, no rounding)"""
    b1 = v[..., 4]
    b3 = v[..., 2] + v[..., 6]
    b4 = v[..., 5] - v[..., 3]
    t1 = v[..., 1] + v[..., 7]
    t2 = v[..., 3] + v[..., 5]
    b6 = v[..., 1] - v[..., 7]
    b7 = t1 + t2
    m0 = v[..., 0]
    x4 = (b6 * 473 - b4 * 196) / 256 - b7
    x0 = x4 - (t1 - t2) * 362 / 256
    x1 = m0 - b1
    x2 = (v[..., 2] - v[..., 6]) * 362 / 256 - b3
    x3 = m0 + b1
    y3, y4, y5, y6 = x1 + x2, x3 + b3, x1 - x2, x3 - b3
    y7 = -x0 - (b4 * 473 + b6 * 196) / 256
    return np.stack([b7 + y4, x4 + y3, y5 - x0, y6 - y7, y6 + y7, x0 + y5, y3 - x4, y4 - b7], axis=-1)


def linear_idct(blocks):
    """blocks float64 [..., 8 rows, 8 cols] -> the unrounded transform (columns first, then rows, then / 256)"""
    cols = np.swapaxes(_pass(np.swapaxes(blocks, -1, -2)), -1, -2)
    return _pass(cols) / 256.0


def matrix():
    eye = np.eye(64).reshape(64, 8, 8)
    return linear_idct(eye).reshape(64, 64).T.copy()      # [out][in]
