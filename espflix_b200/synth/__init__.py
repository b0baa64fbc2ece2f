"""espflix_b200.synth — ctypes binding of the synthetic MPEG-1 I+P stream generator (efsynth.cpp).
Workload generator for tests and bench.py; not on the decode path."""
import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MBQUANT, FULLPEL, FCODE3, BIGLEVELS, MATRICES, INTRA_IN_P, STATIC, OVERDRIVE = 1, 2, 4, 8, 16, 32, 64, 128
SEED0 = 0x45535046          # "ESPF" 
BENCH_QSCALE, BENCH_NOISE = 6, 10   # tuned so that the mean picture is ~7.4 KB (1.5 Mbit/s at 24 fps ~ 7.8 KB)

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(os.path.join(_HERE, "libefsynth.so"))
        lib.efs_generate.restype = ctypes.c_size_t
        lib.efs_generate.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        lib.efs_wrap_ts.restype = ctypes.c_size_t
        lib.efs_wrap_ts.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        _lib = lib
    return _lib


def generate(seed, n_pictures=12, gop=12, slices=12, qscale=BENCH_QSCALE, flags=0, noise=BENCH_NOISE):
    """-> (es bytes as uint8 array, picture offsets uint32[n_pictures+1])"""
    lib = _load()
    cap = n_pictures * 131072 + 4096
    es = np.empty(cap, dtype=np.uint8)
    off = np.zeros(n_pictures + 1, dtype=np.uint32)
    n = lib.efs_generate(seed, n_pictures, gop, slices, qscale, flags, noise, es.ctypes.data, cap, off.ctypes.data)
    if n == 0:
        raise RuntimeError("efs_generate: capacity exceeded")
    return es[:n].copy(), off


def wrap_ts(es, off):
    lib = _load()
    n_pictures = len(off) - 1
    cap = int(es.size * 1.2) + 188 * 8 * n_pictures + 4096
    ts = np.empty(cap, dtype=np.uint8)
    es = np.ascontiguousarray(es)
    off = np.ascontiguousarray(off, dtype=np.uint32)
    n = lib.efs_wrap_ts(es.ctypes.data, off.ctypes.data, n_pictures, ts.ctypes.data, cap)
    if n == 0:
        raise RuntimeError("efs_wrap_ts: capacity exceeded")
    return ts[:n].copy()


def generate_many(count, first_index=0, workers=None, **kw):
    """count streams with seeds SEED0 + index; returns list of (es, off)."""
    workers = workers or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(workers) as ex:
        return list(ex.map(lambda i: generate(SEED0 + first_index + i, **kw), range(count)))
