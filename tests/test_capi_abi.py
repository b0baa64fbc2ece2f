"""The C-ABI library loads on a CPU-only box, exports every symbol include/espflix_b200.h declares,
and refuses loudly to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import espflix_b200
from espflix_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "espflix_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(ef_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(espflix_b200.lib_path())
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert set(capi._SIGNATURES) == set(names), set(capi._SIGNATURES) ^ set(names)


def test_version_and_error_strings():
    lib = espflix_b200.load_library()
    assert b"sm_100a" in lib.ef_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(espflix_b200.EspflixError) as e:
        espflix_b200.Context(n_streams=1, max_pictures=2)
    assert e.value.code == capi.EF_ECUDA
    assert "no CPU" in str(e.value) or "CUDA" in str(e.value)


def test_index_entry_points_have_no_cpu_fallback_either():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(espflix_b200.EspflixError) as e:
        espflix_b200.tsidx_scan([b"\x47" + b"\x00" * 187])
    assert e.value.code == capi.EF_ECUDA


def test_product_does_not_touch_the_oracle():
    # nothing under espflix_b200/ may import, link or call anything under oracle/
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "espflix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                s = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"ef_oracle|oracle_lib|efref_|libef_oracle|oracle/", s) and f != "build.py":
                    bad.append(os.path.join(d, f))
    assert not bad, bad
