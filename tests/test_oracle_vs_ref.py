"""The oracle restatement against the UNMODIFIED reference (oracle/_ref) on the synthetic coverage
streams, plus assertions that those streams really exercise the paths the reference's fixtures
lack. Skipped where oracle/_ref is not built (it is built in the agent container and travels to
the GPU box as binaries)."""
import numpy as np
import pytest

from espflix_b200 import synth
from tests import oracle_lib
from tests.synth_cases import COVERAGE, make

pytestmark = pytest.mark.skipif(not oracle_lib.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("idx", range(len(COVERAGE)), ids=[c[0] for c in COVERAGE])
def test_port_equals_reference_on_synthetic(oracle, idx):
    name, kw = COVERAGE[idx]
    es, off = make(idx, kw)
    ts = synth.wrap_ts(es, off)
    info, ref = oracle_lib.ref_decode_ts(ts)
    got = oracle.decode_ts(ts)
    assert info["frames"] == kw["n_pictures"] == got.shape[0]
    assert np.array_equal(got, ref), name
    assert np.array_equal(oracle.decode_es(es), ref), "ES path"
    assert np.array_equal(oracle.demux_ts(ts), es), "TS wrapper round trip"


def test_coverage_set_reaches_the_missing_paths(oracle):
    oracle.stats_reset()
    for idx, (name, kw) in enumerate(COVERAGE):
        oracle.decode_es(make(idx, kw)[0])
    s = oracle.stats()
    t = list(s.mb_type)
    assert t[0x11] > 0 and t[0x12] > 0 and t[0x1A] > 0, "macroblock-level quantiser changes"
    assert t[0x01] > 0 and t[0x02] > 0 and t[0x08] > 0 and t[0x0A] > 0
    assert s.full_pel_mbs > 0
    assert s.f_code[3] > 0 and s.f_code[1] > 0
    assert s.escapes16 > 0
    assert s.q2_zero > 0, "quirk Q2 (zero coefficient becomes +1)"
    assert all(v > 0 for v in s.mocomp_xy), "all four half-pel cases"
    assert s.skipped > 0 and s.blocks_dc_only > 0
    assert s.pictures[1] > 0 and s.pictures[2] > 0
    assert s.pin_out_of_domain == 0, "coverage streams must stay inside the reference's defined clamp domain"


def test_reference_video_equals_port_on_random_frames(oracle):
    rv = oracle_lib.RefVideo()
    rng = np.random.default_rng(7)
    for ntsc in (1, 0):
        for fc in (0, 1, 2):
            fr = rng.integers(0, 249, 101376, dtype=np.uint8)
            assert np.array_equal(oracle.field(fr, ntsc, fc), rv.field(fr, ntsc, fc)), (ntsc, fc)
        fr = rng.integers(0, 256, 101376, dtype=np.uint8)          # bytes above 248: dither carries cross bytes
        assert np.array_equal(oracle.field(fr, ntsc, 1), rv.field(fr, ntsc, 1))
        for line, x, w in ((0, 0, 352), (191, 0, 352), (77, 16, 64), (100, 8, 344)):
            assert np.array_equal(oracle.blit(fr, ntsc, line, x, w, 1), rv.blit(fr, ntsc, line, x, w, 1)), (ntsc, line, x, w)


PRESENTATION_CASES = [(0, 0, 0), (8, 0, 0), (-8, 0, 0), (176, 0, 0), (-344, 0, 0), (344, -1, 100), (0, 32, 0), (0, 5, 239), (0, 31, 17), (24, 1, 300)]


def test_presentation_extras_port_equals_reference(oracle):
    """_hscroll two-frame scroll and the composite() overlay / progress bar / fade."""
    rv = oracle_lib.RefVideo()
    rng = np.random.default_rng(11)
    a, b = rng.integers(0, 249, 101376, dtype=np.uint8), rng.integers(0, 249, 101376, dtype=np.uint8)
    bm = rng.integers(0, 256, 1280, dtype=np.uint8)
    for ntsc in (1, 0):
        for hs, blend, prog in PRESENTATION_CASES:
            want = rv.field_ex(a, b, ntsc, 1, hs, bm, blend, prog)
            assert np.array_equal(oracle.field_ex(a, b, ntsc, 1, hs, bm, blend, prog), want), (ntsc, hs, blend, prog)
