"""Parity of the CUDA composite path (K2) against the pins taken from the reference's video_isr and
against the oracle restatement. Sample-exact (uint16, including the deterministic low-byte junk)."""
import hashlib
import json
import os

import numpy as np
import pytest

import espflix_b200

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def ctx():
    c = espflix_b200.Context(n_streams=4, max_pictures=2, es_capacity=4096, fields=True)
    yield c
    c.close()


def test_golden_fields(ctx):
    """config 3: whole fields vs SHA-256 pins produced by the unmodified reference."""
    pins = json.load(open(os.path.join(G, "composite_pins.json")))
    for ntsc in (1, 0):
        ctx.video_init(ntsc)
        assert list(ctx.geometry()) == pins["geometry"]["ntsc" if ntsc else "pal"][:2]
        for f in [p for p in pins["fields"] if p["ntsc"] == ntsc]:
            fr = np.fromfile(os.path.join(G, "frame_%s_%d.i420" % (f["src"], f["frame"])), dtype=np.uint8)
            ctx.write_frame_i420(1, 0, fr)
            ctx.composite_field(fb=0, frame_counter=f["frame_counter"])
            out = ctx.read_field(1)
            assert out.nbytes == f["bytes"]
            assert hashlib.sha256(out.tobytes()).hexdigest() == f["sha256"], f
            assert (out >> 8).max() == f["max_hi"]


def test_random_frames_vs_oracle(ctx, oracle):
    rng = np.random.default_rng(3)
    for ntsc in (1, 0):
        ctx.video_init(ntsc)
        frames = [rng.integers(0, 249, 101376, dtype=np.uint8) for _ in range(3)]
        frames.append(rng.integers(0, 256, 101376, dtype=np.uint8))     # > 248: dither carries cross byte lanes
        for s, fr in enumerate(frames):
            ctx.write_frame_i420(s, 1, fr)
        for fc in (0, 1, 5):
            ctx.composite_field(fb=1, frame_counter=fc)
            for s, fr in enumerate(frames):
                got, want = ctx.read_field(s), oracle.field(fr, ntsc, fc)
                d = np.nonzero(got != want)[0]
                assert d.size == 0, "std %d fc %d stream %d: %d samples differ, first %d" % (ntsc, fc, s, d.size, d[0])


def test_video_isr_line_fetch_and_blit(ctx, oracle):
    rng = np.random.default_rng(4)
    fr = rng.integers(0, 249, 101376, dtype=np.uint8)
    for ntsc in (1, 0):
        ctx.video_init(ntsc)
        ctx.write_frame_i420(2, 0, fr)
        ctx.composite_field(fb=0, frame_counter=1)
        want = oracle.field(fr, ntsc, 1)
        w, n = ctx.geometry()
        for line in (0, 31, 32, 100, 223, 224, n - 1):
            assert np.array_equal(ctx.video_isr(2, line), want[line * w:(line + 1) * w]), line
        for line, x, width in ((0, 0, 352), (191, 0, 352), (77, 16, 64), (100, 8, 344), (50, 4, 20), (9, 330, 3)):
            assert np.array_equal(ctx.blit(2, 0, line, x, width, 1), oracle.blit(fr, ntsc, line, x, width, 1)), (ntsc, line, x, width)
        import espflix_b200
        with pytest.raises(espflix_b200.EspflixError):      # the last 8-pixel group would run past the line (width is rounded up to 8)
            ctx.blit(2, 0, 10, 348, 4, 1)


def test_decode_then_composite_latest_frame(oracle):
    """configs 2+3 chained: composite of the most recently decoded picture (fb = -1)."""
    from espflix_b200 import synth
    es, _ = synth.generate(synth.SEED0 + 5, n_pictures=3, gop=3)
    c = espflix_b200.Context(n_streams=1, max_pictures=3, es_capacity=1 << 20, fields=True)
    frames = c.decode_sequence([es])[0]
    c.composite_field(fb=-1, frame_counter=2)
    assert np.array_equal(c.read_field(0), oracle.field(frames[2], 1, 2))
    c.close()


def test_scroll_and_overlay(oracle):
    """video_isr's two-frame scroll (_hscroll) and the composite() overlay, sample-exact."""
    rng = np.random.default_rng(12)
    a, b = rng.integers(0, 249, 101376, dtype=np.uint8), rng.integers(0, 249, 101376, dtype=np.uint8)
    bm = rng.integers(0, 256, 1280, dtype=np.uint8)
    c = espflix_b200.Context(n_streams=2, max_pictures=2, es_capacity=4096, fields=True)
    c.write_frame_i420(1, 0, a)
    c.write_frame_i420(1, 1, b)
    cases = [(0, 0, 0), (8, 0, 0), (-8, 0, 0), (176, 0, 0), (-344, 0, 0), (344, -1, 100), (0, 32, 0), (0, 5, 239), (0, 31, 17), (24, 1, 300)]
    for ntsc in (1, 0):
        c.video_init(ntsc)
        for hs, blend, prog in cases:
            c.set_scroll(hs)
            c.set_overlay(bm, blend, prog)
            c.composite_field(fb=0, frame_counter=1)
            got, want = c.read_field(1), oracle.field_ex(a, b, ntsc, 1, hs, bm, blend, prog)
            d = np.nonzero(got != want)[0]
            assert d.size == 0, "std %d hscroll %d blend %d progress %d: %d samples differ, first at %d" % (ntsc, hs, blend, prog, d.size, d[0])
    with pytest.raises(espflix_b200.EspflixError):
        c.set_scroll(12)
    c.close()
