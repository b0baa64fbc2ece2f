"""PTS -> field pacing of push_video (video.cpp:1023-1057, 1165-1177), CPU side: the C restatement
of the schedule against the pins the unmodified reference produced (tools/make_pacing_golden.py) and, where
oracle/_ref exists, against the reference itself (real push_video / video_isr on two threads) for irregular PTS."""
import json
import os

import numpy as np
import pytest

from tests.oracle_lib import Oracle, RefVideo, have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PINS = json.load(open(os.path.join(ROOT, "tests", "golden", "pacing_pins.json")))


@pytest.mark.parametrize("name", sorted(PINS))
def test_schedule_matches_reference_pins(name):
    p = PINS[name]
    pts = 129003 + 3003 * np.arange(p["pictures"], dtype=np.int64)
    fields, ff, fl, hs = Oracle().paced_schedule(pts, p["ntsc"], p["frame_counter0"], p["max_fields"], modes=p["modes"], tail_fields=p["tail_fields"], want_hscroll=True)
    assert fields == p["fields"] and ff.tolist() == p["flip_field"] and fl.tolist() == p["flip_line"]
    assert hs.tolist() == p["hscroll"]                    # the poster scroll (_animate / _easd) field by field


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_schedule_matches_reference_on_irregular_pts():
    """Jitter, repeated and decreasing PTS (late frames, 'resetting v timing'), long gaps, both standards."""
    o, rv = Oracle(), RefVideo()
    rng = np.random.default_rng(2024)
    frames = np.zeros((2, 101376), dtype=np.uint8)                      # content is irrelevant to the schedule
    for case in range(24):
        ntsc = case % 2
        n = int(rng.integers(2, 14))
        step = rng.choice([3003, 3003, 3003, 1501, 6006, 0, -4000, 45045], size=n)
        pts = (129003 + np.cumsum(step)).astype(np.int64)
        pts = np.maximum(pts, 0)
        fc0 = int(rng.integers(0, 4)) if case < 8 else int(rng.integers(1, 100000))
        fr = np.ascontiguousarray(np.broadcast_to(frames[0], (n, 101376)))
        modes = np.where(rng.random(n) < 0.2, rng.integers(1, 4, size=n), 0).astype(np.int32)      # 1 at once, 2 / 3 poster scroll
        tail = int(rng.integers(0, 20))
        rf, rff, rfl, _, rhs = rv.paced(fr, pts, ntsc, fc0, 400, want_fields=False, modes=modes, tail_fields=tail, want_hscroll=True)
        f, ff, fl, hs = o.paced_schedule(pts, ntsc, fc0, 400, modes=modes, tail_fields=tail, want_hscroll=True)
        assert (f, ff.tolist(), fl.tolist()) == (rf, rff.tolist(), rfl.tolist()), (case, ntsc, fc0, pts.tolist())
        assert hs.tolist() == rhs.tolist(), (case, modes.tolist())
