#!/usr/bin/env python3
"""tools/make_pacing_golden.py — golden vectors of the PTS -> field pacing, generated HERE by
the UNMODIFIED reference (push_video + video_isr in oracle/_ref/libefref_vid.so, driven under the
instant-decoder model by oracle/ref_video_harness.cpp:efref_paced): tests/golden/pacing_pins.json holds, per
case, the flip list and the SHA-256 of the emitted field stream. The frames come from the decode oracle (pinned
to the reference decoder), the PTS from the synthetic wrapper's rule PTS_k = 129003 + 3003 k; the last picture is
pushed with mode 1, which is what MpegDecoder::flush_picture(1) does at the end of a stream (player.cpp:694)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from espflix_b200 import synth  # noqa: E402
from tests.oracle_lib import Oracle, RefVideo  # noqa: E402

CASES = [  # name, synth seed offset, pictures, ntsc, frame_counter0, max_fields, mode of the last picture (1 = flush_picture(1); 2 / 3 = load_poster)
    ("ntsc_fc1", 31, 12, 1, 1, 64, 1),
    ("pal_fc5", 32, 12, 0, 5, 64, 1),
    ("ntsc_fc0_quirk", 33, 6, 1, 0, 8, 1),
    ("ntsc_poster3", 34, 4, 1, 7, 64, 3),
    ("pal_poster2", 35, 4, 0, 3, 64, 2),
]


def case_stream(seed, n):
    es, off = synth.generate(synth.SEED0 + seed, n_pictures=n)
    return synth.wrap_ts(es, off)


def main():
    o, rv = Oracle(), RefVideo()
    pins = {}
    for name, seed, n, ntsc, fc0, maxf, last_mode in CASES:
        frames = o.decode_ts(case_stream(seed, n))
        assert frames.shape[0] == n
        pts = 129003 + 3003 * np.arange(n, dtype=np.int64)
        modes = [0] * (n - 1) + [last_mode]
        tail = 17 if last_mode > 1 else 0                 # the poster scroll (_easd) runs for 16 fields after the flip
        fields, ff, fl, stream, hs = rv.paced(frames, pts, ntsc, fc0, maxf, modes=modes, tail_fields=tail, want_hscroll=True)
        pins[name] = {"seed": seed, "pictures": n, "ntsc": ntsc, "frame_counter0": fc0, "max_fields": maxf, "fields": fields, "modes": modes,
                      "tail_fields": tail, "hscroll": [int(x) for x in hs],
                      "flip_field": [int(x) for x in ff], "flip_line": [int(x) for x in fl],
                      "stream_bytes": int(stream.nbytes), "stream_sha256": hashlib.sha256(stream.tobytes()).hexdigest()}
        print(name, fields, pins[name]["flip_field"], pins[name]["flip_line"])
    json.dump(pins, open(os.path.join(ROOT, "tests", "golden", "pacing_pins.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
