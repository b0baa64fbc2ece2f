#!/usr/bin/env python3
"""tools/make_index_golden.py — golden vectors of the trick-mode index, generated HERE by the
UNMODIFIED reference tool's own functions (oracle/_ref/libefref_idx.so = indexer/indexer.cpp compiled where it
lies): tests/golden/index_pins.json (sequence-header tables of the reference's embedded streams, hashes of
video.idx images for fixture and fuzz combinations) and tests/golden/video_idx_fixture.bin (the image for
main = vmedia.ts, fast-forward = splash.ts, rewind = vmedia.ts, struct padding zeroed).
The GPU box has no /root/reference: tests there use these files."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.oracle_lib import RefIndexer  # noqa: E402
from tests.ts_cases import make_ts  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def main():
    ref = RefIndexer()
    pins = {"tables": {}, "images": {}}
    for name in ("splash", "vmedia"):
        r = ref.make_index(open(os.path.join(G, name + ".ts"), "rb").read())
        pins["tables"][name] = {"first_pts": r["first_pts"], "last_pts": r["last_pts"],
                                "seq_pts": [int(x) for x in r["seq_pts"]], "seq_pos": [int(x) for x in r["seq_pos"]]}
    fx = [open(os.path.join(G, n + ".ts"), "rb").read() for n in ("vmedia", "splash", "vmedia")]
    img = ref.build_idx(fx)
    open(os.path.join(G, "video_idx_fixture.bin"), "wb").write(img)
    pins["images"]["fixture:vmedia,splash,vmedia"] = {"bytes": len(img), "sha256": hashlib.sha256(img).hexdigest()}
    for seeds, mono in (((100, 101, 102), True), ((7, 8, 9), False), ((20, 21, 22), True)):
        files = [make_ts(s, monotonic=mono) for s in seeds]
        img = ref.build_idx(files)
        pins["images"]["fuzz:%s:%d" % (",".join(map(str, seeds)), int(mono))] = {"bytes": len(img), "sha256": hashlib.sha256(img).hexdigest()}
    json.dump(pins, open(os.path.join(G, "index_pins.json"), "w"), indent=1)
    print(json.dumps(pins["images"], indent=1))


if __name__ == "__main__":
    main()
