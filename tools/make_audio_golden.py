#!/usr/bin/env python3
"""tools/make_audio_golden.py — pins of the audio path, generated HERE by the UNMODIFIED reference
(oracle/_ref/efref_audio: player.cpp demux -> video.cpp push_audio/decode_audio -> sbc_decoder.cpp -> the
pdm_second_order of espflix.ino) on the audio that the reference's own splash.ts / vmedia.ts carry, plus the
reference's constant tables (sbc_decoder.cpp:23-97) for tests/test_tables.py. Writes tests/golden/audio_pins.json
and tests/golden/sbc_tables.json. Needs /root/reference (run in the build container, not on the GPU box).

Quirk Q13 (found while pinning): decode_audio() reads a frame linearly out of its 4 KB ring; with a frame size that does
not divide 4096 (vmedia: 48 bytes) the straddling frame is read past the ring (undefined). Those frames and the 72
samples after them (the filter memory) are listed as `undefined` and masked out of the PCM hash; the PDM stream is a
chaotic recurrence of everything before it and is pinned only up to the first undefined sample."""
import hashlib
import json
import os
import re
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def run_ref(ts_path):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "a.bin")
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "efref_audio"), ts_path, out], check=True, capture_output=True)
        raw = open(out, "rb").read()
    nes, npcm = [int(x) for x in np.frombuffer(raw[:16], dtype=np.uint64)]
    es = np.frombuffer(raw[16:16 + nes], dtype=np.uint8)
    pcm = np.frombuffer(raw[16 + nes:16 + nes + 2 * npcm], dtype=np.int16)
    pdm = np.frombuffer(raw[16 + nes + 2 * npcm:], dtype=np.uint16)
    return es, pcm, pdm


def undefined_ranges(frame_size, n_frames):
    """sample ranges the reference's ring read makes undefined: straddling frame + 9 blocks of filter memory"""
    out = []
    for k in range(n_frames):
        if (frame_size * k) % 4096 + frame_size > 4096:
            out.append([k * 128, min(n_frames * 128, k * 128 + 128 + 72)])
    return out


def masked(pcm, ranges):
    p = pcm.copy()
    for a, b in ranges:
        p[a:b] = 0
    return p


def main():
    pins = {}
    for name in ("splash", "vmedia"):
        es, pcm, pdm = run_ref(os.path.join(G, name + ".ts"))
        n_frames = len(pcm) // 128
        frame_size = len(es) // n_frames
        und = undefined_ranges(frame_size, n_frames)
        first = und[0][0] if und else len(pcm)
        pins[name] = {
            "es_bytes": int(len(es)), "es_sha256": hashlib.sha256(es.tobytes()).hexdigest(),
            "frame_size": frame_size, "n_frames": n_frames, "undefined": und,
            "pcm_sha256_masked": hashlib.sha256(masked(pcm, und).tobytes()).hexdigest(),
            "pdm_defined_words": 2 * first, "pdm_sha256_defined": hashlib.sha256(pdm[:2 * first].tobytes()).hexdigest(),
            "pcm_head": [int(x) for x in pcm[:16]], "pcm_abs_max": int(np.abs(pcm.astype(np.int32)).max()),
        }
        print(name, pins[name]["es_bytes"], frame_size, n_frames, len(und), "undefined ranges")
    json.dump(pins, open(os.path.join(G, "audio_pins.json"), "w"), indent=1)

    src = open("/root/reference/src/sbc_decoder.cpp").read()

    def arr(name, signed32=True):
        m = re.search(name + r"\[.*?\] = \{(.*?)\};", src, re.S)
        v = np.array([int(x, 16) for x in re.findall(r"0x([0-9A-Fa-f]+)", m.group(1))], dtype=np.uint32)
        return [int(x) for x in v.astype(np.int32)]

    m = re.search(r"SBC_offset8\[4\]\[8\] = \{(.*?)\};", src, re.S)
    off8 = [int(x) for x in re.findall(r"-?\d+", m.group(1))]
    json.dump({"SBC_syn_8": arr("SBC_syn_8"), "SBC_proto_8": arr("SBC_proto_8"), "SBC_offset8": off8}, open(os.path.join(G, "sbc_tables.json"), "w"))


if __name__ == "__main__":
    main()
