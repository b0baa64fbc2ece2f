"""The oracle (oracle/ef_oracle.c) against the pins taken from the UNMODIFIED reference
(tests/golden/*.json, produced by tools/make_golden.py from oracle/_ref). CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name", ["splash", "vmedia"])
def test_decode_fixture_matches_reference_pins(oracle, name):
    pins = json.load(open(os.path.join(G, "decode_pins.json")))[name]
    ts = open(os.path.join(G, name + ".ts"), "rb").read()
    assert hashlib.sha256(ts).hexdigest() == pins["ts_sha256"]
    frames = oracle.decode_ts(ts)
    assert frames.shape[0] == pins["frames"]
    for k in range(frames.shape[0]):
        assert hashlib.sha256(frames[k].tobytes()).hexdigest() == pins["frame_sha256"][k], "frame %d" % k
    assert hashlib.sha256(frames.tobytes()).hexdigest() == pins["i420_sha256"]


def test_demux_then_es_decode_equals_ts_decode(oracle):
    ts = open(os.path.join(G, "splash.ts"), "rb").read()
    es = oracle.demux_ts(ts)
    assert 100000 < es.size < len(ts)
    assert np.array_equal(oracle.decode_es(es), oracle.decode_ts(ts))


def test_first_frame_is_flat_and_clamped(oracle):
    # splash frame 0 is flat Y=16, U=V=128; global max sample is 248 (quirk Q1)
    frames = oracle.decode_ts(open(os.path.join(G, "splash.ts"), "rb").read())
    f0 = frames[0]
    assert (f0[:352 * 192] == 16).all() and (f0[352 * 192:] == 128).all()
    assert frames.max() == 248


def test_composite_geometry_luts_and_fields(oracle):
    pins = json.load(open(os.path.join(G, "composite_pins.json")))
    for ntsc, name in ((1, "ntsc"), (0, "pal")):
        v = oracle.video(ntsc)
        geo = [v.line_width, v.line_count, v.hsync, v.hsync_long, v.hsync_short, v.burst_start, v.burst_width, v.active_start]
        assert geo == pins["geometry"][name]
        tab = np.frombuffer(v.color_tab, dtype=np.uint32)
        assert hashlib.sha256(tab.tobytes()).hexdigest() == pins["color_tab_sha256"][name]
        assert np.array_equal(tab, np.fromfile(os.path.join(G, "color_tab_%s.u32" % name), dtype=np.uint32))
    v = oracle.video(0)
    assert list(v.burst0)[:44] == pins["pal_burst"]["burst0"] and list(v.burst1)[:44] == pins["pal_burst"]["burst1"]
    for f in pins["fields"]:
        fr = np.fromfile(os.path.join(G, "frame_%s_%d.i420" % (f["src"], f["frame"])), dtype=np.uint8)
        out = oracle.field(fr, f["ntsc"], f["frame_counter"])
        assert out.nbytes == f["bytes"]
        assert hashlib.sha256(out.tobytes()).hexdigest() == f["sha256"], f


def test_field_spot_values(oracle):
    # SURVEY.md 8c spot values, NTSC, vmedia frame 0, frame_counter 0
    fr = np.fromfile(os.path.join(G, "frame_vmedia_0.i420"), dtype=np.uint8)
    f = oracle.field(fr, 1, 0).reshape(262, 912)
    assert (f[0, :64] == 0).all()
    assert list(f[0, 64:72]) == [0x1E00, 0x1400, 0x0A00, 0x1400] * 2
    assert (f[0, 104:] == 0x1800).all()
    assert list(f[32, 158:168]) == [0x1800, 0x1800, 0x1C18, 0x1A18, 0x1C00, 0x1C18, 0x1C1C, 0x1C1C, 0x1C04, 0x1C18]
    assert (f[259, :840] == 0).all() and (f[259, 840:] == 0x1400).all()
    assert (f >> 8).max() == 84


def test_idct_is_the_reference_transform_not_the_ideal_one(oracle):
    # a flat DC block: b[0] = dc << 8 through both passes gives dc everywhere
    b = np.zeros(64, dtype=np.int32)
    b[0] = 100 << 8
    assert (oracle.idct(b) == 100).all()
    # entry (7,7) of the prescale is 2 (exact 2.44): impulse response differs from an exact IDCT
    b = np.zeros(64, dtype=np.int32)
    b[63] = 1000 * 2
    out = oracle.idct(b).reshape(8, 8)
    assert out[0, 0] != 0 and abs(int(out[0, 0])) < 64


def test_strip_layout_roundtrip(oracle):
    rng = np.random.default_rng(1)
    i420 = rng.integers(0, 249, 101376, dtype=np.uint8)
    s = oracle.i420_to_strips(i420)
    assert np.array_equal(oracle.strips_to_i420(s), i420)
    # Y(x,y) = y*528 + x ; block-4 chroma row yc at strip rows 0-7, block-5 at 8-15 (player.cpp:33-46)
    assert s[5 * 528 + 7] == i420[5 * 352 + 7]
    assert s[(16 * 2 + 3) * 528 + 352 + 9] == i420[352 * 192 + (8 * 2 + 3) * 176 + 9]
    assert s[(16 * 2 + 8 + 3) * 528 + 352 + 9] == i420[352 * 192 + 176 * 96 + (8 * 2 + 3) * 176 + 9]
