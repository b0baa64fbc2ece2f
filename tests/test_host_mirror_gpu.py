"""The host-side mirror of the reference interface (espflix_b200/host: class MpegDecoder / Frame,
push_video, video_isr) driven like the reference's own app drives the original: TS Buffers in,
push_video frames out. Reads like the oracle harness because it is the same protocol."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from espflix_b200 import build as ef_build
from espflix_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def _run_cli(ts_path, tmp_path, ntsc=None):
    out = os.path.join(str(tmp_path), "out.i420")
    cmd = [ef_build.HOST_CLI, ts_path, out]
    field = None
    if ntsc is not None:
        field = os.path.join(str(tmp_path), "field.u16")
        cmd += [field, str(ntsc)]
    r = subprocess.run(cmd, capture_output=True, timeout=600, check=True)
    frames = np.fromfile(out, dtype=np.uint8).reshape(-1, 101376)
    assert json.loads(r.stdout.decode().strip().splitlines()[-1])["frames"] == frames.shape[0]
    return frames, (np.fromfile(field, dtype=np.uint16) if field else None)


@pytest.mark.parametrize("name", ["splash", "vmedia"])
def test_mirrored_decoder_on_reference_fixture(name, tmp_path, oracle):
    pins = json.load(open(os.path.join(G, "decode_pins.json")))[name]
    frames, field = _run_cli(os.path.join(G, name + ".ts"), tmp_path, ntsc=1)
    assert frames.shape[0] == pins["frames"]
    assert hashlib.sha256(frames.tobytes()).hexdigest() == pins["i420_sha256"]
    assert np.array_equal(field, oracle.field(frames[-1], 1, 0))


def test_mirrored_decoder_on_synthetic_pal(tmp_path, oracle):
    es, off = synth.generate(synth.SEED0 + 9, n_pictures=12, slices=5, flags=synth.MBQUANT)
    ts = synth.wrap_ts(es, off)
    p = os.path.join(str(tmp_path), "in.ts")
    open(p, "wb").write(ts.tobytes())
    frames, field = _run_cli(p, tmp_path, ntsc=0)
    want = oracle.decode_ts(ts)
    assert np.array_equal(frames, want)
    assert np.array_equal(field, oracle.field(frames[-1], 0, 0))


def test_mirrored_index_builder_on_reference_fixture(tmp_path):
    """make_index() x 3 + merge_index() with the reference tool's own signatures (espflix_b200/host/ef_indexer.h):
    the video.idx it writes equals the image the unmodified reference tool produced for the same three streams."""
    paths = [os.path.join(G, n + ".ts") for n in ("vmedia", "splash", "vmedia")]
    subprocess.run([ef_build.INDEXER_CLI] + paths + [str(tmp_path)], capture_output=True, timeout=600, check=True)
    got = open(os.path.join(str(tmp_path), "video.idx"), "rb").read()
    assert got == open(os.path.join(G, "video_idx_fixture.bin"), "rb").read()


@pytest.mark.parametrize("name", ["ntsc_fc1", "pal_fc5", "ntsc_fc0_quirk", "ntsc_poster3", "pal_poster2"])
def test_mirrored_presentation_pacing(name, tmp_path):
    """push_video with the offline PTS -> field pacing (ef_set_video_pacing): the decoder mirror pushes its
    frames, push_video drives video_isr until each frame has flipped, and the stream of fields that reaches the
    sink is - sample for sample - what the unmodified reference's push_video + video_isr emit for the same
    pictures and PTS (tests/golden/pacing_pins.json, instant-decoder model)."""
    p = json.load(open(os.path.join(G, "pacing_pins.json")))[name]
    es, off = synth.generate(synth.SEED0 + p["seed"], n_pictures=p["pictures"])
    ts_path = os.path.join(str(tmp_path), "in.ts")
    open(ts_path, "wb").write(synth.wrap_ts(es, off).tobytes())
    out, fields = os.path.join(str(tmp_path), "out.i420"), os.path.join(str(tmp_path), "fields.u16")
    r = subprocess.run([ef_build.HOST_CLI, ts_path, out, "--paced", fields, str(p["ntsc"]), str(p["frame_counter0"]), str(p["max_fields"]), str(p["modes"][-1])],
                       capture_output=True, timeout=600, check=True)       # the last argument: flush_picture(mode); 2 / 3 = a poster scrolling in (load_poster)
    info = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert info["frames"] == p["pictures"] and info["fields"] == p["fields"]
    stream = np.fromfile(fields, dtype=np.uint16)
    assert stream.nbytes == p["stream_bytes"]
    assert hashlib.sha256(stream.tobytes()).hexdigest() == p["stream_sha256"]


@pytest.mark.parametrize("name", ["splash", "vmedia"])
def test_mirrored_program_with_audio(name, tmp_path, oracle):
    """A whole TS program through the mirror: pictures via push_video (batched decode underneath), the SBC audio of
    PID 0x101/0x102 via push_audio -> ef_audio_drain (the offline audio thread) -> sink: PCM and PDM equal the pins
    the unmodified reference produced, and the CLI reports the level-1 drop-in rate of one stream."""
    vp = json.load(open(os.path.join(G, "decode_pins.json")))[name]
    ap = json.load(open(os.path.join(G, "audio_pins.json")))[name]
    out, pcm_p, pdm_p = [os.path.join(str(tmp_path), n) for n in ("out.i420", "a.pcm", "a.pdm")]
    r = subprocess.run([ef_build.HOST_CLI, os.path.join(G, name + ".ts"), out, "--audio", pcm_p, pdm_p], capture_output=True, timeout=600, check=True)
    info = json.loads(r.stdout.decode().strip().splitlines()[-1])
    frames = np.fromfile(out, dtype=np.uint8)
    assert info["frames"] == vp["frames"] and hashlib.sha256(frames.tobytes()).hexdigest() == vp["i420_sha256"]
    assert info["frames_per_s"] > 0 and info["audio_frames"] == ap["n_frames"]
    pcm, pdm = np.fromfile(pcm_p, dtype=np.int16), np.fromfile(pdm_p, dtype=np.uint16)
    for a, b in ap["undefined"]:
        pcm[a:b] = 0
    assert hashlib.sha256(pcm.tobytes()).hexdigest() == ap["pcm_sha256_masked"]
    assert hashlib.sha256(pdm[:ap["pdm_defined_words"]].tobytes()).hexdigest() == ap["pdm_sha256_defined"]
    print("level-1 drop-in, %s: %.0f frames/s (one stream, %d pictures)" % (name, info["frames_per_s"], info["frames"]))
