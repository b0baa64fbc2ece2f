"""Audio path on the CPU: the oracle restatement (oracle/ef_oracle_audio.c) against the pins the
unmodified reference produced on its own fixtures (tests/golden/audio_pins.json, tools/make_audio_golden.py), against
the reference itself on synthetic transport streams (wherever oracle/_ref exists), and the constant tables."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests import audio_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
REF_AUDIO = os.path.join(ROOT, "oracle", "_ref", "efref_audio")


def _masked(pcm, ranges):
    p = pcm.copy()
    for a, b in ranges:
        p[a:b] = 0
    return p


@pytest.mark.parametrize("name", ["splash", "vmedia"])
def test_oracle_matches_reference_pins(oracle, name):
    pins = json.load(open(os.path.join(G, "audio_pins.json")))[name]
    es = oracle.demux_audio_ts(open(os.path.join(G, name + ".ts"), "rb").read())
    assert es.size == pins["es_bytes"] and hashlib.sha256(es.tobytes()).hexdigest() == pins["es_sha256"]
    pcm = oracle.sbc_decode(es)
    assert pcm.size == pins["n_frames"] * 128 and [int(x) for x in pcm[:16]] == pins["pcm_head"]
    assert hashlib.sha256(_masked(pcm, pins["undefined"]).tobytes()).hexdigest() == pins["pcm_sha256_masked"]
    pdm = oracle.pdm(pcm)
    k = pins["pdm_defined_words"]
    assert hashlib.sha256(pdm[:k].tobytes()).hexdigest() == pins["pdm_sha256_defined"]


def test_sbc_tables_match_reference_arrays():
    """espflix_b200/csrc/ef_sbc_tables.h (closed-form matrix, embedded spec window, offsets) against the reference's
    arrays as committed by tools/make_audio_golden.py: SBC_syn_8[i][j] = matrix[i][j]; SBC_proto_8[i][t] = window[t][i]."""
    import re
    g = json.load(open(os.path.join(G, "sbc_tables.json")))
    hdr = open(os.path.join(ROOT, "espflix_b200", "csrc", "ef_sbc_tables.h")).read()

    def table(name):
        i = hdr.index(name)
        return [int(x) for x in re.findall(r"-?\d+", hdr[hdr.index("{", i):hdr.index("};", i)])]

    assert table("ef_sbc_matrix[16][8]") == g["SBC_syn_8"]
    w = np.array(table("ef_sbc_window[10][8]")).reshape(10, 8)
    assert w.T.reshape(-1).tolist() == g["SBC_proto_8"]
    assert table("ef_sbc_offset8[4][8]") == g["SBC_offset8"]


@pytest.mark.skipif(not os.path.exists(REF_AUDIO), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("case", ["bp28", "snr_bp12_f0", "bp60_loud", "rejected_frames", "muted_pes", "pid102"])
def test_oracle_matches_reference_on_synthetic_streams(oracle, tmp_path, case):
    kw = {"bp28": dict(bitpool=28), "snr_bp12_f0": dict(bitpool=12, allocation=1, frequency=0), "bp60_loud": dict(bitpool=60, loud=True, frequency=3),
          "rejected_frames": dict(bitpool=28, bad_frames=(3, 4, 17)), "muted_pes": dict(bitpool=28), "pid102": dict(bitpool=28, frequency=1)}[case]
    es = audio_cases.sbc_stream(1000 + len(case), 40, **kw)
    ts = audio_cases.mux_audio_ts(es, pid=0x102 if case == "pid102" else 0x101, drop_pts_on=(1,) if case == "muted_pes" else ())
    p = tmp_path / "a.ts"
    p.write_bytes(ts.tobytes())
    out = tmp_path / "a.bin"
    subprocess.run([REF_AUDIO, str(p), str(out)], check=True, capture_output=True, timeout=120)   # one process per run: the reference keeps its state in globals
    raw = out.read_bytes()
    nes, npcm = [int(x) for x in np.frombuffer(raw[:16], dtype=np.uint64)]
    ref_es = np.frombuffer(raw[16:16 + nes], dtype=np.uint8)
    ref_pcm = np.frombuffer(raw[16 + nes:16 + nes + 2 * npcm], dtype=np.int16)
    ref_pdm = np.frombuffer(raw[16 + nes + 2 * npcm:], dtype=np.uint16)
    got_es = oracle.demux_audio_ts(ts)
    assert np.array_equal(got_es, ref_es)
    if case == "muted_pes":
        assert got_es.size == es.size - 1024              # the second PES (no PTS) is dropped whole; the third one opens the stream again
    pcm = oracle.sbc_decode(got_es)
    assert npcm > 0 and np.array_equal(pcm, ref_pcm)
    assert np.array_equal(oracle.pdm(pcm), ref_pdm)
