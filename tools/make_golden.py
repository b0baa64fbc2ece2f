#!/usr/bin/env python3
"""tools/make_golden.py — regenerates tests/golden/ from the reference. Runs ONLY in the build
container (needs /root/reference and oracle/_ref built by `make -C oracle ref`).

Outputs (committed):
  tests/golden/splash.ts, vmedia.ts   the reference's embedded media fixtures
                                       (src/splash.h:12 `splash_ts`, src/vmedia.h:1 `vmedia`), byte-for-byte
  tests/golden/decode_pins.json       per-frame SHA-256 of the I420 dump produced by the UNMODIFIED
                                       reference decoder (oracle/_ref/efref_decode) + whole-dump hashes
  tests/golden/composite_pins.json    SHA-256 of whole composite fields produced by the reference
                                       video_isr (oracle/_ref/libefref_vid.so), geometry and colour LUT hashes
  tests/golden/frames_*.i420          a few raw oracle frames (inputs for the composite tests on the GPU box)
  tests/golden/vlc_codes.json         the reference's VLC trees (player.cpp:59-148) enumerated as
                                       code-string -> value through the compiled reference arrays
"""
import ctypes, hashlib, json, os, re, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
G = os.path.join(ROOT, "tests", "golden")
I420 = 352 * 192 * 3 // 2


def c_array_bytes(path, name):
    s = open(path, encoding="utf-8", errors="replace").read()
    i = s.index(name)
    i = s.index("{", i)
    j = s.index("};", i)
    vals = re.findall(r"0x([0-9A-Fa-f]{2})", s[i:j])
    return bytes(int(v, 16) for v in vals)


def c_array_u32(path, name):
    s = open(path, encoding="utf-8", errors="replace").read()
    i = s.index(name)
    i = s.index("{", i)
    j = s.index("};", i)
    return [int(v, 16) for v in re.findall(r"0x([0-9A-Fa-f]{8})", s[i:j])]


def enumerate_tree(vlc):
    """Walk a reference VLC tree (node word: bits31-24 next-on-0, 23-16 next-on-1, 0xFF invalid;
    leaf iff top byte == 0, value = low 16 bits as int16; player.cpp:516-530)."""
    out = {}
    def rec(state, prefix):
        for bit in (0, 1):
            nxt = (vlc[state] >> (16 if bit else 24)) & 0xFF
            if nxt == 0xFF:
                continue
            code = prefix + str(bit)
            if (vlc[nxt] >> 24) == 0:
                v = vlc[nxt] & 0xFFFF
                out[code] = v - 65536 if v >= 32768 else v
            else:
                rec(nxt, code)
    rec(0, "")
    return out


def main():
    os.makedirs(G, exist_ok=True)
    pins = {}
    frames = {}
    for name, hdr, sym in (("splash", "splash.h", "splash_ts"), ("vmedia", "vmedia.h", "vmedia")):
        ts = c_array_bytes(os.path.join(REF, hdr), sym)
        open(os.path.join(G, name + ".ts"), "wb").write(ts)
        tmp = os.path.join("/tmp", name + ".i420")
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "efref_decode"), os.path.join(G, name + ".ts"), tmp],
                           capture_output=True, timeout=120, check=True)
        n = json.loads(r.stdout)["frames"]
        out = np.fromfile(tmp, dtype=np.uint8)
        assert out.size == n * I420
        frames[name] = out.reshape(n, I420)
        pins[name] = {
            "ts_bytes": len(ts), "ts_sha256": hashlib.sha256(ts).hexdigest(),
            "frames": int(n), "i420_sha256": hashlib.sha256(out.tobytes()).hexdigest(),
            "frame_sha256": [hashlib.sha256(frames[name][k].tobytes()).hexdigest() for k in range(n)],
        }
        print(name, len(ts), n, pins[name]["i420_sha256"])
    json.dump(pins, open(os.path.join(G, "decode_pins.json"), "w"), indent=1)

    # a few raw frames for composite tests: vmedia 0, 5, 40 ; splash 20
    keep = {"vmedia": [0, 5, 40], "splash": [20]}
    for name, ks in keep.items():
        for k in ks:
            frames[name][k].tofile(os.path.join(G, f"frame_{name}_{k}.i420"))

    vid = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libefref_vid.so"))
    vid.efref_field.restype = ctypes.c_long
    vid.efref_field.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    comp = {"fields": [], "geometry": {}, "color_tab_sha256": {}, "pal_burst": {}}
    for std in (1, 0):
        vid.efref_video_init(std)
        g = (ctypes.c_int * 8)()
        vid.efref_geometry(g)
        comp["geometry"]["ntsc" if std else "pal"] = list(g)
        tab = np.zeros(768, dtype=np.uint32)
        vid.efref_color_tab(tab.ctypes.data)
        comp["color_tab_sha256"]["ntsc" if std else "pal"] = hashlib.sha256(tab.tobytes()).hexdigest()
        tab.tofile(os.path.join(G, f"color_tab_{'ntsc' if std else 'pal'}.u32"))
        if not std:
            b0 = np.zeros(64, dtype=np.int16); b1 = np.zeros(64, dtype=np.int16)
            w = vid.efref_pal_burst(b0.ctypes.data, b1.ctypes.data)
            comp["pal_burst"] = {"width": int(w), "burst0": [int(x) for x in b0[:w]], "burst1": [int(x) for x in b1[:w]]}
        lw, lc = g[0], g[1]
        for name, k in (("vmedia", 0), ("vmedia", 5), ("vmedia", 40), ("splash", 20)):
            for fc in (0, 1):
                out = np.zeros(lw * lc, dtype=np.uint16)
                fr = np.ascontiguousarray(frames[name][k])
                vid.efref_field(fr.ctypes.data, None, fc, 0, out.ctypes.data)
                comp["fields"].append({"src": name, "frame": k, "ntsc": std, "frame_counter": fc,
                                       "bytes": int(out.nbytes), "sha256": hashlib.sha256(out.tobytes()).hexdigest(),
                                       "max_hi": int((out >> 8).max())})
    json.dump(comp, open(os.path.join(G, "composite_pins.json"), "w"), indent=1)
    for f in comp["fields"]:
        print(f)

    trees = {
        "macroblock_address_increment": "macroblock_address_increment[75]",
        "macroblock_type_I": "macroblock_type_I[4]",
        "macroblock_type_P": "macroblock_type_P[14]",
        "coded_block_pattern": "coded_block_pattern[126]",
        "motion_vec": "motion_vec[67]",
        "dct_coeff": "dct_coeff[224]",
    }
    codes = {k: enumerate_tree(c_array_u32(os.path.join(REF, "player.cpp"), v)) for k, v in trees.items()}
    json.dump(codes, open(os.path.join(G, "vlc_codes.json"), "w"), indent=0, sort_keys=True)
    print({k: len(v) for k, v in codes.items()})


if __name__ == "__main__":
    main()
