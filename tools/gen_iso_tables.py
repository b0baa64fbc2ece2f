#!/usr/bin/env python3
"""tools/gen_iso_tables.py — emits espflix_b200/csrc/ef_iso11172_tables.h.

The MPEG-1 video VLC tables (ISO/IEC 11172-2 Annex B, tables B.1-B.5) written out as
(code-string, value) lists, plus the small numeric tables of the decode path. The reference
holds the same tables as binary trees (player.cpp:59-148) and a nested-prefix fast decoder
(player.cpp:532-644); tests/test_tables.py checks this file against the enumeration of those
trees committed in tests/golden/vlc_codes.json.
"""
import math, os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# B.1 macroblock_address_increment (34 = stuffing, 35 = escape; values as returned by the
# reference tree, player.cpp:59-70)
MBA = {
    1: "1", 2: "011", 3: "010", 4: "0011", 5: "0010", 6: "00011", 7: "00010",
    8: "0000111", 9: "0000110", 10: "00001011", 11: "00001010", 12: "00001001", 13: "00001000",
    14: "00000111", 15: "00000110", 16: "0000010111", 17: "0000010110", 18: "0000010101",
    19: "0000010100", 20: "0000010011", 21: "0000010010",
}
for k in range(22, 34):            # 22..33: 11-bit codes 00000100011 .. 00000011000 descending
    MBA[k] = format(0b00000100011 - (k - 22), "011b")
MBA[34] = "00000001111"
MBA[35] = "00000001000"

# B.2a / B.2b macroblock_type: flags 0x10 quant, 0x08 motion forward, 0x02 pattern, 0x01 intra
MBTYPE_I = {0x01: "1", 0x11: "01"}
MBTYPE_P = {0x0A: "1", 0x02: "01", 0x08: "001", 0x01: "00011", 0x1A: "00010", 0x12: "00001", 0x11: "000001"}

# B.3 coded_block_pattern
CBP = {
    60: "111", 4: "1101", 8: "1100", 16: "1011", 32: "1010", 12: "10011", 48: "10010", 20: "10001",
    40: "10000", 28: "01111", 44: "01110", 52: "01101", 56: "01100", 1: "01011", 61: "01010",
    2: "01001", 62: "01000", 24: "001111", 36: "001110", 3: "001101", 63: "001100",
    5: "0010111", 9: "0010110", 17: "0010101", 33: "0010100", 6: "0010011", 10: "0010010",
    18: "0010001", 34: "0010000", 7: "00011111", 11: "00011110", 19: "00011101", 35: "00011100",
    13: "00011011", 49: "00011010", 21: "00011001", 41: "00011000", 14: "00010111", 50: "00010110",
    22: "00010101", 42: "00010100", 15: "00010011", 51: "00010010", 23: "00010001", 43: "00010000",
    25: "00001111", 37: "00001110", 26: "00001101", 38: "00001100", 29: "00001011", 45: "00001010",
    53: "00001001", 57: "00001000", 30: "00000111", 46: "00000110", 54: "00000101", 58: "00000100",
    31: "000000111", 47: "000000110", 55: "000000101", 59: "000000100", 27: "000000011", 39: "000000010",
}

# B.4 motion vector codes: magnitude -> prefix, then sign bit (0 = +, 1 = -)
MV_MAG = {
    1: "01", 2: "001", 3: "0001", 4: "000011", 5: "0000101", 6: "0000100", 7: "0000011",
    8: "000001011", 9: "000001010", 10: "000001001", 11: "0000010001", 12: "0000010000",
    13: "0000001111", 14: "0000001110", 15: "0000001101", 16: "0000001100",
}
MV = {0: "1"}
for m, p in MV_MAG.items():
    MV[m] = p + "0"
    MV[-m] = p + "1"

# B.5c-g dct_coeff_next: (run, level) -> code WITHOUT the sign bit. "10" = end of block,
# "000001" = escape. (0,1) is "11" here ("1" when it is the first coefficient of a block).
DCT = {
    (0, 1): "11", (1, 1): "011", (0, 2): "0100", (2, 1): "0101", (0, 3): "00101", (3, 1): "00111",
    (4, 1): "00110", (1, 2): "000110", (5, 1): "000111", (6, 1): "000101", (7, 1): "000100",
    (0, 4): "0000110", (2, 2): "0000100", (8, 1): "0000111", (9, 1): "0000101",
    (0, 5): "00100110", (0, 6): "00100001", (1, 3): "00100101", (3, 2): "00100100",
    (10, 1): "00100111", (11, 1): "00100011", (12, 1): "00100010", (13, 1): "00100000",
    (0, 7): "0000001010", (1, 4): "0000001100", (2, 3): "0000001011", (4, 2): "0000001111",
    (5, 2): "0000001001", (14, 1): "0000001110", (15, 1): "0000001101", (16, 1): "0000001000",
}
# long codes: (7+z) zeros, a one, then four bits; rows listed for the four bits 0000..1111
LONG = [
    [(0, 11), (8, 2), (4, 3), (0, 10), (2, 4), (7, 2), (21, 1), (20, 1), (0, 9), (19, 1), (18, 1), (1, 5), (3, 3), (0, 8), (6, 2), (17, 1)],
    [(10, 2), (9, 2), (5, 3), (3, 4), (2, 5), (1, 7), (1, 6), (0, 15), (0, 14), (0, 13), (0, 12), (26, 1), (25, 1), (24, 1), (23, 1), (22, 1)],
    [(0, 31 - i) for i in range(16)],
    [(0, 40 - i) for i in range(9)] + [(1, 14 - i) for i in range(7)],
    [(1, 18), (1, 17), (1, 16), (1, 15), (6, 3), (16, 2), (15, 2), (14, 2), (13, 2), (12, 2), (11, 2), (31, 1), (30, 1), (29, 1), (28, 1), (27, 1)],
]
for z, row in enumerate(LONG):
    for i, rl in enumerate(row):
        DCT[rl] = "0" * (7 + z) + "1" + format(i, "04b")

# B.5a / B.5b dct_dc_size
DC_LUMA = {0: "100", 1: "00", 2: "01", 3: "101", 4: "110", 5: "1110", 6: "11110", 7: "111110", 8: "1111110"}
DC_CHROMA = {0: "00", 1: "01", 2: "10", 3: "110", 4: "1110", 5: "11110", 6: "111110", 7: "1111110", 8: "11111110"}

ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,
          7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
          39, 46, 53, 60, 61, 54, 47, 55, 62, 63]

DEFAULT_INTRA_Q = [
    8, 16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34, 34, 38,
    22, 22, 26, 27, 29, 34, 37, 40, 22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83]


def aan_prescale():
    """The reference's 8-bit AAN prescale (player.cpp:161-170): round(32 * s_i * s_j) with
    s_0 = 1, s_k = sqrt(2) * cos(k*pi/16). 32 == 1.0."""
    s = [1.0] + [math.sqrt(2.0) * math.cos(k * math.pi / 16) for k in range(1, 8)]
    return [int(math.floor(32 * s[i] * s[j] + 0.5)) for i in range(8) for j in range(8)]


def emit_table(name, d, keyfmt):
    lines = [f"static const ef_vlc_code {name}[] = {{"]
    row = []
    for k, code in d.items():
        row.append(f'{{"{code}", {keyfmt(k)}}}')
        if len(row) == 4:
            lines.append("    " + ", ".join(row) + ",")
            row = []
    if row:
        lines.append("    " + ", ".join(row) + ",")
    lines.append("};")
    lines.append(f"#define {name.upper()}_COUNT {len(d)}")
    return "\n".join(lines)


def emit_u8(name, vals, per=16):
    lines = [f"static const unsigned char {name}[{len(vals)}] = {{"]
    for i in range(0, len(vals), per):
        lines.append("    " + ", ".join(f"{v:3d}" for v in vals[i:i + per]) + ",")
    lines.append("};")
    return "\n".join(lines)


def main():
    out = os.path.join(ROOT, "espflix_b200", "csrc", "ef_iso11172_tables.h")
    parts = [
        "// GENERATED by tools/gen_iso_tables.py — do not edit.",
        "// MPEG-1 video (ISO/IEC 11172-2 Annex B) VLC tables as (code string, value) lists and the",
        "// small numeric tables of the decode path. Host-side only: ef_tables.cpp turns these into",
        "// the clz-indexed lookup tables the kernels read. Checked against the reference's VLC trees",
        "// (player.cpp:59-148) by tests/test_tables.py.",
        "#ifndef EF_ISO11172_TABLES_H",
        "#define EF_ISO11172_TABLES_H",
        "",
        "typedef struct { const char* code; int value; } ef_vlc_code;",
        "",
        "// B.1 macroblock_address_increment; 34 = macroblock_stuffing, 35 = macroblock_escape",
        emit_table("ef_vlc_mba", MBA, str),
        "// B.2a macroblock_type, I pictures (0x10 quant | 0x08 fwd | 0x02 pattern | 0x01 intra)",
        emit_table("ef_vlc_mbtype_i", MBTYPE_I, lambda k: f"0x{k:02X}"),
        "// B.2b macroblock_type, P pictures",
        emit_table("ef_vlc_mbtype_p", MBTYPE_P, lambda k: f"0x{k:02X}"),
        "// B.3 coded_block_pattern",
        emit_table("ef_vlc_cbp", CBP, str),
        "// B.4 motion_horizontal/vertical_forward_code (sign bit included)",
        emit_table("ef_vlc_mv", MV, str),
        "// B.5c-g dct_coeff_next WITHOUT sign bit; value = (run << 8) | level. End of block is",
        "// \"10\", escape is \"000001\"; neither is listed. (0,1) is \"1\" as dct_coeff_first.",
        emit_table("ef_vlc_dct", DCT, lambda k: f"0x{(k[0] << 8) | k[1]:04X}"),
        "// B.5a dct_dc_size_luminance / B.5b dct_dc_size_chrominance",
        emit_table("ef_vlc_dc_luma", DC_LUMA, str),
        emit_table("ef_vlc_dc_chroma", DC_CHROMA, str),
        "",
        "// zig-zag scan position -> raster index (ISO 11172-2 2.4.4.1; reference player.cpp:150)",
        emit_u8("ef_zigzag", ZIGZAG),
        "// default intra quantiser matrix, raster order (reference player.cpp:172)",
        emit_u8("ef_default_intra_q", DEFAULT_INTRA_Q),
        "// the reference's 8-bit AAN IDCT prescale, round(32*s_i*s_j) (player.cpp:161)",
        emit_u8("ef_aan_prescale", aan_prescale(), 8),
        "",
        "#endif",
        "",
    ]
    open(out, "w").write("\n".join(parts))
    print("wrote", out)


if __name__ == "__main__":
    main()
