#!/usr/bin/env python3
"""tools/dump_sass.py <tag> — SASS listings of the hot kernels of espflix_b200/libespflix_b200.so into profiles/<tag>_sass_*.txt
(cuobjdump -sass, encodings stripped) plus profiles/<tag>_sass_summary.txt: instruction count and the count of every
mnemonic that shows Blackwell-native machinery (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk,
UTMALDG/UTMASTG = tensor TMA, SYNCS = mbarrier, LDGSTS = cp.async, VIADDMNMX = DPX) per kernel. Runs here, no GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "espflix_b200", "libespflix_b200.so")
KERNELS = {"parse": "ef_parse_kernel", "recon": "ef_recon_kernel", "composite_ntsc": "ef_composite_kernelILb1", "composite_pal": "ef_composite_kernelILb0",
           "scan": "ef_scan_kernel", "idct_tc": "ef_idct_tc_kernel", "sbc_matrix": "ef_sbc_matrix_kernel"}
NATIVE = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "VIADDMNMX", "FENCE", "HMMA", "IMMA"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, name = collections.OrderedDict(), None
    for ln in text.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", ln)
        if m and name:
            funcs[name].append("/*%s*/ %s" % (m.group(1), m.group(2).rstrip(" ;") + " ;"))
    out_dir = os.path.join(ROOT, "profiles")
    summary = ["SASS of espflix_b200/libespflix_b200.so (sm_100a), per kernel: instructions and Blackwell-native mnemonics", ""]
    for short, pat in KERNELS.items():
        hit = [n for n in funcs if pat in n]
        if not hit:
            summary.append("%-16s not found" % short)
            continue
        ins = funcs[hit[0]]
        with open(os.path.join(out_dir, "%s_sass_%s.txt" % (tag, short)), "w") as f:
            f.write("// %s  (%d instructions)\n" % (hit[0], len(ins)))
            f.write("\n".join(ins) + "\n")
        cnt = collections.Counter()
        for i in ins:
            op = re.sub(r"^/\*\w+\*/\s+(@!?U?P\d+\s+)?", "", i).split()[0].split(".")[0]
            cnt[op] += 1
        native = ", ".join("%s x%d" % (k, cnt[k]) for k in NATIVE if cnt[k])
        summary.append("%-16s %5d instructions; %s" % (short, len(ins), native or "-"))
    with open(os.path.join(out_dir, "%s_sass_summary.txt" % tag), "w") as f:
        f.write("\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    main()
