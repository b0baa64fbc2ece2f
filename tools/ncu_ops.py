#!/usr/bin/env python3
"""tools/ncu_ops.py <report.ncu-rep> <units> — executed-instruction histogram by SASS opcode of the first launch
in the report, per unit of work (e.g. macroblocks per launch)."""
import csv, collections, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = None; ops = collections.Counter(); tot = 0; launch = 0
for r in rows:
    if r and r[0] == "Address":
        h = r; launch += 1; continue
    if h is None or launch != 1 or len(r) < len(h):
        continue
    try:
        c = int(r[h.index("Instructions Executed")])
    except ValueError:
        continue
    s = r[h.index("Source")].strip().split()
    op = s[1] if s[0].startswith("@") else s[0]
    full = op
    op = op.split(".")[0]
    if op == "IMAD":
        op = "IMAD.MOV/IADD/SHL" if any(k in full for k in (".MOV", ".IADD", ".SHL")) else "IMAD"
    ops[op] += c; tot += c
print("total %d = %.1f per unit" % (tot, tot / units))
for k, v in ops.most_common(32):
    print("%-18s %5.1f%%  %6.1f" % (k, 100 * v / tot, v / units))
