#!/usr/bin/env python3
"""tools/ncu_lines.py <report.ncu-rep> [kernel-substring] — per source line: warp instructions executed and
stall samples of the first matching launch (ncu --page source, cuda+sass view). Run here, no GPU needed."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
per = {}
order = []
kernel = None
done = False
fname = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if kernel is not None and r[1] != kernel and per:
            pass
        kernel = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        continue
    if hdr is None or (flt and flt not in (kernel or "")):
        continue
    if r[0] != "":
        key = (fname, int(r[0]), r[1].strip()[:110])
        try:
            per[key] = per.get(key, [0, 0])
            per[key][0] += int(r[ii]); per[key][1] += int(r[isamp])
            if key not in order:
                order.append(key)
        except ValueError:
            pass
tot = sum(v[0] for v in per.values()) or 1
ts = sum(v[1] for v in per.values()) or 1
print("total warp instructions %d, samples %d" % (tot, ts))
for k in sorted(order, key=lambda k: (k[0], k[1])):
    v = per[k]
    if v[0] * 1000 >= tot or v[1] * 1000 >= ts:
        print("%-16s %4d  inst %5.1f%%  samp %5.1f%%  %s" % (k[0][:16], k[1], 100.0 * v[0] / tot, 100.0 * v[1] / ts, k[2]))
