#!/usr/bin/env python3
"""tools/summarize_profiles.py — turn ncu captures brought back in gpurun_out/ into the tracked
summaries under profiles/ (run here, no GPU needed).

  python tools/summarize_profiles.py <tag> [--launches gpurun_out/launches.csv]
                                     [--k1a gpurun_out/k1a.ncu-rep --k1b gpurun_out/k1b.ncu-rep] [--k2 gpurun_out/k2.ncu-rep] [--bench gpurun_out/bench.json]

Writes profiles/<tag>_launches.csv + _launches_summary.txt (per-kernel share of the step),
profiles/<tag>_k1_ncu.json / _k2_ncu.json (DRAM bytes, instruction counts, issue/stall metrics, per
launch), profiles/<tag>_k1a_details.txt / _k1b_details.txt (ncu --page details), profiles/<tag>_bench.json, and refreshes
profiles/k1_traffic.json, which bench.py reads for roofline.traffic.
"""
import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.per_cycle_active", "sm__cycles_active.avg",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def raw_summary(rep, kernel_filter):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    h, units = rows[0], rows[1]
    ik = h.index("Kernel Name")
    out = []
    for r in rows[2:]:
        if kernel_filter not in r[ik]:
            continue
        d = {"kernel": r[ik].split("(")[0]}
        for n in KEEP:
            if n in h:
                i = h.index(n)
                try:
                    d[n] = float(r[i])
                    d[n + ".unit"] = units[i]
                except ValueError:
                    pass
        stalls = {}
        for i, n in enumerate(h):
            if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued"):
                try:
                    stalls[n.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i])
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1.0
        d["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]}
        out.append(d)
    return out


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * mult.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--k1a", help="ncu --set full report holding ef_parse_kernel launches (one per 12-picture step)")
    ap.add_argument("--k1b", help="ncu --set full report holding ef_recon_kernel launches (one per picture index)")
    ap.add_argument("--pictures", type=int, default=12, help="picture indices covered by one ef_parse_kernel launch")
    ap.add_argument("--k2")
    ap.add_argument("--bench")
    a = ap.parse_args()
    os.makedirs(P, exist_ok=True)
    if a.launches:
        shutil.copy(a.launches, os.path.join(P, a.tag + "_launches.csv"))
        lines = open(a.launches).read().splitlines()
        start = [i for i, line in enumerate(lines) if line.startswith('"ID"')][0]
        rows = list(csv.reader(lines[start:]))
        hdr = rows[0]
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        t = collections.defaultdict(list)
        for r in rows[1:]:
            if len(r) > iv:
                try:
                    t[r[ik].split("(")[0]].append(float(r[iv]))
                except ValueError:
                    pass
        tot = sum(sum(v) for v in t.values())
        out = ["# %s - ncu launch list (gpu__time_duration.sum; cold-cache, serialised: compare SHARES)" % a.tag, ""]
        for k, v in sorted(t.items(), key=lambda x: -sum(x[1])):
            out.append("%-44s launches %4d  total %11.1f us  avg %9.1f us  share %5.1f%%" % (k, len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
        open(os.path.join(P, a.tag + "_launches_summary.txt"), "w").write("\n".join(out) + "\n")
        print("\n".join(out))
    if a.k1a and a.k1b:
        pa = raw_summary(a.k1a, "ef_parse_kernel")
        pb = raw_summary(a.k1b, "ef_recon_kernel")
        json.dump({"ef_parse_kernel": pa, "ef_recon_kernel": pb}, open(os.path.join(P, a.tag + "_k1_ncu.json"), "w"), indent=1)
        open(os.path.join(P, a.tag + "_k1a_details.txt"), "w").write(ncu(["-i", a.k1a, "--page", "details"]))
        open(os.path.join(P, a.tag + "_k1b_details.txt"), "w").write(ncu(["-i", a.k1b, "--page", "details"]))
        if pa and pb:
            def avg(rows, key):
                return sum(to_bytes(d[key], d[key + ".unit"]) for d in rows) / len(rows)
            ra, wa = avg(pa, "dram__bytes_read.sum"), avg(pa, "dram__bytes_write.sum")
            rb, wb = avg(pb, "dram__bytes_read.sum"), avg(pb, "dram__bytes_write.sum")
            per_pic = (ra + wa) / a.pictures + rb + wb
            json.dump({"kernel": "K1 = ef_parse_kernel (1 launch per %d picture indices) + ef_recon_kernel (1 launch per picture index)" % a.pictures,
                       "source": a.tag + "_k1_ncu.json (ncu --set full; 4096 streams)",
                       "parse_dram_bytes_per_launch": ra + wa, "recon_dram_bytes_per_launch": rb + wb,
                       "dram_bytes_per_launch": per_pic,
                       "note": "per picture index over the batch = parse launch / %d + one recon launch, the unit bench.py uses for roofline.achieved" % a.pictures},
                      open(os.path.join(P, "k1_traffic.json"), "w"), indent=1)
            print("K1a dram/launch: read %.1f MB write %.1f MB; K1b: read %.1f MB write %.1f MB; per picture index %.1f MB" % (ra / 1e6, wa / 1e6, rb / 1e6, wb / 1e6, per_pic / 1e6))
    if a.k2:
        s = raw_summary(a.k2, "ef_composite_kernel")
        json.dump(s, open(os.path.join(P, a.tag + "_k2_ncu.json"), "w"), indent=1)
        open(os.path.join(P, a.tag + "_k2_details.txt"), "w").write(ncu(["-i", a.k2, "--page", "details"]))
    if a.bench:
        line = open(a.bench).read().strip().splitlines()[-1]
        json.dump(json.loads(line), open(os.path.join(P, a.tag + "_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
