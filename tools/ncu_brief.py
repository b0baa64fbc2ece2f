#!/usr/bin/env python3
"""tools/ncu_brief.py <report.ncu-rep> [kernel-substring] — the handful of ncu raw metrics that matter here
(duration, instructions, IPC, occupancy, DRAM bytes, pipe use, stall shares), per captured launch."""
import csv, io, subprocess, sys
rep = sys.argv[1]; flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); h = rows[0]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__cycles_active.avg", "gpc__cycles_elapsed.max",
        "sm__warps_active.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if flt not in name:
        continue
    print(name.split("(")[0])
    for w in want:
        if w in h:
            print("   %-75s %s %s" % (w, r[h.index(w)], rows[1][h.index(w)]))
    st = {n.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i]) for i, n in enumerate(h)
          if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued") and r[i]}
    tot = sum(st.values()) or 1
    print("   stalls:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
