#!/usr/bin/env python
"""Print the metrics we track from an `ncu --page raw --csv` dump (one column per captured launch)."""
import csv
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active"]
STALL = "smsp__average_warps_issue_stalled_"

rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
for w in WANT:
    if w in h:
        i = h.index(w)
        print(w[-64:].ljust(66), [r[i][:22] for r in rows[2:]])
for i, name in enumerate(h):
    if name.startswith(STALL) and name.endswith("_per_issue_active.ratio"):
        vals = [r[i] for r in rows[2:]]
        try:
            if max(float(v.replace(",", "")) for v in vals) >= 0.15:
                print(("stall " + name[len(STALL):-len("_per_issue_active.ratio")]).ljust(66), [v[:6] for v in vals])
        except ValueError:
            pass
