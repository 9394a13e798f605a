#!/usr/bin/env python
"""Fold one `ncu --set full --page raw --csv` capture of the three traversal launches of a bench step (primary, incoherent, shadow —
tools/ncu_trace.sh / tools/r2_ncu.sh) into profiles/traffic.json, which bench.py reads for roofline.traffic / l2_traffic /
l1_global_load_traffic.   usage: tools/traffic_from_ncu.py WORKLOAD RAW.csv "kernel version" "capture command" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["k_trace<closest> primary", "k_trace<closest> incoherent", "k_trace<any> shadow"]      # bench.py KERNEL_NAMES, launch order
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "sector": 32.0}


def column(rows, name):
    i = rows[0].index(name)
    return [float(r[i].replace(",", "")) * UNIT[rows[1][i]] for r in rows[2:5]]


def traffic_of(path):
    """(captured kernel names, {launch: DRAM bytes}, {launch: L2 bytes}, {launch: L1 global-load bytes}) of one raw page."""
    rows = list(csv.reader(open(path)))
    dram = [a + b for a, b in zip(column(rows, "dram__bytes_read.sum"), column(rows, "dram__bytes_write.sum"))]
    l2 = column(rows, "lts__t_sectors.sum")
    l1 = column(rows, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
    kernels = [r[rows[0].index("Kernel Name")].split("(")[0].replace("void ", "") for r in rows[2:5]]
    return kernels, dict(zip(NAMES, dram)), dict(zip(NAMES, l2)), dict(zip(NAMES, l1))


def main():
    key, path, version, capture = sys.argv[1:5]
    kernels, dram, l2, l1 = traffic_of(path)
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    tj = json.load(open(tp))
    note = tj.get(key, {}).get("note")
    tj[key] = {"kernel_version": version, "capture": capture, "captured_kernels": kernels,
               "per_launch_dram_bytes": dram, "per_launch_l2_bytes": l2, "per_launch_l1_global_load_bytes": l1}
    if note:
        tj[key]["note"] = note
    json.dump(tj, open(tp, "w"), indent=1)
    print(key, {n: "%.1f MB DRAM, %.1f MB L2, %.1f MB L1" % (dram[n] * 1e-6, l2[n] * 1e-6, l1[n] * 1e-6) for n in NAMES})


if __name__ == "__main__":
    main()
