#!/usr/bin/env python
"""Segments the SASS of one profiled kernel (ncu --page source --csv) into runs of equal execution count and prints,
per run, its share of issued instructions and the average number of active threads (SIMT utilisation)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; ia = h.index("Source"); ie = h.index("Instructions Executed"); it = h.index("Thread Instructions Executed"); isamp = h.index("# Samples")
data = []
seen = set()
for r in rows[2:]:
    try:
        if r[0] in seen:
            continue
        seen.add(r[0])
        data.append((r[ia].strip(), int(r[ie]), int(r[it]), int(r[isamp])))
    except (ValueError, IndexError):
        pass
tot = sum(d[1] for d in data); ts = sum(d[3] for d in data)
print("instructions", len(data), "executed", tot, "avg threads/inst %.2f" % (sum(d[2] for d in data) / tot))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][1] - data[i - 1][1]) > 0.02 * max(data[i][1], data[i - 1][1], 1):
        s = data[start:i]; e = sum(x[1] for x in s); t = sum(x[2] for x in s); sm = sum(x[3] for x in s)
        if e > thresh * tot:
            ops = {}
            for x in s:
                tok = x[0].split()
                op = tok[1] if tok[0].startswith('@') else tok[0]
                ops[op] = ops.get(op, 0) + 1
            top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
            print("sass %4d-%4d n=%3d exec %9d share %5.1f%% thr/inst %5.1f stall-samples %5.1f%% %s" % (start, i - 1, len(s), s[0][1], 100 * e / tot, t / e, 100 * sm / ts, top))
        start = i
