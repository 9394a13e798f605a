#!/usr/bin/env python
"""The three traversal kernels on the bench workloads at FULL size (3 x 2 073 600 rays each on c2, big and c5): closest-hit records
and occlusion answers must be byte-identical (run on the GPU box).  usage: python tools/kernel_identity.py [c2 big c5]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
import bench_workloads as bw
import bench

for name in (sys.argv[1:] or ["c2", "big", "c5"]):
    w = bw.Workload(name).load(helpers.Fixture)
    sc = bench.product_scene(w, None, 0)
    prim = w.primary(); inco = w.incoherent(1)
    out = {}
    for k in ("warp", "flat", "pool"):
        sc.set_trace_kernel(k)
        hp = sc.trace_closest(prim); hi = sc.trace_closest(inco)
        sh = w.shadow(inco, hi["t"], hi["prim"] >= 0)
        out[k] = (hp.tobytes(), hi.tobytes(), sc.trace_any(sh).tobytes(), sc.trace_any(inco).tobytes())
    same = {k: [a == b for a, b in zip(out["warp"], out[k])] for k in ("flat", "pool")}
    print(name, "rays", len(prim) + 2 * len(inco) + len(sh), "hit fraction %.3f" % (np.frombuffer(out["warp"][1], dtype=hi.dtype)["prim"] >= 0).mean(), same, flush=True)
    assert all(all(v) for v in same.values()), same
    sc.close()
print("identical")
