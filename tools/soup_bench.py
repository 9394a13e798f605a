#!/usr/bin/env python
"""Traversal throughput on a seeded synthetic triangle soup too large for L1 (and, at 4 M+, a good part of L2).
usage (GPU box): tools/soup_bench.py [n_triangles ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import test_synthetic_gpu as T
import miro_b200 as mb

for n in [int(a) for a in sys.argv[1:]] or [1_000_000]:
    v, f = T.soup(n, 7 + n, 0.6 / n ** (1 / 3), False)
    sc = T.scene_of(v, f).attach(0)
    rays = T.rays_for(v, 1 << 21, 5)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(len(rays), -1)).cuda()
    d_hits = torch.empty((len(rays), 20), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream(); sc.set_stream(stream.cuda_stream)
    sc.enable_counting(True); sc.reset_counters(); sc.trace_closest_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr()); c = sc.counters(); sc.enable_counting(False)
    for _ in range(3):
        sc.trace_closest_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        sc.trace_closest_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr())
    e1.record(stream); stream.synchronize()
    ms = e0.elapsed_time(e1) / 10
    st = sc.bvh_stats()
    by = c["nodes_fetched"] * 64 + c["tris_tested"] * 48 + 68 * len(rays)
    print(json.dumps({"triangles": n, "nodes": st["nodes"], "device_MB": (st["nodes"] * 64 + n * 48) / 1e6, "rays": len(rays), "ms": ms, "Mrays_per_s": len(rays) / ms * 1e-3,
                      "nodes_per_ray": c["nodes_fetched"] / len(rays), "tris_per_ray": c["tris_tested"] / len(rays), "algorithmic_GBps": by / ms * 1e-6}))
    sc.set_stream(None); sc.close()
