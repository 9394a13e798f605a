#!/usr/bin/env python
"""Traversal throughput on the C5-scale instanced scene (201 x 201 ProxyObject instances of testGrass.obj + MB bullets).
usage (GPU box): tools/instance_bench.py [grid_n]"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import helpers
import miro_b200 as mb

spec = importlib.util.spec_from_file_location("make_scenes", os.path.join(ROOT, "tools", "make_scenes.py"))
ms = importlib.util.module_from_spec(spec); spec.loader.exec_module(ms)
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 201
fx = helpers.Fixture(helpers.fixture_path("c5_mb_instances"))
sc = fx.scene(script_override=ms.c5(grid, name=None)).attach(0)
rng = np.random.default_rng(5)
n = 1 << 21
o = np.stack([rng.uniform(-25, 17, n), rng.uniform(0.05, 14, n), rng.uniform(-20, 20, n)], 1)
d = rng.normal(size=(n, 3)); d[:, 1] = -np.abs(d[:, 1]) * 0.5; d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = mb.make_rays(o, d); rays["time"] = rng.uniform(0, 1, n).astype(np.float32)
cam = np.array([-4, 12, 26.0]); tgt = np.stack([rng.uniform(-19, 11, n), np.zeros(n), rng.uniform(-15, 15, n)], 1)
dd = tgt - cam; dd /= np.linalg.norm(dd, axis=1, keepdims=True)
prim = mb.make_rays(np.tile(cam, (n, 1)), dd)
stream = torch.cuda.Stream(); sc.set_stream(stream.cuda_stream)
for name, r in [("incoherent", rays), ("camera-to-field", prim)]:
    d_rays = torch.from_numpy(r.view(np.uint8).reshape(n, -1)).cuda()
    d_hits = torch.empty((n, 20), dtype=torch.uint8, device="cuda")
    sc.enable_counting(True); sc.reset_counters(); sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr()); c = sc.counters(); sc.enable_counting(False)
    for _ in range(3):
        sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr())
    e1.record(stream); stream.synchronize()
    ms_ = e0.elapsed_time(e1) / 10
    print(json.dumps({"batch": name, "instances": grid * grid, "rays": n, "ms": ms_, "Mrays_per_s": n / ms_ * 1e-3, "nodes_per_ray": c["nodes_fetched"] / n,
                      "tris_per_ray": c["tris_tested"] / n, "instances_per_ray": c["insts_entered"] / n}))
sc.set_stream(None); sc.close()
