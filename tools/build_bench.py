#!/usr/bin/env python
"""Host SAH build vs device LBVH build: build time and traversal throughput of the resulting trees.
usage (GPU box): tools/build_bench.py"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np
import torch
import helpers
import miro_b200 as mb
import test_synthetic_gpu as T


def throughput(sc, rays):
    n = len(rays)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(n, -1)).cuda()
    d_hits = torch.empty((n, 20), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream(); sc.set_stream(stream.cuda_stream)
    sc.enable_counting(True); sc.reset_counters(); sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr()); c = sc.counters(); sc.enable_counting(False)
    for _ in range(3):
        sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr())
    e1.record(stream); stream.synchronize()
    sc.set_stream(None)
    ms = e0.elapsed_time(e1) / 10
    return {"Mrays_per_s": n / ms * 1e-3, "nodes_per_ray": c["nodes_fetched"] / n, "tris_per_ray": c["tris_tested"] / n}


def soup_scene(v, f, device_build):
    sc = mb.MiroScene(); sc.preload_mesh("m", v, f)
    with tempfile.NamedTemporaryFile("w", suffix=".miro", delete=False) as fh:
        fh.write("image 64 64\nscene devicebuild %d\nmaterial g lambert kd 0.7 0.7 0.7\nmesh m m.obj\nobject m g\n" % device_build)
    t0 = time.time(); sc.load_script(fh.name, "/nonexistent"); host_s = time.time() - t0
    os.unlink(fh.name)
    return sc, host_s


def upload_time(sc):
    """Second miro_gpu_upload_scene call (context, allocator and cub temp storage warm): H2D + (device build)."""
    d = sc.desc()
    torch.cuda.synchronize(); t0 = time.time()
    rc = sc.L.miro_gpu_upload_scene(sc.ctx, C.byref(d)); torch.cuda.synchronize()
    assert rc == 0
    return time.time() - t0


cases = []
fx = helpers.Fixture(helpers.fixture_path("c2_explosion", full=True) or helpers.fixture_path("c2_explosion"))
allv = np.concatenate([fx.mesh(k)["vertices"] for k in range(len(fx.names))])
for name, maker, rays in [("explosion01 (86 914 tris)", lambda db: (fx.scene(script_override=fx.script.replace("scene ", "scene devicebuild %d " % db, 1)), None),
                          T.rays_for(allv, 1 << 21, 3)),
                         ("soup 1 M tris", None, None)]:
    if maker is None:
        v, f = T.soup(1_000_000, 99, 0.004, False)
        maker = lambda db: soup_scene(v, f, db)
        rays = T.rays_for(v, 1 << 21, 3)
    for db in ((1,) if os.environ.get("LBVH_ONLY") else (0, 1)):
        t0 = time.time(); sc, host_s = maker(db); host_s = host_s if host_s is not None else time.time() - t0
        sc.attach(0)
        up = upload_time(sc)
        r = throughput(sc, rays)
        print(json.dumps({"scene": name, "build": "device LBVH" if db else "host SAH", "host_preCalc_s": round(host_s, 3), "upload_s": round(up, 4), **{k: round(v, 2) for k, v in r.items()}}))
        sc.close()
