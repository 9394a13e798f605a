#!/usr/bin/env python
"""Quick GPU probe: parity + kernel timing of the traversal kernels on the full fixtures (development aid)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import helpers
import miro_b200 as mb

def time_trace(sc, rays, any_hit=False, iters=10):
    n = len(rays)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(n, 48)).cuda()
    d_out = torch.empty((n, 20) if not any_hit else ((n + 31) // 32 * 4,), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    sc.set_stream(stream.cuda_stream)
    f = sc.trace_any_device if any_hit else sc.trace_closest_device
    torch.cuda.synchronize()
    for _ in range(3): f(d_rays.data_ptr(), n, d_out.data_ptr())
    stream.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(stream); f(d_rays.data_ptr(), n, d_out.data_ptr()); b.record(stream); b.synchronize()
        ts.append(a.elapsed_time(b))
    sc.set_stream(None)
    return float(np.median(ts)), float(np.min(ts))

for scene in sys.argv[1:] or ["c1_cornell", "c2_explosion"]:
    p = helpers.fixture_path(scene, full=True) or helpers.fixture_path(scene)
    fx = helpers.Fixture(p)
    t0 = time.time(); sc = fx.scene(); t1 = time.time(); sc.attach(0); t2 = time.time()
    print(scene, "host preCalc %.2fs attach %.2fs" % (t1 - t0, t2 - t1), sc.bvh_stats())
    hits = sc.trace_closest(fx.rays)
    st = helpers.compare_hits(sc, hits, fx.hits)
    print("  parity vs reference:", {k: v for k, v in st.items() if k != "hard_idx"})
    sc.enable_counting(True); sc.reset_counters(); sc.trace_closest(fx.rays); c = sc.counters(); sc.enable_counting(False)
    n = len(fx.rays)
    print("  per ray: nodes %.2f tris %.2f" % (c["nodes_fetched"] / n, c["tris_tested"] / n))
    bytes_per_ray = c["nodes_fetched"] / n * 128 + c["tris_tested"] / n * 48 + 48 + 20
    med, best = time_trace(sc, fx.rays)
    print("  closest: %.3f ms (best %.3f) -> %.1f Mrays/s, %.1f GB/s algorithmic (%.0f B/ray)" % (med, best, n / med * 1e-3, n * bytes_per_ray / med * 1e-6, bytes_per_ray))
    med, best = time_trace(sc, fx.rays, any_hit=True)
    print("  any:     %.3f ms (best %.3f) -> %.1f Mrays/s" % (med, best, n / med * 1e-3))
    # sub-batches: primary (coherent) vs incoherent part
    npix = int(fx.z["image_shape"][0]) * int(fx.z["image_shape"][1])
    if 0 < npix < n:
        for label, r in (("primary", fx.rays[:npix]), ("incoherent", fx.rays[npix:])):
            med, best = time_trace(sc, r)
            print("  %s closest: %.3f ms -> %.1f Mrays/s" % (label, med, len(r) / med * 1e-3))
    t0 = time.time(); h = sc.trace_closest(fx.rays); t1 = time.time()
    print("  e2e host->host: %.1f ms -> %.1f Mrays/s" % ((t1 - t0) * 1e3, n / (t1 - t0) * 1e-6))
    sc.close()
