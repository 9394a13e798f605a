#!/usr/bin/env python
"""GPU-vs-reference parity table, config by config (run on the GPU box; writes gpurun_out/parity_<tag>.json, copied to
profiles/ by hand).  Traversal configs: hit identities / distances against the reference's hit records (golden fixture
samples, and BASELINE-size batches traced by oracle/_ref/miro_ref on the spot), ties split by cause
(tests/reference_arm.py::compare_with_reference).  Render configs: the GPU image against the reference's radiance
(as estimators where random numbers differ) and against the oracle sample by sample.
usage: tools/parity_report.py [tag]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
import test_trace_gpu as T

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out = {"what": "GPU (libmiro_gpu.so, warp kernel; flat and pool kernels asserted byte-identical; the kernel MIRO_GPU_KERNEL_AUTO picks is named per config) vs the unmodified reference", "trace": {}, "render": {}}


def clean(st):
    return {k: (float(v) if isinstance(v, (np.floating, float)) else int(v)) for k, v in st.items() if k != "hard_idx"}


def trace_stats(fx, sc, rays, ref):
    sc.set_trace_kernel("auto"); auto_kernel = sc.trace_kernel()
    sc.set_trace_kernel("warp"); hits = sc.trace_closest(rays)
    sc.set_trace_kernel("pool"); pool = sc.trace_closest(rays)
    sc.set_trace_kernel("flat"); flat = sc.trace_closest(rays); flat_occ = sc.trace_any(rays)
    sc.set_trace_kernel("warp")
    raw = helpers.compare_hits(sc, hits, ref, t_rel=1e-5, rays=rays)
    adj = helpers.adjudicate_hard(fx, sc, hits, ref, raw, rays)
    st = clean(raw)
    st["unclassed_by_the_barycentric_heuristic"] = st.pop("hard")
    st["unclassed_in_float64"] = {k: v for k, v in adj.items() if k != "hard_idx"}
    st["hard"] = adj["product_missed"] + adj["unexplained"]
    st["pool_kernel_identical"] = bool(hits.tobytes() == pool.tobytes())
    st["flat_kernel_identical"] = bool(hits.tobytes() == flat.tobytes())
    st["default_kernel"] = auto_kernel
    occ = sc.trace_any(rays)
    st["flat_kernel_any_hit_identical"] = bool((occ == flat_occ).all())
    st["any_hit_agrees_with_reference"] = float((occ == (ref["mesh"] >= 0)).mean())
    st["any_hit_agrees_with_own_closest"] = float((occ == (hits["prim"] >= 0)).mean())
    return st


for name in T.SCENES:
    fx = helpers.Fixture(helpers.fixture_path(name))
    sc = fx.scene().attach(0)
    out["trace"][name + " (golden sample)"] = trace_stats(fx, sc, fx.rays, fx.hits)
    sc.close()
    print(name, out["trace"][name + " (golden sample)"], flush=True)
for name in ("c2_explosion", "c5_mb_instances"):
    fx, rays, ref = T.full_size_batch(name)
    sc = fx.scene().attach(0)
    key = name + " (1920x1080 primary + 1 Mi incoherent, reference run on the spot)"
    out["trace"][key] = trace_stats(fx, sc, rays, ref)
    sc.close()
    print(key, out["trace"][key], flush=True)

import test_render_gpu as R
for name in sorted(f[:-4] for f in os.listdir(helpers.GOLDEN) if f.endswith(".npz")):
    fx = helpers.Fixture(helpers.fixture_path(name))
    if fx.radiance is None:
        continue
    sc = fx.scene().attach(0)
    img = sc.render()
    row = {"gpu_mean": float(np.minimum(img, 4).mean()), "reference_mean": float(np.minimum(fx.radiance, 4).mean()),
           "pixels_within_2e-3_of_reference": float(R.pixel_agreement(img, fx.radiance).mean()),
           "rmse_vs_reference_image": R.rmse(img, fx.radiance)}
    if fx.radiance_converged is not None:
        row["rmse_gpu_vs_reference_converged"] = R.rmse(img, fx.radiance_converged)
        row["rmse_reference_equal_spp_vs_reference_converged"] = R.rmse(fx.radiance, fx.radiance_converged)
        row["reference_converged_mean"] = float(np.minimum(fx.radiance_converged, 4).mean())
    p = sc.render_params()
    if p.width * p.height * max(1, p.num_paths) <= (1 << 22):
        oimg, orays = helpers.oracle_render(sc)
        c = sc.counters()
        row["pixels_within_5e-3_of_oracle"] = float(R.pixel_agreement(img, oimg, rel=5e-3, ab=1e-3).mean())
        row["scene_trace_calls_gpu"] = int(c["rays_closest"] + c["rays_any"]); row["scene_trace_calls_oracle"] = int(orays)
    out["render"][name] = row
    sc.close()
    print(name, row, flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_%s.json" % tag), "w"), indent=1)
