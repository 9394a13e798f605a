#!/usr/bin/env python
"""bench.py — Mrays/s of the ray-casting hot path on B200 (BASELINE.json metric), with roofline, in-run parity against the
reference and the reference's CPU path timed beside it.

Headline workload (`value`, `e2e`, `roofline`): BASELINE config C2 on stand-in geometry (bunny / dragon_2 are missing from the
reference mount): Models/Final/explosion01.obj, 86 914 triangles, 1920x1080.  One STEP = one pass of the hot path over
    (i)   2 073 600 coherent primary rays at the pixel centres      -> closest hit  (Scene::trace)
    (ii)  2 073 600 seeded incoherent rays                          -> closest hit
    (iii) 2 073 600 PointLight shadow rays from the hits of (ii)    -> any hit
A "ray" is one Scene::trace query.  `value` is device throughput with the ray buffers resident in HBM; `e2e` is the same step
through the host-pointer C ABI (miro_gpu_trace_closest / _any) from PINNED host buffers, H2D + kernels + D2H inside the
timed region.  The same three batches are also measured (`workloads`) on the scenes the metric names at THEIR scale — `big`
(20 placed copies of explosion01 = 1.74 M triangles, the dragon-scale stand-in) and `c5` (motion-blur bullets + 40 401
ProxyObject instances) —, each with its own roofline, device counters, in-run parity against the reference's hit records for a
1/8 sample of the rays, and the reference's Mrays/s on that sample (tests/bench_workloads.py).  `render` holds whole frames
(Scene::raytraceImage through miro_host_raytrace_image, host frame out) for C1..C5 with the reference's own render timed
beside each.  Under torchrun each rank traces its own batches (weak scaling, no data-path collective; the scene is
replicated) and `render_scaling` holds the strong-scaling record of one tile-sharded frame + ONE NCCL reduce.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--legs c2,big,c5,render] [--no-cpu]
"""
import argparse
import ctypes
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reference_arm as ra          # noqa: E402  (numpy + the reference binary; never imports the product)
import bench_workloads as bw        # noqa: E402

N_BATCH = bw.N_BATCH
NODE_BYTES = 64      # device node (csrc/traverse.cuh DeviceNode)
TRI_BYTES = 48       # bytes a triangle test reads = the device record
MBTRI_BYTES = 96
INST_BYTES = 64
SEED = 0x5EED
L2_BYTES = 126e6


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def l2_copy_bandwidth():
    """Measured on the spot: device-to-device copies of a 32 MB buffer into another one (64 MB in all: L2-resident on the 126 MB
    L2), read + written bytes per second, CUDA events around 200 copies.  The denominator of roofline.frac_l2 — the launch whose
    DRAM traffic is a twelfth of its algorithmic bytes is served by L1 and L2, so the HBM figure alone says little about it."""
    import torch
    n = 8 << 20
    x = torch.ones(n, dtype=torch.float32, device="cuda"); y = torch.empty_like(x)
    for _ in range(20):
        y.copy_(x)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(200):
        y.copy_(x)
    b.record(); torch.cuda.synchronize()
    return 2.0 * x.numel() * 4 * 200 / (a.elapsed_time(b) * 1e-3) * 1e-9


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.rows = []; self.stop_flag = False; self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self, t0=None, t1=None):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        rows = [r for (t, r) in self.rows if (t0 is None or (t0 <= t <= t1)) and len(r) >= 6]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "covers": "the timed region and the sustained leg that follows it (same launches, back to back)"}


# ------------------------------------------------------------------------------------------------ the reference on the host cores
def write_hdr(path, rgbe):
    """Radiance .hdr (new-style RLE scanlines made of literal chunks) from RGBE bytes [h, w, 4] — what src/hdrloader.cpp reads."""
    h, w, _ = rgbe.shape
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w))
        for y in range(h):
            f.write(bytes([2, 2, w >> 8, w & 255]))
            for c in range(4):
                row = rgbe[y, :, c].tobytes()
                for i in range(0, w, 128):
                    chunk = row[i:i + 128]
                    f.write(bytes([len(chunk)])); f.write(chunk)
        f.write(b"\n")


def materialise(fx, script, tmp, workload=None):
    """Scene script + OBJ files (+ .hdr light maps) under tmp, as the reference's own loaders read them."""
    sp = workload.materialise(tmp) if workload is not None else ra.write_obj_scene(fx, tmp, script)
    text = open(sp).read()
    out = []
    for line in text.splitlines():
        tok = line.split()
        if len(tok) >= 3 and tok[0] == "texture" and ("texrgbe_" + tok[1]) in fx.z.files:
            hp = os.path.join(tmp, tok[1] + ".hdr")
            write_hdr(hp, fx.z["texrgbe_" + tok[1]])
            line = "texture %s %s" % (tok[1], hp)
        out.append(line)
    open(sp, "w").write("\n".join(out) + "\n")
    return sp


def reference_workload(w, threads, repeat, warmup, want_meshes):
    """The reference's own Scene::trace over the bounded sample of workload w (oracle/_ref/miro_ref: ONE process, scene load and
    BVH::build outside the timed region, `warmup` untimed and `repeat` timed passes on `threads` OpenMP threads; the hit
    records come from a single-threaded pass — see ra.run_reference)."""
    prim, inco = w.primary(), w.incoherent(SEED)
    p, q = w.sample(prim, inco)
    rays = np.concatenate([p, q])
    ex = {}
    with tempfile.TemporaryDirectory() as tmp:
        materialise(w.fx, w.script, tmp, workload=w)
        res = ra.run_reference(w.fx, rays, threads=threads, repeat=max(repeat, 1), warmup=max(warmup, 0), want_hits=True, scene_dir=tmp,
                               shadow=(w.light, len(p), len(q)), extras=ex, dump_meshes=want_meshes)
    ev = res[0]
    tr = [e for e in ev if e.get("event") == "trace"][0]; sh = [e for e in ev if e.get("event") == "shadow"][0]
    sc = [e for e in ev if e.get("event") == "scene"][0]
    n = len(rays) + len(q)
    return {"n": n, "n_primary": len(p), "n_incoherent": len(q), "rays": rays, "hits": res[1], "meshes": res[2] if want_meshes else None,
            "shadow_rays": ex["shadow_rays"], "shadow_hits": ex["shadow_hits"],
            "seconds_mean": tr["mean_seconds"] + sh["mean_seconds"], "seconds_best": tr["seconds"] + sh["seconds"], "threads": tr["threads"],
            "build_s": sc.get("build_s"), "qbvh_nodes": sc.get("qbvh_nodes"),
            "sample": "1/8 of each of the 3 ray batches of one step (%d rays): every 2nd row x 4th column of the primary batch, every 8th "
                      "incoherent ray, the shadow rays the reference casts from those hits" % n}


def reference_main(args, rank):
    """--impl reference: the UNMODIFIED reference (oracle/_ref/miro_ref) on the host cores; no product code is imported."""
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 16)          # Ray::counter[128 * tid] caps the reference at 16 threads (src/Ray.h:30,74)
    legs = [l for l in (args.legs or "c2,big,c5").split(",") if l in ("c2", "big", "c5")]
    if "c2" not in legs:
        legs.insert(0, "c2")
    if not ra.have_reference():
        emit({"impl": "reference", "unavailable": "oracle/_ref/miro_ref is not built (run __graft_entry__.build() where /root/reference exists)"})
        return
    rows = {}
    for name in legs:
        w = bw.Workload(name).load()
        k, wu = (max(args.steps, 1), max(args.warmup, 0)) if name == "c2" else (min(max(args.steps, 1), 3), min(max(args.warmup, 0), 1))
        r = reference_workload(w, threads, k, wu, want_meshes=False)
        rows[name] = {"workload": w.label, "Mrays_per_s": r["n"] / r["seconds_mean"] * 1e-6, "ms_per_step": r["seconds_mean"] * 1e3, "rays_per_step": r["n"],
                      "cores": r["threads"], "steps": k, "warmup": wu, "scene_load_and_bvh_build_s": r["build_s"], "qbvh_nodes": r["qbvh_nodes"], "sample": r["sample"]}
    c2 = rows["c2"]
    val = c2["Mrays_per_s"]
    line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": c2["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(bw.Workload("c2").load()),
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": c2["cores"], "kind": "reference", "sample": c2["sample"]},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "workloads": rows}
    emit(line)


def bench_config(w):
    return {"workload": w.label + ": 2073600 primary + 2073600 incoherent closest-hit + 2073600 shadow any-hit rays per step",
            "scene": w.fixture, "triangles": w.triangles(), "rays_per_step_per_gpu": 3 * N_BATCH,
            "l2_policy": "inputs_larger_than_L2 (298 MB of rays per step; the BVH of the headline workload is L2-resident by nature of the "
                         "workload — see roofline.bound and workloads.big for a structure of L2 size)",
            "sharding": "rays per rank, scene replicated, no collective",
            "launch_chaining": "the 3 traversal launches of a step are chained with programmatic dependent launch (miro_gpu_set_trace_chaining) in "
                               "the timed region; per-launch times come from a separate, unchained pass"}


# ------------------------------------------------------------------------------------------------ the GPU arm
def product_scene(w, meshes, device):
    """The product's scene for workload w: the geometry the reference holds after ITS load (ref dump) when there is one, else
    the fixture's meshes (transformed by the script's ctm in numpy for scripts that place copies)."""
    import miro_b200 as mb
    sc = mb.MiroScene()
    names = w.mesh_names()
    if meshes is not None:
        for n in names:
            sc.preload_mesh(n, **meshes[n])
    else:
        own = {n: w.fx.mesh(k) for k, n in enumerate(w.fx.names)}
        for line in w.script.splitlines():
            t = line.split()
            if t[:1] != ["mesh"]:
                continue
            m = dict(own[t[1]] if t[1] in own else own[t[2].lstrip("@")])
            if "ctm" in t:
                M = np.array([float(x) for x in t[t.index("ctm") + 1:t.index("ctm") + 17]], np.float32).reshape(4, 4)
                m["vertices"] = (m["vertices"] @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
            sc.preload_mesh(t[1], **m)
    with tempfile.NamedTemporaryFile("w", suffix=".miro", delete=False) as f:
        f.write(w.script.replace("@", "")); path = f.name
    try:
        sc.load_script(path, "/nonexistent-asset-root")
    finally:
        os.unlink(path)
    return sc.attach(device)


def algorithmic_bytes(c, n_rays, closest, mb_share):
    """N_node x node + N_tri x triangle + N_inst x instance + ray in + hit out, with the byte sizes of the SHIPPED
    layout (64 B quantized device node, 48 B read per triangle test / 96 B per motion-blur test, 64 B instance record,
    48 B ray, 20 B hit | 1 bit)."""
    tri_bytes = c["tris_tested"] * (TRI_BYTES + (MBTRI_BYTES - TRI_BYTES) * mb_share)
    out = 20 * n_rays if closest else 4 * ((n_rays + 31) // 32)
    return c["nodes_fetched"] * NODE_BYTES + tri_bytes + c["insts_entered"] * INST_BYTES + 48 * n_rays + out


def trace_leg(w, sc, args, rank, world, local, barrier, sampler=None, sustain_s=0.0):
    """Device-resident throughput of the three batches of workload w + per-launch roofline numbers."""
    import torch
    import miro_b200 as mb
    import torch.distributed as dist
    prim = w.primary()
    inco = w.incoherent(SEED + rank)
    inco_hits = sc.trace_closest(inco)
    shad = w.shadow(inco, inco_hits["t"], inco_hits["prim"] >= 0)
    batches = [prim, inco, shad]
    stream = torch.cuda.Stream()
    sc.set_stream(stream.cuda_stream)
    d_rays = [torch.from_numpy(b.view(np.uint8).reshape(len(b), -1)).cuda() for b in batches]
    d_hits = [torch.empty((N_BATCH, 20), dtype=torch.uint8, device="cuda") for _ in range(2)]
    d_bits = torch.empty(((N_BATCH + 31) // 32,), dtype=torch.int32, device="cuda")
    calls = [(sc.trace_closest_device, d_rays[0], d_hits[0]), (sc.trace_closest_device, d_rays[1], d_hits[1]), (sc.trace_any_device, d_rays[2], d_bits)]

    def step():
        for f, r, o in calls:
            f(r.data_ptr(), N_BATCH, o.data_ptr())

    d = sc.desc()
    mb_share = 0.0       # share of motion-blur tests among triangle tests is not counted separately: bound it by the scene's share of MB triangles
    if d.n_mbtris:
        mb_share = d.n_mbtris / float(d.n_mbtris + d.n_tris)
    sc.enable_counting(True)
    per_launch = []
    for i, (f, r, o) in enumerate(calls):
        sc.reset_counters(); f(r.data_ptr(), N_BATCH, o.data_ptr()); c = sc.counters()
        per_launch.append({"nodes": c["nodes_fetched"], "tris": c["tris_tested"], "insts": c["insts_entered"], "bytes": algorithmic_bytes(c, N_BATCH, i < 2, mb_share)})
    sc.enable_counting(False)

    sc.set_trace_chaining(True)          # all ray buffers of the step were complete long before: the contract holds
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # timed region: K steps back to back, bracketed by two events on the launching stream (no events between the launches of a
    # step: consecutive traversal launches are chained by programmatic dependent launch, which an event record would break)
    e_first, e_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e_first.record(stream)
    for _ in range(args.steps):
        step()
    e_last.record(stream)
    stream.synchronize()
    barrier()
    total_ms = e_first.elapsed_time(e_last)
    sustained = None
    if sustain_s > 0:
        # the same launches for >= sustain_s seconds, so that the 20 ms clock sampler sees the GPU under this load
        n_sus = max(args.steps, int(sustain_s * 1e3 / max(total_ms / args.steps, 1e-3)) + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n_sus):
            step()
        e1.record(stream); stream.synchronize()
        sus_ms = e0.elapsed_time(e1)
        sustained = {"value": 3 * N_BATCH * n_sus / sus_ms * 1e-3, "unit": "Mrays/s (this rank)", "steps": n_sus, "seconds": sus_ms * 1e-3}
    t_wall1 = time.time()
    sc.set_trace_chaining(False)
    # the chained launches must have produced what the unchained host-pointer call produced
    chk = d_hits[1].cpu().numpy().view(mb.HIT_DTYPE).reshape(-1)
    assert np.array_equal(chk["prim"], inco_hits["prim"]) and np.array_equal(chk["t"], inco_hits["t"]), "chained launches changed the hits"
    # per-launch durations (roofline of the dominant kernel): a second pass with an event after every launch
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for k in range(args.steps):
        ev[k][0].record(stream)
        for i, (f, r, o) in enumerate(calls):
            f(r.data_ptr(), N_BATCH, o.data_ptr()); ev[k][i + 1].record(stream)
    stream.synchronize()
    launch_ms = [float(np.mean([ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(args.steps)])) for i in range(3)]
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = 3 * N_BATCH * args.steps * world / total_ms * 1e-3
    sc.set_stream(None)
    return {"value": value, "total_ms": total_ms, "launch_ms": launch_ms, "per_launch": per_launch, "batches": batches, "inco_hits": inco_hits,
            "sustained": sustained, "wall": (t_wall0, t_wall1), "structure_bytes": int(d.n_nodes) * NODE_BYTES + int(d.n_tris) * TRI_BYTES + int(d.n_mbtris) * 96 + int(d.n_instances) * 64,
            "n_nodes": int(d.n_nodes), "n_tris_device": int(d.n_tris), "n_instances_device": int(d.n_instances)}


KERNEL_NAMES = ["k_trace<closest> primary", "k_trace<closest> incoherent", "k_trace<any> shadow"]


_L2_PEAK = []


def roofline_of(leg, hbm_peak, peak_src, traffic_key=None):
    if not _L2_PEAK:
        _L2_PEAK.append(l2_copy_bandwidth())
    launch_ms, per_launch = leg["launch_ms"], leg["per_launch"]
    dom = int(np.argmax(launch_ms))
    achieved = per_launch[dom]["bytes"] / (launch_ms[dom] * 1e-3) * 1e-9
    traffic = l2_bytes = l1_bytes = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if traffic_key and os.path.exists(tp):      # per launch, from the committed ncu --set full capture: DRAM read+write, L2 sectors, L1 global-load sectors
        tj = json.load(open(tp)).get(traffic_key, {})
        traffic = tj.get("per_launch_dram_bytes", {}).get(KERNEL_NAMES[dom])
        l2_bytes = tj.get("per_launch_l2_bytes", {}).get(KERNEL_NAMES[dom])
        l1_bytes = tj.get("per_launch_l1_global_load_bytes", {}).get(KERNEL_NAMES[dom])
    alg = per_launch[dom]["bytes"]
    # where the algorithmic bytes are served: a structure that fits L2 leaves DRAM only the ray / hit streams — the SURVEY 8(d)
    # fraction against HBM bandwidth is then an L1/L2-served figure, and the kernel is bound by issue slots and L1 wavefronts
    if traffic is not None:
        bound = "hbm" if traffic >= 0.5 * alg else "l2/issue (ncu DRAM traffic is %.2f x the algorithmic bytes: L1 and L2 serve the rest; %s)" % (
            traffic / alg, "the structure fits L2" if leg["structure_bytes"] <= L2_BYTES else "the structure exceeds L2, the launch waits on the latency of its L2 misses, not on bandwidth")
    else:
        bound = "hbm" if leg["structure_bytes"] > L2_BYTES else "l2/issue (structure of %.1f MB fits the 126 MB L2)" % (leg["structure_bytes"] * 1e-6)
    return {"bound": bound, "kernel": KERNEL_NAMES[dom], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "peak_source": peak_src, "traffic": traffic, "l2_traffic": l2_bytes, "l1_global_load_traffic": l1_bytes,
            "l2_copy_bandwidth_measured_GBps": _L2_PEAK[0],
            "frac_l2": (l2_bytes / (launch_ms[dom] * 1e-3) * 1e-9 / _L2_PEAK[0]) if l2_bytes else None,
            "frac_l2_note": "L2 sector traffic of the launch (committed ncu capture, lts__t_sectors x 32 B) / its live duration / the L2-resident copy "
                            "bandwidth measured in this run: how much of the L2 the launch uses — it is bound by issue slots and dependent-fetch "
                            "latency (profiles/r2_ncu_summary.md), not by L2 bandwidth either",
            "algorithmic_bytes_per_launch": alg, "structure_bytes_on_device": leg["structure_bytes"],
            "nodes_per_ray": per_launch[dom]["nodes"] / N_BATCH, "tris_per_ray": per_launch[dom]["tris"] / N_BATCH, "insts_per_ray": per_launch[dom]["insts"] / N_BATCH,
            "launch_ms": launch_ms[dom], "share_of_step": launch_ms[dom] / sum(launch_ms),
            "all_launches": [{"kernel": KERNEL_NAMES[i], "ms": launch_ms[i], "Mrays_per_s": N_BATCH / launch_ms[i] * 1e-3,
                              "GBps": per_launch[i]["bytes"] / launch_ms[i] * 1e-6, "bytes_per_ray": per_launch[i]["bytes"] / N_BATCH,
                              "nodes_per_ray": per_launch[i]["nodes"] / N_BATCH, "tris_per_ray": per_launch[i]["tris"] / N_BATCH,
                              "insts_per_ray": per_launch[i]["insts"] / N_BATCH} for i in range(3)]}


def parity_of(w, sc, ref):
    """The GPU's answers for the reference's sample against the reference's own hit records (single-threaded pass of the
    unmodified binary): hit identity, ties split by cause, distances; occlusion of the shadow rays the reference cast."""
    hits = sc.trace_closest(ref["rays"])
    mesh, tri, proxy = sc.resolve_hits(hits)
    geom = ra.SceneGeometry(w.script, ref["meshes"])
    out = {}
    parts = (("primary", slice(0, ref["n_primary"])), ("incoherent", slice(ref["n_primary"], None)))
    for name, sl in parts:
        st = ra.compare_with_reference(mesh[sl], tri[sl], proxy[sl], hits["t"][sl], hits["a"][sl], hits["b"][sl], ref["hits"][sl], rays=ref["rays"][sl], t_rel=1e-5)
        # what the barycentric heuristic cannot class as a tie is re-computed in float64 on both sides' triangles
        adj = ra.adjudicate_mismatches(geom, ref["rays"][sl], st["hard_idx"], mesh[sl], tri[sl], proxy[sl], ref["hits"][sl])
        out[name] = {"n": st["n"], "id_match": st["id_match"], "ties": st["ties"], "ties_equal_t": st["ties_equal_t"], "ties_own_edge": st["ties_own_edge"],
                     "ties_ref_edge": st["ties_ref_edge"], "unclassed": st["hard"],
                     "unclassed_in_float64": {k: v for k, v in adj.items() if k != "hard_idx"},
                     "hard": adj["product_missed"] + adj["unexplained"], "max_rel_t": st["max_rel_t"], "frac_t_within_1e-5": st["frac_t_within"]}
    occ = sc.trace_any(ref["shadow_rays"])
    r_occ = ref["shadow_hits"]["mesh"] >= 0
    dis = occ != r_occ
    # a disagreement is a tie when the reference's blocker sits at the very start or end of the interval or the ray grazes an edge of it
    sh = ref["shadow_hits"]; sr = ref["shadow_rays"]
    wmin = np.minimum(np.minimum(sh["a"], sh["b"]), 1.0 - sh["a"] - sh["b"])
    graze = r_occ & ((wmin <= 1e-4) | (sh["t"] <= sr["tmin"] * (1 + 1e-3)) | (sh["t"] >= sr["tmax"] * (1 - 1e-5)))
    closest_of_shadow = sc.trace_closest(sr)
    gw = np.minimum(np.minimum(closest_of_shadow["a"], closest_of_shadow["b"]), 1.0 - closest_of_shadow["a"] - closest_of_shadow["b"])
    own_graze = (closest_of_shadow["prim"] >= 0) & (gw <= 1e-4)
    out["shadow"] = {"n": int(len(occ)), "any_hit_agree": float((~dis).mean()), "disagree": int(dis.sum()),
                     "disagree_ties": int((dis & (graze | own_graze)).sum()), "hard": int((dis & ~(graze | own_graze)).sum()),
                     "any_hit_equals_own_closest_hit": bool(np.array_equal(occ, closest_of_shadow["prim"] >= 0))}
    out["hard_total"] = out["primary"]["hard"] + out["incoherent"]["hard"] + out["shadow"]["hard"]
    out["against"] = "hit records of oracle/_ref/miro_ref (the unmodified reference), single-threaded pass, same rays, same geometry as its loader left it"
    return out


def e2e_legs(sc, leg, args, world, barrier):
    """The step through the host-pointer ABI: H2D of the rays, kernels, D2H of hits / occlusion bits inside the timed region."""
    import torch
    import torch.distributed as dist
    import miro_b200 as mb
    batches, inco_hits = leg["batches"], leg["inco_hits"]
    L = sc.L
    out_hits = [torch.empty((N_BATCH, 20), dtype=torch.uint8).pin_memory() for _ in range(2)]
    out_bits = torch.empty(((N_BATCH + 31) // 32,), dtype=torch.int32).pin_memory()
    d2h = 2 * N_BATCH * 20 + 4 * ((N_BATCH + 31) // 32)
    e2e_steps = max(3, min(args.steps, 40))

    def run(fn_c, fn_a, bufs, hits_out, bits_out, ray_bytes, steps):
        def one():
            fn_c(sc.ctx, bufs[0], N_BATCH, hits_out[0]); fn_c(sc.ctx, bufs[1], N_BATCH, hits_out[1]); fn_a(sc.ctx, bufs[2], N_BATCH, bits_out)
        for _ in range(2):
            one()
        barrier()
        t0 = time.time()
        for _ in range(steps):
            one()
        torch.cuda.synchronize()
        s = time.time() - t0
        s_rank = s
        if world > 1:
            t = torch.tensor([s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s = float(t.item())
        return {"value": 3 * N_BATCH * steps * world / s * 1e-6, "unit": "Mrays/s", "h2d_bytes_per_step": 3 * N_BATCH * ray_bytes, "d2h_bytes_per_step": d2h, "steps": steps,
                "this_rank_h2d_GBps": 3 * N_BATCH * ray_bytes * steps / s_rank * 1e-9, "this_rank_d2h_GBps": d2h * steps / s_rank * 1e-9}

    hp = [h.data_ptr() for h in out_hits]
    pinned = [torch.from_numpy(b.view(np.uint8).reshape(len(b), -1).copy()).pin_memory() for b in batches]
    e2e = run(L.miro_gpu_trace_closest, L.miro_gpu_trace_any, [p.data_ptr() for p in pinned], hp, out_bits.data_ptr(), 48, e2e_steps)
    e2e["ray_format"] = "miro_gpu_ray, 48 B (the general ABI record), pinned host buffers"
    got = out_hits[1].numpy().view(mb.HIT_DTYPE).reshape(-1)
    assert np.array_equal(got["prim"], inco_hits["prim"]) and np.array_equal(got["t"], inco_hits["t"])
    # 32-byte packed rays (static scenes: time / flags / user words carry nothing): a third less over PCIe, which bounds these calls
    pinned32 = [torch.from_numpy(mb.pack_rays(b).view(np.uint8).reshape(len(b), -1).copy()).pin_memory() for b in batches]
    packed = run(L.miro_gpu_trace_closest_packed, L.miro_gpu_trace_any_packed, [p.data_ptr() for p in pinned32], hp, out_bits.data_ptr(), 32, e2e_steps)
    packed["ray_format"] = "miro_gpu_ray32, 32 B (static scenes: no time / flags words), pinned host buffers; identical hits asserted"
    got = out_hits[1].numpy().view(mb.HIT_DTYPE).reshape(-1)
    assert np.array_equal(got["prim"], inco_hits["prim"]) and np.array_equal(got["t"], inco_hits["t"])
    # pageable memory, as an unmodified Miro caller would hand it over (plain malloc'ed arrays): the library stages through its own pinned ring
    plain = [np.ascontiguousarray(b) for b in batches]
    ph = [np.empty(N_BATCH, mb.HIT_DTYPE) for _ in range(2)]; pb = np.empty((N_BATCH + 31) // 32, np.uint32)
    pageable = run(L.miro_gpu_trace_closest, L.miro_gpu_trace_any, [p.ctypes.data for p in plain], [h.ctypes.data for h in ph], pb.ctypes.data, 48, max(3, e2e_steps // 2))
    pageable["ray_format"] = "miro_gpu_ray, 48 B, PAGEABLE host buffers (what an unmodified Miro caller holds)"
    assert np.array_equal(ph[1]["prim"], inco_hits["prim"])
    # the same ordinary arrays page-locked in place, once, by the caller (miro_gpu_pin_host_buffer): what the patched Miro does with
    # the ray / hit arrays it reuses from frame to frame
    t0 = time.time()
    for a in plain + ph + [pb]:
        assert L.miro_gpu_pin_host_buffer(sc.ctx, a.ctypes.data, a.nbytes) == 0
    pin_ms = (time.time() - t0) * 1e3
    registered = run(L.miro_gpu_trace_closest, L.miro_gpu_trace_any, [p.ctypes.data for p in plain], [h.ctypes.data for h in ph], pb.ctypes.data, 48, e2e_steps)
    registered["ray_format"] = "miro_gpu_ray, 48 B, the caller's own (malloc'ed) arrays pinned in place once with miro_gpu_pin_host_buffer"
    registered["pin_once_ms"] = pin_ms
    assert np.array_equal(ph[1]["prim"], inco_hits["prim"])
    for a in plain + ph + [pb]:
        L.miro_gpu_unpin_host_buffer(sc.ctx, a.ctypes.data)
    # camera rays made on the device (miro_gpu_trace_primary): a renderer's primary batch needs no upload at all; the other two
    # batches as packed rays
    cam = sc.camera()

    def prim_call(ctx, _rays, _n, out):
        return L.miro_gpu_trace_primary(ctx, ctypes.byref(cam), bw.WIDTH, bw.HEIGHT, 0, out, None)

    def one_dc():
        prim_call(sc.ctx, None, N_BATCH, hp[0]); L.miro_gpu_trace_closest_packed(sc.ctx, pinned32[1].data_ptr(), N_BATCH, hp[1]); L.miro_gpu_trace_any_packed(sc.ctx, pinned32[2].data_ptr(), N_BATCH, out_bits.data_ptr())
    for _ in range(2):
        one_dc()
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        one_dc()
    torch.cuda.synchronize()
    s_dc = time.time() - t0
    if world > 1:
        t = torch.tensor([s_dc], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s_dc = float(t.item())
    devcam = {"value": 3 * N_BATCH * e2e_steps * world / s_dc * 1e-6, "unit": "Mrays/s", "h2d_bytes_per_step": 2 * N_BATCH * 32, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
              "ray_format": "primary batch: camera rays generated on the device (miro_gpu_trace_primary, hits only travel); incoherent and shadow batches: 32 B packed rays, pinned"}
    return e2e, packed, pageable, devcam, registered


def pcie_probe(world, barrier):
    """What the host gives this rank when ALL ranks copy at once: plain pinned H2D and D2H of 256 MB, 5 rounds, all ranks between
    the same barriers — the ceiling of the host-pointer calls (they move 48 + 13 bytes per ray), measured rather than assumed."""
    import torch
    import torch.distributed as dist
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    out = {}
    for name, (src, dst) in (("h2d", (h, d)), ("d2h", (d, h))):
        dst.copy_(src, non_blocking=True); barrier()
        t0 = time.time()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        s = time.time() - t0
        if world > 1:
            t = torch.tensor([s], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); s = float(t.item())
        out[name + "_GBps_per_rank_all_ranks_concurrent"] = 5 * n / s * 1e-9
        barrier()
    out["e2e_ceiling_Mrays_per_s"] = world * 1e3 / (48.0 / out["h2d_GBps_per_rank_all_ranks_concurrent"])      # the upload alone: 48 B per ray
    return out


def render_legs(local, with_reference):
    """Whole frames — Scene::raytraceImage through miro_host_raytrace_image (float frame on the host at the end) — for the
    BASELINE configs, next to the reference's own render loop (miro_ref --render-float: adaptiveSampleScene over the
    reference's bucket order on all host threads).  rays = Scene::trace queries, counted on both sides."""
    import helpers
    rows = []
    threads = min(os.cpu_count() or 1, 16)
    configs = [("C1", "c1_cornell", None, "Cornell box, Lambert + PointLight, 512x512, 1 spp primary + shadow"),
               ("C2", "c2_explosion", None, "explosion01.obj (stand-in), Lambert + PointLight, 1920x1080, 1 spp primary + shadow"),
               ("C3", "c3_dome_pt", (512, 512), "teapot + floor (stand-in), Blinn path tracing, DomeLight (Arches_E_PineTree.hdr) importance sampling, 64 paths, 512x512"),
               ("C4", "c4_cornell_pt", (512, 512), "Cornell box (Sponza stand-in), Blinn, RectangleLight x4 soft-shadow samples + emitter, 16 paths, 4 indirect segments, 512x512"),
               ("C5", "c5_mb_instances", (512, 512), "motion-blur bullets + 961 ProxyObject instances of testGrass.obj, 2 subdivision levels (5 camera samples), 512x512"),
               ("C5 at makeProxyGrid scale", "c5_mb_instances", (512, 512), "motion-blur bullets + 201 x 201 = 40401 ProxyObject instances (src/main.cpp:37-52), 2 subdivision levels (5 camera samples), 512x512")]
    for tag, name, size, label in configs:
        path = helpers.fixture_path(name)
        if path is None:
            continue
        fx = helpers.Fixture(path)
        base_script = bw.Workload("c5").script if tag == "C5 at makeProxyGrid scale" else fx.script
        script = base_script if size is None else re.sub(r"image \d+ \d+", "image %d %d" % size, base_script)
        row = {"config": tag, "workload": label}
        ref_runs = {}
        if with_reference and ra.have_reference():
            with tempfile.TemporaryDirectory() as tmp:
                materialise(fx, script, tmp)
                for th in (threads, 1):
                    # all host threads first; ONE thread as well when that is affordable: with several threads the reference's
                    # traversals corrupt each other (QBVH_Node::boxHit) — on small trees most of them, and
                    # paths that wrongly miss end early, so the multi-threaded frame is faster than a correct one and wrong
                    if th == 1 and (threads == 1 or ref_runs[threads].get("ms", 1e9) * threads > 60e3):
                        continue
                    try:
                        ev, _ = ra.run_reference(fx, None, threads=th, scene_dir=tmp, extra_args=["--render-float", os.path.join(tmp, "out.f32")])
                        rf = [e for e in ev if e.get("event") == "render_float"][0]
                        ref_runs[th] = {"rays": rf["rays"], "ms": rf["seconds"] * 1e3, "Mrays_per_s": rf["mrays_per_s"], "cores": rf["threads"]}
                    except Exception as e:      # a reference failure must not cost the GPU numbers
                        ref_runs[th] = {"error": str(e)[-300:]}
            row["reference"] = ref_runs[threads]
            if 1 in ref_runs:
                row["reference_one_thread"] = ref_runs[1]
        sc = fx.scene(script_override=script).attach(local)
        sc.render_in_place()
        best = 1e30
        for _ in range(5):      # the frame as Scene::raytraceImage leaves it: in the scene's Image (float radiance + 8-bit pixels, on the host)
            sc.reset_counters()
            t0 = time.time(); img, _img8 = sc.render_in_place(); dt = time.time() - t0
            best = min(best, dt)
        c = sc.counters(); rays = int(c["rays_closest"] + c["rays_any"])
        row["gpu"] = {"rays": rays, "ms": best * 1e3, "Mrays_per_s": rays / best * 1e-6, "kernel_launches": int(c["kernel_launches"]), "frame_mean": float(np.minimum(img, 4).mean())}
        if "reference" in row and "ms" in row["reference"]:
            row["frame_speedup"] = row["reference"]["ms"] / row["gpu"]["ms"]
            row["reference"]["rays_vs_gpu"] = row["reference"]["rays"] / max(rays, 1)
        if "reference_one_thread" in row and "ms" in row["reference_one_thread"]:
            row["frame_speedup_vs_one_correct_thread"] = row["reference_one_thread"]["ms"] / row["gpu"]["ms"]
            row["reference_one_thread"]["rays_vs_gpu"] = row["reference_one_thread"]["rays"] / max(rays, 1)
        sc.close()
        rows.append(row)
    return rows


def render_scaling_leg(rank, world, local, barrier):
    """Strong scaling of ONE frame (north_star: tile x sample sharding, accumulation buffers combined by an NCCL reduce): the C4
    stand-in and the C3 stand-in at 2048x2048; every rank renders its 32x32 buckets (bucket order of src/Scene.cpp:160-175) into
    its own frame, then ONE ncclReduce(SUM) to rank 0.  Frame time = max over ranks, the collective included."""
    import torch
    import torch.distributed as dist
    import helpers
    rows = []
    for tag, name in (("C4", "c4_cornell_pt"), ("C3", "c3_dome_pt")):
        path = helpers.fixture_path(name)
        if path is None:
            continue
        fx = helpers.Fixture(path)
        script = re.sub(r"image \d+ \d+", "image 2048 2048", fx.script)
        sc = fx.scene(script_override=script).attach(local)
        p = sc.render_params(); cam = sc.camera()
        p.shard_index, p.shard_count = rank, world
        frame = torch.zeros((p.height, p.width, 3), dtype=torch.float32, device="cuda")

        def one():
            frame.zero_()
            sc.render_device(frame.data_ptr(), params=p, camera=cam)
            if world > 1:
                dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
        one()
        best, rays = 1e30, 0
        for _ in range(3):
            sc.reset_counters()
            barrier()
            t0 = time.time(); one(); dt = time.time() - t0
            c = sc.counters()
            t = torch.tensor([dt, float(c["rays_closest"] + c["rays_any"])], dtype=torch.float64, device="cuda")
            if world > 1:
                tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX); dist.all_reduce(t, op=dist.ReduceOp.SUM)
                dt = float(tm[0].item())
            if dt < best:
                best, rays = dt, int(t[1].item())
        rows.append({"config": tag, "scene": name, "size": [p.width, p.height], "num_paths": p.num_paths, "max_bounces": p.max_bounces, "n_gpus": world,
                     "sharding": "32x32 buckets round-robin over ranks", "collective": "one NCCL reduce(SUM) of the %d MB frame to rank 0" % (p.width * p.height * 12 // 1000000),
                     "rays": rays, "ms": best * 1e3, "Mrays_per_s": rays / best * 1e-6, "frame_mean": float(frame.mean().item()) if rank == 0 else None, "scaling": "strong"})
        sc.close()
    return rows


def bind_to_gpu_numa_node(device):
    """Pin this rank to the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE any pinned host
    buffer is allocated, so the e2e leg's H2D / D2H copies do not cross the socket interconnect.  Best effort."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        node = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"gpu_numa_node": int(node), "cpus_bound": len(cpus)}
    except Exception:
        return {"gpu_numa_node": None, "cpus_bound": 0}


def gpu_main(args, rank, world, local):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU baseline)")
    import __graft_entry__ as g
    g.ensure_built()
    import helpers
    import miro_b200 as mb      # noqa: F401
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL's version banner goes to stdout, which carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    default_legs = "c2,big,c5,render" if world == 1 else "c2,render_scale"
    legs = (args.legs or default_legs).split(",")      # without c2 (profiling runs) there is no headline line: only `workloads` is printed
    do_cpu = world == 1 and not args.no_cpu
    threads = min(os.cpu_count() or 1, 16)
    hbm_peak, peak_src = peaks()
    line = None
    workloads = {}
    for name in [l for l in legs if l in ("c2", "big", "c5")]:
        w = bw.Workload(name).load(helpers.Fixture)
        ref = None
        if do_cpu and ra.have_reference():
            ref = reference_workload(w, threads, 3, 0, want_meshes=True)      # first: the product traces the geometry the reference holds
        t_build = time.time()
        sc = product_scene(w, ref["meshes"] if ref else None, local)
        t_build = time.time() - t_build
        sampler = None
        if name == "c2":
            sampler = ClockSampler(local); sampler.start()
        leg = trace_leg(w, sc, args, rank, world, local, barrier, sampler, sustain_s=1.5 if name == "c2" else 0.0)
        clocks = sampler.finish(*leg["wall"]) if sampler else None
        roof = roofline_of(leg, hbm_peak, peak_src, traffic_key=name)
        parity = parity_of(w, sc, ref) if ref else None
        row = {"workload": w.label, "traversal_kernel": sc.trace_kernel(), "triangles": w.triangles(), "device_nodes": leg["n_nodes"], "device_triangle_slots": leg["n_tris_device"],
               "device_instances": leg["n_instances_device"], "structure_MB_on_device": leg["structure_bytes"] * 1e-6,
               "host_bvh_build_flatten_and_upload_s": t_build,
               "Mrays_per_s": leg["value"], "ms_per_step": leg["total_ms"] / args.steps, "roofline": roof, "parity": parity}
        if ref:
            row["cpu_baseline"] = {"value": ref["n"] / ref["seconds_best"] * 1e-6, "unit": "Mrays/s", "cores": ref["threads"], "kind": "reference",
                                   "sample": ref["sample"] + ", best of 3", "scene_load_and_bvh_build_s": ref["build_s"]}
        if name == "c2":
            e2e, packed, pageable, devcam, registered = e2e_legs(sc, leg, args, world, barrier)
            numa.update(pcie_probe(world, barrier))
            line = {"metric": "Mrays/s", "value": leg["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                    "ms_per_step": leg["total_ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f32", "data": "synthetic", "config": bench_config(w), "traversal_kernel": sc.trace_kernel(), "clocks": clocks, "gpu_launches": 3 * args.steps,
                    "e2e": e2e, "e2e_packed": packed, "e2e_pageable": pageable, "e2e_registered": registered, "e2e_device_camera": devcam, "host": numa, "sustained": leg["sustained"], "roofline": roof}
            line["roofline"]["frac_abi_node_layout"] = (leg["per_launch"][int(np.argmax(leg["launch_ms"]))]["bytes"] + leg["per_launch"][int(np.argmax(leg["launch_ms"]))]["nodes"] * (128 - NODE_BYTES)) / (max(leg["launch_ms"]) * 1e-3) * 1e-9 / hbm_peak
            line["roofline"]["note"] = ("algorithmic bytes count the SHIPPED 64-byte device node; frac_abi_node_layout counts the 128-byte ABI node as kernel "
                                        "versions <= v4 fetched it (comparable with profiles/bench_r1_v1..v4.json).")
            if parity:
                line["parity"] = parity
            if "cpu_baseline" in row:
                line["cpu_baseline"] = row["cpu_baseline"]
        workloads[name] = row
        sc.close()
        if parity and rank == 0:
            if parity["hard_total"] > 0 or min(parity["primary"]["id_match"], parity["incoherent"]["id_match"]) < (0.9999 if name != "c5" else 0.999):
                raise SystemExit("bench.py: parity against the reference FAILED on workload %s: %s" % (name, json.dumps(parity)))
    if line is None:
        line = {"note": "no headline workload (c2) among the legs: a tooling run"}
    if rank == 0:
        line["workloads"] = {k: v for k, v in workloads.items() if k != "c2"}
    if "render" in legs and world == 1:
        line["render"] = render_legs(local, do_cpu)
    if "render_scale" in legs:
        rs = render_scaling_leg(rank, world, local, barrier)
        if rank == 0:
            line["render_scaling"] = rs
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="miro_gpu", choices=["miro_gpu", "reference"])
    ap.add_argument("--legs", default=None, help="comma list of c2,big,c5,render,render_scale (default: all single-GPU legs; under torchrun c2,render_scale)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the reference legs (kernel tuning runs): no cpu_baseline, no parity")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: whatever a library prints to file descriptor 1 meanwhile (NCCL's version banner, build
    # output) is sent to stderr, and the line is written to the real stdout at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_main(args, rank)
    else:
        gpu_main(args, rank, world, local)


if __name__ == "__main__":
    main()
