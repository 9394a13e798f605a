#!/usr/bin/env python
"""bench.py — Mrays/s of the ray-casting hot path on B200 (BASELINE.json metric), with roofline and CPU baseline.

Workload (BASELINE config C2, stand-in geometry because bunny/dragon are missing from the reference mount):
Models/Final/explosion01.obj (86 914 triangles, geometry as the reference's loader left it, carried by the
fixture files), 1920x1080.  One STEP = one pass of the hot path over one batch of rays:
    (i)   2 073 600 coherent primary rays at pixel centres      -> closest-hit  (Scene::trace)
    (ii)  2 073 600 seeded incoherent rays (origins ~U(AABB), directions ~U(S^2))  -> closest-hit
    (iii) 2 073 600 shadow rays (from the incoherent hits towards the point light)  -> any-hit
A "ray" is one Scene::trace query.  `value` is device throughput with the ray buffers resident in HBM;
`e2e` is the same step through the host-pointer C ABI (miro_gpu_trace_closest / _any) from PINNED host
buffers, H2D + kernels + D2H inside the timed region.  Under torchrun each rank traces its own batch
(weak scaling, no data-path collective; the scene is replicated).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WIDTH, HEIGHT = 1920, 1080
N_BATCH = WIDTH * HEIGHT
NODE_BYTES = 64      # device node (csrc/traverse.cuh DeviceNode)
LIGHT_POS = np.array([-2.0, 4.0, 3.0], np.float32)
SCENE = "c2_explosion"
WORKLOAD = ("C2 stand-in: explosion01.obj 86914 tris, 1920x1080: 2073600 primary + 2073600 incoherent closest-hit "
            "+ 2073600 shadow any-hit rays per step")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def primary_rays(cam, w, h):
    """Pinhole rays at pixel centres (Camera::eyeRayAdaptive with 0.5 offsets, src/Camera.cpp:116-158), row 0 = bottom."""
    from miro_b200 import RAY_DTYPE
    eye = np.array(cam.eye[:], np.float32); vd = np.array(cam.view_dir[:], np.float32); up = np.array(cam.up[:], np.float32)
    wv = -vd / np.linalg.norm(vd); u = np.cross(up, wv); u /= np.linalg.norm(u); v = np.cross(wv, u)
    top = np.tan(np.float32(cam.fov_deg) * np.float32(3.1415926 / 360.0)); right = top * w / h
    xs = (-right + 2 * right * (np.arange(w, dtype=np.float32) + 0.5) / w)[None, :, None]
    ys = (-top + 2 * top * (np.arange(h, dtype=np.float32) + 0.5) / h)[:, None, None]
    d = xs * u[None, None, :] + ys * v[None, None, :] - wv[None, None, :]
    d = (d / np.linalg.norm(d, axis=2, keepdims=True)).reshape(-1, 3).astype(np.float32)
    r = np.zeros(w * h, RAY_DTYPE)
    r["o"] = eye; r["d"] = d; r["tmin"] = 1e-3; r["tmax"] = 1e12
    return r


def incoherent_rays(lo, hi, n, seed):
    from miro_b200 import RAY_DTYPE
    rng = np.random.default_rng(seed)
    c, e = 0.5 * (lo + hi), 0.55 * (hi - lo) + 1e-3
    r = np.zeros(n, RAY_DTYPE)
    r["o"] = (c + e * rng.uniform(-1, 1, (n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["d"] = d.astype(np.float32); r["tmin"] = 1e-3; r["tmax"] = 1e12
    return r


def shadow_rays(src, hits, light):
    """Shadow rays as PointLight::sampleLight casts them (src/PointLight.cpp:20-48): from the hit point (or, for a
    miss, from the ray origin) towards the light, tmin 1e-3, tmax = distance."""
    from miro_b200 import RAY_DTYPE
    t = np.where(hits["prim"] >= 0, hits["t"], 0.0).astype(np.float32)
    p = src["o"] + t[:, None] * src["d"]
    L = light[None, :] - p
    dist = np.linalg.norm(L, axis=1).astype(np.float32)
    r = np.zeros(len(src), RAY_DTYPE)
    r["o"] = p; r["d"] = (L / np.maximum(dist, 1e-20)[:, None]).astype(np.float32); r["tmin"] = 1e-3; r["tmax"] = dist
    return r


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.rows = []; self.stop_flag = False; self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def write_obj_scene(fx, tmp):
    import helpers
    return helpers.write_obj_scene(fx, tmp)


def reference_trace(fx, batches, threads, repeat, warmup=0, mean=False):
    """Time the reference's own Scene::trace (oracle/_ref/miro_ref) on the host cores — ONE process: scene load and BVH
    build once, then `warmup` untimed and `repeat` timed passes over the batch (best, or mean when `mean`).
    Falls back to the oracle port when the reference binary is absent."""
    import helpers
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "miro_ref")
    rays = np.concatenate(batches)
    if os.path.exists(ref_bin):
        with tempfile.TemporaryDirectory() as tmp:
            sp = write_obj_scene(fx, tmp)
            rp = os.path.join(tmp, "rays.bin"); rays.tofile(rp)
            p = subprocess.run([ref_bin, "--scene", sp, "--assets", tmp, "--threads", str(threads), "--repeat", str(repeat), "--warmup", str(warmup), "--trace", rp],
                               stderr=subprocess.PIPE, text=True)
            ev = [json.loads(l) for l in p.stderr.splitlines() if l.startswith("{")]
            tr = [e for e in ev if e.get("event") == "trace"]
            if p.returncode == 0 and tr:
                return tr[0]["mean_seconds" if mean else "seconds"], len(rays), "reference", tr[0]["threads"]
    sc = fx.scene()
    ts = []
    for it in range(-warmup, repeat):
        t0 = time.time(); helpers.oracle_trace_closest(sc, rays); t1 = time.time()
        if it >= 0:
            ts.append(t1 - t0)
    return (float(np.mean(ts)) if mean else min(ts)), len(rays), "port", 1


def bind_to_gpu_numa_node(device):
    """Pin this rank to the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE any pinned host
    buffer is allocated, so the e2e leg's H2D / D2H copies do not cross the socket interconnect.  Best effort."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def render_leg(device):
    """Scene::raytraceImage through miro_gpu_render (wavefront path tracer) on the C4 stand-in at 1024x1024, 16 paths,
    4 indirect segments: rays = Scene::trace queries counted on the device.  Extra evidence beside the trace metric."""
    import re
    import torch
    import helpers
    path = helpers.fixture_path("c4_cornell_pt")
    if path is None:
        return None
    fx = helpers.Fixture(path)
    sc = fx.scene(script_override=re.sub(r"image \d+ \d+", "image 1024 1024", fx.script)).attach(device)
    p = sc.render_params()
    out = torch.zeros((p.height, p.width, 3), dtype=torch.float32, device="cuda")
    sc.render_device(out.data_ptr()); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        sc.reset_counters()
        t0 = time.time(); sc.render_device(out.data_ptr()); torch.cuda.synchronize(); best = min(best, time.time() - t0)
    c = sc.counters(); rays = c["rays_closest"] + c["rays_any"]
    sc.close()
    return {"workload": "C4 stand-in (Cornell box, Blinn, rectangle light x4 samples, emitter) 1024x1024, 16 paths, maxBounces 5",
            "rays": int(rays), "ms": best * 1e3, "Mrays_per_s": rays / best * 1e-6, "kernel_launches": int(c["kernel_launches"])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="miro_gpu", choices=["miro_gpu", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (kernel tuning runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "miro_gpu" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    import __graft_entry__ as g
    import helpers
    path = helpers.fixture_path(SCENE, full=True) or helpers.fixture_path(SCENE)
    fx = helpers.Fixture(path)
    allv = np.concatenate([fx.mesh(k)["vertices"] for k in range(len(fx.names))])
    lo, hi = allv.min(0), allv.max(0)
    hbm_peak, peak_src = peaks()
    config = {"workload": WORKLOAD, "scene": SCENE, "triangles": int(sum(len(fx.mesh(k)["vidx"]) for k in range(len(fx.names)))),
              "rays_per_step_per_gpu": 3 * N_BATCH, "l2_policy": "inputs_larger_than_L2 (298 MB of rays per step; the 6 MB BVH stays L2-resident by nature of the workload)",
              "sharding": "rays per rank, scene replicated, no collective",
              "launch_chaining": "the 3 traversal launches of a step are chained with programmatic dependent launch (miro_gpu_set_trace_chaining) in the timed region; per-launch times come from a separate, unchained pass"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        threads = min(os.cpu_count() or 1, 16)
        sc_cpu = fx.scene()
        cam = sc_cpu.camera()
        prim = primary_rays(cam, WIDTH, HEIGHT)
        inco = incoherent_rays(lo, hi, N_BATCH, 0x5EED)
        # bounded sample of the step: 1/8 of each batch (strided, keeps the coherence pattern of the rows)
        stride = 8
        sample = [prim.reshape(HEIGHT, WIDTH)[::2, ::4].reshape(-1).copy(), inco[::stride].copy()]
        ohits, _ = helpers.oracle_trace_closest(sc_cpu, sample[1])
        sample.append(shadow_rays(sample[1], ohits, LIGHT_POS))
        # K timed + W untimed passes inside one run of the reference binary (scene load / BVH build outside the timed region)
        sec, n, kind, used = reference_trace(fx, sample, threads, max(args.steps, 1), max(args.warmup, 0), mean=True)
        val = n / sec * 1e-6
        line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": used, "kind": kind,
                                                   "sample": "1/8 of each of the 3 ray batches of one step (%d rays) per step" % n},
                "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ GPU arm
    import torch
    import miro_b200 as mb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL's version banner goes to stdout, which carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = fx.scene().attach(local)
    cam = sc.camera()
    prim = primary_rays(cam, WIDTH, HEIGHT)
    inco = incoherent_rays(lo, hi, N_BATCH, 0x5EED + rank)
    inco_hits = sc.trace_closest(inco)
    shad = shadow_rays(inco, inco_hits, LIGHT_POS)
    batches = [prim, inco, shad]
    stream = torch.cuda.Stream()
    sc.set_stream(stream.cuda_stream)

    def dev(a):
        return torch.from_numpy(a.view(np.uint8).reshape(len(a), -1)).cuda()
    d_rays = [dev(b) for b in batches]
    d_hits = [torch.empty((N_BATCH, 20), dtype=torch.uint8, device="cuda") for _ in range(2)]
    d_bits = torch.empty(((N_BATCH + 31) // 32,), dtype=torch.int32, device="cuda")

    def step():
        sc.trace_closest_device(d_rays[0].data_ptr(), N_BATCH, d_hits[0].data_ptr())
        sc.trace_closest_device(d_rays[1].data_ptr(), N_BATCH, d_hits[1].data_ptr())
        sc.trace_any_device(d_rays[2].data_ptr(), N_BATCH, d_bits.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # algorithmic bytes: device counters of the SHIPPED layout (64 B quantized device nodes — the 128 B ABI node is
    # re-encoded at upload —, 48 B triangles, 64 B instances, 48 B ray in, 20 B hit / 1 bit out)
    sc.enable_counting(True)
    per_launch = []
    for i, f in enumerate((sc.trace_closest_device, sc.trace_closest_device, sc.trace_any_device)):
        sc.reset_counters()
        f(d_rays[i].data_ptr(), N_BATCH, (d_hits[min(i, 1)] if i < 2 else d_bits).data_ptr())
        c = sc.counters()
        out_bytes = 20 * N_BATCH if i < 2 else 4 * ((N_BATCH + 31) // 32)
        per_launch.append({"nodes": c["nodes_fetched"], "tris": c["tris_tested"],
                           "bytes": c["nodes_fetched"] * NODE_BYTES + c["tris_tested"] * 48 + c["insts_entered"] * 64 + 48 * N_BATCH + out_bytes})
    sc.enable_counting(False)

    sampler = ClockSampler(local); sampler.start()
    sc.set_trace_chaining(True)
    for _ in range(args.warmup):
        step()
    barrier()
    # timed region: K steps back to back, bracketed by two events on the launching stream (no events between the launches of a
    # step: consecutive traversal launches are chained by programmatic dependent launch, which an event record would break)
    sc.set_trace_chaining(True)       # all ray buffers of the step were complete long before: the contract holds
    e_first, e_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e_first.record(stream)
    for k in range(args.steps):
        step()
    e_last.record(stream)
    stream.synchronize()
    barrier()
    clocks = sampler.finish()
    total_ms = e_first.elapsed_time(e_last)
    sc.set_trace_chaining(False)
    # the chained launches must have produced what the unchained ones produce
    chk = d_hits[1].cpu().numpy().view(mb.HIT_DTYPE).reshape(-1)
    assert np.array_equal(chk["prim"], inco_hits["prim"]) and np.array_equal(chk["t"], inco_hits["t"])
    # per-launch durations (roofline of the dominant kernel): a second pass with an event after every launch
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for k in range(args.steps):
        ev[k][0].record(stream)
        sc.trace_closest_device(d_rays[0].data_ptr(), N_BATCH, d_hits[0].data_ptr()); ev[k][1].record(stream)
        sc.trace_closest_device(d_rays[1].data_ptr(), N_BATCH, d_hits[1].data_ptr()); ev[k][2].record(stream)
        sc.trace_any_device(d_rays[2].data_ptr(), N_BATCH, d_bits.data_ptr()); ev[k][3].record(stream)
    stream.synchronize()
    launch_ms = [float(np.mean([ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(args.steps)])) for i in range(3)]
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    rays_total = 3 * N_BATCH * args.steps * world
    value = rays_total / total_ms * 1e-3

    # end to end through the host-pointer ABI, pinned host buffers
    pinned = [torch.from_numpy(b.view(np.uint8).reshape(len(b), -1).copy()).pin_memory() for b in batches]
    out_hits = [torch.empty((N_BATCH, 20), dtype=torch.uint8).pin_memory() for _ in range(2)]
    out_bits = torch.empty(((N_BATCH + 31) // 32,), dtype=torch.int32).pin_memory()
    L = sc.L

    def e2e_step():
        L.miro_gpu_trace_closest(sc.ctx, pinned[0].data_ptr(), N_BATCH, out_hits[0].data_ptr())
        L.miro_gpu_trace_closest(sc.ctx, pinned[1].data_ptr(), N_BATCH, out_hits[1].data_ptr())
        L.miro_gpu_trace_any(sc.ctx, pinned[2].data_ptr(), N_BATCH, out_bits.data_ptr())
    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.time()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.time() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = 3 * N_BATCH * e2e_steps * world / e2e_s * 1e-6
    assert np.array_equal(out_hits[1].numpy().view(mb.HIT_DTYPE).reshape(-1)["prim"], inco_hits["prim"])

    # the same step with 32-byte packed rays (miro_gpu_trace_*_packed: the scene is static, so time / flags / user words carry
    # nothing): a third less over PCIe, which is what bounds the host-pointer calls
    pinned32 = [torch.from_numpy(mb.pack_rays(b).view(np.uint8).reshape(len(b), -1).copy()).pin_memory() for b in batches]

    def e2e_packed_step():
        L.miro_gpu_trace_closest_packed(sc.ctx, pinned32[0].data_ptr(), N_BATCH, out_hits[0].data_ptr())
        L.miro_gpu_trace_closest_packed(sc.ctx, pinned32[1].data_ptr(), N_BATCH, out_hits[1].data_ptr())
        L.miro_gpu_trace_any_packed(sc.ctx, pinned32[2].data_ptr(), N_BATCH, out_bits.data_ptr())
    for _ in range(2):
        e2e_packed_step()
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        e2e_packed_step()
    torch.cuda.synchronize()
    e2e32_s = time.time() - t0
    if world > 1:
        t = torch.tensor([e2e32_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e32_s = float(t.item())
    e2e32_val = 3 * N_BATCH * e2e_steps * world / e2e32_s * 1e-6
    chk32 = out_hits[1].numpy().view(mb.HIT_DTYPE).reshape(-1)
    assert np.array_equal(chk32["prim"], inco_hits["prim"]) and np.array_equal(chk32["t"], inco_hits["t"])

    if rank == 0:
        dom = int(np.argmax(launch_ms))
        kernel_names = ["k_trace<closest> primary", "k_trace<closest> incoherent", "k_trace<any> shadow"]
        traffic = l2_bytes = l1_bytes = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):      # per launch, from the committed ncu --set full capture: DRAM read+write, L2 sectors, L1 global-load sectors
            tj = json.load(open(tp))
            traffic = tj.get("per_launch_dram_bytes", {}).get(kernel_names[dom])
            l2_bytes = tj.get("per_launch_l2_bytes", {}).get(kernel_names[dom])
            l1_bytes = tj.get("per_launch_l1_global_load_bytes", {}).get(kernel_names[dom])
        achieved = per_launch[dom]["bytes"] / (launch_ms[dom] * 1e-3) * 1e-9
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
                "gpu_launches": 3 * args.steps,
                "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": 3 * N_BATCH * 48,
                        "d2h_bytes_per_step": 2 * N_BATCH * 20 + 4 * ((N_BATCH + 31) // 32), "steps": e2e_steps,
                        "ray_format": "miro_gpu_ray, 48 B (the general ABI record)"},
                "e2e_packed": {"value": e2e32_val, "unit": "Mrays/s", "h2d_bytes_per_step": 3 * N_BATCH * 32,
                               "d2h_bytes_per_step": 2 * N_BATCH * 20 + 4 * ((N_BATCH + 31) // 32), "steps": e2e_steps,
                               "ray_format": "miro_gpu_ray32, 32 B (static scenes: no time / flags words); identical hits asserted"},
                "roofline": {"bound": "hbm", "kernel": kernel_names[dom],
                             "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src,
                             "traffic": traffic, "l2_traffic": l2_bytes, "l1_global_load_traffic": l1_bytes,
                             "algorithmic_bytes_per_launch": per_launch[dom]["bytes"],
                             "nodes_per_ray": per_launch[dom]["nodes"] / N_BATCH, "tris_per_ray": per_launch[dom]["tris"] / N_BATCH,
                             "launch_ms": launch_ms[dom], "share_of_step": launch_ms[dom] / sum(launch_ms),
                             "all_launches": [{"kernel": kernel_names[i], "ms": launch_ms[i], "Mrays_per_s": N_BATCH / launch_ms[i] * 1e-3,
                                               "GBps": per_launch[i]["bytes"] / launch_ms[i] * 1e-6, "bytes_per_ray": per_launch[i]["bytes"] / N_BATCH,
                                               "nodes_per_ray": per_launch[i]["nodes"] / N_BATCH, "tris_per_ray": per_launch[i]["tris"] / N_BATCH}
                                              for i in range(3)]}}
        line["roofline"]["frac_abi_node_layout"] = (per_launch[dom]["bytes"] + per_launch[dom]["nodes"] * (128 - NODE_BYTES)) / (launch_ms[dom] * 1e-3) * 1e-9 / hbm_peak
        line["roofline"]["note"] = ("algorithmic bytes count the SHIPPED 64-byte device node; frac_abi_node_layout counts the 128-byte ABI node as kernel "
                                    "versions <= v4 fetched it (comparable with profiles/bench_r1_v1..v4.json). The BVH is L2-resident: see traffic.")
        line["render"] = render_leg(local)
        if world == 1 and not args.no_cpu:
            threads = min(os.cpu_count() or 1, 16)
            sample = [prim.reshape(HEIGHT, WIDTH)[::2, ::4].reshape(-1).copy(), inco[::8].copy(), shad[::8].copy()]
            sec, n, kind, used = reference_trace(fx, sample, threads, 3)
            line["cpu_baseline"] = {"value": n / sec * 1e-6, "unit": "Mrays/s", "cores": used, "kind": kind,
                                    "sample": "1/8 of each of the 3 ray batches of one step (%d rays), best of 3" % n}
        print(json.dumps(line))
    sc.set_stream(None)
    sc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
