#!/usr/bin/env python
"""Scene::raytraceImage across N GPUs of one box (BASELINE config C4: path tracing with tile / sample sharding).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/render_scale.py
One process per GPU, scene replicated, each rank renders its 32x32 buckets (mode tiles) or its paths (mode samples) into a
torch CUDA frame, ONE NCCL all_reduce(SUM) of the frame at the end (inside the timed region).  Rank 0 prints one JSON line per
(scene, mode): frame time = max over ranks of the device-synchronised wall time, rays = Scene::trace queries of all ranks."""
import argparse
import json
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scenes", nargs="*", default=["c4_cornell_pt", "c3_dome_pt"])
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import helpers
    from miro_b200 import distributed as md
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for name in args.scenes:
        fx = helpers.Fixture(helpers.fixture_path(name, full=True) or helpers.fixture_path(name))
        script = re.sub(r"image \d+ \d+", "image %d %d" % (args.size, args.size), fx.script)
        script = re.sub(r"minsubdivs \d+ maxsubdivs \d+", "minsubdivs 1 maxsubdivs 1", script)      # fixed level: both sharding modes apply
        sc = fx.scene(script_override=script).attach(local)
        p = sc.render_params()
        for mode in ("tiles", "samples"):
            md.render_scene_distributed(sc, rank, world, mode=mode)       # warm-up (queues, NCCL communicator)
            best, rays, mean = 1e30, 0, 0.0
            for _ in range(args.reps):
                sc.reset_counters()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.time()
                frame = md.render_scene_distributed(sc, rank, world, mode=mode)
                torch.cuda.synchronize()
                dt = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
                c = sc.counters()
                r = torch.tensor([c["rays_closest"] + c["rays_any"]], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(r, op=dist.ReduceOp.SUM)
                if float(dt.item()) < best:
                    best, rays, mean = float(dt.item()), int(r.item()), float(frame.mean())
            if rank == 0:
                print(json.dumps({"scene": name, "mode": mode, "n_gpus": world, "size": [p.width, p.height], "num_paths": p.num_paths, "max_bounces": p.max_bounces,
                                  "rays": rays, "ms": best * 1e3, "Mrays_per_s": rays / best * 1e-6, "frame_mean": mean,
                                  "collective": "one NCCL all_reduce(SUM) of %d MB" % (p.width * p.height * 12 // 1000000)}), flush=True)
        sc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
