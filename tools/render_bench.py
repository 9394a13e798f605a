#!/usr/bin/env python
"""Times miro_gpu_render (Scene::raytraceImage on the GPU) on the fixture scenes at a given resolution and reports
Mrays/s (rays = Scene::trace queries, counted by the device), next to the reference's CPU render of the same script
when oracle/_ref/miro_ref is present.  usage: tools/render_bench.py [scene ...] [--size N] [--paths N] [--cpu]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scenes", nargs="*", default=["c1_cornell", "c4_cornell_pt", "c3_dome_pt", "c5_mb_instances"])
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--paths", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import helpers
    for name in args.scenes:
        fx = helpers.Fixture(helpers.fixture_path(name, full=True) or helpers.fixture_path(name))
        script = fx.script
        if args.size:
            script = re.sub(r"image \d+ \d+", "image %d %d" % (args.size, args.size), script)
        if args.paths:
            script = re.sub(r"numpaths \d+", "numpaths %d" % args.paths, script)
        sc = fx.scene(script_override=script).attach(0)
        p = sc.render_params()
        out = torch.zeros((p.height, p.width, 3), dtype=torch.float32, device="cuda")
        sc.render_device(out.data_ptr()); torch.cuda.synchronize()
        best = 1e30
        for _ in range(args.reps):
            sc.reset_counters()
            t0 = time.time(); sc.render_device(out.data_ptr()); torch.cuda.synchronize(); dt = time.time() - t0
            best = min(best, dt)
        c = sc.counters()
        rays = c["rays_closest"] + c["rays_any"]
        print(json.dumps({"scene": name, "size": [p.width, p.height], "num_paths": p.num_paths, "rays": rays, "ms": best * 1e3,
                          "Mrays_per_s": rays / best * 1e-6, "kernel_launches": c["kernel_launches"], "mean": float(out.mean())}))
        sc.close()


if __name__ == "__main__":
    main()
