#!/usr/bin/env python
"""Strong scaling of ONE frame through miro_gpu_group_* (one caller, N GPUs of one box): the C4 and C3 stand-ins at 2048x2048, bucket
sharding (and sample sharding for C4), N = 1, 2, 4, 8 as far as the box has devices.  Frame time = wall time of
miro_host_raytrace_image with the frame left in the scene's Image (render on every device + combine over peer memory + download of
the float and the 8-bit frame into page-locked host memory), best of 3.
usage: tools/group_scale.py [--size 2048] [scene ...]"""
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
    args = ap.parse_args()
    import numpy as np
    import torch
    import helpers
    n_dev = torch.cuda.device_count()
    for name in args.scenes:
        fx = helpers.Fixture(helpers.fixture_path(name))
        script = re.sub(r"image \d+ \d+", "image %d %d" % (args.size, args.size), fx.script)
        base = None
        for mode in ("buckets", "samples"):
            if mode == "samples" and name != "c4_cornell_pt":
                continue
            for n in [k for k in (1, 2, 4, 8) if k <= n_dev]:
                sc = fx.scene(script_override=script).attach_devices(list(range(n)), sample_sharding=(mode == "samples"))
                img, _ = sc.render_in_place()
                best = 1e30
                for _ in range(3):      # the frame as Scene::raytraceImage leaves it: in the scene's (page-locked) Image, float + 8-bit
                    sc.L.miro_gpu_group_reset_counters(sc.group)
                    t0 = time.time(); img, _ = sc.render_in_place(); best = min(best, time.time() - t0)
                img = img.copy()
                c = sc.group_counters(); rays = int(c["rays_closest"] + c["rays_any"])
                peers = [int(sc.L.miro_gpu_group_peer_access(sc.group, i)) for i in range(n)]
                if n == 1:
                    base = best
                    ref_img = img
                err = float(np.abs(img - ref_img).max() / max(float(ref_img.max()), 1e-9))
                print(json.dumps({"scene": name, "sharding": mode, "n_gpus": n, "size": args.size, "rays": rays, "ms": best * 1e3, "Mrays_per_s": rays / best * 1e-6,
                                  "speedup": base / best, "efficiency": base / best / n, "peer_access": peers, "max_rel_diff_vs_one_gpu": err,
                                  "frame_mean": float(np.minimum(img, 4).mean())}), flush=True)
                sc.close()


if __name__ == "__main__":
    main()
