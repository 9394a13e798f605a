#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer: trace (closest / any, ragged sizes), instanced + motion-blur scene,
and a small path-traced render.  usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers

for name in ["c1_cornell", "c5_mb_instances", "c4_cornell_pt"]:
    fx = helpers.Fixture(helpers.fixture_path(name))
    script = re.sub(r"image \d+ \d+", "image 48 40", fx.script)
    script = re.sub(r"numpaths \d+", "numpaths 2", script)
    sc = fx.scene(script_override=script).attach(0)
    for n in (1, 33, 1000):
        h = sc.trace_closest(fx.rays[:n]); o = sc.trace_any(fx.rays[:n])
        assert (o == (h["prim"] >= 0)).all()
    img = sc.render()
    assert np.isfinite(img).all()
    print(name, "ok", float(img.mean()))
    sc.close()
