#!/usr/bin/env python
"""Diagnostic: per-chunk device timeline of one host-pointer trace call (MIRO_GPU_TIMELINE=1 makes the library print it)."""
import os, sys, time
os.environ["MIRO_GPU_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers, bench
import miro_b200 as mb
fx = helpers.Fixture(helpers.fixture_path("c2_explosion", full=True) or helpers.fixture_path("c2_explosion"))
sc = fx.scene().attach(0)
allv = np.concatenate([fx.mesh(k)["vertices"] for k in range(len(fx.names))])
inco = bench.incoherent_rays(allv.min(0), allv.max(0), bench.N_BATCH, 0x5EED)
pin = torch.from_numpy(inco.view(np.uint8).reshape(len(inco), -1).copy()).pin_memory()
out = torch.empty((bench.N_BATCH, 20), dtype=torch.uint8).pin_memory()
for rep in range(3):
    t0 = time.time()
    sc.L.miro_gpu_trace_closest(sc.ctx, pin.data_ptr(), bench.N_BATCH, out.data_ptr())
    print("call wall ms", (time.time() - t0) * 1e3, file=sys.stderr)
