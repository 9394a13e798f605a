"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm prints ONE JSON line with the keys
the driver reads, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

import helpers

BENCH = os.path.join(helpers.ROOT, "bench.py")
REF_BIN = os.path.join(helpers.ROOT, "oracle", "_ref", "miro_ref")


def test_reference_arm_prints_one_json_line():
    if not os.path.exists(REF_BIN):
        pytest.skip("reference binary not built (oracle/_ref)")
    p = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["config"]["workload"].startswith("C2")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
    assert not any(l.startswith("{") for l in p.stdout.splitlines())


def test_bench_workloads_are_deterministic_and_named():
    """Both arms of bench.py build their scenes and ray batches from tests/bench_workloads.py: the same script text and the same
    seeded rays on every call (the reference arm and the product arm run in different processes)."""
    import numpy as np
    import bench_workloads as bw
    for name, tris in (("c2", 86914), ("big", 20 * 86914)):
        a, b = bw.Workload(name).load(), bw.Workload(name).load()
        assert a.script == b.script and a.triangles() == tris, name
        assert np.array_equal(a.primary(), b.primary()) and len(a.primary()) == 1920 * 1080
        assert a.incoherent(7).tobytes() == b.incoherent(7).tobytes() and a.incoherent(7).tobytes() != a.incoherent(8).tobytes()
        p, q = a.sample(a.primary(), a.incoherent(7))
        assert len(p) == len(q) == 1920 * 1080 // 8
    c5 = bw.Workload("c5").load()
    assert c5.script.count("\ninstance ") == 201 * 201 and c5.times
    assert (c5.incoherent(1)["time"] > 0).any()
    big = bw.Workload("big").load()
    assert len(big.mesh_names()) == 20 and big.label.startswith("C2 at dragon scale")
    lo, hi = big.bounds()
    assert (hi - lo > np.array([4.0, 0.9, 2.0])).all()


def test_committed_traffic_figures_come_from_the_committed_ncu_pages():
    """profiles/traffic.json (read by bench.py for roofline.traffic / l2_traffic / l1_global_load_traffic) is what
    tools/traffic_from_ncu.py derives from the committed `ncu --set full --page raw` captures of the shipped kernels."""
    import json
    ROOT = helpers.ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import traffic_from_ncu as T
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for key, page in (("c2", "r2_c2_flat_ncu_raw.csv"), ("big", "r2_big_flat_ncu_raw.csv")):
        kernels, dram, l2, l1 = T.traffic_of(os.path.join(ROOT, "profiles", page))
        assert all("k_trace_flat" in k for k in kernels), kernels
        assert tj[key]["captured_kernels"] == kernels
        for name in T.NAMES:
            assert abs(tj[key]["per_launch_dram_bytes"][name] - dram[name]) < 1.0
            assert abs(tj[key]["per_launch_l2_bytes"][name] - l2[name]) < 1.0
            assert abs(tj[key]["per_launch_l1_global_load_bytes"][name] - l1[name]) < 1.0
            assert dram[name] < l2[name]      # L2-resident structures: DRAM carries the ray / hit streams only
