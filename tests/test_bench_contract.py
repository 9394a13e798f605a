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
