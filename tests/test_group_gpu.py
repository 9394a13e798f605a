"""One caller, several GPUs behind the C ABI (miro_gpu_group_*, csrc/multi.cu): the frame of a group equals the single-GPU
frame, batched Scene::trace split over a group gives the single-GPU hits.  A group may list one device twice (two contexts on
one GPU), so the whole code path — worker threads, shard parameters, the combine kernel — runs on a 1-GPU box; with two or more
GPUs the same tests run across devices (peer-memory loads)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers
from miro_b200 import capi

pytestmark = pytest.mark.gpu


def device_lists():
    import torch
    n = torch.cuda.device_count()
    out = [[0, 0], [0, 0, 0]]
    if n >= 2:
        out.append(list(range(min(n, 8))))
    return out


@pytest.mark.parametrize("name", ["c1_cornell", "c5_mb_instances", "c4_cornell_pt"])
def test_group_frame_equals_the_single_gpu_frame(name):
    fx = helpers.Fixture(helpers.fixture_path(name))
    one = fx.scene().attach(0)
    whole = one.render()
    one.close()
    for devs in device_lists():
        sc = fx.scene().attach_devices(devs)
        img = sc.render()
        # bucket sharding renders every pixel exactly as the whole frame does; what differs between ANY two renders of a path-traced
        # frame is the order in which a light loop's samples are added to its accumulator (float atomics)
        if name == "c1_cornell":
            assert np.array_equal(img, whole), (name, devs, float(np.abs(img - whole).max()))
        assert np.allclose(img, whole, rtol=1e-4, atol=1e-5), (name, devs, float(np.abs(img - whole).max()))
        c = sc.group_counters()
        assert c["rays_closest"] > 0 and c["kernel_launches"] > 0
        sc.close()


def test_group_sample_sharding_sums_to_the_whole_frame():
    fx = helpers.Fixture(helpers.fixture_path("c4_cornell_pt"))      # 16 paths, one subdivision level
    one = fx.scene().attach(0)
    whole = one.render()
    one.close()
    for devs in device_lists():
        sc = fx.scene().attach_devices(devs, sample_sharding=True)
        img = sc.render()
        err = np.abs(img - whole).max(axis=2) / np.maximum(whole.max(axis=2), 1e-3)
        assert (err < 1e-4).mean() > 0.9999, (devs, float(err.max()))      # the same samples, summed in another order
        sc.close()


def test_group_trace_splits_the_batch():
    fx = helpers.Fixture(helpers.fixture_path("c2_explosion"))
    one = fx.scene().attach(0)
    for n in (len(fx.rays), 1000, 33, 1):
        rays = fx.rays[:n]
        hits = one.trace_closest(rays); occ = one.trace_any(rays)
        for devs in device_lists():
            sc = fx.scene().attach_devices(devs)
            assert sc.trace_closest(rays).tobytes() == hits.tobytes(), (n, devs)
            assert np.array_equal(sc.trace_any(rays), occ), (n, devs)
            sc.close()
    one.close()


def test_group_errors():
    L = capi.lib()
    g = C.c_void_p()
    ids = (C.c_int * 2)(0, 99)
    assert L.miro_gpu_group_create(C.byref(g), ids, 2) == capi.EINVAL and not g.value
    assert L.miro_gpu_group_create(C.byref(g), ids, 0) == capi.EINVAL
    fx = helpers.Fixture(helpers.fixture_path("c1_cornell"))
    sc = fx.scene().attach_devices([0, 0])
    p = sc.render_params(); cam = sc.camera()
    p.shard_count = 2
    out = np.zeros((p.height, p.width, 3), np.float32)
    assert L.miro_gpu_group_render(sc.group, C.byref(cam), C.byref(p), 0, out.ctypes.data, None) == capi.EINVAL
    assert b"shards the frame itself" in L.miro_gpu_group_last_error(sc.group)
    sc.close()


def test_headless_cli_over_a_group(tmp_path):
    """miro_render --devices 0,0 writes the PPM miro_render --device 0 writes."""
    exe = os.path.join(helpers.ROOT, "rendering-algorithms-raytracer_b200", "miro_render")
    fx = helpers.Fixture(helpers.fixture_path("c1_cornell"))
    sp = helpers.write_obj_scene(fx, str(tmp_path))
    a, b = str(tmp_path / "one.ppm"), str(tmp_path / "group.ppm")
    subprocess.run([exe, sp, a, "--assets", str(tmp_path)], check=True)
    p = subprocess.run([exe, sp, b, "--assets", str(tmp_path), "--devices", "0,0", "--stats"], check=True, stderr=subprocess.PIPE, text=True)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert '"rays"' in p.stderr
