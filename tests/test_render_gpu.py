"""GPU parity of miro_gpu_render (Scene::raytraceImage on the GPU) through the C ABI.

The product and the oracle share the random-number ADDRESSING (counter-based Philox) but not their structure
(wavefront queues vs. the reference's recursion), so non-dome images are compared pixel by pixel; dome-lit images
(alias table vs. CDF inversion of the same pmf) and everything against the reference itself are compared as
estimators (RMSE against the reference's converged render at equal spp)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def load(name):
    path = helpers.fixture_path(name)
    if path is None:
        pytest.skip("fixture %s not generated" % name)
    fx = helpers.Fixture(path)
    return fx, fx.scene().attach(0)


def rmse(a, b, clamp=4.0):
    return float(np.sqrt(np.mean((np.minimum(a, clamp) - np.minimum(b, clamp)) ** 2)))


def pixel_agreement(img, ref, rel=2e-3, ab=2e-4):
    err = np.abs(img - ref).max(axis=2)
    return (err <= rel * np.maximum(ref.max(axis=2), 1e-3) + ab)


def test_c1_image_matches_reference_and_oracle():
    fx, sc = load("c1_cornell")
    img, img8 = sc.render(want_bytes=True)
    oimg, orays = helpers.oracle_render(sc)
    ok = pixel_agreement(img, oimg, rel=1e-4, ab=1e-5)
    print("c1: gpu vs oracle agreement", ok.mean())
    assert ok.mean() > 0.998                             # the rest: edge ties (the oracle keeps the reference's non-watertight test)
    ok_ref = pixel_agreement(img, fx.radiance)           # float16 fixture
    assert ok_ref.mean() > 0.998
    # final 8-bit image against the reference's stock render: +-1 LSB except crack / tie pixels
    d8 = np.abs(img8.astype(int) - fx.image8.astype(int)).max(axis=2)
    assert (d8 <= 1).mean() > 0.998
    c = sc.counters()
    assert abs(int(c["rays_closest"] + c["rays_any"]) - orays) <= 2e-3 * orays
    sc.close()


@pytest.mark.parametrize("name", ["c4_cornell_pt", "c5_mb_instances", "c6_cornell_glass", "c7_foliage", "c9_texmaps", "c10_full_shadows"])
def test_image_matches_oracle_sample_by_sample(name):
    fx, sc = load(name)
    img = sc.render()
    oimg, orays = helpers.oracle_render(sc)
    assert np.isfinite(img).all()
    ok = pixel_agreement(img, oimg, rel=5e-3, ab=1e-3)
    print(name, "gpu vs oracle pixel agreement", ok.mean(), "means", img.mean(), oimg.mean())
    assert ok.mean() > 0.97, ok.mean()                  # paths that graze an edge diverge (watertight vs. reference test)
    assert abs(img.mean() - oimg.mean()) <= 0.01 * oimg.mean()
    c = sc.counters()
    assert abs(int(c["rays_closest"] + c["rays_any"]) - orays) <= 5e-3 * orays
    sc.close()


@pytest.mark.parametrize("name,mean_tol", [("c4_cornell_pt", 0.08), ("c3_dome_pt", 0.02), ("c6_cornell_glass", 0.02), ("c8_dispersion", 0.03), ("c10_full_shadows", 0.02), ("c11_dome_full_shadows", 0.02)])
def test_path_traced_estimator_matches_reference(name, mean_tol):
    """RMSE(gpu_N, ref_converged) <= 1.1 * RMSE(ref_N, ref_converged) at equal spp."""
    fx, sc = load(name)
    img = sc.render()
    ref, conv = fx.radiance, fx.radiance_converged
    e_g, e_r = rmse(img, conv), rmse(ref, conv)
    print(name, "rmse gpu/conv %.4f ref/conv %.4f" % (e_g, e_r), "means", img.mean(), ref.mean(), conv.mean())
    assert np.isfinite(img).all()
    assert e_g <= 1.1 * e_r, (e_g, e_r)
    # firefly-heavy scenes (bright dome samples without a cosine): the clamp biases a 16-path mean, so compare at equal spp
    target = ref if name in ("c11_dome_full_shadows",) else conv
    assert abs(np.minimum(img, 4).mean() - np.minimum(target, 4).mean()) <= mean_tol * np.minimum(target, 4).mean()
    sc.close()


def test_full_shadow_method_under_a_dome_light_matches_oracle():
    """Light::setFastShadows(false) on a DomeLight (DomeLight.cpp:123-146): shadow rays are walked hit by hit through a
    half-transparent teapot (refract 0.5).  GPU and oracle draw dome cells differently (alias table vs. CDF inversion), so the
    images are compared as estimators of each other; the walk must also differ visibly from the any-hit shadows."""
    path = helpers.fixture_path("c11_dome_full_shadows")
    if path is None:
        pytest.skip("fixture c11_dome_full_shadows not generated")
    fx = helpers.Fixture(path)
    full = fx.scene().attach(0)
    fast = fx.scene(script_override=fx.script.replace(" fastshadows 0", "")).attach(0)
    img_fast, img_full = fast.render(), full.render()
    oimg, orays = helpers.oracle_render(full)
    assert np.isfinite(img_full).all()
    blk = lambda a: np.minimum(a, 4).reshape(32, 8, 32, 8, 3).mean(axis=(1, 3))
    d_full = np.abs(blk(img_full) - blk(oimg)).mean(); d_fast = np.abs(blk(img_fast) - blk(oimg)).mean()
    print("dome full shadows: block diff gpu-full/oracle %.4f gpu-fast/oracle %.4f" % (d_full, d_fast), "means", img_full.mean(), oimg.mean(), img_fast.mean())
    assert abs(np.minimum(img_full, 4).mean() - np.minimum(oimg, 4).mean()) <= 0.02 * np.minimum(oimg, 4).mean()
    assert img_full.mean() > 1.02 * img_fast.mean()          # light leaks through the teapot
    assert d_full < 0.6 * d_fast
    c = full.counters()
    assert abs(int(c["rays_closest"] + c["rays_any"]) - orays) <= 0.02 * orays
    fast.close(); full.close()


def test_full_shadow_method_through_alpha_cutouts_matches_oracle():
    """The shadow walk (k_walk_shadows, one thread per ray) in a scene WITH alpha maps — its traversal evaluates the cut-outs like
    the persistent-warp kernels — through half-transparent leaves (refract 0.4), rectangle light with two samples per loop."""
    fx = helpers.Fixture(helpers.fixture_path("c7_foliage"))
    script = fx.script.replace("light point pos 15 60 -40 power 60000", "light rect v1 10 60 -40 v2 20 60 -40 v3 10 60 -30 power 60000 samples 2 fastshadows 0")
    script = script.replace("translucency 0.6 colormap", "translucency 0.6 refract 0.4 reflect 0 colormap")
    assert script != fx.script
    sc = fx.scene(script_override=script).attach(0)
    fast = fx.scene(script_override=script.replace(" fastshadows 0", "")).attach(0)
    img, img_fast = sc.render(), fast.render()
    oimg, orays = helpers.oracle_render(sc)
    ok = pixel_agreement(img, oimg, rel=5e-3, ab=1e-3)
    print("full shadows through alpha cut-outs: agreement", ok.mean(), "means", img.mean(), oimg.mean(), "any-hit method", img_fast.mean())
    assert np.isfinite(img).all() and ok.mean() > 0.97
    assert abs(img.mean() - oimg.mean()) <= 0.01 * oimg.mean()
    c, cf = sc.counters(), fast.counters()
    assert abs(int(c["rays_closest"] + c["rays_any"]) - orays) <= 5e-3 * orays
    assert int(c["rays_closest"] + c["rays_any"]) > 1.05 * int(cf["rays_closest"] + cf["rays_any"])      # the walk re-traces from every hit
    assert img.mean() > img_fast.mean()                  # and some light passes the leaves
    sc.close(); fast.close()


def test_sharded_render_equals_whole():
    """Buckets b % shard_count == shard_index; the union of the shards is the whole image (RNG keyed by pixel)."""
    fx, sc = load("c4_cornell_pt")
    whole = sc.render()
    import torch
    p = sc.render_params()
    ys, xs = np.mgrid[0:p.height, 0:p.width]
    bucket = (ys // 32) * ((p.width + 31) // 32) + xs // 32
    parts = np.zeros_like(whole)
    for i in range(3):
        out = torch.full((p.height, p.width, 3), -1.0, dtype=torch.float32, device="cuda")
        p.shard_index = i; p.shard_count = 3
        sc.render_device(out.data_ptr(), params=p)
        torch.cuda.synchronize()
        part = out.cpu().numpy()
        own = bucket % 3 == i
        assert (part[~own] == -1.0).all()              # pixels of other shards are left untouched
        assert (part[own] >= 0).all()
        parts[own] = part[own]
    assert np.allclose(parts, whole, rtol=1e-4, atol=1e-5)
    # host-pointer variant through Scene::raytraceImage
    hp = sc.render(shard_index=1, shard_count=3)
    assert np.allclose(hp[bucket % 3 == 1], whole[bucket % 3 == 1], rtol=1e-4, atol=1e-5)
    sc.close()


def test_sample_sharded_render_sums_to_whole():
    """path_shard_index / path_shard_count: the images of the path shards add up to the whole image."""
    import torch
    fx, sc = load("c4_cornell_pt")
    whole = sc.render()
    p = sc.render_params()
    acc = np.zeros_like(whole)
    for i in range(3):
        out = torch.zeros((p.height, p.width, 3), dtype=torch.float32, device="cuda")
        p.path_shard_index = i; p.path_shard_count = 3
        sc.render_device(out.data_ptr(), params=p)
        torch.cuda.synchronize()
        acc += out.cpu().numpy()
    assert np.allclose(acc, whole, rtol=2e-4, atol=2e-5)
    # more shards than paths: the surplus shards render nothing
    p.num_paths = 2; p.path_shard_index = 2; p.path_shard_count = 4
    out = torch.zeros((p.height, p.width, 3), dtype=torch.float32, device="cuda")
    sc.render_device(out.data_ptr(), params=p); torch.cuda.synchronize()
    assert float(out.abs().max()) == 0.0
    import miro_b200 as mb
    p = sc.render_params(); p.path_shard_count = 2; p.min_subdivs = 1; p.max_subdivs = 2
    with pytest.raises(mb.MiroError):
        sc.render_device(out.data_ptr(), params=p)
    sc.close()


def test_adaptive_levels_and_lens():
    """min/max subdivs > 1 (stratified levels, gamma-space cut-off) and a thin lens, against the oracle."""
    fx, sc = load("c1_cornell")
    p = sc.render_params(); p.width = p.height = 128; p.min_subdivs = 2; p.max_subdivs = 4; p.noise_threshold = 0.5
    cam = sc.camera(); cam.aperture = 0.05; cam.focus_plane = 6.0
    import torch
    out = torch.zeros((128, 128, 3), dtype=torch.float32, device="cuda")
    sc.render_device(out.data_ptr(), params=p, camera=cam)
    torch.cuda.synchronize()
    img = out.cpu().numpy()
    oimg, _ = helpers.oracle_render(sc, params=p, camera=cam)
    ok = pixel_agreement(img, oimg, rel=5e-3, ab=1e-3)
    print("adaptive: agreement", ok.mean())
    assert ok.mean() > 0.97
    assert abs(img.mean() - oimg.mean()) < 0.005 * oimg.mean()
    sc.close()


def test_dispersion_matches_oracle_sample_by_sample():
    """The c8 prism sphere with the dome light replaced by a rectangle light (the dome's alias table draws different cells
    than the oracle's CDF inversion): three refraction rays per split, masked throughput, the all-three-missed environment
    rule, IOR history — pixel by pixel against the oracle's recursion."""
    fx = helpers.Fixture(helpers.fixture_path("c8_dispersion"))
    script = fx.script.replace("light dome tex sky power 0.15 samples 6", "light rect v1 -1 6 -1 v2 1 6 -1 v3 -1 6 1 power 40 samples 2")
    assert "light rect" in script
    sc = fx.scene(script_override=script).attach(0)
    img = sc.render()
    oimg, orays = helpers.oracle_render(sc)
    ok = pixel_agreement(img, oimg, rel=5e-3, ab=1e-3)
    print("dispersion: gpu vs oracle pixel agreement", ok.mean(), "means", img.mean(), oimg.mean())
    assert np.isfinite(img).all() and ok.mean() > 0.97
    assert abs(img.mean() - oimg.mean()) <= 0.01 * oimg.mean()
    c = sc.counters()
    assert abs(int(c["rays_closest"] + c["rays_any"]) - orays) <= 5e-3 * orays
    sc.close()


def test_render_errors():
    import miro_b200 as mb
    fx, sc = load("c1_cornell")
    p = sc.render_params(); p.width = 0
    with pytest.raises(mb.MiroError):
        sc.render_device(0, params=p)
    p = sc.render_params(); p.shard_index = 3; p.shard_count = 2
    import torch
    out = torch.zeros((p.height, p.width, 3), dtype=torch.float32, device="cuda")
    with pytest.raises(mb.MiroError):
        sc.render_device(out.data_ptr(), params=p)
    sc.close()


def test_two_gpu_render_nccl(tmp_path):
    """Two ranks, one GPU each: bucket-sharded miro_gpu_render + one NCCL all_reduce == the single-GPU frame."""
    import os, subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    worker = r'''
import os, sys
sys.path.insert(0, os.environ["MIRO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["MIRO_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import helpers
from miro_b200 import distributed as md
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % os.environ["MIRO_PORT"], rank=rank, world_size=2, device_id=torch.device("cuda", rank))
fx = helpers.Fixture(helpers.fixture_path("c4_cornell_pt")); sc = fx.scene().attach(rank)
full = md.render_scene_distributed(sc, rank, 2)
full_s = md.render_scene_distributed(sc, rank, 2, mode="samples")
if rank == 0:
    np.save(os.environ["MIRO_OUT"], full.cpu().numpy()); np.save(os.environ["MIRO_OUT"] + ".samples.npy", full_s.cpu().numpy())
dist.destroy_process_group(); sc.close()
'''
    out = tmp_path / "full.npy"
    env = dict(os.environ, MIRO_ROOT=helpers.ROOT, MIRO_PORT=str(29600 + os.getpid() % 2000), MIRO_OUT=str(out))
    procs = [subprocess.Popen([sys.executable, "-c", worker], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    fx, sc = load("c4_cornell_pt")
    whole = sc.render()
    assert np.allclose(np.load(out), whole, rtol=1e-4, atol=1e-5)
    assert np.allclose(np.load(str(out) + ".samples.npy"), whole, rtol=2e-4, atol=2e-5)
    sc.close()


def test_headless_cli_writes_the_reference_ppm(tmp_path):
    """miro_render scene.miro out.ppm: OBJ loader + script + GPU render + Image::Map + bottom-up PPM, against the
    8-bit image of the reference's stock render (+-1 LSB except crack / tie pixels)."""
    import os, subprocess
    exe = os.path.join(helpers.ROOT, "rendering-algorithms-raytracer_b200", "miro_render")
    if not os.path.exists(exe):
        pytest.skip("miro_render not built")
    fx = helpers.Fixture(helpers.fixture_path("c1_cornell"))
    sp = helpers.write_obj_scene(fx, str(tmp_path))
    out = tmp_path / "o.ppm"
    p = subprocess.run([exe, sp, str(out), "--assets", str(tmp_path), "--stats"], stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert '"rays":' in p.stderr
    raw = open(out, "rb").read().split(b"\n", 3)
    w, h = map(int, raw[1].split())
    img = np.frombuffer(raw[3], np.uint8).reshape(h, w, 3)[::-1]       # PPM rows are top-down
    d8 = np.abs(img.astype(int) - fx.image8.astype(int)).max(axis=2)
    assert (d8 <= 1).mean() > 0.998


def test_reference_binding_renders_the_references_own_scene_on_the_gpu(tmp_path):
    """INTEGRATION.md made real: the UNMODIFIED reference (oracle/_ref/miro_ref) loads the scene with its own loaders, builds its
    own QBVH, and its --render-gpu glue flattens the reference's objects / materials / lights / tree into a
    miro_gpu_scene_desc and calls libmiro_gpu.so (dlopen) for Scene::raytraceImage.  The float radiance must be the
    reference's own CPU render (golden fixture), and the 8-bit image through the reference's Image::Map its stock render."""
    import os, subprocess
    exe = os.path.join(helpers.ROOT, "oracle", "_ref", "miro_ref")
    lib = os.path.join(helpers.ROOT, "rendering-algorithms-raytracer_b200", "libmiro_gpu.so")
    if not os.path.exists(exe):
        pytest.skip("reference binary not built (oracle/_ref)")
    fx = helpers.Fixture(helpers.fixture_path("c1_cornell"))
    sp = helpers.write_obj_scene(fx, str(tmp_path))
    out, ppm = tmp_path / "gpu.f32", tmp_path / "gpu.ppm"
    p = subprocess.run([exe, "--scene", sp, "--assets", str(tmp_path), "--render-gpu", str(out), "--gpu-lib", lib, "--gpu-ppm", str(ppm)],
                       stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert '"event":"render_gpu"' in p.stderr
    H, W = fx.radiance.shape[:2]
    img = np.fromfile(out, np.float32).reshape(H, W, 3)
    ok = pixel_agreement(img, fx.radiance)               # float16 fixture; the rest: crack / tie pixels of the reference
    print("reference binding: gpu vs reference radiance agreement", ok.mean())
    assert ok.mean() > 0.998
    raw = open(ppm, "rb").read().split(b"\n", 3)
    img8 = np.frombuffer(raw[3], np.uint8).reshape(H, W, 3)[::-1]
    assert (np.abs(img8.astype(int) - fx.image8.astype(int)).max(axis=2) <= 1).mean() > 0.998


def test_reference_binding_renders_motion_blur_and_instances_on_the_gpu(tmp_path):
    """The same binding on c5: MBObjects and 961 ProxyObjects sharing one bottom-level tree, mixed TriCache4 packets — the
    reference's own object model flattened by the glue, rendered by libmiro_gpu.so; compared with the reference's CPU render as
    in test_oracle_render.py (the jittered camera samples draw different random numbers)."""
    import json, os, subprocess
    exe = os.path.join(helpers.ROOT, "oracle", "_ref", "miro_ref")
    lib = os.path.join(helpers.ROOT, "rendering-algorithms-raytracer_b200", "libmiro_gpu.so")
    if not os.path.exists(exe):
        pytest.skip("reference binary not built (oracle/_ref)")
    fx = helpers.Fixture(helpers.fixture_path("c5_mb_instances"))
    sp = helpers.write_obj_scene(fx, str(tmp_path))
    out = tmp_path / "gpu.f32"
    p = subprocess.run([exe, "--scene", sp, "--assets", str(tmp_path), "--render-gpu", str(out), "--gpu-lib", lib], stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    ev = [json.loads(l) for l in p.stderr.splitlines() if l.startswith("{")]
    ref_rays = [e for e in json.loads(fx.events) if e["event"] == "render_float"][0]["rays"]
    gpu = [e for e in ev if e["event"] == "render_gpu"][0]
    assert abs(gpu["rays"] - ref_rays) <= 0.01 * ref_rays
    H, W = fx.radiance.shape[:2]
    img = np.fromfile(out, np.float32).reshape(H, W, 3)
    ref = fx.radiance
    blk = lambda a: a.reshape(32, 8, 32, 8, 3).mean(axis=(1, 3))
    print("reference binding c5: means", img.mean(), ref.mean(), "block diff", np.abs(blk(img) - blk(ref)).mean())
    assert abs(img.mean() - ref.mean()) <= 0.02 * ref.mean()
    assert np.abs(blk(img) - blk(ref)).mean() < 0.01


def test_render_image_delivers_the_8_bit_frame():
    """miro_gpu_render_image: rgb8 = every channel of the float frame through Image::setPixel's Map (src/Image.cpp:71-87: clamp,
    32 769-entry 2.2-gamma table), applied on the device; host and device destinations, either output alone, shards."""
    import ctypes as C
    import torch
    fx, sc = load("c4_cornell_pt")
    p = sc.render_params(); cam = sc.camera()
    L = sc.L
    n = p.width * p.height * 3
    rgb = np.zeros(n, np.float32); rgb8 = np.zeros(n, np.uint8)
    assert L.miro_gpu_render_image(sc.ctx, C.byref(cam), C.byref(p), rgb.ctypes.data, rgb8.ctypes.data) == 0
    # the table as Image::generateGammaTables builds it (float arithmetic), indexed as Map indexes it
    lut = (np.power(np.arange(32769, dtype=np.float32) / np.float32(32768.0), np.float32(1 / 2.2)) * np.float32(255.0) + np.float32(0.5)).astype(np.int32)
    idx = np.where(rgb * np.float32(32768.0) > 32768.0, 32768, np.maximum(rgb * np.float32(32768.0), 0).astype(np.int32))
    want = lut[idx].astype(np.uint8)
    diff = np.abs(want.astype(int) - rgb8.astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3, (diff.max(), (diff != 0).mean())      # powf vs numpy's float32 power at a table entry or two
    assert rgb8.max() == 255 and rgb8.min() == 0
    # bytes only; bytes into device memory; nothing at all is an error
    only8 = np.zeros(n, np.uint8)
    assert L.miro_gpu_render_image(sc.ctx, C.byref(cam), C.byref(p), None, only8.ctypes.data) == 0
    d8 = torch.zeros(n, dtype=torch.uint8, device="cuda")
    assert L.miro_gpu_render_image(sc.ctx, C.byref(cam), C.byref(p), None, d8.data_ptr()) == 0
    torch.cuda.synchronize()
    for other in (only8, d8.cpu().numpy()):      # path-traced frames differ between renders in the last bits (float atomics): a byte step here and there
        assert (np.abs(other.astype(int) - rgb8.astype(int)) > 1).mean() == 0 and (other != rgb8).mean() < 0.02
    assert L.miro_gpu_render_image(sc.ctx, C.byref(cam), C.byref(p), None, None) != 0
    # a shard touches its own pixels only, in both outputs
    p.shard_index, p.shard_count = 1, 3
    part = np.full(n, 7, np.uint8); partf = np.full(n, -1.0, np.float32)
    assert L.miro_gpu_render_image(sc.ctx, C.byref(cam), C.byref(p), partf.ctypes.data, part.ctypes.data) == 0
    ys, xs = np.mgrid[0:p.height, 0:p.width]
    own = ((ys // 32) * ((p.width + 31) // 32) + xs // 32) % 3 == 1
    part = part.reshape(p.height, p.width, 3); partf = partf.reshape(p.height, p.width, 3)
    assert (part[~own] == 7).all() and (partf[~own] == -1.0).all()
    assert (np.abs(part[own].astype(int) - rgb8.reshape(p.height, p.width, 3)[own].astype(int)) <= 1).all()
    sc.close()


def test_frame_lands_in_the_scenes_image():
    """miro_host_raytrace_image without copies: the frame is in the scene's Image (float radiance + 8-bit pixels), as
    Scene::raytraceImage leaves it in the reference; the copying variant returns the same frame."""
    fx, sc = load("c1_cornell")
    rgb, rgb8 = sc.render_in_place()
    rgb = rgb.copy(); rgb8 = rgb8.copy()
    again, again8 = sc.render(want_bytes=True)
    assert np.array_equal(rgb, again) and np.array_equal(rgb8, again8)
    assert rgb8.max() > 100
    sc.close()
