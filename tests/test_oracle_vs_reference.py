"""Pins the oracle (oracle/miro_oracle.c, the CPU restatement) against hit dumps of the UNMODIFIED reference
(tests/golden/*.npz, made by tools/make_fixtures.py with oracle/_ref/miro_ref).  CPU only."""
import numpy as np
import pytest

import helpers

SCENES = ["c1_cornell", "c2_explosion", "c5_mb_instances", "c7_foliage"]      # c7: alpha cut-outs inside Scene::trace


@pytest.fixture(scope="module", params=SCENES)
def loaded(request):
    path = helpers.fixture_path(request.param)
    if path is None:
        pytest.skip("fixture %s not generated" % request.param)
    fx = helpers.Fixture(path)
    sc = fx.scene()
    yield request.param, fx, sc
    sc.close()


def as_ref(sc, ohits):
    m, t, p = sc.resolve_hits(ohits)
    r = np.zeros(len(ohits), helpers.REFHIT)
    r["t"], r["a"], r["b"], r["mesh"], r["tri"], r["proxy"] = ohits["t"], ohits["a"], ohits["b"], m, t, p
    return r


def test_oracle_reproduces_reference_hits(loaded):
    name, fx, sc = loaded
    ohits, ctr = helpers.oracle_trace_closest(sc, fx.rays)
    st = helpers.compare_hits(sc, ohits, fx.hits, t_rel=1e-5, rays=fx.rays)
    print(name, {k: v for k, v in st.items() if k != "hard_idx"}, "nodes/ray %.2f tris/ray %.2f" % (ctr[0] / len(fx.rays), ctr[1] / len(fx.rays)))
    assert st["hard"] == 0, st
    assert st["id_match"] >= 0.9995, st            # the rest: exact-t ties resolved by a different (but legal) visiting order
    # instanced and motion-blurred hits (c5) included: the object-space ray is formed with the reference's own rounding
    # (reference-order inverse, dpps-order transform, recipps(w): DESIGN.md section 4), so t agrees to a few 1e-7 everywhere
    assert st["frac_t_within"] == 1.0 and st["max_rel_t"] < 2e-6, st
    assert st["max_abs_a"] < 1e-3 and st["max_abs_b"] < 1e-3, st


def test_oracle_any_is_closest_as_boolean(loaded):
    """The reference's shadow query is a closest-hit query used as a boolean (src/PointLight.cpp:44, BVH.cpp:1156-1160)."""
    name, fx, sc = loaded
    r = fx.rays[:4096]
    occ = helpers.oracle_trace_any(sc, r)
    assert (occ == (fx.hits["mesh"][:4096] >= 0)).mean() >= 0.999


def test_fixture_is_self_consistent(loaded):
    name, fx, sc = loaded
    h = fx.hits
    hit = h["mesh"] >= 0
    assert hit.any() and (~hit).any() or name == "c1_cornell"
    assert (h["t"][hit] >= fx.rays["tmin"][hit]).all() and (h["t"][hit] < fx.rays["tmax"][hit]).all()
    assert (h["a"][hit] >= 0).all() and (h["b"][hit] >= 0).all() and (h["a"][hit] + h["b"][hit] <= 1 + 1e-6).all()
    # the recorded hit point lies on the recorded triangle (reference geometry carried by the fixture)
    mesh_of, tri_of, _ = sc.prim_table()
    d = sc.desc()
    assert d.n_tris + d.n_mbtris == len(mesh_of)


@pytest.mark.parametrize("name", ["c1_cornell", "c2_explosion"])
def test_oracle_on_the_references_own_tree_is_exact(name):
    """The reference's own QBVH flattened 1:1 (helpers.ReferenceTreeScene): the oracle visits the same nodes in the same order
    with the same packet semantics, so even exact-t ties resolve the way the reference resolved them."""
    fx = helpers.Fixture(helpers.fixture_path(name))
    sc = helpers.ReferenceTreeScene(fx)
    ohits, ctr = helpers.oracle_trace_closest(sc, fx.rays)
    st = helpers.compare_hits(sc, ohits, fx.hits, t_rel=1e-5, rays=fx.rays)
    print(name, {k: v for k, v in st.items() if k != "hard_idx"}, "nodes/ray %.2f packets-as-tris/ray %.2f" % (ctr[0] / len(fx.rays), ctr[1] / len(fx.rays)))
    # what is left: rays through an edge where 1/det by division (here) and by rcpps + one Newton step (reference) round apart
    assert st["hard"] == 0 and st["ties"] <= 8, st
    assert st["id_match"] >= 0.9997, st
    assert st["frac_t_within"] >= 0.9999, st


def test_reference_binding_glue_through_the_oracle(tmp_path):
    """The --render-gpu glue of oracle/ref_harness.cpp (the binding INTEGRATION.md describes: the reference's own objects,
    materials, lights and QBVH flattened into a miro_gpu_scene_desc) checked without a GPU: pointed at the CPU oracle's
    oracle_render instead of libmiro_gpu.so, the description it builds must render the reference's own image."""
    import os, subprocess
    exe = os.path.join(helpers.ROOT, "oracle", "_ref", "miro_ref")
    if not os.path.exists(exe):
        pytest.skip("reference binary not built (oracle/_ref)")
    fx = helpers.Fixture(helpers.fixture_path("c1_cornell"))
    sp = helpers.write_obj_scene(fx, str(tmp_path))
    out = tmp_path / "glue.f32"
    p = subprocess.run([exe, "--scene", sp, "--assets", str(tmp_path), "--render-gpu", str(out), "--gpu-lib", helpers.ORACLE_LIB],
                       stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    H, W = fx.radiance.shape[:2]
    img = np.fromfile(out, np.float32).reshape(H, W, 3)
    err = np.abs(img - fx.radiance).max(axis=2)
    assert (err <= 2e-3 * np.maximum(fx.radiance.max(axis=2), 1e-3) + 1e-4).mean() > 0.998
    # and the product refuses to pretend: without a CUDA device the same call through libmiro_gpu.so fails loudly
    import torch
    if not torch.cuda.is_available():
        lib = os.path.join(helpers.ROOT, "rendering-algorithms-raytracer_b200", "libmiro_gpu.so")
        q = subprocess.run([exe, "--scene", sp, "--assets", str(tmp_path), "--render-gpu", str(out), "--gpu-lib", lib], stderr=subprocess.PIPE, text=True)
        assert q.returncode != 0 and "no CUDA device" in q.stderr


@pytest.mark.parametrize("scene,block_tol", [("c5_mb_instances", 0.01), ("c7_foliage", 0.004), ("c9_texmaps", 0.03)])
def test_reference_binding_glue_covers_the_reference_object_model(scene, block_tol, tmp_path):
    """The same glue on scenes with MBObjects + ProxyObjects (shared bottom-level trees, mixed TriCache4 packets), alpha-mapped
    foliage, and normal / specular / reflect / refract maps with the tangent frame: the description built from the reference's
    own objects, rendered by the CPU oracle, against the reference's own render of the same process (same Scene::trace count,
    same image up to the random numbers).  Needs the reference's asset tree (this container only)."""
    import json, os, subprocess
    exe = os.path.join(helpers.ROOT, "oracle", "_ref", "miro_ref")
    assets = os.environ.get("MIRO_REFERENCE_ROOT", "/root/reference")
    if not os.path.exists(exe) or not os.path.isdir(os.path.join(assets, "Models")):
        pytest.skip("reference binary or asset tree not present")
    script = os.path.join(helpers.ROOT, "tests", "scenes", scene + ".miro")
    glue, ref = tmp_path / "glue.f32", tmp_path / "ref.f32"
    p = subprocess.run([exe, "--scene", script, "--assets", assets, "--render-gpu", str(glue), "--gpu-lib", helpers.ORACLE_LIB, "--render-float", str(ref)],
                       stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    ev = {e["event"]: e for e in (json.loads(l) for l in p.stderr.splitlines() if l.startswith("{"))}
    assert abs(ev["render_gpu"]["rays"] - ev["render_float"]["rays"]) <= 2e-3 * ev["render_float"]["rays"]
    W, H = ev["render_float"]["width"], ev["render_float"]["height"]
    a = np.fromfile(glue, np.float32).reshape(H, W, 3); b = np.fromfile(ref, np.float32).reshape(H, W, 3)
    blk = lambda x: np.minimum(x, 4).reshape(32, H // 32, 32, W // 32, 3).mean(axis=(1, 3))
    print(scene, "means", a.mean(), b.mean(), "block diff", np.abs(blk(a) - blk(b)).mean())
    assert abs(np.minimum(a, 4).mean() - np.minimum(b, 4).mean()) <= 0.01 * np.minimum(b, 4).mean()
    assert np.abs(blk(a) - blk(b)).mean() < block_tol


def test_host_instance_matrices_are_the_references_bit_for_bit(tmp_path):
    """ProxyObject::intersect moves the ray by ProxyMatrix::m_inverse (src/ProxyObject.cpp:78-79) and multiplyAndDivideByW scales
    the origin by recipps(w) (src/Matrix4x4.h:728-733).  For instanced hit distances to be the reference's, the host layer must hand
    the GPU exactly those numbers: rows 0..2 of the inverse as Matrix4x4::invert rounds it (host/miro_math.h invertedAsReference)
    and recipps(m44) from this host's SSE unit (referenceRecip) — checked against a dump from the reference binary itself."""
    import ctypes as C
    import os
    import reference_arm as ra
    if not ra.have_reference():
        pytest.skip("reference binary not built (oracle/_ref)")
    fx = helpers.Fixture(helpers.fixture_path("c5_mb_instances"))
    out = str(tmp_path / "inst.bin")
    ra.run_reference(fx, None, extra_args=["--dump-instances", out])
    ref = np.fromfile(out, np.float32).reshape(-1, 17)
    sc = fx.scene()
    d = sc.desc()
    inst = np.ctypeslib.as_array(C.cast(d.instances, C.POINTER(C.c_uint32)), shape=(d.n_instances, 16)).copy()
    assert d.n_instances >= len(ref) == 961
    seen = set()
    for row in inst:                         # an instance may appear as several records (sub-trees); all carry the proxy's ordinal
        o = int(row[13]); seen.add(o)
        assert row[:12].tobytes() == ref[o, :12].tobytes(), o          # rows 0..2 of m_inverse
        assert row[14:15].tobytes() == ref[o, 16:17].tobytes(), o      # recipps(m44)
    assert seen == set(range(961))
    assert (ref[:, 12:15] == 0).all()        # affine: w = m44 exactly
    assert (ref[:, 15] != 1.0).any()         # ... and m44 is not always 1: the reason w_recip exists
    sc.close()
