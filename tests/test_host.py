"""Host layer (C++): BVH build + flatten invariants, scene scripts, loaders, Image::Map.  CPU only."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import helpers
import miro_b200 as mb
from miro_b200 import capi


def flat(sc):
    d = sc.desc()
    nodes = np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_float)), shape=(d.n_nodes, 32)).copy() if d.n_nodes else np.zeros((0, 32), np.float32)
    child = nodes[:, 24:28].view(np.int32)
    tris = np.ctypeslib.as_array(C.cast(d.tris, C.POINTER(C.c_float)), shape=(d.n_tris, 12)).copy() if d.n_tris else np.zeros((0, 12), np.float32)
    return d, nodes, child, tris


def leaf_fields(ref):
    u = int(ref) & 0xffffffff
    return (u >> 29) & 3, ((u >> 26) & 7) + 1, u & ((1 << 26) - 1)


def walk(d, nodes, child, tris, ref, lo, hi, seen, depth=0, inside=True):
    """Every child box is inside its parent's box.  Object splits only (inside=True): every triangle of a leaf lies inside the
    leaf's box.  With spatial splits a leaf holds the part of a triangle that lies inside its box: the triangle's box must at
    least overlap it."""
    if ref == capi.CHILD_EMPTY:
        return 0
    if ref < 0:
        kind, count, first = leaf_fields(ref)
        assert 1 <= count <= 4
        if kind == capi.KIND_TRI:
            for i in range(first, first + count):
                seen[i] += 1
                v = tris[i].reshape(3, 4)[:, :3]
                if inside:
                    assert (v.min(0) >= lo - 1e-6).all() and (v.max(0) <= hi + 1e-6).all()
                else:
                    assert (v.max(0) >= lo - 1e-6).all() and (v.min(0) <= hi + 1e-6).all()
        return depth
    n = nodes[ref]
    best = depth
    for i in range(4):
        c = int(child[ref, i])
        if c == capi.CHILD_EMPTY:
            continue
        clo = np.array([n[0 + i], n[4 + i], n[8 + i]]); chi = np.array([n[12 + i], n[16 + i], n[20 + i]])
        assert (clo <= chi).all()
        assert (clo >= lo - 1e-6).all() and (chi <= hi + 1e-6).all()
        best = max(best, walk(d, nodes, child, tris, c, clo, chi, seen, depth + 1, inside))
    return best


@pytest.mark.parametrize("scene", ["c1_cornell", "c2_explosion"])
@pytest.mark.parametrize("spatial", [False, True])
def test_flatten_invariants(scene, spatial, monkeypatch):
    """spatial=False: object splits only (MIRO_BVH_SPATIAL=0) — the flattening is a permutation of the source triangles.
    spatial=True (the default build): spatial splits may reference a triangle from several leaves (at most 2x the source
    count); every source triangle is still referenced, and every copy carries its (mesh, tri) identity."""
    if not spatial:
        monkeypatch.setenv("MIRO_BVH_SPATIAL", "0")
    else:
        monkeypatch.setenv("MIRO_BVH_SPATIAL_MIN", "0")       # also on the 36-triangle Cornell box (the default skips tiny scenes)
    fx = helpers.Fixture(helpers.fixture_path(scene))
    sc = fx.scene()
    d, nodes, child, tris = flat(sc)
    n_src = sum(len(fx.mesh(k)["vidx"]) for k in range(len(fx.names)))
    assert d.n_tris == n_src if not spatial else n_src <= d.n_tris <= 2 * n_src
    seen = np.zeros(d.n_tris, np.int64)
    big = np.float32(3e38)
    depth = walk(d, nodes, child, tris, d.root, -np.full(3, big), np.full(3, big), seen, inside=not spatial)
    assert (seen == 1).all()          # every leaf slot belongs to exactly one leaf
    st = sc.bvh_stats()
    assert st["nodes"] == d.n_nodes and st["max_depth"] >= depth >= 1
    # (mesh, tri) identity maps ONTO the source triangles (a bijection without spatial splits)
    mesh_of, tri_of, _ = sc.prim_table()
    assert len(set(zip(mesh_of.tolist(), tri_of.tolist()))) == n_src
    # leaf-ordered vertices equal the source mesh's vertices
    m = fx.mesh(0)
    k = np.nonzero(mesh_of == 0)[0][:500]
    src = m["vertices"][m["vidx"][tri_of[k]]]
    assert np.array_equal(tris[k].reshape(-1, 3, 4)[:, :, :3], src)
    sc.close()


def test_parallel_build_is_deterministic_and_valid(monkeypatch):
    """Sub-trees of >= MIRO_BVH_PARALLEL_MIN references are built as concurrent tasks; whether a node's children run concurrently
    depends on the node alone, so two builds (any thread timing) give byte-identical node and triangle arrays."""
    monkeypatch.setenv("MIRO_BVH_PARALLEL_MIN", "1024")       # many tasks
    fx = helpers.Fixture(helpers.fixture_path("c2_explosion"))
    builds = []
    for _ in range(2):
        sc = fx.scene()
        d, nodes, child, tris = flat(sc)
        builds.append((d.root, nodes.copy(), child.copy(), tris.copy(), sc.bvh_stats()))
        seen = np.zeros(d.n_tris, np.int64)
        big = np.float32(3e38)
        walk(d, nodes, child, tris, d.root, -np.full(3, big), np.full(3, big), seen, inside=False)
        assert (seen == 1).all()
        sc.close()
    a, b = builds
    assert a[0] == b[0] and a[4] == b[4]
    u = lambda x: x.view(np.uint32)      # compare bit patterns: child references read as floats are NaNs
    assert np.array_equal(u(a[1]), u(b[1])) and np.array_equal(a[2], b[2]) and np.array_equal(u(a[3]), u(b[3]))


def _script_scene(text, meshes=None):
    sc = mb.MiroScene()
    for name, (v, f) in (meshes or {}).items():
        sc.preload_mesh(name, v, f)
    with tempfile.NamedTemporaryFile("w", suffix=".miro", delete=False) as fh:
        fh.write(text)
    try:
        sc.load_script(fh.name, "/nonexistent")
    finally:
        os.unlink(fh.name)
    return sc


QUAD = (np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32))


def test_script_errors_are_reported():
    for bad, frag in [("bogus 1 2 3\n", "unknown command"), ("object nomesh nomat\n", "unknown mesh"),
                      ("material m phong\n", "unknown material kind"), ("light dome power 1\n", "without tex")]:
        with pytest.raises(mb.MiroError) as e:
            _script_scene(bad)
        assert frag in str(e.value)


def test_single_leaf_scene_and_camera_defaults():
    sc = _script_scene("image 64 32\ncamera eye 0 0 3 lookat 0 0 0 fov 40\nmaterial m lambert kd 1 0 0\nmesh q q.obj\nobject q m\n", {"q": QUAD})
    d, nodes, child, tris = flat(sc)
    assert d.n_nodes == 0 and d.root < 0           # whole scene is one leaf (src/BVH.cpp:118-132)
    assert leaf_fields(d.root) == (capi.KIND_TRI, 2, 0)
    cam = sc.camera(); p = sc.render_params()
    assert (p.width, p.height) == (64, 32) and p.num_paths == 1 and p.max_bounces == 10 and p.seed == 3163513
    assert np.allclose(cam.view_dir[:], [0, 0, -1]) and abs(cam.fov_deg - 40) < 1e-6
    assert abs(cam.shutter_speed - 1e-3) < 1e-9 and cam.aperture == 0.0
    sc.close()


def test_motion_blur_and_instances_flatten():
    v2 = QUAD[0] + np.float32([0, 0, 1])
    ident = "1 0 0 0  0 1 0 0  0 0 1 0  0 0 0 1"
    shifted = "2 0 0 5  0 2 0 0  0 0 2 0  0 0 0 1"
    sc = _script_scene("material m blinn kd .5 .5 .5\nmesh a a.obj\nmesh b b.obj\nmbobject a b m\nblas g a m\n"
                       f"instance g {ident}\ninstance g {shifted}\n", {"a": QUAD, "b": (v2, QUAD[1])})
    d = sc.desc()
    assert d.n_mbtris == 2 and d.n_instances == 2 and d.n_tris == 2
    inst = np.ctypeslib.as_array(C.cast(d.instances, C.POINTER(C.c_float)), shape=(2, 16)).copy()
    ords = inst[:, 13].view(np.int32) if False else np.ctypeslib.as_array(C.cast(d.instances, C.POINTER(C.c_uint32)), shape=(2, 16))[:, 13]
    k = int(np.nonzero(ords == 1)[0][0])
    assert np.allclose(inst[k, :12].reshape(3, 4), [[.5, 0, 0, -2.5], [0, .5, 0, 0], [0, 0, .5, 0]])   # rows of M^-1
    nx = np.ctypeslib.as_array(d.inst_normal_xform, shape=(2, 9))[k].reshape(3, 3)
    assert np.allclose(nx, np.eye(3) * .5)                                                                 # (M^-1)^T
    mb_ = np.ctypeslib.as_array(C.cast(d.mbtris, C.POINTER(C.c_float)), shape=(2, 24))
    assert np.allclose(mb_[:, 12:].reshape(2, 3, 4)[:, :, 2], 1.0) and np.allclose(mb_[:, :12].reshape(2, 3, 4)[:, :, 2], 0.0)
    sc.close()


def test_specular_material_fields_travel_in_the_desc():
    # Blinn::setIor(ior, i) sets ONE entry (src/Blinn.h:38); the script's `ior` sets all three, `ior_i` one
    sc = _script_scene("material m blinn kd .5 .5 .5 reflect 0.5 refract 0.25 gloss 0.9 ior_i 0 2.2 translucency 0.5\nmesh a a.obj\nobject a m\n", {"a": QUAD})
    d = sc.desc()
    m = d.materials[0]
    assert abs(m.reflect_amt - 0.5) < 1e-7 and abs(m.refract_amt - 0.25) < 1e-7 and abs(m.spec_gloss - 0.9) < 1e-7
    assert abs(m.ior[0] - 2.2) < 1e-6 and abs(m.ior[1] - 1.5) < 1e-6 and abs(m.ior[2] - 1.5) < 1e-6 and m.disperse == 0
    assert abs(m.translucency - 0.5) < 1e-7    # travels in the desc; miro_gpu_upload_scene refuses it (GPU test)
    sc.close()


def test_light_shadow_flags_travel_in_the_desc():
    # Light::setCastShadows / setFastShadows (src/Light.h:22-24): script keys `shadows`, `fastshadows`; defaults cast + fast
    sc = _script_scene("material m lambert kd .5 .5 .5\nlight point pos 0 1 0 power 1\nlight point pos 0 2 0 power 1 shadows 0 fastshadows 0\n"
                       "light rect v1 0 3 0 v2 1 3 0 v3 0 3 1 power 2 samples 3 fastshadows 0\nmesh a a.obj\nobject a m\n", {"a": QUAD})
    d = sc.desc()
    assert d.n_lights == 3
    assert (d.lights[0].cast_shadows, d.lights[0].full_shadows) == (1, 0)
    assert (d.lights[1].cast_shadows, d.lights[1].full_shadows) == (0, 1)
    assert (d.lights[2].kind, d.lights[2].cast_shadows, d.lights[2].full_shadows, d.lights[2].num_samples) == (1, 1, 1, 3)
    sc.close()


def test_tangent_frame_and_texture_maps_travel_in_the_desc(tmp_path):
    """TriangleMesh::preCalc (src/TriangleMesh.cpp:107-150): per NORMAL index, the reference's tangent made orthogonal to the
    normal, bitangent = cross(tangent, normal); a triangle with a degenerate uv mapping writes nothing; meshes without uvs
    get zeros (src/Ray.cpp:44-45).  Material map indices (src/Blinn.cpp:120-142) travel; Lambert ignores them."""
    obj = tmp_path / "t.obj"
    # quad in the xy plane, u along +x, v along +y, two normals; third triangle has a degenerate uv mapping and its own normal
    obj.write_text("v 0 0 0\nv 2 0 0\nv 2 2 0\nv 0 2 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\nvn 0 0 1\nvn 0 1 0\n"
                   "f 1/1/1 2/2/1 3/3/2\nf 1/1/1 3/3/2 4/4/2\nf 1/1/3 2/1/3 3/1/3\n")
    import struct
    tga = tmp_path / "x.tga"
    tga.write_bytes(struct.pack("<BBBHHBHHHHBB", 0, 0, 2, 0, 0, 0, 0, 0, 2, 2, 24, 0) + bytes([128] * 12))
    script = tmp_path / "s.miro"
    script.write_text("image 8 4\ntexture x x.tga\nmaterial m blinn kd .5 .5 .5 normalmap x specularmap x reflectmap x refractmap x\n"
                      "material l lambert colormap x\nmesh t t.obj\nobject t m\n")
    sc = mb.MiroScene(); sc.load_script(script, tmp_path)
    d = sc.desc()
    assert d.n_normals == 3 and bool(d.tangents) and bool(d.bitangents)
    T = np.ctypeslib.as_array(d.tangents, shape=(3, 3)); B = np.ctypeslib.as_array(d.bitangents, shape=(3, 3))
    # the reference's tangent is (AB * -du2 + AC * dv1) / (dv1 du2 - du1 dv2) — not the textbook d(position)/du; reproduced, not
    # corrected: triangle 1 gives (1,0,0), triangle 2 gives (0,-1,0) and, coming later, overwrites both normal indices
    assert np.allclose(T[:2], [[0, -1, 0], [0, -1, 0]], atol=1e-6)
    assert np.allclose(B[:2], [[-1, 0, 0], [-1, 0, 0]], atol=1e-6)          # cross(T, N)
    assert np.allclose(T[2], 0) and np.allclose(B[2], 0)                      # degenerate uv mapping: never written
    m = d.materials[0]
    assert m.kind == 1 and m.normal_map == m.specular_map == m.reflect_map == m.refract_map == 0 and m.color_map == -1
    sc.close()
    script.write_text("image 8 4\ntexture x x.tga\nmaterial l lambert colormap x\nmesh t t.obj\nobject t l\n")
    sc = mb.MiroScene(); sc.load_script(script, tmp_path)
    m = sc.desc().materials[0]
    assert m.kind == 0 and m.color_map == 0 and m.normal_map == m.specular_map == m.reflect_map == m.refract_map == -1
    sc.close()
    # no texture coordinates: zero tangents
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    sc = mb.MiroScene(); sc.load_script(script, tmp_path)
    d = sc.desc()
    assert not d.tangents or np.allclose(np.ctypeslib.as_array(d.tangents, shape=(d.n_normals, 3)), 0)
    sc.close()


def test_obj_loader_and_ppm_writer(tmp_path):
    obj = tmp_path / "t.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1\nf 2/2/1 4/3/1 3/3/1\n")
    script = tmp_path / "s.miro"
    script.write_text("image 8 4\nmaterial m lambert\nmesh t t.obj\nobject t m\n")
    sc = mb.MiroScene(); sc.load_script(script, tmp_path)
    d = sc.desc()
    assert d.n_tris == 2 and d.n_normals == 1 and d.n_uvs == 3
    prims = np.ctypeslib.as_array(C.cast(d.prims, C.POINTER(C.c_uint32)), shape=(2, 12))
    assert (prims[:, 0:3] == 0).all() and prims[:, 3:6].max() == 2
    sc.close()
    # flat normals when the file has none (src/TriangleMeshLoad.cpp:194-206)
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    sc = mb.MiroScene(); sc.load_script(script, tmp_path)
    d = sc.desc()
    assert np.allclose(np.ctypeslib.as_array(d.normals, shape=(3,)), [0, 0, 1])
    prims = np.ctypeslib.as_array(C.cast(d.prims, C.POINTER(C.c_uint32)), shape=(1, 12))
    assert (prims[0, 3:6] == 0xffffffff).all()
    sc.close()
