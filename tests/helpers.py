"""Shared test plumbing: golden fixtures -> scenes, the oracle library, comparison helpers."""
import ctypes as C
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FULL = os.path.join(ROOT, "oracle", "_ref", "fixtures")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libmiro_oracle.so")

import miro_b200 as mb  # noqa: E402
from miro_b200 import capi  # noqa: E402

REFHIT = np.dtype([("t", "f4"), ("a", "f4"), ("b", "f4"), ("mesh", "i4"), ("tri", "i4"), ("proxy", "i4")])


def fixture_path(scene, full=False):
    p = os.path.join(FULL if full else GOLDEN, scene + ".npz")
    return p if os.path.exists(p) else None


class Fixture:
    """A golden file: the reference's geometry, rays, hits and images for one scene script."""

    def __init__(self, path):
        z = np.load(path, allow_pickle=False)
        self.events = str(z["events"])
        self.script = str(z["script"])
        self.radiance = z["radiance"].astype(np.float32) if "radiance" in z.files else None
        self.radiance_converged = z["radiance_converged"].astype(np.float32) if "radiance_converged" in z.files else None
        if "overlay_of" in z.files:       # an overlay: its own script and reference images, geometry / textures of another fixture
            z = np.load(os.path.join(os.path.dirname(path), str(z["overlay_of"]) + ".npz"), allow_pickle=False)
            self.z = z
            self.names = [str(n) for n in z["mesh_names"]]
            self.rays = self.hits = self.ray_index = self.image8 = None
            return
        self.z = z
        self.names = [str(n) for n in z["mesh_names"]]
        self.rays = z["rays"]; self.hits = z["hits"]; self.ray_index = z["ray_index"]
        self.radiance = z["radiance"].astype(np.float32) if "radiance" in z.files else None
        self.image8 = z["image8"] if "image8" in z.files else None
        self.radiance_converged = z["radiance_converged"].astype(np.float32) if "radiance_converged" in z.files else None

    def mesh(self, k):
        g = lambda key: self.z[f"m{k}_{key}"]
        ti = g("ti")
        return dict(vertices=g("v"), vidx=g("vi").astype(np.uint32), normals=g("n"), nidx=g("ni").astype(np.uint32),
                    uvs=g("t") if len(ti) else None, tidx=ti.astype(np.uint32) if len(ti) else None)

    def scene(self, script_override=None, images=None):
        """Build a MiroScene from the fixture: the reference's own geometry, the same scene script."""
        sc = mb.MiroScene()
        for k, name in enumerate(self.names):
            sc.preload_mesh(name, **self.mesh(k))
        for key in self.z.files:                       # textures exactly as the reference's loaders decoded them
            if key.startswith("tex_"):
                sc.preload_image(key[4:], self.z[key], hdr=int(self.z["texkind_" + key[4:]]) == 3)
            elif key.startswith("texrgbe_"):
                b = self.z[key]; e = b[..., 3].astype(np.int32)
                tex = (b[..., :3].astype(np.float32) * np.ldexp(1.0, e - 136).astype(np.float32)[..., None]) * (e > 0)[..., None]
                sc.preload_image(key[8:], tex.astype(np.float32), hdr=True)
        for name, tex in (images or {}).items():
            sc.preload_image(name, tex)
        with tempfile.NamedTemporaryFile("w", suffix=".miro", delete=False) as f:
            f.write(script_override or self.script)
            path = f.name
        try:
            sc.load_script(path, "/nonexistent-asset-root")
        finally:
            os.unlink(path)
        return sc


_oracle = None


def oracle():
    """The CPU restatement (oracle/miro_oracle.c), compiled by __graft_entry__.build()."""
    global _oracle
    if _oracle is None:
        L = C.CDLL(ORACLE_LIB)
        L.oracle_trace_closest.argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.oracle_trace_closest.restype = C.c_int
        L.oracle_trace_any.argtypes = [C.POINTER(capi.SceneDesc), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.oracle_trace_any.restype = C.c_int
        _oracle = L
    return _oracle


def oracle_trace_closest(scene, rays):
    rays = np.ascontiguousarray(rays, mb.RAY_DTYPE)
    hits = np.empty(len(rays), mb.HIT_DTYPE)
    ctr = np.zeros(2, np.uint64)
    d = scene.desc()
    oracle().oracle_trace_closest(C.byref(d), rays.ctypes.data, len(rays), hits.ctypes.data, ctr.ctypes.data)
    return hits, ctr


def oracle_trace_any(scene, rays):
    rays = np.ascontiguousarray(rays, mb.RAY_DTYPE)
    bits = np.zeros((len(rays) + 31) // 32, np.uint32)
    ctr = np.zeros(2, np.uint64)
    d = scene.desc()
    oracle().oracle_trace_any(C.byref(d), rays.ctypes.data, len(rays), bits.ctypes.data, ctr.ctypes.data)
    return np.unpackbits(bits.view(np.uint8), bitorder="little")[:len(rays)].astype(bool)


def compare_hits(scene, hits, ref, t_rel=1e-5, edge_eps=1e-4, rays=None):
    """Compare product hits with reference-identity hits (mesh, tri, proxy, t).

    Returns a dict of statistics.  A primitive-id mismatch is an *edge case* ("tie") when
      * both sides hit at the same distance (|dt| <= t_rel * t): an edge / vertex shared by two triangles,
        where the winner depends on visiting order; or
      * either side's hit lies on a triangle edge (a barycentric weight within edge_eps of 0): the ray grazes
        a silhouette edge and the two intersection tests round differently (the reference's test is not
        watertight, so it also misses triangles through cracks).
    Everything else is a *hard* mismatch.
    """
    mesh, tri, proxy = scene.resolve_hits(hits) if hits.dtype == mb.HIT_DTYPE else (hits["mesh"], hits["tri"], hits["proxy"])
    r_hit = ref["mesh"] >= 0; g_hit = mesh >= 0
    same = (mesh == ref["mesh"]) & (tri == ref["tri"]) & (proxy == ref["proxy"])
    both = r_hit & g_hit
    dt = np.abs(hits["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-30)

    def on_edge(h, valid):
        w = np.minimum(np.minimum(h["a"], h["b"]), 1.0 - h["a"] - h["b"])
        return valid & (w <= edge_eps)
    tie = ~same & ((both & (dt <= t_rel)) | on_edge(hits, g_hit) | on_edge(ref, r_hit))
    hard = ~same & ~tie
    ok = same & both
    # instanced hits: the ray is moved into object space (ProxyObject.cpp:78-79), so t carries the rounding of
    # coordinates of magnitude |o|; for t << |o| "relative to t" is not attainable by ANY float32 implementation.
    # frac_t_within_pos measures |dt| against max(t, |o|) instead.
    pos = 1.0
    if rays is not None and ok.any():
        scale = np.maximum(np.abs(ref["t"]), np.linalg.norm(rays["o"], axis=1))
        pos = float((np.abs(hits["t"] - ref["t"])[ok] <= t_rel * scale[ok]).mean())
    return dict(frac_t_within_pos=pos, n=len(ref), id_match=float(same.mean()), ties=int(tie.sum()), hard=int(hard.sum()),
                hard_idx=np.nonzero(hard)[0], max_rel_t=float(dt[ok].max()) if ok.any() else 0.0,
                frac_t_within=float((dt[ok] <= t_rel).mean()) if ok.any() else 1.0,
                max_abs_a=float(np.abs(hits["a"] - ref["a"])[ok].max()) if ok.any() else 0.0,
                max_abs_b=float(np.abs(hits["b"] - ref["b"])[ok].max()) if ok.any() else 0.0,
                closer=int((~same & g_hit & ((hits["t"] < ref["t"]) | ~r_hit)).sum()))


def oracle_render(scene, params=None, camera=None, mask=None):
    """CPU restatement of Scene::raytraceImage (oracle/miro_oracle_shade.c).  Returns (float image [h,w,3], Scene::trace calls)."""
    L = oracle()
    L.oracle_render.argtypes = [C.POINTER(capi.SceneDesc), C.POINTER(capi.Camera), C.POINTER(capi.RenderParams), C.c_void_p, C.c_void_p]
    L.oracle_render.restype = C.c_uint64
    p = params or scene.render_params(); c = camera or scene.camera(); d = scene.desc()
    img = np.zeros((p.height, p.width, 3), np.float32)
    m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
    n = L.oracle_render(C.byref(d), C.byref(c), C.byref(p), img.ctypes.data, m.ctypes.data if m is not None else None)
    return img, int(n)


def write_obj_scene(fx, tmp):
    """Materialise the fixture's geometry as OBJ + script so the reference binary can load it with its own loader (positions,
    and normals / texture coordinates with their own index triples where the mesh has them)."""
    script = fx.script
    for k, name in enumerate(fx.names):
        m = fx.mesh(k)
        path = os.path.join(tmp, name + ".obj")
        has_n = len(m["normals"]) > 0 and not np.array_equal(m["nidx"], np.arange(3 * len(m["vidx"]), dtype=np.uint32).reshape(-1, 3))
        has_t = m["uvs"] is not None
        with open(path, "w") as f:
            for v in m["vertices"]:
                f.write("v %.9g %.9g %.9g\n" % tuple(v))
            if has_t:
                for t in m["uvs"]:
                    f.write("vt %.9g %.9g\n" % tuple(t))
            if has_n:
                for n in m["normals"]:
                    f.write("vn %.9g %.9g %.9g\n" % tuple(n))
            for i, t in enumerate(m["vidx"]):
                if has_n and has_t:
                    f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % tuple(x for j in range(3) for x in (t[j] + 1, m["tidx"][i][j] + 1, m["nidx"][i][j] + 1)))
                elif has_n:
                    f.write("f %d//%d %d//%d %d//%d\n" % tuple(x for j in range(3) for x in (t[j] + 1, m["nidx"][i][j] + 1)))
                elif has_t:
                    f.write("f %d/%d %d/%d %d/%d\n" % tuple(x for j in range(3) for x in (t[j] + 1, m["tidx"][i][j] + 1)))
                else:
                    f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))
        lines = []
        for line in script.splitlines():
            tok = line.split()
            if len(tok) >= 3 and tok[0] == "mesh" and tok[1] == name:
                line = "mesh %s %s" % (name, path) + ("" if len(tok) == 3 else " " + " ".join(tok[3:]))
            lines.append(line)
        script = "\n".join(lines) + "\n"
    sp = os.path.join(tmp, "scene.miro")
    open(sp, "w").write(script)
    return sp


class ReferenceTreeScene:
    """The reference's OWN QBVH (dumped by oracle/ref_harness.cpp --dump-qbvh: QBVH_Node bounds / children / TriCache4 lanes,
    src/BVH.h:89-104) flattened 1:1 into the miro_gpu_node / leaf-ordered triangle layout of include/miro_gpu.h — what the
    flattenQ() glue of INTEGRATION.md produces inside Miro — held as numpy arrays behind a miro_gpu_scene_desc."""

    def __init__(self, fx):
        z = fx.z
        bounds, child, leaves = z["qbvh_bounds"], z["qbvh_child"], z["qbvh_leaves"]
        meshes = [fx.mesh(k) for k in range(len(fx.names))]
        nbase = np.cumsum([0] + [len(m["normals"]) for m in meshes])
        tris, prims = [], []
        leaf_ref = np.zeros(len(leaves), np.int64)
        for li, lanes in enumerate(leaves):
            first, count = len(tris), 0
            for mesh, tri, kind in lanes:
                if mesh < 0:
                    continue
                assert kind == 0, "only plain Objects in these fixtures"
                m = meshes[mesh]
                tris.append(m["vertices"][m["vidx"][tri]])
                prims.append(list(nbase[mesh] + m["nidx"][tri]) + [0xffffffff] * 3 + [0, mesh, tri, 0, 0, 0])
                count += 1
            leaf_ref[li] = (0x80000000 | ((count - 1) << 26) | first) - (1 << 32)      # MIRO_GPU_LEAF(KIND_TRI, first, count)
        self.nodes = np.zeros((len(bounds), 32), np.float32)
        self.nodes[:, :24] = bounds
        ch = self.nodes[:, 24:28].view(np.int32)
        for i in range(len(child)):
            for k in range(4):
                c = int(child[i, k])
                ch[i, k] = capi.CHILD_EMPTY if c == -(1 << 31) else (c if c >= 0 else leaf_ref[~c])
        t = np.zeros((len(tris), 3, 4), np.float32); t[:, :, :3] = np.array(tris, np.float32)
        self.tris = t
        self.prims = np.array(prims, np.uint32)
        self.normals = np.concatenate([m["normals"] for m in meshes]).astype(np.float32)
        self.material = capi.Material(); self.material.kind = 0
        self.material.kd[:] = [0.8, 0.8, 0.8]; self.material.spec_gloss = 1.0; self.material.color_map = -1; self.material.alpha_map = -1
        self.material.normal_map = self.material.specular_map = self.material.reflect_map = self.material.refract_map = -1
        d = capi.SceneDesc(); d.abi_version = capi.lib().miro_gpu_abi_version()
        d.nodes = self.nodes.ctypes.data_as(C.POINTER(capi.Node)); d.n_nodes = len(self.nodes); d.root = 0
        d.tris = self.tris.ctypes.data_as(C.POINTER(capi.Tri)); d.n_tris = len(self.tris)
        d.prims = self.prims.ctypes.data_as(C.POINTER(capi.Prim))
        d.normals = self.normals.ctypes.data_as(C.POINTER(C.c_float)); d.n_normals = len(self.normals)
        d.materials = C.pointer(self.material); d.n_materials = 1
        d.env_map = -1; d.env_exposure = 1.0
        self.d = d
        self.mesh_of = self.prims[:, 7].astype(np.int64); self.tri_of = self.prims[:, 8].astype(np.int64)
        self.ctx = None

    def desc(self):
        return self.d

    def resolve_hits(self, hits):
        prim = hits["prim"]; ok = prim >= 0
        mesh = np.full(len(hits), -1, np.int64); tri = np.full(len(hits), -1, np.int64)
        mesh[ok] = self.mesh_of[prim[ok]]; tri[ok] = self.tri_of[prim[ok]]
        return mesh, tri, np.full(len(hits), -1, np.int64)

    def attach(self, device=0):
        L = capi.lib()
        ctx = C.c_void_p()
        rc = L.miro_gpu_create(C.byref(ctx), device)
        if rc:
            raise mb.MiroError("miro_gpu_create: " + L.miro_gpu_last_error(None).decode())
        self.ctx, self.L = ctx, L
        rc = L.miro_gpu_upload_scene(ctx, C.byref(self.d))
        if rc:
            raise mb.MiroError("miro_gpu_upload_scene: " + L.miro_gpu_last_error(ctx).decode())
        return self

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, mb.RAY_DTYPE); hits = np.empty(len(rays), mb.HIT_DTYPE)
        rc = self.L.miro_gpu_trace_closest(self.ctx, rays.ctypes.data, len(rays), hits.ctypes.data)
        if rc:
            raise mb.MiroError(self.L.miro_gpu_last_error(self.ctx).decode())
        return hits

    def close(self):
        if self.ctx:
            self.L.miro_gpu_destroy(self.ctx); self.ctx = None


def same_primitive(sc, a, b):
    """Boolean array: hit records a and b name the same SOURCE primitive (or both miss).  Spatial splits of the host builder
    store a triangle once per leaf that references it, so two traversal orders may report different copies of one triangle;
    identity is (mesh ordinal, triangle index), as it is against the reference."""
    mesh_of, tri_of, inst_ord = sc.prim_table()
    def ident(h):
        p = h["prim"].astype(np.int64); ok = p >= 0
        out = np.full(len(p), -1, np.int64)
        out[ok] = mesh_of[p[ok]] * (1 << 32) + tri_of[p[ok]]
        return out
    def instance(h):        # an instance is entered through several records (braiding): compare ordinals, not record indices
        i = h["inst"].astype(np.int64); ok = i >= 0
        out = np.full(len(i), -1, np.int64)
        out[ok] = inst_ord[i[ok]]
        return out
    return (ident(a) == ident(b)) & (instance(a) == instance(b))
