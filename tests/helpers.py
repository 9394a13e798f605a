"""Shared test plumbing: golden fixtures -> scenes, the oracle library, comparison helpers."""
import ctypes as C
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libmiro_oracle.so")

import miro_b200 as mb  # noqa: E402
from miro_b200 import capi  # noqa: E402
import reference_arm as ra  # noqa: E402
from reference_arm import GOLDEN, FULL, REFHIT, fixture_path, write_obj_scene  # noqa: E402,F401


class Fixture(ra.FixtureData):
    """A golden file (reference_arm.FixtureData) that can also build the product's scene from it."""

    def scene(self, script_override=None, images=None):
        """Build a MiroScene from the fixture: the reference's own geometry, the same scene script."""
        sc = mb.MiroScene()
        for k, name in enumerate(self.names):
            sc.preload_mesh(name, **self.mesh(k))
        for name, (tex, hdr) in self.textures().items():      # textures exactly as the reference's loaders decoded them
            sc.preload_image(name, tex, hdr=hdr)
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
    """Compare product hits with reference-identity hits (mesh, tri, proxy, t): reference_arm.compare_with_reference after
    resolving the product's primitive / instance indices to the reference's identities.  Returns its dict of statistics
    (ties split by cause: ties_equal_t / ties_own_edge / ties_ref_edge; everything else is `hard`)."""
    mesh, tri, proxy = scene.resolve_hits(hits) if hits.dtype == mb.HIT_DTYPE else (hits["mesh"], hits["tri"], hits["proxy"])
    return ra.compare_with_reference(mesh, tri, proxy, hits["t"], hits["a"], hits["b"], ref, rays=rays, t_rel=t_rel, edge_eps=edge_eps)


def adjudicate_hard(fx, scene, hits, ref, st, rays, script=None):
    """The mismatches compare_hits could not class as ties (st["hard_idx"]) re-computed in float64 on both sides' triangles
    (reference_arm.adjudicate_mismatches).  Returns the class counts; product_missed + unexplained are the true failures."""
    geom = ra.SceneGeometry(script or fx.script, {n: fx.mesh(k) for k, n in enumerate(fx.names)})
    mesh, tri, proxy = scene.resolve_hits(hits)
    adj = ra.adjudicate_mismatches(geom, rays, st["hard_idx"], mesh, tri, proxy, ref)
    for i in adj["hard_idx"][:8]:      # what a failing assertion needs to show
        g = geom.intersect(rays[i], int(mesh[i]), int(tri[i]), int(proxy[i])) if mesh[i] >= 0 else None
        f = geom.intersect(rays[i], int(ref["mesh"][i]), int(ref["tri"][i]), int(ref["proxy"][i])) if ref["mesh"][i] >= 0 else None
        print("unresolved ray", i, rays[i], "product", (mesh[i], tri[i], proxy[i]), hits[i], "float64 (t, edge distance, reach)", g,
              "reference", ref[i], "float64", f)
    return adj


def reference_hits(fx, rays, threads=1):
    """The reference's own Scene::trace over `rays` on the fixture's scene, run on the spot (oracle/_ref/miro_ref travels to
    the GPU box); the fixture's geometry is REPLACED by what that run loaded (FixtureData.use_meshes), so call this before
    fx.scene().  Single-threaded by default: with several OpenMP threads the reference's traversals corrupt each other
    (QBVH_Node::boxHit is a mutable member of the shared node)."""
    _, hits, meshes = ra.run_reference(fx, rays, threads=threads, want_hits=True, dump_meshes=True)
    fx.use_meshes(meshes)       # from here on fx.scene() builds the geometry the reference traced (its loader is not idempotent)
    return hits


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
