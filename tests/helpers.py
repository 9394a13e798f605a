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
        self.z = z
        self.names = [str(n) for n in z["mesh_names"]]
        self.script = str(z["script"])
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
    """Materialise the fixture's geometry as OBJ + script so the reference binary can load it with its own loader."""
    script = fx.script
    for k, name in enumerate(fx.names):
        m = fx.mesh(k)
        path = os.path.join(tmp, name + ".obj")
        with open(path, "w") as f:
            for v in m["vertices"]:
                f.write("v %.9g %.9g %.9g\n" % tuple(v))
            for t in m["vidx"]:
                f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))
        lines = []
        for line in script.splitlines():
            tok = line.split()
            if len(tok) >= 3 and tok[0] == "mesh" and tok[1] == name:
                line = "mesh %s %s" % (name, path)
            lines.append(line)
        script = "\n".join(lines) + "\n"
    sp = os.path.join(tmp, "scene.miro")
    open(sp, "w").write(script)
    return sp
