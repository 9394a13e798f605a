"""The reference side of tests and benchmarks: golden fixtures, ray batches and the reference binary (oracle/_ref/miro_ref).

Nothing in this module imports the product (no miro_b200, no libmiro_gpu.so): `bench.py --impl reference` is built from it
alone, so the reference arm never maps the library it is compared with.  Test / benchmark infrastructure, not product.
"""
import json
import os
import re
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FULL = os.path.join(ROOT, "oracle", "_ref", "fixtures")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "miro_ref")

# miro_gpu_ray / the harness's RayRec (48 bytes) and the harness's hit record (oracle/ref_harness.cpp RefHit)
RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32),
                      ("time", np.float32), ("flags", np.uint32), ("user", np.uint32, 2)])
REFHIT = np.dtype([("t", "f4"), ("a", "f4"), ("b", "f4"), ("mesh", "i4"), ("tri", "i4"), ("proxy", "i4")])


def fixture_path(scene, full=False):
    p = os.path.join(FULL if full else GOLDEN, scene + ".npz")
    return p if os.path.exists(p) else None


class FixtureData:
    """A golden file: the reference's geometry, rays, hits and images for one scene script (numpy only)."""

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
        self.image8 = z["image8"] if "image8" in z.files else None

    def mesh(self, k):
        if getattr(self, "_mesh_override", None):
            return self._mesh_override[self.names[k]]
        g = lambda key: self.z[f"m{k}_{key}"]
        ti = g("ti")
        return dict(vertices=g("v"), vidx=g("vi").astype(np.uint32), normals=g("n"), nidx=g("ni").astype(np.uint32),
                    uvs=g("t") if len(ti) else None, tidx=ti.astype(np.uint32) if len(ti) else None)

    def use_meshes(self, meshes):
        """Replace the geometry by what a reference run dumped (run_reference(..., dump_meshes=True)).  Needed whenever the
        reference loads this fixture's geometry AGAIN from the OBJ files write_obj_scene writes: its loader is not idempotent
        (every vertex goes through Matrix4x4::multiplyAndDivideByW with an approximate reciprocal, normals through an approximate
        rsqrt: src/TriangleMeshLoad.cpp:143-172), so the re-loaded vertices differ from the fixture's by an ulp — enough to move
        t by 1e-4 relative on sliver triangles.  Both sides must trace the geometry the reference actually holds."""
        self._mesh_override = meshes
        return self

    def bounds(self):
        allv = np.concatenate([self.mesh(k)["vertices"] for k in range(len(self.names))])
        return allv.min(0), allv.max(0)

    def n_triangles(self):
        return int(sum(len(self.mesh(k)["vidx"]) for k in range(len(self.names))))

    def textures(self):
        """name -> (float texels [h, w, c], is_hdr), exactly as the reference's loaders decoded them."""
        out = {}
        for key in self.z.files:
            if key.startswith("tex_"):
                out[key[4:]] = (self.z[key], int(self.z["texkind_" + key[4:]]) == 3)
            elif key.startswith("texrgbe_"):
                b = self.z[key]; e = b[..., 3].astype(np.int32)
                tex = (b[..., :3].astype(np.float32) * np.ldexp(1.0, e - 136).astype(np.float32)[..., None]) * (e > 0)[..., None]
                out[key[8:]] = (tex.astype(np.float32), True)
        return out


def write_obj_scene(fx, tmp, script=None):
    """Materialise the fixture's geometry as OBJ + script so the reference binary can load it with its own loader (positions,
    and normals / texture coordinates with their own index triples where the mesh has them)."""
    script = script or fx.script
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
                line = "mesh %s %s" % (name, path)      # without the script's `ctm`: the fixture holds the vertices AFTER the reference's load, transform applied
            lines.append(line)
        script = "\n".join(lines) + "\n"
    sp = os.path.join(tmp, "scene.miro")
    open(sp, "w").write(script)
    return sp


# ---------------------------------------------------------------------------------------------- ray batches (numpy)
def script_camera(script):
    """The `camera` and `image` lines of a scene script -> dict(eye, view_dir, up, fov_deg, width, height), with the
    normalisations of the host layer's Camera setters (the reference's Camera::setLookAt / setUp, src/Camera.h)."""
    cam = dict(eye=np.zeros(3, np.float32), view_dir=np.array([0, 0, -1], np.float32), up=np.array([0, 1, 0], np.float32), fov_deg=45.0, width=0, height=0)
    lookat = None
    for line in script.splitlines():
        tok = line.split("#")[0].split()
        if not tok:
            continue
        if tok[0] == "image":
            cam["width"], cam["height"] = int(tok[1]), int(tok[2])
        elif tok[0] == "camera":
            i = 1
            while i < len(tok):
                k = tok[i]
                if k in ("eye", "lookat", "viewdir", "up"):
                    v = np.array([float(x) for x in tok[i + 1:i + 4]], np.float32); i += 4
                    if k == "eye":
                        cam["eye"] = v
                    elif k == "lookat":
                        lookat = v
                    elif k == "viewdir":
                        cam["view_dir"] = v; lookat = None
                    else:
                        cam["up"] = v
                else:
                    if k == "fov":
                        cam["fov_deg"] = float(tok[i + 1])
                    i += 2
    if lookat is not None:
        cam["view_dir"] = (lookat - cam["eye"]).astype(np.float32)
    cam["view_dir"] = (cam["view_dir"] / np.linalg.norm(cam["view_dir"])).astype(np.float32)
    cam["up"] = (cam["up"] / np.linalg.norm(cam["up"])).astype(np.float32)
    return cam


def primary_rays(cam, w, h):
    """Pinhole rays at pixel centres (Camera::eyeRayAdaptive with 0.5 offsets, src/Camera.cpp:116-158), row 0 = bottom.
    cam: dict as script_camera() returns (or any object with eye / view_dir / up / fov_deg members)."""
    g = (lambda k: cam[k]) if isinstance(cam, dict) else (lambda k: getattr(cam, k))
    eye = np.array(g("eye")[:], np.float32); vd = np.array(g("view_dir")[:], np.float32); up = np.array(g("up")[:], np.float32)
    wv = -vd / np.linalg.norm(vd); u = np.cross(up, wv); u /= np.linalg.norm(u); v = np.cross(wv, u)
    top = np.tan(np.float32(g("fov_deg")) * np.float32(3.1415926 / 360.0)); right = top * w / h
    xs = (-right + 2 * right * (np.arange(w, dtype=np.float32) + 0.5) / w)[None, :, None]
    ys = (-top + 2 * top * (np.arange(h, dtype=np.float32) + 0.5) / h)[:, None, None]
    d = xs * u[None, None, :] + ys * v[None, None, :] - wv[None, None, :]
    d = (d / np.linalg.norm(d, axis=2, keepdims=True)).reshape(-1, 3).astype(np.float32)
    r = np.zeros(w * h, RAY_DTYPE)
    r["o"] = eye; r["d"] = d; r["tmin"] = 1e-3; r["tmax"] = 1e12
    return r


def incoherent_rays(lo, hi, n, seed, times=False):
    """Seeded incoherent batch: origins ~U(scene AABB inflated 10 %), directions ~U(S^2) (SURVEY 8d, C2 ii)."""
    rng = np.random.default_rng(seed)
    c, e = 0.5 * (lo + hi), 0.55 * (hi - lo) + 1e-3
    r = np.zeros(n, RAY_DTYPE)
    r["o"] = (c + e * rng.uniform(-1, 1, (n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["d"] = d.astype(np.float32); r["tmin"] = 1e-3; r["tmax"] = 1e12
    if times:
        r["time"] = rng.uniform(0, 1, n).astype(np.float32)
    return r


def shadow_rays(src, hit_t, hit_mask, light):
    """Shadow rays as PointLight::sampleLight casts them (src/PointLight.cpp:20-48): from the hit point (or, for a
    miss, from the ray origin) towards the light, tmin 1e-3, tmax = distance."""
    t = np.where(hit_mask, hit_t, 0.0).astype(np.float32)
    p = src["o"] + t[:, None] * src["d"]
    L = np.asarray(light, np.float32)[None, :] - p
    dist = np.linalg.norm(L, axis=1).astype(np.float32)
    r = np.zeros(len(src), RAY_DTYPE)
    r["o"] = p; r["d"] = (L / np.maximum(dist, 1e-20)[:, None]).astype(np.float32); r["tmin"] = 1e-3; r["tmax"] = dist
    r["time"] = src["time"]
    return r


# ---------------------------------------------------------------------------------------------- the reference binary
def have_reference():
    return os.path.exists(REF_BIN)


def read_mesh_dump(path):
    """One file of the harness's --dump-meshes: (ordinal, mesh dict as FixtureData.mesh() returns it)."""
    b = open(path, "rb").read()
    ordinal, nv, nn, nt, nf = np.frombuffer(b[:20], np.int32)
    off = [20]

    def take(n, dt, w):
        a = np.frombuffer(b[off[0]:off[0] + n * w * 4], dt).reshape(n, w).copy()
        off[0] += n * w * 4
        return a
    v = take(nv, np.float32, 3); n = take(nn, np.float32, 3); t = take(nt, np.float32, 2)
    vi = take(nf, np.uint32, 3); ni = take(nf, np.uint32, 3)
    ti = take(nf, np.uint32, 3) if nt else None
    return int(ordinal), dict(vertices=v, vidx=vi, normals=n, nidx=ni, uvs=t if nt else None, tidx=ti)


def run_reference(fx, rays=None, threads=1, repeat=1, warmup=0, want_hits=False, script=None, extra_args=(), scene_dir=None, dump_meshes=False,
                  shadow=None, extras=None):
    """One run of the unmodified reference behind its headless harness on the fixture's scene: Scene::trace over `rays`
    (timed inside the binary, scene load and BVH::build outside the timed region) and / or whatever `extra_args` ask for.
    Returns (events, hits): the harness's JSON event lines and, with want_hits, the reference's hit records (REFHIT); with
    dump_meshes a third value, {mesh name: mesh dict} — the geometry as the reference holds it after ITS load of the OBJ files
    (see FixtureData.use_meshes).  shadow = (light xyz, first, count): after the batch the harness casts PointLight shadow rays
    from the hits of rays [first, first + count) and traces them as a second timed batch ("shadow" event); `extras` (a dict)
    receives "shadow_rays" (RAY_DTYPE) and "shadow_hits" (REFHIT).  With threads > 1 the HIT RECORDS still come from a
    single-threaded pass (the reference's concurrent traversals corrupt each other)."""
    if not have_reference():
        raise FileNotFoundError(REF_BIN)
    with tempfile.TemporaryDirectory() as tmp:
        sdir = scene_dir or tmp
        sp = os.path.join(sdir, "scene.miro")
        if scene_dir is None or not os.path.exists(sp):
            sp = write_obj_scene(fx, sdir, script)      # texture-free scenes only: the fixtures carry decoded texels, not the image files
        cmd = [REF_BIN, "--scene", sp, "--assets", sdir, "--threads", str(threads), "--repeat", str(max(repeat, 1)), "--warmup", str(max(warmup, 0))]
        hp = os.path.join(tmp, "out.hits")
        if rays is not None:
            rp = os.path.join(tmp, "in.rays")
            np.ascontiguousarray(rays, RAY_DTYPE).tofile(rp)
            cmd += ["--trace", rp]
            if want_hits:
                cmd += ["--hits", hp]
            if shadow is not None:
                light, first, count = shadow
                cmd += ["--shadow-light"] + ["%.9g" % float(x) for x in light] + [str(int(first)), str(int(count)),
                        "--shadow-rays", os.path.join(tmp, "shadow.rays"), "--shadow-hits", os.path.join(tmp, "shadow.hits")]
        cmd += list(extra_args)
        md = os.path.join(tmp, "meshdump")
        if dump_meshes:
            os.makedirs(md)
            cmd += ["--dump-meshes", md]
        p = subprocess.run(cmd, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL, text=True)
        if p.returncode != 0:
            raise RuntimeError("miro_ref failed (%d): %s" % (p.returncode, p.stderr[-2000:]))
        events = [json.loads(l) for l in p.stderr.splitlines() if l.startswith("{")]
        hits = np.fromfile(hp, REFHIT) if (rays is not None and want_hits) else None
        if shadow is not None and extras is not None:
            extras["shadow_rays"] = np.fromfile(os.path.join(tmp, "shadow.rays"), RAY_DTYPE)
            extras["shadow_hits"] = np.fromfile(os.path.join(tmp, "shadow.hits"), REFHIT)
        if dump_meshes:
            meshes = {f[:-5]: read_mesh_dump(os.path.join(md, f))[1] for f in os.listdir(md) if f.endswith(".mesh")}
            return events, hits, meshes
    return events, hits


def compare_with_reference(mesh, tri, proxy, t, a, b, ref, rays=None, t_rel=1e-5, edge_eps=1e-4):
    """Product hits (already resolved to the reference's identities: mesh ordinal, triangle index, proxy ordinal) against
    the reference's hit records.  An id mismatch is a TIE — north_star: "the remainder only at edge/vertex ties" — when
      equal_t : both sides hit at the same distance (|dt| <= t_rel * t): an edge / vertex / coplanar overlap shared by two
                triangles, where the winner depends on visiting order;
      own_edge: the product's hit lies on a triangle edge (a barycentric weight within edge_eps of 0) — the ray grazes a
                silhouette edge and the two intersection tests round differently;
      ref_edge: the reference's hit does (its test is not watertight: it also misses through cracks).
    Everything else is a HARD mismatch.  The three causes are counted separately (first match in that order)."""
    r_hit = ref["mesh"] >= 0; g_hit = mesh >= 0
    same = (mesh == ref["mesh"]) & (tri == ref["tri"]) & (proxy == ref["proxy"])
    both = r_hit & g_hit
    dt = np.abs(t - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-30)

    def on_edge(aa, bb, valid):
        w = np.minimum(np.minimum(aa, bb), 1.0 - aa - bb)
        return valid & (w <= edge_eps)
    tie_t = ~same & both & (dt <= t_rel)
    tie_own = ~same & ~tie_t & on_edge(a, b, g_hit)
    tie_ref = ~same & ~tie_t & ~tie_own & on_edge(ref["a"], ref["b"], r_hit)
    tie = tie_t | tie_own | tie_ref
    hard = ~same & ~tie
    ok = same & both
    # instanced hits: the ray is moved into object space (ProxyObject.cpp:78-79), so t carries the rounding of
    # coordinates of magnitude |o|; for t << |o| "relative to t" is not attainable by ANY float32 implementation.
    # frac_t_within_pos measures |dt| against max(t, |o|) instead.
    pos = 1.0; max_pos = 0.0
    if rays is not None and ok.any():
        scale = np.maximum(np.abs(ref["t"]), np.linalg.norm(rays["o"], axis=1))
        e = np.abs(t - ref["t"])[ok] / scale[ok]
        pos = float((e <= t_rel).mean()); max_pos = float(e.max())
    return dict(frac_t_within_pos=pos, max_t_err_over_scale=max_pos, n=int(len(ref)), id_match=float(same.mean()), ties=int(tie.sum()),
                ties_equal_t=int(tie_t.sum()), ties_own_edge=int(tie_own.sum()), ties_ref_edge=int(tie_ref.sum()), hard=int(hard.sum()),
                hard_idx=np.nonzero(hard)[0], max_rel_t=float(dt[ok].max()) if ok.any() else 0.0,
                frac_t_within=float((dt[ok] <= t_rel).mean()) if ok.any() else 1.0,
                max_abs_a=float(np.abs(a - ref["a"])[ok].max()) if ok.any() else 0.0,
                max_abs_b=float(np.abs(b - ref["b"])[ok].max()) if ok.any() else 0.0,
                closer=int((~same & g_hit & ((t < ref["t"]) | ~r_hit)).sum()))


# ---------------------------------------------------------------------------------------------- float64 adjudication of mismatches
class SceneGeometry:
    """Triangles and instance transforms of a scene script by the identities hit records carry (mesh ordinal, triangle index,
    proxy ordinal), for re-computing single intersections in float64.  meshes: {mesh name: mesh dict}."""

    def __init__(self, script, meshes):
        self.names = []; self.mb_pair = {}; self.inv = []
        for line in script.splitlines():
            t = line.split("#")[0].split()
            if not t:
                continue
            if t[0] == "mesh":
                self.names.append(t[1])
            elif t[0] == "mbobject":
                self.mb_pair[self.names.index(t[1])] = self.names.index(t[2])
            elif t[0] == "instance":
                M = np.array([np.float32(x) for x in t[2:18]], np.float64).reshape(4, 4)
                self.inv.append(np.linalg.inv(M))
        self.meshes = [meshes[n] for n in self.names]

    def triangle(self, mesh, tri, time):
        m = self.meshes[mesh]
        v = m["vertices"][m["vidx"][tri]].astype(np.float64)
        if mesh in self.mb_pair:      # time * pose2 + (1 - time) * pose1, src/BVH.cpp:1320-1335
            m2 = self.meshes[self.mb_pair[mesh]]
            v = float(time) * m2["vertices"][m2["vidx"][tri]].astype(np.float64) + (1.0 - float(time)) * v
        return v

    def intersect(self, ray, mesh, tri, proxy):
        """Exact-arithmetic (float64) crossing of `ray` with one triangle: (t, edge_distance, reach) or None when the ray is
        parallel to the plane.  edge_distance: how far inside the triangle the crossing lies, measured in the plane
        perpendicular to the ray (where both sides' triangle tests decide) — negative outside; reach: the largest distance
        from the ray origin to a vertex, the magnitude at which float32 implementations round the translated vertices."""
        o = ray["o"].astype(np.float64); d = ray["d"].astype(np.float64)
        if proxy >= 0:
            Mi = self.inv[proxy]
            o = Mi[:3, :3] @ o + Mi[:3, 3]; d = Mi[:3, :3] @ d
        v = self.triangle(mesh, tri, ray["time"])
        e0, e1 = v[1] - v[0], v[2] - v[0]
        p = np.cross(d, e1); det = e0 @ p
        if det == 0.0:
            return None
        tv = o - v[0]; a = (tv @ p) / det; q = np.cross(tv, e0); b = (d @ q) / det; t = (e1 @ q) / det
        dh = d / np.linalg.norm(d)
        f0, f1 = e0 - (e0 @ dh) * dh, e1 - (e1 @ dh) * dh              # the triangle's edges projected along the ray
        area2 = np.linalg.norm(np.cross(f0, f1))
        dist = min(a * area2 / max(np.linalg.norm(f1), 1e-300), b * area2 / max(np.linalg.norm(f0), 1e-300),
                   (1.0 - a - b) * area2 / max(np.linalg.norm(f1 - f0), 1e-300))
        return t, dist, float(np.max(np.linalg.norm(v - o, axis=1)))


def adjudicate_mismatches(geom, rays, idx, mesh, tri, proxy, ref, ulps=16.0, t_eps=1e-6):
    """Every id mismatch in `idx` re-computed in float64 on both sides' triangles.  A crossing is CLEAR when it lies inside its
    triangle by more than `ulps` float32 roundings of the translated vertices (edge_distance > ulps * 2^-24 * reach: both the
    watertight edge test and Moller-Trumbore decide on vertices translated to the ray origin, rounded at that magnitude);
    nearer than that to an edge it is an edge / vertex tie.  Classes (first match):
      coincident       : both sides' crossings are real and their distances agree to t_eps: overlapping / touching surfaces;
      reference_missed : the product's crossing is real, clear, within [tmin, tmax) and NEARER than what the reference reports
                         (or the reference reports a miss) — the reference's traversal lost a hit (its slab test is not
                         conservative and its triangle test not watertight);
      product_missed   : the same with the roles swapped — a defect of the product: HARD;
      edge             : the nearer real crossing is not clear: an edge / vertex tie in float32;
      unexplained      : anything else: HARD."""
    out = dict(reference_missed=0, product_missed=0, coincident=0, edge=0, unexplained=0, hard_idx=[])
    eps32 = 2.0 ** -24
    for i in idx:
        r = rays[i]
        g = geom.intersect(r, int(mesh[i]), int(tri[i]), int(proxy[i])) if mesh[i] >= 0 else None
        f = geom.intersect(r, int(ref["mesh"][i]), int(ref["tri"][i]), int(ref["proxy"][i])) if ref["mesh"][i] >= 0 else None
        tol = lambda x: ulps * eps32 * x[2]
        ok = lambda x: x is not None and r["tmin"] <= x[0] < r["tmax"] and x[1] > -tol(x)
        clear = lambda x: x[1] > tol(x)
        gv, fv = ok(g), ok(f)
        if gv and fv and abs(g[0] - f[0]) <= t_eps * abs(f[0]):
            k = "coincident"
        elif gv and clear(g) and (not fv or g[0] < f[0]):
            k = "reference_missed"
        elif fv and clear(f) and (not gv or f[0] < g[0]):
            k = "product_missed"
        elif (gv and not clear(g)) or (fv and not clear(f)):
            k = "edge"
        else:
            k = "unexplained"
        out[k] += 1
        if k in ("product_missed", "unexplained"):
            out["hard_idx"].append(int(i))
    return out
