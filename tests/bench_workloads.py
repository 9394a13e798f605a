"""Benchmark workloads shared by both arms of bench.py (numpy + tests/reference_arm.py only — no product import, so the
reference arm never maps the library it is compared with).

A workload = a scene script (read by the product's host layer AND by the harness around the unmodified reference), the
golden fixture its meshes come from, and three seeded ray batches of 1920x1080 rays each: coherent primary rays at the pixel
centres, incoherent closest-hit rays, and the PointLight shadow rays cast from the incoherent batch's hits (any-hit).

  c2   BASELINE config C2 stand-in: Models/Final/explosion01.obj, 86 914 triangles (bunny / dragon_2 are absent from the mount)
  big  C2 at dragon scale: 20 placed copies of explosion01.obj (the 20-copy pattern of makeBunny20Scene, src/assignment2.h:131-345)
       = 1 738 280 triangles, ~110 MB of nodes + triangles on the device
  c5   BASELINE config C5 at the reference's own scale: motion-blur bullets + makeProxyGrid's 201 x 201 = 40 401 ProxyObject
       instances of testGrass.obj (src/main.cpp:37-52); rays carry random times
"""
import importlib.util
import os

import numpy as np

import reference_arm as ra

ROOT = ra.ROOT
WIDTH, HEIGHT = 1920, 1080
N_BATCH = WIDTH * HEIGHT


def _make_scenes():
    spec = importlib.util.spec_from_file_location("make_scenes", os.path.join(ROOT, "tools", "make_scenes.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def _big_script():
    rng = np.random.default_rng(20)
    lines = ["# C2 at dragon scale: 20 placed copies of explosion01.obj (cf. makeBunny20Scene, src/assignment2.h:131-345)",
             "image %d %d" % (WIDTH, HEIGHT),
             "camera eye 0.3 2.4 5.6 lookat 0 0.45 0 up 0 1 0 fov 40",
             "scene bgcolor 0 0 0 pathtrace 0 numpaths 1 minsubdivs 1 maxsubdivs 1",
             "material grey lambert kd 0.7 0.7 0.7",
             "light point pos -3.0 6.0 4.0 power 900"]
    k = 0
    for i in range(5):
        for j in range(4):
            a = rng.uniform(-0.7, 0.7); s = rng.uniform(0.8, 1.2, 3)
            c, sn = np.cos(a), np.sin(a)
            tx = (i - 2) * 1.15 + rng.uniform(-0.1, 0.1); tz = (j - 1.5) * 0.8 + rng.uniform(-0.1, 0.1)
            M = np.array([[c * s[0], 0, sn * s[2], tx], [0, s[1], 0, 0], [-sn * s[0], 0, c * s[2], tz], [0, 0, 0, 1]])
            lines.append("mesh e%02d @explosion ctm " % k + " ".join("%.7g" % v for v in M.reshape(-1)))
            k += 1
    lines += ["object e%02d grey" % k for k in range(20)]
    return "\n".join(lines) + "\n"


class Workload:
    def __init__(self, name):
        self.name = name
        if name == "c2":
            self.fixture = "c2_explosion"; self.script = None
            self.light = np.array([-2.0, 4.0, 3.0], np.float32); self.times = False
            self.label = "C2 stand-in: explosion01.obj 86914 tris, 1920x1080"
        elif name == "big":
            self.fixture = "c2_explosion"; self.script = _big_script()
            self.light = np.array([-3.0, 6.0, 4.0], np.float32); self.times = False
            self.label = "C2 at dragon scale: 20 placed copies of explosion01.obj = 1738280 tris, 1920x1080"
        elif name == "c5":
            self.fixture = "c5_mb_instances"; self.script = _make_scenes().c5(201, name=None, res=256)
            self.script = self.script.replace("image 256 256", "image %d %d" % (WIDTH, HEIGHT))
            self.light = np.array([-4.0, 30.0, 20.0], np.float32); self.times = True
            self.label = "C5 at makeProxyGrid scale: MB bullets + 201x201 = 40401 ProxyObject instances of testGrass.obj (5172 tris each), 1920x1080"
        else:
            raise ValueError(name)
        self.fx = None

    def load(self, fixture_class=ra.FixtureData):
        path = ra.fixture_path(self.fixture, full=True) or ra.fixture_path(self.fixture)
        self.fx = fixture_class(path)
        if self.script is None:
            self.script = self.fx.script
        return self

    # ---- the scene as files the reference's own loaders read
    def materialise(self, tmp):
        sp = ra.write_obj_scene(self.fx, tmp, self.script)
        text = open(sp).read()
        for name in self.fx.names:
            text = text.replace("@" + name, os.path.join(tmp, name + ".obj"))
        open(sp, "w").write(text)
        return sp

    def mesh_names(self):
        return [l.split()[1] for l in self.script.splitlines() if l.split()[:1] == ["mesh"]]

    def triangles(self):
        per = {n: len(self.fx.mesh(k)["vidx"]) for k, n in enumerate(self.fx.names)}
        n = 0
        for l in self.script.splitlines():
            t = l.split()
            if t[:1] == ["mesh"]:
                n += per.get(t[1], per.get(t[2].lstrip("@"), 0))
        return n

    # ---- ray batches
    def bounds(self):
        """World bounds the incoherent origins are drawn from."""
        if self.name == "c5":
            return np.array([-25.0, 0.05, -20.0], np.float32), np.array([17.0, 14.0, 20.0], np.float32)
        lo, hi = self.fx.bounds()
        if self.name == "big":
            pts = []
            corners = np.array([[(hi if (c >> k) & 1 else lo)[k] for k in range(3)] for c in range(8)], np.float64)
            for l in self.script.splitlines():
                t = l.split()
                if t[:1] == ["mesh"] and "ctm" in t:
                    M = np.array([float(x) for x in t[t.index("ctm") + 1:t.index("ctm") + 17]]).reshape(4, 4)
                    pts.append(corners @ M[:3, :3].T + M[:3, 3])
            pts = np.concatenate(pts)
            lo, hi = pts.min(0).astype(np.float32), pts.max(0).astype(np.float32)
        return lo, hi

    def primary(self):
        return ra.primary_rays(ra.script_camera(self.script), WIDTH, HEIGHT)

    def incoherent(self, seed):
        lo, hi = self.bounds()
        if self.name == "c5":      # origins over the field, directions biased downwards so they meet it (tools/instance_bench.py)
            rng = np.random.default_rng(seed)
            r = np.zeros(N_BATCH, ra.RAY_DTYPE)
            r["o"] = rng.uniform(lo, hi, (N_BATCH, 3)).astype(np.float32)
            d = rng.normal(size=(N_BATCH, 3)); d[:, 1] = -np.abs(d[:, 1]) * 0.5; d /= np.linalg.norm(d, axis=1, keepdims=True)
            r["d"] = d.astype(np.float32); r["tmin"] = 1e-3; r["tmax"] = 1e12
            r["time"] = rng.uniform(0, 1, N_BATCH).astype(np.float32)
            return r
        return ra.incoherent_rays(lo, hi, N_BATCH, seed, times=self.times)

    def shadow(self, src, hit_t, hit_mask):
        return ra.shadow_rays(src, hit_t, hit_mask, self.light)

    def sample(self, prim, inco):
        """The bounded sample the reference traces: 1/8 of the primary batch (every 2nd row, every 4th column: keeps the
        coherence of the rows) and 1/8 of the incoherent batch; the harness casts the shadow sample from the latter's hits."""
        p = prim.reshape(HEIGHT, WIDTH)[::2, ::4].reshape(-1).copy()
        q = inco[::8].copy()
        return p, q
