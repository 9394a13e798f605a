#!/usr/bin/env python
"""Generate parity fixtures with the reference oracle (oracle/_ref/miro_ref).

Runs HERE (where /root/reference exists).  For each scene script in tests/scenes it records what the
UNMODIFIED reference produced:
  * the meshes exactly as the reference's loader left them (TriangleMeshLoad.cpp),
  * the reference's own camera rays at pixel centres (Camera::eyeRayAdaptive) — or a seeded synthetic
    incoherent batch — and the reference's Scene::trace result for every one of them,
  * the float radiance image of Scene::adaptiveSampleScene and the stock 8-bit render.
Two files per scene:
  oracle/_ref/fixtures/<scene>.npz   full size (git-ignored; travels to the GPU box with the snapshot)
  tests/golden/<scene>.npz           committed subset: <= GOLDEN_RAYS rays, float16 radiance
usage: tools/make_fixtures.py [scene ...]
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "miro_ref")
ASSETS = os.environ.get("MIRO_REFERENCE_ROOT", "/root/reference")
FULL = os.path.join(ROOT, "oracle", "_ref", "fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_RAYS = 32768

RAY = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32),
                ("time", np.float32), ("flags", np.uint32), ("user", np.uint32, 2)])
REFHIT = np.dtype([("t", "f4"), ("a", "f4"), ("b", "f4"), ("mesh", "i4"), ("tri", "i4"), ("proxy", "i4")])

# per scene: render the float image? run the stock 8-bit render? how many incoherent rays to add?
SCENES = {
    "c1_cornell": dict(render=True, stock=True, incoherent=0, qbvh=True),
    "c2_explosion": dict(render=True, stock=True, incoherent=1 << 20, qbvh=True),
    "c5_mb_instances": dict(render=True, stock=False, incoherent=1 << 19, threads=1),
    "c3_dome_pt": dict(render=True, stock=False, incoherent=0, threads=1, converged=32),
    "c4_cornell_pt": dict(render=True, stock=False, incoherent=0, threads=1, converged=64),
    "c7_foliage": dict(render=True, stock=False, incoherent=1 << 18, threads=1),
    "c8_dispersion": dict(render=True, stock=False, incoherent=0, threads=1, converged=16),
    "c6_cornell_glass": dict(render=True, stock=False, incoherent=0, threads=1, converged=16),
    "c9_texmaps": dict(render=True, stock=False, incoherent=0, threads=1, converged=16),
    "c10_full_shadows": dict(render=True, stock=False, incoherent=0, threads=1, converged=32),
    # same geometry and light map as c3: the committed file holds only the script and the reference's images (overlay of c3_dome_pt)
    "c11_dome_full_shadows": dict(render=True, stock=False, incoherent=0, threads=1, converged=16, overlay_of="c3_dome_pt"),
}


def ref_events(stderr):
    return [json.loads(l) for l in stderr.splitlines() if l.startswith("{")]


def read_mesh(path):
    b = open(path, "rb").read()
    ordinal, nv, nn, nt, nf = np.frombuffer(b[:20], np.int32)
    off = [20]

    def take(n, dt, w):
        a = np.frombuffer(b[off[0]:off[0] + n * w * 4], dt).reshape(n, w).copy()
        off[0] += n * w * 4
        return a
    v = take(nv, np.float32, 3); n = take(nn, np.float32, 3); t = take(nt, np.float32, 2)
    vi = take(nf, np.uint32, 3); ni = take(nf, np.uint32, 3)
    ti = take(nf, np.uint32, 3) if nt else np.zeros((0, 3), np.uint32)
    return int(ordinal), dict(v=v, n=n, t=t, vi=vi, ni=ni, ti=ti)


def incoherent_rays(meshes, n, seed=0x5EED):
    """Seeded synthetic incoherent batch: origins ~U(scene AABB inflated 10%), directions ~U(S^2) (SURVEY 8d C2 ii)."""
    rng = np.random.default_rng(seed)
    allv = np.concatenate([m["v"] for m in meshes.values()])
    lo, hi = allv.min(0), allv.max(0)
    c, e = 0.5 * (lo + hi), 0.55 * (hi - lo) + 1e-3
    r = np.zeros(n, RAY)
    r["o"] = (c + e * rng.uniform(-1, 1, (n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["d"] = d.astype(np.float32); r["tmin"] = 1e-3; r["tmax"] = 1e12
    r["time"] = rng.uniform(0, 1, n).astype(np.float32)
    return r


def trace(script, rays, tmp, threads=1):
    rp, hp = os.path.join(tmp, "in.rays"), os.path.join(tmp, "out.hits")
    rays.tofile(rp)
    p = subprocess.run([REF, "--scene", script, "--assets", ASSETS, "--threads", str(threads), "--trace", rp, "--hits", hp],
                       stderr=subprocess.PIPE, text=True, check=True)
    return np.fromfile(hp, REFHIT), ref_events(p.stderr)


def to_rgbe(a):
    """float RGB -> shared-exponent RGBE bytes (lossless for texels that were decoded from a Radiance .hdr file)."""
    m = a.max(axis=-1)
    e = np.where(m > 1e-38, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, -128).astype(np.int32)
    scale = np.where(m > 1e-38, np.ldexp(1.0, 8 - e), 0.0)
    rgb = np.clip(np.floor(a * scale[..., None] + 0.5), 0, 255)
    # a mantissa that rounds to 256 would need the next exponent
    bump = (rgb.max(axis=-1) > 255)
    out = np.concatenate([rgb, (e + 128)[..., None]], axis=-1).astype(np.uint8)
    assert not bump.any()
    return out


def from_rgbe(b):
    e = b[..., 3].astype(np.int32)
    return (b[..., :3].astype(np.float32) * np.ldexp(1.0, e - 136).astype(np.float32)[..., None]) * (e > 0)[..., None]


def pack(out, meshes, names, script_text, events, rays, hits, ray_index, radiance, image8, shape, radiance_dtype, textures=None, extra=None):
    d = dict(mesh_names=np.array(names), script=np.array(script_text), events=np.array(json.dumps(events)),
             rays=rays, hits=hits, ray_index=ray_index, image_shape=np.array(shape, np.int32))
    for k, name in enumerate(names):
        for key, arr in meshes[name].items():
            if key in ("vi", "ni", "ti") and arr.size and arr.max() < 65536:
                arr = arr.astype(np.uint16)
            d[f"m{k}_{key}"] = arr
    if radiance is not None:
        d["radiance"] = radiance.astype(radiance_dtype)
    if image8 is not None:
        d["image8"] = image8
    for name, (tex, kind) in (textures or {}).items():
        d["texkind_" + name] = np.array(kind, np.int32)
        if radiance_dtype == np.float16 and kind == 3 and np.array_equal(from_rgbe(to_rgbe(tex)), tex):
            d["texrgbe_" + name] = to_rgbe(tex)       # committed fixture: HDR texels as the RGBE bytes they were decoded from
        else:
            d["tex_" + name] = tex
    for k, v in (extra or {}).items():
        d[k] = v.astype(radiance_dtype) if (v.dtype == np.float32 and not k.startswith("qbvh")) else v
    np.savez_compressed(out, **d)
    print("wrote", out, "%.2f MB" % (os.path.getsize(out) / 1e6))


def run(scene, opt):
    script = os.path.join(ROOT, "tests", "scenes", scene + ".miro")
    text = open(script).read()
    threads = opt.get("threads", 1)
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [REF, "--scene", script, "--assets", ASSETS, "--threads", str(threads), "--dump-meshes", tmp, "--dump-textures", tmp, "--dump-qbvh", os.path.join(tmp, "qbvh.bin"),
               "--dump-primary", os.path.join(tmp, "primary.rays")]
        if opt["render"]:
            cmd += ["--render-float", os.path.join(tmp, "radiance.f32")]
        p = subprocess.run(cmd, stderr=subprocess.PIPE, text=True, check=True)
        events = ref_events(p.stderr)
        meshes, order = {}, {}
        for f in os.listdir(tmp):
            if f.endswith(".mesh"):
                o, m = read_mesh(os.path.join(tmp, f)); meshes[f[:-5]] = m; order[f[:-5]] = o
        names = sorted(meshes, key=lambda k: order[k])
        textures = {}
        for f in os.listdir(tmp):
            if f.endswith(".tex"):
                b = open(os.path.join(tmp, f), "rb").read()
                tw, th, tc, kind = np.frombuffer(b[:16], np.int32)
                textures[f[:-4]] = (np.frombuffer(b[16:], np.float32).reshape(th, tw, tc).copy(), int(kind))
        extra = {}
        if opt.get("qbvh"):           # the reference's own QBVH (BVH.cpp:100-389), walked by the harness: bounds, children, leaf lanes
            b = open(os.path.join(tmp, "qbvh.bin"), "rb").read()
            nn, nl = np.frombuffer(b[:8], np.int32)
            off = 8
            extra["qbvh_bounds"] = np.frombuffer(b[off:off + nn * 96], np.float32).reshape(nn, 24).copy(); off += nn * 96
            extra["qbvh_child"] = np.frombuffer(b[off:off + nn * 16], np.int32).reshape(nn, 4).copy(); off += nn * 16
            extra["qbvh_leaves"] = np.frombuffer(b[off:off + nl * 48], np.int32).reshape(nl, 4, 3).copy()
        if opt.get("converged"):      # a second, converged render of the same scene: numpaths multiplied
            conv = os.path.join(tmp, "converged.miro")
            import re
            n0 = int(re.search(r"numpaths (\d+)", text).group(1))
            open(conv, "w").write(re.sub(r"numpaths \d+", "numpaths %d" % (n0 * opt["converged"]), text))
            cf = os.path.join(tmp, "converged.f32")
            p2 = subprocess.run([REF, "--scene", conv, "--assets", ASSETS, "--threads", str(threads), "--render-float", cf],
                                stderr=subprocess.PIPE, text=True, check=True)
            ev2 = [e for e in ref_events(p2.stderr) if e["event"] == "render_float"][0]
            extra["radiance_converged"] = np.fromfile(cf, np.float32).reshape(ev2["height"], ev2["width"], 3)
            extra["converged_numpaths"] = np.array(n0 * opt["converged"], np.int32)
        rays = np.fromfile(os.path.join(tmp, "primary.rays"), RAY)
        image8 = None
        if opt["stock"]:
            ppm = os.path.join(tmp, "stock.ppm")
            p = subprocess.run([REF, "--scene", script, "--assets", ASSETS, "--threads", "1", "--render-stock", ppm],
                               stderr=subprocess.PIPE, text=True, check=True)
            events += ref_events(p.stderr)
            raw = open(ppm, "rb").read()
            hdr = raw.split(b"\n", 3)
            pw, ph = map(int, hdr[1].split())
            image8 = np.frombuffer(hdr[3], np.uint8).reshape(ph, pw, 3)[::-1].copy()   # PPM is top-down; row 0 = bottom here
        radiance = None
        if opt["render"]:
            ev = [e for e in events if e["event"] == "render_float"][0]
            radiance = np.fromfile(os.path.join(tmp, "radiance.f32"), np.float32).reshape(ev["height"], ev["width"], 3)
            shape = (ev["height"], ev["width"])
        else:
            shape = (0, 0)
        n_primary = len(rays)
        if opt["incoherent"]:
            rays = np.concatenate([rays, incoherent_rays(meshes, opt["incoherent"])])
        hits, ev = trace(script, rays, tmp)
        events += ev
    os.makedirs(FULL, exist_ok=True); os.makedirs(GOLDEN, exist_ok=True)
    idx = np.arange(len(rays), dtype=np.int64)
    pack(os.path.join(FULL, scene + ".npz"), meshes, names, text, events, rays, hits, idx, radiance, image8, shape, np.float32, textures, extra)
    # committed subset: every ray that hit has the same chance as a miss; keep primary and incoherent halves
    rng = np.random.default_rng(12345)
    if len(rays) > GOLDEN_RAYS:
        hit = np.nonzero(hits["mesh"] >= 0)[0]; miss = np.nonzero(hits["mesh"] < 0)[0]
        k_hit = min(len(hit), GOLDEN_RAYS * 3 // 4); k_miss = min(len(miss), GOLDEN_RAYS - k_hit)
        sel = np.sort(np.concatenate([rng.choice(hit, k_hit, replace=False), rng.choice(miss, k_miss, replace=False)]))
    else:
        sel = idx
    small_rad = radiance
    small_img = image8
    if radiance is not None and radiance.shape[0] * radiance.shape[1] > 512 * 512:
        small_rad = None; small_img = None      # large images stay in the full fixture only
    if opt.get("overlay_of"):
        np.savez_compressed(os.path.join(GOLDEN, scene + ".npz"), overlay_of=np.array(opt["overlay_of"]), script=np.array(text), events=np.array(json.dumps(events)),
                            image_shape=np.array(shape, np.int32), radiance=radiance.astype(np.float16), radiance_converged=extra["radiance_converged"].astype(np.float16))
    else:
        pack(os.path.join(GOLDEN, scene + ".npz"), meshes, names, text, events, rays[sel], hits[sel], sel, small_rad, small_img, shape, np.float16, textures, extra)
    print(scene, "primary", n_primary, "total rays", len(rays), "hit fraction %.3f" % (hits["mesh"] >= 0).mean())


if __name__ == "__main__":
    for s in (sys.argv[1:] or ["c1_cornell", "c2_explosion"]):
        run(s, SCENES[s])
