"""Thin Python handle on the native host layer (include/miro_host.h) and the GPU ABI (include/miro_gpu.h).

Used by tests/ and bench.py.  All geometry processing, BVH construction, flattening and every ray
is handled by libmiro_gpu.so; this file only marshals numpy / torch buffers into C pointers.
"""
import ctypes as C
import numpy as np

from . import capi

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32),
                      ("time", np.float32), ("flags", np.uint32), ("user", np.uint32, 2)])
HIT_DTYPE = np.dtype([("t", np.float32), ("a", np.float32), ("b", np.float32), ("prim", np.int32), ("inst", np.int32)])
RAY32_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32)])      # miro_gpu_ray32
assert RAY_DTYPE.itemsize == 48 and HIT_DTYPE.itemsize == 20 and RAY32_DTYPE.itemsize == 32


def pack_rays(rays):
    """miro_gpu_ray records -> miro_gpu_ray32 (drops time / flags / user words: static scenes)."""
    r = np.empty(len(rays), RAY32_DTYPE)
    for k in ("o", "tmin", "d", "tmax"):
        r[k] = rays[k]
    return r


class MiroError(RuntimeError):
    pass


def make_rays(origins, directions, tmin=capi.EPSILON, tmax=capi.TMAX, time=0.0):
    """Pack (n,3) origins / directions into the 48-byte miro_gpu_ray layout."""
    o = np.asarray(origins, np.float32); d = np.asarray(directions, np.float32)
    r = np.zeros(o.shape[0], RAY_DTYPE)
    r["o"] = o; r["d"] = d; r["tmin"] = tmin; r["tmax"] = tmax; r["time"] = time
    return r


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class MiroScene:
    """Scene described by a .miro script, pre-processed on the host, traced / rendered on the GPU."""

    def __init__(self):
        self.L = capi.lib()
        self.h = self.L.miro_host_new()
        self._keep = []
        self.loaded = False
        self.attached = False

    def close(self):
        if self.h:
            self.L.miro_host_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self.L.miro_host_error(self.h).decode()
            ctx = self.L.miro_host_ctx(self.h) if self.loaded else None
            if ctx:
                gm = self.L.miro_gpu_last_error(ctx).decode()
                if gm:
                    msg = f"{msg} | {gm}"
            raise MiroError(f"{what} failed (rc={rc}): {msg}")

    # ---- host side -------------------------------------------------------------------------
    def preload_mesh(self, name, vertices, vidx, normals=None, nidx=None, uvs=None, tidx=None):
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        vi = np.ascontiguousarray(vidx, np.uint32).reshape(-1, 3)
        n = np.ascontiguousarray(normals, np.float32).reshape(-1, 3) if normals is not None and nidx is not None else None
        ni = np.ascontiguousarray(nidx, np.uint32).reshape(-1, 3) if n is not None else None
        t = np.ascontiguousarray(uvs, np.float32).reshape(-1, 2) if uvs is not None and tidx is not None else None
        ti = np.ascontiguousarray(tidx, np.uint32).reshape(-1, 3) if t is not None else None
        rc = self.L.miro_host_preload_mesh(self.h, name.encode(), _ptr(v), len(v), _ptr(vi), len(vi),
                                           _ptr(n), 0 if n is None else len(n), _ptr(ni),
                                           _ptr(t), 0 if t is None else len(t), _ptr(ti))
        self._check(rc, "preload_mesh")

    def preload_image(self, name, texels, hdr=True):
        a = np.ascontiguousarray(texels, np.float32)
        if a.ndim == 2:
            a = a[:, :, None]
        h, w, c = a.shape
        self._check(self.L.miro_host_preload_image(self.h, name.encode(), _ptr(a), w, h, c, 1 if hdr else 0), "preload_image")

    def load_script(self, path, asset_root="."):
        self._check(self.L.miro_host_load_script(self.h, str(path).encode(), str(asset_root).encode()), "load_script")
        self.loaded = True
        return self

    def desc(self):
        d = capi.SceneDesc()
        self._check(self.L.miro_host_get_desc(self.h, C.byref(d)), "get_desc")
        return d

    def camera(self):
        c = capi.Camera()
        self._check(self.L.miro_host_get_camera(self.h, C.byref(c)), "get_camera")
        return c

    def render_params(self):
        p = capi.RenderParams()
        self._check(self.L.miro_host_get_render_params(self.h, C.byref(p)), "get_render_params")
        return p

    def bvh_stats(self):
        n, l, d, s = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_double()
        self._check(self.L.miro_host_bvh_stats(self.h, C.byref(n), C.byref(l), C.byref(d), C.byref(s)), "bvh_stats")
        return {"nodes": n.value, "leaves": l.value, "max_depth": d.value, "sah_cost": s.value}

    def prim_table(self):
        """(n_prims, 2) int array: (mesh ordinal, triangle index) per primitive id, and instance ordinals."""
        d = self.desc()
        n = d.n_tris + d.n_mbtris
        prims = np.ctypeslib.as_array(C.cast(d.prims, C.POINTER(C.c_uint32)), shape=(n, 12)) if n else np.zeros((0, 12), np.uint32)
        inst = (np.ctypeslib.as_array(C.cast(d.instances, C.POINTER(C.c_uint32)), shape=(d.n_instances, 16))[:, 13].copy()
                if d.n_instances else np.zeros(0, np.uint32))
        return prims[:, 7].astype(np.int64), prims[:, 8].astype(np.int64), inst.astype(np.int64)

    def resolve_hits(self, hits):
        """Map GPU hit records to the reference's identity (mesh ordinal, triangle index, proxy ordinal)."""
        mesh_of, tri_of, inst_ord = self.prim_table()
        prim = hits["prim"]; ok = prim >= 0
        mesh = np.full(len(hits), -1, np.int64); tri = np.full(len(hits), -1, np.int64); proxy = np.full(len(hits), -1, np.int64)
        mesh[ok] = mesh_of[prim[ok]]; tri[ok] = tri_of[prim[ok]]
        hi = ok & (hits["inst"] >= 0)
        proxy[hi] = inst_ord[hits["inst"][hi]]
        return mesh, tri, proxy

    # ---- GPU side --------------------------------------------------------------------------
    def attach(self, device=0):
        self._check(self.L.miro_host_attach(self.h, device), "attach")
        self.attached = True
        return self

    def attach_devices(self, devices, sample_sharding=False):
        """One caller, several GPUs (miro_gpu_group_*): render() shards the frame over `devices` and combines it on the first;
        trace_closest() / trace_any() split their batch.  The same device may be listed twice."""
        ids = (C.c_int * len(devices))(*devices)
        self._check(self.L.miro_host_attach_devices(self.h, ids, len(devices), 1 if sample_sharding else 0), "attach_devices")
        self.attached = True
        return self

    @property
    def group(self):
        return self.L.miro_host_group(self.h)

    def group_counters(self):
        c = capi.Counters()
        rc = self.L.miro_gpu_group_get_counters(self.group, C.byref(c))
        if rc != 0:
            raise MiroError("group_get_counters failed: " + self.L.miro_gpu_group_last_error(self.group).decode())
        return {k: getattr(c, k) for k, _ in capi.Counters._fields_}

    @property
    def ctx(self):
        return self.L.miro_host_ctx(self.h)

    def _gpu_check(self, rc, what):
        if rc != 0:
            raise MiroError(f"{what} failed (rc={rc}): {self.L.miro_gpu_last_error(self.ctx).decode()}")

    def trace_closest(self, rays):
        """Host buffers in, host buffers out (H2D + kernel + D2H inside the call)."""
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.empty(len(rays), HIT_DTYPE)
        if self.group:
            self._check(self.L.miro_host_trace(self.h, _ptr(rays), len(rays), _ptr(hits)), "trace (group)")
            return hits
        self._gpu_check(self.L.miro_gpu_trace_closest(self.ctx, _ptr(rays), len(rays), _ptr(hits)), "trace_closest")
        return hits

    def trace_any(self, rays):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        bits = np.zeros((len(rays) + 31) // 32, np.uint32)
        if self.group:
            self._check(self.L.miro_host_trace_any(self.h, _ptr(rays), len(rays), _ptr(bits)), "trace_any (group)")
            return np.unpackbits(bits.view(np.uint8), bitorder="little")[:len(rays)].astype(bool)
        self._gpu_check(self.L.miro_gpu_trace_any(self.ctx, _ptr(rays), len(rays), _ptr(bits)), "trace_any")
        return np.unpackbits(bits.view(np.uint8), bitorder="little")[:len(rays)].astype(bool)

    def trace_closest_packed(self, rays):
        """32-byte rays (pack_rays): miro_gpu_trace_closest_packed — the same hits as trace_closest of the same rays at time 0."""
        rays = np.ascontiguousarray(rays, RAY32_DTYPE)
        hits = np.empty(len(rays), HIT_DTYPE)
        self._gpu_check(self.L.miro_gpu_trace_closest_packed(self.ctx, _ptr(rays), len(rays), _ptr(hits)), "trace_closest_packed")
        return hits

    def trace_any_packed(self, rays):
        rays = np.ascontiguousarray(rays, RAY32_DTYPE)
        bits = np.zeros((len(rays) + 31) // 32, np.uint32)
        self._gpu_check(self.L.miro_gpu_trace_any_packed(self.ctx, _ptr(rays), len(rays), _ptr(bits)), "trace_any_packed")
        return np.unpackbits(bits.view(np.uint8), bitorder="little")[:len(rays)].astype(bool)

    def trace_primary(self, width=None, height=None, camera=None, seed=None, d_rays_out=None):
        """Primary rays generated on the device at the pixel centres and traced (miro_gpu_trace_primary): hit records [h * w]."""
        p = self.render_params(); c = camera or self.camera()
        w = width or p.width; h = height or p.height
        hits = np.empty(w * h, HIT_DTYPE)
        self._gpu_check(self.L.miro_gpu_trace_primary(self.ctx, C.byref(c), w, h, p.seed if seed is None else seed, _ptr(hits), d_rays_out), "trace_primary")
        return hits

    def trace_closest_device(self, d_rays_ptr, n, d_hits_ptr):
        self._gpu_check(self.L.miro_gpu_trace_closest_device(self.ctx, d_rays_ptr, n, d_hits_ptr), "trace_closest_device")

    def trace_any_device(self, d_rays_ptr, n, d_bits_ptr):
        self._gpu_check(self.L.miro_gpu_trace_any_device(self.ctx, d_rays_ptr, n, d_bits_ptr), "trace_any_device")

    def set_stream(self, cuda_stream_ptr):
        self._gpu_check(self.L.miro_gpu_set_stream(self.ctx, cuda_stream_ptr), "set_stream")

    def render(self, shard_index=0, shard_count=1, want_bytes=False, out=None):
        """Scene::raytraceImage through the host layer.  out = (rgb float32 [h, w, 3], rgb8 uint8 [h, w, 3] or None): buffers to
        reuse from frame to frame (fresh arrays cost a page fault per 4 KB, which at 1080p is as long as the frame itself)."""
        p = self.render_params()
        if out is not None:
            rgb, rgb8 = out
        else:
            rgb = np.zeros((p.height, p.width, 3), np.float32)
            rgb8 = np.zeros((p.height, p.width, 3), np.uint8) if want_bytes else None
        self._check(self.L.miro_host_raytrace_image(self.h, _ptr(rgb), _ptr(rgb8), shard_index, shard_count), "raytrace_image")
        return (rgb, rgb8) if want_bytes else rgb

    def render_in_place(self, shard_index=0, shard_count=1):
        """Scene::raytraceImage with the frame left in the scene's Image (no copies out): returns views (rgb float32 [h, w, 3],
        rgb8 uint8 [h, w, 3]) of the Image's own page-locked buffers, valid until the next render / resize."""
        self._check(self.L.miro_host_raytrace_image(self.h, None, None, shard_index, shard_count), "raytrace_image")
        f, b, w, h = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
        self._check(self.L.miro_host_image(self.h, C.byref(f), C.byref(b), C.byref(w), C.byref(h)), "image")
        rgb = np.ctypeslib.as_array(C.cast(f, C.POINTER(C.c_float)), shape=(h.value, w.value, 3))
        rgb8 = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_ubyte)), shape=(h.value, w.value, 3))
        return rgb, rgb8

    def render_device(self, d_rgb_ptr, params=None, camera=None):
        p = params or self.render_params(); c = camera or self.camera()
        self._gpu_check(self.L.miro_gpu_render(self.ctx, C.byref(c), C.byref(p), d_rgb_ptr), "render")

    def set_trace_chaining(self, on=True):
        """Let consecutive trace_*_device launches overlap (see miro_gpu_set_trace_chaining for the contract)."""
        self._gpu_check(self.L.miro_gpu_set_trace_chaining(self.ctx, 1 if on else 0), "set_trace_chaining")

    def set_trace_kernel(self, kind):
        """'auto' (the default: per scene), 'warp' (persistent warps, one ray per lane), 'pool' (64-ray pool per warp,
        csrc/trace_pool.cuh) or 'flat' (the warp kernel with leaf rounds dealt out over the warp, csrc/trace_flat.cuh)."""
        self._gpu_check(self.L.miro_gpu_set_trace_kernel(self.ctx, {"auto": -1, "warp": 0, "pool": 1, "flat": 2}[kind]), "set_trace_kernel")

    def trace_kernel(self):
        """The traversal kernel in effect on this context."""
        return {-1: "auto", 0: "warp", 1: "pool", 2: "flat"}[self.L.miro_gpu_get_trace_kernel(self.ctx)]

    def enable_counting(self, on=True):
        self._gpu_check(self.L.miro_gpu_enable_counting(self.ctx, 1 if on else 0), "enable_counting")

    def reset_counters(self):
        self._gpu_check(self.L.miro_gpu_reset_counters(self.ctx), "reset_counters")

    def counters(self):
        c = capi.Counters()
        self._gpu_check(self.L.miro_gpu_get_counters(self.ctx, C.byref(c)), "get_counters")
        return {k: getattr(c, k) for k, _ in capi.Counters._fields_}
