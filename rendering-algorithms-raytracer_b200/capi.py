"""ctypes view of the two C ABIs exported by libmiro_gpu.so (include/miro_gpu.h, include/miro_host.h).

Plumbing only: every function here forwards to the native library; there is no Python or CPU
implementation of the ray-casting path.  Importing this module fails loudly when the shared
library has not been built (run `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MIRO_GPU_LIB") or os.path.join(_HERE, "libmiro_gpu.so")   # MIRO_GPU_LIB: kernel-tuning builds (tools/tune.sh)

OK, EINVAL, ENODEVICE, ECUDA, ENOSCENE, EUNSUPPORTED, ENOMEM = 0, -1, -2, -3, -4, -5, -6
TMAX = 1e12
EPSILON = 0.001
CHILD_EMPTY = 0x7FFFFFFF
ROOT_BUILD_ON_DEVICE = 0x7FFFFFFD
KIND_TRI, KIND_MBTRI, KIND_INST = 0, 1, 2
LEAF_INDEX_BITS = 26


class Ray(C.Structure):
    _fields_ = [("ox", C.c_float), ("oy", C.c_float), ("oz", C.c_float), ("tmin", C.c_float),
                ("dx", C.c_float), ("dy", C.c_float), ("dz", C.c_float), ("tmax", C.c_float),
                ("time", C.c_float), ("flags", C.c_uint32), ("user0", C.c_uint32), ("user1", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("t", C.c_float), ("a", C.c_float), ("b", C.c_float), ("prim", C.c_int32), ("inst", C.c_int32)]


class Node(C.Structure):
    _fields_ = [("lo_x", C.c_float * 4), ("lo_y", C.c_float * 4), ("lo_z", C.c_float * 4),
                ("hi_x", C.c_float * 4), ("hi_y", C.c_float * 4), ("hi_z", C.c_float * 4),
                ("child", C.c_int32 * 4), ("reserved", C.c_uint32 * 4)]


class Tri(C.Structure):
    _fields_ = [("v0", C.c_float * 3), ("pad0", C.c_uint32), ("v1", C.c_float * 3), ("pad1", C.c_uint32),
                ("v2", C.c_float * 3), ("pad2", C.c_uint32)]


class MBTri(C.Structure):
    _fields_ = [("pose", Tri * 2)]


class Instance(C.Structure):
    _fields_ = [("inv", C.c_float * 12), ("blas_root", C.c_int32), ("ordinal", C.c_uint32), ("w_recip", C.c_float), ("reserved", C.c_uint32)]


class Prim(C.Structure):
    _fields_ = [("n", C.c_uint32 * 3), ("uv", C.c_uint32 * 3), ("material", C.c_uint32), ("mesh", C.c_uint32),
                ("tri", C.c_uint32), ("reserved", C.c_uint32 * 3)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("kd", C.c_float * 3), ("ka", C.c_float * 3), ("ks", C.c_float * 3),
                ("spec_exp", C.c_float), ("spec_amt", C.c_float), ("emit_intensity", C.c_float), ("le", C.c_float * 3),
                ("color_map", C.c_int32), ("alpha_map", C.c_int32), ("reflect_amt", C.c_float), ("refract_amt", C.c_float),
                ("spec_gloss", C.c_float), ("translucency", C.c_float), ("sample_env", C.c_uint32), ("ior", C.c_float * 3), ("disperse", C.c_uint32),
                ("normal_map", C.c_int32), ("specular_map", C.c_int32), ("reflect_map", C.c_int32), ("refract_map", C.c_int32), ("reserved", C.c_uint32)]


class Light(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("p0", C.c_float * 3), ("p1", C.c_float * 3), ("p2", C.c_float * 3),
                ("power", C.c_float), ("num_samples", C.c_int32), ("noise_threshold", C.c_float),
                ("cast_shadows", C.c_uint32), ("texture", C.c_int32), ("full_shadows", C.c_uint32)]


class Texture(C.Structure):
    _fields_ = [("texels", C.POINTER(C.c_float)), ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
                ("reserved", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32),
                ("nodes", C.POINTER(Node)), ("n_nodes", C.c_uint32), ("root", C.c_int32),
                ("tris", C.POINTER(Tri)), ("n_tris", C.c_uint32),
                ("mbtris", C.POINTER(MBTri)), ("n_mbtris", C.c_uint32),
                ("instances", C.POINTER(Instance)), ("n_instances", C.c_uint32),
                ("prims", C.POINTER(Prim)),
                ("normals", C.POINTER(C.c_float)), ("n_normals", C.c_uint32),
                ("uvs", C.POINTER(C.c_float)), ("n_uvs", C.c_uint32),
                ("inst_normal_xform", C.POINTER(C.c_float)),
                ("tangents", C.POINTER(C.c_float)), ("bitangents", C.POINTER(C.c_float)),
                ("materials", C.POINTER(Material)), ("n_materials", C.c_uint32),
                ("lights", C.POINTER(Light)), ("n_lights", C.c_uint32),
                ("textures", C.POINTER(Texture)), ("n_textures", C.c_uint32),
                ("env_map", C.c_int32), ("env_exposure", C.c_float), ("bg_color", C.c_float * 3)]


class Camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("view_dir", C.c_float * 3), ("up", C.c_float * 3), ("fov_deg", C.c_float),
                ("focus_plane", C.c_float), ("aperture", C.c_float), ("shutter_speed", C.c_float)]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("min_subdivs", C.c_int32), ("max_subdivs", C.c_int32),
                ("noise_threshold", C.c_float), ("num_paths", C.c_int32), ("max_bounces", C.c_int32),
                ("path_trace", C.c_uint32), ("sample_env", C.c_uint32), ("seed", C.c_uint64),
                ("shard_index", C.c_int32), ("shard_count", C.c_int32), ("path_shard_index", C.c_int32), ("path_shard_count", C.c_int32),
                ("reserved", C.c_uint32 * 2)]


class Counters(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_any", C.c_uint64), ("nodes_fetched", C.c_uint64),
                ("tris_tested", C.c_uint64), ("insts_entered", C.c_uint64), ("trace_ms", C.c_double),
                ("total_ms", C.c_double), ("kernel_launches", C.c_uint64)]


# every symbol include/miro_gpu.h and include/miro_host.h declare (checked by tests/test_abi.py)
GPU_SYMBOLS = ["miro_gpu_create", "miro_gpu_destroy", "miro_gpu_last_error", "miro_gpu_abi_version", "miro_gpu_sizeof", "miro_gpu_set_stream", "miro_gpu_set_trace_chaining", "miro_gpu_set_trace_kernel", "miro_gpu_get_trace_kernel",
               "miro_gpu_upload_scene", "miro_gpu_trace_closest", "miro_gpu_trace_any", "miro_gpu_trace_closest_packed", "miro_gpu_trace_any_packed", "miro_gpu_trace_closest_device",
               "miro_gpu_trace_any_device", "miro_gpu_trace_primary", "miro_gpu_render", "miro_gpu_render_image", "miro_gpu_pin_host_buffer", "miro_gpu_unpin_host_buffer", "miro_gpu_enable_counting", "miro_gpu_get_counters",
               "miro_gpu_reset_counters",
               "miro_gpu_group_create", "miro_gpu_group_destroy", "miro_gpu_group_size", "miro_gpu_group_ctx", "miro_gpu_group_last_error", "miro_gpu_group_peer_access",
               "miro_gpu_group_upload_scene", "miro_gpu_group_render", "miro_gpu_group_trace_closest", "miro_gpu_group_trace_any", "miro_gpu_group_get_counters",
               "miro_gpu_group_reset_counters"]
HOST_SYMBOLS = ["miro_host_new", "miro_host_free", "miro_host_error", "miro_host_preload_mesh", "miro_host_preload_image",
                "miro_host_load_script", "miro_host_get_desc", "miro_host_get_camera", "miro_host_get_render_params",
                "miro_host_bvh_stats", "miro_host_attach", "miro_host_attach_devices", "miro_host_group", "miro_host_trace", "miro_host_trace_any",
                "miro_host_ctx", "miro_host_raytrace_image", "miro_host_image", "miro_host_write_ppm"]

_lib = None


def lib():
    """Load libmiro_gpu.so (once) and declare prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with rendering-algorithms-raytracer_b200/build.sh "
                           "(or __graft_entry__.build()); there is no Python/CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, cp, i32, u32, sz = C.c_void_p, C.c_char_p, C.c_int, C.c_uint32, C.c_size_t
    L.miro_gpu_create.argtypes = [C.POINTER(vp), i32]; L.miro_gpu_create.restype = i32
    L.miro_gpu_destroy.argtypes = [vp]; L.miro_gpu_destroy.restype = None
    L.miro_gpu_last_error.argtypes = [vp]; L.miro_gpu_last_error.restype = cp
    L.miro_gpu_abi_version.argtypes = []; L.miro_gpu_abi_version.restype = i32
    L.miro_gpu_set_stream.argtypes = [vp, vp]; L.miro_gpu_set_stream.restype = i32
    L.miro_gpu_upload_scene.argtypes = [vp, C.POINTER(SceneDesc)]; L.miro_gpu_upload_scene.restype = i32
    for name in ("miro_gpu_trace_closest", "miro_gpu_trace_any", "miro_gpu_trace_closest_device", "miro_gpu_trace_any_device",
                 "miro_gpu_trace_closest_packed", "miro_gpu_trace_any_packed"):
        f = getattr(L, name); f.argtypes = [vp, vp, sz, vp]; f.restype = i32
    L.miro_gpu_trace_primary.argtypes = [vp, C.POINTER(Camera), i32, i32, C.c_uint64, vp, vp]; L.miro_gpu_trace_primary.restype = i32
    L.miro_gpu_render.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), vp]; L.miro_gpu_render.restype = i32
    L.miro_gpu_enable_counting.argtypes = [vp, i32]; L.miro_gpu_enable_counting.restype = i32
    L.miro_gpu_set_trace_chaining.argtypes = [vp, i32]; L.miro_gpu_set_trace_chaining.restype = i32
    L.miro_gpu_set_trace_kernel.argtypes = [vp, i32]; L.miro_gpu_set_trace_kernel.restype = i32
    L.miro_gpu_get_trace_kernel.argtypes = [vp]; L.miro_gpu_get_trace_kernel.restype = i32
    L.miro_gpu_get_counters.argtypes = [vp, C.POINTER(Counters)]; L.miro_gpu_get_counters.restype = i32
    L.miro_gpu_reset_counters.argtypes = [vp]; L.miro_gpu_reset_counters.restype = i32
    L.miro_host_new.argtypes = []; L.miro_host_new.restype = vp
    L.miro_host_free.argtypes = [vp]; L.miro_host_free.restype = None
    L.miro_host_error.argtypes = [vp]; L.miro_host_error.restype = cp
    L.miro_host_preload_mesh.argtypes = [vp, cp, vp, u32, vp, u32, vp, u32, vp, vp, u32, vp]; L.miro_host_preload_mesh.restype = i32
    L.miro_host_preload_image.argtypes = [vp, cp, vp, i32, i32, i32, i32]; L.miro_host_preload_image.restype = i32
    L.miro_host_load_script.argtypes = [vp, cp, cp]; L.miro_host_load_script.restype = i32
    L.miro_host_get_desc.argtypes = [vp, C.POINTER(SceneDesc)]; L.miro_host_get_desc.restype = i32
    L.miro_host_get_camera.argtypes = [vp, C.POINTER(Camera)]; L.miro_host_get_camera.restype = i32
    L.miro_host_get_render_params.argtypes = [vp, C.POINTER(RenderParams)]; L.miro_host_get_render_params.restype = i32
    L.miro_host_bvh_stats.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_double)]; L.miro_host_bvh_stats.restype = i32
    L.miro_host_attach.argtypes = [vp, i32]; L.miro_host_attach.restype = i32
    L.miro_host_ctx.argtypes = [vp]; L.miro_host_ctx.restype = vp
    L.miro_host_attach_devices.argtypes = [vp, C.POINTER(i32), i32, i32]; L.miro_host_attach_devices.restype = i32
    L.miro_host_group.argtypes = [vp]; L.miro_host_group.restype = vp
    L.miro_host_trace.argtypes = [vp, vp, sz, vp]; L.miro_host_trace.restype = i32
    L.miro_host_trace_any.argtypes = [vp, vp, sz, vp]; L.miro_host_trace_any.restype = i32
    L.miro_gpu_group_create.argtypes = [C.POINTER(vp), C.POINTER(i32), i32]; L.miro_gpu_group_create.restype = i32
    L.miro_gpu_group_destroy.argtypes = [vp]; L.miro_gpu_group_destroy.restype = None
    L.miro_gpu_group_size.argtypes = [vp]; L.miro_gpu_group_size.restype = i32
    L.miro_gpu_group_ctx.argtypes = [vp, i32]; L.miro_gpu_group_ctx.restype = vp
    L.miro_gpu_group_last_error.argtypes = [vp]; L.miro_gpu_group_last_error.restype = cp
    L.miro_gpu_group_peer_access.argtypes = [vp, i32]; L.miro_gpu_group_peer_access.restype = i32
    L.miro_gpu_group_upload_scene.argtypes = [vp, C.POINTER(SceneDesc)]; L.miro_gpu_group_upload_scene.restype = i32
    L.miro_gpu_group_render.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), i32, vp, vp]; L.miro_gpu_group_render.restype = i32
    L.miro_gpu_pin_host_buffer.argtypes = [vp, vp, sz]; L.miro_gpu_pin_host_buffer.restype = i32
    L.miro_gpu_unpin_host_buffer.argtypes = [vp, vp]; L.miro_gpu_unpin_host_buffer.restype = i32
    L.miro_gpu_render_image.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), vp, vp]; L.miro_gpu_render_image.restype = i32
    L.miro_gpu_group_trace_closest.argtypes = [vp, vp, sz, vp]; L.miro_gpu_group_trace_closest.restype = i32
    L.miro_gpu_group_trace_any.argtypes = [vp, vp, sz, vp]; L.miro_gpu_group_trace_any.restype = i32
    L.miro_gpu_group_get_counters.argtypes = [vp, C.POINTER(Counters)]; L.miro_gpu_group_get_counters.restype = i32
    L.miro_gpu_group_reset_counters.argtypes = [vp]; L.miro_gpu_group_reset_counters.restype = i32
    L.miro_host_raytrace_image.argtypes = [vp, vp, vp, i32, i32]; L.miro_host_raytrace_image.restype = i32
    L.miro_host_image.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]; L.miro_host_image.restype = i32
    L.miro_host_write_ppm.argtypes = [vp, cp]; L.miro_host_write_ppm.restype = i32
    _lib = L
    return L
