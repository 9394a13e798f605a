/* miro_gpu.h — C ABI of the B200-native ray-casting core for the Miro ray tracer.
 *
 * This is the drop-in boundary.  The reference (bitfrozen/rendering-algorithms-raytracer)
 * has no FFI: its seams are three C++ member functions.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference tree):
 *
 *   Scene::preCalc()                 src/Scene.cpp:63-79  (-> BVH::build, src/BVH.h:130)
 *        after it has run, the host flattens what it built and calls miro_gpu_upload_scene
 *   Scene::trace(tid, hit, ray, tMin) src/Scene.h:32, src/Scene.cpp:295-298 (-> BVH::intersect,
 *        src/BVH.h:147, src/BVH.cpp:1112-1178; intersect4, src/BVH.cpp:1298-1459)
 *        -> miro_gpu_trace_closest / miro_gpu_trace_any (batched)
 *   Scene::raytraceImage(cam, img)    src/Scene.h:31, src/Scene.cpp:86-217
 *        -> miro_gpu_render (float radiance, before Image::Map, src/Image.cpp:71-87)
 *
 * Conventions: plain C, POD structs, host pointers unless a *_device variant says otherwise,
 * no torch / C++ types.  Every call returns 0 (MIRO_GPU_OK) or a negative MIRO_GPU_E* code;
 * the message is available through miro_gpu_last_error().  No exceptions cross the boundary.
 * Buffers passed in are borrowed for the duration of the call only; device memory is owned
 * by the context.  A context is not re-entrant (one host thread at a time); calls are
 * synchronous unless stated.  There is NO CPU fallback: without a CUDA device every compute
 * entry point fails with MIRO_GPU_ENODEVICE.
 */
#ifndef MIRO_GPU_H
#define MIRO_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIRO_GPU_ABI_VERSION 2

enum {
    MIRO_GPU_OK = 0,
    MIRO_GPU_EINVAL = -1,       /* bad argument / malformed scene description */
    MIRO_GPU_ENODEVICE = -2,    /* no CUDA device, or device is not sm_100 */
    MIRO_GPU_ECUDA = -3,        /* CUDA runtime error (message has the details) */
    MIRO_GPU_ENOSCENE = -4,     /* trace/render before upload_scene */
    MIRO_GPU_EUNSUPPORTED = -5, /* scene uses a feature outside the hot-path scope */
    MIRO_GPU_ENOMEM = -6
};

/* ---- constants shared with the reference (src/Miro.h:35-68) ------------------------- */
#define MIRO_GPU_TMAX 1e12f      /* MIRO_TMAX */
#define MIRO_GPU_EPSILON 0.001f  /* epsilon   */

/* ---- rays and hits (reference: Ray, src/Ray.h:27-178; HitInfo, src/Ray.h:185-200) ---- */
typedef struct miro_gpu_ray {   /* 48 bytes, 16-byte aligned: three 16-byte vector loads */
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
    float time;                 /* motion-blur lerp weight, Ray::time */
    uint32_t flags;             /* reserved, 0 */
    uint32_t user0, user1;      /* opaque to the tracer */
} miro_gpu_ray;

/* Packed ray for static scenes: the first two 16-byte words of miro_gpu_ray (time = 0, no flags / user words).  A third less
 * to move over PCIe, which is what bounds the host-pointer calls. */
typedef struct miro_gpu_ray32 {  /* 32 bytes, 16-byte aligned */
    float o[3]; float tmin;
    float d[3]; float tmax;
} miro_gpu_ray32;

typedef struct miro_gpu_hit {   /* 20 bytes */
    float t, a, b;              /* distance; barycentric weights of vertex 1 and vertex 2 (HitInfo::a,b) */
    int32_t prim;               /* index into miro_gpu_scene_desc::prims, -1 = miss */
    int32_t inst;               /* index into ::instances (HitInfo::m_proxy), -1 = none */
} miro_gpu_hit;

/* ---- acceleration structure ------------------------------------------------------------
 * 4-wide BVH node, 128 bytes, 128-byte aligned (one L1/L2 line), SoA bounds as in the
 * reference's QBVH_Node (src/BVH.h:89-96) so its tree flattens 1:1.  All BVHs of a scene
 * (the top level and every instanced bottom level) live in ONE node array.
 * child[i]:  >= 0            index of an inner node
 *            MIRO_GPU_CHILD_EMPTY  unused slot (flagsIsValid[i] == false)
 *            otherwise (bit 31 set) a leaf reference, see MIRO_GPU_LEAF().             */
#define MIRO_GPU_CHILD_EMPTY ((int32_t)0x7fffffff)
/* miro_gpu_scene_desc::root value asking miro_gpu_upload_scene to BUILD the acceleration structure on the GPU (LBVH) over
 * `tris`, given in any order, with nodes == NULL / n_nodes == 0.  Static triangles only (n_mbtris == n_instances == 0,
 * MIRO_GPU_EUNSUPPORTED otherwise).  Hit records keep referring to the caller's triangle order.  Replaces the host
 * BVH::build (src/BVH.cpp:457-1106) for scenes that change every frame; the host-built SAH tree traverses faster. */
#define MIRO_GPU_ROOT_BUILD_ON_DEVICE ((int32_t)0x7ffffffd)
#define MIRO_GPU_KIND_TRI 0u    /* static triangle      (reference: Object,      objectType OBJECT) */
#define MIRO_GPU_KIND_MBTRI 1u  /* motion-blur triangle (reference: MBObject,    MB_OBJECT)         */
#define MIRO_GPU_KIND_INST 2u   /* instance             (reference: ProxyObject, PROXY_OBJECT)      */
#define MIRO_GPU_MAX_LEAF 4u    /* MAX_LEAF_SIZE, src/Miro.h:38 */
#define MIRO_GPU_LEAF_INDEX_BITS 26
/* leaf reference: bit31 | kind<<29 | (count-1)<<26 | first ; `first` indexes tris / mbtris / instances */
#define MIRO_GPU_LEAF(kind, first, count) \
    ((int32_t)(0x80000000u | ((uint32_t)(kind) << 29) | (((uint32_t)(count) - 1u) << 26) | (uint32_t)(first)))

typedef struct miro_gpu_node {
    float lo_x[4], lo_y[4], lo_z[4];
    float hi_x[4], hi_y[4], hi_z[4];
    int32_t child[4];
    uint32_t reserved[4];
} miro_gpu_node;

typedef struct miro_gpu_tri {       /* 48 bytes: three vertices, padded to float4 */
    float v0[3]; uint32_t pad0;
    float v1[3]; uint32_t pad1;
    float v2[3]; uint32_t pad2;
} miro_gpu_tri;

typedef struct miro_gpu_mbtri {     /* 96 bytes: pose 1 (time 0) and pose 2 (time 1), src/BVH.cpp:1316-1335 */
    miro_gpu_tri pose[2];
} miro_gpu_mbtri;

typedef struct miro_gpu_instance {  /* 64 bytes: what traversal needs of a ProxyObject */
    float inv[12];                  /* rows 0..2 of M^-1 (row-major 3x4), src/ProxyObject.cpp:78-79 */
    int32_t blas_root;              /* node index of the instanced BVH's root */
    uint32_t ordinal;               /* caller's ProxyObject ordinal (reported back by the host layer, not interpreted) */
    float w_recip;                  /* what Matrix4x4::multiplyAndDivideByW multiplies the transformed origin by
                                       (src/Matrix4x4.h:728-741): recipps(w), w = row 4 of the inverse . [o 1] — for an
                                       affine matrix the inverse's m44 as Matrix4x4::invert rounds it (often 1 - 2^-24) —,
                                       i.e. rcpps(w) + one Newton step ON THE HOST'S SSE UNIT; 0 = not supplied, treated as 1 */
    uint32_t reserved;
} miro_gpu_instance;

/* Shading record of one primitive; prims[0..n_tris) describe tris[], prims[n_tris..n_tris+n_mbtris)
 * describe mbtris[].  (reference: Object{m_material,m_mesh,m_index}, src/Object.h:73-76)        */
typedef struct miro_gpu_prim {      /* 48 bytes */
    uint32_t n[3];                  /* indices into normals[]  (TriangleMesh::m_normalIndices)   */
    uint32_t uv[3];                 /* indices into uvs[], or 0xffffffff: (u,v) = (a,b), src/Ray.cpp:42-48 */
    uint32_t material;
    uint32_t mesh;                  /* caller's mesh ordinal (reported back, not interpreted)     */
    uint32_t tri;                   /* Object::m_index                                             */
    uint32_t reserved[3];
} miro_gpu_prim;

/* ---- materials (reference: Lambert src/Lambert.cpp:19-53, Blinn src/Blinn.cpp:39-236,335) ---- */
#define MIRO_GPU_MAT_LAMBERT 0u
#define MIRO_GPU_MAT_BLINN 1u
typedef struct miro_gpu_material {  /* 128 bytes */
    uint32_t kind;
    float kd[3];
    float ka[3];
    float ks[3];
    float spec_exp, spec_amt;
    float emit_intensity;           /* Blinn::m_lightEmitted */
    float le[3];                    /* Blinn::m_Le           */
    int32_t color_map;              /* texture index or -1   */
    int32_t alpha_map;              /* texture index or -1: hits where its alpha channel reads < 0.5 are ignored by every trace (src/BVH.cpp:1401-1435) */
    float reflect_amt, refract_amt; /* Blinn::m_reflectAmt / m_refractAmt: mirror reflection / refraction, chosen by Fresnel-weighted
                                       Russian roulette (src/Blinn.cpp:188-204,238-331) */
    float spec_gloss;               /* Blinn::m_specGloss: < 1 blends the reflection vector with a cosine sample (src/Blinn.cpp:160-165) */
    float translucency;             /* Material::m_translucency: > 0.01 adds the lights seen from the back side (src/Blinn.cpp:223-236) */
    uint32_t sample_env;            /* Material::m_sampleEnv */
    float ior[3];                   /* Blinn::m_ior[0..2]; a non-dispersive material refracts with ior[1] (src/Blinn.cpp:183) */
    uint32_t disperse;              /* Material::m_disperse: a refraction splits into one ray per colour channel with ior[0..2] (src/Blinn.cpp:275-302) */
    /* Blinn only (src/Blinn.cpp:120-142), texture index or -1.  normal_map: N = texel.x*T + texel.y*BT + texel.z*N with the
     * texel as stored (no [0,1] -> [-1,1] remap, no renormalisation — the reference's behaviour); the other three scale
     * spec_amt / reflect_amt / refract_amt by the mean of the texel's RGB. */
    int32_t normal_map, specular_map, reflect_map, refract_map;
    uint32_t reserved;
} miro_gpu_material;

/* ---- lights (reference: src/PointLight.cpp:8-82, src/RectangleLight.cpp:14-137, src/DomeLight.cpp:8-161) */
#define MIRO_GPU_LIGHT_POINT 0u
#define MIRO_GPU_LIGHT_RECT 1u
#define MIRO_GPU_LIGHT_DOME 2u
#define MIRO_GPU_MAX_LIGHTS 8u
typedef struct miro_gpu_light {     /* 64 bytes */
    uint32_t kind;
    float p0[3];                    /* point: position; rect: v1 */
    float p1[3];                    /* rect: v2 */
    float p2[3];                    /* rect: v3 */
    float power;                    /* point: m_power; rect: m_power AFTER setPower's 1/area (RectangleLight.cpp:39); dome: gain */
    int32_t num_samples;            /* Light::m_numSamples */
    float noise_threshold;          /* Light::m_noiseThreshold */
    uint32_t cast_shadows;
    int32_t texture;                /* dome: light map texture index */
    uint32_t full_shadows;          /* 0: Light::m_fastShadows (the default, src/Light.h:16) — a shadow ray is an any-hit query.
                                       1: setFastShadows(false), the "full method" of src/PointLight.cpp:49-70, RectangleLight.cpp:93-118,
                                       DomeLight.cpp:123-146: the ray is walked hit by hit and attenuated by the refractAmt of every
                                       surface it enters (was `reserved`, must-be-zero, up to ABI v2: same layout) */
} miro_gpu_light;

/* ---- textures (reference: src/Texture.cpp:12-125; float texels, row-major, already linearised) */
typedef struct miro_gpu_texture {
    const float* texels;            /* width*height*channels floats */
    int32_t width, height, channels; /* channels: 1 (GRAYSCALE), 3 (RGB / HDR), 4 (RGBA) */
    int32_t reserved;
} miro_gpu_texture;

typedef struct miro_gpu_scene_desc {
    uint32_t abi_version;           /* MIRO_GPU_ABI_VERSION */
    /* acceleration structure */
    const miro_gpu_node* nodes;       uint32_t n_nodes;
    int32_t root;                   /* child-style reference to the top-level root: a node index, or a leaf reference
                                       when the whole scene is a single leaf (src/BVH.cpp:118-132) */
    const miro_gpu_tri* tris;         uint32_t n_tris;
    const miro_gpu_mbtri* mbtris;     uint32_t n_mbtris;
    const miro_gpu_instance* instances; uint32_t n_instances;
    /* shading data */
    const miro_gpu_prim* prims;       /* n_tris + n_mbtris records */
    const float* normals;             uint32_t n_normals;   /* xyz triples */
    const float* uvs;                 uint32_t n_uvs;       /* uv pairs (may be NULL/0) */
    const float* inst_normal_xform;   /* n_instances x 9: rows 0..2 of (M^-1)^T, src/Ray.cpp:27-31 */
    const float* tangents;            /* n_normals xyz triples, indexed by a primitive's NORMAL indices (src/Ray.cpp:22,35-36; */
    const float* bitangents;          /*  TriangleMesh::preCalc, src/TriangleMesh.cpp:107-150); both may be NULL: T = BT = 0   */
    const miro_gpu_material* materials; uint32_t n_materials;
    const miro_gpu_light* lights;       uint32_t n_lights;
    const miro_gpu_texture* textures;   uint32_t n_textures;
    int32_t env_map;                /* Scene::m_envMap texture index or -1 */
    float env_exposure;             /* Scene::m_envExposure */
    float bg_color[3];              /* Scene::m_BGColor */
} miro_gpu_scene_desc;

/* ---- camera and render parameters (reference: Camera src/Camera.h:26-69; Scene src/Scene.h:40-64) */
typedef struct miro_gpu_camera {
    float eye[3];
    float view_dir[3];              /* normalised, Camera::m_viewDir */
    float up[3];                    /* normalised, Camera::m_up      */
    float fov_deg;
    float focus_plane, aperture, shutter_speed;
} miro_gpu_camera;

typedef struct miro_gpu_render_params {
    int32_t width, height;
    int32_t min_subdivs, max_subdivs;   /* Scene::m_minSubdivs / m_maxSubdivs (src/Scene.cpp:252-293) */
    float noise_threshold;              /* Scene::m_noiseThreshold */
    int32_t num_paths;                  /* Scene::m_numPaths   */
    int32_t max_bounces;                /* Scene::m_maxBounces */
    uint32_t path_trace;                /* Scene::m_pathTrace  */
    uint32_t sample_env;                /* Scene::m_sampleLightFromEnv */
    uint64_t seed;                      /* counter-based RNG seed (replaces the global MT19937, src/Scene.cpp:26-47) */
    /* work sharding (multi-GPU): this call renders the 32x32 buckets b with b % shard_count == shard_index
       (bucket order of src/Scene.cpp:160-175).  shard_count <= 1 renders everything.  Pixels outside the
       shard are left untouched in rgb_out. */
    int32_t shard_index, shard_count;
    /* sample sharding (multi-GPU, path-traced configs): this call traces the paths p of every camera sample with
       p % path_shard_count == path_shard_index, each still weighted 1 / num_paths, so the SUM of the shards' images is the
       whole image (better balance than tiles when parts of the image are empty).  path_shard_count <= 1: all paths.
       Needs min_subdivs == max_subdivs (the adaptive cut-off looks at the complete pixel value): EINVAL otherwise. */
    int32_t path_shard_index, path_shard_count;
    uint32_t reserved[2];
} miro_gpu_render_params;

typedef struct miro_gpu_counters {
    uint64_t rays_closest;      /* closest-hit queries since the last reset (one Scene::trace call = one ray) */
    uint64_t rays_any;          /* any-hit (shadow) queries */
    uint64_t nodes_fetched;     /* 128-byte nodes fetched (only counted when counting is enabled) */
    uint64_t tris_tested;       /* triangles tested */
    uint64_t insts_entered;     /* instance transforms fetched */
    double trace_ms;            /* device time of the traversal kernels (CUDA events), accumulated while counting is enabled */
    double total_ms;            /* device time of whole trace / render calls, accumulated while counting is enabled */
    uint64_t kernel_launches;   /* kernels launched by this context since the last reset */
} miro_gpu_counters;

typedef struct miro_gpu_ctx miro_gpu_ctx;

/* Create a context on one CUDA device (device_id as in cudaSetDevice).  One context per GPU.  Several GPUs are driven either by
 * one process per GPU (torch.distributed: NCCL combines the frame buffers, bench.py) or by ONE caller through a group
 * (miro_gpu_group_*, below). */
int miro_gpu_create(miro_gpu_ctx** out, int device_id);
void miro_gpu_destroy(miro_gpu_ctx* ctx);
const char* miro_gpu_last_error(const miro_gpu_ctx* ctx);   /* ctx may be NULL: last create() error */
int miro_gpu_abi_version(void);
/* sizeof() of the k-th struct of this header, in declaration order (ray, hit, node, tri, mbtri, instance, prim,
 * material, light, texture, scene_desc, camera, render_params, counters); 0 for k out of range.  Lets a binding
 * verify its mirror of the layouts at load time. */
size_t miro_gpu_sizeof(int k);

/* Use an existing CUDA stream (cudaStream_t passed as void*) for all work of this context;
 * NULL restores the context's own stream.  Lets the caller order work against torch streams. */
int miro_gpu_set_stream(miro_gpu_ctx* ctx, void* cuda_stream);

/* Trace chaining (default off).  When on, consecutive miro_gpu_trace_*_device calls of this context are launched with
 * programmatic dependent launch: the next traversal kernel starts filling the SMs while the previous one drains its last rays
 * (each kernel has a start-up ramp and a tail at falling occupancy; on the 3-launch benchmark step this is worth ~20 %).
 * CONTRACT while it is on: the inputs of a trace call (ray buffer, device-side count) must already be complete when the
 * PREVIOUS trace call of the context is issued — i.e. they must not be produced by work enqueued on the stream between the two
 * calls, because the later kernel no longer waits for the complete end of everything before it — and two consecutive calls
 * must not write overlapping output ranges with different values (their order of writing is not defined).  Outputs are complete, as
 * always, for any work enqueued after the call that does not use this mechanism (copies, other kernels, events).
 * miro_gpu_render never chains its own launches.  Timing events (enable_counting) break the chain. */
int miro_gpu_set_trace_chaining(miro_gpu_ctx* ctx, int on);

/* Which traversal kernel serves Scene::trace on this context (all give the same hits, byte for byte; tests run all of them):
 *   MIRO_GPU_KERNEL_WARP  persistent warps, one ray per lane in registers, majority vote per round between a node and a leaf step
 *   MIRO_GPU_KERNEL_POOL  a warp owns a pool of 64 rays in shared memory and advances, per round, up to 32 of them that wait for
 *                         the same kind of step (csrc/trace_pool.cuh)
 *   MIRO_GPU_KERNEL_FLAT  the persistent-warp kernel with the triangle tests of a leaf round dealt out over all 32 lanes of the warp
 *                         (csrc/trace_flat.cuh)
 * The default, MIRO_GPU_KERNEL_AUTO, picks per uploaded scene the kernel measured faster on that kind of scene (DESIGN.md section
 * 3.1): FLAT for scenes of static and motion-blur triangles from 16 384 triangles up, WARP for smaller ones and when the scene has
 * instances or alpha cut-outs.  The environment variable
 * MIRO_GPU_TRACE_KERNEL (auto | warp | pool | flat), read at miro_gpu_create, overrides the default;
 * miro_gpu_get_trace_kernel returns the kernel in effect (never AUTO once a scene is uploaded). */
#define MIRO_GPU_KERNEL_AUTO (-1)
#define MIRO_GPU_KERNEL_WARP 0
#define MIRO_GPU_KERNEL_POOL 1
#define MIRO_GPU_KERNEL_FLAT 2
int miro_gpu_set_trace_kernel(miro_gpu_ctx* ctx, int kind);
int miro_gpu_get_trace_kernel(const miro_gpu_ctx* ctx);

/* Copy a flattened scene to the device (replaces any previous scene of this context). */
int miro_gpu_upload_scene(miro_gpu_ctx* ctx, const miro_gpu_scene_desc* desc);

/* Batched Scene::trace.  Host buffers: copied in, traced, copied out (timed end to end by callers). */
int miro_gpu_trace_closest(miro_gpu_ctx* ctx, const miro_gpu_ray* rays, size_t n, miro_gpu_hit* hits);
/* occluded_bits: (n+31)/32 words, bit i set iff ray i hits anything in [tmin, tmax). */
int miro_gpu_trace_any(miro_gpu_ctx* ctx, const miro_gpu_ray* rays, size_t n, uint32_t* occluded_bits);
/* Same, buffers already resident in device memory (e.g. torch CUDA tensors); asynchronous on the
 * context's stream — the caller synchronises. */
int miro_gpu_trace_closest_device(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, miro_gpu_hit* d_hits);
int miro_gpu_trace_any_device(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, uint32_t* d_occluded_bits);

/* Host buffers of the calls above may be ordinary (pageable) memory, but then every copy is staged by the driver (measured on
 * B200: 8 GB/s against 47 GB/s from pinned memory — the host-pointer calls are bound by PCIe either way).  A caller that reuses
 * its ray / hit arrays (a renderer does) page-locks them once: miro_gpu_pin_host_buffer wraps cudaHostRegister for memory the
 * caller allocated itself (malloc / new / std::vector), miro_gpu_unpin_host_buffer must be called before that memory is freed. */
int miro_gpu_pin_host_buffer(miro_gpu_ctx* ctx, void* ptr, size_t bytes);
int miro_gpu_unpin_host_buffer(miro_gpu_ctx* ctx, void* ptr);

/* The same two queries for packed rays (host pointers; pipelined like miro_gpu_trace_closest / _any).  Results are those of the
 * 48-byte call with time = 0. */
int miro_gpu_trace_closest_packed(miro_gpu_ctx* ctx, const miro_gpu_ray32* rays, size_t n, miro_gpu_hit* hits);
int miro_gpu_trace_any_packed(miro_gpu_ctx* ctx, const miro_gpu_ray32* rays, size_t n, uint32_t* occluded_bits);

/* Primary rays made where they are traced: Camera::eyeRayAdaptive at the pixel centres (src/Camera.cpp:116-174 — the level-1 sample of
 * Scene::adaptiveSampleScene, src/Scene.cpp:254 — with the time sample and lens of miro_gpu_render at the same seed) are generated on
 * the device and traced; only the hit records travel (20 B per ray instead of 48 B up + 20 B down, and PCIe bounds the
 * host-pointer calls).  hits[y * width + x], row 0 = bottom; hits: host or device pointer.  d_rays_out: NULL, or a device
 * buffer of width * height rays that receives the generated rays (origin, direction, tmin = 1e-3, tmax = MIRO_GPU_TMAX, time). */
int miro_gpu_trace_primary(miro_gpu_ctx* ctx, const miro_gpu_camera* cam, int width, int height, uint64_t seed, miro_gpu_hit* hits, miro_gpu_ray* d_rays_out);

/* Scene::raytraceImage.  rgb_out: width*height*3 floats, row 0 = bottom row (src/Image.cpp:150-151),
 * linear radiance before Image::Map.  rgb_out may be a host or a device pointer. */
int miro_gpu_render(miro_gpu_ctx* ctx, const miro_gpu_camera* cam, const miro_gpu_render_params* params, float* rgb_out);
/* The same, also (or only) delivering what the reference's Image holds after the frame: rgb8_out = width*height*3 bytes, every
 * channel through Image::setPixel's Map (src/Image.cpp:71-87: clamp to [0, 1], 32 769-entry 2.2-gamma table) on the device.
 * Either pointer may be NULL (not both), each may be a host or a device pointer. */
int miro_gpu_render_image(miro_gpu_ctx* ctx, const miro_gpu_camera* cam, const miro_gpu_render_params* params, float* rgb_out, unsigned char* rgb8_out);

/* Counters.  enable != 0 switches the traversal kernels to their instrumented variant (node / triangle / instance fetch
 * counts, used for the roofline's algorithmic bytes) and brackets the library's launches with timing events; ray counts
 * and launch counts are always kept. */
int miro_gpu_enable_counting(miro_gpu_ctx* ctx, int enable);
int miro_gpu_get_counters(miro_gpu_ctx* ctx, miro_gpu_counters* out);
int miro_gpu_reset_counters(miro_gpu_ctx* ctx);

/* ---- several GPUs behind one caller ("one ctx owns 1..8 GPUs"; replaces the OpenMP bucket loop of
 * Scene::raytraceImage, src/Scene.cpp:160-175, across devices).  A group holds one context per listed device (the same device
 * may be listed twice: two contexts on one GPU).  The scene is replicated; miro_gpu_group_render deals the frame's 32x32
 * buckets (MIRO_GPU_SHARD_BUCKETS: bucket b of the reference's bucket order to member b % n — any configuration) or the paths of
 * every camera sample (MIRO_GPU_SHARD_SAMPLES: path p to member p % n — path-traced configurations with min_subdivs ==
 * max_subdivs) to the members, and combines the members' frames on the first device with one kernel that reads the others'
 * memory over NVLink peer access (staged by cudaMemcpyPeerAsync where peer access is unavailable).  Random numbers are keyed by
 * (pixel, sample, path), so the frame equals the single-GPU frame (to the rounding of a light loop's accumulation order, which also differs between two
 * single-GPU renders: float atomics).  rgb_out: host pointer,
 * or device pointer on the first device.  The trace calls split the batch into contiguous parts, one per member (host
 * pointers; no exchange).  A group is used from one host thread at a time; it fans out to one worker thread per member
 * for the duration of a call. */
typedef struct miro_gpu_group miro_gpu_group;
#define MIRO_GPU_SHARD_BUCKETS 0
#define MIRO_GPU_SHARD_SAMPLES 1
int miro_gpu_group_create(miro_gpu_group** out, const int* device_ids, int n_devices);
void miro_gpu_group_destroy(miro_gpu_group* g);
int miro_gpu_group_size(const miro_gpu_group* g);
miro_gpu_ctx* miro_gpu_group_ctx(miro_gpu_group* g, int i);          /* member i's context (counters, kernel selection, ...) */
const char* miro_gpu_group_last_error(const miro_gpu_group* g);
int miro_gpu_group_peer_access(const miro_gpu_group* g, int i);      /* 1: the first member reads member i's frame directly */
int miro_gpu_group_upload_scene(miro_gpu_group* g, const miro_gpu_scene_desc* desc);
int miro_gpu_group_render(miro_gpu_group* g, const miro_gpu_camera* cam, const miro_gpu_render_params* params, int sharding, float* rgb_out,
                          unsigned char* rgb8_out /* NULL, or the 8-bit image as miro_gpu_render_image delivers it (host pointer) */);
int miro_gpu_group_trace_closest(miro_gpu_group* g, const miro_gpu_ray* rays, size_t n, miro_gpu_hit* hits);
int miro_gpu_group_trace_any(miro_gpu_group* g, const miro_gpu_ray* rays, size_t n, uint32_t* occluded_bits);
int miro_gpu_group_get_counters(miro_gpu_group* g, miro_gpu_counters* out);      /* sums over members (times: the slowest member) */
int miro_gpu_group_reset_counters(miro_gpu_group* g);

#ifdef __cplusplus
}
#endif
#endif /* MIRO_GPU_H */
