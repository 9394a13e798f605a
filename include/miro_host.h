/* miro_host.h — C view of the product's C++ host layer (rendering-algorithms-raytracer_b200/host),
 * for bindings (the Python tests/bench use it through ctypes).  The host layer mirrors the
 * reference's scene API (Scene/Camera/Image/TriangleMesh/Material/Light, src/ headers); a scene is
 * described by a ".miro" script (one reference API call per line, see host/miro_script.cpp).
 * All GPU work goes through include/miro_gpu.h; nothing here renders on the CPU.               */
#ifndef MIRO_HOST_H
#define MIRO_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "miro_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct miro_host_scene miro_host_scene;

/* Begin a scene; meshes / images registered before load_script replace the files named in the script
 * (arrays are copied).  nidx/normals and tidx/uvs may be NULL. */
miro_host_scene* miro_host_new(void);
void miro_host_free(miro_host_scene* s);
const char* miro_host_error(const miro_host_scene* s);
int miro_host_preload_mesh(miro_host_scene* s, const char* name, const float* vertices, uint32_t nv, const uint32_t* vidx, uint32_t nf,
                           const float* normals, uint32_t nn, const uint32_t* nidx, const float* uvs, uint32_t nt, const uint32_t* tidx);
int miro_host_preload_image(miro_host_scene* s, const char* name, const float* texels, int width, int height, int channels, int is_hdr);
/* Parse the script, run Scene::preCalc (BVH build + flatten).  Host only — works without a GPU. */
int miro_host_load_script(miro_host_scene* s, const char* script_path, const char* asset_root);
/* The flattened scene (valid until the scene is freed / reloaded) and the derived call arguments. */
int miro_host_get_desc(const miro_host_scene* s, miro_gpu_scene_desc* out);
int miro_host_get_camera(const miro_host_scene* s, miro_gpu_camera* out);
int miro_host_get_render_params(const miro_host_scene* s, miro_gpu_render_params* out);
int miro_host_bvh_stats(const miro_host_scene* s, uint32_t* nodes, uint32_t* leaves, uint32_t* max_depth, double* sah_cost);
/* Scene::attach: create the GPU context on `device` and upload.  Fails loudly without a GPU. */
int miro_host_attach(miro_host_scene* s, int device);
/* The same over several GPUs of one box (miro_gpu_group_*, include/miro_gpu.h): the scene is replicated, miro_host_raytrace_image
 * deals the frame's buckets (sample_sharding != 0: the paths of every camera sample) to the devices and combines the frame on
 * the first; miro_host_trace / _trace_any (Scene::trace, src/Scene.h:32, batched) split their batch over the devices. */
int miro_host_attach_devices(miro_host_scene* s, const int* device_ids, int n_devices, int sample_sharding);
miro_gpu_group* miro_host_group(miro_host_scene* s);
int miro_host_trace(miro_host_scene* s, const miro_gpu_ray* rays, size_t n, miro_gpu_hit* hits);
int miro_host_trace_any(miro_host_scene* s, const miro_gpu_ray* rays, size_t n, uint32_t* occluded_bits);
miro_gpu_ctx* miro_host_ctx(miro_host_scene* s);
/* Scene::raytraceImage (src/Scene.h:31): the frame lands in the scene's Image — 8-bit pixels through Image::Map and the float
 * radiance, both written by the GPU into page-locked buffers — and is copied to rgb (width*height*3 floats, row 0 = bottom) and
 * rgb8 (bytes) where those are not NULL.  miro_host_image hands out the Image's own buffers (valid until the next resize). */
int miro_host_raytrace_image(miro_host_scene* s, float* rgb, unsigned char* rgb8, int shard_index, int shard_count);
int miro_host_image(miro_host_scene* s, const float** rgb, const unsigned char** rgb8, int* width, int* height);
/* Standalone BVH build over triangle soup (testing the builder): returns node count, fills order (n entries). */
int miro_host_write_ppm(miro_host_scene* s, const char* path);

#ifdef __cplusplus
}
#endif
#endif
