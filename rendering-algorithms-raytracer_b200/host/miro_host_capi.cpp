// miro_host_capi.cpp — extern "C" shim over the C++ host layer (include/miro_host.h).
#include "miro_host.h"
#include "../../include/miro_host.h"
#include <string.h>

using namespace miro;

struct miro_host_scene {
    LoadedScene loaded;
    std::vector<std::unique_ptr<TriangleMesh>> preMeshes;
    std::vector<std::pair<std::string, TriangleMesh*>> preMeshList;
    std::vector<std::unique_ptr<RawImage>> preImages;
    std::vector<std::pair<std::string, RawImage*>> preImageList;
    std::string error;
    bool ready = false;
};

extern "C" {

miro_host_scene* miro_host_new(void) { return new miro_host_scene(); }
void miro_host_free(miro_host_scene* s) { delete s; }
const char* miro_host_error(const miro_host_scene* s) { return s ? s->error.c_str() : "null scene"; }

int miro_host_preload_mesh(miro_host_scene* s, const char* name, const float* vertices, uint32_t nv, const uint32_t* vidx, uint32_t nf,
                           const float* normals, uint32_t nn, const uint32_t* nidx, const float* uvs, uint32_t nt, const uint32_t* tidx) {
    if (!s || !name || !vertices || !vidx) return MIRO_GPU_EINVAL;
    for (uint32_t i = 0; i < 3 * nf; ++i) if (vidx[i] >= nv) { s->error = "preload_mesh: vertex index out of range"; return MIRO_GPU_EINVAL; }
    if (nidx) for (uint32_t i = 0; i < 3 * nf; ++i) if (nidx[i] >= nn) { s->error = "preload_mesh: normal index out of range"; return MIRO_GPU_EINVAL; }
    if (tidx) for (uint32_t i = 0; i < 3 * nf; ++i) if (tidx[i] >= nt) { s->error = "preload_mesh: uv index out of range"; return MIRO_GPU_EINVAL; }
    s->preMeshes.emplace_back(new TriangleMesh);
    s->preMeshes.back()->setGeometry(vertices, nv, vidx, nf, normals, nn, nidx, uvs, nt, tidx);
    s->preMeshList.emplace_back(name, s->preMeshes.back().get());
    return MIRO_GPU_OK;
}

int miro_host_preload_image(miro_host_scene* s, const char* name, const float* texels, int width, int height, int channels, int is_hdr) {
    if (!s || !name || !texels || width <= 0 || height <= 0) return MIRO_GPU_EINVAL;
    ImageType t = channels == 1 ? GRAYSCALE : (channels == 4 ? RGBA : (is_hdr ? HDR : RGB));
    if (channels != 1 && channels != 3 && channels != 4) { s->error = "preload_image: channels must be 1, 3 or 4"; return MIRO_GPU_EINVAL; }
    s->preImages.emplace_back(new RawImage(width, height, texels, t));
    s->preImageList.emplace_back(name, s->preImages.back().get());
    return MIRO_GPU_OK;
}

int miro_host_load_script(miro_host_scene* s, const char* script_path, const char* asset_root) {
    if (!s || !script_path) return MIRO_GPU_EINVAL;
    s->ready = false;
    s->loaded = LoadedScene();
    if (!loadSceneScript(script_path, asset_root, s->loaded, s->error, &s->preMeshList, &s->preImageList)) return MIRO_GPU_EINVAL;
    if (!s->loaded.scene->preCalc()) { s->error = s->loaded.scene->lastError(); return MIRO_GPU_EUNSUPPORTED; }
    s->ready = true;
    return MIRO_GPU_OK;
}

int miro_host_get_desc(const miro_host_scene* s, miro_gpu_scene_desc* out) {
    if (!s || !out || !s->ready) return MIRO_GPU_EINVAL;
    *out = s->loaded.scene->flat().desc();
    return MIRO_GPU_OK;
}
int miro_host_get_camera(const miro_host_scene* s, miro_gpu_camera* out) {
    if (!s || !out || !s->ready) return MIRO_GPU_EINVAL;
    s->loaded.camera->fill(*out);
    return MIRO_GPU_OK;
}
int miro_host_get_render_params(const miro_host_scene* s, miro_gpu_render_params* out) {
    if (!s || !out || !s->ready) return MIRO_GPU_EINVAL;
    s->loaded.scene->renderParams(s->loaded.image.get(), *out);
    return MIRO_GPU_OK;
}
int miro_host_bvh_stats(const miro_host_scene* s, uint32_t* nodes, uint32_t* leaves, uint32_t* max_depth, double* sah_cost) {
    if (!s || !s->ready) return MIRO_GPU_EINVAL;
    const BvhStats& st = s->loaded.scene->flat().top_stats;
    if (nodes) *nodes = st.nodes;
    if (leaves) *leaves = st.leaves;
    if (max_depth) *max_depth = st.max_depth;
    if (sah_cost) *sah_cost = st.sah_cost;
    return MIRO_GPU_OK;
}
int miro_host_attach(miro_host_scene* s, int device) {
    if (!s || !s->ready) return MIRO_GPU_EINVAL;
    if (!s->loaded.scene->attach(device)) { s->error = s->loaded.scene->lastError(); return MIRO_GPU_ECUDA; }
    return MIRO_GPU_OK;
}
int miro_host_attach_devices(miro_host_scene* s, const int* device_ids, int n_devices, int sample_sharding) {
    if (!s || !s->ready || !device_ids || n_devices < 1) return MIRO_GPU_EINVAL;
    s->loaded.scene->setSampleSharding(sample_sharding != 0);
    if (!s->loaded.scene->attachDevices(device_ids, n_devices)) { s->error = s->loaded.scene->lastError(); return MIRO_GPU_ECUDA; }
    return MIRO_GPU_OK;
}
miro_gpu_group* miro_host_group(miro_host_scene* s) { return (s && s->ready) ? s->loaded.scene->group() : nullptr; }
int miro_host_trace(miro_host_scene* s, const miro_gpu_ray* rays, size_t n, miro_gpu_hit* hits) {
    if (!s || !s->ready) return MIRO_GPU_EINVAL;
    if (!s->loaded.scene->trace(rays, n, hits)) { s->error = s->loaded.scene->lastError(); return MIRO_GPU_ECUDA; }
    return MIRO_GPU_OK;
}
int miro_host_trace_any(miro_host_scene* s, const miro_gpu_ray* rays, size_t n, uint32_t* occluded_bits) {
    if (!s || !s->ready) return MIRO_GPU_EINVAL;
    if (!s->loaded.scene->traceAny(rays, n, occluded_bits)) { s->error = s->loaded.scene->lastError(); return MIRO_GPU_ECUDA; }
    return MIRO_GPU_OK;
}
miro_gpu_ctx* miro_host_ctx(miro_host_scene* s) { return (s && s->ready) ? s->loaded.scene->context() : nullptr; }

int miro_host_raytrace_image(miro_host_scene* s, float* rgb, unsigned char* rgb8, int shard_index, int shard_count) {
    if (!s || !s->ready) return MIRO_GPU_EINVAL;
    Image* img = s->loaded.image.get();
    if (!s->loaded.scene->raytraceImage(s->loaded.camera.get(), img, shard_index, shard_count)) { s->error = s->loaded.scene->lastError(); return MIRO_GPU_ECUDA; }
    if (rgb) memcpy(rgb, img->m_radiance.data(), img->m_radiance.size() * sizeof(float));
    if (rgb8) memcpy(rgb8, img->getCharPixels(), (size_t)img->width() * img->height() * 3);
    return MIRO_GPU_OK;
}
int miro_host_image(miro_host_scene* s, const float** rgb, const unsigned char** rgb8, int* width, int* height) {
    if (!s || !s->ready) return MIRO_GPU_EINVAL;
    Image* img = s->loaded.image.get();
    if (rgb) *rgb = img->m_radiance.data();
    if (rgb8) *rgb8 = img->getCharPixels();
    if (width) *width = img->width();
    if (height) *height = img->height();
    return MIRO_GPU_OK;
}
int miro_host_write_ppm(miro_host_scene* s, const char* path) {
    if (!s || !s->ready || !path) return MIRO_GPU_EINVAL;
    s->loaded.image->writePPM(path);
    return MIRO_GPU_OK;
}

}  // extern "C"
