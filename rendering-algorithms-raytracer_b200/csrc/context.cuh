// context.cuh — the device context behind miro_gpu_ctx (shared by the API and the renderer).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/miro_gpu.h"
#include "traverse.cuh"

namespace miro {

// Device-side view of everything shading needs (the traversal subset is DeviceScene).
struct DeviceTexture {
    const float* texels;
    int32_t width, height, channels, pad;
};

struct DeviceDome {            // per DomeLight: alias table over the nu x nv cells (DomeLight.cpp:8-78)
    const float2* alias;       // {acceptance threshold, alias index as float bits}
    const float4* cell_E;      // per cell: gain-free  L(dir)/pdf  (rgb), w = 1 if usable else 0
    const float* cos_u; const float* sin_u;   // nu+1 entries (DomeLight.cpp:59-66)
    const float* cos_v; const float* sin_v;   // nv+1 entries (DomeLight.cpp:68-76)
    int32_t nu, nv;
};

struct DeviceShading {
    const miro_gpu_prim* prims;
    const float* normals;
    const float* tangents;     // n_normals x 3, indexed like normals; NULL: T = BT = 0
    const float* bitangents;
    const float* uvs;
    const float* inst_nxf;     // n_instances x 9
    const miro_gpu_material* materials;
    const miro_gpu_light* lights;
    const DeviceTexture* textures;
    const DeviceDome* domes;   // indexed by light ordinal (unused slots zero)
    uint32_t n_lights, n_materials, n_textures, n_prims;
    int32_t env_map;
    float env_exposure;
    float bg[3];
};

template <class T>
struct DeviceBuffer {
    T* ptr = nullptr;
    size_t cap = 0;   // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; cap = 0;
        size_t want = n + n / 4;
        cudaError_t e = cudaMalloc((void**)&ptr, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};

struct EventPair { cudaEvent_t a, b; bool trace; };
struct PoolScratch { unsigned long long* ptr = nullptr; size_t entries = 0; };      // stack overflow of the pool kernel's slots

}  // namespace miro

struct miro_gpu_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string error;
    bool has_scene = false;
    bool counting = false;
    bool chain_traces = false;    // miro_gpu_set_trace_chaining: consecutive *_device trace launches may overlap (PDL)
    bool in_api_trace = false;    // set around the launches of miro_gpu_trace_*_device (the only ones that may chain)
    bool has_alpha = false;       // some material has an alpha map with an alpha channel: traversal kernels evaluate cut-outs

    // scene storage (device)
    std::vector<void*> scene_allocs;
    miro::DeviceScene scene{};
    miro::DeviceShading shading{};
    uint32_t n_nodes = 0, n_tris = 0, n_mbtris = 0, n_insts = 0;
    std::vector<miro_gpu_light> host_lights;
    std::vector<miro_gpu_material> host_materials;

    // counters
    miro::TraceCounters* d_counters = nullptr;
    uint32_t* d_work = nullptr;                 // [0] next unclaimed ray of the running traversal kernel, [1] blocks that have left
    int sm_count = 148;
    uint64_t work_slot[4] = {0, 0, 0, 0};       // next pair of each lane's work-counter ring
    int work_lane = 0;                          // which ring the next traversal launch uses: callers that spread launches over
                                                // several streams give every stream its own lane (launches of one lane are
                                                // stream-ordered or PDL-chained, so a ring never wraps onto a live launch)
    int build_levels = 0;                       // depth of the last device-built wide tree
    int stack_need = 0;                         // deepest traversal stack the uploaded trees can ask for (entries)
    int trace_kernel_request = MIRO_GPU_KERNEL_AUTO;   // miro_gpu_set_trace_kernel / MIRO_GPU_TRACE_KERNEL
    int trace_kernel = MIRO_GPU_KERNEL_FLAT;    // the kernel in effect (resolve_trace_kernel: the request, or per scene when it is AUTO)
    miro::PoolScratch pool_ovf[4];              // per work lane
    std::vector<miro::EventPair> events;        // pending (not yet summed) timing pairs
    std::vector<miro::EventPair> event_pool;
    double trace_ms = 0.0, total_ms = 0.0;
    uint64_t launches = 0;

    // scratch for the host-pointer trace entry points
    miro::DeviceBuffer<miro_gpu_ray> d_rays;
    miro::DeviceBuffer<miro_gpu_hit> d_hits;
    miro::DeviceBuffer<uint32_t> d_bits;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;      // H2D / D2H streams of the pipelined host-pointer calls
    cudaStream_t trace_aux[3] = {nullptr, nullptr, nullptr}; // further kernel streams of those calls (chunk kernels overlap their tails)
    cudaEvent_t fork_event = nullptr;
    std::vector<cudaEvent_t> pipe_events;

    // renderer state (render.cu)
    void* render_state = nullptr;
};

namespace miro {

int set_error(miro_gpu_ctx* ctx, int code, const std::string& msg);
int cuda_fail(miro_gpu_ctx* ctx, cudaError_t e, const char* what);
#define MIRO_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return miro::cuda_fail(ctx, e__, #call); } while (0)

// timing helpers: bracket a group of launches with events on ctx->stream
EventPair begin_timing(miro_gpu_ctx* ctx, bool trace);
void end_timing(miro_gpu_ctx* ctx, EventPair p);
void drain_timing(miro_gpu_ctx* ctx);

// launches the traversal kernel over device buffers (count either n, or *d_count when non-null)
void launch_trace_closest(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, const uint32_t* d_count, miro_gpu_hit* d_hits);
void launch_trace_any(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, const uint32_t* d_count, uint32_t* d_bits);
void launch_trace_closest_packed(miro_gpu_ctx* ctx, const miro_gpu_ray32* d_rays, size_t n, miro_gpu_hit* d_hits);
void launch_trace_any_packed(miro_gpu_ctx* ctx, const miro_gpu_ray32* d_rays, size_t n, uint32_t* d_bits);
// any-hit traversal of shadow rays; an unoccluded ray adds d_E[i] to d_slots[4 * ray.user0] (see render.cu)
void launch_trace_shadow(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, const uint32_t* d_count, const float4* d_E, float4* d_slots);

void resolve_trace_kernel(miro_gpu_ctx* ctx);
void render_state_free(miro_gpu_ctx* ctx);
int map_frame_to_bytes(miro_gpu_ctx* ctx, const float* d_rgb, size_t pixels, unsigned char* d_rgb8, cudaStream_t s);      // render.cu: Image::setPixel on the device

// build.cu: LBVH over static triangles on the device
int build_lbvh_on_device(miro_gpu_ctx* ctx, const float4* d_tris_in, uint32_t n, const DeviceNode** out_nodes, uint32_t* out_n_nodes,
                         int32_t* out_root, const float4** out_tris, const uint32_t** out_perm);
int reorder_prims_on_device(miro_gpu_ctx* ctx, const miro_gpu_prim* d_in, const uint32_t* d_perm, uint32_t n, const miro_gpu_prim** out);

}  // namespace miro
