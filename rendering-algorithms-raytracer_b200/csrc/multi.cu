// multi.cu — one caller, several GPUs: miro_gpu_group_* (include/miro_gpu.h).
//
// The reference is ONE process whose frame entry is Scene::raytraceImage (src/Scene.h:31); its bucket loop (src/Scene.cpp:160-175)
// deals 32x32 buckets to OpenMP threads.  A group deals the same buckets (or the paths of every camera sample) to the GPUs of one
// box: the scene is replicated, every member renders its share into a frame in ITS OWN memory with miro_gpu_render's shard
// parameters, and the shares are combined on the first member's GPU by ONE kernel that reads the other members' frames through
// peer memory (NVLink / NVSwitch loads; 1 / N of the frame from each peer for bucket sharding) — the only exchange of the path.
// Where peer access is not available the frames are staged with cudaMemcpyPeerAsync first.  Batched Scene::trace over a group
// splits the ray array into contiguous parts, one per member, with no exchange at all.
//
// Host side: the calling thread fans out to one worker thread per member for the duration of a call (miro_gpu_render is
// host-driven and synchronous per context); contexts stay single-threaded as the ABI requires.
#include <cuda_runtime.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>
#include "context.cuh"

using namespace miro;

struct miro_gpu_group {
    std::vector<miro_gpu_ctx*> ctx;
    std::vector<int> device;
    std::vector<float*> frame;          // per member: w * h * 3 floats in that member's device memory
    std::vector<float*> staged;         // on member 0's device, for members it cannot read directly
    std::vector<char> peer_ok;          // member 0 can load from member i's memory
    float* combined = nullptr;          // on member 0's device
    unsigned char* combined8 = nullptr;
    size_t frame_pixels = 0;
    std::string error;
};

namespace {

constexpr int MAX_MEMBERS = 16;
struct FramePtrs { const float* p[MAX_MEMBERS]; };

// out[pixel] = the frame of the member that owns the pixel's 32x32 bucket (bucket order of src/Scene.cpp:160-175: row-major over
// buckets, bucket b belongs to member b % n) — the gather of the owned tiles, as loads over peer memory.
__global__ void k_gather_buckets(FramePtrs f, int n, int W, int H, float* __restrict__ out) {
    const int nbx = (W + 31) / 32;
    const size_t total = (size_t)W * H;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(p % W), y = (int)(p / W);
        const int owner = ((y >> 5) * nbx + (x >> 5)) % n;
        const float* src = f.p[owner] + p * 3;
        out[p * 3] = src[0]; out[p * 3 + 1] = src[1]; out[p * 3 + 2] = src[2];
    }
}

// out = sum over members, in member order (deterministic): sample sharding, every member holds the whole frame at weight
// (its paths) / numPaths.
__global__ void k_sum_frames(FramePtrs f, int n, size_t floats, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < floats; i += (size_t)gridDim.x * blockDim.x) {
        float s = f.p[0][i];
        for (int k = 1; k < n; ++k) s += f.p[k][i];
        out[i] = s;
    }
}

int group_fail(miro_gpu_group* g, int code, const std::string& msg) { if (g) g->error = msg; return code; }

// run fn(member index) on one thread per member; returns the first non-zero result
template <class F>
int for_each_member(miro_gpu_group* g, F fn) {
    const int n = (int)g->ctx.size();
    std::vector<int> rc(n, 0);
    if (n == 1) { rc[0] = fn(0); }
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n; ++i) th.emplace_back([&, i] { cudaSetDevice(g->device[i]); rc[i] = fn(i); });
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < n; ++i) if (rc[i]) { g->error = std::string("member ") + std::to_string(i) + " (device " + std::to_string(g->device[i]) + "): " + miro_gpu_last_error(g->ctx[i]); return rc[i]; }
    return MIRO_GPU_OK;
}

int ensure_frames(miro_gpu_group* g, size_t pixels) {
    if (g->frame_pixels >= pixels) return MIRO_GPU_OK;
    const int n = (int)g->ctx.size();
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(g->device[i]);
        if (g->frame[i]) cudaFree(g->frame[i]);
        g->frame[i] = nullptr;
        if (cudaMalloc((void**)&g->frame[i], pixels * 3 * sizeof(float)) != cudaSuccess) return group_fail(g, MIRO_GPU_ENOMEM, "group: frame allocation failed");
    }
    cudaSetDevice(g->device[0]);
    for (int i = 0; i < n; ++i) {
        if (g->staged[i]) cudaFree(g->staged[i]);
        g->staged[i] = nullptr;
        if (!g->peer_ok[i] && cudaMalloc((void**)&g->staged[i], pixels * 3 * sizeof(float)) != cudaSuccess) return group_fail(g, MIRO_GPU_ENOMEM, "group: staging allocation failed");
    }
    if (g->combined) cudaFree(g->combined);
    if (g->combined8) cudaFree(g->combined8);
    g->combined = nullptr; g->combined8 = nullptr;
    if (cudaMalloc((void**)&g->combined, pixels * 3 * sizeof(float)) != cudaSuccess || cudaMalloc((void**)&g->combined8, pixels * 3) != cudaSuccess)
        return group_fail(g, MIRO_GPU_ENOMEM, "group: frame allocation failed");
    g->frame_pixels = pixels;
    return MIRO_GPU_OK;
}

}  // namespace

extern "C" {

int miro_gpu_group_create(miro_gpu_group** out, const int* device_ids, int n_devices) {
    if (!out) return MIRO_GPU_EINVAL;
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > MAX_MEMBERS) return set_error(nullptr, MIRO_GPU_EINVAL, "miro_gpu_group_create: 1..16 devices");
    miro_gpu_group* g = new miro_gpu_group();
    for (int i = 0; i < n_devices; ++i) {
        miro_gpu_ctx* c = nullptr;
        const int rc = miro_gpu_create(&c, device_ids[i]);
        if (rc) { for (miro_gpu_ctx* k : g->ctx) miro_gpu_destroy(k); delete g; return rc; }      // message: miro_gpu_last_error(NULL)
        g->ctx.push_back(c); g->device.push_back(device_ids[i]);
    }
    g->frame.assign(n_devices, nullptr); g->staged.assign(n_devices, nullptr); g->peer_ok.assign(n_devices, 0);
    // member 0 combines: it needs to read the other members' frames
    cudaSetDevice(g->device[0]);
    for (int i = 0; i < n_devices; ++i) {
        if (g->device[i] == g->device[0]) { g->peer_ok[i] = 1; continue; }       // the same GPU (two contexts on one device): plain loads
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, g->device[0], g->device[i]) == cudaSuccess && can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(g->device[i], 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) g->peer_ok[i] = 1;
            cudaGetLastError();
        }
    }
    *out = g;
    return MIRO_GPU_OK;
}

void miro_gpu_group_destroy(miro_gpu_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->ctx.size(); ++i) { cudaSetDevice(g->device[i]); if (g->frame[i]) cudaFree(g->frame[i]); }
    cudaSetDevice(g->device[0]);
    for (float* p : g->staged) if (p) cudaFree(p);
    if (g->combined) cudaFree(g->combined);
    if (g->combined8) cudaFree(g->combined8);
    for (miro_gpu_ctx* c : g->ctx) miro_gpu_destroy(c);
    delete g;
}

int miro_gpu_group_size(const miro_gpu_group* g) { return g ? (int)g->ctx.size() : 0; }
miro_gpu_ctx* miro_gpu_group_ctx(miro_gpu_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }
const char* miro_gpu_group_last_error(const miro_gpu_group* g) { return g ? g->error.c_str() : miro_gpu_last_error(nullptr); }
int miro_gpu_group_peer_access(const miro_gpu_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? (int)g->peer_ok[i] : 0; }

int miro_gpu_group_upload_scene(miro_gpu_group* g, const miro_gpu_scene_desc* desc) {
    if (!g || !desc) return MIRO_GPU_EINVAL;
    return for_each_member(g, [&](int i) { return miro_gpu_upload_scene(g->ctx[i], desc); });
}

int miro_gpu_group_render(miro_gpu_group* g, const miro_gpu_camera* cam, const miro_gpu_render_params* rp, int sharding, float* rgb_out, unsigned char* rgb8_out) {
    if (!g || !cam || !rp || !rgb_out) return MIRO_GPU_EINVAL;
    const int n = (int)g->ctx.size();
    if (rp->shard_count > 1 || rp->path_shard_count > 1) return group_fail(g, MIRO_GPU_EINVAL, "miro_gpu_group_render: the group shards the frame itself (pass shard counts of 0 / 1)");
    if (sharding != MIRO_GPU_SHARD_BUCKETS && sharding != MIRO_GPU_SHARD_SAMPLES) return group_fail(g, MIRO_GPU_EINVAL, "miro_gpu_group_render: unknown sharding");
    if (rp->width <= 0 || rp->height <= 0) return group_fail(g, MIRO_GPU_EINVAL, "bad image size");
    const size_t pixels = (size_t)rp->width * rp->height;
    int rc = ensure_frames(g, pixels);
    if (rc) return rc;
    rc = for_each_member(g, [&](int i) {
        miro_gpu_render_params p = *rp;
        if (sharding == MIRO_GPU_SHARD_SAMPLES) { p.path_shard_index = i; p.path_shard_count = n; }
        else { p.shard_index = i; p.shard_count = n; }
        return miro_gpu_render(g->ctx[i], cam, &p, g->frame[i]);       // synchronous: the member's share is complete on return
    });
    if (rc) return rc;
    // ---- the one exchange: combine on member 0's GPU
    cudaSetDevice(g->device[0]);
    cudaStream_t s = g->ctx[0]->stream;
    FramePtrs f;
    for (int i = 0; i < n; ++i) {
        if (g->peer_ok[i]) f.p[i] = g->frame[i];
        else {
            if (cudaMemcpyPeerAsync(g->staged[i], g->device[0], g->frame[i], g->device[i], pixels * 3 * sizeof(float), s) != cudaSuccess) return group_fail(g, MIRO_GPU_ECUDA, "group: cudaMemcpyPeerAsync failed");
            f.p[i] = g->staged[i];
        }
    }
    cudaPointerAttributes attr;
    const bool out_is_device = cudaPointerGetAttributes(&attr, rgb_out) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    float* target = out_is_device ? rgb_out : g->combined;
    const int grid = g->ctx[0]->sm_count * 8;
    if (n == 1) { if (cudaMemcpyAsync(target, g->frame[0], pixels * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess) return group_fail(g, MIRO_GPU_ECUDA, "group: frame copy failed"); }
    else if (sharding == MIRO_GPU_SHARD_SAMPLES) k_sum_frames<<<grid, 256, 0, s>>>(f, n, pixels * 3, target);
    else k_gather_buckets<<<grid, 256, 0, s>>>(f, n, rp->width, rp->height, target);
    g->ctx[0]->launches++;
    if (rgb8_out) {      // Image::setPixel over the combined frame, on the device
        const int mrc = map_frame_to_bytes(g->ctx[0], target, pixels, g->combined8, s);
        if (mrc) return group_fail(g, mrc, miro_gpu_last_error(g->ctx[0]));
        if (cudaMemcpyAsync(rgb8_out, g->combined8, pixels * 3, cudaMemcpyDeviceToHost, s) != cudaSuccess) return group_fail(g, MIRO_GPU_ECUDA, "group: frame download failed");
    }
    if (!out_is_device && cudaMemcpyAsync(rgb_out, g->combined, pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess) return group_fail(g, MIRO_GPU_ECUDA, "group: frame download failed");
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return group_fail(g, MIRO_GPU_ECUDA, std::string("group: combine failed: ") + cudaGetErrorString(e));
    return MIRO_GPU_OK;
}

// Batched Scene::trace over the group: member i traces rays [i * per, (i + 1) * per) (per rounded to 32 so that any-hit result
// words are not shared between members); no exchange, the results land in the caller's arrays.
static size_t part_size(size_t n, int members) { const size_t per = (n + members - 1) / members; return (per + 31) / 32 * 32; }

int miro_gpu_group_trace_closest(miro_gpu_group* g, const miro_gpu_ray* rays, size_t n, miro_gpu_hit* hits) {
    if (!g || (n && (!rays || !hits))) return MIRO_GPU_EINVAL;
    const size_t per = part_size(n, (int)g->ctx.size());
    return for_each_member(g, [&](int i) {
        const size_t a = std::min(n, per * i), b = std::min(n, per * (i + 1));
        return a < b ? miro_gpu_trace_closest(g->ctx[i], rays + a, b - a, hits + a) : MIRO_GPU_OK;
    });
}

int miro_gpu_group_trace_any(miro_gpu_group* g, const miro_gpu_ray* rays, size_t n, uint32_t* occluded_bits) {
    if (!g || (n && (!rays || !occluded_bits))) return MIRO_GPU_EINVAL;
    const size_t per = part_size(n, (int)g->ctx.size());
    return for_each_member(g, [&](int i) {
        const size_t a = std::min(n, per * i), b = std::min(n, per * (i + 1));
        return a < b ? miro_gpu_trace_any(g->ctx[i], rays + a, b - a, occluded_bits + a / 32) : MIRO_GPU_OK;
    });
}

int miro_gpu_group_get_counters(miro_gpu_group* g, miro_gpu_counters* out) {
    if (!g || !out) return MIRO_GPU_EINVAL;
    miro_gpu_counters sum = {};
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        miro_gpu_counters c;
        cudaSetDevice(g->device[i]);
        const int rc = miro_gpu_get_counters(g->ctx[i], &c);
        if (rc) return group_fail(g, rc, miro_gpu_last_error(g->ctx[i]));
        sum.rays_closest += c.rays_closest; sum.rays_any += c.rays_any; sum.nodes_fetched += c.nodes_fetched; sum.tris_tested += c.tris_tested;
        sum.insts_entered += c.insts_entered; sum.kernel_launches += c.kernel_launches;
        sum.trace_ms = std::max(sum.trace_ms, c.trace_ms); sum.total_ms = std::max(sum.total_ms, c.total_ms);
    }
    *out = sum;
    return MIRO_GPU_OK;
}

int miro_gpu_group_reset_counters(miro_gpu_group* g) {
    if (!g) return MIRO_GPU_EINVAL;
    for (size_t i = 0; i < g->ctx.size(); ++i) { cudaSetDevice(g->device[i]); const int rc = miro_gpu_reset_counters(g->ctx[i]); if (rc) return rc; }
    return MIRO_GPU_OK;
}

}  // extern "C"
