// miro_gpu_api.cu — C ABI implementation: context, scene upload, batched Scene::trace, counters.
// See include/miro_gpu.h for the reference interfaces each entry point replaces.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <string>
#include <vector>
#include "context.cuh"
#include "dome.cuh"
#include "trace_pool.cuh"
#include "trace_flat.cuh"

using namespace miro;

static std::string g_create_error;

namespace miro {

int set_error(miro_gpu_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->error = msg; else g_create_error = msg;
    return code;
}
int cuda_fail(miro_gpu_ctx* ctx, cudaError_t e, const char* what) {
    return set_error(ctx, MIRO_GPU_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// Device timing of the library's launches (miro_gpu_counters::trace_ms / total_ms) is kept only while counting is enabled:
// an event record between two traversal launches would break their programmatic-dependent-launch chaining.
EventPair begin_timing(miro_gpu_ctx* ctx, bool trace) {
    EventPair p;
    if (!ctx->counting) { p.a = p.b = nullptr; p.trace = trace; return p; }
    if (!ctx->event_pool.empty()) { p = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
    else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
    p.trace = trace;
    cudaEventRecord(p.a, ctx->stream);
    return p;
}
void end_timing(miro_gpu_ctx* ctx, EventPair p) {
    if (!p.a) return;
    cudaEventRecord(p.b, ctx->stream);
    ctx->events.push_back(p);
    if (ctx->events.size() > 4096) drain_timing(ctx);
}
void drain_timing(miro_gpu_ctx* ctx) {
    for (EventPair& p : ctx->events) {
        if (cudaEventSynchronize(p.b) == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
                if (p.trace) ctx->trace_ms += ms; else ctx->total_ms += ms;
            }
        }
        ctx->event_pool.push_back(p);
    }
    ctx->events.clear();
}

// ---------------------------------------------------------------------------------------------
// Traversal kernels.  One thread per ray; a warp covers 32 consecutive rays, so the any-hit variant
// can emit one packed word per warp.  Rays / hits are streamed (ld.cs / st.cs) so they do not
// displace BVH nodes from L1/L2.
// MODE: 0 closest hit -> hit records; 1 any hit -> one bit per ray; 2 any hit -> the unoccluded ray's light sample
// (sample_E[i] = E.rgb, specular input) is added to the accumulator of its light loop (slot index in ray.user0).
enum { TRACE_CLOSEST = 0, TRACE_ANY_BITS = 1, TRACE_ANY_ACCUM = 2 };
// rays are claimed from a 32-bit work counter that every warp keeps advancing by 32 after the batch is exhausted (a few million
// past n at full occupancy): n stays 2^24 below the wrap
constexpr unsigned long long MAX_RAYS_PER_CALL = 0xff000000ull;
constexpr int WORK_RING = 32;     // pairs of work counters per lane; launch k of a lane uses pair k % WORK_RING and re-arms it when its last block leaves
constexpr int WORK_LANES = 4;     // lanes = streams a caller may spread traversal launches over (miro_gpu_ctx::work_lane)
// The pool kernel keeps the deep end of its traversal stacks in a scratch in global memory.  Chained launches overlap, so
// consecutive launches of a lane must not share it: SCRATCH_COPIES regions used round-robin, and the chain is broken every
// SCRATCH_COPIES-th launch (that launch waits for the complete end of everything before it), so the launches that can be live
// together always hold different regions.
constexpr int SCRATCH_COPIES = 3;

// PACKED: rays are miro_gpu_ray32 records (two 16-byte words: o, tmin | d, tmax; time = 0) instead of miro_gpu_ray (three).
template <int MODE, bool COUNT, bool ALPHA, bool PACKED>
__global__ void __launch_bounds__(TRACE_BLOCK, TRACE_MIN_BLOCKS)
k_trace(DeviceScene s, const float4* __restrict__ rays, uint32_t n_static, const uint32_t* __restrict__ d_count, uint32_t chunk,
        miro_gpu_hit* __restrict__ hits, uint32_t* __restrict__ bits, const float4* __restrict__ sample_E, float4* __restrict__ slots,
        TraceCounters* __restrict__ ctr, uint32_t* __restrict__ work) {
    constexpr bool ANY = MODE != TRACE_CLOSEST;
    constexpr uint32_t RAY_F4 = PACKED ? 2u : 3u;       // 16-byte words per ray record
    __shared__ unsigned long long stack[SMEM_STACK * TRACE_BLOCK];
    const uint32_t n = d_count ? min(*d_count, n_static) : n_static;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t c_nodes = 0, c_tris = 0, c_insts = 0, c_rays = 0;
    uint32_t chunk_next = 0, chunk_end = 0;      // warp-uniform: the warp's claimed range of ray indices
    bool exhausted = false;                      // warp-uniform: the global counter has run past n
    Lane L; L.cur = MIRO_GPU_CHILD_EMPTY; L.ray_idx = 0; L.cur_inst = -1; L.tmin = 0.f; L.time = 0.f;
    L.hit.t = 0.f; L.hit.a = L.hit.b = 0.f; L.hit.prim = -1; L.hit.inst = -1;
    L.set_ray(0.f, 0.f, 0.f, 0.f, 0.f, 1.f);
    unsigned long long overflow[LMEM_STACK];
    TraversalStack st; st.init(stack + threadIdx.x, overflow, LMEM_STACK);

    auto write_result = [&]() {
        const uint32_t i = L.ray_idx;
        const bool hit = L.hit.prim >= 0;
        if (MODE == TRACE_ANY_BITS) { if (hit) atomicOr(bits + (i >> 5), 1u << (i & 31u)); }
        else if (MODE == TRACE_ANY_ACCUM) {
            if (!hit) { const float4 r2 = __ldcs(rays + (size_t)i * 3 + 2); atomicAdd(slots + (size_t)__float_as_uint(r2.z) * 4, __ldcs(sample_E + i)); }
        } else {
            float* o = reinterpret_cast<float*>(hits + i);
            __stcs(o + 0, hit ? L.hit.t : -1.0f); __stcs(o + 1, hit ? L.hit.a : 0.f); __stcs(o + 2, hit ? L.hit.b : 0.f);
            __stcs(reinterpret_cast<int*>(o) + 3, L.hit.prim); __stcs(reinterpret_cast<int*>(o) + 4, hit ? L.hit.inst : -1);
        }
    };
    // A slot is idle when its `cur` is EMPTY.  pending: the slot holds a ray whose result has not been written yet — it is written
    // (with the other idle lanes, converged) at the next refill, not when the ray finishes (divergent).
    bool pending = false;

    while (true) {
        // ---- refill: idle slots write their results, then take the next rays of the warp's chunk
        const uint32_t idle = __ballot_sync(0xffffffffu, L.cur == MIRO_GPU_CHILD_EMPTY);
        int n_idle = __popc(idle);
        if (n_idle >= TRACE_REFILL) {      // (all 32 idle included)
            if (pending && L.cur == MIRO_GPU_CHILD_EMPTY) { write_result(); pending = false; }
            if (exhausted) { if (idle == 0xffffffffu) break; }
            else {
                if (chunk_next == chunk_end) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(work, chunk);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    chunk_next = min(base, n); chunk_end = min(base + chunk, n);
                    if (chunk_next == chunk_end) {
                        exhausted = true;
                        // programmatic dependent launch: from here on this block only drains its last rays, so the next
                        // traversal launch on the stream may start filling the SMs as blocks of this one leave
                        if (lane == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
                    } else if (MODE == TRACE_ANY_BITS) {
                        // a 32-ray claim is exactly one result word, and only this warp ever touches it: clear it here
                        // (no memset node between two launches, which would serialise them)
                        if (lane == 0) bits[chunk_next >> 5] = 0u;
                        __syncwarp();
                    }
                }
                const uint32_t take = min((uint32_t)__popc(idle), chunk_end - chunk_next);
                const uint32_t rank = __popc(idle & lt_mask);
                if (L.cur == MIRO_GPU_CHILD_EMPTY && rank < take) {
                    L.ray_idx = chunk_next + rank;
                    const float4* rp = rays + (size_t)L.ray_idx * RAY_F4;
                    const float4 r0 = __ldcs(rp), r1 = __ldcs(rp + 1);
                    L.set_ray(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
                    L.tmin = r0.w; L.time = PACKED ? 0.f : __ldcs(reinterpret_cast<const float*>(rp + 2));
                    L.hit.t = r1.w; L.hit.a = 0.f; L.hit.b = 0.f; L.hit.prim = -1; L.hit.inst = -1;
                    L.cur = s.root; L.cur_inst = -1; st.sp = 0; pending = true;
                    c_rays += (r0.w <= r1.w) ? 1u : 0u;      // an empty interval (a light that casts no shadow) is not a Scene::trace call
                }
                chunk_next += take; n_idle -= (int)take;
                if (idle == 0xffffffffu && take == 0) continue;      // nothing claimed this round: try the next chunk (or leave)
            }
        }
        // ---- vote: a live lane waits either at an inner node or at a leaf (leaf reference / instance-exit marker)
        const bool at_node = ref_is_inner(L.cur);
        const bool at_leaf = !at_node && L.cur != MIRO_GPU_CHILD_EMPTY;
        const int n_node = __popc(__ballot_sync(0xffffffffu, at_node)), n_leaf = 32 - n_idle - n_node;
        bool finished = false;
        if (n_node * TRACE_NODE_BIAS_DEN >= n_leaf * TRACE_NODE_BIAS_NUM || (TRACE_LEAF_MIN > 0 && n_node > 0 && n_leaf < TRACE_LEAF_MIN)) {
            // ---- node round
            if (at_node) { node_step<COUNT>(s, L, st, c_nodes); finished = L.cur == MIRO_GPU_CHILD_EMPTY; }
        } else if (at_leaf) {
            // ---- leaf round
            if (L.cur == STACK_SENTINEL) {            // leaving an instance: back to the world-space ray
                const float4 w0 = __ldg(rays + (size_t)L.ray_idx * RAY_F4), w1 = __ldg(rays + (size_t)L.ray_idx * RAY_F4 + 1);
                L.set_ray(w0.x, w0.y, w0.z, w1.x, w1.y, w1.z);
                L.cur_inst = -1; L.cur = MIRO_GPU_CHILD_EMPTY;
            } else finished = intersect_leaf<ANY, COUNT, ALPHA>(s, L, st, rays, RAY_F4, c_tris, c_insts);     // true: any-hit found its occluder
            if (!finished && L.cur == MIRO_GPU_CHILD_EMPTY) { pop_next(L, st); finished = L.cur == MIRO_GPU_CHILD_EMPTY; }
        }
        if (finished) L.cur = MIRO_GPU_CHILD_EMPTY;      // (an any-hit query that found its occluder stops with work left)
    }
    // counters: warp-reduce then one atomic per warp
    unsigned long long v_rays = c_rays, v_nodes = c_nodes, v_tris = c_tris, v_insts = c_insts;
    for (int o = 16; o > 0; o >>= 1) {
        v_rays += __shfl_down_sync(0xffffffffu, v_rays, o);
        if (COUNT) {
            v_nodes += __shfl_down_sync(0xffffffffu, v_nodes, o);
            v_tris += __shfl_down_sync(0xffffffffu, v_tris, o);
            v_insts += __shfl_down_sync(0xffffffffu, v_insts, o);
        }
    }
    if (lane == 0 && v_rays) {
        atomicAdd(ANY ? &ctr->rays_any : &ctr->rays_closest, v_rays);
        if (COUNT) { atomicAdd(&ctr->nodes, v_nodes); atomicAdd(&ctr->tris, v_tris); atomicAdd(&ctr->insts, v_insts); }
    }
    // the last block to leave re-arms the work counter for the next launch on this context
    __syncthreads();
    if (threadIdx.x == 0) {
        // chained launches: no block of this launch retires before the launch chained in front of it has completed and
        // flushed (a no-op without the launch attribute), so "this launch is complete" implies "all earlier ones are".
        // A block waiting here still holds its SM slot; the host breaks the chain every WORK_RING - 1 launches
        // (launch_trace), so the ring of work-counter pairs cannot wrap onto a live launch.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        __threadfence();
        if (atomicAdd(work + 1, 1u) == gridDim.x - 1) { work[0] = 0; work[1] = 0; __threadfence(); }      // this launch's own pair (ring of pairs)
    }
}

// MIRO_GPU_KERNEL_AUTO: the kernel measured faster on the kind of scene that is uploaded (profiles/r2_ncu_summary.md section 4).  The
// flat kernel deals out the tests of static-triangle leaves: +10 % on the 87 k-triangle C2 step, +12 % on 1.74 M triangles; on the
// 40 401-instance field, where instance entries / exits and the nodes of the bottom-level trees dominate, it is 2 - 6 % behind the
// warp kernel, and with alpha cut-outs its leaves take the sequential path anyway.  Scenes of a few thousand triangles (the Cornell
// box and teapot frames: leaves of one or two triangles, everything in L1) are 2 % faster on the warp kernel.
constexpr uint32_t FLAT_MIN_TRIANGLES = 16384;
void resolve_trace_kernel(miro_gpu_ctx* ctx) {
    if (ctx->trace_kernel_request != MIRO_GPU_KERNEL_AUTO) ctx->trace_kernel = ctx->trace_kernel_request;
    else ctx->trace_kernel = (ctx->n_insts > 0 || ctx->has_alpha || ctx->n_tris + ctx->n_mbtris < FLAT_MIN_TRIANGLES) ? MIRO_GPU_KERNEL_WARP : MIRO_GPU_KERNEL_FLAT;
}

// Device-built trees: hit records leave the library in the CALLER's triangle numbering.
__global__ void k_translate_hits(miro_gpu_hit* __restrict__ hits, uint32_t n, const uint32_t* __restrict__ prim_map) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && hits[i].prim >= 0) hits[i].prim = (int32_t)prim_map[hits[i].prim];
}

template <int MODE>
static int trace_grid(miro_gpu_ctx* ctx, size_t n) {
    // persistent grid: every SM holds as many blocks as fit (asked of the occupancy calculator once per kernel)
    static int per_sm[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // (group worker threads may fill an entry concurrently: with the same value)
    const bool flat = ctx->trace_kernel == MIRO_GPU_KERNEL_FLAT;
    int& v = per_sm[(ctx->counting ? 1 : 0) + (ctx->has_alpha ? 2 : 0) + (flat ? 4 : 0)];
    if (v == 0 && flat) {
        if (ctx->has_alpha) {
            if (ctx->counting) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace_flat<MODE, true, true, false, false>, TRACE_BLOCK, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace_flat<MODE, false, true, false, false>, TRACE_BLOCK, 0);
        } else {
            if (ctx->counting) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace_flat<MODE, true, false, false, false>, TRACE_BLOCK, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace_flat<MODE, false, false, false, false>, TRACE_BLOCK, 0);
        }
        if (v <= 0) v = 1;
    }
    if (v == 0) {
        if (ctx->has_alpha) {
            if (ctx->counting) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace<MODE, true, true, false>, TRACE_BLOCK, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace<MODE, false, true, false>, TRACE_BLOCK, 0);
        } else {
            if (ctx->counting) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace<MODE, true, false, false>, TRACE_BLOCK, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_trace<MODE, false, false, false>, TRACE_BLOCK, 0);
        }
        if (v <= 0) v = 1;
    }
    const size_t blocks_needed = (n + TRACE_BLOCK - 1) / TRACE_BLOCK;
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->sm_count * v, blocks_needed));
}

// Pool kernel (trace_pool.cuh): grid = SMs x resident blocks, dynamic shared memory = the warps' pools; the per-slot stack
// overflow scratch (global memory; one per work lane, because launches of different lanes run concurrently) is sized from the
// depth of the uploaded trees and allocated on first use.
template <int MODE, bool PACKED>
static cudaError_t launch_trace_pool(miro_gpu_ctx* ctx, const cudaLaunchConfig_t& base_cfg, const float4* r, size_t n, const uint32_t* d_count, uint32_t chunk,
                                     miro_gpu_hit* d_hits, uint32_t* d_bits, const float4* d_E, float4* d_slots, uint32_t* work, uint64_t scratch_copy) {
    static int per_sm[4] = {0, 0, 0, 0};
    int& v = per_sm[(ctx->counting ? 1 : 0) + (ctx->has_alpha ? 2 : 0)];
#define MIRO_POOL_KERNEL(COUNT, ALPHA) k_trace_pool<MODE, COUNT, ALPHA, PACKED>
#define MIRO_POOL_DISPATCH(WHAT) \
    if (ctx->has_alpha) { if (ctx->counting) { WHAT(true, true); } else { WHAT(false, true); } } \
    else { if (ctx->counting) { WHAT(true, false); } else { WHAT(false, false); } }
    if (v == 0) {
#define MIRO_POOL_SETUP(COUNT, ALPHA) \
        cudaFuncSetAttribute(MIRO_POOL_KERNEL(COUNT, ALPHA), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POOL_SMEM_BYTES); \
        cudaFuncSetAttribute(MIRO_POOL_KERNEL(COUNT, ALPHA), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, MIRO_POOL_KERNEL(COUNT, ALPHA), POOL_BLOCK, POOL_SMEM_BYTES)
        MIRO_POOL_DISPATCH(MIRO_POOL_SETUP)
#undef MIRO_POOL_SETUP
        if (v <= 0) v = 1;
    }
    const int full_grid = ctx->sm_count * v;
    const size_t blocks_needed = (n + POOL_WARPS * POOL_SLOTS - 1) / (POOL_WARPS * POOL_SLOTS);
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)full_grid, blocks_needed));
    const int cap = std::max(0, ctx->stack_need - POOL_STACK);
    const size_t need = (size_t)full_grid * POOL_WARPS * POOL_SLOTS * (size_t)cap;
    PoolScratch& sc = ctx->pool_ovf[ctx->work_lane];
    if (need * SCRATCH_COPIES > sc.entries) {
        if (sc.ptr) { cudaStreamSynchronize(ctx->stream); cudaFree(sc.ptr); sc.ptr = nullptr; sc.entries = 0; }
        cudaError_t e = cudaMalloc((void**)&sc.ptr, need * SCRATCH_COPIES * sizeof(unsigned long long));
        if (e != cudaSuccess) return e;
        sc.entries = need * SCRATCH_COPIES;
    }
    unsigned long long* const scratch = sc.ptr ? sc.ptr + need * (size_t)(scratch_copy % SCRATCH_COPIES) : nullptr;
    cudaLaunchConfig_t cfg = base_cfg;
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(POOL_BLOCK); cfg.dynamicSmemBytes = POOL_SMEM_BYTES;
    cudaError_t e = cudaSuccess;
#define MIRO_POOL_LAUNCH(COUNT, ALPHA) e = cudaLaunchKernelEx(&cfg, MIRO_POOL_KERNEL(COUNT, ALPHA), ctx->scene, r, (uint32_t)n, d_count, chunk, d_hits, d_bits, d_E, d_slots, ctx->d_counters, work, scratch, cap)
    MIRO_POOL_DISPATCH(MIRO_POOL_LAUNCH)
#undef MIRO_POOL_LAUNCH
#undef MIRO_POOL_DISPATCH
#undef MIRO_POOL_KERNEL
    return e;
}

template <int MODE, bool PACKED = false>
static void launch_trace(miro_gpu_ctx* ctx, const void* d_rays, size_t n, const uint32_t* d_count, miro_gpu_hit* d_hits, uint32_t* d_bits,
                         const float4* d_E, float4* d_slots) {
    if (n == 0) return;
    const int grid = trace_grid<MODE>(ctx, n);
    // rays are claimed one warp-load (32) at a time: measured on the coherent 1080p batch, 64-ray chunks left a tail worth 15 %
    // of the launch (a warp stuck with two heavy chunks while the rest of the GPU had drained); 16 is no better than 32
    const uint32_t chunk = 32u;      // exactly one any-hit result word per claim: the claiming warp clears it in-kernel
    static_assert(MODE != TRACE_ANY_BITS || true, "");
    const float4* r = reinterpret_cast<const float4*>(d_rays);
    // Every launch has its own pair of work counters out of a ring, so consecutive traversal launches can overlap.  When the
    // caller has switched trace chaining on (miro_gpu_set_trace_chaining: it vouches that the inputs do not depend on work
    // enqueued since the previous trace call), the launch carries the programmatic-dependent-launch attribute: the tail of
    // launch k, where warps drain their last rays at falling occupancy, overlaps the start of launch k+1.
    const uint64_t slot = ctx->work_slot[ctx->work_lane]++;
    uint32_t* work = ctx->d_work + 2 * ((size_t)ctx->work_lane * WORK_RING + slot % WORK_RING);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TRACE_BLOCK); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    // The chain is broken at every (WORK_RING - 1)-th launch of a lane (that launch waits for the complete end of everything
    // before it, as an unchained one does): short launches pass their launch_dependents point almost at once and park at
    // griddepcontrol.wait while still resident, so without the break any number of them could be live behind one long launch
    // and launch k + WORK_RING would claim rays from the not yet re-armed counter pair of launch k.  With the break at most
    // WORK_RING - 1 consecutive launches of a lane are ever live together.
    const bool scratch_kernel = ctx->trace_kernel == MIRO_GPU_KERNEL_POOL;      // see SCRATCH_COPIES
    const bool chained = ctx->chain_traces && ctx->in_api_trace && (slot % (WORK_RING - 1)) != 0 && !(scratch_kernel && slot % SCRATCH_COPIES == 0);
    cfg.attrs = attr; cfg.numAttrs = chained ? 1 : 0;
    ctx->launches++;
    if (ctx->trace_kernel == MIRO_GPU_KERNEL_POOL) {
        static_assert((int)TRACE_CLOSEST == (int)POOL_TRACE_CLOSEST && (int)TRACE_ANY_BITS == (int)POOL_TRACE_ANY_BITS && (int)TRACE_ANY_ACCUM == (int)POOL_TRACE_ANY_ACCUM, "mode numbering");
        const cudaError_t e = launch_trace_pool<MODE, PACKED>(ctx, cfg, r, n, d_count, chunk, d_hits, d_bits, d_E, d_slots, work, slot);
        if (e != cudaSuccess) ctx->error = std::string("pool traversal launch: ") + cudaGetErrorString(e);      // surfaces through the caller's cudaGetLastError check
        return;
    }
#define MIRO_LAUNCH(COUNT, ALPHA) do { if (ctx->trace_kernel == MIRO_GPU_KERNEL_FLAT) \
        { if (ctx->n_mbtris) cudaLaunchKernelEx(&cfg, k_trace_flat<MODE, COUNT, ALPHA, PACKED, true>, ctx->scene, r, (uint32_t)n, d_count, chunk, d_hits, d_bits, d_E, d_slots, ctx->d_counters, work); \
          else cudaLaunchKernelEx(&cfg, k_trace_flat<MODE, COUNT, ALPHA, PACKED, false>, ctx->scene, r, (uint32_t)n, d_count, chunk, d_hits, d_bits, d_E, d_slots, ctx->d_counters, work); } \
    else cudaLaunchKernelEx(&cfg, k_trace<MODE, COUNT, ALPHA, PACKED>, ctx->scene, r, (uint32_t)n, d_count, chunk, d_hits, d_bits, d_E, d_slots, ctx->d_counters, work); } while (0)
    if (ctx->has_alpha) { if (ctx->counting) MIRO_LAUNCH(true, true); else MIRO_LAUNCH(false, true); }
    else { if (ctx->counting) MIRO_LAUNCH(true, false); else MIRO_LAUNCH(false, false); }
#undef MIRO_LAUNCH
}
void launch_trace_closest(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, const uint32_t* d_count, miro_gpu_hit* d_hits) {
    launch_trace<TRACE_CLOSEST>(ctx, d_rays, n, d_count, d_hits, nullptr, nullptr, nullptr);
}
void launch_trace_any(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, const uint32_t* d_count, uint32_t* d_bits) {
    launch_trace<TRACE_ANY_BITS>(ctx, d_rays, n, d_count, nullptr, d_bits, nullptr, nullptr);
}
void launch_trace_closest_packed(miro_gpu_ctx* ctx, const miro_gpu_ray32* d_rays, size_t n, miro_gpu_hit* d_hits) {
    launch_trace<TRACE_CLOSEST, true>(ctx, d_rays, n, nullptr, d_hits, nullptr, nullptr, nullptr);
}
void launch_trace_any_packed(miro_gpu_ctx* ctx, const miro_gpu_ray32* d_rays, size_t n, uint32_t* d_bits) {
    launch_trace<TRACE_ANY_BITS, true>(ctx, d_rays, n, nullptr, nullptr, d_bits, nullptr, nullptr);
}
void launch_trace_shadow(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, const uint32_t* d_count, const float4* d_E, float4* d_slots) {
    launch_trace<TRACE_ANY_ACCUM>(ctx, d_rays, n, d_count, nullptr, nullptr, d_E, d_slots);
}

}  // namespace miro

// ---------------------------------------------------------------------------------------------
extern "C" {

int miro_gpu_abi_version(void) { return MIRO_GPU_ABI_VERSION; }

static_assert(sizeof(miro_gpu_ray32) == 32, "include/miro_gpu.h layout changed");
static_assert(sizeof(miro_gpu_ray) == 48 && sizeof(miro_gpu_hit) == 20 && sizeof(miro_gpu_node) == 128 && sizeof(miro_gpu_tri) == 48 &&
              sizeof(miro_gpu_mbtri) == 96 && sizeof(miro_gpu_instance) == 64 && sizeof(miro_gpu_prim) == 48 &&
              sizeof(miro_gpu_material) == 128 && sizeof(miro_gpu_light) == 64, "include/miro_gpu.h layout changed");
size_t miro_gpu_sizeof(int k) {
    static const size_t sizes[] = {sizeof(miro_gpu_ray), sizeof(miro_gpu_hit), sizeof(miro_gpu_node), sizeof(miro_gpu_tri), sizeof(miro_gpu_mbtri),
                                   sizeof(miro_gpu_instance), sizeof(miro_gpu_prim), sizeof(miro_gpu_material), sizeof(miro_gpu_light),
                                   sizeof(miro_gpu_texture), sizeof(miro_gpu_scene_desc), sizeof(miro_gpu_camera), sizeof(miro_gpu_render_params),
                                   sizeof(miro_gpu_counters)};
    return (k >= 0 && k < (int)(sizeof(sizes) / sizeof(sizes[0]))) ? sizes[k] : 0;
}

const char* miro_gpu_last_error(const miro_gpu_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int miro_gpu_create(miro_gpu_ctx** out, int device_id) {
    if (!out) return set_error(nullptr, MIRO_GPU_EINVAL, "miro_gpu_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(nullptr, MIRO_GPU_ENODEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU fallback)");
    if (device_id < 0 || device_id >= n) return set_error(nullptr, MIRO_GPU_EINVAL, "miro_gpu_create: bad device id");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return set_error(nullptr, MIRO_GPU_ENODEVICE, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                         "; this build contains sm_100a code only");
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    miro_gpu_ctx* ctx = new miro_gpu_ctx();
    ctx->device = device_id;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) { delete ctx; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
    ctx->stream = ctx->own_stream;
    if ((e = cudaMalloc((void**)&ctx->d_counters, sizeof(TraceCounters))) != cudaSuccess) { delete ctx; return cuda_fail(nullptr, e, "cudaMalloc(counters)"); }
    cudaMemset(ctx->d_counters, 0, sizeof(TraceCounters));
    if ((e = cudaMalloc((void**)&ctx->d_work, 2 * WORK_RING * WORK_LANES * sizeof(uint32_t))) != cudaSuccess) { delete ctx; return cuda_fail(nullptr, e, "cudaMalloc(work counter)"); }
    cudaMemset(ctx->d_work, 0, 2 * WORK_RING * WORK_LANES * sizeof(uint32_t));
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* k = getenv("MIRO_GPU_TRACE_KERNEL"))
        ctx->trace_kernel_request = strcmp(k, "pool") == 0 ? MIRO_GPU_KERNEL_POOL : strcmp(k, "flat") == 0 ? MIRO_GPU_KERNEL_FLAT : strcmp(k, "warp") == 0 ? MIRO_GPU_KERNEL_WARP : MIRO_GPU_KERNEL_AUTO;
    resolve_trace_kernel(ctx);
    // the traversal kernels keep their stacks in shared memory and want the rest of the 256 KB as L1
    *out = ctx;
    return MIRO_GPU_OK;
}

static void free_scene(miro_gpu_ctx* ctx) {
    for (void* p : ctx->scene_allocs) cudaFree(p);
    ctx->scene_allocs.clear();
    ctx->has_scene = false;
}

void miro_gpu_destroy(miro_gpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drain_timing(ctx);
    for (EventPair& p : ctx->event_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    render_state_free(ctx);
    free_scene(ctx);
    ctx->d_rays.release(); ctx->d_hits.release(); ctx->d_bits.release();
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_work) cudaFree(ctx->d_work);
    for (PoolScratch& sc : ctx->pool_ovf) if (sc.ptr) cudaFree(sc.ptr);
    for (cudaEvent_t e : ctx->pipe_events) cudaEventDestroy(e);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    for (cudaStream_t a : ctx->trace_aux) if (a) cudaStreamDestroy(a);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int miro_gpu_set_stream(miro_gpu_ctx* ctx, void* cuda_stream) {
    if (!ctx) return MIRO_GPU_EINVAL;
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return MIRO_GPU_OK;
}

}  // extern "C"

template <class T>
static int upload_array(miro_gpu_ctx* ctx, const T* host, size_t n, const T** dev, size_t min_elems = 1) {
    size_t bytes = std::max(n, min_elems) * sizeof(T);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(scene)");
    ctx->scene_allocs.push_back(p);
    if (n == 0 || !host) cudaMemsetAsync(p, 0, bytes, ctx->stream);
    else if ((e = cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMemcpy(scene)");
    *dev = (const T*)p;
    return MIRO_GPU_OK;
}

// Depth of the (sub)tree behind a child-style reference; -1 on a malformed tree.  memo[node] caches the depth of a node's
// sub-tree (-2: not visited, -3: on the current path, i.e. a cycle), so shared sub-trees — every instance of one bottom-level
// tree — and DAG-shaped input cost one visit per node.  in_blas: the walk is below an instance; the reference has ONE level
// of instancing (a ProxyObject's BVH holds plain Objects, src/ProxyObject.cpp:131-166) and so has the traversal kernel.
static int tree_depth(const miro_gpu_scene_desc* d, int32_t ref, int level, bool in_blas, std::vector<int>& memo, std::string& err, int& code) {
    if (ref == MIRO_GPU_CHILD_EMPTY) return 0;
    if (ref < 0) {
        uint32_t u = (uint32_t)ref, kind = (u >> 29) & 3u, count = ((u >> MIRO_GPU_LEAF_INDEX_BITS) & 7u) + 1u, first = u & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u);
        uint32_t limit = kind == MIRO_GPU_KIND_TRI ? d->n_tris : kind == MIRO_GPU_KIND_MBTRI ? d->n_mbtris : kind == MIRO_GPU_KIND_INST ? d->n_instances : 0;
        if (kind > 2 || first + count > limit) { err = "leaf reference out of range"; return -1; }
        if (kind == MIRO_GPU_KIND_INST && in_blas) { err = "instance inside an instanced tree (one level of instancing, as the reference)"; code = MIRO_GPU_EUNSUPPORTED; return -1; }
        return 0;
    }
    if ((uint32_t)ref >= d->n_nodes) { err = "child node index out of range"; return -1; }
    if (level > 64) { err = "BVH deeper than 64 levels"; return -1; }
    int& m = memo[(size_t)ref];
    if (m == -3) { err = "cycle in the node array"; return -1; }
    if (m >= 0) return m;
    m = -3;
    int best = 0;
    for (int i = 0; i < 4; ++i) {
        int c = tree_depth(d, d->nodes[ref].child[i], level + 1, in_blas, memo, err, code);
        if (c < 0) return -1;
        best = std::max(best, c);
    }
    return memo[(size_t)ref] = best + 1;
}

extern "C" {

int miro_gpu_upload_scene(miro_gpu_ctx* ctx, const miro_gpu_scene_desc* d) {
    if (!ctx || !d) return MIRO_GPU_EINVAL;
    if (d->abi_version != MIRO_GPU_ABI_VERSION) return set_error(ctx, MIRO_GPU_EINVAL, "scene desc abi_version mismatch");
    if ((d->n_nodes && !d->nodes) || (d->n_tris && !d->tris) || (d->n_mbtris && !d->mbtris) || (d->n_instances && !d->instances))
        return set_error(ctx, MIRO_GPU_EINVAL, "scene desc: NULL array with non-zero count");
    if (d->n_tris >= (1u << MIRO_GPU_LEAF_INDEX_BITS) || d->n_mbtris >= (1u << MIRO_GPU_LEAF_INDEX_BITS) || d->n_instances >= (1u << MIRO_GPU_LEAF_INDEX_BITS))
        return set_error(ctx, MIRO_GPU_EUNSUPPORTED, "scene desc: more than 2^26 primitives of one kind");
    if (d->n_lights > MIRO_GPU_MAX_LIGHTS) return set_error(ctx, MIRO_GPU_EUNSUPPORTED, "more than MIRO_GPU_MAX_LIGHTS lights");
    const bool device_build = d->root == MIRO_GPU_ROOT_BUILD_ON_DEVICE;
    if (device_build && (d->n_mbtris || d->n_instances || d->n_nodes))
        return set_error(ctx, MIRO_GPU_EUNSUPPORTED, "device BVH build handles static triangles only (no motion-blur triangles, instances or host nodes)");
    if ((d->n_materials && !d->materials) || (d->n_lights && !d->lights) || (d->n_textures && !d->textures) || (d->n_normals && !d->normals) || (d->n_uvs && !d->uvs))
        return set_error(ctx, MIRO_GPU_EINVAL, "scene desc: NULL material / light / texture / normal / uv array with non-zero count");
    // validate the trees and bound the traversal stack: <= 3 pushes per level, + 2 for an instance hop.  The top-level walk
    // and the bottom-level walks keep separate memos: a node reached both ways would be an instance's tree containing an
    // instance, or the top level entering a bottom-level tree directly — the second is legal, the first is reported.
    std::string err;
    int code = MIRO_GPU_EINVAL;
    std::vector<int> memo_top(d->n_nodes, -2), memo_blas(d->n_nodes, -2);
    int top = device_build ? 0 : tree_depth(d, d->root, 0, false, memo_top, err, code);
    if (top < 0) return set_error(ctx, code, "scene desc: " + err);
    int blas = 0;
    for (uint32_t i = 0; i < d->n_instances; ++i) {
        int b = tree_depth(d, d->instances[i].blas_root, 0, true, memo_blas, err, code);
        if (b < 0) return set_error(ctx, code, "scene desc: instance " + std::to_string(i) + ": " + err);
        blas = std::max(blas, b);
    }
    if (3 * top + 2 + 3 * blas > SMEM_STACK + LMEM_STACK)
        return set_error(ctx, MIRO_GPU_EUNSUPPORTED, "BVH too deep for the traversal stack (top " + std::to_string(top) + ", instanced " + std::to_string(blas) + " levels)");
    int stack_need = 3 * top + 2 + 3 * blas;
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const miro_gpu_material& m = d->materials[i];
        if (m.kind > MIRO_GPU_MAT_BLINN) return set_error(ctx, MIRO_GPU_EINVAL, "unknown material kind");
        if (m.alpha_map >= (int32_t)d->n_textures) return set_error(ctx, MIRO_GPU_EINVAL, "material alpha_map out of range");
        if (m.color_map >= (int32_t)d->n_textures) return set_error(ctx, MIRO_GPU_EINVAL, "material color_map out of range");
        if (m.normal_map >= (int32_t)d->n_textures || m.specular_map >= (int32_t)d->n_textures || m.reflect_map >= (int32_t)d->n_textures ||
            m.refract_map >= (int32_t)d->n_textures) return set_error(ctx, MIRO_GPU_EINVAL, "material normal/specular/reflect/refract map out of range");
    }
    for (uint32_t i = 0; i < d->n_tris + d->n_mbtris && d->prims; ++i) {
        const miro_gpu_prim& pr = d->prims[i];
        if (pr.material >= d->n_materials) return set_error(ctx, MIRO_GPU_EINVAL, "prim material out of range");
        for (int k = 0; k < 3; ++k) {
            if (pr.n[k] >= d->n_normals) return set_error(ctx, MIRO_GPU_EINVAL, "prim " + std::to_string(i) + ": normal index out of range");
            if (pr.uv[k] != 0xffffffffu && pr.uv[k] >= d->n_uvs) return set_error(ctx, MIRO_GPU_EINVAL, "prim " + std::to_string(i) + ": uv index out of range");
        }
    }

    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_scene(ctx);
    int rc;
    const DeviceNode* dn; const miro_gpu_tri* dt; const miro_gpu_mbtri* dm; const miro_gpu_instance* di;
    const uint32_t* d_perm = nullptr;
    int32_t root = d->root;
    if (device_build) {
        // triangles in the caller's order go up once; the build sorts them into leaf order on the device
        const miro_gpu_tri* d_in;
        if ((rc = upload_array(ctx, d->tris, d->n_tris, &d_in))) return rc;
        uint32_t n_nodes = 0; const float4* sorted = nullptr;
        if ((rc = build_lbvh_on_device(ctx, reinterpret_cast<const float4*>(d_in), d->n_tris, &dn, &n_nodes, &root, &sorted, &d_perm))) return rc;
        dt = reinterpret_cast<const miro_gpu_tri*>(sorted);
        if (3 * ctx->build_levels + 2 > SMEM_STACK + LMEM_STACK) return set_error(ctx, MIRO_GPU_EUNSUPPORTED, "device-built BVH too deep for the traversal stack");
        stack_need = 3 * ctx->build_levels + 2;
        ctx->n_nodes = n_nodes;
    } else {
        std::vector<DeviceNode> cnodes(d->n_nodes);          // 128-byte ABI nodes -> 64-byte device nodes (traverse.cuh)
        for (uint32_t i = 0; i < d->n_nodes; ++i) cnodes[i] = compress_node(d->nodes[i]);
        if ((rc = upload_array(ctx, cnodes.data(), cnodes.size(), &dn))) return rc;
        // triangles: as they are, or (TRI_F4 == 4) as 64-byte aligned device records (traverse.cuh)
#if MIRO_TRI_F4 == 4
        struct Tri64 { miro_gpu_tri t; uint32_t pad[4]; };
        static_assert(sizeof(Tri64) == TRI_F4 * sizeof(float4), "device triangle record");
        std::vector<Tri64> ctris(d->n_tris);
        for (uint32_t i = 0; i < d->n_tris; ++i) { ctris[i].t = d->tris[i]; ctris[i].pad[0] = ctris[i].pad[1] = ctris[i].pad[2] = ctris[i].pad[3] = 0; }
        const Tri64* dt64;
        if ((rc = upload_array(ctx, ctris.data(), ctris.size(), &dt64))) return rc;
#else
        const miro_gpu_tri* dt64;
        if ((rc = upload_array(ctx, d->tris, d->n_tris, &dt64))) return rc;
#endif
        MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the staging vectors go out of scope
        dt = reinterpret_cast<const miro_gpu_tri*>(dt64);
    }
    if ((rc = upload_array(ctx, d->mbtris, d->n_mbtris, &dm))) return rc;
    if ((rc = upload_array(ctx, d->instances, d->n_instances, &di))) return rc;
    ctx->scene.nodes = reinterpret_cast<const float4*>(dn);
    ctx->scene.tris = reinterpret_cast<const float4*>(dt);
    ctx->scene.mbtris = reinterpret_cast<const float4*>(dm);
    ctx->scene.insts = reinterpret_cast<const float4*>(di);
    ctx->scene.root = root;
    ctx->scene.n_tris = d->n_tris;
    ctx->scene.prim_map = d_perm;
    if (!device_build) ctx->n_nodes = d->n_nodes;
    ctx->stack_need = stack_need;
    ctx->n_tris = d->n_tris; ctx->n_mbtris = d->n_mbtris; ctx->n_insts = d->n_instances;

    DeviceShading& sh = ctx->shading;
    memset(&sh, 0, sizeof(sh));
    const uint32_t n_prims = d->prims ? d->n_tris + d->n_mbtris : 0;
    if ((rc = upload_array(ctx, d->prims, n_prims, &sh.prims))) return rc;
    if (device_build && n_prims) {       // shading records follow the triangles into leaf order
        const miro_gpu_prim* sorted_prims;
        if ((rc = reorder_prims_on_device(ctx, sh.prims, d_perm, n_prims, &sorted_prims))) return rc;
        sh.prims = sorted_prims;
    }
    if ((rc = upload_array(ctx, d->normals, (size_t)d->n_normals * 3, &sh.normals))) return rc;
    if ((rc = upload_array(ctx, d->tangents, d->tangents ? (size_t)d->n_normals * 3 : 0, &sh.tangents))) return rc;
    if ((rc = upload_array(ctx, d->bitangents, d->bitangents ? (size_t)d->n_normals * 3 : 0, &sh.bitangents))) return rc;
    if ((rc = upload_array(ctx, d->uvs, (size_t)d->n_uvs * 2, &sh.uvs))) return rc;
    if ((rc = upload_array(ctx, d->inst_normal_xform, d->inst_normal_xform ? (size_t)d->n_instances * 9 : 0, &sh.inst_nxf))) return rc;
    if ((rc = upload_array(ctx, d->materials, d->n_materials, &sh.materials))) return rc;
    if ((rc = upload_array(ctx, d->lights, d->n_lights, &sh.lights))) return rc;
    sh.n_lights = d->n_lights; sh.n_materials = d->n_materials; sh.n_textures = d->n_textures; sh.n_prims = n_prims;
    sh.env_map = d->env_map; sh.env_exposure = d->env_exposure;
    sh.bg[0] = d->bg_color[0]; sh.bg[1] = d->bg_color[1]; sh.bg[2] = d->bg_color[2];
    ctx->host_lights.assign(d->lights, d->lights + d->n_lights);
    ctx->host_materials.assign(d->materials, d->materials + d->n_materials);
    if (d->env_map >= (int32_t)d->n_textures) return set_error(ctx, MIRO_GPU_EINVAL, "env_map out of range");
    // textures
    std::vector<DeviceTexture> tex(d->n_textures);
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const miro_gpu_texture& t = d->textures[i];
        if (!t.texels || t.width <= 0 || t.height <= 0 || (t.channels != 1 && t.channels != 3 && t.channels != 4))
            return set_error(ctx, MIRO_GPU_EINVAL, "bad texture " + std::to_string(i));
        const float* p;
        if ((rc = upload_array(ctx, t.texels, (size_t)t.width * t.height * t.channels, &p))) return rc;
        tex[i].texels = p; tex[i].width = t.width; tex[i].height = t.height; tex[i].channels = t.channels; tex[i].pad = 0;
    }
    if ((rc = upload_array(ctx, tex.data(), tex.size(), &sh.textures))) return rc;
    // alpha cut-outs are part of Scene::trace (intersect4, src/BVH.cpp:1401-1435): the traversal kernels get what they need
    ctx->has_alpha = false;
    for (uint32_t i = 0; i < d->n_materials; ++i) if (d->materials[i].alpha_map >= 0 && d->textures[d->materials[i].alpha_map].channels == 4) ctx->has_alpha = true;
    static_assert(sizeof(AlphaTexture) == sizeof(DeviceTexture), "AlphaTexture mirrors DeviceTexture");
    ctx->scene.alpha = AlphaData{sh.prims, sh.uvs, sh.materials, reinterpret_cast<const AlphaTexture*>(sh.textures)};
    if (ctx->has_alpha && !d->prims) return set_error(ctx, MIRO_GPU_EINVAL, "alpha-mapped materials need the prims table");
    resolve_trace_kernel(ctx);
    // dome lights: importance tables (DomeLight::setTexture, src/DomeLight.cpp:8-78)
    std::vector<DeviceDome> domes(std::max<uint32_t>(d->n_lights, 1));
    memset(domes.data(), 0, domes.size() * sizeof(DeviceDome));
    for (uint32_t i = 0; i < d->n_lights; ++i) {
        const miro_gpu_light& l = d->lights[i];
        if (l.kind > MIRO_GPU_LIGHT_DOME) return set_error(ctx, MIRO_GPU_EINVAL, "unknown light kind");
        if (l.kind != MIRO_GPU_LIGHT_DOME) continue;
        if (l.texture < 0 || l.texture >= (int32_t)d->n_textures) return set_error(ctx, MIRO_GPU_EINVAL, "dome light without a texture");
        if ((rc = build_dome_tables(ctx, d->textures[l.texture], &domes[i]))) return rc;
    }
    if ((rc = upload_array(ctx, domes.data(), domes.size(), &sh.domes))) return rc;
    MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_scene = true;
    return MIRO_GPU_OK;
}

int miro_gpu_trace_closest_device(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, miro_gpu_hit* d_hits) {
    if (!ctx) return MIRO_GPU_EINVAL;
    if (!ctx->has_scene) return set_error(ctx, MIRO_GPU_ENOSCENE, "trace before upload_scene");
    if (n > MAX_RAYS_PER_CALL) return set_error(ctx, MIRO_GPU_EINVAL, "more than 2^32 - 2^24 rays in one call");
    if (n == 0) return MIRO_GPU_OK;
    if (!d_rays || !d_hits) return set_error(ctx, MIRO_GPU_EINVAL, "NULL ray/hit buffer");
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    EventPair p = begin_timing(ctx, true);
    ctx->in_api_trace = true;
    launch_trace_closest(ctx, d_rays, n, nullptr, d_hits);
    ctx->in_api_trace = false;
    if (ctx->scene.prim_map) { k_translate_hits<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_hits, (uint32_t)n, ctx->scene.prim_map); ctx->launches++; }
    end_timing(ctx, p);
    MIRO_CUDA(ctx, cudaGetLastError());
    return MIRO_GPU_OK;
}

int miro_gpu_trace_any_device(miro_gpu_ctx* ctx, const miro_gpu_ray* d_rays, size_t n, uint32_t* d_bits) {
    if (!ctx) return MIRO_GPU_EINVAL;
    if (!ctx->has_scene) return set_error(ctx, MIRO_GPU_ENOSCENE, "trace before upload_scene");
    if (n > MAX_RAYS_PER_CALL) return set_error(ctx, MIRO_GPU_EINVAL, "more than 2^32 - 2^24 rays in one call");
    if (n == 0) return MIRO_GPU_OK;
    if (!d_rays || !d_bits) return set_error(ctx, MIRO_GPU_EINVAL, "NULL ray/bit buffer");
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    EventPair p = begin_timing(ctx, true);
    ctx->in_api_trace = true;
    launch_trace_any(ctx, d_rays, n, nullptr, d_bits);
    ctx->in_api_trace = false;
    end_timing(ctx, p);
    MIRO_CUDA(ctx, cudaGetLastError());
    return MIRO_GPU_OK;
}

// Host-pointer entry points: the batch is cut into chunks and pipelined over three streams — chunk k+1 is on its way
// up (H2D) and chunk k-1 on its way down (D2H) while chunk k is traversed — so the call costs max(copy, compute), not
// their sum.  PCIe is full duplex, so the two copy directions overlap as well.
static int trace_host(miro_gpu_ctx* ctx, const void* rays_v, size_t n, miro_gpu_hit* hits, uint32_t* bits, bool packed = false) {
    const char* rays = static_cast<const char*>(rays_v);
    const size_t ray_bytes = packed ? sizeof(miro_gpu_ray32) : sizeof(miro_gpu_ray);
    if (!ctx) return MIRO_GPU_EINVAL;
    if (!ctx->has_scene) return set_error(ctx, MIRO_GPU_ENOSCENE, "trace before upload_scene");
    if (n > MAX_RAYS_PER_CALL) return set_error(ctx, MIRO_GPU_EINVAL, "more than 2^32 - 2^24 rays in one call");
    if (n == 0) return MIRO_GPU_OK;
    if (!rays || (!hits && !bits)) return set_error(ctx, MIRO_GPU_EINVAL, "NULL ray/result buffer");
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    // Chunked pipeline: H2D of chunk k+1, traversal of chunk k and D2H of chunk k-1 overlap.  Measured on B200 (tools/e2e_sweep.sh,
    // tools/e2e_timeline.py): the call is bound by the H2D stream (48 B per ray over PCIe, 43-54 GB/s with the D2H running
    // against it); 256 K-ray chunks are the optimum (smaller: per-chunk kernel tails and copy latencies; larger: fill / drain).
    // MIRO_GPU_CHUNK (log2 rays), MIRO_GPU_KSTREAMS and MIRO_GPU_TIMELINE are tuning / diagnostic switches.
    static const int chunk_log2 = getenv("MIRO_GPU_CHUNK") ? std::min(26, std::max(5, atoi(getenv("MIRO_GPU_CHUNK")))) : 18;
    const size_t chunk = (size_t)1 << chunk_log2;              // rays per chunk (a multiple of 32: whole result words)
    std::vector<size_t> bounds;                                // chunk k = [bounds[k], bounds[k+1])
    for (size_t off = 0; off < n; off += chunk) bounds.push_back(off);
    bounds.push_back(n);
    const size_t n_chunks = bounds.size() - 1;
    MIRO_CUDA(ctx, ctx->d_rays.reserve(n));
    if (hits) MIRO_CUDA(ctx, ctx->d_hits.reserve(n)); else MIRO_CUDA(ctx, ctx->d_bits.reserve((n + 31) / 32));
    if (!ctx->copy_in) {
        MIRO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        MIRO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        for (cudaStream_t& a : ctx->trace_aux) MIRO_CUDA(ctx, cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        MIRO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    }
    // The chunk kernels rotate over k_streams streams: a traversal launch ends with a tail of ~150 us in which its last warps
    // walk their longest rays alone (a chain of dependent node fetches); on one stream that tail is paid once per chunk and made
    // the kernel stage, not the PCIe copy, the slowest stage of the pipeline.  On rotating streams the next chunk's kernel fills
    // the SMs the tail leaves idle.  Work the caller enqueued on the context's stream before this call is waited for by all of them.
    static const int k_streams = getenv("MIRO_GPU_KSTREAMS") ? std::min(4, std::max(1, atoi(getenv("MIRO_GPU_KSTREAMS")))) : 2;
    cudaStream_t const user_stream = ctx->stream;
    struct Restore { miro_gpu_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; c->work_lane = 0; } } restore{ctx, user_stream};      // also on error returns
    if (k_streams > 1) {
        MIRO_CUDA(ctx, cudaEventRecord(ctx->fork_event, user_stream));
        for (int a = 0; a + 1 < k_streams; ++a) MIRO_CUDA(ctx, cudaStreamWaitEvent(ctx->trace_aux[a], ctx->fork_event, 0));
    }
    while (ctx->pipe_events.size() < 2 * n_chunks) {
        cudaEvent_t e; MIRO_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    EventPair tot = begin_timing(ctx, false);
    static const bool timeline = getenv("MIRO_GPU_TIMELINE") != nullptr;      // diagnostic: per-chunk device timeline on stderr
    std::vector<cudaEvent_t> tl;
    auto mark = [&](cudaStream_t st) { if (timeline) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tl.push_back(e); } };
    mark(ctx->copy_in);
    for (size_t k = 0; k < n_chunks; ++k) {
        const size_t off = bounds[k], m = bounds[k + 1] - off;
        cudaEvent_t up = ctx->pipe_events[2 * k], done = ctx->pipe_events[2 * k + 1];
        char* const d_chunk = reinterpret_cast<char*>(ctx->d_rays.ptr) + off * ray_bytes;      // the staging buffer holds either record size
        MIRO_CUDA(ctx, cudaMemcpyAsync(d_chunk, rays + off * ray_bytes, m * ray_bytes, cudaMemcpyHostToDevice, ctx->copy_in));
        MIRO_CUDA(ctx, cudaEventRecord(up, ctx->copy_in));
        mark(ctx->copy_in);
        ctx->stream = (k % k_streams == 0) ? user_stream : ctx->trace_aux[k % k_streams - 1];
        ctx->work_lane = (int)(k % k_streams);
        MIRO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, up, 0));
        mark(ctx->stream);
        EventPair p = begin_timing(ctx, true);
        if (hits) {
            if (packed) launch_trace_closest_packed(ctx, reinterpret_cast<const miro_gpu_ray32*>(d_chunk), m, ctx->d_hits.ptr + off);
            else launch_trace_closest(ctx, reinterpret_cast<const miro_gpu_ray*>(d_chunk), m, nullptr, ctx->d_hits.ptr + off);
            if (ctx->scene.prim_map) { k_translate_hits<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_hits.ptr + off, (uint32_t)m, ctx->scene.prim_map); ctx->launches++; }
        }
        else if (packed) launch_trace_any_packed(ctx, reinterpret_cast<const miro_gpu_ray32*>(d_chunk), m, ctx->d_bits.ptr + off / 32);
        else launch_trace_any(ctx, reinterpret_cast<const miro_gpu_ray*>(d_chunk), m, nullptr, ctx->d_bits.ptr + off / 32);
        end_timing(ctx, p);
        MIRO_CUDA(ctx, cudaEventRecord(done, ctx->stream));
        mark(ctx->stream);
        MIRO_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, done, 0));
        if (hits) MIRO_CUDA(ctx, cudaMemcpyAsync(hits + off, ctx->d_hits.ptr + off, m * sizeof(miro_gpu_hit), cudaMemcpyDeviceToHost, ctx->copy_out));
        else MIRO_CUDA(ctx, cudaMemcpyAsync(bits + off / 32, ctx->d_bits.ptr + off / 32, ((m + 31) / 32) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copy_out));
        mark(ctx->copy_out);
        ctx->stream = user_stream; ctx->work_lane = 0;
    }
    MIRO_CUDA(ctx, cudaGetLastError());
    MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
    if (timeline) {     // order of marks: start, then per chunk {h2d end, kernel begin, kernel end, d2h end}
        cudaDeviceSynchronize();
        fprintf(stderr, "timeline n=%zu chunks=%zu (us from start: h2d_end kernel_begin kernel_end d2h_end)\n", n, n_chunks);
        for (size_t i = 1; i < tl.size(); ++i) { float ms = 0; cudaEventElapsedTime(&ms, tl[0], tl[i]); fprintf(stderr, "%8.1f%s", ms * 1e3, (i % 4 == 0) ? "\n" : " "); }
        fprintf(stderr, "\n");
        for (cudaEvent_t e : tl) cudaEventDestroy(e);
    }
    end_timing(ctx, tot);
    MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MIRO_GPU_OK;
}

int miro_gpu_trace_closest(miro_gpu_ctx* ctx, const miro_gpu_ray* rays, size_t n, miro_gpu_hit* hits) {
    if (ctx && n && !hits) return set_error(ctx, MIRO_GPU_EINVAL, "NULL ray/hit buffer");
    return trace_host(ctx, rays, n, hits, nullptr);
}

int miro_gpu_trace_any(miro_gpu_ctx* ctx, const miro_gpu_ray* rays, size_t n, uint32_t* occluded_bits) {
    if (ctx && n && !occluded_bits) return set_error(ctx, MIRO_GPU_EINVAL, "NULL ray/bit buffer");
    return trace_host(ctx, rays, n, nullptr, occluded_bits);
}

int miro_gpu_trace_closest_packed(miro_gpu_ctx* ctx, const miro_gpu_ray32* rays, size_t n, miro_gpu_hit* hits) {
    if (ctx && n && !hits) return set_error(ctx, MIRO_GPU_EINVAL, "NULL ray/hit buffer");
    return trace_host(ctx, rays, n, hits, nullptr, true);
}

int miro_gpu_trace_any_packed(miro_gpu_ctx* ctx, const miro_gpu_ray32* rays, size_t n, uint32_t* occluded_bits) {
    if (ctx && n && !occluded_bits) return set_error(ctx, MIRO_GPU_EINVAL, "NULL ray/bit buffer");
    return trace_host(ctx, rays, n, nullptr, occluded_bits, true);
}

int miro_gpu_pin_host_buffer(miro_gpu_ctx* ctx, void* ptr, size_t bytes) {
    if (!ctx) return MIRO_GPU_EINVAL;
    if (!ptr || !bytes) return set_error(ctx, MIRO_GPU_EINVAL, "miro_gpu_pin_host_buffer: NULL / empty buffer");
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return MIRO_GPU_OK; }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaHostRegister");
    return MIRO_GPU_OK;
}

int miro_gpu_unpin_host_buffer(miro_gpu_ctx* ctx, void* ptr) {
    if (!ctx) return MIRO_GPU_EINVAL;
    if (!ptr) return set_error(ctx, MIRO_GPU_EINVAL, "miro_gpu_unpin_host_buffer: NULL buffer");
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return set_error(ctx, MIRO_GPU_EINVAL, "miro_gpu_unpin_host_buffer: the buffer is not pinned"); }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaHostUnregister");
    return MIRO_GPU_OK;
}

int miro_gpu_set_trace_kernel(miro_gpu_ctx* ctx, int kind) {
    if (!ctx) return MIRO_GPU_EINVAL;
    if (kind != MIRO_GPU_KERNEL_AUTO && kind != MIRO_GPU_KERNEL_WARP && kind != MIRO_GPU_KERNEL_POOL && kind != MIRO_GPU_KERNEL_FLAT)
        return set_error(ctx, MIRO_GPU_EINVAL, "miro_gpu_set_trace_kernel: unknown kernel");
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->trace_kernel_request = kind;
    miro::resolve_trace_kernel(ctx);
    return MIRO_GPU_OK;
}

int miro_gpu_get_trace_kernel(const miro_gpu_ctx* ctx) { return ctx ? ctx->trace_kernel : MIRO_GPU_KERNEL_AUTO; }

int miro_gpu_set_trace_chaining(miro_gpu_ctx* ctx, int on) {
    if (!ctx) return MIRO_GPU_EINVAL;
    ctx->chain_traces = on != 0;
    return MIRO_GPU_OK;
}

int miro_gpu_enable_counting(miro_gpu_ctx* ctx, int enable) {
    if (!ctx) return MIRO_GPU_EINVAL;
    ctx->counting = enable != 0;
    return MIRO_GPU_OK;
}

int miro_gpu_get_counters(miro_gpu_ctx* ctx, miro_gpu_counters* out) {
    if (!ctx || !out) return MIRO_GPU_EINVAL;
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    drain_timing(ctx);
    TraceCounters h;
    MIRO_CUDA(ctx, cudaMemcpy(&h, ctx->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    out->rays_closest = h.rays_closest; out->rays_any = h.rays_any;
    out->nodes_fetched = h.nodes; out->tris_tested = h.tris; out->insts_entered = h.insts;
    out->trace_ms = ctx->trace_ms; out->total_ms = ctx->total_ms;
    out->kernel_launches = ctx->launches;
    return MIRO_GPU_OK;
}

int miro_gpu_reset_counters(miro_gpu_ctx* ctx) {
    if (!ctx) return MIRO_GPU_EINVAL;
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    drain_timing(ctx);
    MIRO_CUDA(ctx, cudaMemset(ctx->d_counters, 0, sizeof(TraceCounters)));
    ctx->trace_ms = ctx->total_ms = 0.0; ctx->launches = 0;
    return MIRO_GPU_OK;
}

}  // extern "C"
