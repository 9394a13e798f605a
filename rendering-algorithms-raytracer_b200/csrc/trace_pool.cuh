// trace_pool.cuh — the POOL traversal kernel (round 2): a warp owns a pool of POOL_SLOTS rays in shared memory and every round
// runs ONE step (node step | leaf step) on up to 32 rays picked out of the pool by phase, so both code paths run with (almost)
// all lanes active.
//
// Why: the persistent-warp kernel (k_trace, miro_gpu_api.cu) keeps one ray per lane in registers and votes per round between a
// node step and a leaf step; the lanes of the minority phase idle.  ncu on the C2 step (profiles/r1_v10_ncu_summary.md,
// profiles/r2_v12_ncu_summary.md): 13 of 32 threads active per instruction — node step 17 lanes, triangle test 8, — issue slots
// 70 % busy, ALU pipe 62 %: the kernel is bound by the instructions it issues for idle lanes.  With twice as many rays as lanes
// per warp the majority phase almost always has >= 32 rays waiting (tools/sched_sim.py: 31 lanes per node round, -39 %
// instructions per ray before the pool's own overhead).
//
// State of a ray = one COLUMN of a structure-of-arrays block in shared memory, [field][slot] (23 words) + its traversal stack
// [entry][slot] (POOL_STACK x 8 bytes; deeper entries go to a per-slot scratch in global memory).  Lane l OWNS slots l and l + 32:
// it reads their references every round (classification by ballot), refills them from the work counter when they are idle and
// writes their results.  WHICH lane advances a slot changes from round to round: the slots of the winning phase are compacted
// by prefix popcount into sel[0..32), lane j loads the fields the step needs of slot sel[j] (10 words for a node step, 21 for a
// leaf step) into registers, runs the SAME node_step / intersect_leaf as the persistent-warp kernel, and stores what changed.
#pragma once
#include "traverse.cuh"

namespace miro {

#ifndef MIRO_POOL_SLOTS
#define MIRO_POOL_SLOTS 48
#endif
#ifndef MIRO_POOL_STACK
#define MIRO_POOL_STACK 8
#endif
#ifndef MIRO_POOL_BLOCK
#define MIRO_POOL_BLOCK 128
#endif
#ifndef MIRO_POOL_MIN_BLOCKS
#define MIRO_POOL_MIN_BLOCKS 6
#endif
#ifndef MIRO_POOL_REFILL
#define MIRO_POOL_REFILL 16
#endif
constexpr int POOL_SLOTS = MIRO_POOL_SLOTS;            // rays per warp (33..64; lane l owns slots l and l + 32)
constexpr int POOL_STACK = MIRO_POOL_STACK;            // stack entries per ray kept in shared memory
constexpr int POOL_BLOCK = MIRO_POOL_BLOCK;            // threads per block
constexpr int POOL_WARPS = POOL_BLOCK / 32;
constexpr int POOL_MIN_BLOCKS = MIRO_POOL_MIN_BLOCKS;  // resident blocks per SM the kernel is compiled for
constexpr int POOL_REFILL = MIRO_POOL_REFILL;          // idle slots that trigger a refill
static_assert(POOL_SLOTS > 32 && POOL_SLOTS <= 64, "lane l owns slot l and, when it exists, slot l + 32");

enum PoolField {
    F_OX, F_OY, F_OZ, F_IX, F_IY, F_IZ, F_TMIN, F_T, F_CUR, F_SP,                         // what a node step reads
    F_DX, F_DY, F_DZ, F_SB, F_SC, F_KZ, F_TIME, F_A, F_B, F_PRIM, F_INST, F_CURINST, F_RAY,   // + what a leaf step reads
    POOL_FIELDS
};
constexpr int POOL_WARP_WORDS = POOL_FIELDS * POOL_SLOTS + 2 * POOL_STACK * POOL_SLOTS + 32;      // state + stack + sel[32]
constexpr size_t POOL_SMEM_BYTES = (size_t)POOL_WARPS * POOL_WARP_WORDS * 4;
using PoolStackT = TraversalStackT<POOL_SLOTS * 8u, POOL_STACK>;

enum { POOL_TRACE_CLOSEST = 0, POOL_TRACE_ANY_BITS = 1, POOL_TRACE_ANY_ACCUM = 2 };

template <int MODE, bool COUNT, bool ALPHA, bool PACKED>
__global__ void __launch_bounds__(POOL_BLOCK, POOL_MIN_BLOCKS)
k_trace_pool(DeviceScene s, const float4* __restrict__ rays, uint32_t n_static, const uint32_t* __restrict__ d_count, uint32_t chunk,
             miro_gpu_hit* __restrict__ hits, uint32_t* __restrict__ bits, const float4* __restrict__ sample_E, float4* __restrict__ slots,
             TraceCounters* __restrict__ ctr, uint32_t* __restrict__ work, unsigned long long* __restrict__ ovf, int ovf_cap) {
    constexpr bool ANY = MODE != POOL_TRACE_CLOSEST;
    constexpr uint32_t RAY_F4 = PACKED ? 2u : 3u;
    extern __shared__ __align__(16) uint32_t pool_smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* const S = pool_smem + warp * POOL_WARP_WORDS;                                  // S[field * POOL_SLOTS + slot]
    unsigned long long* const stack0 = reinterpret_cast<unsigned long long*>(S + POOL_FIELDS * POOL_SLOTS);   // [entry][slot]
    uint32_t* const sel = S + POOL_FIELDS * POOL_SLOTS + 2 * POOL_STACK * POOL_SLOTS;
    unsigned long long* const ovf_warp = ovf + ((size_t)blockIdx.x * POOL_WARPS + warp) * POOL_SLOTS * (size_t)ovf_cap;
    const uint32_t n = d_count ? min(*d_count, n_static) : n_static;
    uint32_t c_nodes = 0, c_tris = 0, c_insts = 0, c_rays = 0;
    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = false;
    bool pending_lo = false, pending_hi = false;       // the owned slot holds a ray whose result has not been written
    const bool has_hi = lane + 32u < (uint32_t)POOL_SLOTS;      // pools of fewer than 64 slots: the upper lanes own one slot only
    S[F_CUR * POOL_SLOTS + lane] = (uint32_t)MIRO_GPU_CHILD_EMPTY;
    if (has_hi) S[F_CUR * POOL_SLOTS + 32 + lane] = (uint32_t)MIRO_GPU_CHILD_EMPTY;
    __syncwarp();

    auto write_result = [&](uint32_t slot) {
        const uint32_t i = S[F_RAY * POOL_SLOTS + slot];
        const int prim = (int)S[F_PRIM * POOL_SLOTS + slot];
        const bool hit = prim >= 0;
        if (MODE == POOL_TRACE_ANY_BITS) { if (hit) atomicOr(bits + (i >> 5), 1u << (i & 31u)); }
        else if (MODE == POOL_TRACE_ANY_ACCUM) {
            if (!hit) { const float4 r2 = __ldcs(rays + (size_t)i * 3 + 2); atomicAdd(slots + (size_t)__float_as_uint(r2.z) * 4, __ldcs(sample_E + i)); }
        } else {
            float* o = reinterpret_cast<float*>(hits + i);
            __stcs(o + 0, hit ? __uint_as_float(S[F_T * POOL_SLOTS + slot]) : -1.0f);
            __stcs(o + 1, hit ? __uint_as_float(S[F_A * POOL_SLOTS + slot]) : 0.f);
            __stcs(o + 2, hit ? __uint_as_float(S[F_B * POOL_SLOTS + slot]) : 0.f);
            __stcs(reinterpret_cast<int*>(o) + 3, prim);
            __stcs(reinterpret_cast<int*>(o) + 4, hit ? (int)S[F_INST * POOL_SLOTS + slot] : -1);
        }
    };
    auto store_ray_space = [&](uint32_t slot, const RaySpace& r) {
        S[F_OX * POOL_SLOTS + slot] = __float_as_uint(r.ox); S[F_OY * POOL_SLOTS + slot] = __float_as_uint(r.oy); S[F_OZ * POOL_SLOTS + slot] = __float_as_uint(r.oz);
        S[F_IX * POOL_SLOTS + slot] = __float_as_uint(r.ix); S[F_IY * POOL_SLOTS + slot] = __float_as_uint(r.iy); S[F_IZ * POOL_SLOTS + slot] = __float_as_uint(r.iz);
        S[F_DX * POOL_SLOTS + slot] = __float_as_uint(r.dx); S[F_DY * POOL_SLOTS + slot] = __float_as_uint(r.dy); S[F_DZ * POOL_SLOTS + slot] = __float_as_uint(r.dz);
        S[F_SB * POOL_SLOTS + slot] = __float_as_uint(r.Sb); S[F_SC * POOL_SLOTS + slot] = __float_as_uint(r.Sc); S[F_KZ * POOL_SLOTS + slot] = (uint32_t)r.kz;
    };

    while (true) {
        // ---- the owner's view of its two slots
        int cur_lo = (int)S[F_CUR * POOL_SLOTS + lane], cur_hi = has_hi ? (int)S[F_CUR * POOL_SLOTS + 32 + lane] : MIRO_GPU_CHILD_EMPTY;
        const uint32_t idle_lo = __ballot_sync(0xffffffffu, cur_lo == MIRO_GPU_CHILD_EMPTY);
        const uint32_t idle_hi = __ballot_sync(0xffffffffu, has_hi && cur_hi == MIRO_GPU_CHILD_EMPTY);
        const int n_idle = __popc(idle_lo) + __popc(idle_hi);
        if (n_idle >= POOL_REFILL || (exhausted && n_idle == POOL_SLOTS)) {
            // results of finished rays leave here, converged, not when each ray finishes
            if (pending_lo && cur_lo == MIRO_GPU_CHILD_EMPTY) { write_result(lane); pending_lo = false; }
            if (pending_hi && cur_hi == MIRO_GPU_CHILD_EMPTY) { write_result(lane + 32u); pending_hi = false; }
            if (exhausted) { if (n_idle == POOL_SLOTS) break; }
            else {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t m = half ? idle_hi : idle_lo;
                    if (m == 0u || exhausted) continue;
                    if (chunk_next == chunk_end) {
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(work, chunk);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        chunk_next = min(base, n); chunk_end = min(base + chunk, n);
                        if (chunk_next == chunk_end) {
                            exhausted = true;
                            if (lane == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
                            continue;
                        } else if (MODE == POOL_TRACE_ANY_BITS) {
                            if (lane == 0) bits[chunk_next >> 5] = 0u;      // a 32-ray claim is one result word, touched by this warp only
                            __syncwarp();
                        }
                    }
                    const uint32_t take = min((uint32_t)__popc(m), chunk_end - chunk_next);
                    const uint32_t rank = __popc(m & lt_mask);
                    const bool mine = ((m >> lane) & 1u) && rank < take;
                    if (mine) {
                        const uint32_t slot = lane + 32u * half;
                        const uint32_t ri = chunk_next + rank;
                        const float4* rp = rays + (size_t)ri * RAY_F4;
                        const float4 r0 = __ldcs(rp), r1 = __ldcs(rp + 1);
                        RaySpace r; r.set(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
                        store_ray_space(slot, r);
                        S[F_TMIN * POOL_SLOTS + slot] = __float_as_uint(r0.w);
                        S[F_TIME * POOL_SLOTS + slot] = PACKED ? 0u : __float_as_uint(__ldcs(reinterpret_cast<const float*>(rp + 2)));
                        S[F_T * POOL_SLOTS + slot] = __float_as_uint(r1.w);
                        S[F_A * POOL_SLOTS + slot] = 0u; S[F_B * POOL_SLOTS + slot] = 0u;
                        S[F_PRIM * POOL_SLOTS + slot] = 0xffffffffu; S[F_INST * POOL_SLOTS + slot] = 0xffffffffu;
                        S[F_CURINST * POOL_SLOTS + slot] = 0xffffffffu; S[F_RAY * POOL_SLOTS + slot] = ri;
                        S[F_SP * POOL_SLOTS + slot] = 0u; S[F_CUR * POOL_SLOTS + slot] = (uint32_t)s.root;
                        if (half) { cur_hi = s.root; pending_hi = true; } else { cur_lo = s.root; pending_lo = true; }
                        c_rays += (r0.w <= r1.w) ? 1u : 0u;      // an empty interval (a light that casts no shadow) is not a Scene::trace call
                    }
                    chunk_next += take;
                }
            }
        }
        // ---- vote over the 64 slots: at an inner node | at a leaf (leaf reference / instance-exit marker) | idle
        const bool node_lo = ref_is_inner(cur_lo), node_hi = ref_is_inner(cur_hi);
        const uint32_t mn_lo = __ballot_sync(0xffffffffu, node_lo), mn_hi = __ballot_sync(0xffffffffu, node_hi);
        const uint32_t ml_lo = __ballot_sync(0xffffffffu, !node_lo && cur_lo != MIRO_GPU_CHILD_EMPTY);
        const uint32_t ml_hi = __ballot_sync(0xffffffffu, !node_hi && cur_hi != MIRO_GPU_CHILD_EMPTY);
        const int n_node = __popc(mn_lo) + __popc(mn_hi), n_leaf = __popc(ml_lo) + __popc(ml_hi);
        if (n_node + n_leaf == 0) continue;            // everything idle: back to the refill (or out)
        const bool node_round = n_node >= n_leaf;
        const uint32_t m_lo = node_round ? mn_lo : ml_lo, m_hi = node_round ? mn_hi : ml_hi;
        // ---- compaction: the first 32 slots of the winning phase, in slot order
        {
            const uint32_t r_lo = __popc(m_lo & lt_mask), r_hi = __popc(m_lo) + __popc(m_hi & lt_mask);
            if (((m_lo >> lane) & 1u)) sel[r_lo] = lane;                 // r_lo < 32 always
            if (((m_hi >> lane) & 1u) && r_hi < 32u) sel[r_hi] = lane + 32u;
        }
        __syncwarp();
        const uint32_t n_sel = min(32u, (uint32_t)(node_round ? n_node : n_leaf));
        if (lane < n_sel) {
            const uint32_t slot = sel[lane];
            Lane L;
            PoolStackT st;
            st.base = (uint32_t)__cvta_generic_to_shared(stack0 + slot);
            st.overflow = ovf_warp + (size_t)slot * (size_t)ovf_cap; st.cap = ovf_cap;
            st.sp = (int)S[F_SP * POOL_SLOTS + slot];
            L.cur = (int)S[F_CUR * POOL_SLOTS + slot];
            L.r.ox = __uint_as_float(S[F_OX * POOL_SLOTS + slot]); L.r.oy = __uint_as_float(S[F_OY * POOL_SLOTS + slot]); L.r.oz = __uint_as_float(S[F_OZ * POOL_SLOTS + slot]);
            L.tmin = __uint_as_float(S[F_TMIN * POOL_SLOTS + slot]);
            L.hit.t = __uint_as_float(S[F_T * POOL_SLOTS + slot]);
            if (node_round) {
                L.r.ix = __uint_as_float(S[F_IX * POOL_SLOTS + slot]); L.r.iy = __uint_as_float(S[F_IY * POOL_SLOTS + slot]); L.r.iz = __uint_as_float(S[F_IZ * POOL_SLOTS + slot]);
                node_step<COUNT>(s, L, st, c_nodes);
            } else {
                L.r.dx = __uint_as_float(S[F_DX * POOL_SLOTS + slot]); L.r.dy = __uint_as_float(S[F_DY * POOL_SLOTS + slot]); L.r.dz = __uint_as_float(S[F_DZ * POOL_SLOTS + slot]);
                L.r.Sb = __uint_as_float(S[F_SB * POOL_SLOTS + slot]); L.r.Sc = __uint_as_float(S[F_SC * POOL_SLOTS + slot]); L.r.kz = (int)S[F_KZ * POOL_SLOTS + slot];
                L.time = __uint_as_float(S[F_TIME * POOL_SLOTS + slot]);
                L.cur_inst = (int)S[F_CURINST * POOL_SLOTS + slot];
                L.ray_idx = S[F_RAY * POOL_SLOTS + slot];
                L.hit.prim = -1;                         // "no new hit in this step" (the slot keeps the ray's best so far)
                const float t_in = L.hit.t;
                bool finished = false, moved = false;
                const uint32_t kind = ((uint32_t)L.cur >> 29) & 3u;
                if (L.cur == STACK_SENTINEL) {            // leaving an instance: back to the world-space ray
                    const float4 w0 = __ldg(rays + (size_t)L.ray_idx * RAY_F4), w1 = __ldg(rays + (size_t)L.ray_idx * RAY_F4 + 1);
                    L.set_ray(w0.x, w0.y, w0.z, w1.x, w1.y, w1.z);
                    L.cur_inst = -1; L.cur = MIRO_GPU_CHILD_EMPTY; moved = true;
                } else {
                    moved = kind == MIRO_GPU_KIND_INST;
                    finished = intersect_leaf<ANY, COUNT, ALPHA>(s, L, st, rays, RAY_F4, c_tris, c_insts);     // true: any-hit found its occluder
                }
                if (!finished && L.cur == MIRO_GPU_CHILD_EMPTY) pop_next(L, st);
                if (finished) L.cur = MIRO_GPU_CHILD_EMPTY;
                if (L.hit.prim >= 0) {                   // a nearer hit (or the occluder) was found in this step
                    S[F_T * POOL_SLOTS + slot] = __float_as_uint(L.hit.t); S[F_A * POOL_SLOTS + slot] = __float_as_uint(L.hit.a); S[F_B * POOL_SLOTS + slot] = __float_as_uint(L.hit.b);
                    S[F_PRIM * POOL_SLOTS + slot] = (uint32_t)L.hit.prim; S[F_INST * POOL_SLOTS + slot] = (uint32_t)L.hit.inst;
                }
                (void)t_in;
                if (moved) { store_ray_space(slot, L.r); S[F_CURINST * POOL_SLOTS + slot] = (uint32_t)L.cur_inst; }
            }
            S[F_CUR * POOL_SLOTS + slot] = (uint32_t)L.cur;
            S[F_SP * POOL_SLOTS + slot] = (uint32_t)st.sp;
        }
        __syncwarp();
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
    __syncthreads();
    if (threadIdx.x == 0) {      // see k_trace: ordered retirement of chained launches, re-arming of this launch's counter pair
        asm volatile("griddepcontrol.wait;" ::: "memory");
        __threadfence();
        if (atomicAdd(work + 1, 1u) == gridDim.x - 1) { work[0] = 0; work[1] = 0; __threadfence(); }
    }
}

}  // namespace miro
