// trace_flat.cuh — the FLAT traversal kernel (round 2): the persistent-warp kernel (k_trace, miro_gpu_api.cu) with the leaf round's
// triangle tests DEALT OUT OVER THE WARP.
//
// Why: in k_trace a leaf round runs `for (i < count) test triangle i` in every lane that waits at a leaf — the loop runs at the
// length of the largest leaf with the lanes of the minority phase idle: ncu on the C2 step (profiles/r2_ncu_summary.md) has the
// triangle test at 8.4 of 32 lanes (1 277 k warp-level executions for 10.7 M tests) and the reference-order Moller-Trumbore block
// behind it at 1.7 lanes — 39 % of the instructions the kernel issues.  The pool / duo kernels bought lanes by putting more rays
// behind a warp and paid in resident warps; here the number of rays per warp stays 32 and 64 registers x 8 blocks stay.  What moves is
// the WORK of a leaf round: the (ray, triangle) pairs of all waiting leaves are numbered by a prefix sum over the leaves' counts,
// lane w tests pair w (+ 32 per pass) for whichever ray owns it, and every owner gathers the nearest accepted distance of its <= 4
// pairs by shuffles.  ~30 pairs wait in an average leaf round: one or two passes at (almost) full width instead of ~3.6 loop
// iterations at 8 lanes.
//
// What a worker needs of the owner's ray: the origin (3 shuffles), the shear constants Sb, Sc, kz of the watertight test and the
// direction for the reference-order t / a / b — the last six live in SHARED memory ([field][thread]: the owner's column is read by
// its workers, conflict free), which also takes them out of the registers every lane holds across node rounds.  The range test
// tMin <= t < hit.t is the owner's (it scans its pairs in triangle order with a strict compare, so the result is the sequential
// loop's: nearest crossing, lowest triangle on exact ties); hits are byte-identical to k_trace's.
// Leaves of other kinds (motion-blur triangles, instance entry / exit), leaves of more than four triangles and scenes with alpha
// cut-outs take the sequential path of traverse.cuh.
#pragma once
#include "traverse.cuh"

namespace miro {

// Leaves of up to 4 triangles are dealt out (the host builder's MAX_LEAF_SIZE).  Dealing leaves of up to 8 (the ABI's limit: a
// fourth ballot over bit 2 of count - 1, eight table entries, an eight-step gather) was built and measured on trees with larger
// leaves — MIRO_BVH_MAX_LEAF 4 / 5 / 6 / 8: 7 543 / 7 603 / 7 599 / 7 459 Mrays/s on the C2 step against 7 792 for this build on
// leaves of <= 4 (smaller leaves lose as well: <= 3: 7 303, <= 2: 6 487) — and removed.
constexpr int FLAT_MAX_LEAF = 4;                                 // leaves of up to this many triangles are dealt out
// A leaf round is run when  n_leaf * NUM > n_node * DEN: dealt leaf rounds are cheap at any number of waiting leaves and return
// their lanes to the node phase, so they are favoured (measured on the C2 step: 1:1 7 107, 3:2 7 489, 2:1 7 604, 5:2 7 587,
// 3:1 7 529 Mrays/s).
#ifndef MIRO_FLAT_BIAS_NUM
#define MIRO_FLAT_BIAS_NUM 2
#endif
#ifndef MIRO_FLAT_BIAS_DEN
#define MIRO_FLAT_BIAS_DEN 1
#endif
// With more than this many pairs waiting (a coherent batch: most lanes of the warp reach their leaves together) the leaves take
// the sequential per-lane loop of traverse.cuh, which is efficient exactly then; dealing serves 32 pairs per round.
#ifndef MIRO_FLAT_OTHER_NUM
#define MIRO_FLAT_OTHER_NUM 1
#endif
#ifndef MIRO_FLAT_OTHER_DEN
#define MIRO_FLAT_OTHER_DEN 1
#endif
#ifndef MIRO_FLAT_KIND_VOTE
#define MIRO_FLAT_KIND_VOTE 0
#endif
#ifndef MIRO_FLAT_SEQ_PAIRS
#define MIRO_FLAT_SEQ_PAIRS 48
#endif
// Shared memory of a thread = one 8-byte COLUMN of [row][thread of the block] (the traversal stack's layout, TraversalStack::STRIDE
// between rows): rows 0 .. SMEM_STACK-1 the stack, then three rows of ray constants and one row for the pair table, so every
// address is st.base + a compile-time offset (+ 8 x lane distance for another lane's column) and costs no register.
constexpr uint32_t FLAT_ROW = TraversalStack::STRIDE;
constexpr uint32_t FLAT_ROW_DXY = SMEM_STACK * FLAT_ROW;          // {dx, dy}
constexpr uint32_t FLAT_ROW_DZSB = (SMEM_STACK + 1) * FLAT_ROW;   // {dz, Sb}
constexpr uint32_t FLAT_ROW_SCKZ = (SMEM_STACK + 2) * FLAT_ROW;   // {Sc, kz}
constexpr uint32_t FLAT_ROW_PAIR = (SMEM_STACK + 3) * FLAT_ROW;   // pair table of the warp: entry p in the column of lane p (low word)
constexpr int FLAT_ROWS = SMEM_STACK + 4;

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) { asm volatile("st.shared.v2.b32 [%0], {%1, %2};" :: "r"(addr), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t& a, uint32_t& b) { asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void sts32_if(bool c, uint32_t addr, uint32_t v) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.shared.b32 [%0], %1; }" :: "r"(addr), "r"(v), "r"((int)c) : "memory");
}

struct FlatRay { float ox, oy, oz, ix, iy, iz; };      // what the node step reads

// One ray slot of the flat kernel: the node step's part of the ray in registers, the triangle test's part in shared memory.
struct FlatLane {
    FlatRay r;
    float tmin, time;
    HitRec hit;
    int32_t cur, cur_inst;
    uint32_t ray_idx;
    uint32_t sbase;        // shared address of this thread's column (== the traversal stack's base)
    __device__ __forceinline__ void set_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
        RaySpace s; s.set(ox, oy, oz, dx, dy, dz);
        r.ox = ox; r.oy = oy; r.oz = oz; r.ix = s.ix; r.iy = s.iy; r.iz = s.iz;
        sts64(sbase + FLAT_ROW_DXY, __float_as_uint(dx), __float_as_uint(dy));
        sts64(sbase + FLAT_ROW_DZSB, __float_as_uint(dz), __float_as_uint(s.Sb));
        sts64(sbase + FLAT_ROW_SCKZ, __float_as_uint(s.Sc), (uint32_t)s.kz);
    }
    // the whole ray, for the sequential leaf paths
    __device__ __forceinline__ RaySpace ray_space() const {
        RaySpace s;
        s.ox = r.ox; s.oy = r.oy; s.oz = r.oz; s.ix = r.ix; s.iy = r.iy; s.iz = r.iz;
        uint32_t a, b;
        lds64(sbase + FLAT_ROW_DXY, a, b); s.dx = __uint_as_float(a); s.dy = __uint_as_float(b);
        lds64(sbase + FLAT_ROW_DZSB, a, b); s.dz = __uint_as_float(a); s.Sb = __uint_as_float(b);
        lds64(sbase + FLAT_ROW_SCKZ, a, b); s.Sc = __uint_as_float(a); s.kz = (int)b;
        return s;
    }
};

// One (ray, triangle) pair: the watertight edge test of intersect_tri (traverse.cuh) for the ray whose column is `ocol`, origin
// (ox, oy, oz).  Returns the crossing's distance when it lies in [tmin, tmax) — the reference-order Moller-Trumbore t (with a, b),
// or for T_ONLY queries the edge functions' own — and +inf otherwise.
template <bool T_ONLY>
__device__ __forceinline__ float flat_pair_test(uint32_t ocol, float ox, float oy, float oz, float tmin, float tmax, float4 p0, float4 p1, float4 p2, float& a_out, float& b_out) {
    const float INF = __int_as_float(0x7f800000);
    uint32_t w0, w1, w2, w3;
    lds64(ocol + FLAT_ROW_DZSB, w0, w1);
    lds64(ocol + FLAT_ROW_SCKZ, w2, w3);
    const float dz = __uint_as_float(w0), Sb = __uint_as_float(w1), Sc = __uint_as_float(w2);
    const int kz = (int)w3;
    float Aa = p0.x - ox, Ab = p0.y - oy, Ac = p0.z - oz;
    float Ba = p1.x - ox, Bb = p1.y - oy, Bc = p1.z - oz;
    float Ca = p2.x - ox, Cb = p2.y - oy, Cc = p2.z - oz;
    rotate_to_ray_axes(kz, Aa, Ab, Ac, Ba, Bb, Bc, Ca, Cb, Cc);
    const float Ax = __fmaf_rn(-Sb, Aa, Ab), Ay = __fmaf_rn(-Sc, Aa, Ac);
    const float Bx = __fmaf_rn(-Sb, Ba, Bb), By = __fmaf_rn(-Sc, Ba, Bc);
    const float Cx = __fmaf_rn(-Sb, Ca, Cb), Cy = __fmaf_rn(-Sc, Ca, Cc);
    const float U = diff_of_products(Cx, By, Cy, Bx);
    const float V = diff_of_products(Ax, Cy, Ay, Cx);
    const float W = diff_of_products(Bx, Ay, By, Ax);
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return INF;
    const float det = U + V + W;
    if (det == 0.0f) return INF;
    uint32_t wx, wy;
    lds64(ocol + FLAT_ROW_DXY, wx, wy);
    const float dx = __uint_as_float(wx), dy = __uint_as_float(wy);
    float t;
    if (T_ONLY) {
        const float dkz = fsel(kz == 0, dx, fsel(kz == 1, dy, dz));
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(det * dkz));
        a_out = 0.f; b_out = 0.f;
        t = __fmaf_rn(U, Aa, __fmaf_rn(V, Ba, W * Ca)) * inv;
    } else {
        RaySpace rs;
        rs.ox = ox; rs.oy = oy; rs.oz = oz; rs.dx = dx; rs.dy = dy; rs.dz = dz;
        if (!moller_trumbore_reference(rs, p0, p1, p2, t, a_out, b_out)) return INF;
    }
    return (t >= tmin && t < tmax) ? t : INF;
}

enum { FLAT_TRACE_CLOSEST = 0, FLAT_TRACE_ANY_BITS = 1, FLAT_TRACE_ANY_ACCUM = 2 };

// Same contract, parameters, work claiming, result writing and launch chaining as k_trace (miro_gpu_api.cu); the rounds differ.
// MB: the scene holds motion-blur triangles (their leaves are dealt as well; a separate instantiation, so that scenes without
// them do not carry the registers of the two-pose fetch).
template <int MODE, bool COUNT, bool ALPHA, bool PACKED, bool MB>
__global__ void __launch_bounds__(TRACE_BLOCK, TRACE_MIN_BLOCKS)
k_trace_flat(DeviceScene s, const float4* __restrict__ rays, uint32_t n_static, const uint32_t* __restrict__ d_count, uint32_t chunk,
             miro_gpu_hit* __restrict__ hits, uint32_t* __restrict__ bits, const float4* __restrict__ sample_E, float4* __restrict__ slots,
             TraceCounters* __restrict__ ctr, uint32_t* __restrict__ work) {
    constexpr bool ANY = MODE != FLAT_TRACE_CLOSEST;
    constexpr bool DEAL = !ALPHA;                       // alpha cut-outs are evaluated per candidate in triangle order: sequential path
    constexpr uint32_t RAY_F4 = PACKED ? 2u : 3u;
    __shared__ unsigned long long stack[FLAT_ROWS * TRACE_BLOCK];      // [row][thread]: stack entries, ray constants, pair table (see FLAT_ROW_*)
    const uint32_t n = d_count ? min(*d_count, n_static) : n_static;
    constexpr bool has_mb = MB;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t c_nodes = 0, c_tris = 0, c_insts = 0, c_rays = 0;
    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = false;
    FlatLane L; L.cur = MIRO_GPU_CHILD_EMPTY; L.ray_idx = 0; L.cur_inst = -1; L.tmin = 0.f; L.time = 0.f;
    L.hit.t = 0.f; L.hit.a = L.hit.b = 0.f; L.hit.prim = -1; L.hit.inst = -1;
    unsigned long long overflow[LMEM_STACK];
    TraversalStack st; st.init(stack + threadIdx.x, overflow, LMEM_STACK);
    L.sbase = st.base;
    L.set_ray(0.f, 0.f, 0.f, 0.f, 0.f, 1.f);

    auto write_result = [&]() {
        const uint32_t i = L.ray_idx;
        const bool hit = L.hit.prim >= 0;
        if (MODE == FLAT_TRACE_ANY_BITS) { if (hit) atomicOr(bits + (i >> 5), 1u << (i & 31u)); }
        else if (MODE == FLAT_TRACE_ANY_ACCUM) {
            if (!hit) { const float4 r2 = __ldcs(rays + (size_t)i * 3 + 2); atomicAdd(slots + (size_t)__float_as_uint(r2.z) * 4, __ldcs(sample_E + i)); }
        } else {
            float* o = reinterpret_cast<float*>(hits + i);
            __stcs(o + 0, hit ? L.hit.t : -1.0f); __stcs(o + 1, hit ? L.hit.a : 0.f); __stcs(o + 2, hit ? L.hit.b : 0.f);
            __stcs(reinterpret_cast<int*>(o) + 3, L.hit.prim); __stcs(reinterpret_cast<int*>(o) + 4, hit ? L.hit.inst : -1);
        }
    };
    bool pending = false;

    while (true) {
        // ---- refill (as k_trace)
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
                        if (lane == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
                    } else if (MODE == FLAT_TRACE_ANY_BITS) {
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
                    c_rays += (r0.w <= r1.w) ? 1u : 0u;
                }
                chunk_next += take; n_idle -= (int)take;
                if (idle == 0xffffffffu && take == 0) continue;
            }
        }
        // ---- vote
        const bool at_node = ref_is_inner(L.cur);
        const bool at_leaf = !at_node && L.cur != MIRO_GPU_CHILD_EMPTY;
        const int n_node = __popc(__ballot_sync(0xffffffffu, at_node)), n_leaf = 32 - n_idle - n_node;
        bool finished = false;
#if MIRO_FLAT_KIND_VOTE
        // only the leaves that are dealt (static triangles) get the bias; instance entries / exits and motion-blur leaves run at the
        // lane count they have, like a node step, and weigh 1:1
        const int n_tri = __popc(__ballot_sync(0xffffffffu, DEAL && (((uint32_t)L.cur >> 28) == 8u || (has_mb && ((uint32_t)L.cur >> 28) == 10u))));
        if (n_node * (MIRO_FLAT_BIAS_DEN * MIRO_FLAT_OTHER_DEN) >= n_tri * (MIRO_FLAT_BIAS_NUM * MIRO_FLAT_OTHER_DEN) + (n_leaf - n_tri) * (MIRO_FLAT_OTHER_NUM * MIRO_FLAT_BIAS_DEN)) {
#else
        if (n_node * MIRO_FLAT_BIAS_DEN >= n_leaf * MIRO_FLAT_BIAS_NUM) {
#endif
            if (at_node) { node_step<COUNT>(s, L, st, c_nodes); finished = L.cur == MIRO_GPU_CHILD_EMPTY; }
        } else {
            // ---- leaf round: every lane of the warp takes part as a worker
            const uint32_t u = (uint32_t)L.cur;
            // bits 31..28 == 1000: a leaf reference (bit 31; the instance-exit marker and EMPTY lack it) of static triangles (kind 0)
            // with count - 1 < 4 (larger leaves: sequential path below)
            // ... or of motion-blur triangles (kind 1: bits 31..28 == 1010), dealt the same way: the worker lerps the two poses at the
            // owner's time first (src/BVH.cpp:1320-1335)
            const bool mb_leaf = DEAL && has_mb && (u >> 28) == 10u;
            const bool tri_leaf = (DEAL && (u >> 28) == 8u) || mb_leaf;
            const uint32_t cnt = tri_leaf ? ((u >> MIRO_GPU_LEAF_INDEX_BITS) & 3u) + 1u : 0u;
            const bool c0 = tri_leaf && (u & (1u << MIRO_GPU_LEAF_INDEX_BITS)) != 0u, c1 = tri_leaf && (u & (2u << MIRO_GPU_LEAF_INDEX_BITS)) != 0u;      // bits of count - 1
            bool dealt = false, sequential = !DEAL;
            if (DEAL) {
                // pair numbering, owner-major: S = pairs of the lanes below = sum of 1 + (count - 1) over them, from three ballots
                const uint32_t BT = __ballot_sync(0xffffffffu, tri_leaf), B0 = __ballot_sync(0xffffffffu, c0), B1 = __ballot_sync(0xffffffffu, c1);
                const uint32_t total = __popc(BT) + __popc(B0) + 2u * __popc(B1);
                if (total > (uint32_t)MIRO_FLAT_SEQ_PAIRS) sequential = true;
                else if (total != 0u) {                         // (0: a round of instance entries / exits / motion-blur leaves only, nothing to deal)
                    const uint32_t S = __popc(BT & lt_mask) + __popc(B0 & lt_mask) + 2u * __popc(B1 & lt_mask);
                    // ONE pass of 32 pairs per round: an owner whose pairs do not all fit keeps waiting at its leaf (the lowest
                    // waiting lane always fits, so every leaf is served)
                    dealt = tri_leaf && S + cnt <= 32u;
                    // pair table: entry p (in the column of lane p) = triangle index | owner lane << 26
                    const uint32_t word = (u & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u)) | (lane << MIRO_GPU_LEAF_INDEX_BITS) | (mb_leaf ? 0x80000000u : 0u);
                    const uint32_t ent = st.base + FLAT_ROW_PAIR + (S - lane) * 8u;      // entry S, relative to this lane's own column
                    // (the one owner that straddles entry 32 writes its first pairs too: tested for nothing, which is cheaper than
                    // finding out where the served pairs end)
                    sts32_if(tri_leaf && S < 32u, ent, word);
                    sts32_if(cnt >= 2u && S + 1u < 32u, ent + 8u, word + 1u);
                    sts32_if(cnt >= 3u && S + 2u < 32u, ent + 16u, word + 2u);
                    sts32_if(cnt >= 4u && S + 3u < 32u, ent + 24u, word + 3u);
                    __syncwarp();
                    const bool valid = lane < total;
                    const uint32_t pw = valid ? lds32(st.base + FLAT_ROW_PAIR) : 0u;
                    const uint32_t owner = (pw >> MIRO_GPU_LEAF_INDEX_BITS) & 31u;
                    const float wox = __shfl_sync(0xffffffffu, L.r.ox, owner), woy = __shfl_sync(0xffffffffu, L.r.oy, owner), woz = __shfl_sync(0xffffffffu, L.r.oz, owner);
                    const float wtmin = __shfl_sync(0xffffffffu, L.tmin, owner), wtmax = __shfl_sync(0xffffffffu, L.hit.t, owner);
                    float ht = __int_as_float(0x7f800000), ha = 0.f, hb = 0.f;
                    const float wtime = has_mb ? __shfl_sync(0xffffffffu, L.time, owner) : 0.f;
                    if (valid) {
                        float4 p0, p1, p2;
                        const uint32_t idx = pw & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u);
                        if (has_mb && (int32_t)pw < 0) {
                            // time * pose2 + (1 - time) * pose1 with both products rounded, as intersect_leaf has it
                            const float4* t = s.mbtris + (size_t)idx * 6;
                            const float w1 = wtime, w0 = __fsub_rn(1.0f, wtime);
#define MIRO_LERP(B, A) __fadd_rn(__fmul_rn(w1, B), __fmul_rn(w0, A))
                            { const float4 a = __ldg(t), b = __ldg(t + 3); p0.x = MIRO_LERP(b.x, a.x); p0.y = MIRO_LERP(b.y, a.y); p0.z = MIRO_LERP(b.z, a.z); }
                            { const float4 a = __ldg(t + 1), b = __ldg(t + 4); p1.x = MIRO_LERP(b.x, a.x); p1.y = MIRO_LERP(b.y, a.y); p1.z = MIRO_LERP(b.z, a.z); }
                            { const float4 a = __ldg(t + 2), b = __ldg(t + 5); p2.x = MIRO_LERP(b.x, a.x); p2.y = MIRO_LERP(b.y, a.y); p2.z = MIRO_LERP(b.z, a.z); }
#undef MIRO_LERP
                        } else load_tri(s.tris + (size_t)idx * TRI_F4, p0, p1, p2);
                        ht = flat_pair_test<ANY>(st.base + (owner - lane) * 8u, wox, woy, woz, wtmin, wtmax, p0, p1, p2, ha, hb);
                    }
                    const uint32_t hitmask = __ballot_sync(0xffffffffu, ht != __int_as_float(0x7f800000));
                    if (ANY) {
                        // an occluder among the owner's pairs (lanes S .. S + cnt - 1) ends the query
                        if (dealt && ((hitmask >> S) & ((1u << cnt) - 1u)) != 0u) L.hit.prim = (int32_t)(u & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u));
                    } else if (hitmask != 0u) {
                        // the owner scans its pairs in triangle order (strict compare: the sequential loop's result — nearest
                        // crossing, lowest triangle on exact ties); lanes beyond its count re-read its last pair
                        const uint32_t last = cnt ? cnt - 1u : 0u;
                        float best = L.hit.t; uint32_t bsrc = S;
#pragma unroll
                        for (uint32_t k = 0; k < (uint32_t)FLAT_MAX_LEAF; ++k) {
                            const uint32_t src = S + min(k, last);
                            const float tk = __shfl_sync(0xffffffffu, ht, src);
                            if (tk < best) { best = tk; bsrc = src; }
                        }
                        const float fa = __shfl_sync(0xffffffffu, ha, bsrc), fb = __shfl_sync(0xffffffffu, hb, bsrc);
                        if (dealt && best < L.hit.t) {
                            L.hit.t = best; L.hit.a = fa; L.hit.b = fb;
                            L.hit.prim = (int32_t)((mb_leaf ? s.n_tris : 0u) + (u & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u)) + (bsrc - S)); L.hit.inst = L.cur_inst;
                        }
                    }
                    __syncwarp();      // the pair table is rewritten by the next leaf round
                }
            }
            if (dealt) {
                if (COUNT) c_tris += cnt;
                L.cur = MIRO_GPU_CHILD_EMPTY;
                finished = ANY && L.hit.prim >= 0;
            } else if (at_leaf && (!tri_leaf || sequential)) {
                if (L.cur == STACK_SENTINEL) {            // leaving an instance: back to the world-space ray
                    const float4 w0 = __ldg(rays + (size_t)L.ray_idx * RAY_F4), w1 = __ldg(rays + (size_t)L.ray_idx * RAY_F4 + 1);
                    L.set_ray(w0.x, w0.y, w0.z, w1.x, w1.y, w1.z);
                    L.cur_inst = -1; L.cur = MIRO_GPU_CHILD_EMPTY;
                } else finished = intersect_leaf<ANY, COUNT, ALPHA>(s, L, st, rays, RAY_F4, c_tris, c_insts);
            }
            if (at_leaf && !finished && L.cur == MIRO_GPU_CHILD_EMPTY) { pop_next(L, st); finished = L.cur == MIRO_GPU_CHILD_EMPTY; }
        }
        if (finished) L.cur = MIRO_GPU_CHILD_EMPTY;
    }
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
    if (threadIdx.x == 0) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        __threadfence();
        if (atomicAdd(work + 1, 1u) == gridDim.x - 1) { work[0] = 0; work[1] = 0; __threadfence(); }
    }
}

}  // namespace miro
