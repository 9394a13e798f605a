// traverse.cuh — wide-BVH traversal and watertight ray/triangle intersection (device code, sm_100a).
//
// Replaces, on the GPU:
//   BVH::intersect (QBVH branch)   reference src/BVH.cpp:1128-1178
//   QBVH_Node::intersect           reference src/BVH.cpp:391-414   (4-wide slab test)
//   intersect4                     reference src/BVH.cpp:1298-1459 (Moller-Trumbore, 4 triangles)
//   ProxyObject::intersect         reference src/ProxyObject.cpp:76-95 (instancing, t shared)
//   MB lanes of intersect4         reference src/BVH.cpp:1316-1335 (two-pose lerp at ray.time)
//
// Design (B200-first, not a translation):
//   * PERSISTENT WARPS with dynamic ray fetch: a warp owns 32 ray slots; lanes whose ray has
//     finished are refilled from a global work counter (claimed in per-warp chunks) as soon as
//     enough of them are idle, so a warp never runs at the length of its longest ray with the
//     other lanes empty (v1's one-thread-per-ray kernel executed 2-9 of 32 lanes per instruction);
//   * MAJORITY-PHASE scheduling: every lane is either at an inner node or at a leaf; each round
//     the warp votes (__ballot_sync) and executes ONE step of the kind most lanes wait for —
//     a node step (4 slab tests, sort, push) or a leaf step (<= 4 triangles / instance entry) —
//     so both code paths run with most lanes active instead of a while-while loop whose inner
//     loop runs at the length of the slowest lane (measured: 5.5 of 32 lanes in the node test);
//   * 128-byte BVH4 nodes fetched with seven 16-byte vector loads through the read-only path (the
//     whole tree lives in the 126 MB L2; hot top levels in L1);
//   * children are visited nearest-first (4-element sorting network) and the deferred ones go
//     to a per-thread stack in SHARED memory laid out [entry][lane] (conflict-free, "warp
//     coherent"); entries carry their entry distance so popped sub-trees behind the current
//     hit are culled without a fetch.  The reference visits children unordered (0..3);
//     closest-hit results are order independent except for exact-t ties;
//   * the triangle test is a watertight edge-function test in ray space (shear + scale as in
//     Woop/Benthin/Wald 2013) evaluated in FP32 with error-free products (Kahan) so the SIGN of
//     every edge function is exact: neighbouring triangles agree on shared edges and no ray
//     slips between them (the reference's test is not watertight — crack pixels in its Cornell
//     render).  Once a triangle is accepted, t/a/b are computed with the reference's own
//     Moller-Trumbore expressions so they agree with it to rounding.  Same acceptance set as the reference: two-sided, edges inclusive,
//     tMin <= t < current hit.t; barycentrics a,b are the weights of vertex 1 and vertex 2;
//   * instances: one level (as the reference), ray transformed by the 3x4 inverse, direction NOT
//     renormalised so t is shared with the parent space; a sentinel on the stack restores the
//     world-space ray;
//   * zero direction components use the reference's +-1e12 reciprocal (src/Ray.h:79-90) to
//     avoid 0*inf NaNs in the slab test.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/miro_gpu.h"

namespace miro {

constexpr int TRACE_BLOCK = 128;        // threads per block of the traversal kernels
constexpr int TRACE_MIN_BLOCKS = 6;     // resident blocks per SM the kernels are compiled for (register budget)
#ifndef MIRO_TRACE_REFILL
#define MIRO_TRACE_REFILL 8
#endif
constexpr int TRACE_REFILL = MIRO_TRACE_REFILL;
// a node round is run when  n_node * DEN >= n_leaf * NUM  (NUM/DEN < 1 favours node rounds: leaf rounds cost more and fill up while waiting)
#ifndef MIRO_NODE_BIAS_NUM
#define MIRO_NODE_BIAS_NUM 1
#endif
#ifndef MIRO_NODE_BIAS_DEN
#define MIRO_NODE_BIAS_DEN 1
#endif
constexpr int TRACE_NODE_BIAS_NUM = MIRO_NODE_BIAS_NUM, TRACE_NODE_BIAS_DEN = MIRO_NODE_BIAS_DEN;   // idle lanes in a warp that trigger a refill from the work counter
constexpr int SMEM_STACK = 24;          // per-thread stack entries kept in shared memory
constexpr int LMEM_STACK = 72;          // overflow entries (local memory, touched only by very deep trees)
constexpr int32_t STACK_SENTINEL = 0x7ffffffe;   // "leave instance" marker

struct DeviceScene {
    const float4* nodes;     // 8 x float4 per node
    const float4* tris;      // 3 x float4 per triangle
    const float4* mbtris;    // 6 x float4 per motion-blur triangle
    const float4* insts;     // 4 x float4 per instance
    int32_t root;
    uint32_t n_tris;
};

struct TraceCounters {
    unsigned long long rays_closest, rays_any, nodes, tris, insts;
};

struct HitRec {
    float t, a, b;
    int32_t prim, inst;
};

__device__ __forceinline__ float safe_rcp_dir(float d) {
    // src/Ray.h:79-90: 1/d, with d == 0 mapped to +-MIRO_TMAX by the sign of the IEEE quotient
    if (d == 0.0f) return (__float_as_uint(d) >> 31) ? -MIRO_GPU_TMAX : MIRO_GPU_TMAX;
    return 1.0f / d;
}

__device__ __forceinline__ float sel3(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

// error-free a*b - c*d (Kahan): relative error <= 1.5 ulp, sign always exact, exactly 0 when the true value is 0
__device__ __forceinline__ float diff_of_products(float a, float b, float c, float d) {
    float w = __fmul_rn(c, d);
    float e = __fmaf_rn(-c, d, w);
    float f = __fmaf_rn(a, b, -w);
    return __fadd_rn(f, e);
}

struct RaySpace {
    float ox, oy, oz;
    float dx, dy, dz;
    float ix, iy, iz;      // reciprocal direction (slab test)
    float Sx, Sy, Sz;      // shear / scale of the watertight test
    int kx, ky, kz;

    __device__ __forceinline__ void set(float ox_, float oy_, float oz_, float dx_, float dy_, float dz_) {
        ox = ox_; oy = oy_; oz = oz_; dx = dx_; dy = dy_; dz = dz_;
        ix = safe_rcp_dir(dx); iy = safe_rcp_dir(dy); iz = safe_rcp_dir(dz);
        float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
        kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
        kx = kz + 1; if (kx == 3) kx = 0;
        ky = kx + 1; if (ky == 3) ky = 0;
        float dkz = sel3(dx, dy, dz, kz);
        if (dkz < 0.0f) { int t = kx; kx = ky; ky = t; }
        float rz = 1.0f / dkz;
        Sx = sel3(dx, dy, dz, kx) * rz;
        Sy = sel3(dx, dy, dz, ky) * rz;
        Sz = rz;
    }
};

// Watertight two-sided ray/triangle test.  Returns true and updates (t,a,b) when tmin <= t < tmax.
__device__ __forceinline__ bool intersect_tri(const RaySpace& r, float tmin, float tmax,
                                              float4 p0, float4 p1, float4 p2, float& t_out, float& a_out, float& b_out) {
    const float Ax_ = p0.x - r.ox, Ay_ = p0.y - r.oy, Az_ = p0.z - r.oz;
    const float Bx_ = p1.x - r.ox, By_ = p1.y - r.oy, Bz_ = p1.z - r.oz;
    const float Cx_ = p2.x - r.ox, Cy_ = p2.y - r.oy, Cz_ = p2.z - r.oz;
    const float Akz = sel3(Ax_, Ay_, Az_, r.kz), Bkz = sel3(Bx_, By_, Bz_, r.kz), Ckz = sel3(Cx_, Cy_, Cz_, r.kz);
    const float Ax = __fmaf_rn(-r.Sx, Akz, sel3(Ax_, Ay_, Az_, r.kx));
    const float Ay = __fmaf_rn(-r.Sy, Akz, sel3(Ax_, Ay_, Az_, r.ky));
    const float Bx = __fmaf_rn(-r.Sx, Bkz, sel3(Bx_, By_, Bz_, r.kx));
    const float By = __fmaf_rn(-r.Sy, Bkz, sel3(Bx_, By_, Bz_, r.ky));
    const float Cx = __fmaf_rn(-r.Sx, Ckz, sel3(Cx_, Cy_, Cz_, r.kx));
    const float Cy = __fmaf_rn(-r.Sy, Ckz, sel3(Cx_, Cy_, Cz_, r.ky));
    const float U = diff_of_products(Cx, By, Cy, Bx);
    const float V = diff_of_products(Ax, Cy, Ay, Cx);
    const float W = diff_of_products(Bx, Ay, By, Ax);
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    // The ray passes through the triangle (decided watertight, above).  Distance and barycentrics are then
    // evaluated with the reference's Moller-Trumbore expressions in the reference's operation order
    // (src/BVH.cpp:1343-1369; SoADot = x*x' + (y*y' + z*z'), no FMA contraction), so t, a, b agree with it
    // to rounding instead of differing by the conditioning of two different algorithms on sliver triangles.
    const float e0x = __fsub_rn(p1.x, p0.x), e0y = __fsub_rn(p1.y, p0.y), e0z = __fsub_rn(p1.z, p0.z);
    const float e1x = __fsub_rn(p2.x, p0.x), e1y = __fsub_rn(p2.y, p0.y), e1z = __fsub_rn(p2.z, p0.z);
    const float px = __fsub_rn(__fmul_rn(r.dy, e1z), __fmul_rn(r.dz, e1y));
    const float py = -__fsub_rn(__fmul_rn(r.dx, e1z), __fmul_rn(r.dz, e1x));
    const float pz = __fsub_rn(__fmul_rn(r.dx, e1y), __fmul_rn(r.dy, e1x));
    const float mdet = __fadd_rn(__fmul_rn(e0x, px), __fadd_rn(__fmul_rn(e0y, py), __fmul_rn(e0z, pz)));
    if (mdet == 0.0f) return false;
    const float inv = __frcp_rn(mdet);
    const float tx = __fsub_rn(r.ox, p0.x), ty = __fsub_rn(r.oy, p0.y), tz = __fsub_rn(r.oz, p0.z);
    const float qx = __fsub_rn(__fmul_rn(ty, e0z), __fmul_rn(tz, e0y));
    const float qy = -__fsub_rn(__fmul_rn(tx, e0z), __fmul_rn(tz, e0x));
    const float qz = __fsub_rn(__fmul_rn(tx, e0y), __fmul_rn(ty, e0x));
    const float t = __fmul_rn(inv, __fadd_rn(__fmul_rn(e1x, qx), __fadd_rn(__fmul_rn(e1y, qy), __fmul_rn(e1z, qz))));
    if (!(t >= tmin && t < tmax)) return false;
    const float a = __fmul_rn(inv, __fadd_rn(__fmul_rn(tx, px), __fadd_rn(__fmul_rn(ty, py), __fmul_rn(tz, pz))));
    const float b = __fmul_rn(inv, __fadd_rn(__fmul_rn(r.dx, qx), __fadd_rn(__fmul_rn(r.dy, qy), __fmul_rn(r.dz, qz))));
    t_out = t;
    a_out = fminf(fmaxf(a, 0.0f), 1.0f);     // inside by the edge test: clamp the rounding residue
    b_out = fminf(fmaxf(b, 0.0f), 1.0f - a_out);
    return true;
}

struct StackEntry { int32_t ref; float t; };

// Per-lane traversal stack: first SMEM_STACK entries in shared memory ([entry][lane] layout, conflict
// free), the rest in local memory (touched only by very deep trees; depth is validated at upload).
struct TraversalStack {
    unsigned long long* smem;   // base + threadIdx.x, stride = TRACE_BLOCK
    unsigned long long lmem[LMEM_STACK];
    int sp;
    __device__ __forceinline__ void push(int32_t ref, float t) {
        unsigned long long v = ((unsigned long long)__float_as_uint(t) << 32) | (uint32_t)ref;
        if (sp < SMEM_STACK) smem[sp * TRACE_BLOCK] = v;
        else if (sp - SMEM_STACK < LMEM_STACK) lmem[sp - SMEM_STACK] = v;
        ++sp;
    }
    __device__ __forceinline__ StackEntry pop() {
        --sp;
        StackEntry e;
        if (sp >= SMEM_STACK + LMEM_STACK) { e.ref = MIRO_GPU_CHILD_EMPTY; e.t = 0.f; return e; }
        unsigned long long v = (sp < SMEM_STACK) ? smem[sp * TRACE_BLOCK] : lmem[sp - SMEM_STACK];
        e.ref = (int32_t)(uint32_t)v; e.t = __uint_as_float((uint32_t)(v >> 32));
        return e;
    }
};

#define MIRO_CSWAP(ta, ca, tb, cb) { if (tb < ta) { float tt = ta; ta = tb; tb = tt; int32_t cc = ca; ca = cb; cb = cc; } }

__device__ __forceinline__ bool ref_is_inner(int32_t ref) { return ref >= 0 && ref != MIRO_GPU_CHILD_EMPTY && ref != STACK_SENTINEL; }

// One ray slot of a persistent warp.
struct Lane {
    RaySpace r;          // current-space ray (world, or object space inside an instance)
    float tmin, time;
    HitRec hit;          // hit.t = current tmax
    int32_t cur;         // node / leaf reference being processed, or MIRO_GPU_CHILD_EMPTY
    int32_t cur_inst;
    uint32_t ray_idx;
    bool done;           // slot is empty
};

// Pops the next candidate that can still beat the current hit; `cur` = MIRO_GPU_CHILD_EMPTY when the stack runs dry.
__device__ __forceinline__ void pop_next(Lane& L, TraversalStack& st) {
    L.cur = MIRO_GPU_CHILD_EMPTY;
    while (st.sp > 0) {
        const StackEntry e = st.pop();
        if (e.ref == STACK_SENTINEL || e.t < L.hit.t) { L.cur = e.ref; break; }
    }
}

// Node step: test the four children of inner node `cur`, continue with the nearest, defer the others (far to near).
template <bool COUNT>
__device__ __forceinline__ void node_step(const DeviceScene& s, Lane& L, TraversalStack& st, uint32_t& n_nodes) {
    const float inf = __int_as_float(0x7f800000);
    const float4* n = s.nodes + (size_t)L.cur * 8;
    const float4 lox = __ldg(n + 0), loy = __ldg(n + 1), loz = __ldg(n + 2);
    const float4 hix = __ldg(n + 3), hiy = __ldg(n + 4), hiz = __ldg(n + 5);
    const int4 ch = __ldg(reinterpret_cast<const int4*>(n + 6));
    if (COUNT) ++n_nodes;
    float tn0, tn1, tn2, tn3;
#define MIRO_SLAB(LX, LY, LZ, HX, HY, HZ, CH, TN) { \
    float ax = (LX - L.r.ox) * L.r.ix, bx = (HX - L.r.ox) * L.r.ix; \
    float ay = (LY - L.r.oy) * L.r.iy, by = (HY - L.r.oy) * L.r.iy; \
    float az = (LZ - L.r.oz) * L.r.iz, bz = (HZ - L.r.oz) * L.r.iz; \
    float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), L.tmin)); \
    float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), L.hit.t)); \
    TN = (tn <= tf && CH != MIRO_GPU_CHILD_EMPTY) ? tn : inf; }
    int32_t c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
    MIRO_SLAB(lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, c0, tn0)
    MIRO_SLAB(lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, c1, tn1)
    MIRO_SLAB(lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, c2, tn2)
    MIRO_SLAB(lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, c3, tn3)
#undef MIRO_SLAB
    MIRO_CSWAP(tn0, c0, tn1, c1) MIRO_CSWAP(tn2, c2, tn3, c3)
    MIRO_CSWAP(tn0, c0, tn2, c2) MIRO_CSWAP(tn1, c1, tn3, c3)
    MIRO_CSWAP(tn1, c1, tn2, c2)
    if (tn3 < inf) st.push(c3, tn3);
    if (tn2 < inf) st.push(c2, tn2);
    if (tn1 < inf) st.push(c1, tn1);
    if (tn0 < inf) L.cur = c0;
    else pop_next(L, st);
}

// Leaf phase.  Returns true when an ANY query has found its occluder.
template <bool ANY, bool COUNT>
__device__ __forceinline__ bool intersect_leaf(const DeviceScene& s, Lane& L, TraversalStack& st, const float4* __restrict__ rays,
                                               uint32_t& n_tris, uint32_t& n_insts) {
    const uint32_t u = (uint32_t)L.cur;
    const uint32_t kind = (u >> 29) & 3u;
    const uint32_t count = ((u >> MIRO_GPU_LEAF_INDEX_BITS) & 7u) + 1u;
    const uint32_t first = u & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u);
    if (kind == MIRO_GPU_KIND_TRI) {
        for (uint32_t i = 0; i < count; ++i) {
            const float4* t = s.tris + (size_t)(first + i) * 3;
            const float4 p0 = __ldg(t), p1 = __ldg(t + 1), p2 = __ldg(t + 2);
            if (COUNT) ++n_tris;
            if (intersect_tri(L.r, L.tmin, L.hit.t, p0, p1, p2, L.hit.t, L.hit.a, L.hit.b)) {
                L.hit.prim = (int32_t)(first + i); L.hit.inst = L.cur_inst;
                if (ANY) return true;
            }
        }
    } else if (kind == MIRO_GPU_KIND_MBTRI) {
        const float w1 = L.time, w0 = 1.0f - L.time;    // src/BVH.cpp:1323-1334
        for (uint32_t i = 0; i < count; ++i) {
            const float4* t = s.mbtris + (size_t)(first + i) * 6;
            const float4 a0 = __ldg(t), a1 = __ldg(t + 1), a2 = __ldg(t + 2);
            const float4 b0 = __ldg(t + 3), b1 = __ldg(t + 4), b2 = __ldg(t + 5);
            if (COUNT) ++n_tris;
            float4 p0, p1, p2;
            p0.x = w1 * b0.x + w0 * a0.x; p0.y = w1 * b0.y + w0 * a0.y; p0.z = w1 * b0.z + w0 * a0.z;
            p1.x = w1 * b1.x + w0 * a1.x; p1.y = w1 * b1.y + w0 * a1.y; p1.z = w1 * b1.z + w0 * a1.z;
            p2.x = w1 * b2.x + w0 * a2.x; p2.y = w1 * b2.y + w0 * a2.y; p2.z = w1 * b2.z + w0 * a2.z;
            if (intersect_tri(L.r, L.tmin, L.hit.t, p0, p1, p2, L.hit.t, L.hit.a, L.hit.b)) {
                L.hit.prim = (int32_t)(s.n_tris + first + i); L.hit.inst = L.cur_inst;
                if (ANY) return true;
            }
        }
    } else {   // MIRO_GPU_KIND_INST: enter the first instance, defer the others
        const float ninf = -__int_as_float(0x7f800000);
        if (count > 1u) st.push(MIRO_GPU_LEAF(MIRO_GPU_KIND_INST, first + 1u, count - 1u), ninf);
        st.push(STACK_SENTINEL, ninf);
        const float4* m = s.insts + (size_t)first * 4;
        const float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
        const int4 meta = __ldg(reinterpret_cast<const int4*>(m + 3));
        if (COUNT) ++n_insts;
        // the world-space ray is not kept in registers: instance entry / exit re-read it (L2-resident, rare)
        const float4 w0 = __ldg(rays + (size_t)L.ray_idx * 3), w1 = __ldg(rays + (size_t)L.ray_idx * 3 + 1);
        const float wox = w0.x, woy = w0.y, woz = w0.z, wdx = w1.x, wdy = w1.y, wdz = w1.z;
        // o' = M^-1 [o 1], d' = M^-1 [d 0]  (src/ProxyObject.cpp:78-79; affine, so w = 1)
        L.r.set(r0.x * wox + r0.y * woy + r0.z * woz + r0.w, r1.x * wox + r1.y * woy + r1.z * woz + r1.w, r2.x * wox + r2.y * woy + r2.z * woz + r2.w,
                r0.x * wdx + r0.y * wdy + r0.z * wdz, r1.x * wdx + r1.y * wdy + r1.z * wdz, r2.x * wdx + r2.y * wdy + r2.z * wdz);
        L.cur_inst = (int32_t)first;
        L.cur = meta.x;
        return false;
    }
    L.cur = MIRO_GPU_CHILD_EMPTY;
    return false;
}

}  // namespace miro
