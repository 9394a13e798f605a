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
//   * 64-byte quantized BVH4 nodes (re-encoded from the 128-byte ABI node at upload) fetched with two 32-byte vector loads
//     through the read-only path (the whole tree lives in the 126 MB L2; hot top levels in L1);
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
#include <string.h>
#include <math.h>
#include "../../include/miro_gpu.h"

namespace miro {

#ifndef MIRO_TRACE_BLOCK
#define MIRO_TRACE_BLOCK 128
#endif
#ifndef MIRO_TRACE_MIN_BLOCKS
#define MIRO_TRACE_MIN_BLOCKS 8
#endif
#ifndef MIRO_SMEM_STACK
#define MIRO_SMEM_STACK 16
#endif
constexpr int TRACE_BLOCK = MIRO_TRACE_BLOCK;            // threads per block of the traversal kernels
constexpr int TRACE_MIN_BLOCKS = MIRO_TRACE_MIN_BLOCKS;  // resident blocks per SM the kernels are compiled for (register budget)
#ifndef MIRO_TRACE_REFILL
#define MIRO_TRACE_REFILL 8
#endif
constexpr int TRACE_REFILL = MIRO_TRACE_REFILL;
static_assert(TRACE_REFILL >= 1 && TRACE_REFILL <= 32, "a warp whose 32 slots are idle must reach the refill block (that is where it leaves)");
// a node round is run when  n_node * DEN >= n_leaf * NUM  (NUM/DEN < 1 favours node rounds: leaf rounds cost more and fill up while waiting)
#ifndef MIRO_NODE_BIAS_NUM
#define MIRO_NODE_BIAS_NUM 1
#endif
#ifndef MIRO_NODE_BIAS_DEN
#define MIRO_NODE_BIAS_DEN 1
#endif
constexpr int TRACE_NODE_BIAS_NUM = MIRO_NODE_BIAS_NUM, TRACE_NODE_BIAS_DEN = MIRO_NODE_BIAS_DEN;
#ifndef MIRO_LEAF_MIN
#define MIRO_LEAF_MIN 0
#endif
constexpr int TRACE_LEAF_MIN = MIRO_LEAF_MIN;   // tuning knob: while some lane waits at a node, a leaf round needs at least this many lanes (0: the ratio alone decides)
   // idle lanes in a warp that trigger a refill from the work counter
// Byte -> float decode of the quantized node bounds.  0: all 24 planes by I2F.U8 (the quarter-rate XU pipe); 1: the 12 far planes by
// one byte permute each (PRMT drops the byte into the mantissa of 1.0f: ALU pipe), the 12 near planes by I2F — splits the decode
// over two pipes; 2: all 24 by PRMT (measured slower than 0 in round 2: the ALU pipe becomes the bound).
#ifndef MIRO_PRMT_PLANES
#define MIRO_PRMT_PLANES 0
#endif
constexpr int SMEM_STACK = MIRO_SMEM_STACK;   // per-thread stack entries kept in shared memory
constexpr int LMEM_STACK = 96 - MIRO_SMEM_STACK;          // overflow entries (local memory, touched only by very deep trees)
constexpr int32_t STACK_SENTINEL = 0x7ffffffe;   // "leave instance" marker

// What the traversal needs to evaluate alpha cut-outs (intersect4's alpha-map test, src/BVH.cpp:1401-1435); all NULL
// when no material has an alpha map (the kernels are then instantiated without the test).
struct AlphaTexture { const float* texels; int32_t width, height, channels, pad; };
struct AlphaData {
    const miro_gpu_prim* prims;
    const float* uvs;
    const miro_gpu_material* materials;
    const AlphaTexture* textures;
};

struct DeviceScene {
    AlphaData alpha;
    const float4* nodes;     // DeviceNode: 4 x float4 per node (the 64-byte compressed form, see compress_node)
    const float4* tris;      // TRI_F4 x float4 per triangle (v0, v1, v2 [, pad]), leaf order
    const float4* mbtris;    // 6 x float4 per motion-blur triangle (96 bytes: three 32-byte loads)
    const float4* insts;     // 4 x float4 per instance
    int32_t root;
    uint32_t n_tris;
    const uint32_t* prim_map;   // device-built trees only: caller's triangle index of device triangle i (NULL otherwise)
};

// 32-byte read-only global load (sm_100: LDG.E.256): one L1 wavefront where two 16-byte loads take two.  The traversal
// kernels are bound by L1 wavefronts (every lane reads its own node), so a 128-byte node is fetched as 4 x 32 bytes.
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

// Device triangle record.  3 (default): the 48-byte ABI triangle (three float4 vertices) as it is, three 16-byte loads.
// 4: padded to 64 bytes and 64-byte aligned — one 32-byte + one 16-byte load out of one 128-byte line and exactly two 32-byte
// sectors (the 48-byte stride straddles a line with every fourth triangle; ncu: L1 sector traffic 1.23 x algorithmic).  Built in
// round 2 and measured against each other: the padded record is 0.6 % slower on the 87 k-triangle scene and 2.2 % slower on
// the 1.74 M-triangle one (a third more footprint for the same data: 237 MB against 187 MB) — the sectors it saves were never
// the limit.  Either way a test reads 48 bytes.
#ifndef MIRO_TRI_F4
#define MIRO_TRI_F4 3
#endif
constexpr int TRI_F4 = MIRO_TRI_F4;
__device__ __forceinline__ void load_tri(const float4* __restrict__ t, float4& p0, float4& p1, float4& p2) {
#if MIRO_TRI_F4 == 4
    ldg256(t, p0, p1);
    p2 = __ldg(t + 2);
#else
    p0 = __ldg(t); p1 = __ldg(t + 1); p2 = __ldg(t + 2);
#endif
}

struct TraceCounters {
    unsigned long long rays_closest, rays_any, nodes, tris, insts;
};

struct HitRec {
    float t, a, b;
    int32_t prim, inst;
};

// branch-free float select (the compiler turns chains of ternaries on the dominant axis into divergent branches)
__device__ __forceinline__ float fsel(bool c, float a, float b) {
    float r;
    asm("{ .reg .pred p; setp.ne.b32 p, %3, 0; selp.f32 %0, %1, %2, p; }" : "=f"(r) : "f"(a), "f"(b), "r"((int)c));
    return r;
}

__device__ __forceinline__ float safe_rcp_dir(float d) {
    // The reciprocal direction only feeds the (conservative) slab test: one MUFU.RCP (<= 1 ulp) instead of the correctly rounded
    // reciprocal with its slow path, branch-free.  Components below 1e-12 in magnitude (0 included) get the reference's +-MIRO_TMAX
    // (src/Ray.h:79-90); the slab test's widening (node_step) covers the reciprocal's error.
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return fsel(fabsf(d) < 1.0e-12f, copysignf(MIRO_GPU_TMAX, d), r);
}


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
    // The watertight test works in a sheared space whose z axis is the ray (Woop, Benthin, Wald 2013): with kz the dominant axis
    // of the direction, a translated vertex v maps to  x' = v[kz+1] - Sb v[kz],  y' = v[kz+2] - Sc v[kz]  (indices mod 3),
    // Sb = d[kz+1] / d[kz], Sc = d[kz+2] / d[kz].  The paper also swaps x' and y' when d[kz] < 0 to keep the winding for back-face
    // culling; this test is two-sided (only the signs' agreement matters, and they are exact), so the swap is dropped.
    float Sb, Sc;
    int kz;

    __device__ __forceinline__ void set(float ox_, float oy_, float oz_, float dx_, float dy_, float dz_) {
        ox = ox_; oy = oy_; oz = oz_; dx = dx_; dy = dy_; dz = dz_;
        ix = safe_rcp_dir(dx); iy = safe_rcp_dir(dy); iz = safe_rcp_dir(dz);
        const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
        const bool z0 = (ax > ay) && (ax > az);                 // kz == 0
        const bool z1 = !(ax > ay) && (ay > az);                // kz == 1; otherwise kz == 2
        kz = z0 ? 0 : (z1 ? 1 : 2);
        const float rz = __frcp_rn(fsel(z0, dx, fsel(z1, dy, dz)));
        Sb = fsel(z0, dy, fsel(z1, dz, dx)) * rz;
        Sc = fsel(z0, dz, fsel(z1, dx, dy)) * rz;
    }
};

// The three translated vertices of a triangle rotated into the ray's axis order: (a, b, c) = (v[kz], v[kz+1], v[kz+2]).  One
// block of predicated selects (2 setp + 18 selp): as C++ ternaries this compiled to ~48 instructions of divergent branches per
// triangle test, 15 % of all instructions the traversal kernel issued.
__device__ __forceinline__ void rotate_to_ray_axes(int kz, float& ax, float& ay, float& az, float& bx, float& by, float& bz, float& cx, float& cy, float& cz) {
    float a0, b0, c0, a1, b1, c1, a2, b2, c2;
    asm("{ .reg .pred p0, p1; .reg .f32 t;\n\t"
        "setp.eq.s32 p0, %18, 0; setp.eq.s32 p1, %18, 1;\n\t"
        "selp.f32 t, %10, %11, p1; selp.f32 %0, %9, t, p0; selp.f32 t, %11, %9, p1; selp.f32 %1, %10, t, p0; selp.f32 t, %9, %10, p1; selp.f32 %2, %11, t, p0;\n\t"
        "selp.f32 t, %13, %14, p1; selp.f32 %3, %12, t, p0; selp.f32 t, %14, %12, p1; selp.f32 %4, %13, t, p0; selp.f32 t, %12, %13, p1; selp.f32 %5, %14, t, p0;\n\t"
        "selp.f32 t, %16, %17, p1; selp.f32 %6, %15, t, p0; selp.f32 t, %17, %15, p1; selp.f32 %7, %16, t, p0; selp.f32 t, %15, %16, p1; selp.f32 %8, %17, t, p0; }"
        : "=f"(a0), "=f"(b0), "=f"(c0), "=f"(a1), "=f"(b1), "=f"(c1), "=f"(a2), "=f"(b2), "=f"(c2)
        : "f"(ax), "f"(ay), "f"(az), "f"(bx), "f"(by), "f"(bz), "f"(cx), "f"(cy), "f"(cz), "r"(kz));
    ax = a0; ay = b0; az = c0; bx = a1; by = b1; bz = c1; cx = a2; cy = b2; cz = c2;
}

// Distance and barycentrics of a crossing the watertight test has accepted, evaluated with the reference's Moller-Trumbore
// expressions in the reference's operation order (src/BVH.cpp:1343-1369; SoADot = x*x' + (y*y' + z*z'), no FMA contraction), so
// t, a, b agree with it to rounding instead of differing by the conditioning of two different algorithms on sliver triangles.
// a, b are clamped to the triangle (inside by the edge test: only the rounding residue is cut).  False: degenerate (det == 0).
__device__ __forceinline__ bool moller_trumbore_reference(const RaySpace& r, float4 p0, float4 p1, float4 p2, float& t_out, float& a_out, float& b_out) {
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
    t_out = __fmul_rn(inv, __fadd_rn(__fmul_rn(e1x, qx), __fadd_rn(__fmul_rn(e1y, qy), __fmul_rn(e1z, qz))));
    const float a = __fmul_rn(inv, __fadd_rn(__fmul_rn(tx, px), __fadd_rn(__fmul_rn(ty, py), __fmul_rn(tz, pz))));
    const float b = __fmul_rn(inv, __fadd_rn(__fmul_rn(r.dx, qx), __fadd_rn(__fmul_rn(r.dy, qy), __fmul_rn(r.dz, qz))));
    a_out = fminf(fmaxf(a, 0.0f), 1.0f);
    b_out = fminf(fmaxf(b, 0.0f), 1.0f - a_out);
    return true;
}

// Watertight two-sided ray/triangle test.  Returns true and updates (t,a,b) when tmin <= t < tmax.
// T_ONLY (any-hit queries without alpha maps): the caller only asks WHETHER the crossing lies in [tmin, tmax), so the distance is
// taken from the edge functions themselves and the reference-order Moller-Trumbore block (53 instructions run by 2 of 32 lanes,
// 8 % of a shadow launch) is skipped; a, b are not produced.
template <bool T_ONLY = false>
__device__ __forceinline__ bool intersect_tri(const RaySpace& r, float tmin, float tmax,
                                              float4 p0, float4 p1, float4 p2, float& t_out, float& a_out, float& b_out) {
    float Aa = p0.x - r.ox, Ab = p0.y - r.oy, Ac = p0.z - r.oz;
    float Ba = p1.x - r.ox, Bb = p1.y - r.oy, Bc = p1.z - r.oz;
    float Ca = p2.x - r.ox, Cb = p2.y - r.oy, Cc = p2.z - r.oz;
    rotate_to_ray_axes(r.kz, Aa, Ab, Ac, Ba, Bb, Bc, Ca, Cb, Cc);      // (a, b, c) = components along kz, kz+1, kz+2
    const float Ax = __fmaf_rn(-r.Sb, Aa, Ab), Ay = __fmaf_rn(-r.Sc, Aa, Ac);
    const float Bx = __fmaf_rn(-r.Sb, Ba, Bb), By = __fmaf_rn(-r.Sc, Ba, Bc);
    const float Cx = __fmaf_rn(-r.Sb, Ca, Cb), Cy = __fmaf_rn(-r.Sc, Ca, Cc);
    const float U = diff_of_products(Cx, By, Cy, Bx);
    const float V = diff_of_products(Ax, Cy, Ay, Cx);
    const float W = diff_of_products(Bx, Ay, By, Ax);
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    if (T_ONLY) {
        // t = (U A + V B + W C)[kz] / (det d[kz]): the barycentric mean of the vertices' distances along the dominant axis
        const float dkz = fsel(r.kz == 0, r.dx, fsel(r.kz == 1, r.dy, r.dz));
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(det * dkz));
        const float t = __fmaf_rn(U, Aa, __fmaf_rn(V, Ba, W * Ca)) * inv;
        if (!(t >= tmin && t < tmax)) return false;
        t_out = t; a_out = 0.f; b_out = 0.f;
        return true;
    }
    float t, a, b;
    if (!moller_trumbore_reference(r, p0, p1, p2, t, a, b)) return false;
    if (!(t >= tmin && t < tmax)) return false;
    t_out = t; a_out = a; b_out = b;
    return true;
}

// order-preserving map float -> int (signed integer compare == float compare, any sign)
__device__ __forceinline__ int float_key(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }

struct StackEntry { int32_t ref; int key; };     // key: float_key(entry distance), low bits cleared (never later than the truth)

// Per-ray traversal stack.  Entries are {reference, entry distance}; the first NSMEM of them live in SHARED memory with a
// fixed byte STRIDE between consecutive entries of one stack — [entry][lane] for the persistent-warp kernel, [entry][slot] for
// the pool kernel: conflict free when the lanes of a warp address distinct columns —, addressed with explicit 32-bit shared
// addresses (ld/st.shared — a generic pointer here made the compiler emit generic LD/ST plus a stack pointer in local memory).
// Deeper entries (very deep trees only; depth is validated at upload) overflow into `overflow` (local memory of the lane, or a
// per-slot scratch in global memory for the pool kernel, whose rays change lanes) behind one rarely taken branch.
template <uint32_t STRIDE_, int NSMEM_>
struct TraversalStackT {
    static constexpr uint32_t STRIDE = STRIDE_;
    static constexpr int NSMEM = NSMEM_;
    uint32_t base;                       // shared address of entry 0 of this stack
    unsigned long long* overflow;        // `cap` further entries
    int cap;
    int sp;
    __device__ __forceinline__ void init(unsigned long long* smem_entry0, unsigned long long* ovf, int ovf_cap) {
        base = (uint32_t)__cvta_generic_to_shared(smem_entry0); overflow = ovf; cap = ovf_cap; sp = 0;
    }
    __device__ __forceinline__ void store(int slot, int32_t ref, int key) {
        if (slot < NSMEM)
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" :: "r"(base + (uint32_t)slot * STRIDE), "r"(ref), "r"(key) : "memory");
        else if (slot - NSMEM < cap)
            overflow[slot - NSMEM] = ((unsigned long long)(uint32_t)key << 32) | (uint32_t)ref;
    }
    __device__ __forceinline__ void push(int32_t ref, int key) { store(sp, ref, key); ++sp; }
    __device__ __forceinline__ StackEntry pop() {
        --sp;
        StackEntry e;
        if (sp < NSMEM)
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(e.ref), "=r"(e.key) : "r"(base + (uint32_t)sp * STRIDE) : "memory");
        else if (sp - NSMEM < cap) { const unsigned long long v = overflow[sp - NSMEM]; e.ref = (int32_t)(uint32_t)v; e.key = (int)(uint32_t)(v >> 32); }
        else { e.ref = MIRO_GPU_CHILD_EMPTY; e.key = 0x7fffffff; }
        return e;
    }
};
using TraversalStack = TraversalStackT<TRACE_BLOCK * 8u, SMEM_STACK>;      // the persistent-warp kernel: [entry][lane of the block]

// four-way select by a 2-bit slot index, forced to predicated selects (the compiler turns the ternary chain into branches)
__device__ __forceinline__ int sel4(int sl, int a, int b, int c, int d) {
    int r;
    asm("{ .reg .pred p0, p1; .reg .b32 lo, hi;\n\t"
        "and.b32 lo, %1, 1; setp.ne.b32 p0, lo, 0; and.b32 hi, %1, 2; setp.ne.b32 p1, hi, 0;\n\t"
        "selp.b32 lo, %3, %2, p0; selp.b32 hi, %5, %4, p0; selp.b32 %0, hi, lo, p1; }"
        : "=r"(r) : "r"(sl), "r"(a), "r"(b), "r"(c), "r"(d));
    return r;
}
// predicated 8-byte shared store
__device__ __forceinline__ void sts_if(int cond, uint32_t addr, int ref, int key) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %3, 0; @q st.shared.v2.b32 [%0], {%1, %2}; }" :: "r"(addr), "r"(ref), "r"(key), "r"(cond) : "memory");
}

// inner-node references are 0 .. 0x7ffffffd; the instance-exit marker, EMPTY and the leaf references (bit 31) lie above
__device__ __forceinline__ bool ref_is_inner(int32_t ref) { return (uint32_t)ref < (uint32_t)STACK_SENTINEL; }

// One ray slot of a persistent warp.
struct Lane {
    RaySpace r;          // current-space ray (world, or object space inside an instance)
    float tmin, time;
    HitRec hit;          // hit.t = current tmax
    int32_t cur;         // node / leaf reference being processed, or MIRO_GPU_CHILD_EMPTY
    int32_t cur_inst;
    uint32_t ray_idx;
    __device__ __forceinline__ void set_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
        r.set(ox, oy, oz, dx, dy, dz);
    }
    __device__ __forceinline__ const RaySpace& ray_space() const { return r; }      // (the flat kernel's lane keeps part of it in shared memory, trace_flat.cuh)
};

// Pops the next candidate that can still beat the current hit; `cur` = MIRO_GPU_CHILD_EMPTY when the stack runs dry.
template <class LN, class ST>
__device__ __forceinline__ void pop_next(LN& L, ST& st) {
    L.cur = MIRO_GPU_CHILD_EMPTY;
    const int limit = float_key(L.hit.t);
    while (st.sp > 0) {
        const StackEntry e = st.pop();
        if (e.key < limit) { L.cur = e.ref; break; }       // instance markers carry the smallest key
    }
}

// Device node: the 128-byte ABI node (include/miro_gpu.h) is re-encoded at upload into 64 bytes — the traversal kernels
// are bound by L1 data-pipe wavefronts, which for divergent loads scale with the BYTES each lane fetches:
//   word 0..2   p = grid origin, a little below the min corner of the union of the children's boxes (float)
//   word 3      biased power-of-two exponents of the per-axis grid step: ex | ey << 8 | ez << 16 (step = 2^(e-127))
//   word 4..7   child references (as in miro_gpu_node)
//   word 8..10  lower bounds of the 4 children on x, y, z: one byte per child, grid units, rounded DOWN
//   word 11..13 upper bounds, rounded UP            word 14..15 unused
// A child's box only ever grows (by less than 1 + 2/32 grid steps = extent/240 per side), so no hit is lost; the
// uncompressed node would only have culled a few more candidates.  Every quantized plane keeps a MARGIN of >= 1/32 grid step
// to the true box (NODE_GRID_MARGIN): that margin is what absorbs the absolute rounding error of the slab arithmetic in
// node_step (see there), so the boxes stay conservative for every ray, not only up to rounding.
struct DeviceNode { uint32_t w[16]; };
static_assert(sizeof(DeviceNode) == 64, "DeviceNode layout");
constexpr double NODE_GRID_MARGIN = 1.0 / 32.0;

__host__ __device__ inline DeviceNode compress_node(const miro_gpu_node& n) {
    DeviceNode o;
    for (int i = 0; i < 16; ++i) o.w[i] = 0;
    const float* lo[3] = {n.lo_x, n.lo_y, n.lo_z};
    const float* hi[3] = {n.hi_x, n.hi_y, n.hi_z};
    for (int k = 0; k < 3; ++k) {
        float pmin = 3.0e38f, pmax = -3.0e38f;
        for (int c = 0; c < 4; ++c) if (n.child[c] != MIRO_GPU_CHILD_EMPTY) { pmin = fminf(pmin, lo[k][c]); pmax = fmaxf(pmax, hi[k][c]); }
        if (!(pmin <= pmax)) { pmin = 0.f; pmax = 0.f; }
        // grid: origin p0 (a float at or below pmin - margin), step = the smallest power of two with
        // p0 + 255 * step >= pmax + margin, margin = step / 32.  Evaluated in double (exact for these operands).
        float p0 = pmin;
        int e = 1;
        double step = 0.0;
        for (int iter = 0; iter < 8; ++iter) {
            const double ext = (double)pmax - (double)p0;
            e = 1;
            if (ext > 0.0) { int ex; frexp(ext / 255.0, &ex); e = ex + 126; e = e < 1 ? 1 : (e > 254 ? 254 : e); }    // 2^(ex-1) <= ext/255 < 2^ex
            while (e < 254 && (255.0 - NODE_GRID_MARGIN) * ldexp(1.0, e - 127) < ext) ++e;
            step = ldexp(1.0, e - 127);
            if ((double)pmin - (double)p0 >= step * NODE_GRID_MARGIN) break;
            // lower the origin by ~step/16 (at least one float below) and size the grid again
            const float want = (float)((double)pmin - step * (1.0 / 16.0));
            p0 = want < p0 ? want : nextafterf(p0, -3.0e38f);
        }
        uint32_t qlo = 0, qhi = 0;
        for (int c = 0; c < 4; ++c) {
            uint32_t a = 255u, b = 0u;                     // empty slot: never consulted (its reference is EMPTY)
            if (n.child[c] != MIRO_GPU_CHILD_EMPTY) {
                double fa = floor(((double)lo[k][c] - (double)p0) / step - NODE_GRID_MARGIN), fb = ceil(((double)hi[k][c] - (double)p0) / step + NODE_GRID_MARGIN);
                fa = fa < 0.0 ? 0.0 : (fa > 255.0 ? 255.0 : fa); fb = fb < 0.0 ? 0.0 : (fb > 255.0 ? 255.0 : fb);
                a = (uint32_t)fa; b = (uint32_t)fb;
            }
            qlo |= a << (8 * c); qhi |= b << (8 * c);
        }
        union { float f; uint32_t u; } cv; cv.f = p0; o.w[k] = cv.u;
        o.w[3] |= (uint32_t)e << (8 * k);
        o.w[8 + k] = qlo; o.w[11 + k] = qhi;
    }
    for (int c = 0; c < 4; ++c) o.w[4 + c] = (uint32_t)n.child[c];
    return o;
}

// Node step: test the four children of inner node `cur`, continue with the nearest, defer the others (far to near).
// Straight-line code: two 32-byte loads, byte -> float conversions (I2F.U8: the XU pipe, which nothing else here uses — decoding
// the bytes with PRMT into a float's mantissa instead was measured in round 2 and is slower, because it moves 24 instructions per
// step onto the ALU pipe, the busiest one at ~60 %), slab tests in FMA form (grid unit * step/d + (p - o)/d), a 5-comparator
// sorting network on integer keys (entry distance with the child slot in its two low mantissa bits — truncation only makes an
// entry look nearer, which is conservative for culling), predicated pushes.
//
// Error budget (what keeps the test conservative): the relative errors of t (reciprocal 1 ulp, p - o, products, FMA: < 8 ulp
// on the near and the far bound together) are covered by widening the far bound by MIRO_SLAB_WIDEN; what is absolute — the
// rounding of (p - o)/d when the plane is much nearer than the grid origin, <= 2^-23 x 255 grid steps of t — is covered by the
// >= 1/32 grid step every quantized plane keeps to the true box (compress_node, NODE_GRID_MARGIN).
// The 64 bytes of a device node as the two 32-byte loads deliver them.  node_step = node_fetch + node_test.  (The split served a
// round-2 variant of the pool kernel that requested a node AND a leaf's first triangle before using either — two steps per round,
// measured 25 % slower than one: serving every waiting slot at once empties the rounds, the leaf part ran at ~20 lanes.)
struct NodeData { float4 h0, chf, q0, q1; };
__device__ __forceinline__ void node_fetch(const DeviceScene& s, int32_t cur, NodeData& d) {
    const float4* n = s.nodes + (size_t)cur * 4;
    ldg256(n + 0, d.h0, d.chf); ldg256(n + 2, d.q0, d.q1);
}

template <class LN, class ST>
__device__ __forceinline__ void node_test(const NodeData& d, LN& L, ST& st) {
    const float4 h0 = d.h0, chf = d.chf, q0 = d.q0, q1 = d.q1;
    const uint32_t ex = __float_as_uint(h0.w);
    // per axis: t(q) = q * (step / d) + (p - o) / d   (the subtraction first, so the error of the second term is relative to it)
    const float ax = __uint_as_float((ex & 0xffu) << 23) * L.r.ix, bx = (h0.x - L.r.ox) * L.r.ix;
    const float ay = __uint_as_float((ex & 0xff00u) << 15) * L.r.iy, by = (h0.y - L.r.oy) * L.r.iy;
    const float az = __uint_as_float((ex & 0xff0000u) << 7) * L.r.iz, bz = (h0.z - L.r.oz) * L.r.iz;
    // near / far planes by the sign of the ray direction (the same for all four children): whole-word selects, so the
    // per-child test needs no min/max of plane pairs
    const uint32_t lx = __float_as_uint(q0.x), ly = __float_as_uint(q0.y), lz = __float_as_uint(q0.z);
    const uint32_t hx = __float_as_uint(q0.w), hy = __float_as_uint(q1.x), hz = __float_as_uint(q1.y);
    const bool ngx = L.r.ix < 0.f, ngy = L.r.iy < 0.f, ngz = L.r.iz < 0.f;
    const uint32_t nx = ngx ? hx : lx, fx = ngx ? lx : hx;
    const uint32_t ny = ngy ? hy : ly, fy = ngy ? ly : hy;
    const uint32_t nz = ngz ? hz : lz, fz = ngz ? lz : hz;
    const int INF_KEY = 0x7fffffff;
    const float tmax = L.hit.t;
    int k0, k1, k2, k3;
#define MIRO_SLAB_WIDEN 1.6e-6f      /* relative: ~13 ulp, the near and the far bound's rounding together (see above) */
#define MIRO_BYTE(W, C) ((float)(((W) >> (8 * (C))) & 0xffu))
#if MIRO_PRMT_PLANES
    // PRMT decode (see MIRO_PRMT_PLANES): f = 1 + q 2^-15, t = f A + B with A = 2^15 a, B = b - A
    const float Ax = ax * 32768.0f, Bx = __fsub_rn(bx, Ax), Ay = ay * 32768.0f, By = __fsub_rn(by, Ay), Az = az * 32768.0f, Bz = __fsub_rn(bz, Az);
#define MIRO_BYTE_P(W, C) __uint_as_float(__byte_perm((W), 0x3f800000u, 0x7604u | ((C) << 4)))
#define MIRO_FAR(W, C, K) __fmaf_rn(MIRO_BYTE_P(W, C), A##K, B##K)
#else
#define MIRO_FAR(W, C, K) __fmaf_rn(MIRO_BYTE(W, C), a##K, b##K)
#endif
#if MIRO_PRMT_PLANES == 2
#define MIRO_NEAR(W, C, K) __fmaf_rn(MIRO_BYTE_P(W, C), A##K, B##K)
#else
#define MIRO_NEAR(W, C, K) __fmaf_rn(MIRO_BYTE(W, C), a##K, b##K)
#endif
#define MIRO_SLAB(C, CH, KEY) { \
    const float tn = fmaxf(fmaxf(MIRO_NEAR(nx, C, x), MIRO_NEAR(ny, C, y)), fmaxf(MIRO_NEAR(nz, C, z), L.tmin)); \
    const float tf0 = fminf(fminf(MIRO_FAR(fx, C, x), MIRO_FAR(fy, C, y)), fminf(MIRO_FAR(fz, C, z), tmax)); \
    const float tf = __fmaf_rn(fabsf(tf0), MIRO_SLAB_WIDEN, tf0); \
    KEY = (tn <= tf && __float_as_int(CH) != MIRO_GPU_CHILD_EMPTY) ? ((float_key(tn) & ~3) | C) : INF_KEY; }
    MIRO_SLAB(0, chf.x, k0) MIRO_SLAB(1, chf.y, k1) MIRO_SLAB(2, chf.z, k2) MIRO_SLAB(3, chf.w, k3)
#undef MIRO_NEAR
#undef MIRO_FAR
#undef MIRO_SLAB
#undef MIRO_BYTE
#define MIRO_KSWAP(a, b) { const int lo_ = min(a, b), hi_ = max(a, b); a = lo_; b = hi_; }
    MIRO_KSWAP(k0, k1) MIRO_KSWAP(k2, k3) MIRO_KSWAP(k0, k2) MIRO_KSWAP(k1, k3) MIRO_KSWAP(k1, k2)
#undef MIRO_KSWAP
    const int ch0 = __float_as_int(chf.x), ch1 = __float_as_int(chf.y), ch2 = __float_as_int(chf.z), ch3 = __float_as_int(chf.w);
    auto child_of = [&](int key) { return sel4(key, ch0, ch1, ch2, ch3); };
    const int hits = (k0 != INF_KEY) + (k1 != INF_KEY) + (k2 != INF_KEY) + (k3 != INF_KEY);
    // deferred children go on the stack far to near: k3 (if hit) lowest, k1 on top
    const int sp = st.sp;
    if (sp + 3 <= ST::NSMEM) {          // the usual case: three predicated shared stores, no branches
        const uint32_t a0 = st.base + (uint32_t)sp * ST::STRIDE;
        sts_if(hits > 3, a0, child_of(k3), k3 & ~3);
        sts_if(hits > 2, a0 + (uint32_t)(hits - 3) * ST::STRIDE, child_of(k2), k2 & ~3);
        sts_if(hits > 1, a0 + (uint32_t)(hits - 2) * ST::STRIDE, child_of(k1), k1 & ~3);
    } else {
        if (hits > 3) st.store(sp, child_of(k3), k3 & ~3);
        if (hits > 2) st.store(sp + hits - 3, child_of(k2), k2 & ~3);
        if (hits > 1) st.store(sp + hits - 2, child_of(k1), k1 & ~3);
    }
    st.sp = sp + max(hits - 1, 0);
    if (hits > 0) L.cur = child_of(k0);
    else pop_next(L, st);
}

template <bool COUNT, class LN, class ST>
__device__ __forceinline__ void node_step(const DeviceScene& s, LN& L, ST& st, uint32_t& n_nodes) {
    NodeData d;
    node_fetch(s, L.cur, d);
    if (COUNT) ++n_nodes;
    node_test(d, L, st);
}

// Texture::getLookupAlpha (src/Texture.cpp:12-41) for the hit (prim, a, b): the bilinearly filtered alpha channel at the
// interpolated uv; 1 when the primitive's material has no alpha map (or the map has no alpha channel, Texture.cpp:109-111).
__device__ __forceinline__ float hit_alpha(const AlphaData& A, uint32_t prim, float a, float b) {
    const miro_gpu_prim* pr = A.prims + prim;
    const int32_t am = A.materials[__ldg(&pr->material)].alpha_map;
    if (am < 0) return 1.0f;
    const AlphaTexture t = A.textures[am];
    if (t.channels != 4) return 1.0f;
    float u = a, v = b;
    const uint32_t i0 = __ldg(&pr->uv[0]);
    if (i0 != 0xffffffffu) {
        const float c = 1.0f - a - b;
        const float* t0 = A.uvs + (size_t)i0 * 2; const float* t1 = A.uvs + (size_t)__ldg(&pr->uv[1]) * 2; const float* t2 = A.uvs + (size_t)__ldg(&pr->uv[2]) * 2;
        u = __ldg(t0) * c + __ldg(t1) * a + __ldg(t2) * b;
        v = __ldg(t0 + 1) * c + __ldg(t1 + 1) * a + __ldg(t2 + 1) * b;
    }
    u = u - float(int(u)); v = v - float(int(v));
    if (u < 0.0f) u = u + 1.0f;
    if (v < 0.0f) v = v + 1.0f;
    v = 1.0f - v;
    const float px = u * t.width, py = v * t.height;
    const float x1 = floorf(px), y1 = floorf(py), dx = px - x1, dy = py - y1;
    auto texel = [&](int x, int y) { x = x % t.width; y = y % t.height; return __ldg(t.texels + ((size_t)y * t.width + x) * 4 + 3); };
    const float q1 = texel((int)x1, (int)y1) * (1.0f - dx) + texel((int)x1 + 1, (int)y1) * dx;
    const float q2 = texel((int)x1, (int)y1 + 1) * (1.0f - dx) + texel((int)x1 + 1, (int)y1 + 1) * dx;
    return q1 * (1.0f - dy) + q2 * dy;
}

// Instance leaf {first, count}: the lane enters instance `first` with the world-space ray (wo, wd) moved into its object space and
// defers the others; a marker on the stack restores the world-space ray when the instance's sub-tree is exhausted.
template <class LN, class ST>
__device__ __forceinline__ void enter_instance(const DeviceScene& s, LN& L, ST& st, uint32_t first, uint32_t count,
                                               float wox, float woy, float woz, float wdx, float wdy, float wdz) {
    const int ninf = (int)0x80000000;
    if (count > 1u) st.push(MIRO_GPU_LEAF(MIRO_GPU_KIND_INST, first + 1u, count - 1u), ninf);
    st.push(STACK_SENTINEL, ninf);
    const float4* m = s.insts + (size_t)first * 4;
    const float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
    const int4 meta = __ldg(reinterpret_cast<const int4*>(m + 3));
    // o' = (M^-1 [o 1]) * wRecip, d' = M^-1 [d 0]  (src/ProxyObject.cpp:78-79) in the REFERENCE's rounding, so the object-space ray —
    // and with it t, a, b of an instanced hit — is the reference's to the last bit instead of differing by ulp(|o|):
    //   origin    dpps over [o 1] (src/Matrix4x4.h:728-733): the four products rounded, summed as (p0 + p1) + (p2 + p3), then
    //             multiplied by recipps(w), w = 1 for an affine matrix (miro_gpu_instance::w_recip, evaluated by the host);
    //   direction scalar code (src/Matrix4x4.h:699-701): (m0 dx + m1 dy) + m2 dz, no contraction.
    const float wr = meta.z == 0 ? 1.0f : __int_as_float(meta.z);
#define MIRO_ROW_POINT(R) __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(R.x, wox), __fmul_rn(R.y, woy)), __fadd_rn(__fmul_rn(R.z, woz), R.w)), wr)
#define MIRO_ROW_VECTOR(R) __fadd_rn(__fadd_rn(__fmul_rn(R.x, wdx), __fmul_rn(R.y, wdy)), __fmul_rn(R.z, wdz))
    L.set_ray(MIRO_ROW_POINT(r0), MIRO_ROW_POINT(r1), MIRO_ROW_POINT(r2), MIRO_ROW_VECTOR(r0), MIRO_ROW_VECTOR(r1), MIRO_ROW_VECTOR(r2));
#undef MIRO_ROW_POINT
#undef MIRO_ROW_VECTOR
    L.cur_inst = (int32_t)first;
    L.cur = meta.x;
}

// Leaf phase.  Returns true when an ANY query has found its occluder.
template <bool ANY, bool COUNT, bool ALPHA, class LN, class ST>
__device__ __forceinline__ bool intersect_leaf(const DeviceScene& s, LN& L, ST& st, const float4* __restrict__ rays, const uint32_t ray_f4,
                                               uint32_t& n_tris, uint32_t& n_insts) {
    const uint32_t u = (uint32_t)L.cur;
    const uint32_t kind = (u >> 29) & 3u;
    const uint32_t count = ((u >> MIRO_GPU_LEAF_INDEX_BITS) & 7u) + 1u;
    const uint32_t first = u & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u);
    if (kind == MIRO_GPU_KIND_TRI) {
        const auto& rs = L.ray_space();
        for (uint32_t i = 0; i < count; ++i) {
            float4 p0, p1, p2;
            load_tri(s.tris + (size_t)(first + i) * TRI_F4, p0, p1, p2);
            if (COUNT) ++n_tris;
            float ht, ha, hb;
            if (intersect_tri<ANY && !ALPHA>(rs, L.tmin, L.hit.t, p0, p1, p2, ht, ha, hb) && (!ALPHA || hit_alpha(s.alpha, first + i, ha, hb) >= 0.5f)) {
                L.hit.t = ht; L.hit.a = ha; L.hit.b = hb;
                L.hit.prim = (int32_t)(first + i); L.hit.inst = L.cur_inst;
                if (ANY) return true;
            }
        }
    } else if (kind == MIRO_GPU_KIND_MBTRI) {
        const float w1 = L.time, w0 = __fsub_rn(1.0f, L.time);    // src/BVH.cpp:1320-1321
        const auto& rs = L.ray_space();
        for (uint32_t i = 0; i < count; ++i) {
            const float4* t = s.mbtris + (size_t)(first + i) * 6;
            float4 a0, a1, a2, b0, b1, b2;
            ldg256(t, a0, a1); ldg256(t + 2, a2, b0); ldg256(t + 4, b1, b2);
            if (COUNT) ++n_tris;
            float4 p0, p1, p2;
            // time * pose2 + (1 - time) * pose1 with both products rounded (src/BVH.cpp:1327-1335 is scalar code without contraction)
#define MIRO_LERP(B, A) __fadd_rn(__fmul_rn(w1, B), __fmul_rn(w0, A))
            p0.x = MIRO_LERP(b0.x, a0.x); p0.y = MIRO_LERP(b0.y, a0.y); p0.z = MIRO_LERP(b0.z, a0.z);
            p1.x = MIRO_LERP(b1.x, a1.x); p1.y = MIRO_LERP(b1.y, a1.y); p1.z = MIRO_LERP(b1.z, a1.z);
            p2.x = MIRO_LERP(b2.x, a2.x); p2.y = MIRO_LERP(b2.y, a2.y); p2.z = MIRO_LERP(b2.z, a2.z);
#undef MIRO_LERP
            float ht, ha, hb;
            if (intersect_tri<ANY && !ALPHA>(rs, L.tmin, L.hit.t, p0, p1, p2, ht, ha, hb) && (!ALPHA || hit_alpha(s.alpha, s.n_tris + first + i, ha, hb) >= 0.5f)) {
                L.hit.t = ht; L.hit.a = ha; L.hit.b = hb;
                L.hit.prim = (int32_t)(s.n_tris + first + i); L.hit.inst = L.cur_inst;
                if (ANY) return true;
            }
        }
    } else {   // MIRO_GPU_KIND_INST: enter the first instance, defer the others
        // the world-space ray is not kept in registers: instance entry / exit re-read it (L2-resident, rare)
        const float4 w0 = __ldg(rays + (size_t)L.ray_idx * ray_f4), w1 = __ldg(rays + (size_t)L.ray_idx * ray_f4 + 1);
        if (COUNT) ++n_insts;
        enter_instance(s, L, st, first, count, w0.x, w0.y, w0.z, w1.x, w1.y, w1.z);
        return false;
    }
    L.cur = MIRO_GPU_CHILD_EMPTY;
    return false;
}

// One thread walks one ray to its closest hit, without warp cooperation — for the renderer's rare in-kernel queries (the "full"
// shadow method re-traces from every hit point, render.cu).  (wo, wd) is the world-space ray; L.tmin / L.time / L.hit.t (= tMax)
// are set by the caller.  Same node step, leaf step and instance handling as the persistent-warp kernel.
template <bool ALPHA>
__device__ inline void trace_closest_thread(const DeviceScene& s, Lane& L, TraversalStack& st, float wox, float woy, float woz, float wdx, float wdy, float wdz) {
    uint32_t unused = 0;
    L.set_ray(wox, woy, woz, wdx, wdy, wdz);
    L.hit.a = L.hit.b = 0.f; L.hit.prim = -1; L.hit.inst = -1;
    L.cur = s.root; L.cur_inst = -1; st.sp = 0;
    while (L.cur != MIRO_GPU_CHILD_EMPTY) {
        if (ref_is_inner(L.cur)) { node_step<false>(s, L, st, unused); continue; }
        if (L.cur == STACK_SENTINEL) { L.set_ray(wox, woy, woz, wdx, wdy, wdz); L.cur_inst = -1; L.cur = MIRO_GPU_CHILD_EMPTY; }
        else {
            const uint32_t u = (uint32_t)L.cur;
            if (((u >> 29) & 3u) == MIRO_GPU_KIND_INST)
                enter_instance(s, L, st, u & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u), ((u >> MIRO_GPU_LEAF_INDEX_BITS) & 7u) + 1u, wox, woy, woz, wdx, wdy, wdz);
            else intersect_leaf<false, false, ALPHA>(s, L, st, nullptr, 0u, unused, unused);
        }
        if (L.cur == MIRO_GPU_CHILD_EMPTY) pop_next(L, st);
    }
}

}  // namespace miro
