/* oracle/miro_oracle.c — CPU restatement of the reference's ray-casting path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this file; the product (libmiro_gpu.so) never does.
 *
 * Parity status: PINNED against the reference itself — tests/test_oracle_vs_reference.py compares
 * this restatement with hit dumps produced by the unmodified reference built headless here
 * (oracle/_ref/miro_ref, recipe oracle/build_ref.sh) and committed as tests/golden fixtures.  The
 * reference ships no tests / golden vectors of its own.
 *
 * It restates, in plain scalar C operating on the flattened scene of include/miro_gpu.h:
 *   BVH::intersect, QBVH branch        reference src/BVH.cpp:1128-1178  (stack DFS, children 0..3,
 *                                      leaves intersected immediately, inner children pushed in order)
 *   QBVH_Node::intersect               reference src/BVH.cpp:391-414    (slab test, tmin<=tmax)
 *   intersect4                         reference src/BVH.cpp:1298-1459  (Moller-Trumbore, the 4 lanes of a
 *                                      packet all tested against the tMax at packet entry, nearest lane wins,
 *                                      lowest lane index on ties)
 *   MB lanes                           reference src/BVH.cpp:1316-1335
 *   ProxyObject::intersect             reference src/ProxyObject.cpp:76-95
 *   Ray reciprocal convention          reference src/Ray.h:79-90
 * Deliberate difference: 1/det and 1/w use an exact FP32 division where the reference uses
 * rcpps + one Newton step (src/SSE.h:67-86), whose low bits are CPU-vendor specific.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "../include/miro_gpu.h"

#define ORACLE_STACK 256   /* QBVH_Node* BVH_Stack[256], src/BVH.cpp:1133 */

typedef struct { float o[3], d[3], id[3], time; } oray;
typedef struct { float t, a, b; int32_t prim, inst; } ohit;

static float rcp_dir(float d) {
    float id = 1.0f / d;
    if (d == 0.f) id = (id < -0.f) ? -MIRO_GPU_TMAX : MIRO_GPU_TMAX;   /* src/Ray.h:79-90 */
    return id;
}
static void ray_set(oray* r, const float* o, const float* d, float time) {
    for (int k = 0; k < 3; ++k) { r->o[k] = o[k]; r->d[k] = d[k]; r->id[k] = rcp_dir(d[k]); }
    r->time = time;
}
static float minf(float a, float b) { return a < b ? a : b; }   /* minps/maxps semantics: second operand on NaN */
static float maxf(float a, float b) { return a > b ? a : b; }

/* one lane of intersect4; returns 1 and the candidate (t,a,b) when the lane passes all masks */
static int mt_lane(const oray* r, const float* A, const float* B, const float* C, float tMin, float tMax, float* t, float* a, float* b) {
    const float e0x = B[0] - A[0], e0y = B[1] - A[1], e0z = B[2] - A[2];
    const float e1x = C[0] - A[0], e1y = C[1] - A[1], e1z = C[2] - A[2];
    const float px = r->d[1] * e1z - r->d[2] * e1y;
    const float py = -1.0f * (r->d[0] * e1z - r->d[2] * e1x);
    const float pz = r->d[0] * e1y - r->d[1] * e1x;
    const float det = e0x * px + (e0y * py + e0z * pz);          /* SoADot, src/SSE.h:110-113 */
    const float inv = 1.0f / det;
    const float tx = r->o[0] - A[0], ty = r->o[1] - A[1], tz = r->o[2] - A[2];
    const float av = inv * (tx * px + (ty * py + tz * pz));
    if (!(av >= 0.f && av <= 1.f)) return 0;
    const float qx = ty * e0z - tz * e0y;
    const float qy = -1.0f * (tx * e0z - tz * e0x);
    const float qz = tx * e0y - ty * e0x;
    const float bv = inv * (r->d[0] * qx + (r->d[1] * qy + r->d[2] * qz));
    if (!(bv >= 0.f && bv <= 1.f && av + bv <= 1.f)) return 0;
    const float tv = inv * (e1x * qx + (e1y * qy + e1z * qz));
    if (!(tv >= tMin && tv < tMax)) return 0;
    *t = tv; *a = av; *b = bv;
    return 1;
}

static int traverse(const miro_gpu_scene_desc* s, int32_t root, const oray* r, float tMin, ohit* hit, int32_t cur_inst,
                    uint64_t* n_nodes, uint64_t* n_tris);

/* Texture::getLookupAlpha at the hit's interpolated uv (src/Texture.cpp:12-41, src/BVH.cpp:1403-1421); 1 without an alpha map */
static float hit_alpha(const miro_gpu_scene_desc* s, uint32_t prim, float a, float b) {
    if (!s->prims) return 1.0f;
    const miro_gpu_prim* pr = &s->prims[prim];
    const int32_t am = s->materials[pr->material].alpha_map;
    if (am < 0) return 1.0f;
    const miro_gpu_texture* t = &s->textures[am];
    if (t->channels != 4) return 1.0f;                    /* Texture::getPixel: alpha 1 for RGB / GRAYSCALE images */
    float u = a, v = b;
    if (pr->uv[0] != 0xffffffffu) {
        const float c = 1.0f - a - b;
        const float *t0 = s->uvs + (size_t)pr->uv[0] * 2, *t1 = s->uvs + (size_t)pr->uv[1] * 2, *t2 = s->uvs + (size_t)pr->uv[2] * 2;
        u = t0[0] * c + t1[0] * a + t2[0] * b; v = t0[1] * c + t1[1] * a + t2[1] * b;
    }
    u = u - (float)(int)u; v = v - (float)(int)v;
    if (u < 0.0f) u = u + 1.0f;
    if (v < 0.0f) v = v + 1.0f;
    v = 1.0f - v;
    const float px = u * t->width, py = v * t->height;
    const float x1 = floorf(px), y1 = floorf(py), dx = px - x1, dy = py - y1;
    #define TEXA(X, Y) t->texels[((size_t)((Y) % t->height) * t->width + ((X) % t->width)) * 4 + 3]
    const float q1 = TEXA((int)x1, (int)y1) * (1.0f - dx) + TEXA((int)x1 + 1, (int)y1) * dx;
    const float q2 = TEXA((int)x1, (int)y1 + 1) * (1.0f - dx) + TEXA((int)x1 + 1, (int)y1 + 1) * dx;
    #undef TEXA
    return q1 * (1.0f - dy) + q2 * dy;
}

/* a leaf = one TriCache4 packet */
static int intersect_leaf(const miro_gpu_scene_desc* s, int32_t ref, const oray* r, float tMin, ohit* hit, int32_t cur_inst,
                          uint64_t* n_nodes, uint64_t* n_tris) {
    const uint32_t u = (uint32_t)ref, kind = (u >> 29) & 3u, count = ((u >> MIRO_GPU_LEAF_INDEX_BITS) & 7u) + 1u;
    const uint32_t first = u & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u);
    int any = 0;
    if (kind == MIRO_GPU_KIND_INST) {
        for (uint32_t i = 0; i < count; ++i) {   /* proxy lanes are dispatched first, src/BVH.cpp:1306-1315 */
            const miro_gpu_instance* in = &s->instances[first + i];
            float no[3], nd[3];
            for (int k = 0; k < 3; ++k) {
                const float* m = &in->inv[4 * k];
                /* src/Matrix4x4.h:728-733 (dpps over [o 1]: (p0 + p1) + (p2 + p3), times recipps(w), w = 1) and :699-701 */
                const float wr = in->w_recip == 0.f ? 1.0f : in->w_recip;
                no[k] = ((m[0] * r->o[0] + m[1] * r->o[1]) + (m[2] * r->o[2] + m[3])) * wr;
                nd[k] = (m[0] * r->d[0] + m[1] * r->d[1]) + m[2] * r->d[2];
            }
            oray nr; ray_set(&nr, no, nd, r->time);
            ohit nh = *hit;   /* newHit.t = result.t */
            if (traverse(s, in->blas_root, &nr, tMin, &nh, (int32_t)(first + i), n_nodes, n_tris)) { *hit = nh; any = 1; }
        }
        return any;
    }
    const float tMax = hit->t;   /* all lanes of the packet see the tMax at entry */
    float ct[4], ca[4], cb[4]; int cvalid[4] = {0, 0, 0, 0};     /* per-lane candidates of the packet */
    for (uint32_t i = 0; i < count; ++i) {
        float A[3], B[3], C[3], t, a, b;
        if (kind == MIRO_GPU_KIND_TRI) {
            const miro_gpu_tri* tr = &s->tris[first + i];
            memcpy(A, tr->v0, 12); memcpy(B, tr->v1, 12); memcpy(C, tr->v2, 12);
        } else {
            const miro_gpu_mbtri* tr = &s->mbtris[first + i];
            const float w1 = r->time, w0 = 1.f - r->time;
            for (int k = 0; k < 3; ++k) {
                A[k] = w1 * tr->pose[1].v0[k] + w0 * tr->pose[0].v0[k];
                B[k] = w1 * tr->pose[1].v1[k] + w0 * tr->pose[0].v1[k];
                C[k] = w1 * tr->pose[1].v2[k] + w0 * tr->pose[0].v2[k];
            }
        }
        (*n_tris)++;
        if (mt_lane(r, A, B, C, tMin, tMax, &t, &a, &b)) { ct[i] = t; ca[i] = a; cb[i] = b; cvalid[i] = 1; }
    }
    /* nearest valid lane (lowest index on ties); a lane whose alpha map says "cut out" (< 0.5) is dropped and the next
     * nearest is tried (src/BVH.cpp:1387-1435) */
    float bt = MIRO_GPU_TMAX, ba = 0.f, bb = 0.f; int bl = -1;
    for (int tries = 0; tries < 4; ++tries) {
        int li = -1; float lowest = MIRO_GPU_TMAX;
        for (uint32_t i = 0; i < count; ++i) if (cvalid[i] && ct[i] < lowest) { lowest = ct[i]; li = (int)i; }
        if (li < 0) break;
        cvalid[li] = 0;
        const uint32_t prim = (kind == MIRO_GPU_KIND_TRI ? 0u : s->n_tris) + first + (uint32_t)li;
        if (hit_alpha(s, prim, ca[li], cb[li]) < 0.5f) continue;
        bt = lowest; ba = ca[li]; bb = cb[li]; bl = li;
        break;
    }
    if (bl >= 0 && bt < hit->t) {
        hit->t = bt; hit->a = ba; hit->b = bb; hit->inst = cur_inst;
        hit->prim = (int32_t)((kind == MIRO_GPU_KIND_TRI ? 0u : s->n_tris) + first + (uint32_t)bl);
        return 1;
    }
    return 0;
}

static int traverse(const miro_gpu_scene_desc* s, int32_t root, const oray* r, float tMin, ohit* hit, int32_t cur_inst,
                    uint64_t* n_nodes, uint64_t* n_tris) {
    int any = 0;
    if (root == MIRO_GPU_CHILD_EMPTY) return 0;
    if (root < 0) return intersect_leaf(s, root, r, tMin, hit, cur_inst, n_nodes, n_tris);
    int32_t stack[ORACLE_STACK];
    int sp = 1;
    stack[0] = root;
    while (--sp >= 0) {
        const miro_gpu_node* n = &s->nodes[stack[sp]];
        (*n_nodes)++;
        int boxHit = 0;
        for (int i = 0; i < 4; ++i) {
            const float t0x = (n->lo_x[i] - r->o[0]) * r->id[0], t1x = (n->hi_x[i] - r->o[0]) * r->id[0];
            const float t0y = (n->lo_y[i] - r->o[1]) * r->id[1], t1y = (n->hi_y[i] - r->o[1]) * r->id[1];
            const float t0z = (n->lo_z[i] - r->o[2]) * r->id[2], t1z = (n->hi_z[i] - r->o[2]) * r->id[2];
            const float t0 = maxf(minf(t0x, t1x), maxf(minf(t0y, t1y), minf(t0z, t1z)));
            const float t1 = minf(maxf(t0x, t1x), minf(maxf(t0y, t1y), maxf(t0z, t1z)));
            if (maxf(t0, tMin) <= minf(t1, hit->t)) boxHit |= 1 << i;
        }
        int32_t tmp[4]; int childHit = 0;
        for (int i = 0; i < 4; ++i) {
            if (!(boxHit & (1 << i))) continue;
            const int32_t c = n->child[i];
            if (c == MIRO_GPU_CHILD_EMPTY) continue;
            if (c < 0) { if (intersect_leaf(s, c, r, tMin, hit, cur_inst, n_nodes, n_tris)) any = 1; }
            else tmp[childHit++] = c;
        }
        for (int i = 0; i < childHit && sp + i < ORACLE_STACK; ++i) stack[sp + i] = tmp[i];
        sp += childHit;
    }
    return any;
}

/* Scene::trace over a ray buffer.  counters (optional): [0] nodes visited, [1] triangles tested. */
int oracle_trace_closest(const miro_gpu_scene_desc* s, const miro_gpu_ray* rays, size_t n, miro_gpu_hit* hits, uint64_t* counters) {
    uint64_t nn = 0, nt = 0;
    for (size_t i = 0; i < n; ++i) {
        oray r; const float o[3] = {rays[i].ox, rays[i].oy, rays[i].oz}, d[3] = {rays[i].dx, rays[i].dy, rays[i].dz};
        ray_set(&r, o, d, rays[i].time);
        ohit h; h.t = rays[i].tmax; h.a = h.b = 0.f; h.prim = -1; h.inst = -1;
        const int hit = traverse(s, s->root, &r, rays[i].tmin, &h, -1, &nn, &nt);
        if (hit && h.prim >= 0) { hits[i].t = h.t; hits[i].a = h.a; hits[i].b = h.b; hits[i].prim = h.prim; hits[i].inst = h.inst; }
        else { hits[i].t = -1.f; hits[i].a = hits[i].b = 0.f; hits[i].prim = -1; hits[i].inst = -1; }
    }
    if (counters) { counters[0] += nn; counters[1] += nt; }
    return 0;
}

/* Shadow query: the reference traces a full closest-hit traversal and uses the result as a boolean. */
int oracle_trace_any(const miro_gpu_scene_desc* s, const miro_gpu_ray* rays, size_t n, uint32_t* bits, uint64_t* counters) {
    uint64_t nn = 0, nt = 0;
    for (size_t w = 0; w < (n + 31) / 32; ++w) bits[w] = 0;
    for (size_t i = 0; i < n; ++i) {
        oray r; const float o[3] = {rays[i].ox, rays[i].oy, rays[i].oz}, d[3] = {rays[i].dx, rays[i].dy, rays[i].dz};
        ray_set(&r, o, d, rays[i].time);
        ohit h; h.t = rays[i].tmax; h.a = h.b = 0.f; h.prim = -1; h.inst = -1;
        if (traverse(s, s->root, &r, rays[i].tmin, &h, -1, &nn, &nt)) bits[i >> 5] |= 1u << (i & 31);
    }
    if (counters) { counters[0] += nn; counters[1] += nt; }
    return 0;
}

/* Scene::trace for one ray (used by oracle/miro_oracle_shade.c).  Returns 1 on a hit in [tmin, tmax). */
int oracle_trace_one(const miro_gpu_scene_desc* s, const float o[3], const float d[3], float time, float tmin, float tmax, miro_gpu_hit* out) {
    uint64_t nn = 0, nt = 0;
    oray r; ray_set(&r, o, d, time);
    ohit h; h.t = tmax; h.a = h.b = 0.f; h.prim = -1; h.inst = -1;
    const int hit = traverse(s, s->root, &r, tmin, &h, -1, &nn, &nt);
    if (hit && h.prim >= 0) { out->t = h.t; out->a = h.a; out->b = h.b; out->prim = h.prim; out->inst = h.inst; return 1; }
    out->t = -1.f; out->a = out->b = 0.f; out->prim = -1; out->inst = -1;
    return 0;
}
