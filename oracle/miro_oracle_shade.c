/* oracle/miro_oracle_shade.c — CPU restatement of the reference's render loop, shading and light sampling.
 *
 * TEST INFRASTRUCTURE ONLY (see the header of miro_oracle.c): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call this file.
 *
 * Unlike the product (a wavefront of queues), this file keeps the reference's own control flow — shade() calling
 * sampleLight() and calculatePathTracing(), which traces and calls shade() again — so that two independently
 * structured implementations can be compared pixel by pixel.  It restates:
 *   Scene::adaptiveSampleScene / sampleScene     reference src/Scene.cpp:219-293
 *   Camera::eyeRayAdaptive, getTimeSample        reference src/Camera.cpp:116-174, src/Camera.h:46
 *   HitInfo::getAllInfos                         reference src/Ray.cpp:5-50
 *   Lambert::shade                               reference src/Lambert.cpp:19-53
 *   Blinn::shade (diffuse + highlight), calculatePathTracing   reference src/Blinn.cpp:39-236,335
 *   Material::getCosineDistributedSamples, getEnvironmentColor reference src/Material.cpp:14-63
 *   PointLight / RectangleLight / DomeLight::sampleLight       reference src/PointLight.cpp:8-82,
 *                                                src/RectangleLight.cpp:42-137, src/DomeLight.cpp:8-161, src/DomeLight.h:10-42
 *   Texture::getLookup / getLookupXYZ3 / getPixel               reference src/Texture.cpp:43-125
 *   Image gamma table                            reference src/Image.cpp:19-35
 *
 * Parity status: pinned STATISTICALLY against float radiance images rendered by the unmodified reference
 * (tests/golden, tests/test_oracle_vs_reference.py) and exactly (to FP32 rounding) for the deterministic C1 config.
 * Deliberate differences from the reference, shared with the product so the two can be compared sample by sample:
 *   * random numbers: the reference draws from one global MT19937 through per-thread blocks (src/Scene.cpp:26-47), so
 *     the numbers a sample sees depend on thread scheduling.  Here every draw has an address
 *     (pixel, camera-sample ordinal, path, depth, purpose, light, pass, sample, attempt) hashed with Philox4x32-10;
 *   * rcpps/rsqrtss + one Newton step (src/SSE.h:67-101) are replaced by exact FP32 division / sqrtf;
 *   * the dome light draws its cell by inverting the reference's marginal/conditional CDFs (Distribution1D::sample);
 *     the product uses an alias table over the same probability mass function, so dome-lit images agree in
 *     distribution, not sample by sample.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/miro_gpu.h"

int oracle_trace_one(const miro_gpu_scene_desc* s, const float o[3], const float d[3], float time, float tmin, float tmax, miro_gpu_hit* out);

#define O_PI 3.1415926f                      /* src/Miro.h:57 */
#define O_1_PI (1.0f / O_PI)
#define O_1_4PI (0.25f / O_PI)
#define O_2_PI2 (2.f * (O_PI * O_PI))
#define O_EPS MIRO_GPU_EPSILON

typedef struct { float x, y, z; } v3;
static v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static v3 scl(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static v3 normalize(v3 a) { return scl(a, 1.0f / sqrtf(dot(a, a))); }
static float average(v3 a) { return (a.x + a.y + a.z) * 0.333333f; }          /* src/Vector3.h:258 */

/* ---- counter-based random numbers (same addressing as csrc/shading.cuh) ------------------------------------- */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static float unit(uint32_t u) { return ((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f); }
enum { RP_CAMERA = 0, RP_LENS = 1, RP_COSINE = 2, RP_LIGHT = 3, RP_ROULETTE = 4, RP_GLOSS = 5 };
typedef struct { uint32_t pixel, sample, path_depth; uint64_t seed; } raddr;
static void rand4(const raddr* a, uint32_t purpose, uint32_t light, uint32_t pass, uint32_t sample, uint32_t attempt, float out[4]) {
    uint32_t c[4] = {a->pixel, a->sample, a->path_depth, (purpose << 28) | ((pass >> 1) << 27) | (light << 24) | ((pass & 1u) << 23) | ((sample & 0x7ffu) << 12) | (attempt & 0xfffu)};
    philox4x32_10(c, (uint32_t)a->seed, (uint32_t)(a->seed >> 32));
    for (int k = 0; k < 4; ++k) out[k] = unit(c[k]);
}

/* ---- textures ------------------------------------------------------------------------------------------------- */
static void tex_pixel(const miro_gpu_texture* t, int x, int y, float out[4]) {
    x = x % t->width; y = y % t->height;
    if (t->channels == 1) { const float g = t->texels[(size_t)y * t->width + x]; out[0] = out[1] = out[2] = g; out[3] = 1.f; return; }
    const float* p = t->texels + ((size_t)y * t->width + x) * t->channels;
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = t->channels == 4 ? p[3] : 1.f;
}
static void tex_lookup(const miro_gpu_texture* t, float u, float v, float out[4]) {
    u = u - (float)(int)u; v = v - (float)(int)v;
    if (u < 0.0f) u = u + 1.0f;
    if (v < 0.0f) v = v + 1.0f;
    v = 1.0f - v;
    const float px = u * t->width, py = v * t->height;
    const float x1 = floorf(px), x2 = x1 + 1.0f, dx = px - x1, y1 = floorf(py), y2 = y1 + 1.0f, dy = py - y1;
    float a[4], b[4], c[4], d[4];
    tex_pixel(t, (int)x1, (int)y1, a); tex_pixel(t, (int)x2, (int)y1, b); tex_pixel(t, (int)x1, (int)y2, c); tex_pixel(t, (int)x2, (int)y2, d);
    for (int k = 0; k < 4; ++k) out[k] = (a[k] * (1.0f - dx) + b[k] * dx) * (1.0f - dy) + (c[k] * (1.0f - dx) + d[k] * dx) * dy;
}
static v3 tex_lookup_dir(const miro_gpu_texture* t, v3 d) {
    const float theta = atan2f(d.z, d.x) + O_PI;
    float y = d.y; if (y > 1.f) y = 1.f; if (y < -1.f) y = -1.f;
    const float phi = acosf(y);
    float c[4];
    tex_lookup(t, theta * 0.5f * O_1_PI, 1.0f - (phi * O_1_PI), c);
    return V(c[0], c[1], c[2]);
}

/* ---- dome light tables (DomeLight::setTexture + Distribution1D) ---------------------------------------------- */
typedef struct { float* func; float* cdf; float funcInt, invFuncInt; int count; } dist1d;
static void dist_init(dist1d* d, const float* f, int n) {
    d->func = (float*)malloc(sizeof(float) * n); d->cdf = (float*)malloc(sizeof(float) * (n + 1)); d->count = n;
    memcpy(d->func, f, sizeof(float) * n);
    d->cdf[0] = 0.f;
    for (int i = 1; i < n + 1; ++i) d->cdf[i] = d->cdf[i - 1] + f[i - 1] / n;
    d->funcInt = d->cdf[n];
    for (int i = 1; i < n + 1; ++i) d->cdf[i] /= d->funcInt;
    d->invFuncInt = 1.f / d->funcInt;
}
static float dist_sample(const dist1d* d, float u, float* pdf) {
    int lo = 0, hi = d->count + 1;                        /* std::lower_bound(cdf, cdf+count+1, u) */
    while (lo < hi) { const int mid = (lo + hi) / 2; if (d->cdf[mid] < u) lo = mid + 1; else hi = mid; }
    int offset = lo - 1;
    if (offset < 0) offset = 0;                           /* u == 0: the reference reads cdf[-1]; u is in (0,1) here */
    if (offset > d->count - 1) offset = d->count - 1;
    u = (u - d->cdf[offset]) / (d->cdf[offset + 1] - d->cdf[offset]);
    *pdf = d->func[offset] * d->invFuncInt;
    return offset + u;
}
typedef struct { int nu, nv; dist1d u; dist1d* v; float *cu, *su, *cv, *sv; } dome;
static void dome_init(dome* D, const miro_gpu_texture* t) {
    const int nu = t->width, nv = t->height;
    D->nu = nu; D->nv = nv;
    float* func = (float*)malloc(sizeof(float) * (nu > nv ? nu : nv));
    float* sinVals = (float*)malloc(sizeof(float) * nv);
    for (int i = 0; i < nv; ++i) sinVals[i] = sinf(O_PI * (float)(i + .5) / (float)nv);
    D->v = (dist1d*)malloc(sizeof(dist1d) * nu);
    for (int u = 0; u < nu; ++u) {
        const float up = (float)u / (float)nu;
        for (int v = 0; v < nv; ++v) {
            float c[4]; tex_lookup(t, up, (float)v / (float)nv, c);
            func[v] = average(V(c[0], c[1], c[2])) * sinVals[v];
        }
        dist_init(&D->v[u], func, nv);
    }
    for (int u = 0; u < nu; ++u) func[u] = D->v[u].funcInt;
    dist_init(&D->u, func, nu);
    D->cu = (float*)malloc(sizeof(float) * (nu + 1)); D->su = (float*)malloc(sizeof(float) * (nu + 1));
    D->cv = (float*)malloc(sizeof(float) * (nv + 1)); D->sv = (float*)malloc(sizeof(float) * (nv + 1));
    float inv = 1.f / (float)nu;
    for (int i = 0; i < nu + 1; ++i) { D->cu[i] = cosf(i * inv * 2.f * O_PI); D->su[i] = sinf(i * inv * 2.f * O_PI); }
    inv = 1.f / (float)nv;
    for (int i = 0; i < nv + 1; ++i) { D->cv[i] = cosf(i * inv * O_PI); D->sv[i] = sinf(i * inv * O_PI); }
    free(func); free(sinVals);
}
static void dome_free(dome* D) {
    for (int u = 0; u < D->nu; ++u) { free(D->v[u].func); free(D->v[u].cdf); }
    free(D->v); free(D->u.func); free(D->u.cdf); free(D->cu); free(D->su); free(D->cv); free(D->sv);
}

/* ---- render context ------------------------------------------------------------------------------------------- */
typedef struct {
    const miro_gpu_scene_desc* s;
    const miro_gpu_render_params* p;
    dome* domes;                       /* per light (only dome lights initialised) */
    uint64_t rays;                     /* Scene::trace calls */
} octx;

/* Ray::IORList (src/Ray.h:43-51): the history of refraction indices a ray has traversed; shade() mutates it in place */
typedef struct { float v[12]; unsigned idx; } iorlist;
static void ior_init(iorlist* l) { l->v[0] = 1.0f; l->idx = 0; }
static float ior_top(const iorlist* l) { return l->v[l->idx]; }
static void ior_pop(iorlist* l) { if (l->idx > 0) l->idx--; }
static void ior_push(iorlist* l, float x) { if (l->idx < 11) l->idx++; l->v[l->idx] = x; }

typedef struct { v3 o, d; float time; iorlist ior; int bounces; int is_refract; } oray2;
typedef struct { v3 P, N, geoN, T, BT; float u, v; uint32_t material; } surf;

static int trace(octx* c, v3 o, v3 d, float time, float tmin, float tmax, miro_gpu_hit* h) {
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    c->rays++;
    return oracle_trace_one(c->s, oo, dd, time, tmin, tmax, h);
}

static v3 environment(const octx* c, v3 d) {          /* Scene.cpp:234-240 / Material::getEnvironmentColor */
    if (c->s->env_map >= 0) return scl(tex_lookup_dir(&c->s->textures[c->s->env_map], d), c->s->env_exposure);
    return V(c->s->bg_color[0], c->s->bg_color[1], c->s->bg_color[2]);
}

static surf surface_at(const octx* c, const oray2* r, const miro_gpu_hit* h) {     /* HitInfo::getAllInfos, Ray.cpp:5-50 */
    const miro_gpu_scene_desc* s = c->s;
    surf o;
    o.P = add(r->o, scl(r->d, h->t));
    o.T = V(0, 0, 0); o.BT = V(0, 0, 0);
    const miro_gpu_prim* pr = &s->prims[h->prim];
    o.material = pr->material;
    const miro_gpu_tri* t = (uint32_t)h->prim < s->n_tris ? &s->tris[h->prim] : &s->mbtris[(uint32_t)h->prim - s->n_tris].pose[0];
    const v3 e0 = V(t->v1[0] - t->v0[0], t->v1[1] - t->v0[1], t->v1[2] - t->v0[2]), e1 = V(t->v2[0] - t->v0[0], t->v2[1] - t->v0[1], t->v2[2] - t->v0[2]);
    o.geoN = normalize(cross(e0, e1));
    const float a = h->a, b = h->b, cc = 1.0f - a - b;
    const float *n0 = s->normals + (size_t)pr->n[0] * 3, *n1 = s->normals + (size_t)pr->n[1] * 3, *n2 = s->normals + (size_t)pr->n[2] * 3;
    o.N = normalize(V(n0[0] * cc + n1[0] * a + n2[0] * b, n0[1] * cc + n1[1] * a + n2[1] * b, n0[2] * cc + n1[2] * a + n2[2] * b));
    if (h->inst >= 0) {
        const float* m = s->inst_normal_xform + (size_t)h->inst * 9;
        const v3 r0 = V(m[0], m[1], m[2]), r1 = V(m[3], m[4], m[5]), r2 = V(m[6], m[7], m[8]);
        o.geoN = normalize(V(dot(r0, o.geoN), dot(r1, o.geoN), dot(r2, o.geoN)));
        o.N = normalize(V(dot(r0, o.N), dot(r1, o.N), dot(r2, o.N)));
    }
    if (pr->uv[0] != 0xffffffffu) {
        const float *t0 = s->uvs + (size_t)pr->uv[0] * 2, *t1 = s->uvs + (size_t)pr->uv[1] * 2, *t2 = s->uvs + (size_t)pr->uv[2] * 2;
        o.u = t0[0] * cc + t1[0] * a + t2[0] * b; o.v = t0[1] * cc + t1[1] * a + t2[1] * b;
        if (s->tangents && s->bitangents) {        /* Ray.cpp:35-36: indexed by the NORMAL indices, not transformed by the proxy */
            const float *g0 = s->tangents + (size_t)pr->n[0] * 3, *g1 = s->tangents + (size_t)pr->n[1] * 3, *g2 = s->tangents + (size_t)pr->n[2] * 3;
            o.T = normalize(V(g0[0] * cc + g1[0] * a + g2[0] * b, g0[1] * cc + g1[1] * a + g2[1] * b, g0[2] * cc + g1[2] * a + g2[2] * b));
            const float *h0 = s->bitangents + (size_t)pr->n[0] * 3, *h1 = s->bitangents + (size_t)pr->n[1] * 3, *h2 = s->bitangents + (size_t)pr->n[2] * 3;
            o.BT = normalize(V(h0[0] * cc + h1[0] * a + h2[0] * b, h0[1] * cc + h1[1] * a + h2[1] * b, h0[2] * cc + h1[2] * a + h2[2] * b));
        }
    } else { o.u = a; o.v = b; }
    return o;
}

/* The "full method" of the shadow test, Light::setFastShadows(false) (RectangleLight.cpp:93-118, DomeLight.cpp:123-146): the ray is
 * walked hit by hit; a surface whose INTERPOLATED normal (HitInfo::getInterpolatedNormal, Ray.cpp:52-66: object space, not
 * transformed by a proxy) faces the ray multiplies the visibility by its material's refractAmt.  Kept as the reference has it:
 * sampleHit is never reset, so the previous segment's hit distance is the next segment's tMax. */
static float full_shadow(octx* c, v3 from, v3 dir, float time, float first_tmax, float distance) {
    const miro_gpu_scene_desc* s = c->s;
    float attenuate = 1.0f, traversed = 0.0f, hit_t = first_tmax;
    v3 o = from;
    miro_gpu_hit h;
    for (int guard = 0; traversed < distance && attenuate > O_EPS && guard < 4096; ++guard) {
        if (trace(c, o, dir, time, O_EPS, hit_t, &h)) {
            const miro_gpu_prim* pr = &s->prims[h.prim];
            const float a = h.a, b = h.b, cc = 1.0f - a - b;
            const float *n0 = s->normals + (size_t)pr->n[0] * 3, *n1 = s->normals + (size_t)pr->n[1] * 3, *n2 = s->normals + (size_t)pr->n[2] * 3;
            const v3 N = normalize(V(n0[0] * cc + n1[0] * a + n2[0] * b, n0[1] * cc + n1[1] * a + n2[1] * b, n0[2] * cc + n1[2] * a + n2[2] * b));
            if (dot(N, scl(dir, -1.f)) > 0.0f) attenuate *= s->materials[pr->material].refract_amt;
            o = add(o, scl(dir, h.t)); traversed += h.t; hit_t = h.t;
        } else traversed = distance;
    }
    return attenuate;
}

/* ---- lights: one call = one Light::sampleLight ---------------------------------------------------------------- */
static v3 sample_light(octx* c, uint32_t li, v3 from, v3 normal, float time, v3 rVec, float* outSpec, int isSecondary, uint32_t pass, const raddr* addr) {
    const miro_gpu_light* L = &c->s->lights[li];
    miro_gpu_hit sh;
    *outSpec = 0.f;
    if (L->kind == MIRO_GPU_LIGHT_POINT) {                                   /* PointLight.cpp:8-82 */
        v3 l = sub(V(L->p0[0], L->p0[1], L->p0[2]), from);
        float nDotL = dot(normal, l);
        if (!(nDotL > 0.0f)) return V(0, 0, 0);
        float falloff = dot(l, l);
        const float distance = sqrtf(falloff), distanceRecip = 1.0f / distance;
        falloff = 1.0f / falloff;
        l = scl(l, distanceRecip); nDotL *= distanceRecip;
        float attenuate = 1.0f;
        /* full method (PointLight.cpp:49-70): `sampleHit.t = distance; while (sampleHit.t < distance)` never runs: no shadow at all */
        if (L->cast_shadows && !L->full_shadows && trace(c, from, l, time, 0.001f, distance, &sh)) attenuate = 0.0f;
        attenuate *= nDotL;
        const float rl = dot(rVec, l);
        *outSpec = (rl > 0.f ? rl : 0.f) * attenuate;
        const float e = L->power * falloff * O_1_4PI * attenuate;
        return V(e, e, e);
    }
    if (L->kind == MIRO_GPU_LIGHT_RECT) {                                    /* RectangleLight.cpp:42-137 */
        const v3 v1 = V(L->p0[0], L->p0[1], L->p0[2]), v2 = V(L->p1[0], L->p1[1], L->p1[2]), v3_ = V(L->p2[0], L->p2[1], L->p2[2]);
        v3 tmpResult = V(0, 0, 0); float tmpSpec = 0.f, falloff = 1.0f, samplesDoneRecip = 1.0f;
        int samplesDone = 0, cutOff = 0;
        do {
            float r[4]; rand4(addr, RP_LIGHT, li, pass, (uint32_t)samplesDone, 0, r);
            const float e1 = r[0]; float e2 = r[1]; e2 = (e2 > 0.99f) ? 0.99f : e2;
            v3 dir = sub(add(add(v1, scl(sub(v2, v1), e1)), scl(sub(v3_, v1), e2)), from);
            float nDotL = dot(normal, dir), attenuate = 1.0f;
            if (nDotL > O_EPS) {
                falloff = dot(dir, dir);
                const float distance = sqrtf(falloff), distanceRecip = 1.0f / distance;
                falloff = 1.0f / falloff;
                dir = scl(dir, distanceRecip);
                if (L->cast_shadows && L->full_shadows) attenuate = full_shadow(c, from, dir, time, distance - O_EPS, distance);
                else if (L->cast_shadows && trace(c, from, dir, time, O_EPS, distance - O_EPS, &sh)) attenuate = 0.0f;
            } else attenuate = 0.0f;
            const float e = L->power * falloff * O_1_4PI;
            samplesDone++; samplesDoneRecip = 1.0f / (float)samplesDone;
            cutOff = average(scl(V(e, e, e), samplesDoneRecip)) < L->noise_threshold;
            tmpResult = add(tmpResult, scl(V(e, e, e), attenuate));
            const float rl = dot(rVec, dir);
            tmpSpec += (rl > 0.f ? rl : 0.f) * attenuate;
        } while (samplesDone < L->num_samples && !cutOff);
        *outSpec = tmpSpec * samplesDoneRecip;
        return scl(tmpResult, samplesDoneRecip);
    }
    /* DomeLight.cpp:80-161 */
    const dome* D = &c->domes[li];
    const miro_gpu_texture* tex = &c->s->textures[L->texture];
    v3 tmpResult = V(0, 0, 0); float tmpSpec = 0.f, samplesDoneRecip = 1.0f;
    int samplesDone = 0, cutOff = 0;
    const int numSamples = isSecondary ? 1 : L->num_samples;
    do {
        v3 direction = V(0, 0, 0); float pdfs[2] = {0, 0}, sinTheta = 0.f; int found = 0;
        for (uint32_t attempt = 0; attempt < 64 && !found; ++attempt) {       /* `continue` on a back-facing draw */
            float r[4]; rand4(addr, RP_LIGHT, li, pass, (uint32_t)samplesDone, attempt, r);
            for (int half = 0; half < 2 && !found; ++half) {
                const float e1 = half ? r[2] : r[0], e2 = half ? r[3] : r[1];
                const float fu = dist_sample(&D->u, e1, &pdfs[0]);
                const int u = ((int)fu == D->u.count) ? (int)fu - 1 : (int)fu;
                const float fv = dist_sample(&D->v[u], e2, &pdfs[1]);
                const float cosTheta = D->cv[(int)fv]; sinTheta = D->sv[(int)fv];
                const float sinPhi = D->su[(int)fu], cosPhi = D->cu[(int)fu];
                direction = V(-sinTheta * cosPhi, -cosTheta, -sinTheta * sinPhi);
                if (dot(normal, direction) < 0.0f) continue;
                found = 1;
            }
        }
        v3 E = V(0, 0, 0); float attenuate = 0.0f;
        if (found) {
            const float pdf = (pdfs[0] * pdfs[1]) / (O_2_PI2 * sinTheta);
            const v3 imageSample = tex_lookup_dir(tex, direction);
            attenuate = 1.0f;
            if (L->full_shadows) attenuate = full_shadow(c, from, direction, time, MIRO_GPU_TMAX, MIRO_GPU_TMAX);
            else if (trace(c, from, direction, time, O_EPS, MIRO_GPU_TMAX, &sh)) attenuate = 0.0f;
            E = scl(imageSample, L->power / pdf);
            if (!(pdf > 0.f) || isinf(pdf) || isnan(pdf)) E = V(0, 0, 0);
        }
        samplesDone++; samplesDoneRecip = 1.0f / (float)samplesDone;
        cutOff = average(scl(E, samplesDoneRecip)) < L->noise_threshold;
        tmpResult = add(tmpResult, scl(E, attenuate));
        if (found) tmpSpec += dot(rVec, direction) * attenuate;
    } while (samplesDone < numSamples && !cutOff);
    *outSpec = tmpSpec * samplesDoneRecip;
    return scl(tmpResult, samplesDoneRecip);
}

static v3 cosine_sample(v3 N, float e1, float e2) {                           /* Material.cpp:14-42 */
    e2 = (e2 > 0.99f) ? 0.99f : e2;
    const v3 axis = (fabsf(N.x) > 0.1f) ? V(0, 1, 0) : V(1, 0, 0);
    const v3 u = normalize(cross(axis, N)), v = cross(N, u);
    const float ang = 2 * O_PI * e1, s2 = sqrtf(e2), s1 = sqrtf(fabsf(1.0f - e2));
    return normalize(add(add(scl(u, cosf(ang) * s2), scl(v, sinf(ang) * s2)), scl(N, s1)));
}

static float fresnel(float n1, float n2, float cosThetaI) {                    /* Material::fresnel, src/Material.h:47-55 */
    const float n1CosTh = n1 * cosThetaI;
    const float n1_n2SinTh = n1 * sinf(acosf(cosThetaI)) / n2;
    const float r = sqrtf(1.0f - n1_n2SinTh * n1_n2SinTh);
    const float n2CosTh = n2 * ((0.0f < r) ? r : 0.0f);                      /* max(0.0f, NaN) = 0 */
    const float Rs = (n1CosTh - n2CosTh) / (n1CosTh + n2CosTh);
    return Rs * Rs;
}

/* depth = giBounces; ray->bounces = reflect / refract bounces; isSecondary as in the reference's signature */
static v3 shade(octx* c, oray2* ray, const miro_gpu_hit* hit, uint32_t pixel, uint32_t sample, uint32_t path, int depth, int isSecondary);

static v3 shade(octx* c, oray2* ray, const miro_gpu_hit* hit, uint32_t pixel, uint32_t sample, uint32_t path, int depth, int isSecondary) {
    const miro_gpu_scene_desc* s = c->s;
    const surf sf = surface_at(c, ray, hit);
    const miro_gpu_material* m = &s->materials[sf.material];
    raddr addr; addr.pixel = pixel; addr.sample = sample; addr.path_depth = path | ((uint32_t)(depth + ray->bounces) << 16); addr.seed = c->p->seed;
    v3 kd = V(m->kd[0], m->kd[1], m->kd[2]);
    if (m->color_map >= 0) { float t[4]; tex_lookup(&s->textures[m->color_map], sf.u, sf.v, t); kd = V(t[0], t[1], t[2]); }
    const v3 ka = V(m->ka[0], m->ka[1], m->ka[2]);
    if (m->kind == MIRO_GPU_MAT_LAMBERT) {                                    /* Lambert.cpp:19-53 */
        v3 L = V(0, 0, 0);
        for (uint32_t li = 0; li < s->n_lights; ++li) {
            float discard;
            L = add(L, mul(sample_light(c, li, sf.P, sf.N, ray->time, V(0, 0, 0), &discard, 0, 0, &addr), kd));
        }
        return add(L, ka);
    }
    /* Blinn.cpp:91-335.  Texture maps, Blinn.cpp:120-142: the normal map is applied with the texel as stored. */
    v3 sN = sf.N;
    float spec_amt = m->spec_amt, reflect_amt = m->reflect_amt, refract_amt = m->refract_amt;
    if (m->normal_map >= 0) { float t[4]; tex_lookup(&s->textures[m->normal_map], sf.u, sf.v, t); sN = add(add(scl(sf.T, t[0]), scl(sf.BT, t[1])), scl(sf.N, t[2])); }
    if (m->specular_map >= 0) { float t[4]; tex_lookup(&s->textures[m->specular_map], sf.u, sf.v, t); spec_amt = (t[0] + t[1] + t[2]) * 0.3333333f * spec_amt; }
    if (m->reflect_map >= 0) { float t[4]; tex_lookup(&s->textures[m->reflect_map], sf.u, sf.v, t); reflect_amt = (t[0] + t[1] + t[2]) * 0.3333333f * reflect_amt; }
    if (m->refract_map >= 0) { float t[4]; tex_lookup(&s->textures[m->refract_map], sf.u, sf.v, t); refract_amt = (t[0] + t[1] + t[2]) * 0.3333333f * refract_amt; }
    const v3 viewDir = scl(ray->d, -1.f);
    float vDotN = dot(viewDir, sN);
    const float vDotGeoN = dot(viewDir, sf.geoN);
    const int nEqGeoN = (vDotN * vDotGeoN >= 0.0f);
    v3 theNormal = nEqGeoN ? sN : sf.geoN;
    vDotN = nEqGeoN ? vDotN : vDotGeoN;
    int flip = 0;
    if (vDotN < 0.0f) { flip = 1; vDotN = -vDotN; theNormal = scl(theNormal, -1.f); }
    v3 rVec = add(ray->d, scl(theNormal, 2.f * vDotN));
    if (m->spec_gloss < 1.0f) {                                               /* Blinn.cpp:160-165 */
        float r[4]; rand4(&addr, RP_GLOSS, 0, 0, 0, 0, r);
        const v3 randD = cosine_sample(theNormal, r[0], r[1]);
        rVec = normalize(add(scl(rVec, m->spec_gloss), scl(randD, 1.f - m->spec_gloss)));
    }
    const float inIOR = ior_top(&ray->ior);                                   /* Blinn.cpp:167-186 */
    const int dispersive = m->disperse && !ray->is_refract;
    float outIOR;
    if (dispersive) outIOR = m->ior[0];
    else if (flip) { ior_pop(&ray->ior); outIOR = ior_top(&ray->ior); } else outIOR = m->ior[1];
    float Rs = 0.f, Ts = 0.f;
    if (m->reflect_amt > 0.0f || m->refract_amt > 0.0f) { Rs = fresnel(inIOR, outIOR, vDotN); Ts = 1.0f - Rs; }
    float rr[4]; rand4(&addr, RP_ROULETTE, 0, 0, 0, 0, rr);
    const float rrWeight = 1.0f - Rs * reflect_amt - Ts * refract_amt;
    const float rrWeightRecip = (rrWeight > 0.f) ? 1.f / rrWeight : 1.f;
    const float rrWeightRecipSpec = (1.f - rrWeight > 0.f) ? 1.f / (1.f - rrWeight) : 1.f;
    const v3 Le = V(m->le[0], m->le[1], m->le[2]), ks = V(m->ks[0], m->ks[1], m->ks[2]);
    v3 Ld = V(0, 0, 0), Ls = V(0, 0, 0), Lr = V(0, 0, 0), Lt = V(0, 0, 0);
    if (rr[0] <= rrWeight) {
        if (c->p->path_trace) {                                               /* Blinn::calculatePathTracing, Blinn.cpp:39-89 */
            if (m->emit_intensity > 0.0f || (Le.x + Le.y + Le.z) > 0.0f) Ld = add(Ld, scl(Le, m->emit_intensity));
            else if (depth < c->p->max_bounces - 1) {
                float r[4]; rand4(&addr, RP_COSINE, 0, 0, 0, 0, r);
                oray2 nr; nr.o = sf.P; nr.d = cosine_sample(theNormal, r[0], r[1]); nr.time = ray->time; nr.bounces = ray->bounces; nr.is_refract = 0;
                ior_init(&nr.ior); ior_push(&nr.ior, 1.001f);                  /* Ray randRay(threadID) */
                ior_pop(&nr.ior); ior_push(&nr.ior, ior_top(&ray->ior));       /* randRay.set(..., ray.r_IOR(), ...) */
                miro_gpu_hit nh;
                if (trace(c, nr.o, nr.d, nr.time, O_EPS, MIRO_GPU_TMAX, &nh)) Ld = add(Ld, mul(kd, shade(c, &nr, &nh, pixel, sample, path, depth + 1, 1)));
                else if (m->sample_env && c->p->sample_env) Ld = add(Ld, mul(kd, environment(c, nr.d)));
            } else {
                for (uint32_t li = 0; li < s->n_lights; ++li) {
                    float lightSpec;
                    Ld = add(Ld, mul(sample_light(c, li, sf.P, theNormal, ray->time, V(0, 0, 0), &lightSpec, 1, 1, &addr), kd));
                }
            }
        }
        for (uint32_t li = 0; li < s->n_lights; ++li) {
            float lightSpec = 0.f;
            const v3 lightPower = sample_light(c, li, sf.P, theNormal, ray->time, rVec, &lightSpec, isSecondary, 0, &addr);
            if (spec_amt != 0.f) Ls = add(Ls, scl(mul(lightPower, ks), spec_amt * powf(lightSpec, m->spec_exp)));
            Ld = add(Ld, mul(lightPower, kd));
        }
        if (m->translucency > 0.01f) {                                        /* Blinn.cpp:223-236: lights seen from the back side */
            v3 lightTotal = V(0, 0, 0);
            for (uint32_t li = 0; li < s->n_lights; ++li) {
                float lightSpec = 0.f;
                lightTotal = add(lightTotal, sample_light(c, li, sf.P, scl(theNormal, -1.f), .001f, rVec, &lightSpec, isSecondary, 2, &addr));
            }
            Ld = add(Ld, mul(scl(lightTotal, m->translucency), kd));          /* "translucency" shares Ld's 1/rrWeight */
        }
    } else {
        int doEnv = 1;
        if (rr[1] < reflect_amt * Rs) {                                    /* Blinn.cpp:247-268 */
            if (reflect_amt * Rs > 0.0f && ray->bounces < 5) {
                oray2 nr; nr.o = sf.P; nr.d = rVec; nr.time = ray->time; nr.ior = ray->ior; nr.bounces = ray->bounces + 1; nr.is_refract = 0;
                miro_gpu_hit nh;
                if (trace(c, nr.o, nr.d, nr.time, O_EPS, MIRO_GPU_TMAX, &nh)) { Lr = add(Lr, mul(ks, shade(c, &nr, &nh, pixel, sample, path, depth, 0))); doEnv = 0; }
            }
            if (reflect_amt * Rs > 0.0f && doEnv) Lr = add(Lr, mul(ks, environment(c, rVec)));
        } else if (refract_amt * Ts > 0.0f && dispersive) {                /* Blinn.cpp:275-302: one ray per colour channel */
            v3 tVec = V(0, 0, 0);
            for (int i = 0; i < 3; i++) {
                const float snellsQ = inIOR / m->ior[i];
                const float sq = sqrtf(1.0f - (snellsQ * snellsQ) * (1.0f - vDotN * vDotN));
                const float sqrtPart = (0.0f < sq) ? sq : 0.0f;
                tVec = normalize(add(scl(ray->d, snellsQ), scl(theNormal, snellsQ * vDotN - sqrtPart)));
                if (ray->bounces < 5) {
                    ior_push(&ray->ior, m->ior[i]);
                    oray2 nr; nr.o = sf.P; nr.d = tVec; nr.time = ray->time; nr.ior = ray->ior; nr.bounces = ray->bounces + 1; nr.is_refract = 1;
                    miro_gpu_hit nh;
                    if (trace(c, nr.o, nr.d, nr.time, O_EPS, MIRO_GPU_TMAX, &nh)) {
                        const v3 refraction = shade(c, &nr, &nh, pixel, sample, path, depth, 0);
                        const v3 mask = V(i == 0 ? 1.f : 0.f, i == 1 ? 1.f : 0.f, i == 2 ? 1.f : 0.f);
                        Lt = add(Lt, mul(ks, mul(refraction, mask)));
                        doEnv = 0;
                    }
                    ior_pop(&ray->ior);
                }
            }
            if (doEnv) Lt = add(Lt, mul(ks, environment(c, tVec)));
        } else if (refract_amt * Ts > 0.0f) {                              /* Blinn.cpp:303-322, no dispersion */
            const float snellsQ = inIOR / outIOR;
            const float sq = sqrtf(1.0f - (snellsQ * snellsQ) * (1.0f - vDotN * vDotN));
            const float sqrtPart = (0.0f < sq) ? sq : 0.0f;
            const v3 tVec = normalize(add(scl(ray->d, snellsQ), scl(theNormal, snellsQ * vDotN - sqrtPart)));
            if (ray->bounces < 5) {
                ior_push(&ray->ior, outIOR);
                oray2 nr; nr.o = sf.P; nr.d = tVec; nr.time = ray->time; nr.ior = ray->ior; nr.bounces = ray->bounces + 1; nr.is_refract = 1;
                miro_gpu_hit nh;
                if (trace(c, nr.o, nr.d, nr.time, O_EPS, MIRO_GPU_TMAX, &nh)) { Lt = add(Lt, mul(ks, shade(c, &nr, &nh, pixel, sample, path, depth, 0))); doEnv = 0; }
                ior_pop(&ray->ior);
            }
            if (doEnv) Lt = add(Lt, mul(ks, environment(c, tVec)));
        }
    }
    Ld = add(Ld, ka);
    return add(add(scl(add(Ld, Ls), rrWeightRecip), scl(add(Lr, Lt), rrWeightRecipSpec)), Le);
}

typedef struct { v3 eye, u, v, w; float top, right, focus, aperture, shutter; } ocam;

static oray2 eye_ray(const ocam* cm, int x, int y, float minX, float maxX, float minY, float maxY, int W, int H, const raddr* addr) {
    float r[4]; rand4(addr, RP_CAMERA, 0, 0, 0, 0, r);
    const float xOffset = (maxX - minX) * r[0] + minX, yOffset = (maxY - minY) * r[1] + minY;
    const float left = -cm->right, bottom = -cm->top;
    const float U = left + (cm->right - left) * (((float)x + xOffset) / (float)W);
    const float Vp = bottom + (cm->top - bottom) * (((float)y + yOffset) / (float)H);
    oray2 o; o.time = 1.f - r[2] * r[2] * r[2] * cm->shutter; o.bounces = 0; o.is_refract = 0;
    ior_init(&o.ior); ior_push(&o.ior, 1.001f);                              /* Ray(threadID, o, d, t): IOR = 1.001 pushed, src/Ray.h:70-101 */
    const v3 dir = normalize(sub(add(scl(cm->u, U), scl(cm->v, Vp)), cm->w));
    if (cm->aperture < O_EPS) { o.o = cm->eye; o.d = dir; return o; }
    const v3 focal = add(scl(dir, cm->focus), cm->eye);
    float lu = 0.f, lv = 0.f;
    for (uint32_t attempt = 0; attempt < 64; ++attempt) {
        float q[4]; rand4(addr, RP_LENS, 0, 0, 0, attempt, q);
        lu = 1.0f - 2.f * q[0]; lv = 1.0f - 2.f * q[1]; if (lu * lu + lv * lv <= 1.0f) break;
        lu = 1.0f - 2.f * q[2]; lv = 1.0f - 2.f * q[3]; if (lu * lu + lv * lv <= 1.0f) break;
    }
    o.o = add(scl(add(scl(cm->u, lu), scl(cm->v, lv)), cm->aperture), cm->eye);
    o.d = normalize(sub(focal, o.o));
    return o;
}

static v3 sample_scene(octx* c, oray2* ray, uint32_t pixel, uint32_t sample) {      /* Scene.cpp:219-243 */
    miro_gpu_hit h;
    if (trace(c, ray->o, ray->d, ray->time, O_EPS, MIRO_GPU_TMAX, &h)) {
        v3 result = V(0, 0, 0);
        for (int i = 0; i < c->p->num_paths; i++) result = add(result, scl(shade(c, ray, &h, pixel, sample, (uint32_t)i, 0, 0), 1.0f / (float)c->p->num_paths));
        return result;
    }
    return environment(c, ray->d);
}

static int get_sum(int n) { return (int)(n * (n + 1) * (2 * n + 1) * 0.16666667f); }

/* Scene::raytraceImage / adaptiveSampleScene over the pixels of one shard.  rgb: width*height*3, row 0 = bottom.
 * Returns the number of Scene::trace calls.  pixel_mask (optional, width*height bytes): render only pixels with mask != 0. */
uint64_t oracle_render(const miro_gpu_scene_desc* s, const miro_gpu_camera* cam, const miro_gpu_render_params* p, float* rgb, const uint8_t* pixel_mask) {
    const int W = p->width, H = p->height;
    static float lut[32769]; static int lut_ready = 0;
    if (!lut_ready) { const float GAMMA = 2.2f; for (int i = 0; i < 32769; i++) lut[i] = (float)(powf(i / 32768.0f, 1 / GAMMA) * 255.0 + 0.5); lut_ready = 1; }
    ocam cm;
    {
        const v3 vd = V(cam->view_dir[0], cam->view_dir[1], cam->view_dir[2]), up = V(cam->up[0], cam->up[1], cam->up[2]);
        cm.w = normalize(scl(vd, -1.f)); cm.u = normalize(cross(up, cm.w)); cm.v = cross(cm.w, cm.u);
        cm.eye = V(cam->eye[0], cam->eye[1], cam->eye[2]);
        cm.top = tanf(cam->fov_deg * (O_PI / 360.0f)); cm.right = ((float)W / (float)H) * cm.top;
        cm.focus = cam->focus_plane; cm.aperture = cam->aperture; cm.shutter = cam->shutter_speed;
    }
    dome* domes = (dome*)calloc(s->n_lights ? s->n_lights : 1, sizeof(dome));
    for (uint32_t i = 0; i < s->n_lights; ++i) if (s->lights[i].kind == MIRO_GPU_LIGHT_DOME) dome_init(&domes[i], &s->textures[s->lights[i].texture]);
    const int sc = p->shard_count > 1 ? p->shard_count : 1, si = p->shard_index, nbx = (W + 31) / 32;
    uint64_t total_rays = 0;
    #pragma omp parallel for schedule(dynamic, 16) reduction(+ : total_rays)
    for (int y = 0; y < H; ++y) {
        octx c; c.s = s; c.p = p; c.domes = domes; c.rays = 0;
        for (int x = 0; x < W; ++x) {
            const uint32_t pixel = (uint32_t)(y * W + x);
            if (((y / 32) * nbx + x / 32) % sc != si) continue;
            if (pixel_mask && !pixel_mask[pixel]) continue;
            raddr addr; addr.pixel = pixel; addr.sample = 0; addr.path_depth = 0; addr.seed = p->seed;
            oray2 ray = eye_ray(&cm, x, y, 0.5f, 0.5f, 0.5f, 0.5f, W, H, &addr);
            v3 shadeResult = sample_scene(&c, &ray, pixel, 0);
            int curLevel = 2, cutOff = 0;
            while ((curLevel <= p->max_subdivs && !cutOff) || curLevel <= p->min_subdivs) {         /* Scene.cpp:259-290 */
                v3 curResult = V(0, 0, 0);
                for (int i = 0; i < curLevel; i++) for (int j = 0; j < curLevel; j++) {
                    const float offset = 1.0f / (float)curLevel;
                    addr.sample = (uint32_t)(get_sum(curLevel - 1) + i * curLevel + j);
                    ray = eye_ray(&cm, x, y, i * offset, (i + 1) * offset, j * offset, (j + 1) * offset, W, H, &addr);
                    curResult = add(curResult, sample_scene(&c, &ray, pixel, addr.sample));
                }
                const float pre = (float)get_sum(curLevel - 1), now = (float)(curLevel * curLevel);
                const v3 newResult = scl(add(scl(shadeResult, pre), curResult), 1.0f / (pre + now));
                #define G(v) lut[(int)((((v) > 1.f) ? 1.f : ((v) < 0.f ? 0.f : (v))) * 32767.f)]
                const float tx = fabsf(G(shadeResult.x) - G(newResult.x)), ty = fabsf(G(shadeResult.y) - G(newResult.y)), tz = fabsf(G(shadeResult.z) - G(newResult.z));
                #undef G
                float mx = tx > ty ? tx : ty; mx = mx > tz ? mx : tz;
                cutOff = mx < p->noise_threshold;
                shadeResult = newResult;
                curLevel++;
            }
            rgb[(size_t)pixel * 3] = shadeResult.x; rgb[(size_t)pixel * 3 + 1] = shadeResult.y; rgb[(size_t)pixel * 3 + 2] = shadeResult.z;
        }
        total_rays += c.rays;
    }
    for (uint32_t i = 0; i < s->n_lights; ++i) if (s->lights[i].kind == MIRO_GPU_LIGHT_DOME) dome_free(&domes[i]);
    free(domes);
    return total_rays;
}
