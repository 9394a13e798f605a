// render.cu — Scene::raytraceImage on the GPU: a wavefront renderer over device-resident ray queues.
//
// Replaces (reference, file:line relative to src/):
//   Scene::raytraceImage          Scene.cpp:86-217    pixel loop over 32x32 buckets
//   Scene::adaptiveSampleScene    Scene.cpp:252-293   level-k stratified supersampling with a gamma-space cut-off
//   Scene::sampleScene            Scene.cpp:219-243   trace the camera ray, mean of numPaths shade() calls, miss -> env / BG
//   Camera::eyeRayAdaptive        Camera.cpp:116-174
//   Lambert::shade                Lambert.cpp:19-53
//   Blinn::shade (diffuse, highlight, translucency, reflection / refraction / dispersion with Fresnel-weighted Russian
//   roulette), Blinn::calculatePathTracing   Blinn.cpp:39-335
//   Point/Rectangle/DomeLight::sampleLight   PointLight.cpp:8-82, RectangleLight.cpp:42-137, DomeLight.cpp:80-161
//
// The reference's shade() <-> trace() recursion becomes a loop over path depth; all state lives in queues in HBM:
//
//   per subdivision level k (host loop, one 4-byte read-back per level for the number of still-active pixels):
//     per wave of <= W camera samples (W * numPaths <= wave capacity):
//       k_raygen        camera samples of the active pixels                      -> camera-sample queue (48 B rays)
//       k_trace closest                                                          -> 20 B hits
//       k_shade<PRIMARY> one thread per (camera sample, path): the numPaths shade() calls of sampleScene share the hit
//       for depth = 0 .. maxBounces-1:
//         (k_shade emitted) shadow rays -> k_trace shadow: any-hit traversal whose epilogue adds the unoccluded
//                           sample's irradiance / specular input into the accumulator of its light loop ("slot")
//         k_resolve_slots   per light loop: mean over its samples, kd / ks*pow(spec, specExp) weighting -> pixel sum
//         (k_shade emitted) bounce rays -> k_trace closest -> k_shade<BOUNCE>
//     k_level_resolve   running mean over levels, gamma-space cut-off, compaction of the pixels that go on to level k+1
//
// Queue slots are claimed with warp-aggregated atomics (one atomicAdd per warp per queue): every thread first COUNTS
// what it will emit (the light loops are pure functions of the counter-based RNG), the warp scans the counts, then
// every thread EMITS at its offset.  Pixel sums are float atomics (red.global.add.f32).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "shading.cuh"

namespace miro {

constexpr int SHADE_BLOCK = 128;
#ifndef MIRO_SHADE_MIN_BLOCKS
#define MIRO_SHADE_MIN_BLOCKS 4
#endif
// ray.flags of a queued continuation ray: path 0-15 | giBounces 16-23 | bounces 24-26 | FLAG_SECONDARY | FLAG_SAMPLE_ENV
constexpr uint32_t FLAG_SAMPLE_ENV = 0x80000000u;      // the environment / background is added when the ray leaves the scene
constexpr uint32_t FLAG_SECONDARY = 0x08000000u;       // shade(..., isSecondary = true): reached through calculatePathTracing
constexpr uint32_t FLAG_REFRACT = 0x10000000u;         // IS_REFRACT_RAY (src/Ray.h:15): a dispersive material does not split such a ray again
// bits 29-30: 1 + colour channel of a dispersion ray (Blinn.cpp:275-302); its throughput is masked to that channel when it is shaded
constexpr size_t WAVE_PATHS_MAX = (size_t)1 << 22;     // paths in flight per wave
constexpr size_t WAVE_BYTES_BUDGET = (size_t)12 << 30; // queue memory per context (two queue sets)
constexpr size_t WAVE_SPLIT_MIN = (size_t)1 << 17;     // a frame is cut into two overlapping waves only if each keeps this many paths

struct Slot {            // one light loop (one Light::sampleLight call of the reference): 64 bytes
    float4 acc;          // sum over unoccluded samples: E.rgb, specular input
    float4 tkd;          // throughput * diffuse colour (rgb); w = pixel index (bits)
    float4 tks;          // throughput * ks * specAmt (rgb); w = specExp
    float4 misc;         // x = 1 / samplesDone
};

struct RenderParamsDev {
    int width, height;
    int num_paths, max_bounces;
    uint32_t path_trace, sample_env;
    uint64_t seed;
    float inv_paths;
    uint32_t has_specular;      // some material has reflect_amt / refract_amt > 0: rays carry an IOR history
    uint32_t next_cap;          // capacity of the continuation-ray queues (3x the wave when a material disperses)
    int local_paths;            // paths per camera sample traced by THIS call (sample sharding), path = k * path_stride + path_first
    int path_first, path_stride;
    uint32_t full_shadows;      // some light uses the "full" shadow method (Light::setFastShadows(false)): its shadow rays go to the walk queue
};

struct Queues {
    // camera samples of the wave
    miro_gpu_ray* cs_rays; miro_gpu_hit* cs_hits;
    // bounce queues (ping-pong): rays, throughput, hits
    miro_gpu_ray* q_rays[2]; float4* q_thr[2]; miro_gpu_hit* q_hits;
    float4* q_ior[2];      // 2 x float4 per ray: the IOR history (only allocated when a material reflects / refracts)
    // shadow queue and light-loop slots
    miro_gpu_ray* sh_rays; float4* sh_E; Slot* slots;
    // shadow rays of lights with the "full" shadow method (walked hit by hit, k_walk_shadows); allocated only when such a light exists
    miro_gpu_ray* ws_rays; float4* ws_E;
    // device counters: [0] next bounce count, [1] shadow count, [2] slot count, [3] next active-pixel count, [4] bounce count being
    // traced, [5] dropped continuation rays, [6] walked-shadow count, [7] spare
    uint32_t* counts;
    size_t cap_cs, cap_paths, cap_shadow, cap_slots;
};

struct RenderState {
    // Two queue sets: consecutive waves run on two streams, so the tail of one wave's traversal launches (its last warps
    // walking their longest rays alone, ~100 us per launch, ~90 launches per wave) is filled by the other wave's kernels.
    Queues q[2]{};
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::vector<void*> allocs;
    size_t key_paths = 0, key_shadow_per_path = 0, key_slots_per_path = 0, key_next_mult = 1; bool key_ior = false, key_walk = false;
    // frame buffers
    float4* level_sum = nullptr; float4* result = nullptr; uint32_t* active[2] = {nullptr, nullptr};
    float* rgb_dev = nullptr; unsigned char* rgb8_dev = nullptr; size_t frame_pixels = 0;
    float* gamma_lut = nullptr;
    uint32_t* h_count = nullptr;       // pinned
    uint32_t* bucket_first = nullptr;  // per owned bucket: its first position in the shard's pixel list (+ the total at the end)
    size_t bucket_cap = 0;
};

// The shard's pixel list in the reference's bucket order (Scene.cpp:160-175: buckets row-major, pixels row-major inside a bucket):
// one block per owned bucket, its place in the list from a prefix sum over bucket sizes the host makes (a few thousand entries).
// The frame's pixel list used to be built and uploaded by the host: 8 MB and ~10 ms per 1080p frame, longer than the frame's kernels.
__global__ void __launch_bounds__(1024)
k_own_pixels(const uint32_t* __restrict__ bucket_first, int shard_index, int shard_count, int nbx, int W, int H, uint32_t* __restrict__ out) {
    const int b = shard_index + (int)blockIdx.x * shard_count;
    const int bx = b % nbx, by = b / nbx;
    const int bw = min(32, W - bx * 32), bh = min(32, H - by * 32);
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    if (lx < bw && ly < bh) out[bucket_first[blockIdx.x] + (uint32_t)(ly * bw + lx)] = (uint32_t)((by * 32 + ly) * W + bx * 32 + lx);
}

__device__ __forceinline__ void add_rgb(float4* buf, uint32_t pixel, float3x c) {
    float* p = reinterpret_cast<float*>(buf + pixel);
    if (c.x != 0.f) atomicAdd(p, c.x);
    if (c.y != 0.f) atomicAdd(p + 1, c.y);
    if (c.z != 0.f) atomicAdd(p + 2, c.z);
}

// ---------------------------------------------------------------------------------------------
// Camera samples.  Sample s of level k in pixel p: sub-cell (i, j) = (s / k, s % k), ordinal = getSum(k-1) + s.
__global__ void __launch_bounds__(SHADE_BLOCK)
k_raygen(DeviceCamera cam, RenderParamsDev P, const uint32_t* __restrict__ active, uint32_t first_cs, uint32_t n_cs, int level, uint32_t ordinal_base,
         miro_gpu_ray* __restrict__ rays) {
    const uint32_t i = blockIdx.x * SHADE_BLOCK + threadIdx.x;
    if (i >= n_cs) return;
    const uint32_t cs = first_cs + i;
    const uint32_t k2 = (uint32_t)(level * level);
    const uint32_t a = cs / k2, s = cs - a * k2;
    const uint32_t pixel = active ? __ldg(active + a) : a;      // no list: every pixel of the frame, row-major
    const int x = (int)(pixel % (uint32_t)P.width), y = (int)(pixel / (uint32_t)P.width);
    float minX = 0.5f, maxX = 0.5f, minY = 0.5f, maxY = 0.5f;         // level 1: the pixel centre (Scene.cpp:254)
    if (level > 1) {
        const int si = (int)(s / (uint32_t)level), sj = (int)(s - (uint32_t)si * (uint32_t)level);
        const float offset = 1.0f / (float)level;                     // Scene.cpp:267-268
        minX = si * offset; maxX = (si + 1) * offset; minY = sj * offset; maxY = (sj + 1) * offset;
    }
    RandAddr addr; addr.pixel = pixel; addr.sample = ordinal_base + s; addr.path_depth = 0; addr.seed = P.seed;
    const CameraSample c = camera_ray(cam, x, y, minX, maxX, minY, maxY, P.width, P.height, addr);
    float4* o = reinterpret_cast<float4*>(rays + i);
    o[0] = make_float4(c.o.x, c.o.y, c.o.z, kEps);
    o[1] = make_float4(c.d.x, c.d.y, c.d.z, MIRO_GPU_TMAX);
    o[2] = make_float4(c.time, 0.f, __uint_as_float(pixel), __uint_as_float(ordinal_base + s));
}

// ---------------------------------------------------------------------------------------------
struct ShadeCtx {
    float3x P, N, rVec, kd, tks;      // hit point, shading normal, reflection vector, diffuse colour, throughput*ks*specAmt
    float spec_exp;
    float time;
    bool is_secondary;
};

// Runs every light loop of one shading event; F(light, pass, normal, rVec, is_secondary, with_spec) is called once per loop.
template <class F>
__device__ __forceinline__ void for_each_light_loop(const DeviceShading& sh, bool pt_last_bounce, bool blinn, bool secondary, bool translucent, F&& f) {
    if (pt_last_bounce)                                  // Blinn::calculatePathTracing, last bounce (Blinn.cpp:76-87): rVec = 0, isSecondary = true
        for (uint32_t li = 0; li < sh.n_lights; ++li) f(li, 1u, true, false);
    for (uint32_t li = 0; li < sh.n_lights; ++li)        // Lambert.cpp:41-46 (isSecondary defaults to false) / Blinn.cpp:212-221
        f(li, 0u, blinn ? secondary : false, blinn);
    if (translucent)                                     // Blinn.cpp:223-236: the lights seen from the back side (normal -N, time .001)
        for (uint32_t li = 0; li < sh.n_lights; ++li) f(li, 2u, secondary, false);
}

// Light::m_fastShadows == false: rectangle (when it casts shadows at all) and dome lights walk their shadow rays hit by hit;
// a point light's loop `sampleHit.t = distance; while (sampleHit.t < distance)` never runs (PointLight.cpp:39,52) — it casts no shadow.
__device__ __forceinline__ bool light_walks(const miro_gpu_light& l) {
    return l.full_shadows && (l.kind == MIRO_GPU_LIGHT_DOME || (l.kind == MIRO_GPU_LIGHT_RECT && l.cast_shadows));
}

template <bool PRIMARY>
__global__ void __launch_bounds__(SHADE_BLOCK, MIRO_SHADE_MIN_BLOCKS)
k_shade(DeviceScene sc, DeviceShading sh, RenderParamsDev P, Queues q, int in_q, uint32_t n_static, const uint32_t* __restrict__ d_count,
        float4* __restrict__ level_sum, uint32_t shadow_cap, uint32_t slot_cap) {
    const uint32_t n = PRIMARY ? n_static : min(*d_count, P.next_cap);
    const uint32_t stride = gridDim.x * SHADE_BLOCK;
    for (uint32_t base = blockIdx.x * SHADE_BLOCK + (threadIdx.x & ~31u); base < n; base += stride) {
        const uint32_t idx = base + (threadIdx.x & 31u);
        // ------------------------------------------------------------------ phase 1: evaluate, count
        bool active = idx < n;
        uint32_t pixel = 0, sample = 0, path = 0, gi = 0, bounces = 0, vertex = 0;
        float3x thr = f3(0, 0, 0), thr_d = f3(0, 0, 0), o = f3(0, 0, 0), d = f3(0, 0, 1), bounce_dir = f3(0, 0, 0), bounce_thr = f3(0, 0, 0);
        float time = 0.f;
        bool emit_bounce = false, pt_last = false, blinn = false, diffuse = true, translucent = false, disperse_split = false;
        float disp_in = 1.f, disp_vdn = 0.f; uint32_t disp_mat = 0;
        float transl = 0.f;
        uint32_t bounce_flags = 0;
        IorStack ior; ior.init_camera();
        ShadeCtx c{};
        int n_shadow = 0, n_slots = 0, n_walk = 0;
        RandAddr addr{};
        if (active) {
            const uint32_t ri = PRIMARY ? idx / (uint32_t)P.local_paths : idx;
            const float4* rp = reinterpret_cast<const float4*>((PRIMARY ? q.cs_rays : q.q_rays[in_q]) + ri);
            const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
            const miro_gpu_hit* hp = (PRIMARY ? q.cs_hits : q.q_hits) + ri;
            const float ht = __ldg(&hp->t), ha = __ldg(&hp->a), hb = __ldg(&hp->b);
            const int hprim = __ldg(&hp->prim), hinst = __ldg(&hp->inst);
            o = f3(r0.x, r0.y, r0.z); d = f3(r1.x, r1.y, r1.z); time = r2.x;
            pixel = __float_as_uint(r2.z); sample = __float_as_uint(r2.w);
            const uint32_t flags = __float_as_uint(r2.y);
            bool secondary = false;
            uint32_t channel = 0;
            if (PRIMARY) { path = (idx - ri * (uint32_t)P.local_paths) * (uint32_t)P.path_stride + (uint32_t)P.path_first; thr = f3(P.inv_paths, P.inv_paths, P.inv_paths); }
            else {
                path = flags & 0xffffu; gi = (flags >> 16) & 0xffu; bounces = (flags >> 24) & 7u; secondary = (flags & FLAG_SECONDARY) != 0;
                const float4 t4 = __ldg(q.q_thr[in_q] + idx); thr = f3(t4.x, t4.y, t4.z);
                channel = (flags >> 29) & 3u;
                if (P.has_specular) {
                    const float4 i0 = __ldg(q.q_ior[in_q] + 2 * (size_t)idx), i1 = __ldg(q.q_ior[in_q] + 2 * (size_t)idx + 1);
                    ior.v[0] = i0.x; ior.v[1] = i0.y; ior.v[2] = i0.z; ior.v[3] = i0.w; ior.v[4] = i1.x; ior.v[5] = i1.y; ior.v[6] = i1.z;
                    ior.idx = (int)__float_as_uint(i1.w);
                }
            }
            vertex = gi + bounces;      // ordinal of this vertex along its path: every continuation adds one to gi or to bounces
            addr.pixel = pixel; addr.sample = sample; addr.path_depth = path | (vertex << 16); addr.seed = P.seed;
            if (hprim < 0) {
                // Scene::sampleScene miss (Scene.cpp:234-240): env / BG once per camera sample; continuation-ray miss
                // (Blinn.cpp:70-73 with the material's and the scene's sampleEnv; Blinn.cpp:262,326 unconditionally)
                if (PRIMARY) { if (path == 0) add_rgb(level_sum, pixel, environment(sh, d)); }
                else if (channel) {
                    // dispersion rays: the environment is added ONCE, unmasked, along the last (blue) ray, and only when all
                    // three rays of the split left the scene (Blinn.cpp:281-302,324-327); the three are queue neighbours
                    if (channel == 3u && __ldg(&q.q_hits[idx - 1].prim) < 0 && __ldg(&q.q_hits[idx - 2].prim) < 0) add_rgb(level_sum, pixel, thr * environment(sh, d));
                } else if (flags & FLAG_SAMPLE_ENV) add_rgb(level_sum, pixel, thr * environment(sh, d));
                active = false;
            } else {
                if (channel) thr = thr * f3(channel == 1u ? 1.f : 0.f, channel == 2u ? 1.f : 0.f, channel == 3u ? 1.f : 0.f);      // "refraction * mask"
                const Surface s = surface_at(sc, sh, o, d, ht, ha, hb, hprim, hinst);
                const miro_gpu_material* m = sh.materials + s.material;
                float3x kd = f3(m->kd[0], m->kd[1], m->kd[2]);
                if (m->color_map >= 0) { const float4 t = tex_lookup(sh.textures[m->color_map], s.u, s.v); kd = f3(t.x, t.y, t.z); }
                const float3x ka = f3(m->ka[0], m->ka[1], m->ka[2]);
                c.P = s.P; c.kd = kd; c.time = time; c.is_secondary = secondary; c.spec_exp = m->spec_exp;
                blinn = m->kind == MIRO_GPU_MAT_BLINN;
                thr_d = thr;
                if (!blinn) { c.N = s.N; c.rVec = f3(0, 0, 0); c.tks = f3(0, 0, 0); add_rgb(level_sum, pixel, thr * ka); }
                else {
                    // texture maps of Blinn::shade (Blinn.cpp:120-142): the normal map perturbs N in the tangent frame with the
                    // texel as stored (no remap to [-1,1], no renormalisation); the others scale an amount by the texel's mean RGB
                    float3x sN = s.N;
                    float spec_amt = m->spec_amt, reflect_amt = m->reflect_amt, refract_amt = m->refract_amt;
                    if (m->normal_map >= 0) { const float4 t = tex_lookup(sh.textures[m->normal_map], s.u, s.v); sN = t.x * s.T + t.y * s.BT + t.z * s.N; }
                    if (m->specular_map >= 0) { const float4 t = tex_lookup(sh.textures[m->specular_map], s.u, s.v); spec_amt = (t.x + t.y + t.z) * 0.3333333f * spec_amt; }
                    if (m->reflect_map >= 0) { const float4 t = tex_lookup(sh.textures[m->reflect_map], s.u, s.v); reflect_amt = (t.x + t.y + t.z) * 0.3333333f * reflect_amt; }
                    if (m->refract_map >= 0) { const float4 t = tex_lookup(sh.textures[m->refract_map], s.u, s.v); refract_amt = (t.x + t.y + t.z) * 0.3333333f * refract_amt; }
                    // normal selection / flip towards the viewer (Blinn.cpp:144-155)
                    const float3x viewDir = -d;
                    float vDotN = dot3(viewDir, sN);
                    const float vDotGeoN = dot3(viewDir, s.geoN);
                    const bool nEqGeoN = (vDotN * vDotGeoN >= 0.0f);
                    float3x theNormal = nEqGeoN ? sN : s.geoN;
                    vDotN = nEqGeoN ? vDotN : vDotGeoN;
                    bool flip = false;
                    if (vDotN < 0.0f) { flip = true; vDotN = -vDotN; theNormal = -theNormal; }
                    c.N = theNormal;
                    float3x rVec = d + (2.f * vDotN) * theNormal;                                 // Blinn.cpp:158
                    if (m->spec_gloss < 1.0f) {                                                  // Blinn.cpp:160-165
                        const Rand4 r = rand4(addr, RP_GLOSS, 0, 0, 0, 0);
                        const float3x randD = cosine_sample(theNormal, r.x, r.y);
                        rVec = normalize3(m->spec_gloss * rVec + (1.f - m->spec_gloss) * randD);
                    }
                    c.rVec = rVec;
                    // IOR bookkeeping (Blinn.cpp:167-186).  The reference pops the history of the ray OBJECT it was handed;
                    // sampleScene shades the same camera ray numPaths times, so path i of a back-facing primary hit sees the
                    // history already popped by paths 0..i-1.
                    const bool dispersive = m->disperse && !(flags & FLAG_REFRACT);            // Blinn.cpp:169
                    if (PRIMARY && flip && !dispersive) for (uint32_t k = 0; k < path && k < 2u; ++k) ior.pop();
                    const float inIOR = ior.top();
                    float outIOR;
                    if (dispersive) outIOR = m->ior[0];                                        // no pop on this branch
                    else if (flip) { ior.pop(); outIOR = ior.top(); } else outIOR = m->ior[1];
                    float Rs = 0.f, Ts = 0.f;
                    if (m->reflect_amt > 0.0f || m->refract_amt > 0.0f) { Rs = fresnel(inIOR, outIOR, vDotN); Ts = 1.0f - Rs; }      // the members, not the mapped amounts (Blinn.cpp:189)
                    const Rand4 rr = rand4(addr, RP_ROULETTE, 0, 0, 0, 0);
                    const float rrWeight = 1.0f - Rs * reflect_amt - Ts * refract_amt;       // Blinn.cpp:195-198
                    const float rrWeightRecip = (rrWeight > 0.f) ? 1.f / rrWeight : 1.f;
                    const float rrWeightRecipSpec = (1.f - rrWeight > 0.f) ? 1.f / (1.f - rrWeight) : 1.f;
                    const float3x ks = f3(m->ks[0], m->ks[1], m->ks[2]);
                    const float3x Le = f3(m->le[0], m->le[1], m->le[2]);
                    diffuse = rr.x <= rrWeight;
                    transl = m->translucency; translucent = diffuse && transl > 0.01f;
                    thr_d = thr * rrWeightRecip;                       // (Ld + Ls) / rrWeight, Blinn.cpp:335
                    c.tks = thr_d * ks * spec_amt;
                    float3x constant = thr_d * ka + thr * Le;          // "Ld += m_ka" is on both branches; "+ m_Le" is unscaled
                    if (diffuse) {
                        if (P.path_trace) {                                                      // Blinn::calculatePathTracing
                            if (m->emit_intensity > 0.0f || (Le.x + Le.y + Le.z) > 0.0f) constant = constant + thr_d * (m->emit_intensity * Le);
                            else if ((int)gi < P.max_bounces - 1) {
                                const Rand4 r = rand4(addr, RP_COSINE, 0, 0, 0, 0);
                                bounce_dir = cosine_sample(theNormal, r.x, r.y);
                                bounce_thr = thr_d * kd;
                                bounce_flags = (path | ((gi + 1u) << 16) | (bounces << 24)) | FLAG_SECONDARY | ((m->sample_env && P.sample_env) ? FLAG_SAMPLE_ENV : 0u);
                                // randRay.set(..., ray.r_IOR(), ...) on a fresh Ray: history = {1.0, top}  (Blinn.cpp:61, Ray.h:143-176)
                                const float top = ior.top(); ior.init_camera(); ior.v[1] = top;
                                emit_bounce = true;
                            } else pt_last = true;
                        }
                    } else {
                        // mirror reflection or refraction, Blinn.cpp:238-331; one continuation ray, weight ks / (1 - rrWeight)
                        float3x dir;
                        bool spawn = false;
                        if (rr.y < reflect_amt * Rs) {
                            if (reflect_amt * Rs > 0.0f) { dir = rVec; spawn = true; }
                        } else if (refract_amt * Ts > 0.0f && dispersive) {
                            // one refraction ray per colour channel, each with its own IOR (Blinn.cpp:275-302)
                            bounce_thr = thr * ks * rrWeightRecipSpec;
                            if (bounces < 5u) {
                                disperse_split = true; disp_in = inIOR; disp_vdn = vDotN; disp_mat = s.material;
                                bounce_flags = (path | (gi << 16) | ((bounces + 1u) << 24)) | FLAG_REFRACT;
                            } else {
                                const float snellsQ = inIOR / m->ior[2];
                                const float sqrtPart = fmaxf(0.0f, sqrtf(1.0f - (snellsQ * snellsQ) * (1.0f - vDotN * vDotN)));
                                constant = constant + bounce_thr * environment(sh, normalize3(snellsQ * d + theNormal * (snellsQ * vDotN - sqrtPart)));
                            }
                        } else if (refract_amt * Ts > 0.0f) {
                            const float snellsQ = inIOR / outIOR;
                            const float sqrtPart = fmaxf(0.0f, sqrtf(1.0f - (snellsQ * snellsQ) * (1.0f - vDotN * vDotN)));
                            dir = normalize3(snellsQ * d + theNormal * (snellsQ * vDotN - sqrtPart));
                            ior.push(outIOR);
                            spawn = true; bounce_flags = FLAG_REFRACT;
                        }
                        if (spawn) {
                            bounce_thr = thr * ks * rrWeightRecipSpec;
                            if (bounces < 5u) {
                                bounce_dir = dir; emit_bounce = true;
                                bounce_flags = (bounce_flags & FLAG_REFRACT) | (path | (gi << 16) | ((bounces + 1u) << 24)) | FLAG_SAMPLE_ENV;   // shade(...) with isSecondary = false
                            } else constant = constant + bounce_thr * environment(sh, dir);       // "doEnv": no further bounce
                        }
                    }
                    add_rgb(level_sum, pixel, constant);
                }
                if (diffuse) for_each_light_loop(sh, pt_last, blinn, c.is_secondary, translucent, [&](uint32_t li, uint32_t pass, bool secondary_, bool with_spec) {
                    int lit = 0;
                    light_loop(sh, li, c.P, pass == 2u ? -c.N : c.N, with_spec ? c.rVec : f3(0, 0, 0), secondary_, pass, addr, [&](const LightSample&) { ++lit; });
                    if (lit) { if (P.full_shadows && light_walks(sh.lights[li])) n_walk += lit; else n_shadow += lit; ++n_slots; }
                });
            }
        }
        // ------------------------------------------------------------------ phase 2: claim queue space (warp aggregated)
        const uint32_t lane = threadIdx.x & 31u;
        const uint32_t n_next = disperse_split ? 3u : (emit_bounce ? 1u : 0u);
        uint32_t s_shadow = (uint32_t)n_shadow, s_slots = (uint32_t)n_slots, s_next = n_next;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, s_shadow, off), b = __shfl_up_sync(0xffffffffu, s_slots, off), e = __shfl_up_sync(0xffffffffu, s_next, off);
            if ((int)lane >= off) { s_shadow += a; s_slots += b; s_next += e; }
        }
        uint32_t b_shadow = 0, b_slots = 0, b_next = 0;
        if (lane == 31) {
            if (s_next) b_next = atomicAdd(q.counts + 0, s_next);
            if (s_shadow) b_shadow = atomicAdd(q.counts + 1, s_shadow);
            if (s_slots) b_slots = atomicAdd(q.counts + 2, s_slots);
        }
        b_shadow = __shfl_sync(0xffffffffu, b_shadow, 31) + s_shadow - (uint32_t)n_shadow;
        b_slots = __shfl_sync(0xffffffffu, b_slots, 31) + s_slots - (uint32_t)n_slots;
        b_next = __shfl_sync(0xffffffffu, b_next, 31) + s_next - n_next;
        uint32_t b_walk = 0;
        if (P.full_shadows) {          // uniform: the walk queue exists only in scenes with such a light
            uint32_t s_walk = (uint32_t)n_walk;
            for (int off = 1; off < 32; off <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, s_walk, off); if ((int)lane >= off) s_walk += a; }
            if (lane == 31 && s_walk) b_walk = atomicAdd(q.counts + 6, s_walk);
            b_walk = __shfl_sync(0xffffffffu, b_walk, 31) + s_walk - (uint32_t)n_walk;
        }
        if (!active) continue;
        // ------------------------------------------------------------------ phase 3: emit
        if (n_next && b_next + n_next > P.next_cap) { atomicAdd(q.counts + 5, n_next); }      // queue full: reported as an error by the host
        else if (disperse_split) {
            const miro_gpu_material* m = sh.materials + disp_mat;
            for (uint32_t ch = 0; ch < 3u; ++ch) {
                const float oi = m->ior[ch];
                const float snellsQ = disp_in / oi;
                const float sqrtPart = fmaxf(0.0f, sqrtf(1.0f - (snellsQ * snellsQ) * (1.0f - disp_vdn * disp_vdn)));
                const float3x tv = normalize3(snellsQ * d + c.N * (snellsQ * disp_vdn - sqrtPart));
                IorStack st2 = ior; st2.push(oi);
                const uint32_t k = b_next + ch;
                float4* o4 = reinterpret_cast<float4*>(q.q_rays[in_q ^ 1] + k);
                o4[0] = make_float4(c.P.x, c.P.y, c.P.z, kEps);
                o4[1] = make_float4(tv.x, tv.y, tv.z, MIRO_GPU_TMAX);
                o4[2] = make_float4(time, __uint_as_float(bounce_flags | ((ch + 1u) << 29)), __uint_as_float(pixel), __uint_as_float(sample));
                q.q_thr[in_q ^ 1][k] = make_float4(bounce_thr.x, bounce_thr.y, bounce_thr.z, 0.f);     // unmasked: masked when shaded
                q.q_ior[in_q ^ 1][2 * (size_t)k] = make_float4(st2.v[0], st2.v[1], st2.v[2], st2.v[3]);
                q.q_ior[in_q ^ 1][2 * (size_t)k + 1] = make_float4(st2.v[4], st2.v[5], st2.v[6], __uint_as_float((uint32_t)st2.idx));
            }
        } else if (emit_bounce) {
            float4* o4 = reinterpret_cast<float4*>(q.q_rays[in_q ^ 1] + b_next);
            o4[0] = make_float4(c.P.x, c.P.y, c.P.z, kEps);
            o4[1] = make_float4(bounce_dir.x, bounce_dir.y, bounce_dir.z, MIRO_GPU_TMAX);
            o4[2] = make_float4(time, __uint_as_float(bounce_flags), __uint_as_float(pixel), __uint_as_float(sample));
            q.q_thr[in_q ^ 1][b_next] = make_float4(bounce_thr.x, bounce_thr.y, bounce_thr.z, 0.f);
            if (P.has_specular) {
                q.q_ior[in_q ^ 1][2 * (size_t)b_next] = make_float4(ior.v[0], ior.v[1], ior.v[2], ior.v[3]);
                q.q_ior[in_q ^ 1][2 * (size_t)b_next + 1] = make_float4(ior.v[4], ior.v[5], ior.v[6], __uint_as_float((uint32_t)ior.idx));
            }
        }
        if (n_slots == 0) continue;
        if (b_shadow + (uint32_t)n_shadow > shadow_cap || b_slots + (uint32_t)n_slots > slot_cap || b_walk + (uint32_t)n_walk > shadow_cap) continue;   // cannot happen: capacities are worst case
        uint32_t w_shadow = b_shadow, w_slot = b_slots, w_walk = b_walk;
        for_each_light_loop(sh, pt_last, blinn, c.is_secondary, translucent, [&](uint32_t li, uint32_t pass, bool secondary, bool with_spec) {
            const uint32_t slot = w_slot;
            int lit = 0;
            const bool walk = P.full_shadows && light_walks(sh.lights[li]);
            const bool shadows = sh.lights[li].cast_shadows != 0 && !sh.lights[li].full_shadows;      // full method without a walk: no shadow (see light_walks)
            const float ray_time = pass == 2u ? .001f : c.time;           // the translucency loop passes .001f as the time (Blinn.cpp:232)
            const int done = light_loop(sh, li, c.P, pass == 2u ? -c.N : c.N, with_spec ? c.rVec : f3(0, 0, 0), secondary, pass, addr, [&](const LightSample& ls) {
                if (walk) {
                    float4* o4 = reinterpret_cast<float4*>(q.ws_rays + w_walk);
                    __stcs(o4 + 0, make_float4(c.P.x, c.P.y, c.P.z, ls.tmin));
                    __stcs(o4 + 1, make_float4(ls.dir.x, ls.dir.y, ls.dir.z, ls.tmax));
                    __stcs(o4 + 2, make_float4(ray_time, 0.f, __uint_as_float(slot), ls.dist));
                    __stcs(q.ws_E + w_walk, make_float4(ls.E.x, ls.E.y, ls.E.z, ls.spec));
                    ++w_walk; ++lit;
                    return;
                }
                float4* o4 = reinterpret_cast<float4*>(q.sh_rays + w_shadow);
                // a light that casts no shadows gets an empty interval: never occluded
                __stcs(o4 + 0, make_float4(c.P.x, c.P.y, c.P.z, shadows ? ls.tmin : 1.f));
                __stcs(o4 + 1, make_float4(ls.dir.x, ls.dir.y, ls.dir.z, shadows ? ls.tmax : 0.f));
                __stcs(o4 + 2, make_float4(ray_time, 0.f, __uint_as_float(slot), 0.f));
                __stcs(q.sh_E + w_shadow, make_float4(ls.E.x, ls.E.y, ls.E.z, ls.spec));
                ++w_shadow; ++lit;
            });
            if (lit) {
                Slot* sp = q.slots + slot;
                const float3x tkd = pass == 2u ? thr_d * c.kd * transl : thr_d * c.kd;
                sp->acc = make_float4(0.f, 0.f, 0.f, 0.f);
                sp->tkd = make_float4(tkd.x, tkd.y, tkd.z, __uint_as_float(pixel));
                sp->tks = with_spec ? make_float4(c.tks.x, c.tks.y, c.tks.z, c.spec_exp) : make_float4(0.f, 0.f, 0.f, 1.f);
                sp->misc = make_float4(1.0f / (float)done, 0.f, 0.f, 0.f);
                ++w_slot;
            }
        });
    }
}

// The "full" shadow method (Light::setFastShadows(false); RectangleLight.cpp:93-118, DomeLight.cpp:123-146): one thread walks one
// shadow ray hit by hit.  A surface whose INTERPOLATED normal faces the ray (HitInfo::getInterpolatedNormal, Ray.cpp:52-66: object
// space, not transformed by a proxy) multiplies the visibility by its material's refractAmt; the walk ends when the light is
// reached, nothing is hit, or the visibility falls to epsilon.  Kept as the reference has it: sampleHit is never reset, so the
// previous segment's hit distance is the next segment's tMax.  The sample then enters its light loop's accumulator scaled by
// the visibility, as the any-hit kernel's epilogue does with 0 / 1.
template <bool ALPHA>
__global__ void __launch_bounds__(TRACE_BLOCK)
k_walk_shadows(DeviceScene sc, DeviceShading sh, const miro_gpu_ray* __restrict__ rays, const float4* __restrict__ sample_E,
               const uint32_t* __restrict__ d_count, uint32_t cap, float4* __restrict__ slots, TraceCounters* __restrict__ ctr) {
    __shared__ unsigned long long stack[SMEM_STACK * TRACE_BLOCK];
    unsigned long long overflow[LMEM_STACK];
    TraversalStack st; st.init(stack + threadIdx.x, overflow, LMEM_STACK);
    const uint32_t n = min(*d_count, cap);
    for (uint32_t i = blockIdx.x * TRACE_BLOCK + threadIdx.x; i < n; i += gridDim.x * TRACE_BLOCK) {
        const float4* rp = reinterpret_cast<const float4*>(rays + i);
        const float4 r0 = __ldcs(rp), r1 = __ldcs(rp + 1), r2 = __ldcs(rp + 2);
        float3x o = f3(r0.x, r0.y, r0.z);
        const float3x d = f3(r1.x, r1.y, r1.z);
        const float distance = r2.w;
        float attenuate = 1.0f, traversed = 0.0f;
        uint32_t segments = 0;
        Lane L; L.ray_idx = i; L.tmin = r0.w; L.time = r2.x; L.hit.t = r1.w;
        for (int guard = 0; traversed < distance && attenuate > kEps && guard < 4096; ++guard) {
            trace_closest_thread<ALPHA>(sc, L, st, o.x, o.y, o.z, d.x, d.y, d.z);      // tMax = L.hit.t, carried over from the previous segment
            ++segments;
            if (L.hit.prim < 0) { traversed = distance; break; }
            const miro_gpu_prim* pr = sh.prims + L.hit.prim;
            const float a = L.hit.a, b = L.hit.b, cc = 1.0f - a - b;
            const float* n0 = sh.normals + (size_t)__ldg(&pr->n[0]) * 3; const float* n1 = sh.normals + (size_t)__ldg(&pr->n[1]) * 3; const float* n2 = sh.normals + (size_t)__ldg(&pr->n[2]) * 3;
            const float3x N = normalize3(f3(__ldg(n0) * cc + __ldg(n1) * a + __ldg(n2) * b, __ldg(n0 + 1) * cc + __ldg(n1 + 1) * a + __ldg(n2 + 1) * b,
                                            __ldg(n0 + 2) * cc + __ldg(n1 + 2) * a + __ldg(n2 + 2) * b));
            if (dot3(N, -d) > 0.0f) attenuate *= sh.materials[__ldg(&pr->material)].refract_amt;
            o = o + L.hit.t * d; traversed += L.hit.t;
        }
        if (segments) atomicAdd(&ctr->rays_closest, (unsigned long long)segments);      // every segment is one Scene::trace call
        if (attenuate != 0.0f) {
            const float4 E = __ldcs(sample_E + i);
            atomicAdd(slots + (size_t)__float_as_uint(r2.z) * 4, make_float4(E.x * attenuate, E.y * attenuate, E.z * attenuate, E.w * attenuate));
        }
    }
}

// One light loop's contribution (tail of the sampleLight functions + Blinn.cpp:217-220 / Lambert.cpp:45)
__global__ void __launch_bounds__(SHADE_BLOCK)
k_resolve_slots(const Slot* __restrict__ slots, const uint32_t* __restrict__ d_count, float4* __restrict__ level_sum) {
    const uint32_t n = *d_count;
    for (uint32_t i = blockIdx.x * SHADE_BLOCK + threadIdx.x; i < n; i += gridDim.x * SHADE_BLOCK) {
        const float4* s = reinterpret_cast<const float4*>(slots + i);
        const float4 acc = s[0], tkd = s[1], tks = s[2], misc = s[3];
        const float3x E = f3(acc.x, acc.y, acc.z) * misc.x;
        float3x out = E * f3(tkd.x, tkd.y, tkd.z);
        if (tks.x != 0.f || tks.y != 0.f || tks.z != 0.f) {
            const float spec = powf(acc.w * misc.x, tks.w);
            out = out + E * f3(tks.x, tks.y, tks.z) * spec;
        }
        add_rgb(level_sum, __float_as_uint(tkd.w), out);
    }
}

// Scene::adaptiveSampleScene's level bookkeeping (Scene.cpp:259-290) for the pixels that were sampled at `level`.
__global__ void __launch_bounds__(SHADE_BLOCK)
k_level_resolve(const uint32_t* __restrict__ active, uint32_t n_active, int level, int min_subdivs, int max_subdivs, float noise,
                const float* __restrict__ gamma_lut, float4* __restrict__ level_sum, float4* __restrict__ result,
                uint32_t* __restrict__ next_active, uint32_t* __restrict__ next_count) {
    const uint32_t i = blockIdx.x * SHADE_BLOCK + threadIdx.x;
    bool go_on = false;
    uint32_t pixel = 0;
    if (i < n_active) {
        pixel = active[i];
        const float4 cur = level_sum[pixel];
        level_sum[pixel] = make_float4(0.f, 0.f, 0.f, 0.f);
        bool cutOff = false;
        float4 nr;
        if (level == 1) nr = cur;
        else {
            const float4 old = result[pixel];
            const int km1 = level - 1;
            const float pre = (float)(int)(km1 * (km1 + 1) * (2 * km1 + 1) * 0.16666667f);     // getSum, Scene.cpp:245-248
            const float now = (float)(level * level);
            const float w = 1.0f / (pre + now);
            nr = make_float4((old.x * pre + cur.x) * w, (old.y * pre + cur.y) * w, (old.z * pre + cur.z) * w, 0.f);
            auto g = [&](float v) { return __ldg(gamma_lut + (int)(((v > 1.f) ? 1.f : (v < 0.f ? 0.f : v)) * 32767.f)); };
            const float tx = fabsf(g(old.x) - g(nr.x)), ty = fabsf(g(old.y) - g(nr.y)), tz = fabsf(g(old.z) - g(nr.z));
            cutOff = fmaxf(tx, fmaxf(ty, tz)) < noise;
        }
        result[pixel] = nr;
        const int next = level + 1;
        go_on = (next <= max_subdivs && !cutOff) || next <= min_subdivs;                         // Scene.cpp:259
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, go_on);
    if (ballot) {
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(next_count, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (go_on) next_active[base + __popc(ballot & ((1u << lane) - 1u))] = pixel;
    }
}

// Image::setPixel's tone mapping (Map, src/Image.cpp:71-76): clamp, index the 32 769-entry 2.2-gamma table, truncate to a byte.
__device__ __forceinline__ unsigned char map_byte(const float* __restrict__ lut, float r) {
    const float m = 32768.0f * r;
    const int linear = (m > 32768.0f) ? 32768 : (int)(m > 0.f ? m : 0.f);
    return (unsigned char)(int)__ldg(lut + linear);
}

// The shard's pixels leave the accumulation buffer: float radiance (rgb, may be NULL) and / or the 8-bit image the reference's
// Image holds (rgb8, may be NULL) — mapping 2 M pixels through Image::setPixel on the host cost more than rendering them.
__global__ void k_write_rgb(const uint32_t* __restrict__ pixels, uint32_t n, const float4* __restrict__ result, float* __restrict__ rgb,
                            unsigned char* __restrict__ rgb8, const float* __restrict__ lut) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = pixels[i];
    const float4 r = result[p];
    if (rgb) { rgb[(size_t)p * 3 + 0] = r.x; rgb[(size_t)p * 3 + 1] = r.y; rgb[(size_t)p * 3 + 2] = r.z; }
    if (rgb8) { rgb8[(size_t)p * 3 + 0] = map_byte(lut, r.x); rgb8[(size_t)p * 3 + 1] = map_byte(lut, r.y); rgb8[(size_t)p * 3 + 2] = map_byte(lut, r.z); }
}

__global__ void k_map_frame(const float* __restrict__ rgb, size_t n, unsigned char* __restrict__ rgb8, const float* __restrict__ lut) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) rgb8[i] = map_byte(lut, rgb[i]);
}

// ---------------------------------------------------------------------------------------------
static RenderState* state_of(miro_gpu_ctx* ctx) {
    if (!ctx->render_state) ctx->render_state = new RenderState();
    return static_cast<RenderState*>(ctx->render_state);
}

static void free_queues(RenderState* st) {
    for (void* p : st->allocs) cudaFree(p);
    st->allocs.clear();
    st->q[0] = Queues{}; st->q[1] = Queues{};
    st->key_paths = 0;
}

void render_state_free(miro_gpu_ctx* ctx) {
    if (!ctx->render_state) return;
    RenderState* st = static_cast<RenderState*>(ctx->render_state);
    free_queues(st);
    cudaFree(st->level_sum); cudaFree(st->result); cudaFree(st->active[0]); cudaFree(st->active[1]); cudaFree(st->rgb_dev); cudaFree(st->rgb8_dev); cudaFree(st->gamma_lut); cudaFree(st->bucket_first);
    if (st->h_count) cudaFreeHost(st->h_count);
    if (st->aux) cudaStreamDestroy(st->aux);
    if (st->ev_fork) cudaEventDestroy(st->ev_fork);
    if (st->ev_join) cudaEventDestroy(st->ev_join);
    delete st;
    ctx->render_state = nullptr;
}

template <class T>
static cudaError_t qalloc(RenderState* st, T** p, size_t n) {
    void* v = nullptr;
    cudaError_t e = cudaMalloc(&v, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) { st->allocs.push_back(v); *p = static_cast<T*>(v); }
    return e;
}

static int ensure_queues(miro_gpu_ctx* ctx, RenderState* st, size_t paths, size_t cs, size_t shadow_per_path, size_t slots_per_path, bool with_ior, size_t next_mult, bool with_walk) {
    if (st->key_walk == with_walk && st->key_next_mult == next_mult && st->key_paths == paths && st->q[0].cap_cs >= cs && st->key_shadow_per_path == shadow_per_path && st->key_slots_per_path == slots_per_path && st->key_ior == with_ior) return MIRO_GPU_OK;
    free_queues(st);
    for (int set = 0; set < 2; ++set) {
    Queues& q = st->q[set];
    // a vertex of ANY continuation ray can run every light loop, so the shadow / slot queues scale with the continuation queue
    q.cap_cs = cs; q.cap_paths = paths * next_mult; q.cap_shadow = q.cap_paths * shadow_per_path; q.cap_slots = q.cap_paths * slots_per_path;
    MIRO_CUDA(ctx, qalloc(st, &q.cs_rays, q.cap_cs));
    MIRO_CUDA(ctx, qalloc(st, &q.cs_hits, q.cap_cs));
    for (int i = 0; i < 2; ++i) {
        MIRO_CUDA(ctx, qalloc(st, &q.q_rays[i], q.cap_paths)); MIRO_CUDA(ctx, qalloc(st, &q.q_thr[i], q.cap_paths));
        if (with_ior) MIRO_CUDA(ctx, qalloc(st, &q.q_ior[i], 2 * q.cap_paths));
    }
    MIRO_CUDA(ctx, qalloc(st, &q.q_hits, q.cap_paths));
    MIRO_CUDA(ctx, qalloc(st, &q.sh_rays, q.cap_shadow));
    MIRO_CUDA(ctx, qalloc(st, &q.sh_E, q.cap_shadow));
    if (with_walk) { MIRO_CUDA(ctx, qalloc(st, &q.ws_rays, q.cap_shadow)); MIRO_CUDA(ctx, qalloc(st, &q.ws_E, q.cap_shadow)); }
    MIRO_CUDA(ctx, qalloc(st, &q.slots, q.cap_slots));
    MIRO_CUDA(ctx, qalloc(st, &q.counts, (size_t)8));
    }
    st->key_next_mult = next_mult; st->key_paths = paths; st->key_shadow_per_path = shadow_per_path; st->key_slots_per_path = slots_per_path; st->key_ior = with_ior; st->key_walk = with_walk;
    return MIRO_GPU_OK;
}

static int ensure_frame(miro_gpu_ctx* ctx, RenderState* st, size_t pixels) {
    if (!st->gamma_lut) {
        // Image::generateGammaTables, Image.cpp:19-35: linear_to_gammaF[i] = pow(i/32768, 1/2.2) * 255 + 0.5
        std::vector<float> lut(32769);
        const float GAMMA = 2.2f;
        for (int i = 0; i < 32769; i++) lut[i] = (float)(powf(i / 32768.0f, 1 / GAMMA) * 255.0 + 0.5);
        MIRO_CUDA(ctx, cudaMalloc((void**)&st->gamma_lut, lut.size() * sizeof(float)));
        MIRO_CUDA(ctx, cudaMemcpy(st->gamma_lut, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice));
        MIRO_CUDA(ctx, cudaMallocHost((void**)&st->h_count, 8 * sizeof(uint32_t)));
        MIRO_CUDA(ctx, cudaStreamCreateWithFlags(&st->aux, cudaStreamNonBlocking));
        MIRO_CUDA(ctx, cudaEventCreateWithFlags(&st->ev_fork, cudaEventDisableTiming));
        MIRO_CUDA(ctx, cudaEventCreateWithFlags(&st->ev_join, cudaEventDisableTiming));
    }
    if (st->frame_pixels >= pixels) return MIRO_GPU_OK;
    cudaFree(st->level_sum); cudaFree(st->result); cudaFree(st->active[0]); cudaFree(st->active[1]); cudaFree(st->rgb_dev); cudaFree(st->rgb8_dev);
    st->level_sum = st->result = nullptr; st->active[0] = st->active[1] = nullptr; st->rgb_dev = nullptr; st->rgb8_dev = nullptr; st->frame_pixels = 0;
    MIRO_CUDA(ctx, cudaMalloc((void**)&st->level_sum, pixels * sizeof(float4)));
    MIRO_CUDA(ctx, cudaMalloc((void**)&st->result, pixels * sizeof(float4)));
    MIRO_CUDA(ctx, cudaMalloc((void**)&st->active[0], pixels * sizeof(uint32_t)));
    MIRO_CUDA(ctx, cudaMalloc((void**)&st->active[1], pixels * sizeof(uint32_t)));
    MIRO_CUDA(ctx, cudaMalloc((void**)&st->rgb_dev, pixels * 3 * sizeof(float)));
    MIRO_CUDA(ctx, cudaMalloc((void**)&st->rgb8_dev, pixels * 3));
    st->frame_pixels = pixels;
    return MIRO_GPU_OK;
}

static inline int grid_for(size_t n, int block) { return (int)std::min<size_t>((n + block - 1) / block, 0x7fffffff); }


}  // namespace miro

using namespace miro;

// camera basis (Camera.cpp:123-137) — computed on the host once per frame
static DeviceCamera make_device_camera(const miro_gpu_camera* cam, int W, int H) {
    DeviceCamera dc;
    auto norm = [](float3x a) { const float l = 1.0f / sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); return f3(a.x * l, a.y * l, a.z * l); };
    const float3x vd = f3(cam->view_dir[0], cam->view_dir[1], cam->view_dir[2]), up = f3(cam->up[0], cam->up[1], cam->up[2]);
    dc.w = norm(f3(-vd.x, -vd.y, -vd.z));
    dc.u = norm(cross3(up, dc.w));
    dc.v = cross3(dc.w, dc.u);
    dc.eye = f3(cam->eye[0], cam->eye[1], cam->eye[2]);
    dc.top = tanf(cam->fov_deg * (3.1415926f / 360.0f));
    dc.right = ((float)W / (float)H) * dc.top;
    dc.focus_plane = cam->focus_plane; dc.aperture = cam->aperture; dc.shutter = cam->shutter_speed;
    return dc;
}

// Primary rays made where they are traced: Camera::eyeRayAdaptive at the pixel centres (Camera.cpp:116-174; the level-1 sample of
// Scene::adaptiveSampleScene, Scene.cpp:254) generated on the device, traced, and only the 20-byte hit records travel — a third of
// what miro_gpu_trace_closest moves for the same rays (48 B up + 20 B down per ray over PCIe, which bounds the host-pointer
// calls).  Pixel order: row-major, row 0 = bottom (hits[y * width + x]); the rays are those miro_gpu_render traces at level 1
// with the same seed.  hits: host or device pointer.  rays_out (optional, device pointer or NULL): the generated rays.
extern "C" int miro_gpu_trace_primary(miro_gpu_ctx* ctx, const miro_gpu_camera* cam, int width, int height, uint64_t seed, miro_gpu_hit* hits, miro_gpu_ray* rays_out) {
    if (!ctx) return MIRO_GPU_EINVAL;
    if (!cam || !hits) return set_error(ctx, MIRO_GPU_EINVAL, "miro_gpu_trace_primary: NULL argument");
    if (!ctx->has_scene) return set_error(ctx, MIRO_GPU_ENOSCENE, "trace before upload_scene");
    if (width <= 0 || height <= 0 || (size_t)width * height > 0x7fffffffu) return set_error(ctx, MIRO_GPU_EINVAL, "bad image size");
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)width * height;
    cudaPointerAttributes attr;
    const bool out_is_device = cudaPointerGetAttributes(&attr, hits) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (!rays_out) { MIRO_CUDA(ctx, ctx->d_rays.reserve(n)); }
    miro_gpu_ray* d_rays = rays_out ? rays_out : ctx->d_rays.ptr;
    miro_gpu_hit* d_hits = hits;
    if (!out_is_device) { MIRO_CUDA(ctx, ctx->d_hits.reserve(n)); d_hits = ctx->d_hits.ptr; }
    if (!ctx->copy_out) MIRO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    const DeviceCamera dc = make_device_camera(cam, width, height);
    RenderParamsDev P{};
    P.width = width; P.height = height; P.num_paths = 1; P.seed = seed;
    // chunks of rows: the download of chunk k runs while chunk k + 1 is traced
    const size_t chunk = (size_t)1 << 18;
    const size_t n_chunks = (n + chunk - 1) / chunk;
    while (ctx->pipe_events.size() < 2 * n_chunks) { cudaEvent_t e; MIRO_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->pipe_events.push_back(e); }
    cudaStream_t s = ctx->stream;
    EventPair tot = begin_timing(ctx, false);
    for (size_t k = 0; k < n_chunks; ++k) {
        const size_t off = k * chunk, m = std::min(chunk, n - off);
        k_raygen<<<grid_for(m, SHADE_BLOCK), SHADE_BLOCK, 0, s>>>(dc, P, nullptr, (uint32_t)off, (uint32_t)m, 1, 0u, d_rays + off);
        ctx->launches++;
        EventPair p = begin_timing(ctx, true);
        launch_trace_closest(ctx, d_rays + off, m, nullptr, d_hits + off);
        end_timing(ctx, p);
        if (!out_is_device) {
            MIRO_CUDA(ctx, cudaEventRecord(ctx->pipe_events[k], s));
            MIRO_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->pipe_events[k], 0));
            MIRO_CUDA(ctx, cudaMemcpyAsync(hits + off, d_hits + off, m * sizeof(miro_gpu_hit), cudaMemcpyDeviceToHost, ctx->copy_out));
        }
    }
    end_timing(ctx, tot);
    MIRO_CUDA(ctx, cudaGetLastError());
    if (!out_is_device) MIRO_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
    MIRO_CUDA(ctx, cudaStreamSynchronize(s));
    return MIRO_GPU_OK;
}

namespace miro {
// Image::setPixel over a whole frame that already lies in device memory of ctx's GPU (the combined frame of a group)
int map_frame_to_bytes(miro_gpu_ctx* ctx, const float* d_rgb, size_t pixels, unsigned char* d_rgb8, cudaStream_t s) {
    RenderState* st = state_of(ctx);
    int rc;
    if ((rc = ensure_frame(ctx, st, 1))) return rc;      // the gamma table
    k_map_frame<<<ctx->sm_count * 8, 256, 0, s>>>(d_rgb, pixels * 3, d_rgb8, st->gamma_lut);
    ctx->launches++;
    return MIRO_GPU_OK;
}
}  // namespace miro

extern "C" int miro_gpu_render(miro_gpu_ctx* ctx, const miro_gpu_camera* cam, const miro_gpu_render_params* rp, float* rgb_out) {
    if (ctx && !rgb_out) return set_error(ctx, MIRO_GPU_EINVAL, "miro_gpu_render: NULL argument");
    return miro_gpu_render_image(ctx, cam, rp, rgb_out, nullptr);
}

extern "C" int miro_gpu_render_image(miro_gpu_ctx* ctx, const miro_gpu_camera* cam, const miro_gpu_render_params* rp, float* rgb_out, unsigned char* rgb8_out) {
    if (!ctx) return MIRO_GPU_EINVAL;
    if (!cam || !rp || (!rgb_out && !rgb8_out)) return set_error(ctx, MIRO_GPU_EINVAL, "miro_gpu_render: NULL argument");
    if (!ctx->has_scene) return set_error(ctx, MIRO_GPU_ENOSCENE, "render before upload_scene");
    if (rp->width <= 0 || rp->height <= 0 || (size_t)rp->width * rp->height > 0x7fffffffu) return set_error(ctx, MIRO_GPU_EINVAL, "bad image size");
    if (rp->num_paths < 1 || rp->num_paths > 0xffff) return set_error(ctx, MIRO_GPU_EINVAL, "num_paths must be in 1..65535");
    if (rp->max_bounces > 255) return set_error(ctx, MIRO_GPU_EINVAL, "max_bounces > 255 (the reference keeps the count in 8 bits, src/Ray.h:23)");
    const int max_sub = std::max(1, std::max(rp->max_subdivs, rp->min_subdivs));
    if (max_sub > 32) return set_error(ctx, MIRO_GPU_EINVAL, "more than 32 subdivision levels");
    if (ctx->shading.n_prims == 0 && (ctx->n_tris || ctx->n_mbtris)) return set_error(ctx, MIRO_GPU_EINVAL, "scene has no shading records (prims)");
    MIRO_CUDA(ctx, cudaSetDevice(ctx->device));
    RenderState* st = state_of(ctx);
    const int kPersistentGrid = ctx->sm_count * 8;      // the persistent shading / resolve kernels: 8 blocks per SM
    const int W = rp->width, H = rp->height;
    const size_t pixels = (size_t)W * H;
    int rc;
    if ((rc = ensure_frame(ctx, st, pixels))) return rc;

    const DeviceCamera dc = make_device_camera(cam, W, H);
    RenderParamsDev P;
    P.width = W; P.height = H; P.num_paths = rp->num_paths; P.max_bounces = rp->max_bounces;
    P.path_trace = rp->path_trace; P.sample_env = rp->sample_env; P.seed = rp->seed; P.inv_paths = 1.0f / (float)rp->num_paths;
    P.has_specular = 0;
    const int pcount = std::max(1, rp->path_shard_count), pindex = rp->path_shard_index;
    if (pindex < 0 || pindex >= pcount) return set_error(ctx, MIRO_GPU_EINVAL, "path_shard_index out of range");
    if (pcount > 1 && rp->min_subdivs != rp->max_subdivs) return set_error(ctx, MIRO_GPU_EINVAL, "sample sharding needs min_subdivs == max_subdivs (the adaptive cut-off needs the complete pixel value)");
    P.path_first = pindex; P.path_stride = pcount;
    P.local_paths = (rp->num_paths - pindex + pcount - 1) / pcount;       // paths p < num_paths with p % pcount == pindex
    bool any_disperse = false;
    for (const miro_gpu_material& m : ctx->host_materials) if (m.kind == MIRO_GPU_MAT_BLINN && (m.reflect_amt > 0.f || m.refract_amt > 0.f)) { P.has_specular = 1; if (m.disperse && m.refract_amt > 0.f) any_disperse = true; }
    const size_t next_mult = any_disperse ? 3 : 1;      // a dispersive refraction turns one path into three (Blinn.cpp:275-302)

    // ---- queue capacities: worst case per path
    size_t light_samples = 0;
    for (const miro_gpu_light& l : ctx->host_lights) light_samples += (size_t)std::max(1, l.num_samples);
    bool any_translucent = false;
    for (const miro_gpu_material& m : ctx->host_materials) if (m.kind == MIRO_GPU_MAT_BLINN && m.translucency > 0.01f) any_translucent = true;
    const size_t loops = (rp->path_trace ? 2 : 1) + (any_translucent ? 1 : 0);
    const size_t shadow_per_path = std::max<size_t>(1, loops * light_samples), slots_per_path = std::max<size_t>(1, loops * ctx->host_lights.size());
    P.full_shadows = 0;
    for (const miro_gpu_light& l : ctx->host_lights) if (l.full_shadows && (l.kind == MIRO_GPU_LIGHT_DOME || (l.kind == MIRO_GPU_LIGHT_RECT && l.cast_shadows))) P.full_shadows = 1;
    const size_t bytes_per_path = next_mult * (2 * (48 + 16 + (P.has_specular ? 32 : 0)) + 20 + shadow_per_path * 64 * (P.full_shadows ? 2 : 1) + slots_per_path * 64) + (48 + 20);
    size_t paths = std::min<size_t>(WAVE_PATHS_MAX, std::max<size_t>(WAVE_BYTES_BUDGET / 2 / bytes_per_path, (size_t)rp->num_paths));
    const size_t frame_paths = pixels * (size_t)max_sub * max_sub * rp->num_paths;
    paths = std::min(paths, frame_paths);
    paths = std::max<size_t>((paths / rp->num_paths) * rp->num_paths, (size_t)rp->num_paths);
    const size_t wave_cs = paths / rp->num_paths;
    if ((rc = ensure_queues(ctx, st, paths, wave_cs, shadow_per_path, slots_per_path, P.has_specular != 0, next_mult, P.full_shadows != 0))) return rc;
    P.next_cap = (uint32_t)(paths * next_mult);
    cudaStream_t s = ctx->stream;
    struct Restore { miro_gpu_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; c->work_lane = 0; } } restore{ctx, s};      // also on error returns

    // ---- pixels of this shard: 32x32 buckets in the order of Scene.cpp:160-175, bucket b owned when b % shard_count == shard_index
    const int shard_n = std::max(1, rp->shard_count), shard_i = rp->shard_index;
    if (shard_i < 0 || shard_i >= shard_n) return set_error(ctx, MIRO_GPU_EINVAL, "shard_index out of range");
    const int nbx = (W + 31) / 32, nby = (H + 31) / 32;
    std::vector<uint32_t> first(1, 0u);                // first[k]: position of owned bucket k's first pixel in the list; back(): the total
    for (int b = shard_i; b < nbx * nby; b += shard_n) {
        const int bx = b % nbx, by = b / nbx;
        first.push_back(first.back() + (uint32_t)(std::min(32, W - bx * 32) * std::min(32, H - by * 32)));
    }
    const uint32_t n_owned = (uint32_t)first.size() - 1, n_own = first.back();
    if (first.size() > st->bucket_cap) {
        cudaFree(st->bucket_first); st->bucket_first = nullptr; st->bucket_cap = 0;
        MIRO_CUDA(ctx, cudaMalloc((void**)&st->bucket_first, first.size() * sizeof(uint32_t)));
        st->bucket_cap = first.size();
    }
    uint32_t n_active = n_own;
    EventPair tot = begin_timing(ctx, false);
    MIRO_CUDA(ctx, cudaMemcpyAsync(st->bucket_first, first.data(), first.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    if (n_owned) { k_own_pixels<<<n_owned, 1024, 0, s>>>(st->bucket_first, shard_i, shard_n, nbx, W, H, st->active[0]); ctx->launches++; }
    MIRO_CUDA(ctx, cudaMemsetAsync(st->level_sum, 0, pixels * sizeof(float4), s));
    MIRO_CUDA(ctx, cudaMemsetAsync(st->q[0].counts + 5, 0, sizeof(uint32_t), s));
    MIRO_CUDA(ctx, cudaMemsetAsync(st->q[1].counts + 5, 0, sizeof(uint32_t), s));
    int cur = 0;
    // vertices along a path: up to maxBounces-1 diffuse (GI) continuations and up to 5 reflect / refract continuations (Blinn.cpp:57,247)
    const int last_depth = (rp->path_trace ? std::max(0, rp->max_bounces - 1) : 0) + (P.has_specular ? 5 : 0);
    for (int level = 1; level <= max_sub && n_active > 0; ++level) {
        const uint32_t k2 = (uint32_t)(level * level);
        const int km1 = level - 1;
        const uint32_t ordinal_base = (uint32_t)(km1 * (km1 + 1) * (2 * km1 + 1) / 6);
        const uint64_t total_cs = (uint64_t)n_active * k2;
        // waves alternate between the context's stream (queue set 0) and the auxiliary stream (set 1); both start after
        // everything enqueued so far (active list, previous level) and the level resolve waits for both
        // a level that fits one wave is still cut in two when both halves stay large, so that there is something to overlap
        static const bool overlap = getenv("MIRO_GPU_RENDER_SERIAL") == nullptr;      // A/B switch: one stream, one queue set
        uint64_t level_wave_cs = wave_cs;
        if (overlap && total_cs * (uint64_t)std::max(1, P.local_paths) >= 2 * WAVE_SPLIT_MIN) level_wave_cs = std::min<uint64_t>(level_wave_cs, (total_cs + 1) / 2);
        const bool none = P.local_paths == 0;      // sample sharding: this shard owns none of the num_paths paths (its frame is zero)
        const bool two = total_cs > level_wave_cs && !none && overlap;
        if (two) { MIRO_CUDA(ctx, cudaEventRecord(st->ev_fork, s)); MIRO_CUDA(ctx, cudaStreamWaitEvent(st->aux, st->ev_fork, 0)); }
        int wave = 0;
        for (uint64_t first = 0; first < total_cs && !none; first += level_wave_cs, ++wave) {
            const int set = two ? (wave & 1) : 0;
            Queues& q = st->q[set];
            cudaStream_t ws = set ? st->aux : s;
            ctx->stream = ws; ctx->work_lane = set;
            const uint32_t n_cs = (uint32_t)std::min<uint64_t>(level_wave_cs, total_cs - first);
            k_raygen<<<grid_for(n_cs, SHADE_BLOCK), SHADE_BLOCK, 0, ws>>>(dc, P, st->active[cur], (uint32_t)first, n_cs, level, ordinal_base, q.cs_rays);
            ctx->launches++;
            launch_trace_closest(ctx, q.cs_rays, n_cs, nullptr, q.cs_hits);
            MIRO_CUDA(ctx, cudaMemsetAsync(q.counts, 0, 3 * sizeof(uint32_t), ws));
            if (P.full_shadows) MIRO_CUDA(ctx, cudaMemsetAsync(q.counts + 6, 0, sizeof(uint32_t), ws));
            const size_t n_threads = (size_t)n_cs * P.local_paths;
            k_shade<true><<<std::min(grid_for(n_threads, SHADE_BLOCK), kPersistentGrid * 4), SHADE_BLOCK, 0, ws>>>(
                ctx->scene, ctx->shading, P, q, 0, (uint32_t)n_threads, nullptr, st->level_sum, (uint32_t)q.cap_shadow, (uint32_t)q.cap_slots);
            ctx->launches++;
            int in_q = 1;      // k_shade<PRIMARY> wrote its bounce rays to q_rays[0 ^ 1]
            for (int depth = 0; depth <= last_depth; ++depth) {
                // shadow rays of the vertices at `depth`, then the per-loop resolve
                launch_trace_shadow(ctx, q.sh_rays, q.cap_shadow, q.counts + 1, q.sh_E, reinterpret_cast<float4*>(q.slots));
                if (P.full_shadows) {
                    if (ctx->has_alpha) k_walk_shadows<true><<<kPersistentGrid, TRACE_BLOCK, 0, ws>>>(ctx->scene, ctx->shading, q.ws_rays, q.ws_E, q.counts + 6, (uint32_t)q.cap_shadow, reinterpret_cast<float4*>(q.slots), ctx->d_counters);
                    else k_walk_shadows<false><<<kPersistentGrid, TRACE_BLOCK, 0, ws>>>(ctx->scene, ctx->shading, q.ws_rays, q.ws_E, q.counts + 6, (uint32_t)q.cap_shadow, reinterpret_cast<float4*>(q.slots), ctx->d_counters);
                    ctx->launches++;
                }
                k_resolve_slots<<<kPersistentGrid, SHADE_BLOCK, 0, ws>>>(q.slots, q.counts + 2, st->level_sum);
                ctx->launches++;
                if (depth == last_depth) break;
                // bounce rays spawned at `depth` -> vertices at depth + 1
                MIRO_CUDA(ctx, cudaMemcpyAsync(q.counts + 4, q.counts + 0, sizeof(uint32_t), cudaMemcpyDeviceToDevice, ws));
                MIRO_CUDA(ctx, cudaMemsetAsync(q.counts, 0, 3 * sizeof(uint32_t), ws));
            if (P.full_shadows) MIRO_CUDA(ctx, cudaMemsetAsync(q.counts + 6, 0, sizeof(uint32_t), ws));
                launch_trace_closest(ctx, q.q_rays[in_q], q.cap_paths, q.counts + 4, q.q_hits);
                k_shade<false><<<kPersistentGrid * 4, SHADE_BLOCK, 0, ws>>>(ctx->scene, ctx->shading, P, q, in_q, 0, q.counts + 4, st->level_sum,
                                                                           (uint32_t)q.cap_shadow, (uint32_t)q.cap_slots);
                ctx->launches++;
                in_q ^= 1;
            }
            ctx->stream = s; ctx->work_lane = 0;
        }
        if (two) { MIRO_CUDA(ctx, cudaEventRecord(st->ev_join, st->aux)); MIRO_CUDA(ctx, cudaStreamWaitEvent(s, st->ev_join, 0)); }
        Queues& q = st->q[0];
        MIRO_CUDA(ctx, cudaMemsetAsync(q.counts + 3, 0, sizeof(uint32_t), s));
        k_level_resolve<<<grid_for(n_active, SHADE_BLOCK), SHADE_BLOCK, 0, s>>>(st->active[cur], n_active, level, rp->min_subdivs, rp->max_subdivs, rp->noise_threshold,
                                                                               st->gamma_lut, st->level_sum, st->result, st->active[cur ^ 1], q.counts + 3);
        ctx->launches++;
        if (level < max_sub) {
            MIRO_CUDA(ctx, cudaMemcpyAsync(st->h_count, q.counts + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            MIRO_CUDA(ctx, cudaStreamSynchronize(s));
            n_active = st->h_count[0];
            cur ^= 1;
        }
    }
    // ---- hand the shard's pixels back (row 0 = bottom); pixels of other shards are left untouched
    auto is_device = [](const void* p) { cudaPointerAttributes a; const bool d = p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice; cudaGetLastError(); return d; };
    const bool f_dev = is_device(rgb_out), b_dev = is_device(rgb8_out);
    float* target = rgb_out ? (f_dev ? rgb_out : st->rgb_dev) : nullptr;
    unsigned char* target8 = rgb8_out ? (b_dev ? rgb8_out : st->rgb8_dev) : nullptr;
    if (n_own) {
        k_own_pixels<<<n_owned, 1024, 0, s>>>(st->bucket_first, shard_i, shard_n, nbx, W, H, st->active[cur]);
        k_write_rgb<<<grid_for(n_own, 256), 256, 0, s>>>(st->active[cur], n_own, st->result, target, target8, st->gamma_lut);
        ctx->launches += 2;
    }
    MIRO_CUDA(ctx, cudaGetLastError());
    // host destinations: the whole frame in one copy, or — a shard of it — through a staging copy from which only the shard's
    // bucket rows reach the caller's frame
    auto download = [&](void* host, const void* dev, size_t bytes_per_pixel) -> int {
        if (n_own == pixels) { MIRO_CUDA(ctx, cudaMemcpyAsync(host, dev, pixels * bytes_per_pixel, cudaMemcpyDeviceToHost, s)); return MIRO_GPU_OK; }
        std::vector<unsigned char> tmp(pixels * bytes_per_pixel);
        MIRO_CUDA(ctx, cudaMemcpyAsync(tmp.data(), dev, tmp.size(), cudaMemcpyDeviceToHost, s));
        MIRO_CUDA(ctx, cudaStreamSynchronize(s));
        for (int b = shard_i; b < nbx * nby; b += shard_n) {
            const int bx = b % nbx, by = b / nbx;
            for (int y = by * 32; y < std::min((by + 1) * 32, H); ++y) {
                const size_t a = ((size_t)y * W + bx * 32) * bytes_per_pixel, n = (size_t)(std::min((bx + 1) * 32, W) - bx * 32) * bytes_per_pixel;
                memcpy(static_cast<unsigned char*>(host) + a, tmp.data() + a, n);
            }
        }
        return MIRO_GPU_OK;
    };
    if (rgb_out && !f_dev && (rc = download(rgb_out, st->rgb_dev, 3 * sizeof(float)))) return rc;
    if (rgb8_out && !b_dev && (rc = download(rgb8_out, st->rgb8_dev, 3))) return rc;
    if (any_disperse) {
        MIRO_CUDA(ctx, cudaMemcpyAsync(st->h_count + 1, st->q[0].counts + 5, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        MIRO_CUDA(ctx, cudaMemcpyAsync(st->h_count + 2, st->q[1].counts + 5, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    end_timing(ctx, tot);
    MIRO_CUDA(ctx, cudaStreamSynchronize(s));
    if (any_disperse && (st->h_count[1] || st->h_count[2]))
        return set_error(ctx, MIRO_GPU_ENOMEM, "continuation-ray queue overflow: " + std::to_string(st->h_count[1] + st->h_count[2]) + " dispersion rays were dropped (paths split more than 3x per wave)");
    return MIRO_GPU_OK;
}
