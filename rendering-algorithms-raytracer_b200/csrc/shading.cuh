// shading.cuh — device-side building blocks of the wavefront renderer: counter-based RNG, camera,
// texture lookups, hit-point attributes, light sampling.  Each block cites the reference code it replaces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "context.cuh"

namespace miro {

constexpr float kPI = 3.1415926f;            // src/Miro.h:57
constexpr float k1_PI = 1.0f / kPI;
constexpr float k1_4PI = 0.25f / kPI;
constexpr float kEps = MIRO_GPU_EPSILON;

// ---------------------------------------------------------------------------------------------
// RNG.  The reference pulls floats from one global MT19937 through per-thread 65 536-entry blocks
// (Scene::getRand, src/Scene.cpp:26-47), i.e. the stream a sample sees depends on thread scheduling.
// Here every random decision has an ADDRESS — (pixel, camera-sample ordinal, path, depth, purpose,
// light, pass, sample, attempt) — hashed by Philox4x32-10, so images do not depend on scheduling,
// sharding or wavefront order.  oracle/miro_oracle_shade.c uses the same addressing.
struct Rand4 { float x, y, z, w; };

__host__ __device__ inline uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
    const unsigned long long p = (unsigned long long)a * b;
    *hi = (uint32_t)(p >> 32);
    return (uint32_t)p;
}
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0, hi1;
        const uint32_t lo0 = mulhilo32(0xD2511F53u, c[0], &hi0);
        const uint32_t lo1 = mulhilo32(0xCD9E8D57u, c[2], &hi1);
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__host__ __device__ inline float u32_to_unit(uint32_t u) { return ((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)

enum RandPurpose : uint32_t { RP_CAMERA = 0, RP_LENS = 1, RP_COSINE = 2, RP_LIGHT = 3, RP_ROULETTE = 4, RP_GLOSS = 5 };

struct RandAddr {
    uint32_t pixel, sample, path_depth;      // path_depth = path | depth << 16
    uint64_t seed;
};
__host__ __device__ inline Rand4 rand4(const RandAddr& a, uint32_t purpose, uint32_t light, uint32_t pass, uint32_t sample, uint32_t attempt) {
    uint32_t c[4] = {a.pixel, a.sample, a.path_depth, (purpose << 28) | ((pass >> 1) << 27) | (light << 24) | ((pass & 1u) << 23) | ((sample & 0x7ffu) << 12) | (attempt & 0xfffu)};
    philox4x32_10(c, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    Rand4 r; r.x = u32_to_unit(c[0]); r.y = u32_to_unit(c[1]); r.z = u32_to_unit(c[2]); r.w = u32_to_unit(c[3]);
    return r;
}

// ---------------------------------------------------------------------------------------------
struct float3x { float x, y, z; };
__host__ __device__ inline float3x f3(float x, float y, float z) { float3x r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ inline float3x operator+(float3x a, float3x b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ inline float3x operator-(float3x a, float3x b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ inline float3x operator*(float3x a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ inline float3x operator*(float s, float3x a) { return f3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ inline float3x operator*(float3x a, float3x b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ inline float3x operator-(float3x a) { return f3(-a.x, -a.y, -a.z); }
__host__ __device__ inline float dot3(float3x a, float3x b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline float3x cross3(float3x a, float3x b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__host__ __device__ inline float3x normalize3(float3x a) {
#ifdef __CUDA_ARCH__
    const float l = rsqrtf(dot3(a, a));
#else
    const float l = 1.0f / sqrtf(dot3(a, a));
#endif
    return a * l;
}
__host__ __device__ inline float average3(float3x a) { return (a.x + a.y + a.z) * 0.333333f; }   // Vector3::average, src/Vector3.h:258

// ---------------------------------------------------------------------------------------------
// Camera (Camera::eyeRayAdaptive, src/Camera.cpp:116-174).  The basis is computed once per frame.
struct DeviceCamera {
    float3x eye, u, v, w;
    float top, right;
    float focus_plane, aperture, shutter;
};

struct CameraSample { float3x o, d; float time; };

// x,y pixel; [minX,maxX]x[minY,maxY] sub-cell of the pixel; addr identifies the camera sample
__device__ inline CameraSample camera_ray(const DeviceCamera& c, int x, int y, float minX, float maxX, float minY, float maxY,
                                          int width, int height, const RandAddr& addr) {
    const Rand4 r = rand4(addr, RP_CAMERA, 0, 0, 0, 0);
    const float xOffset = (maxX - minX) * r.x + minX;
    const float yOffset = (maxY - minY) * r.y + minY;
    const float left = -c.right, bottom = -c.top;
    const float U = left + (c.right - left) * (((float)x + xOffset) / (float)width);
    const float V = bottom + (c.top - bottom) * (((float)y + yOffset) / (float)height);
    CameraSample s;
    s.time = 1.f - r.z * r.z * r.z * c.shutter;                      // Camera::getTimeSample, src/Camera.h:46
    const float3x dir = normalize3(U * c.u + V * c.v - c.w);
    if (c.aperture < kEps) { s.o = c.eye; s.d = dir; return s; }
    const float3x focal = dir * c.focus_plane + c.eye;
    float lu = 0.f, lv = 0.f;
    for (uint32_t attempt = 0; attempt < 64; ++attempt) {            // rejection-sample the lens disc (src/Camera.cpp:163-167)
        const Rand4 q = rand4(addr, RP_LENS, 0, 0, 0, attempt);
        lu = 1.0f - 2.f * q.x; lv = 1.0f - 2.f * q.y;
        if (lu * lu + lv * lv <= 1.0f) break;
        lu = 1.0f - 2.f * q.z; lv = 1.0f - 2.f * q.w;
        if (lu * lu + lv * lv <= 1.0f) break;
    }
    s.o = c.aperture * (lu * c.u + lv * c.v) + c.eye;
    s.d = normalize3(focal - s.o);
    return s;
}

// ---------------------------------------------------------------------------------------------
// Textures (Texture::getPixel / getLookup / getLookupXYZ3, src/Texture.cpp:43-125)
__device__ inline float4 tex_pixel(const DeviceTexture& t, int x, int y) {
    x = x % t.width; y = y % t.height;
    if (t.channels == 1) { const float g = __ldg(t.texels + (size_t)y * t.width + x); return make_float4(g, g, g, 1.f); }
    const float* p = t.texels + ((size_t)y * t.width + x) * t.channels;
    return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), t.channels == 4 ? __ldg(p + 3) : 1.f);
}
__device__ inline float4 tex_lookup(const DeviceTexture& t, float u, float v) {
    u = u - float(int(u)); v = v - float(int(v));
    if (u < 0.0f) u = u + 1.0f;
    if (v < 0.0f) v = v + 1.0f;
    v = 1.0f - v;
    const float px = u * t.width, py = v * t.height;
    const float x1 = floorf(px), x2 = x1 + 1.0f, dx = px - x1;
    const float y1 = floorf(py), y2 = y1 + 1.0f, dy = py - y1;
    const float4 a = tex_pixel(t, (int)x1, (int)y1), b = tex_pixel(t, (int)x2, (int)y1);
    const float4 c = tex_pixel(t, (int)x1, (int)y2), d = tex_pixel(t, (int)x2, (int)y2);
    float4 r;
    r.x = (a.x * (1.0f - dx) + b.x * dx) * (1.0f - dy) + (c.x * (1.0f - dx) + d.x * dx) * dy;
    r.y = (a.y * (1.0f - dx) + b.y * dx) * (1.0f - dy) + (c.y * (1.0f - dx) + d.y * dx) * dy;
    r.z = (a.z * (1.0f - dx) + b.z * dx) * (1.0f - dy) + (c.z * (1.0f - dx) + d.z * dx) * dy;
    r.w = (a.w * (1.0f - dx) + b.w * dx) * (1.0f - dy) + (c.w * (1.0f - dx) + d.w * dx) * dy;
    return r;
}
__device__ inline float3x tex_lookup_dir(const DeviceTexture& t, float3x d) {
    const float theta = atan2f(d.z, d.x) + kPI;
    const float phi = acosf(fminf(fmaxf(d.y, -1.f), 1.f));
    const float u = theta * 0.5f * k1_PI;
    const float v = 1.0f - (phi * k1_PI);
    const float4 c = tex_lookup(t, u, v);
    return f3(c.x, c.y, c.z);
}

// environment seen by a ray that leaves the scene (Scene::sampleScene miss branch, src/Scene.cpp:234-240;
// Material::getEnvironmentColor with the scene's map, src/Material.cpp:44-63)
__device__ inline float3x environment(const DeviceShading& sh, float3x d) {
    if (sh.env_map >= 0) return tex_lookup_dir(sh.textures[sh.env_map], d) * sh.env_exposure;
    return f3(sh.bg[0], sh.bg[1], sh.bg[2]);
}

// ---------------------------------------------------------------------------------------------
// Hit attributes (HitInfo::getAllInfos, src/Ray.cpp:5-50)
struct Surface {
    float3x P, N, geoN;
    float3x T, BT;         // interpolated tangent / bitangent; evaluated only for normal-mapped materials (zero otherwise)
    float u, v;
    uint32_t material;
};

__device__ inline Surface surface_at(const DeviceScene& sc, const DeviceShading& sh, float3x o, float3x d, float t, float a, float b,
                                     int32_t prim, int32_t inst) {
    Surface s;
    s.P = o + d * t;                                                  // Ray::getPoint
    const miro_gpu_prim* pr = sh.prims + prim;
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(pr));       // n[0..2], uv[0]
    const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(pr) + 1);   // uv[1..2], material, mesh
    s.material = q1.z;
    // geometric normal from pose-1 vertices (MB objects shade with mesh 1, src/Ray.cpp:12-19)
    const float4* tv = (uint32_t)prim < sc.n_tris ? sc.tris + (size_t)prim * TRI_F4 : sc.mbtris + (size_t)((uint32_t)prim - sc.n_tris) * 6;
    const float4 p0 = __ldg(tv), p1 = __ldg(tv + 1), p2 = __ldg(tv + 2);
    const float3x e0 = f3(p1.x - p0.x, p1.y - p0.y, p1.z - p0.z), e1 = f3(p2.x - p0.x, p2.y - p0.y, p2.z - p0.z);
    s.geoN = normalize3(cross3(e0, e1));
    const float c = 1.0f - a - b;
    const float* n0 = sh.normals + (size_t)q0.x * 3; const float* n1 = sh.normals + (size_t)q0.y * 3; const float* n2 = sh.normals + (size_t)q0.z * 3;
    s.N = normalize3(f3(__ldg(n0) * c + __ldg(n1) * a + __ldg(n2) * b,
                        __ldg(n0 + 1) * c + __ldg(n1 + 1) * a + __ldg(n2 + 1) * b,
                        __ldg(n0 + 2) * c + __ldg(n1 + 2) * a + __ldg(n2 + 2) * b));
    if (inst >= 0) {                                                   // m_invTranspose * n, src/Ray.cpp:27-31
        const float* m = sh.inst_nxf + (size_t)inst * 9;
        const float3x r0 = f3(__ldg(m), __ldg(m + 1), __ldg(m + 2)), r1 = f3(__ldg(m + 3), __ldg(m + 4), __ldg(m + 5)), r2 = f3(__ldg(m + 6), __ldg(m + 7), __ldg(m + 8));
        s.geoN = normalize3(f3(dot3(r0, s.geoN), dot3(r1, s.geoN), dot3(r2, s.geoN)));
        s.N = normalize3(f3(dot3(r0, s.N), dot3(r1, s.N), dot3(r2, s.N)));
    }
    s.T = f3(0, 0, 0); s.BT = f3(0, 0, 0);
    if (q0.w != 0xffffffffu) {
        const float* t0 = sh.uvs + (size_t)q0.w * 2; const float* t1 = sh.uvs + (size_t)q1.x * 2; const float* t2 = sh.uvs + (size_t)q1.y * 2;
        s.u = __ldg(t0) * c + __ldg(t1) * a + __ldg(t2) * b;
        s.v = __ldg(t0 + 1) * c + __ldg(t1 + 1) * a + __ldg(t2 + 1) * b;
        // tangent frame, indexed by the NORMAL index triple and not transformed by an instance (src/Ray.cpp:22,35-36)
        if (sh.tangents && sh.bitangents && sh.materials[s.material].normal_map >= 0) {
            const float* g0 = sh.tangents + (size_t)q0.x * 3; const float* g1 = sh.tangents + (size_t)q0.y * 3; const float* g2 = sh.tangents + (size_t)q0.z * 3;
            s.T = normalize3(f3(__ldg(g0) * c + __ldg(g1) * a + __ldg(g2) * b, __ldg(g0 + 1) * c + __ldg(g1 + 1) * a + __ldg(g2 + 1) * b,
                                __ldg(g0 + 2) * c + __ldg(g1 + 2) * a + __ldg(g2 + 2) * b));
            const float* h0 = sh.bitangents + (size_t)q0.x * 3; const float* h1 = sh.bitangents + (size_t)q0.y * 3; const float* h2 = sh.bitangents + (size_t)q0.z * 3;
            s.BT = normalize3(f3(__ldg(h0) * c + __ldg(h1) * a + __ldg(h2) * b, __ldg(h0 + 1) * c + __ldg(h1 + 1) * a + __ldg(h2 + 1) * b,
                                 __ldg(h0 + 2) * c + __ldg(h1 + 2) * a + __ldg(h2 + 2) * b));
        }
    } else { s.u = a; s.v = b; }
    return s;
}

// cosine-distributed direction around N (Material::getCosineDistributedSamples, src/Material.cpp:14-42)
__device__ inline float3x cosine_sample(float3x N, float e1, float e2) {
    e2 = (e2 > 0.99f) ? 0.99f : e2;
    const float3x axis = (fabsf(N.x) > 0.1f) ? f3(0, 1, 0) : f3(1, 0, 0);
    const float3x u = normalize3(cross3(axis, N));
    const float3x v = cross3(N, u);
    const float ang = 2 * kPI * e1;
    const float s2 = sqrtf(e2), s1 = sqrtf(fabsf(1.0f - e2));
    float sn, cs; sincosf(ang, &sn, &cs);
    return normalize3((cs * s2) * u + (sn * s2) * v + s1 * N);
}

// Fresnel reflectance, full (non-Schlick) form of Material::fresnel (src/Material.h:47-55)
__device__ inline float fresnel(float n1, float n2, float cosThetaI) {
    const float n1CosTh = n1 * cosThetaI;
    const float n1_n2SinTh = n1 * sinf(acosf(cosThetaI)) / n2;
    const float n2CosTh = n2 * fmaxf(0.0f, sqrtf(1.0f - n1_n2SinTh * n1_n2SinTh));     // fmaxf(0, NaN) = 0: total internal reflection
    const float Rs = (n1CosTh - n2CosTh) / (n1CosTh + n2CosTh);
    return Rs * Rs;
}

// The IOR history a ray carries (Ray::IORList, src/Ray.h:43-51): 7 entries are enough for 2 initial + 5 refractions.
struct IorStack {
    float v[7];
    int idx;
    __device__ inline void init_camera() { v[0] = 1.0f; v[1] = 1.001f; idx = 1; for (int i = 2; i < 7; ++i) v[i] = 0.f; }   // IORList() + push(1.001)
    __device__ inline float top() const { return v[idx]; }
    __device__ inline void pop() { if (idx > 0) idx--; }
    __device__ inline void push(float x) { if (idx < 6) ++idx; v[idx] = x; }
};

// ---------------------------------------------------------------------------------------------
// Light sampling (Light::sampleLight of the three light classes).  One call of the reference's
// sampleLight = one LightLoop: a data-dependent number of samples, each either contributing nothing
// or needing one shadow ray.  The loop is pure (counter-based RNG), so the shade kernel runs it twice:
// once to COUNT the shadow rays it will emit (for a warp-aggregated queue allocation) and once to EMIT.
struct LightSample {
    float3x dir;        // unit direction towards the light sample
    float tmin, tmax;   // shadow-ray interval
    float dist;         // distance to the light sample (the walk length of the "full" shadow method; MIRO_GPU_TMAX for the dome)
    float3x E;          // irradiance of the sample if unoccluded (already times the cosine where the reference applies it)
    float spec;         // specular lobe input of the sample if unoccluded
    bool lit;           // false: contributes nothing (back-facing); no shadow ray
};

// Calls f(const LightSample&) for every LIT sample, in order; returns samplesDone (the divisor of the mean).
template <class F>
__device__ inline int light_loop(const DeviceShading& sh, uint32_t li, float3x from, float3x normal, float3x rVec, bool isSecondary,
                                 uint32_t pass, const RandAddr& addr, F&& f) {
    const miro_gpu_light* Lp = sh.lights + li;
    const uint32_t kind = Lp->kind;
    const float power = Lp->power, noise = Lp->noise_threshold;
    const int num_samples = Lp->num_samples;
    if (kind == MIRO_GPU_LIGHT_POINT) {                                // src/PointLight.cpp:8-82
        float3x L = f3(Lp->p0[0], Lp->p0[1], Lp->p0[2]) - from;
        float nDotL = dot3(normal, L);
        if (nDotL > 0.0f) {
            const float d2 = dot3(L, L);
            const float distRecip = rsqrtf(d2), falloff = 1.0f / d2, distance = 1.0f / distRecip;
            L = L * distRecip; nDotL *= distRecip;
            LightSample s; s.dir = L; s.tmin = 0.001f; s.tmax = distance; s.dist = distance; s.lit = true;
            const float att = nDotL;                                    // "attenuate *= nDotL": the cosine
            const float e = power * falloff * k1_4PI * att;
            s.E = f3(e, e, e); s.spec = fmaxf(0.f, dot3(rVec, L)) * att;
            f(s);
        }
        return 1;
    }
    if (kind == MIRO_GPU_LIGHT_RECT) {                                 // src/RectangleLight.cpp:42-137
        const float3x v1 = f3(Lp->p0[0], Lp->p0[1], Lp->p0[2]), v2 = f3(Lp->p1[0], Lp->p1[1], Lp->p1[2]), v3 = f3(Lp->p2[0], Lp->p2[1], Lp->p2[2]);
        float falloff = 1.0f;                                           // persists across iterations in the reference (quirk kept)
        int done = 0; bool cutOff = false;
        do {
            const Rand4 r = rand4(addr, RP_LIGHT, li, pass, (uint32_t)done, 0);
            const float e1 = r.x; float e2 = r.y; e2 = (e2 > 0.99f) ? 0.99f : e2;
            float3x dir = (v1 + e1 * (v2 - v1) + e2 * (v3 - v1)) - from;
            float nDotL = dot3(normal, dir);
            LightSample s; s.lit = false;
            if (nDotL > kEps) {
                const float d2 = dot3(dir, dir);
                const float distRecip = rsqrtf(d2); falloff = 1.0f / d2; const float distance = 1.0f / distRecip;
                dir = dir * distRecip;
                s.lit = true; s.dir = dir; s.tmin = kEps; s.tmax = distance - kEps; s.dist = distance;
                s.spec = fmaxf(0.f, dot3(rVec, dir));
            }
            const float e = power * falloff * k1_4PI;                   // no cosine at the surface (reference behaviour)
            ++done;
            cutOff = (average3(f3(e, e, e)) * (1.0f / (float)done)) < noise;
            if (s.lit) { s.E = f3(e, e, e); f(s); }
        } while (done < num_samples && !cutOff);
        return done;
    }
    // dome light, src/DomeLight.cpp:80-161 (alias table instead of two CDF searches, see dome.cuh)
    const DeviceDome& D = sh.domes[li];
    const int N = D.nu * D.nv;
    const int want = isSecondary ? 1 : num_samples;
    int done = 0; bool cutOff = false;
    do {
        LightSample s; s.lit = false;
        float4 E4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t attempt = 0; attempt < 64 && !s.lit; ++attempt) {  // back-facing draws are redrawn, not counted (DomeLight.cpp:106)
            const Rand4 r = rand4(addr, RP_LIGHT, li, pass, (uint32_t)done, attempt);
            for (int half = 0; half < 2 && !s.lit; ++half) {
                const float ea = half ? r.z : r.x, eb = half ? r.w : r.y;
                int cell = min((int)(ea * (float)N), N - 1);
                const float2 al = __ldg(D.alias + cell);
                if (!(eb < al.x)) cell = (int)__float_as_uint(al.y);
                const int u = cell / D.nv, v = cell - u * D.nv;
                const float cosT = __ldg(D.cos_v + v), sinT = __ldg(D.sin_v + v), sinP = __ldg(D.sin_u + u), cosP = __ldg(D.cos_u + u);
                const float3x dir = f3(-sinT * cosP, -cosT, -sinT * sinP);
                if (dot3(normal, dir) < 0.0f) continue;
                s.lit = true; s.dir = dir; s.tmin = kEps; s.tmax = MIRO_GPU_TMAX; s.dist = MIRO_GPU_TMAX;
                E4 = __ldg(D.cell_E + cell);
            }
        }
        const float3x E = f3(power * E4.x, power * E4.y, power * E4.z);  // gain * L(dir) / pdf, no cosine (reference behaviour)
        ++done;
        cutOff = (average3(E) * (1.0f / (float)done)) < noise;
        if (s.lit) { s.E = E; s.spec = dot3(rVec, s.dir); f(s); }
    } while (done < want && !cutOff);
    return done;
}

}  // namespace miro
