// dome.cuh — importance tables for DomeLight (host-side build at upload time, sampled on the GPU).
//
// Replaces DomeLight::setTexture + Distribution1D (reference src/DomeLight.cpp:8-78,
// src/DomeLight.h:10-42).  The reference draws a lat-long cell with a marginal CDF over columns and a
// conditional CDF over rows (two binary searches).  Here the SAME probability mass function over
// the nu x nv cells — mean(RGB) of the bilinearly resampled texel times sin((v+.5)pi/nv) — is
// encoded as ONE alias table (O(1) per sample, two 8-byte loads), and everything that depends only
// on the cell is precomputed: the snapped direction comes from the same trig tables, and
// L(direction)/pdf is stored per cell, so a dome sample costs no texture filtering at all.
//   P(cell) = f[u][v] / sum(f);  pdf(dir) = (p_u * p_v) / (2 pi^2 sin(theta_v))   (DomeLight.cpp:108)
#pragma once
#include <math.h>
#include <vector>
#include "context.cuh"

namespace miro {

namespace domehost {

constexpr float kPI = 3.1415926f;             // src/Miro.h:57 (the reference's truncated PI)
constexpr float k1_PI = 1.0f / kPI;
constexpr float k2_PI2 = 2.f * (kPI * kPI);

struct Tex {
    const float* t; int w, h, c;
    void pixel(int x, int y, float* out) const {            // Texture::getPixel, src/Texture.cpp:100-125
        x = x % w; y = y % h;
        if (c == 1) { float g = t[y * w + x]; out[0] = out[1] = out[2] = g; }
        else { const float* p = t + ((size_t)y * w + x) * c; out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; }
    }
    void lookup(float u, float v, float* out) const {       // Texture::getLookup, src/Texture.cpp:43-72
        u = u - float(int(u)); v = v - float(int(v));
        if (u < 0.0f) u = u + 1.0f;
        if (v < 0.0f) v = v + 1.0f;
        v = 1.0f - v;
        float px = u * w, py = v * h;
        float x1 = floorf(px), x2 = x1 + 1.0f, dx = px - x1;
        float y1 = floorf(py), y2 = y1 + 1.0f, dy = py - y1;
        float a[3], b[3], c2[3], d[3];
        pixel((int)x1, (int)y1, a); pixel((int)x2, (int)y1, b); pixel((int)x1, (int)y2, c2); pixel((int)x2, (int)y2, d);
        for (int k = 0; k < 3; ++k) {
            float q1 = a[k] * (1.0f - dx) + b[k] * dx;
            float q2 = c2[k] * (1.0f - dx) + d[k] * dx;
            out[k] = q1 * (1.0f - dy) + q2 * dy;
        }
    }
    void lookup_dir(float x, float y, float z, float* out) const {   // Texture::getLookupXYZ3, src/Texture.cpp:90-98
        double theta = atan2((double)z, (double)x) + kPI;
        double phi = acos((double)y);
        float u = (float)(theta * 0.5 * k1_PI);
        float v = (float)(1.0 - (phi * k1_PI));
        lookup(u, v, out);
    }
};

}  // namespace domehost

template <class T>
static int dome_upload(miro_gpu_ctx* ctx, const std::vector<T>& h, const T** dev) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(h.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(dome)");
    ctx->scene_allocs.push_back(p);
    if (!h.empty() && (e = cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMemcpy(dome)");
    *dev = (const T*)p;
    return MIRO_GPU_OK;
}

static int build_dome_tables(miro_gpu_ctx* ctx, const miro_gpu_texture& t, DeviceDome* out) {
    using namespace domehost;
    const int nu = t.width, nv = t.height;
    const size_t N = (size_t)nu * nv;
    Tex tex{t.texels, t.width, t.height, t.channels};
    // f[u*nv+v] = mean(RGB)(u/nu, v/nv) * sin(pi (v+.5)/nv)            (DomeLight.cpp:24-52)
    std::vector<float> f(N);
    std::vector<float> sinVals(nv);
    for (int i = 0; i < nv; ++i) sinVals[i] = sinf(kPI * float(i + .5) / float(nv));
    for (int u = 0; u < nu; ++u) {
        float up = (float)u / (float)nu;
        for (int v = 0; v < nv; ++v) {
            float vp = (float)v / (float)nv, rgb[3];
            tex.lookup(up, vp, rgb);
            f[(size_t)u * nv + v] = ((rgb[0] + rgb[1] + rgb[2]) * 0.333333f) * sinVals[v];
        }
    }
    // Distribution1D integrals (DomeLight.h:21-30): funcInt_v[u] = sum_v f/nv ; funcInt_u = sum_u funcInt_v/nu
    std::vector<float> col(nu);
    for (int u = 0; u < nu; ++u) {
        float c = 0.f;
        for (int v = 0; v < nv; ++v) c = c + f[(size_t)u * nv + v] / nv;
        col[u] = c;
    }
    float total = 0.f;
    for (int u = 0; u < nu; ++u) total = total + col[u] / nu;
    if (!(total > 0.f)) return set_error(ctx, MIRO_GPU_EINVAL, "dome light texture is black");
    // trig tables (DomeLight.cpp:59-76)
    std::vector<float> cu(nu + 1), su(nu + 1), cv(nv + 1), sv(nv + 1);
    { float inv = 1.f / float(nu); for (int i = 0; i <= nu; ++i) { cu[i] = cosf(i * inv * 2.f * kPI); su[i] = sinf(i * inv * 2.f * kPI); } }
    { float inv = 1.f / float(nv); for (int i = 0; i <= nv; ++i) { cv[i] = cosf(i * inv * kPI); sv[i] = sinf(i * inv * kPI); } }
    // per-cell E/gain = L(dir) / pdf with pdf = (p_u p_v) / (2 pi^2 sin theta)       (DomeLight.cpp:96-112,148)
    std::vector<float4> cellE(N);
    std::vector<double> w(N);
    double wsum = 0.0;
    for (int u = 0; u < nu; ++u) {
        const float p_u = col[u] * (1.f / total);
        for (int v = 0; v < nv; ++v) {
            const size_t c = (size_t)u * nv + v;
            const float p_v = col[u] > 0.f ? f[c] * (1.f / col[u]) : 0.f;
            const float sinT = sv[v], cosT = cv[v];
            const float dx = -sinT * cu[u], dy = -cosT, dz = -sinT * su[u];
            const float pdf = (p_u * p_v) / (k2_PI2 * sinT);
            float L[3]; tex.lookup_dir(dx, dy, dz, L);
            float4 e;
            e.x = L[0] / pdf; e.y = L[1] / pdf; e.z = L[2] / pdf; e.w = 1.f;
            if (!(pdf > 0.f) || isinf(pdf) || isnan(pdf)) { e.x = e.y = e.z = 0.f; }   // sin(theta) = 0 row: pdf = inf -> E = 0
            cellE[c] = e;
            w[c] = f[c] > 0.f ? (double)f[c] : 0.0;
            wsum += w[c];
        }
    }
    // Vose alias table over the N cells
    std::vector<float2> alias(N);
    {
        std::vector<double> p(N);
        std::vector<uint32_t> small, large;
        small.reserve(N); large.reserve(N);
        for (size_t i = 0; i < N; ++i) { p[i] = w[i] * (double)N / wsum; (p[i] < 1.0 ? small : large).push_back((uint32_t)i); }
        while (!small.empty() && !large.empty()) {
            uint32_t s = small.back(); small.pop_back();
            uint32_t l = large.back();
            alias[s].x = (float)p[s]; alias[s].y = __builtin_bit_cast(float, l);
            p[l] = (p[l] + p[s]) - 1.0;
            if (p[l] < 1.0) { large.pop_back(); small.push_back(l); }
        }
        for (uint32_t i : large) { alias[i].x = 1.f; alias[i].y = __builtin_bit_cast(float, i); }
        for (uint32_t i : small) { alias[i].x = 1.f; alias[i].y = __builtin_bit_cast(float, i); }
        // a zero-weight cell must never be returned: its threshold is 0 so the alias is always taken
        for (size_t i = 0; i < N; ++i) if (w[i] == 0.0 && alias[i].x > 0.f && __builtin_bit_cast(uint32_t, alias[i].y) == (uint32_t)i) alias[i].x = 0.f;
    }
    int rc;
    if ((rc = dome_upload(ctx, alias, &out->alias))) return rc;
    if ((rc = dome_upload(ctx, cellE, &out->cell_E))) return rc;
    if ((rc = dome_upload(ctx, cu, &out->cos_u))) return rc;
    if ((rc = dome_upload(ctx, su, &out->sin_u))) return rc;
    if ((rc = dome_upload(ctx, cv, &out->cos_v))) return rc;
    if ((rc = dome_upload(ctx, sv, &out->sin_v))) return rc;
    out->nu = nu; out->nv = nv;
    return MIRO_GPU_OK;
}

}  // namespace miro
