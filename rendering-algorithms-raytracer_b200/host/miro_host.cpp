// miro_host.cpp — implementation of the product's C++ host layer (see miro_host.h).
#include "miro_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <algorithm>
#include <chrono>
#include <map>
#include <xmmintrin.h>
#include <cuda_runtime.h>

namespace miro {

float referenceRecip(float w) {
    volatile float wv = w;                             // not a compile-time constant: the instruction itself must run
    const __m128 v = _mm_set1_ps(wv);
    const __m128 x0 = _mm_rcp_ps(v);
    // src/SSE.h:85: 2 * x0 - val * (x0 * x0)
    const __m128 r = _mm_sub_ps(_mm_mul_ps(_mm_set1_ps(2.0f), x0), _mm_mul_ps(v, _mm_mul_ps(x0, x0)));
    return _mm_cvtss_f32(r);
}

// =====================================================================================  meshes
void TriangleMesh::computeTangents(std::vector<Vector3>& tangents, std::vector<Vector3>& bitangents) const {
    tangents.clear(); bitangents.clear();
    if (m_texCoordIndices.empty()) return;
    tangents.assign(m_normals.size(), Vector3(0.f)); bitangents.assign(m_normals.size(), Vector3(0.f));
    for (uint32_t i = 0; i < m_numTris; ++i) {
        const TupleI3 vi = m_vertexIndices[i], ti = m_texCoordIndices[i], ni = m_normalIndices[i];
        const Vector3 A = m_vertices[vi.x], AB = m_vertices[vi.y] - A, AC = m_vertices[vi.z] - A;
        const float e1x = m_texCoords[ti.y].x - m_texCoords[ti.x].x, e1y = m_texCoords[ti.y].y - m_texCoords[ti.x].y;
        const float e2x = m_texCoords[ti.z].x - m_texCoords[ti.x].x, e2y = m_texCoords[ti.z].y - m_texCoords[ti.x].y;
        const float cp = e1y * e2x - e1x * e2y;
        if (cp == 0.0f) continue;
        const float mul = 1.f / cp;
        const Vector3 tangent = ((AB * -e2x + AC * e1y) * mul).normalized();
        const uint32_t idx[3] = {ni.x, ni.y, ni.z};
        for (uint32_t k : idx) {
            const Vector3 normal = m_normals[k];
            tangents[k] = (tangent - normal * dot(normal, tangent)).normalized();
            bitangents[k] = cross(tangents[k], normal);
        }
    }
}

void TriangleMesh::makeFlatNormals() {
    // faces without normals: one flat normal per face (src/TriangleMeshLoad.cpp:194-206)
    m_normals.clear(); m_normalIndices.resize(m_numTris);
    for (uint32_t i = 0; i < m_numTris; ++i) {
        const TupleI3 t = m_vertexIndices[i];
        Vector3 e1 = m_vertices[t.y] - m_vertices[t.x], e2 = m_vertices[t.z] - m_vertices[t.x];
        m_normals.push_back(cross(e1, e2).normalized());
        m_normalIndices[i] = TupleI3{i, i, i};
    }
}

void TriangleMesh::setGeometry(const float* vertices, uint32_t nv, const uint32_t* vidx, uint32_t nf,
                               const float* normals, uint32_t nn, const uint32_t* nidx,
                               const float* uvs, uint32_t nt, const uint32_t* tidx) {
    m_vertices.resize(nv);
    for (uint32_t i = 0; i < nv; ++i) m_vertices[i] = Vector3(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
    m_vertexIndices.resize(nf);
    for (uint32_t i = 0; i < nf; ++i) m_vertexIndices[i] = TupleI3{vidx[3 * i], vidx[3 * i + 1], vidx[3 * i + 2]};
    m_numTris = nf;
    if (normals && nidx && nn) {
        m_normals.resize(nn);
        for (uint32_t i = 0; i < nn; ++i) m_normals[i] = Vector3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
        m_normalIndices.resize(nf);
        for (uint32_t i = 0; i < nf; ++i) m_normalIndices[i] = TupleI3{nidx[3 * i], nidx[3 * i + 1], nidx[3 * i + 2]};
    } else makeFlatNormals();
    m_texCoords.clear(); m_texCoordIndices.clear();
    if (uvs && tidx && nt) {
        m_texCoords.resize(nt);
        for (uint32_t i = 0; i < nt; ++i) m_texCoords[i] = VectorR2{uvs[2 * i], uvs[2 * i + 1]};
        m_texCoordIndices.resize(nf);
        for (uint32_t i = 0; i < nf; ++i) m_texCoordIndices[i] = TupleI3{tidx[3 * i], tidx[3 * i + 1], tidx[3 * i + 2]};
    }
}

static void parseCorner(const char* w, int& v, int& t, int& n) {   // getIndices, src/TriangleMeshLoad.cpp:66-97
    v = atoi(w); t = 0; n = 0;
    const char* s1 = strchr(w, '/');
    if (!s1) return;
    t = atoi(s1 + 1);
    const char* s2 = strchr(s1 + 1, '/');
    if (s2) n = atoi(s2 + 1);
}

bool TriangleMesh::load(const char* file, const Matrix4x4& ctm) {
    FILE* fp = fopen(file, "rb");
    if (!fp) return false;
    Matrix4x4 nctm;
    if (!ctm.inverted(nctm)) { fclose(fp); return false; }
    nctm = nctm.transposed();
    m_vertices.clear(); m_normals.clear(); m_texCoords.clear();
    m_vertexIndices.clear(); m_normalIndices.clear(); m_texCoordIndices.clear();
    std::vector<Vector3> fileNormals;
    bool anyMissingNormal = false, anyTex = false;
    char line[1024];
    while (fgets(line, sizeof(line), fp)) {
        if (line[0] == 'v') {
            if (line[1] == 'n') {
                float x = 0, y = 0, z = 0; sscanf(&line[2], "%f %f %f", &x, &y, &z);
                fileNormals.push_back(nctm.transformVector(Vector3(x, y, z)).normalized());
            } else if (line[1] == 't') {
                float x = 0, y = 0; sscanf(&line[2], "%f %f", &x, &y);
                m_texCoords.push_back(VectorR2{x, y});
            } else {
                float x = 0, y = 0, z = 0; sscanf(&line[1], "%f %f %f", &x, &y, &z);
                m_vertices.push_back(ctm.transformPoint(Vector3(x, y, z)));
            }
        } else if (line[0] == 'f') {
            char s[3][64]; s[0][0] = s[1][0] = s[2][0] = 0;
            sscanf(&line[1], "%63s %63s %63s", s[0], s[1], s[2]);
            int v[3], t[3], n[3];
            for (int k = 0; k < 3; ++k) parseCorner(s[k], v[k], t[k], n[k]);
            m_vertexIndices.push_back(TupleI3{(uint32_t)(v[0] - 1), (uint32_t)(v[1] - 1), (uint32_t)(v[2] - 1)});
            m_normalIndices.push_back(TupleI3{(uint32_t)(n[0] - 1), (uint32_t)(n[1] - 1), (uint32_t)(n[2] - 1)});
            m_texCoordIndices.push_back(TupleI3{(uint32_t)(t[0] - 1), (uint32_t)(t[1] - 1), (uint32_t)(t[2] - 1)});
            if (!n[2]) anyMissingNormal = true;
            if (t[0]) anyTex = true;
        }
    }
    fclose(fp);
    m_numTris = (uint32_t)m_vertexIndices.size();
    for (const TupleI3& t : m_vertexIndices)
        if (t.x >= m_vertices.size() || t.y >= m_vertices.size() || t.z >= m_vertices.size()) return false;
    if (!anyTex || m_texCoords.empty()) { m_texCoords.clear(); m_texCoordIndices.clear(); }
    if (anyMissingNormal || fileNormals.empty()) {
        // the reference appends one flat normal per normal-less face after the file's normals
        // (src/TriangleMeshLoad.cpp:194-206); files in the tree are all-or-nothing, so: all flat
        makeFlatNormals();
    } else {
        m_normals = fileNormals;
        for (const TupleI3& t : m_normalIndices)
            if (t.x >= m_normals.size() || t.y >= m_normals.size() || t.z >= m_normals.size()) return false;
    }
    return true;
}

// =====================================================================================  images
RawImage::RawImage(int w, int h, const float* data, ImageType t) : m_width(w), m_height(h), m_imageType(t) {
    m_rawData.assign(data, data + (size_t)w * h * channels());
}

static unsigned short g_gammaToLinear[256];
static unsigned char g_linearToGamma[32769];
static float g_linearToGammaF[32769];
static bool g_gammaInit = false;
static void initGamma() {   // Image::generateGammaTables, src/Image.cpp:19-35
    if (g_gammaInit) return;
    const float GAMMA = 2.2f;
    for (int i = 0; i < 256; i++) g_gammaToLinear[i] = (unsigned short)(int)(powf(i / 255.0f, GAMMA) * 32768.0 + 0.5);
    for (int i = 0; i < 32769; i++) {
        float r2 = powf(i / 32768.0f, 1 / GAMMA) * 255.0 + 0.5;
        g_linearToGammaF[i] = r2;
        g_linearToGamma[i] = (unsigned char)(int)r2;
    }
    g_gammaInit = true;
}

static bool loadHDR(const char* fn, RawImage& img) {
    // Radiance RGBE, same decode as src/hdrloader.cpp (component = mantissa/256 * 2^(e-128); rows kept in file order)
    FILE* f = fopen(fn, "rb");
    if (!f) return false;
    char hdr[11] = {0};
    if (fread(hdr, 10, 1, f) != 1 || memcmp(hdr, "#?RADIANCE", 10) != 0) { fclose(f); return false; }
    int c = 0, oldc = 0;
    while (true) { oldc = c; c = fgetc(f); if (c == EOF) { fclose(f); return false; } if (c == 0xa && oldc == 0xa) break; }
    char reso[200]; int i = 0;
    while (i < 199) { c = fgetc(f); if (c == EOF) { fclose(f); return false; } reso[i++] = (char)c; if (c == 0xa) break; }
    reso[i] = 0;
    int w = 0, h = 0;
    if (sscanf(reso, "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) { fclose(f); return false; }
    img.m_width = w; img.m_height = h; img.m_imageType = HDR;
    img.m_rawData.assign((size_t)w * h * 3, 0.f);
    std::vector<unsigned char> scan((size_t)w * 4);
    for (int y = 0; y < h; ++y) {
        bool rle = false;
        if (w >= 8 && w <= 0x7fff) {
            int b0 = fgetc(f), b1 = fgetc(f), b2 = fgetc(f), b3 = fgetc(f);
            if (b0 == 2 && b1 == 2 && !(b2 & 128)) rle = true;
            else { fseek(f, -4, SEEK_CUR); (void)b3; }
        }
        if (rle) {
            for (int ch = 0; ch < 4; ++ch)
                for (int j = 0; j < w;) {
                    int code = fgetc(f);
                    if (code == EOF) { fclose(f); return false; }
                    if (code > 128) { code &= 127; int val = fgetc(f); while (code-- && j < w) scan[(size_t)4 * j++ + ch] = (unsigned char)val; }
                    else while (code-- && j < w) scan[(size_t)4 * j++ + ch] = (unsigned char)fgetc(f);
                }
        } else {
            int rshift = 0;
            for (int j = 0; j < w;) {   // flat / old run-length pixels
                unsigned char p[4];
                if (fread(p, 1, 4, f) != 4) { fclose(f); return false; }
                if (p[0] == 1 && p[1] == 1 && p[2] == 1 && j > 0) {
                    for (int k = p[3] << rshift; k > 0 && j < w; --k, ++j) memcpy(&scan[(size_t)4 * j], &scan[(size_t)4 * (j - 1)], 4);
                    rshift += 8;
                } else { memcpy(&scan[(size_t)4 * j], p, 4); ++j; rshift = 0; }
            }
        }
        float* out = &img.m_rawData[(size_t)y * w * 3];
        for (int j = 0; j < w; ++j) {
            const int expo = scan[(size_t)4 * j + 3] - 128;
            const float d = (float)pow(2.0f, expo);
            out[3 * j + 0] = (scan[(size_t)4 * j + 0] / 256.0f) * d;
            out[3 * j + 1] = (scan[(size_t)4 * j + 1] / 256.0f) * d;
            out[3 * j + 2] = (scan[(size_t)4 * j + 2] / 256.0f) * d;
        }
    }
    fclose(f);
    return true;
}

static bool loadPPM(const char* fn, RawImage& img) {   // src/RawImage.cpp:33-87 (P6, values / 255)
    FILE* f = fopen(fn, "rb");
    if (!f) return false;
    char buf[128]; int w = 0, h = 0;
    if (!fgets(buf, sizeof(buf), f)) { fclose(f); return false; }
    do { if (!fgets(buf, sizeof(buf), f)) { fclose(f); return false; } } while (buf[0] == '#');
    if (sscanf(buf, "%d %d", &w, &h) != 2) { fclose(f); return false; }
    do { if (!fgets(buf, sizeof(buf), f)) { fclose(f); return false; } } while (buf[0] == '#');
    std::vector<unsigned char> raw((size_t)w * h * 3);
    if (fread(raw.data(), raw.size(), 1, f) != 1) { fclose(f); return false; }
    fclose(f);
    img.m_width = w; img.m_height = h; img.m_imageType = RGB;
    img.m_rawData.resize(raw.size());
    for (size_t i = 0; i < raw.size(); ++i) img.m_rawData[i] = ((float)raw[i]) / 255;
    return true;
}

static bool loadTGA(const char* fn, RawImage& img) {   // src/RawImage.cpp:89-187 (type 2/3, vertical flip, gamma->linear, BGR->RGB)
    initGamma();
    FILE* f = fopen(fn, "rb");
    if (!f) return false;
    unsigned char h[18];
    if (fread(h, 1, 18, f) != 18) { fclose(f); return false; }
    const int type = h[2], width = h[12] | (h[13] << 8), height = h[14] | (h[15] << 8), mode = h[16] / 8;
    if ((type != 2 && type != 3) || (mode != 1 && mode != 3 && mode != 4)) { fclose(f); return false; }
    const size_t total = (size_t)width * height * mode;
    std::vector<unsigned char> raw(total), flip(total);
    if (fread(raw.data(), 1, total, f) != total) { fclose(f); return false; }
    fclose(f);
    for (int i = 0; i < height; i++) memcpy(&flip[(size_t)(height - i - 1) * width * mode], &raw[(size_t)i * width * mode], (size_t)width * mode);
    img.m_width = width; img.m_height = height;
    img.m_rawData.resize(total);
    for (size_t i = 0; i < total; ++i) img.m_rawData[i] = float(g_gammaToLinear[flip[i]]) / 32768.f;
    if (mode == 4) for (size_t i = 3; i < total; i += 4) img.m_rawData[i] = float(flip[i]) / 255.f;
    img.m_imageType = mode == 1 ? GRAYSCALE : (mode == 3 ? RGB : RGBA);
    if (mode >= 3) for (size_t i = 0; i + 2 < total; i += mode) std::swap(img.m_rawData[i], img.m_rawData[i + 2]);
    return true;
}

bool RawImage::loadImage(const char* filename) {
    std::string fn(filename);
    std::string ext = fn.substr(fn.find_last_of(".") + 1);
    if (ext == "ppm" || ext == "PPM") return loadPPM(filename, *this);
    if (ext == "hdr" || ext == "HDR") return loadHDR(filename, *this);
    if (ext == "tga" || ext == "TGA") return loadTGA(filename, *this);
    return false;
}

Image::Image() { initGamma(); }
Image::~Image() { unpin(); }
void Image::pin() {
    if (m_pinned || m_radiance.empty()) return;
    if (cudaHostRegister(m_radiance.data(), m_radiance.size() * sizeof(float), cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaHostRegister(m_pixels.data(), m_pixels.size(), cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); cudaHostUnregister(m_radiance.data()); return; }
    m_pinned = true;
}
void Image::unpin() {
    if (!m_pinned) return;
    cudaHostUnregister(m_radiance.data()); cudaHostUnregister(m_pixels.data()); cudaGetLastError();
    m_pinned = false;
}
void Image::resize(int width, int height) {
    unpin();
    m_width = width; m_height = height;
    m_pixels.assign((size_t)width * height * 3, 0);
    m_radiance.assign((size_t)width * height * 3, 0.f);
}
unsigned char Image::Map(float r) {
    initGamma();
    float rMap = 32768.0f * r;
    unsigned short linear = (rMap > 32768.0f) ? 32768 : (unsigned short)(rMap > 0.f ? rMap : 0.f);
    return g_linearToGamma[linear];
}
const float* Image::linearToGammaF() { initGamma(); return g_linearToGammaF; }
void Image::setPixel(int x, int y, const Vector3& p) {
    if (x >= 0 && x < m_width && y < m_height && y >= 0) {
        unsigned char* q = &m_pixels[((size_t)y * m_width + x) * 3];
        q[0] = Map(p.x); q[1] = Map(p.y); q[2] = Map(p.z);
    }
}
void Image::writePPM(const char* file) const {
    FILE* fp = fopen(file, "wb");
    if (!fp) { fprintf(stderr, "Couldn't open PPM file %s for writing\n", file); return; }
    fprintf(fp, "P6\n%d %d\n255\n", m_width, m_height);
    const int stride = m_width * 3;
    for (int i = m_height - 1; i >= 0; i--) fwrite(&m_pixels[(size_t)stride * i], stride, 1, fp);
    fclose(fp);
}

// =====================================================================================  materials / lights / camera
static void copy3(float* d, const Vector3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

void Lambert::fill(miro_gpu_material& m) const {
    memset(&m, 0, sizeof(m));
    m.kind = MIRO_GPU_MAT_LAMBERT; copy3(m.kd, m_kd); copy3(m.ka, m_ka);
    m.spec_exp = 1.f; m.spec_gloss = 1.f;
    m.color_map = m_colorMap ? m_colorMap->ordinal : -1; m.alpha_map = m_alphaMap ? m_alphaMap->ordinal : -1;
    m.normal_map = m.specular_map = m.reflect_map = m.refract_map = -1;      // Lambert::shade reads the colour map only (src/Lambert.cpp:19-53)
    m.translucency = m_translucency; m.refract_amt = m_refractAmt; m.sample_env = m_sampleEnv ? 1u : 0u;
}

Blinn::Blinn(const Vector3& kd, const Vector3& ka, const Vector3& ks, const Vector3& kt, float ior, float specExp,
             float specAmt, float reflectAmt, float refractAmt, float specGloss)
    : m_kd(kd), m_ka(ka), m_ks(ks), m_kt(kt), m_specExp(specExp), m_specAmt(specAmt), m_reflectAmt(reflectAmt), m_specGloss(specGloss), m_Le(0.f) {
    m_ior[0] = m_ior[1] = m_ior[2] = ior;
    m_lightEmitted = 0.f; m_refractAmt = refractAmt; m_translucency = 0.f;
}
void Blinn::fill(miro_gpu_material& m) const {
    memset(&m, 0, sizeof(m));
    m.kind = MIRO_GPU_MAT_BLINN; copy3(m.kd, m_kd); copy3(m.ka, m_ka); copy3(m.ks, m_ks);
    m.spec_exp = m_specExp; m.spec_amt = m_specAmt; m.emit_intensity = m_lightEmitted; copy3(m.le, m_Le);
    m.color_map = m_colorMap ? m_colorMap->ordinal : -1; m.alpha_map = m_alphaMap ? m_alphaMap->ordinal : -1;
    m.normal_map = m_normalMap ? m_normalMap->ordinal : -1; m.specular_map = m_specularMap ? m_specularMap->ordinal : -1;
    m.reflect_map = m_reflectMap ? m_reflectMap->ordinal : -1; m.refract_map = m_refractMap ? m_refractMap->ordinal : -1;
    m.reflect_amt = m_reflectAmt; m.refract_amt = m_refractAmt; m.spec_gloss = m_specGloss;
    m.ior[0] = m_ior[0]; m.ior[1] = m_ior[1]; m.ior[2] = m_ior[2]; m.disperse = m_disperse ? 1u : 0u;
    m.translucency = m_translucency; m.sample_env = m_sampleEnv ? 1u : 0u;
}

void PointLight::fill(miro_gpu_light& l) const {
    memset(&l, 0, sizeof(l));
    l.kind = MIRO_GPU_LIGHT_POINT; copy3(l.p0, m_position); l.power = m_power; l.num_samples = 1;
    l.noise_threshold = m_noiseThreshold; l.cast_shadows = m_castShadows ? 1u : 0u; l.texture = -1;
    l.full_shadows = m_fastShadows ? 0u : 1u;
}
void RectangleLight::setPower(float f) {
    Vector3 e0 = m_v2 - m_v1, e1 = m_v3 - m_v1;
    float recip = 1.0f, areaSq;
    if (fabsf(dot(e0, e1)) < MIRO_GPU_EPSILON) areaSq = e0.length2() * e1.length2();
    else areaSq = cross(e0, e1).length2();
    if (areaSq > MIRO_GPU_EPSILON) recip = 1.0f / sqrtf(areaSq);
    m_power = f * recip;
}
void RectangleLight::fill(miro_gpu_light& l) const {
    memset(&l, 0, sizeof(l));
    l.kind = MIRO_GPU_LIGHT_RECT; copy3(l.p0, m_v1); copy3(l.p1, m_v2); copy3(l.p2, m_v3); l.power = m_power;
    l.num_samples = m_numSamples; l.noise_threshold = m_noiseThreshold; l.cast_shadows = m_castShadows ? 1u : 0u; l.texture = -1;
    l.full_shadows = m_fastShadows ? 0u : 1u;
}
void DomeLight::fill(miro_gpu_light& l) const {
    memset(&l, 0, sizeof(l));
    l.kind = MIRO_GPU_LIGHT_DOME; l.power = m_Gain; l.num_samples = m_numSamples; l.noise_threshold = m_noiseThreshold;
    l.cast_shadows = 1u; l.texture = m_lightMap ? m_lightMap->ordinal : -1;
    l.full_shadows = m_fastShadows ? 0u : 1u;
}

Camera::Camera()   // src/Camera.cpp:15-27 (the default fov there is radians-by-mistake; every scene calls setFOV)
    : m_eye(0, 0, 0), m_up(0, 1, 0), m_viewDir(0, 0, -1), m_lookAt(FLT_MAX, FLT_MAX, FLT_MAX),
      m_fov((45.) * (PI / 180.)), m_focusPlane(1.0f), m_aperture(0.0f), m_shutterSpeed(MIRO_GPU_EPSILON) {}
void Camera::fill(miro_gpu_camera& c) const {
    copy3(c.eye, m_eye); copy3(c.view_dir, m_viewDir); copy3(c.up, m_up);
    c.fov_deg = m_fov; c.focus_plane = m_focusPlane; c.aperture = m_aperture; c.shutter_speed = m_shutterSpeed;
}

// =====================================================================================  objects
void makeMeshObjs(Scene* scene, TriangleMesh* mesh, Material* mat) {
    for (int i = (int)mesh->m_numTris - 1; i >= 0; --i) {
        Object o; o.m_mesh = mesh; o.m_index = (uint32_t)i; o.m_material = mat; o.m_objectType = OBJECT;
        scene->addObject(o);
    }
}
void makeMBMeshObjs(Scene* scene, TriangleMesh* mesh, TriangleMesh* mesh2, Material* mat) {
    for (int i = (int)mesh->m_numTris - 1; i >= 0; --i) {
        Object o; o.m_mesh = mesh; o.m_mesh_t2 = mesh2; o.m_index = (uint32_t)i; o.m_material = mat; o.m_objectType = MB_OBJECT;
        scene->addObject(o);
    }
}
void addProxyObject(Scene* scene, ProxyBLAS* blas, const Matrix4x4& m) {
    Object o; o.m_objectType = PROXY_OBJECT; o.m_blas = blas; o.m_transform = m;
    scene->addObject(o);
}
ProxyBLAS* ProxyBLAS::setupProxy(TriangleMesh* mesh, Material* mat) { TriangleMesh* m[1] = {mesh}; Material* a[1] = {mat}; return setupMultiProxy(m, 1, a); }
ProxyBLAS* ProxyBLAS::setupMultiProxy(TriangleMesh* mesh[], int numObjs, Material* mat[]) {
    ProxyBLAS* b = new ProxyBLAS;
    for (int j = 0; j < numObjs; ++j)
        for (int i = (int)mesh[j]->m_numTris - 1; i >= 0; --i) {
            Object o; o.m_mesh = mesh[j]; o.m_index = (uint32_t)i; o.m_material = mat[j]; o.m_objectType = OBJECT;
            b->m_objects.push_back(o);
        }
    return b;
}

// =====================================================================================  scene
miro_gpu_scene_desc FlatScene::desc() const {
    miro_gpu_scene_desc d;
    memset(&d, 0, sizeof(d));
    d.abi_version = MIRO_GPU_ABI_VERSION;
    d.nodes = nodes.data(); d.n_nodes = (uint32_t)nodes.size(); d.root = root;
    d.tris = tris.data(); d.n_tris = (uint32_t)tris.size();
    d.mbtris = mbtris.data(); d.n_mbtris = (uint32_t)mbtris.size();
    d.instances = instances.data(); d.n_instances = (uint32_t)instances.size();
    d.prims = prims.data();
    d.normals = normals.data(); d.n_normals = (uint32_t)(normals.size() / 3);
    d.uvs = uvs.empty() ? nullptr : uvs.data(); d.n_uvs = (uint32_t)(uvs.size() / 2);
    d.inst_normal_xform = inst_nxf.empty() ? nullptr : inst_nxf.data();
    d.tangents = tangents.size() == normals.size() ? tangents.data() : nullptr;
    d.bitangents = bitangents.size() == normals.size() ? bitangents.data() : nullptr;
    d.materials = materials.data(); d.n_materials = (uint32_t)materials.size();
    d.lights = lights.data(); d.n_lights = (uint32_t)lights.size();
    d.textures = textures.data(); d.n_textures = (uint32_t)textures.size();
    d.env_map = env_map; d.env_exposure = env_exposure;
    d.bg_color[0] = bg[0]; d.bg_color[1] = bg[1]; d.bg_color[2] = bg[2];
    return d;
}

Scene::Scene() {}
Scene::~Scene() { if (m_group) miro_gpu_group_destroy(m_group); if (m_ctx) miro_gpu_destroy(m_ctx); }

int Scene::meshOrdinal(TriangleMesh* m) {
    if (m->ordinal < 0) {
        int next = 0;
        for (TriangleMesh* o : meshes) next = std::max(next, o->ordinal + 1);
        m->ordinal = next;
    }
    if (std::find(meshes.begin(), meshes.end(), m) == meshes.end()) {
        meshes.push_back(m);
        // append this mesh's shading attributes to the global arrays
        if (m_meshNormalBase.size() <= (size_t)m->ordinal) { m_meshNormalBase.resize(m->ordinal + 1, 0); m_meshUvBase.resize(m->ordinal + 1, 0); }
        m_meshNormalBase[m->ordinal] = (uint32_t)(m_flat.normals.size() / 3);
        for (const Vector3& n : m->m_normals) { m_flat.normals.push_back(n.x); m_flat.normals.push_back(n.y); m_flat.normals.push_back(n.z); }
        std::vector<Vector3> tg, bt;
        m->computeTangents(tg, bt);
        tg.resize(m->m_normals.size(), Vector3(0.f)); bt.resize(m->m_normals.size(), Vector3(0.f));      // no uvs: T = BT = 0 (src/Ray.cpp:44-45)
        for (const Vector3& t : tg) { m_flat.tangents.push_back(t.x); m_flat.tangents.push_back(t.y); m_flat.tangents.push_back(t.z); }
        for (const Vector3& t : bt) { m_flat.bitangents.push_back(t.x); m_flat.bitangents.push_back(t.y); m_flat.bitangents.push_back(t.z); }
        m_meshUvBase[m->ordinal] = (uint32_t)(m_flat.uvs.size() / 2);
        for (const TriangleMesh::VectorR2& t : m->m_texCoords) { m_flat.uvs.push_back(t.x); m_flat.uvs.push_back(t.y); }
    }
    return m->ordinal;
}
int Scene::textureOrdinal(Texture* t) {
    if (!t) return -1;
    auto it = std::find(m_textureList.begin(), m_textureList.end(), t);
    if (it == m_textureList.end()) { t->ordinal = (int)m_textureList.size(); m_textureList.push_back(t); }
    return t->ordinal;
}
int Scene::materialOrdinal(const Material* m) {
    auto it = std::find(m_materialList.begin(), m_materialList.end(), m);
    if (it != m_materialList.end()) return (int)(it - m_materialList.begin());
    textureOrdinal(m->m_colorMap); textureOrdinal(m->m_alphaMap);
    textureOrdinal(m->m_normalMap); textureOrdinal(m->m_specularMap); textureOrdinal(m->m_reflectMap); textureOrdinal(m->m_refractMap);
    const_cast<Material*>(m)->ordinal = (int)m_materialList.size();
    m_materialList.push_back(m);
    return m->ordinal;
}

static void triBounds(const miro_gpu_tri& t, float lo[3], float hi[3]) {
    for (int k = 0; k < 3; ++k) { lo[k] = std::min(t.v0[k], std::min(t.v1[k], t.v2[k])); hi[k] = std::max(t.v0[k], std::max(t.v1[k], t.v2[k])); }
}
static miro_gpu_tri makeTri(const TriangleMesh* m, uint32_t i) {
    const TriangleMesh::TupleI3 t = m->m_vertexIndices[i];
    miro_gpu_tri r; memset(&r, 0, sizeof(r));
    const Vector3 &a = m->m_vertices[t.x], &b = m->m_vertices[t.y], &c = m->m_vertices[t.z];
    r.v0[0] = a.x; r.v0[1] = a.y; r.v0[2] = a.z; r.v1[0] = b.x; r.v1[1] = b.y; r.v1[2] = b.z; r.v2[0] = c.x; r.v2[1] = c.y; r.v2[2] = c.z;
    return r;
}

bool Scene::appendTriangle(const Object& o, uint32_t& outIndex, float lo[3], float hi[3]) {
    if (!o.m_mesh || o.m_index >= o.m_mesh->m_numTris || !o.m_material) { m_error = "object without mesh/material or index out of range"; return false; }
    const int mo = meshOrdinal(o.m_mesh);
    miro_gpu_prim p; memset(&p, 0, sizeof(p));
    const TriangleMesh::TupleI3 n = o.m_mesh->m_normalIndices[o.m_index];
    p.n[0] = m_meshNormalBase[mo] + n.x; p.n[1] = m_meshNormalBase[mo] + n.y; p.n[2] = m_meshNormalBase[mo] + n.z;
    if (!o.m_mesh->m_texCoordIndices.empty()) {
        const TriangleMesh::TupleI3 t = o.m_mesh->m_texCoordIndices[o.m_index];
        p.uv[0] = m_meshUvBase[mo] + t.x; p.uv[1] = m_meshUvBase[mo] + t.y; p.uv[2] = m_meshUvBase[mo] + t.z;
    } else p.uv[0] = p.uv[1] = p.uv[2] = 0xffffffffu;
    p.material = (uint32_t)materialOrdinal(o.m_material); p.mesh = (uint32_t)mo; p.tri = o.m_index;
    if (o.m_objectType == MB_OBJECT) {
        if (!o.m_mesh_t2 || o.m_mesh_t2->m_numTris != o.m_mesh->m_numTris) { m_error = "MBObject: second pose missing or different topology"; return false; }
        miro_gpu_mbtri t; t.pose[0] = makeTri(o.m_mesh, o.m_index); t.pose[1] = makeTri(o.m_mesh_t2, o.m_index);
        float l2[3], h2[3]; triBounds(t.pose[0], lo, hi); triBounds(t.pose[1], l2, h2);
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], l2[k]); hi[k] = std::max(hi[k], h2[k]); }   // MBObject::getAABB, src/MBObject.cpp:190-198
        outIndex = (uint32_t)m_srcMB.size(); m_srcMB.push_back(t); m_srcMBPrims.push_back(p);
    } else {
        miro_gpu_tri t = makeTri(o.m_mesh, o.m_index);
        triBounds(t, lo, hi);
        outIndex = (uint32_t)m_srcTris.size(); m_srcTris.push_back(t); m_srcPrims.push_back(p);
    }
    return true;
}

// node-array AABB of a child-style reference (used for instance bounds)
static void refBounds(const FlatScene& f, const std::vector<miro_gpu_tri>& srcTris, const std::vector<uint32_t>& triOrder, int32_t ref, float lo[3], float hi[3]) {
    for (int k = 0; k < 3; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
    if (ref == MIRO_GPU_CHILD_EMPTY) return;
    if (ref < 0) {
        const uint32_t u = (uint32_t)ref, count = ((u >> MIRO_GPU_LEAF_INDEX_BITS) & 7u) + 1u, first = u & ((1u << MIRO_GPU_LEAF_INDEX_BITS) - 1u);
        for (uint32_t i = 0; i < count; ++i) {
            float l[3], h[3]; triBounds(srcTris[triOrder[first + i]], l, h);
            for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); }
        }
        return;
    }
    const miro_gpu_node& n = f.nodes[ref];
    for (int i = 0; i < 4; ++i) if (n.child[i] != MIRO_GPU_CHILD_EMPTY) {
        lo[0] = std::min(lo[0], n.lo_x[i]); lo[1] = std::min(lo[1], n.lo_y[i]); lo[2] = std::min(lo[2], n.lo_z[i]);
        hi[0] = std::max(hi[0], n.hi_x[i]); hi[1] = std::max(hi[1], n.hi_y[i]); hi[2] = std::max(hi[2], n.hi_z[i]);
    }
}

// Boxes that together bound a BLAS: the children of its top `depth`+1 levels.  An instance's world-space bounds are the
// union of these boxes transformed one by one — tighter than the eight transformed corners of the single root box the
// reference uses (ProxyObject::getAABB, src/ProxyObject.cpp:16-44), which a rotation inflates by up to sqrt(2) per axis;
// rays over an instanced field enter fewer instances.  Still conservative: every triangle lies inside one of the boxes.
static void collectBoxes(const FlatScene& f, int32_t ref, int depth, std::vector<float>& out) {
    if (ref < 0 || ref == MIRO_GPU_CHILD_EMPTY) return;
    const miro_gpu_node& n = f.nodes[ref];
    for (int i = 0; i < 4; ++i) {
        const int32_t c = n.child[i];
        if (c == MIRO_GPU_CHILD_EMPTY) continue;
        if (c >= 0 && depth > 0) { collectBoxes(f, c, depth - 1, out); continue; }
        const float b[6] = {n.lo_x[i], n.lo_y[i], n.lo_z[i], n.hi_x[i], n.hi_y[i], n.hi_z[i]};
        out.insert(out.end(), b, b + 6);
    }
}

// Sub-trees of a BLAS `depth` levels below its root (child-style references with their boxes): the units an instance is
// entered through when instances are "braided" into the top-level tree (below).
struct SubTree { int32_t ref; float box[6]; };
static void collectSubTrees(const FlatScene& f, int32_t ref, const float* box, int depth, std::vector<SubTree>& out) {
    if (ref == MIRO_GPU_CHILD_EMPTY) return;
    if (ref < 0 || depth == 0) { SubTree s; s.ref = ref; memcpy(s.box, box, sizeof(s.box)); out.push_back(s); return; }
    const miro_gpu_node& n = f.nodes[ref];
    for (int i = 0; i < 4; ++i) {
        if (n.child[i] == MIRO_GPU_CHILD_EMPTY) continue;
        const float b[6] = {n.lo_x[i], n.lo_y[i], n.lo_z[i], n.hi_x[i], n.hi_y[i], n.hi_z[i]};
        collectSubTrees(f, n.child[i], b, depth - 1, out);
    }
}

bool Scene::preCalc() {
    // MIRO_HOST_TIMING=1: phases of the scene hand-off on stderr (diagnostic)
    static const bool timing = getenv("MIRO_HOST_TIMING") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (timing) fprintf(stderr, "[miro_host] preCalc %-28s %8.1f ms since start\n", what,
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };
    m_error.clear();
    m_flat = FlatScene();
    meshes.clear(); m_materialList.clear(); m_textureList.clear();
    m_srcTris.clear(); m_srcPrims.clear(); m_srcMB.clear(); m_srcMBPrims.clear(); m_srcInst.clear(); m_srcInstNxf.clear();
    m_meshNormalBase.clear(); m_meshUvBase.clear();
    std::vector<uint32_t> order[3];
    std::map<ProxyBLAS*, std::pair<int32_t, std::pair<std::vector<float>, std::vector<float>>>> blasInfo;   // root, (lo, hi)
    // Optional (MIRO_BRAID_DEPTH = 1..3, default 0 = off): instances entered through the sub-trees `braid` levels below their BLAS's
    // root, each a top-level primitive of its own with its own (tighter) world box, so the top-level build separates the
    // overlapping boxes of neighbouring instances at sub-tree granularity (cf. Benthin et al., "Improved two-level BVHs using
    // partial re-braiding", 2017).  Measured on the 40 401-instance field (tools/braid_sweep.sh): node visits per grazing ray fall
    // 101 -> 93 / 87 / 74 at depth 1 / 2 / 3, but instance entries rise 16.7 -> 21.7 / 27.6 / 25.3 and an entry (ray transform,
    // reciprocal directions, shear constants, and again on the way out) costs several node visits: 287 -> 278 / 264 / 278 Mrays/s.
    // Not adopted; every record keeps the instance's ordinal either way.
    int braid = 0;
    if (const char* e = getenv("MIRO_BRAID_DEPTH")) { const int v = atoi(e); if (v >= 0 && v <= 3) braid = v; }
    std::map<ProxyBLAS*, std::vector<std::pair<SubTree, std::vector<float>>>> blasSubs;      // per BLAS: entry sub-trees, each with its bounding sub-boxes

    std::vector<BuildPrim> top;
    top.reserve(m_objects.size());
    uint32_t proxyOrdinal = 0;
    for (const Object& o : m_objects) {
        BuildPrim bp;
        if (o.m_objectType == PROXY_OBJECT) {
            if (!o.m_blas) { m_error = "ProxyObject without geometry"; return false; }
            if (!blasInfo.count(o.m_blas)) {
                // ProxyObject::setupProxy: build the shared bottom-level BVH once (src/ProxyObject.cpp:131-146)
                std::vector<BuildPrim> bprims; bprims.reserve(o.m_blas->m_objects.size());
                for (const Object& bo : o.m_blas->m_objects) {
                    if (bo.m_objectType != OBJECT) { m_error = "instanced geometry must be plain triangles (one level of instancing, as the reference)"; return false; }
                    BuildPrim q; q.kind = MIRO_GPU_KIND_TRI;
                    if (!appendTriangle(bo, q.index, q.lo, q.hi)) return false;
                    bprims.push_back(q);
                }
                const int32_t root = build_wide_bvh(bprims, m_flat.nodes, order, nullptr, m_srcTris.data());
                std::vector<float> lo(3), hi(3);
                refBounds(m_flat, m_srcTris, order[MIRO_GPU_KIND_TRI], root, lo.data(), hi.data());
                blasInfo[o.m_blas] = std::make_pair(root, std::make_pair(lo, hi));
                const float rootBox[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
                std::vector<SubTree> subs;
                collectSubTrees(m_flat, root, rootBox, braid, subs);
                for (const SubTree& st : subs) {
                    std::vector<float> boxes;
                    collectBoxes(m_flat, st.ref, 2, boxes);
                    if (boxes.empty()) boxes.insert(boxes.end(), st.box, st.box + 6);      // the sub-tree is a single leaf
                    blasSubs[o.m_blas].push_back(std::make_pair(st, boxes));
                }
                o.m_blas->root_ref = root; o.m_blas->flattened = true;
            }
            if (!o.m_transform.isAffine()) { m_error = "ProxyObject transform is not affine (projective instances are outside the supported scope)"; return false; }
            Matrix4x4 inv;
            if (!o.m_transform.invertedAsReference(inv)) { m_error = "ProxyObject transform is singular"; return false; }
            const Matrix4x4 it = inv.transposed();
            const uint32_t ordinal = proxyOrdinal++;
            // w = row 4 of the inverse . [o 1]: for an affine matrix rows 4's first three entries are (signed) zeros, so w is the
            // inverse's m44 — which Matrix4x4::invert leaves at sd44 * detInv, often 1 - 2^-24 rather than 1
            const float wRecip = referenceRecip(inv.at(3, 3));
            for (const auto& sub : blasSubs[o.m_blas]) {
                miro_gpu_instance in; memset(&in, 0, sizeof(in));
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) in.inv[4 * r + c] = inv.at(r, c);
                in.blas_root = sub.first.ref; in.ordinal = ordinal; in.w_recip = wRecip;
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m_srcInstNxf.push_back(it.at(r, c));
                // world bounds: the transformed corners of the sub-tree's sub-boxes (cf. ProxyObject::getAABB, src/ProxyObject.cpp:16-44,
                // which transforms the single root box)
                const std::vector<float>& boxes = sub.second;
                for (int k = 0; k < 3; ++k) { bp.lo[k] = FLT_MAX; bp.hi[k] = -FLT_MAX; }
                for (size_t b = 0; b + 6 <= boxes.size(); b += 6) {
                    const float* lo = &boxes[b]; const float* hi = lo + 3;
                    for (int c = 0; c < 8; ++c) {
                        Vector3 p((c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2]);
                        p = o.m_transform.transformPoint(p);
                        bp.lo[0] = std::min(bp.lo[0], p.x); bp.lo[1] = std::min(bp.lo[1], p.y); bp.lo[2] = std::min(bp.lo[2], p.z);
                        bp.hi[0] = std::max(bp.hi[0], p.x); bp.hi[1] = std::max(bp.hi[1], p.y); bp.hi[2] = std::max(bp.hi[2], p.z);
                    }
                }
                bp.kind = MIRO_GPU_KIND_INST; bp.index = (uint32_t)m_srcInst.size();
                m_srcInst.push_back(in);
                top.push_back(bp);
            }
            continue;
        } else {
            bp.kind = o.m_objectType == MB_OBJECT ? MIRO_GPU_KIND_MBTRI : MIRO_GPU_KIND_TRI;
            if (!appendTriangle(o, bp.index, bp.lo, bp.hi)) return false;
        }
        top.push_back(bp);
    }
    lap("primitives gathered");
    bool onDevice = m_buildOnDevice;
    for (const BuildPrim& bp : top) if (bp.kind != MIRO_GPU_KIND_TRI) onDevice = false;
    if (onDevice) {
        // hand the triangles over in object order; miro_gpu_upload_scene builds an LBVH over them on the GPU
        m_flat.root = MIRO_GPU_ROOT_BUILD_ON_DEVICE;
        for (const BuildPrim& bp : top) order[MIRO_GPU_KIND_TRI].push_back(bp.index);
    } else m_flat.root = build_wide_bvh(top, m_flat.nodes, order, &m_flat.top_stats, m_srcTris.data());

    lap("acceleration structure built");
    // gather primitives into leaf order
    m_flat.tris.resize(order[0].size()); m_flat.mbtris.resize(order[1].size()); m_flat.instances.resize(order[2].size());
    m_flat.prims.resize(order[0].size() + order[1].size());
    for (size_t i = 0; i < order[0].size(); ++i) { m_flat.tris[i] = m_srcTris[order[0][i]]; m_flat.prims[i] = m_srcPrims[order[0][i]]; }
    for (size_t i = 0; i < order[1].size(); ++i) { m_flat.mbtris[i] = m_srcMB[order[1][i]]; m_flat.prims[order[0].size() + i] = m_srcMBPrims[order[1][i]]; }
    m_flat.inst_nxf.resize(order[2].size() * 9);
    for (size_t i = 0; i < order[2].size(); ++i) {
        m_flat.instances[i] = m_srcInst[order[2][i]];
        memcpy(&m_flat.inst_nxf[i * 9], &m_srcInstNxf[(size_t)order[2][i] * 9], 9 * sizeof(float));
    }
    lap("leaf-order arrays");
    // lights / env (textures they use must get ordinals before the texture table is emitted)
    for (Light* l : m_lights) if (DomeLight* d = dynamic_cast<DomeLight*>(l)) textureOrdinal(d->m_lightMap);
    m_flat.env_map = textureOrdinal(m_envMap);
    m_flat.env_exposure = m_envExposure;
    m_flat.bg[0] = m_BGColor.x; m_flat.bg[1] = m_BGColor.y; m_flat.bg[2] = m_BGColor.z;
    for (const Material* m : m_materialList) { miro_gpu_material gm; m->fill(gm); m_flat.materials.push_back(gm); }
    for (Light* l : m_lights) { miro_gpu_light gl; l->fill(gl); m_flat.lights.push_back(gl); }
    for (Texture* t : m_textureList) {
        miro_gpu_texture gt; gt.texels = t->m_image->m_rawData.data(); gt.width = t->m_image->m_width; gt.height = t->m_image->m_height;
        gt.channels = t->m_image->channels(); gt.reserved = 0;
        m_flat.textures.push_back(gt);
    }
    if (m_flat.lights.size() > MIRO_GPU_MAX_LIGHTS) { m_error = "too many lights"; return false; }
    m_srcTris.clear(); m_srcTris.shrink_to_fit(); m_srcPrims.clear(); m_srcPrims.shrink_to_fit();
    m_srcMB.clear(); m_srcMBPrims.clear(); m_srcInst.clear(); m_srcInstNxf.clear();
    return true;
}

bool Scene::attach(int device_id) {
    if (!m_ctx) {
        int rc = miro_gpu_create(&m_ctx, device_id);
        if (rc) { m_error = std::string("miro_gpu_create: ") + miro_gpu_last_error(nullptr); m_ctx = nullptr; return false; }
    }
    miro_gpu_scene_desc d = m_flat.desc();
    int rc = miro_gpu_upload_scene(m_ctx, &d);
    if (rc) { m_error = std::string("miro_gpu_upload_scene: ") + miro_gpu_last_error(m_ctx); return false; }
    return true;
}

bool Scene::attachDevices(const int* device_ids, int n) {
    if (m_ctx) { miro_gpu_destroy(m_ctx); m_ctx = nullptr; }
    if (!m_group) {
        int rc = miro_gpu_group_create(&m_group, device_ids, n);
        if (rc) { m_error = std::string("miro_gpu_group_create: ") + miro_gpu_last_error(nullptr); m_group = nullptr; return false; }
    }
    miro_gpu_scene_desc d = m_flat.desc();
    int rc = miro_gpu_group_upload_scene(m_group, &d);
    if (rc) { m_error = std::string("miro_gpu_group_upload_scene: ") + miro_gpu_group_last_error(m_group); return false; }
    return true;
}

void Scene::renderParams(const Image* img, miro_gpu_render_params& p) const {
    memset(&p, 0, sizeof(p));
    p.width = img->width(); p.height = img->height();
    p.min_subdivs = m_minSubdivs; p.max_subdivs = m_maxSubdivs; p.noise_threshold = m_noiseThreshold;
    p.num_paths = m_numPaths; p.max_bounces = m_maxBounces; p.path_trace = m_pathTrace ? 1u : 0u;
    p.sample_env = m_sampleLightFromEnv ? 1u : 0u; p.seed = m_seed; p.shard_index = 0; p.shard_count = 1;
}

bool Scene::raytraceImage(const Camera* cam, Image* img, int shard_index, int shard_count) {
    if (!m_ctx && !m_group) { m_error = "raytraceImage: scene is not attached to a GPU (there is no CPU renderer)"; return false; }
    miro_gpu_camera c; cam->fill(c);
    miro_gpu_render_params p; renderParams(img, p);
    img->pin();
    // the frame arrives as the reference's Image holds it — 8-bit pixels through Image::setPixel's mapping, applied on the device —
    // together with the float radiance
    if (m_group) {
        if (shard_count > 1) { m_error = "raytraceImage: a scene attached to several devices shards the frame itself"; return false; }
        int rc = miro_gpu_group_render(m_group, &c, &p, m_sampleSharding ? MIRO_GPU_SHARD_SAMPLES : MIRO_GPU_SHARD_BUCKETS, img->m_radiance.data(), img->charPixels());
        if (rc) { m_error = std::string("miro_gpu_group_render: ") + miro_gpu_group_last_error(m_group); return false; }
    } else {
        p.shard_index = shard_index; p.shard_count = shard_count;
        int rc = miro_gpu_render_image(m_ctx, &c, &p, img->m_radiance.data(), img->charPixels());
        if (rc) { m_error = std::string("miro_gpu_render: ") + miro_gpu_last_error(m_ctx); return false; }
    }
    return true;
}

bool Scene::trace(const miro_gpu_ray* rays, size_t n, miro_gpu_hit* hits) {
    if (m_group) {
        int rc = miro_gpu_group_trace_closest(m_group, rays, n, hits);
        if (rc) { m_error = std::string("miro_gpu_group_trace_closest: ") + miro_gpu_group_last_error(m_group); return false; }
        return true;
    }
    if (!m_ctx) { m_error = "trace: scene is not attached to a GPU (there is no CPU tracer)"; return false; }
    int rc = miro_gpu_trace_closest(m_ctx, rays, n, hits);
    if (rc) { m_error = std::string("miro_gpu_trace_closest: ") + miro_gpu_last_error(m_ctx); return false; }
    return true;
}
bool Scene::traceAny(const miro_gpu_ray* rays, size_t n, uint32_t* bits) {
    if (m_group) {
        int rc = miro_gpu_group_trace_any(m_group, rays, n, bits);
        if (rc) { m_error = std::string("miro_gpu_group_trace_any: ") + miro_gpu_group_last_error(m_group); return false; }
        return true;
    }
    if (!m_ctx) { m_error = "traceAny: scene is not attached to a GPU (there is no CPU tracer)"; return false; }
    int rc = miro_gpu_trace_any(m_ctx, rays, n, bits);
    if (rc) { m_error = std::string("miro_gpu_trace_any: ") + miro_gpu_last_error(m_ctx); return false; }
    return true;
}

}  // namespace miro
