// oracle/ref_harness.cpp — headless driver for the UNMODIFIED reference ray tracer.
//
// TEST INFRASTRUCTURE ONLY.  This file is linked against the reference's own
// sources (compiled where they lie under /root/reference/src by oracle/build_ref.sh)
// and produces oracle/_ref/miro_ref.  Nothing on the product path may call it;
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs execute it as the checker / CPU baseline.
//
// It replaces the reference's GLUT main (src/main.cpp, src/MiroWindow.cpp) with:
//   * a parser for the line-based ".miro" scene script (the same script the
//     product's host library reads), which drives the reference's own
//     Camera / Scene / Material / Light / TriangleMesh / ProxyObject / MBObject API
//     exactly the way the make*Scene() functions do (src/assignment2.h:379-438,
//     src/main.cpp:37-52, src/Assignment3.h);
//   * makeMeshObjs / makeMBMeshObjs, which the reference declares and calls
//     (src/main.cpp:22-23) but never defines (template: src/assignment2.h:717-732,
//     src/ProxyObject.cpp:131-146 — one Object per triangle, reverse index order);
//   * dump / trace / render / timing modes used to create golden vectors.
//
// Documented deviation from the reference: the RNG blocks of threads 1..31 are
// pre-filled (Scene::genRands(t)); upstream only fills thread 0's block in the
// Scene ctor (src/Scene.cpp:23), so other threads would read 65 536 zeros first.

#define protected public   // Scene::m_bvh, BVH::m_baseQNode are protected (src/Scene.h:72, src/BVH.h:149)
#define private public
#include "Miro.h"
#include "Scene.h"
#include "Camera.h"
#include "Image.h"
#include "PointLight.h"
#include "RectangleLight.h"
#include "DomeLight.h"
#include "Object.h"
#include "ProxyObject.h"
#include "MBObject.h"
#include "TriangleMesh.h"
#include "Lambert.h"
#include "Blinn.h"
#include "RawImage.h"
#include "Texture.h"
#include "BVH.h"
#undef protected
#undef private

#include <omp.h>
#include <map>
#include <string>
#include <vector>
#include <sstream>
#include <fstream>
#include <cstdio>
#include <cstdint>
#include <dlfcn.h>
#include "miro_gpu.h"      // include/miro_gpu.h: the C ABI the --render-gpu glue binds (dlopen, no link-time dependency)

unsigned long long g_miro_trace_calls[32 * 16] = {0};
void ParseFile(FILE*) {}
void initOpenGL() {}

namespace {

struct MeshRec { std::string name; TriangleMesh* mesh; int ordinal; };
struct BlasRec { Objects* objs; BVH* bvh; };

std::map<std::string, MeshRec> g_meshes;
std::vector<TriangleMesh*> g_meshByOrdinal;
std::map<TriangleMesh*, int> g_meshOrdinal;
std::map<std::string, Material*> g_materials;
std::map<std::string, Texture*> g_textures;
std::map<std::string, BlasRec> g_blas;
std::map<const ProxyObject*, int> g_proxyOrdinal;
std::string g_assetRoot = ".";

std::string assetPath(const std::string& p) {
    if (!p.empty() && p[0] == '/') return p;
    return g_assetRoot + "/" + p;
}

void makeMeshObjs(TriangleMesh* mesh, Material* mat) {
    int n = mesh->m_numTris;
    Object* t = new Object[n];
    for (int i = n - 1; i >= 0; --i) {
        t[i].setMesh(mesh); t[i].setIndex(i); t[i].setMaterial(mat);
        g_scene->addObject(&t[i]);
    }
}

void makeMBMeshObjs(TriangleMesh* mesh, TriangleMesh* mesh2, Material* mat) {
    int n = mesh->m_numTris;
    for (int i = n - 1; i >= 0; --i) g_scene->addObject(new MBObject(mat, mesh, mesh2, i));
}

Vector3 read3(std::istringstream& ss) { float x, y, z; ss >> x >> y >> z; return Vector3(x, y, z); }

void die(const std::string& m) { fprintf(stderr, "miro_ref: %s\n", m.c_str()); exit(2); }

Texture* getTexture(const std::string& name) {
    if (!g_textures.count(name)) die("unknown texture " + name);
    return g_textures[name];
}
Material* getMaterial(const std::string& name) {
    if (!g_materials.count(name)) die("unknown material " + name);
    return g_materials[name];
}
TriangleMesh* getMesh(const std::string& name) {
    if (!g_meshes.count(name)) die("unknown mesh " + name);
    return g_meshes[name].mesh;
}

void loadScene(const std::string& file) {
    std::ifstream in(file.c_str());
    if (!in) die("cannot open scene " + file);
    g_camera = new Camera; g_scene = new Scene; g_image = new Image;
    g_image->resize(512, 512);
    Vector3 bg(0.f); g_scene->setBGColor(bg);
    std::string line;
    while (std::getline(in, line)) {
        size_t h = line.find('#'); if (h != std::string::npos) line = line.substr(0, h);
        std::istringstream ss(line);
        std::string cmd; if (!(ss >> cmd)) continue;
        if (cmd == "image") { int w, hgt; ss >> w >> hgt; g_image->resize(w, hgt); }
        else if (cmd == "camera") {
            std::string k;
            while (ss >> k) {
                if (k == "eye") g_camera->setEye(read3(ss));
                else if (k == "lookat") g_camera->setLookAt(read3(ss));
                else if (k == "viewdir") g_camera->setViewDir(read3(ss));
                else if (k == "up") g_camera->setUp(read3(ss));
                else if (k == "fov") { float f; ss >> f; g_camera->setFOV(f); }
                else if (k == "focus") { float f; ss >> f; g_camera->setFocusPlane(f); }
                else if (k == "aperture") { float f; ss >> f; g_camera->setAperture(f); }
                else if (k == "shutter") { float f; ss >> f; g_camera->setShutterSpeed(f); }
                else die("camera: unknown key " + k);
            }
        }
        else if (cmd == "scene") {
            std::string k;
            while (ss >> k) {
                if (k == "bgcolor") { Vector3 c = read3(ss); g_scene->setBGColor(c); }
                else if (k == "pathtrace") { int v; ss >> v; g_scene->setPathTrace(v != 0); }
                else if (k == "numpaths") { int v; ss >> v; g_scene->setNumPaths(v); }
                else if (k == "maxbounces") { int v; ss >> v; g_scene->setMaxBounces(v); }
                else if (k == "minsubdivs") { int v; ss >> v; g_scene->setMinSubdivs(v); }
                else if (k == "maxsubdivs") { int v; ss >> v; g_scene->setMaxSubdivs(v); }
                else if (k == "noise") { float v; ss >> v; g_scene->setNoise(v); }
                else if (k == "sampleenv") { int v; ss >> v; g_scene->setSampleEnv(v != 0); }
                else if (k == "envmap") { std::string t; float e; ss >> t >> e; g_scene->setEnvMap(getTexture(t)); g_scene->setEnvExposure(e); }
                else if (k == "seed" || k == "devicebuild") { std::string ignored; ss >> ignored; }      // product-side options
                else die("scene: unknown key " + k);
            }
        }
        else if (cmd == "texture") {
            std::string name, path; ss >> name >> path;
            RawImage* img = new RawImage();
            std::string full = assetPath(path);
            img->m_rawData = 0; img->m_width = 0; img->m_height = 0;
            img->loadImage((char*)full.c_str());
            if (!img->m_rawData || img->m_width <= 0) die("cannot load texture " + full);
            g_textures[name] = new Texture(img);
        }
        else if (cmd == "material") {
            std::string name, kind; ss >> name >> kind;
            std::string k;
            if (kind == "lambert") {
                Lambert* m = new Lambert(Vector3(1.f), Vector3(0.f));
                while (ss >> k) {
                    if (k == "kd") m->setKd(read3(ss));
                    else if (k == "ka") m->setKa(read3(ss));
                    else if (k == "colormap") { std::string t; ss >> t; m->setColorMap(getTexture(t)); }
                    else die("lambert: unknown key " + k);
                }
                g_materials[name] = m;
            } else if (kind == "blinn") {
                Blinn* m = new Blinn(Vector3(1.f));
                while (ss >> k) {
                    if (k == "kd") m->setKd(read3(ss));
                    else if (k == "ka") m->setKa(read3(ss));
                    else if (k == "ks") m->setKs(read3(ss));
                    else if (k == "specexp") { float f; ss >> f; m->setSpecExp(f); }
                    else if (k == "specamt") { float f; ss >> f; m->setSpecAmt(f); }
                    else if (k == "ior") { float f; ss >> f; m->setIor(f, 0); m->setIor(f, 1); m->setIor(f, 2); }
                    else if (k == "ior_i") { int i; float f; ss >> i >> f; m->setIor(f, i); }
                    else if (k == "disperse") { int v; ss >> v; m->m_disperse = v != 0; }
                    else if (k == "reflect") { float f; ss >> f; m->setReflectAmt(f); }
                    else if (k == "refract") { float f; ss >> f; m->setRefractAmt(f); }
                    else if (k == "gloss") { float f; ss >> f; m->setReflectGloss(f); }
                    else if (k == "translucency") { float f; ss >> f; m->setTranslucency(f); }
                    else if (k == "emit") { float i; ss >> i; Vector3 c = read3(ss); m->setLightEmittedIntensity(i); m->setLightEmittedColor(c); }
                    else if (k == "colormap") { std::string t; ss >> t; m->setColorMap(getTexture(t)); }
                    else if (k == "alphamap") { std::string t; ss >> t; m->setAlphaMap(getTexture(t)); }
                    else if (k == "normalmap") { std::string t; ss >> t; m->setNormalMap(getTexture(t)); }
                    else if (k == "specularmap") { std::string t; ss >> t; m->setSpecularMap(getTexture(t)); }
                    else if (k == "reflectmap") { std::string t; ss >> t; m->setReflectMap(getTexture(t)); }
                    else if (k == "refractmap") { std::string t; ss >> t; m->setRefractMap(getTexture(t)); }
                    else if (k == "sampleenv") { int v; ss >> v; m->setSampleEnv(v != 0); }
                    else die("blinn: unknown key " + k);
                }
                g_materials[name] = m;
            } else die("unknown material kind " + kind);
        }
        else if (cmd == "light") {
            std::string kind; ss >> kind; std::string k;
            if (kind == "point") {
                PointLight* l = new PointLight; l->setColor(Vector3(1, 1, 1));
                while (ss >> k) {
                    if (k == "pos") l->setPosition(read3(ss));
                    else if (k == "power") { float f; ss >> f; l->setPower(f); }
                    else if (k == "shadows") { int v; ss >> v; l->setCastShadows(v != 0); }
                    else if (k == "fastshadows") { int v; ss >> v; l->setFastShadows(v != 0); }
                    else die("point light: unknown key " + k);
                }
                g_scene->addLight(l);
            } else if (kind == "rect") {
                RectangleLight* l = new RectangleLight; l->setColor(Vector3(1, 1, 1));
                Vector3 v1(0.f), v2(0.f), v3(0.f); float power = 0.f;
                while (ss >> k) {
                    if (k == "v1") v1 = read3(ss);
                    else if (k == "v2") v2 = read3(ss);
                    else if (k == "v3") v3 = read3(ss);
                    else if (k == "power") ss >> power;
                    else if (k == "samples") { int n; ss >> n; l->setSamples(n); }
                    else if (k == "noise") { float f; ss >> f; l->setNoiseThreshold(f); }
                    else if (k == "shadows") { int v; ss >> v; l->setCastShadows(v != 0); }
                    else if (k == "fastshadows") { int v; ss >> v; l->setFastShadows(v != 0); }
                    else die("rect light: unknown key " + k);
                }
                // same call order as the scene functions: setPower, then setVertices (assignment2.h:404-405)
                l->setPower(power); l->setVertices(v1, v2, v3);
                g_scene->addLight(l);
            } else if (kind == "dome") {
                DomeLight* l = new DomeLight;
                while (ss >> k) {
                    if (k == "tex") { std::string t; ss >> t; l->setTexture(getTexture(t)); }
                    else if (k == "power") { float f; ss >> f; l->setPower(f); }
                    else if (k == "samples") { int n; ss >> n; l->setSamples(n); }
                    else if (k == "noise") { float f; ss >> f; l->setNoiseThreshold(f); }
                    else if (k == "fastshadows") { int v; ss >> v; l->setFastShadows(v != 0); }
                    else die("dome light: unknown key " + k);
                }
                g_scene->addLight(l);
            } else die("unknown light kind " + kind);
        }
        else if (cmd == "mesh") {
            std::string name, path; ss >> name >> path;
            Matrix4x4 ctm; std::string k;
            if (ss >> k) {
                if (k != "ctm") die("mesh: expected ctm");
                float m[16]; for (int i = 0; i < 16; i++) ss >> m[i];
                for (int i = 0; i < 4; i++) { ctm.m1[i] = m[i]; ctm.m2[i] = m[4 + i]; ctm.m3[i] = m[8 + i]; ctm.m4[i] = m[12 + i]; }
            }
            TriangleMesh* mesh = new TriangleMesh;
            mesh->m_tangents = 0; mesh->m_biTangents = 0;
            std::string full = assetPath(path);
            if (!mesh->load((char*)full.c_str(), ctm)) die("cannot load mesh " + full);
            MeshRec r; r.name = name; r.mesh = mesh; r.ordinal = (int)g_meshByOrdinal.size();
            g_meshes[name] = r; g_meshByOrdinal.push_back(mesh); g_meshOrdinal[mesh] = r.ordinal;
        }
        else if (cmd == "object") { std::string m, mat; ss >> m >> mat; makeMeshObjs(getMesh(m), getMaterial(mat)); }
        else if (cmd == "mbobject") { std::string m1, m2, mat; ss >> m1 >> m2 >> mat; makeMBMeshObjs(getMesh(m1), getMesh(m2), getMaterial(mat)); }
        else if (cmd == "blas") {
            std::string name; ss >> name;
            std::vector<TriangleMesh*> ms; std::vector<Material*> mats; std::string m, mat;
            while (ss >> m >> mat) { ms.push_back(getMesh(m)); mats.push_back(getMaterial(mat)); }
            BlasRec b; b.objs = new Objects; b.bvh = new BVH;
            if (ms.size() == 1) ProxyObject::setupProxy(ms[0], mats[0], b.objs, b.bvh);
            else ProxyObject::setupMultiProxy(&ms[0], (int)ms.size(), &mats[0], b.objs, b.bvh);
            g_blas[name] = b;
        }
        else if (cmd == "instance") {
            std::string name; ss >> name;
            if (!g_blas.count(name)) die("unknown blas " + name);
            float m[16]; for (int i = 0; i < 16; i++) ss >> m[i];
            Matrix4x4 M;
            for (int i = 0; i < 4; i++) { M.m1[i] = m[i]; M.m2[i] = m[4 + i]; M.m3[i] = m[8 + i]; M.m4[i] = m[12 + i]; }
            ProxyObject* po = new ProxyObject(g_blas[name].objs, g_blas[name].bvh, M);
            po->setDisplayNum(1000);
            int ord = (int)g_proxyOrdinal.size(); g_proxyOrdinal[po] = ord;
            g_scene->addObject(po);
        }
        else die("unknown command " + cmd);
    }
    g_scene->preCalc();
    for (int t = 1; t < 32; ++t) Scene::genRands(t);   // documented deviation (see header)
}

#pragma pack(push, 1)
struct RayRec { float ox, oy, oz, tmin, dx, dy, dz, tmax, time; uint32_t flags, pad0, pad1; };
struct RefHit { float t, a, b; int32_t mesh, tri, proxy; };
#pragma pack(pop)

RefHit toRefHit(bool hit, const HitInfo& h) {
    RefHit r; r.t = h.t; r.a = h.a; r.b = h.b; r.mesh = r.tri = r.proxy = -1;
    if (hit && h.obj) {
        r.mesh = g_meshOrdinal.count(h.obj->m_mesh) ? g_meshOrdinal[h.obj->m_mesh] : -2;
        r.tri = (int)h.obj->m_index;
        if (h.m_proxy) r.proxy = g_proxyOrdinal[h.m_proxy];
    } else { r.t = -1.f; r.a = r.b = 0.f; }
    return r;
}

unsigned long long traceCalls() { unsigned long long s = 0; for (int i = 0; i < 32; i++) s += g_miro_trace_calls[i * 16]; return s; }
void resetTraceCalls() { for (int i = 0; i < 32 * 16; i++) g_miro_trace_calls[i] = 0; }

template <class T> void writeVec(const std::string& path, const std::vector<T>& v) {
    FILE* f = fopen(path.c_str(), "wb"); if (!f) die("cannot write " + path);
    if (!v.empty()) fwrite(&v[0], sizeof(T), v.size(), f); fclose(f);
}
template <class T> std::vector<T> readVec(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb"); if (!f) die("cannot read " + path);
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T)); if (!v.empty() && fread(&v[0], sizeof(T), v.size(), f) != v.size()) die("short read " + path);
    fclose(f); return v;
}

// --dump-primary: the reference's own camera rays at pixel centres (Camera.cpp:116-174 with
// offsets 0.5..0.5), row 0 = bottom.  One ray per pixel.
void dumpPrimary(const std::string& out) {
    int w = g_image->width(), h = g_image->height();
    std::vector<RayRec> rays((size_t)w * h);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        Ray r = g_camera->eyeRayAdaptive(0, x, y, 0.5f, 0.5f, 0.5f, 0.5f, w, h);
        RayRec& q = rays[(size_t)y * w + x];
        q.ox = r.o[0]; q.oy = r.o[1]; q.oz = r.o[2]; q.tmin = epsilon;
        q.dx = r.d[0]; q.dy = r.d[1]; q.dz = r.d[2]; q.tmax = MIRO_TMAX;
        q.time = r.time; q.flags = 0; q.pad0 = q.pad1 = 0;
    }
    writeVec(out, rays);
}

// --trace: Scene::trace (Scene.cpp:295) over a caller-supplied ray buffer.  Timed passes run on `threads` OpenMP threads; the
// hit records that are written come from one more, SINGLE-threaded pass when threads > 1 — concurrent traversals corrupt each
// other through QBVH_Node::boxHit, a mutable member of the shared node (BVH.h:101, BVH.cpp:413,1151), so only a
// single-threaded pass yields the reference's true answers.
double g_traceMean = 0.0;       // mean seconds over the timed repeats of the last traceBatch call
void tracePass(const std::vector<RayRec>& rays, std::vector<RefHit>& hits, int threads) {
    #pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
    for (long i = 0; i < (long)rays.size(); i++) {
        unsigned tid = omp_get_thread_num();
        const RayRec& q = rays[i];
        Ray r(tid, Vector3(q.ox, q.oy, q.oz), Vector3(q.dx, q.dy, q.dz), q.time);
        HitInfo h; h.t = q.tmax;
        bool hit = g_scene->trace(tid, h, r, q.tmin);
        hits[i] = toRefHit(hit, h);
    }
}
double traceBatch(const std::vector<RayRec>& rays, std::vector<RefHit>& hits, int threads, int repeat, int warmup, bool wantHits) {
    hits.resize(rays.size());
    double best = 1e30, sum = 0.0;
    for (int it = -warmup; it < repeat; ++it) {
        double t0 = omp_get_wtime();
        tracePass(rays, hits, threads);
        double t1 = omp_get_wtime();
        if (it >= 0) { if (t1 - t0 < best) best = t1 - t0; sum += t1 - t0; }
    }
    g_traceMean = sum / (repeat > 0 ? repeat : 1);
    if (wantHits && threads > 1) tracePass(rays, hits, 1);
    return best;
}

// --shadow-light X Y Z FIRST COUNT: the shadow rays PointLight::sampleLight would cast (PointLight.cpp:20-48) from the hits of
// rays [FIRST, FIRST + COUNT) of the traced batch towards a point light — from the hit point, or from the ray origin for a miss;
// tMin = epsilon, tMax = the distance to the light — traced as a second timed batch.
void shadowRaysFromHits(const std::vector<RayRec>& rays, const std::vector<RefHit>& hits, size_t first, size_t count, const float light[3],
                        std::vector<RayRec>& out) {
    out.clear();
    for (size_t i = first; i < first + count && i < rays.size(); i++) {
        const RayRec& q = rays[i];
        const float t = hits[i].mesh >= 0 ? hits[i].t : 0.0f;
        RayRec s = q;
        s.ox = q.ox + t * q.dx; s.oy = q.oy + t * q.dy; s.oz = q.oz + t * q.dz;
        const float lx = light[0] - s.ox, ly = light[1] - s.oy, lz = light[2] - s.oz;
        const float dist = sqrtf(lx * lx + ly * ly + lz * lz), inv = 1.0f / (dist > 1e-20f ? dist : 1e-20f);
        s.dx = lx * inv; s.dy = ly * inv; s.dz = lz * inv; s.tmin = epsilon; s.tmax = dist;
        out.push_back(s);
    }
}

// --render-float: the reference's per-pixel entry point (Scene.cpp:252 adaptiveSampleScene) over
// the reference's bucket order (Scene.cpp:160-175), radiance kept as float before Image::Map.
double renderFloat(const std::string& out, int threads) {
    int w = g_image->width(), h = g_image->height();
    std::vector<float> img((size_t)w * h * 3);
    int nbx = (w + bucket_size - 1) / bucket_size, nby = (h + bucket_size - 1) / bucket_size;
    double t0 = omp_get_wtime();
    #pragma omp parallel num_threads(threads)
    {
        unsigned tid = omp_get_thread_num();
        Ray ray(tid); HitInfo hit;
        #pragma omp for schedule(dynamic)
        for (int b = 0; b < nbx * nby; b++) {
            int bx = b % nbx, by = b / nbx;
            for (int j = by * bucket_size; j < std::min((by + 1) * bucket_size, h); ++j)
                for (int i = bx * bucket_size; i < std::min((bx + 1) * bucket_size, w); ++i) {
                    Vector3 c = g_scene->adaptiveSampleScene(tid, g_camera, g_image, ray, hit, i, j);
                    float* p = &img[((size_t)j * w + i) * 3]; p[0] = c.x; p[1] = c.y; p[2] = c.z;
                }
        }
    }
    double t1 = omp_get_wtime();
    if (!out.empty()) writeVec(out, img);
    return t1 - t0;
}

// --dump-meshes DIR: geometry exactly as the reference's loader left it (TriangleMeshLoad.cpp:100-214).
void dumpMeshes(const std::string& dir) {
    for (std::map<std::string, MeshRec>::iterator it = g_meshes.begin(); it != g_meshes.end(); ++it) {
        TriangleMesh* m = it->second.mesh;
        int nf = m->m_numTris; uint32_t maxv = 0, maxn = 0, maxt = 0;
        for (int i = 0; i < nf; i++) {
            maxv = std::max(maxv, std::max(m->m_vertexIndices[i].x, std::max(m->m_vertexIndices[i].y, m->m_vertexIndices[i].z)));
            maxn = std::max(maxn, std::max(m->m_normalIndices[i].x, std::max(m->m_normalIndices[i].y, m->m_normalIndices[i].z)));
            if (m->m_texCoordIndices) maxt = std::max(maxt, std::max(m->m_texCoordIndices[i].x, std::max(m->m_texCoordIndices[i].y, m->m_texCoordIndices[i].z)));
        }
        int nv = maxv + 1, nn = maxn + 1, nt = m->m_texCoordIndices ? (int)maxt + 1 : 0;
        std::string path = dir + "/" + it->first + ".mesh";
        FILE* f = fopen(path.c_str(), "wb"); if (!f) die("cannot write " + path);
        int32_t hdr[5] = {it->second.ordinal, nv, nn, nt, nf}; fwrite(hdr, 4, 5, f);
        for (int i = 0; i < nv; i++) fwrite(&m->m_vertices[i].x, 4, 3, f);
        for (int i = 0; i < nn; i++) fwrite(&m->m_normals[i].x, 4, 3, f);
        for (int i = 0; i < nt; i++) fwrite(&m->m_texCoords[i].x, 4, 2, f);
        fwrite(m->m_vertexIndices, 12, nf, f);
        fwrite(m->m_normalIndices, 12, nf, f);
        if (nt) fwrite(m->m_texCoordIndices, 12, nf, f);
        fclose(f);
    }
}

// --dump-instances FILE: per ProxyObject, in ordinal order, the 16 floats of its ProxyMatrix::m_inverse (row-major) and what
// multiplyAndDivideByW multiplies a transformed point by for an affine matrix, recipps(m44) — the numbers the product's host layer
// must reproduce bit for bit (host/miro_math.h invertedAsReference, referenceRecip).
void dumpInstances(const std::string& path) {
    std::vector<const ProxyObject*> byOrdinal(g_proxyOrdinal.size(), (const ProxyObject*)0);
    for (std::map<const ProxyObject*, int>::iterator it = g_proxyOrdinal.begin(); it != g_proxyOrdinal.end(); ++it) byOrdinal[it->second] = it->first;
    std::vector<float> out;
    for (size_t i = 0; i < byOrdinal.size(); i++) {
        const Matrix4x4& I = byOrdinal[i]->getMatrix().m_inverse;
        const float m[16] = {I.m11, I.m12, I.m13, I.m14, I.m21, I.m22, I.m23, I.m24, I.m31, I.m32, I.m33, I.m34, I.m41, I.m42, I.m43, I.m44};
        out.insert(out.end(), m, m + 16);
        __attribute__((aligned(16))) float w[4] = {I.m44, I.m44, I.m44, I.m44}; __attribute__((aligned(16))) float r[4];
        storeps(recipps(loadps(w)), r);
        out.push_back(r[0]);
    }
    writeVec(path, out);
}

// --dump-textures DIR: every texture as the reference's loaders left it (RawImage.cpp / hdrloader.cpp): float texels.
void dumpTextures(const std::string& dir) {
    for (std::map<std::string, Texture*>::iterator it = g_textures.begin(); it != g_textures.end(); ++it) {
        RawImage* im = it->second->m_image;
        int ch = im->m_imageType == GRAYSCALE ? 1 : (im->m_imageType == RGBA ? 4 : 3);
        std::string path = dir + "/" + it->first + ".tex";
        FILE* f = fopen(path.c_str(), "wb"); if (!f) die("cannot write " + path);
        int32_t hdr[4] = {im->m_width, im->m_height, ch, (int32_t)im->m_imageType}; fwrite(hdr, 4, 4, f);
        fwrite(im->m_rawData, 4, (size_t)im->m_width * im->m_height * ch, f);
        fclose(f);
    }
}

// --dump-qbvh FILE: the reference's own QBVH (BVH.cpp:100-389) walked and flattened — the data a
// maintainer's flatten() would hand to miro_gpu_upload_scene.  Format:
//   int32 nNodes, nLeaves; nodes: 24 float bounds (minX[4] minY[4] minZ[4] maxX[4] maxY[4] maxZ[4]),
//   int32 child[4] (>=0 node index, <0 = ~leafIndex, INT32_MIN = invalid); leaves: 4 x {int32 mesh, tri, kind}.
struct FlatQ { std::vector<float> bounds; std::vector<int32_t> child; std::vector<int32_t> leaves; };
int flattenQ(const QBVH_Node* n, FlatQ& out) {
    int idx = (int)out.child.size() / 4;
    out.child.resize(out.child.size() + 4, INT32_MIN);
    out.bounds.resize(out.bounds.size() + 24);
    float* b = &out.bounds[(size_t)idx * 24];
    memcpy(b, n->bbMinX, 16); memcpy(b + 4, n->bbMinY, 16); memcpy(b + 8, n->bbMinZ, 16);
    memcpy(b + 12, n->bbMaxX, 16); memcpy(b + 16, n->bbMaxY, 16); memcpy(b + 20, n->bbMaxZ, 16);
    for (int i = 0; i < 4; i++) {
        if (n->flagsIsLeaf[i]) {
            int li = (int)out.leaves.size() / 12;
            const BVH_Node::TriCache4* tc = n->triCaches[i];
            for (int k = 0; k < 4; k++) {
                Object* o = tc->tris[k];
                out.leaves.push_back(o ? (g_meshOrdinal.count(o->m_mesh) ? g_meshOrdinal[o->m_mesh] : -2) : -1);
                out.leaves.push_back(o ? (int)o->m_index : -1);
                out.leaves.push_back(o ? (int)o->m_objectType : -1);
            }
            out.child[(size_t)idx * 4 + i] = ~li;
        } else if (n->flagsIsValid[i]) {
            int c = flattenQ(n->Children[i], out);
            out.child[(size_t)idx * 4 + i] = c;
        }
    }
    return idx;
}
void dumpQBVH(const std::string& path) {
    FlatQ q; flattenQ(g_scene->m_bvh.m_baseQNode, q);
    FILE* f = fopen(path.c_str(), "wb"); if (!f) die("cannot write " + path);
    int32_t hdr[2] = {(int32_t)(q.child.size() / 4), (int32_t)(q.leaves.size() / 12)};
    fwrite(hdr, 4, 2, f); fwrite(&q.bounds[0], 4, q.bounds.size(), f); fwrite(&q.child[0], 4, q.child.size(), f);
    if (!q.leaves.empty()) fwrite(&q.leaves[0], 4, q.leaves.size(), f);
    fclose(f);
}

// --render-gpu FILE --gpu-lib LIB: the binding of INTEGRATION.md made real.  The reference's OWN scene — its Object list, its
// materials and lights as the scene code set them up, and its OWN QBVH (BVH.cpp:100-389) — is flattened into a
// miro_gpu_scene_desc, handed to libmiro_gpu.so (loaded with dlopen: this harness has no link-time dependency on CUDA) through
// miro_gpu_upload_scene, and Scene::raytraceImage's float radiance comes back from miro_gpu_render.  Covers what the reference's
// scene code can build: Objects, MBObjects, ProxyObjects (one level, shared bottom-level trees), Lambert / Blinn with their
// texture maps, point / rectangle / dome lights, the environment map.
struct GpuFlat {
    std::vector<miro_gpu_node> nodes; std::vector<miro_gpu_tri> tris; std::vector<miro_gpu_mbtri> mb; std::vector<miro_gpu_instance> inst;
    std::vector<miro_gpu_prim> prims, mbprims;
    std::vector<float> normals, tangents, bitangents, uvs, nxf;
    bool anyTangents;
    std::map<TriangleMesh*, std::pair<uint32_t, uint32_t> > meshBase;      // first normal / first uv of a mesh
    std::vector<miro_gpu_material> materials; std::map<const Material*, uint32_t> matOrdinal;
    std::vector<miro_gpu_texture> textures; std::map<const Texture*, int32_t> texOrdinal;
    std::map<const BVH*, int32_t> blasRoot;
    GpuFlat() : anyTangents(false) {}
};
int32_t gpuFlattenQ(const QBVH_Node* n, GpuFlat& f);
int32_t gpuTexture(const Texture* t, GpuFlat& f) {              // RawImage::m_rawData: float texels, row-major (borrowed, not copied)
    if (!t) return -1;
    if (f.texOrdinal.count(t)) return f.texOrdinal[t];
    const RawImage* im = t->m_image;
    miro_gpu_texture g; memset(&g, 0, sizeof(g));
    g.texels = im->m_rawData; g.width = im->m_width; g.height = im->m_height;
    g.channels = im->m_imageType == GRAYSCALE ? 1 : (im->m_imageType == RGBA ? 4 : 3);
    const int32_t k = (int32_t)f.textures.size(); f.textures.push_back(g); f.texOrdinal[t] = k;
    return k;
}
uint32_t gpuMaterial(const Material* m, GpuFlat& f) {
    if (f.matOrdinal.count(m)) return f.matOrdinal[m];
    miro_gpu_material g; memset(&g, 0, sizeof(g));
    g.color_map = gpuTexture(m->m_colorMap, f); g.alpha_map = gpuTexture(m->m_alphaMap, f);
    g.normal_map = g.specular_map = g.reflect_map = g.refract_map = -1;
    g.translucency = m->m_translucency; g.sample_env = m->m_sampleEnv ? 1u : 0u; g.disperse = m->m_disperse ? 1u : 0u; g.spec_gloss = 1.0f;
    g.refract_amt = m->m_refractAmt;                              // Material's member: the full shadow method reads it of any material
    if (const Blinn* b = dynamic_cast<const Blinn*>(m)) {
        g.kind = MIRO_GPU_MAT_BLINN;
        g.normal_map = gpuTexture(m->m_normalMap, f); g.specular_map = gpuTexture(m->m_specularMap, f);
        g.reflect_map = gpuTexture(m->m_reflectMap, f); g.refract_map = gpuTexture(m->m_refractMap, f);
        g.kd[0] = b->m_kd.x; g.kd[1] = b->m_kd.y; g.kd[2] = b->m_kd.z; g.ka[0] = b->m_ka.x; g.ka[1] = b->m_ka.y; g.ka[2] = b->m_ka.z;
        g.ks[0] = b->m_ks.x; g.ks[1] = b->m_ks.y; g.ks[2] = b->m_ks.z;
        g.spec_exp = b->m_specExp; g.spec_amt = b->m_specAmt; g.emit_intensity = b->m_lightEmitted;
        g.le[0] = b->m_Le.x; g.le[1] = b->m_Le.y; g.le[2] = b->m_Le.z;
        g.reflect_amt = b->m_reflectAmt; g.refract_amt = b->m_refractAmt; g.spec_gloss = b->m_specGloss;
        g.ior[0] = b->m_ior[0]; g.ior[1] = b->m_ior[1]; g.ior[2] = b->m_ior[2];
    } else if (const Lambert* l = dynamic_cast<const Lambert*>(m)) {
        g.kind = MIRO_GPU_MAT_LAMBERT;                             // Lambert::shade reads the colour map only (Lambert.cpp:19-53)
        g.refract_amt = 0.f;                                       // (the reference leaves Material::m_refractAmt of a Lambert uninitialised)
        g.kd[0] = l->m_kd.x; g.kd[1] = l->m_kd.y; g.kd[2] = l->m_kd.z; g.ka[0] = l->m_ka.x; g.ka[1] = l->m_ka.y; g.ka[2] = l->m_ka.z;
    } else die("--render-gpu: unknown material class");
    const uint32_t k = (uint32_t)f.materials.size(); f.materials.push_back(g); f.matOrdinal[m] = k;
    return k;
}
miro_gpu_tri gpuTri(const TriangleMesh* m, u_int i) {
    const TriangleMesh::TupleI3 vi = m->m_vertexIndices[i];
    miro_gpu_tri t; memset(&t, 0, sizeof(t));
    const Vector3 &a = m->m_vertices[vi.x], &b = m->m_vertices[vi.y], &c = m->m_vertices[vi.z];
    t.v0[0] = a.x; t.v0[1] = a.y; t.v0[2] = a.z; t.v1[0] = b.x; t.v1[1] = b.y; t.v1[2] = b.z; t.v2[0] = c.x; t.v2[1] = c.y; t.v2[2] = c.z;
    return t;
}
// One Object of the reference -> its place in the description (INTEGRATION.md's appendPrimitive).
void gpuAppendPrimitive(const Object* o, GpuFlat& f) {
    if (o->m_objectType == PROXY_OBJECT) {
        const ProxyObject* p = static_cast<const ProxyObject*>(o);
        if (!f.blasRoot.count(p->m_BVH)) f.blasRoot[p->m_BVH] = gpuFlattenQ(p->m_BVH->m_baseQNode, f);      // the shared bottom-level tree, once
        const ProxyMatrix& M = p->getMatrix();
        miro_gpu_instance in; memset(&in, 0, sizeof(in));
        const Matrix4x4& I = M.m_inverse; const Matrix4x4& T = M.m_invTranspose;
        const float inv[12] = {I.m11, I.m12, I.m13, I.m14, I.m21, I.m22, I.m23, I.m24, I.m31, I.m32, I.m33, I.m34};
        memcpy(in.inv, inv, sizeof(inv)); in.blas_root = f.blasRoot[p->m_BVH];
        { __attribute__((aligned(16))) float one[4] = {I.m44, I.m44, I.m44, I.m44}; __attribute__((aligned(16))) float r[4]; storeps(recipps(loadps(one)), r); in.w_recip = r[0]; }      // what multiplyAndDivideByW multiplies by: w = m44 for an affine matrix
        const float nx[9] = {T.m11, T.m12, T.m13, T.m21, T.m22, T.m23, T.m31, T.m32, T.m33};
        f.nxf.insert(f.nxf.end(), nx, nx + 9);
        f.inst.push_back(in);
        return;
    }
    TriangleMesh* m = o->m_mesh;
    if (!f.meshBase.count(m)) {                  // the mesh's normals (tangents, bitangents) / uvs are appended once (counts: highest index used)
        uint32_t maxn = 0, maxt = 0;
        for (u_int i = 0; i < m->m_numTris; i++) {
            maxn = std::max(maxn, std::max(m->m_normalIndices[i].x, std::max(m->m_normalIndices[i].y, m->m_normalIndices[i].z)));
            if (m->m_texCoordIndices) maxt = std::max(maxt, std::max(m->m_texCoordIndices[i].x, std::max(m->m_texCoordIndices[i].y, m->m_texCoordIndices[i].z)));
        }
        f.meshBase[m] = std::make_pair((uint32_t)(f.normals.size() / 3), (uint32_t)(f.uvs.size() / 2));
        const bool tb = m->m_texCoordIndices && m->m_tangents && m->m_biTangents;      // TriangleMesh::preCalc fills them for meshes with uvs
        if (tb) f.anyTangents = true;
        for (uint32_t i = 0; i <= maxn; i++) {
            f.normals.push_back(m->m_normals[i].x); f.normals.push_back(m->m_normals[i].y); f.normals.push_back(m->m_normals[i].z);
            f.tangents.push_back(tb ? m->m_tangents[i].x : 0.f); f.tangents.push_back(tb ? m->m_tangents[i].y : 0.f); f.tangents.push_back(tb ? m->m_tangents[i].z : 0.f);
            f.bitangents.push_back(tb ? m->m_biTangents[i].x : 0.f); f.bitangents.push_back(tb ? m->m_biTangents[i].y : 0.f); f.bitangents.push_back(tb ? m->m_biTangents[i].z : 0.f);
        }
        if (m->m_texCoordIndices) for (uint32_t i = 0; i <= maxt; i++) { f.uvs.push_back(m->m_texCoords[i].x); f.uvs.push_back(m->m_texCoords[i].y); }
    }
    const std::pair<uint32_t, uint32_t> base = f.meshBase[m];
    const TriangleMesh::TupleI3 ni = m->m_normalIndices[o->m_index];
    miro_gpu_prim p; memset(&p, 0, sizeof(p));
    p.n[0] = base.first + ni.x; p.n[1] = base.first + ni.y; p.n[2] = base.first + ni.z;
    if (m->m_texCoordIndices) { const TriangleMesh::TupleI3 ti = m->m_texCoordIndices[o->m_index]; p.uv[0] = base.second + ti.x; p.uv[1] = base.second + ti.y; p.uv[2] = base.second + ti.z; }
    else p.uv[0] = p.uv[1] = p.uv[2] = 0xffffffffu;
    p.material = gpuMaterial(o->m_material, f);
    p.mesh = g_meshOrdinal.count(m) ? (uint32_t)g_meshOrdinal[m] : 0u; p.tri = o->m_index;
    if (o->m_objectType == MB_OBJECT) {          // both poses; shading attributes come from the first mesh (Ray.cpp:12-25)
        const MBObject* mbo = static_cast<const MBObject*>(o);
        miro_gpu_mbtri t; t.pose[0] = gpuTri(m, o->m_index); t.pose[1] = gpuTri(mbo->m_mesh_t2, o->m_index);
        f.mb.push_back(t); f.mbprims.push_back(p);
    } else { f.tris.push_back(gpuTri(m, o->m_index)); f.prims.push_back(p); }
}
uint32_t gpuKind(const Object* o) { return o->m_objectType == MB_OBJECT ? MIRO_GPU_KIND_MBTRI : (o->m_objectType == PROXY_OBJECT ? MIRO_GPU_KIND_INST : MIRO_GPU_KIND_TRI); }
int32_t gpuFlattenQ(const QBVH_Node* n, GpuFlat& f) {               // QBVH_Node, BVH.h:89-104 -> miro_gpu_node, 1:1
    const int32_t me = (int32_t)f.nodes.size(); f.nodes.push_back(miro_gpu_node());
    miro_gpu_node out; memset(&out, 0, sizeof(out));
    memcpy(out.lo_x, n->bbMinX, 16); memcpy(out.lo_y, n->bbMinY, 16); memcpy(out.lo_z, n->bbMinZ, 16);
    memcpy(out.hi_x, n->bbMaxX, 16); memcpy(out.hi_y, n->bbMaxY, 16); memcpy(out.hi_z, n->bbMaxZ, 16);
    for (int i = 0; i < 4; i++) {
        out.child[i] = MIRO_GPU_CHILD_EMPTY;
        if (n->flagsIsLeaf[i]) {
            // One TriCache4 packet (BVH.cpp:64-98).  Its lanes may mix Objects, MBObjects and ProxyObjects; a leaf of the ABI holds one
            // kind, so a mixed packet becomes a small node of homogeneous leaves that all carry the packet's box.
            const BVH_Node::TriCache4* tc = n->triCaches[i];
            int32_t refs[3]; int nrefs = 0;
            for (uint32_t kind = 0; kind < 3; kind++) {
                const uint32_t first = kind == MIRO_GPU_KIND_TRI ? (uint32_t)f.tris.size() : kind == MIRO_GPU_KIND_MBTRI ? (uint32_t)f.mb.size() : (uint32_t)f.inst.size();
                uint32_t count = 0;
                for (int k = 0; k < 4; k++) if (tc->tris[k] && gpuKind(tc->tris[k]) == kind) { gpuAppendPrimitive(tc->tris[k], f); ++count; }
                if (count) refs[nrefs++] = MIRO_GPU_LEAF(kind, first, count);
            }
            if (nrefs == 1) out.child[i] = refs[0];
            else if (nrefs > 1) {
                const int32_t mid = (int32_t)f.nodes.size(); f.nodes.push_back(miro_gpu_node());
                miro_gpu_node m; memset(&m, 0, sizeof(m));
                for (int k = 0; k < 4; k++) {
                    m.child[k] = k < nrefs ? refs[k] : MIRO_GPU_CHILD_EMPTY;
                    m.lo_x[k] = n->bbMinX[i]; m.lo_y[k] = n->bbMinY[i]; m.lo_z[k] = n->bbMinZ[i];
                    m.hi_x[k] = n->bbMaxX[i]; m.hi_y[k] = n->bbMaxY[i]; m.hi_z[k] = n->bbMaxZ[i];
                }
                f.nodes[mid] = m; out.child[i] = mid;
            }
        } else if (n->flagsIsValid[i]) out.child[i] = gpuFlattenQ(n->Children[i], f);
    }
    f.nodes[me] = out;
    return me;
}
double renderGPU(const std::string& out, const std::string& libPath) {
    void* lib = dlopen(libPath.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) die(std::string("--render-gpu: cannot load ") + libPath + ": " + dlerror());
    // LIB is libmiro_gpu.so (the product).  For checking THIS GLUE where there is no GPU, LIB may also be the CPU oracle
    // (oracle/_build/libmiro_oracle.so exports oracle_render(desc, camera, params, rgb, mask)): same description, same image path.
    typedef uint64_t (*oracle_render_fn)(const miro_gpu_scene_desc*, const miro_gpu_camera*, const miro_gpu_render_params*, float*, const uint8_t*);
    oracle_render_fn p_oracle_render = (oracle_render_fn)dlsym(lib, "oracle_render");
    #define GPU_SYM(name) decltype(&name) p_##name = (decltype(&name))dlsym(lib, #name); if (!p_##name && !p_oracle_render) die("--render-gpu: symbol " #name " missing")
    GPU_SYM(miro_gpu_create); GPU_SYM(miro_gpu_destroy); GPU_SYM(miro_gpu_last_error); GPU_SYM(miro_gpu_abi_version);
    GPU_SYM(miro_gpu_upload_scene); GPU_SYM(miro_gpu_render); GPU_SYM(miro_gpu_get_counters);
    #undef GPU_SYM
    // ---- Scene::preCalc() has run (loadScene); flatten what it built
    GpuFlat f;
    const int32_t root = gpuFlattenQ(g_scene->m_bvh.m_baseQNode, f);
    std::vector<miro_gpu_light> lights;
    const Lights* ls = g_scene->lights();
    for (size_t i = 0; i < ls->size(); i++) {
        const Light* l = (*ls)[i];
        miro_gpu_light g; memset(&g, 0, sizeof(g));
        g.num_samples = l->m_numSamples; g.noise_threshold = l->m_noiseThreshold; g.cast_shadows = l->m_castShadows ? 1u : 0u;
        g.full_shadows = l->m_fastShadows ? 0u : 1u; g.texture = -1; g.power = l->m_power;
        if (const PointLight* pl = dynamic_cast<const PointLight*>(l)) {
            g.kind = MIRO_GPU_LIGHT_POINT; g.p0[0] = pl->m_position.x; g.p0[1] = pl->m_position.y; g.p0[2] = pl->m_position.z;
        } else if (const RectangleLight* rl = dynamic_cast<const RectangleLight*>(l)) {
            g.kind = MIRO_GPU_LIGHT_RECT;           // m_power already carries setPower's 1 / area (RectangleLight.cpp:39)
            g.p0[0] = rl->m_v1.x; g.p0[1] = rl->m_v1.y; g.p0[2] = rl->m_v1.z; g.p1[0] = rl->m_v2.x; g.p1[1] = rl->m_v2.y; g.p1[2] = rl->m_v2.z;
            g.p2[0] = rl->m_v3.x; g.p2[1] = rl->m_v3.y; g.p2[2] = rl->m_v3.z;
        } else if (const DomeLight* dl = dynamic_cast<const DomeLight*>(l)) {
            g.kind = MIRO_GPU_LIGHT_DOME; g.power = dl->m_Gain; g.texture = gpuTexture(dl->m_lightMap, f); g.cast_shadows = 1u;
        } else die("--render-gpu: unknown light class");
        lights.push_back(g);
    }
    miro_gpu_scene_desc d; memset(&d, 0, sizeof(d));
    d.abi_version = p_miro_gpu_abi_version ? (uint32_t)p_miro_gpu_abi_version() : (uint32_t)MIRO_GPU_ABI_VERSION;
    d.nodes = f.nodes.data(); d.n_nodes = (uint32_t)f.nodes.size(); d.root = root;
    f.prims.insert(f.prims.end(), f.mbprims.begin(), f.mbprims.end());      // prims[n_tris ..) describe the motion-blur triangles
    d.tris = f.tris.empty() ? NULL : f.tris.data(); d.n_tris = (uint32_t)f.tris.size(); d.prims = f.prims.data();
    d.mbtris = f.mb.empty() ? NULL : f.mb.data(); d.n_mbtris = (uint32_t)f.mb.size();
    d.instances = f.inst.empty() ? NULL : f.inst.data(); d.n_instances = (uint32_t)f.inst.size();
    d.inst_normal_xform = f.nxf.empty() ? NULL : f.nxf.data();
    d.tangents = f.anyTangents ? f.tangents.data() : NULL; d.bitangents = f.anyTangents ? f.bitangents.data() : NULL;
    d.normals = f.normals.data(); d.n_normals = (uint32_t)(f.normals.size() / 3);
    d.uvs = f.uvs.empty() ? NULL : f.uvs.data(); d.n_uvs = (uint32_t)(f.uvs.size() / 2);
    d.materials = f.materials.data(); d.n_materials = (uint32_t)f.materials.size();
    d.lights = lights.empty() ? NULL : lights.data(); d.n_lights = (uint32_t)lights.size();
    d.env_map = gpuTexture(g_scene->m_envMap, f); d.env_exposure = g_scene->m_envExposure;
    d.textures = f.textures.empty() ? NULL : f.textures.data(); d.n_textures = (uint32_t)f.textures.size();
    d.bg_color[0] = g_scene->m_BGColor.x; d.bg_color[1] = g_scene->m_BGColor.y; d.bg_color[2] = g_scene->m_BGColor.z;
    miro_gpu_ctx* ctx = NULL;
    if (!p_oracle_render) {
        if (p_miro_gpu_create(&ctx, 0)) die(std::string("miro_gpu_create: ") + p_miro_gpu_last_error(NULL));
        if (p_miro_gpu_upload_scene(ctx, &d)) die(std::string("miro_gpu_upload_scene: ") + p_miro_gpu_last_error(ctx));
    }
    // ---- Scene::raytraceImage(Camera*, Image*) (Scene.cpp:86-217)
    Camera* cam = g_camera;
    miro_gpu_camera c; memset(&c, 0, sizeof(c));
    c.eye[0] = cam->eye().x; c.eye[1] = cam->eye().y; c.eye[2] = cam->eye().z;
    c.view_dir[0] = cam->viewDir().x; c.view_dir[1] = cam->viewDir().y; c.view_dir[2] = cam->viewDir().z;
    c.up[0] = cam->up().x; c.up[1] = cam->up().y; c.up[2] = cam->up().z;
    c.fov_deg = cam->fov(); c.focus_plane = cam->focusPlane(); c.aperture = cam->aperture(); c.shutter_speed = cam->shutterSpeed();
    miro_gpu_render_params rp; memset(&rp, 0, sizeof(rp));
    rp.width = g_image->width(); rp.height = g_image->height();
    rp.min_subdivs = g_scene->m_minSubdivs; rp.max_subdivs = g_scene->m_maxSubdivs; rp.noise_threshold = g_scene->m_noiseThreshold;
    rp.num_paths = g_scene->m_numPaths; rp.max_bounces = g_scene->m_maxBounces; rp.path_trace = g_scene->m_pathTrace ? 1u : 0u;
    rp.sample_env = g_scene->m_sampleLightFromEnv ? 1u : 0u; rp.seed = 3163513; rp.shard_index = 0; rp.shard_count = 1;
    std::vector<float> rgb((size_t)rp.width * rp.height * 3);
    const double t0 = omp_get_wtime();
    unsigned long long rays = 0;
    if (p_oracle_render) rays = p_oracle_render(&d, &c, &rp, rgb.data(), NULL);
    else if (p_miro_gpu_render(ctx, &c, &rp, rgb.data())) die(std::string("miro_gpu_render: ") + p_miro_gpu_last_error(ctx));
    const double t1 = omp_get_wtime();
    for (int y = 0; y < rp.height; ++y) for (int x = 0; x < rp.width; ++x) {      // row 0 = bottom, as Image expects
        const float* q = &rgb[((size_t)y * rp.width + x) * 3];
        g_image->setPixel(x, y, Vector3(q[0], q[1], q[2]));                       // Image::Map: clamp + gamma LUT
    }
    if (ctx) { miro_gpu_counters ctr; memset(&ctr, 0, sizeof(ctr)); p_miro_gpu_get_counters(ctx, &ctr); rays = ctr.rays_closest + ctr.rays_any; }
    fprintf(stderr, "{\"event\":\"render_gpu\",\"rays\":%llu,\"seconds\":%.6f,\"nodes\":%zu,\"tris\":%zu,\"materials\":%zu,\"lights\":%zu,\"width\":%d,\"height\":%d}\n",
            rays, t1 - t0, f.nodes.size(), f.tris.size(), f.materials.size(), lights.size(), rp.width, rp.height);
    if (!out.empty()) writeVec(out, rgb);
    if (ctx) p_miro_gpu_destroy(ctx);
    return t1 - t0;
}

}  // namespace

int main(int argc, char** argv) {
    std::string scene, dumpPrim, traceIn, traceOut, floatOut, ppmOut, meshDir, qbvhOut, texDir, gpuOut, gpuLib, gpuPpm, shadowRaysOut, shadowHitsOut, instOut;
    int threads = 1, repeat = 1, warmup = 0; bool stock = false, doFloat = false, doShadow = false;
    float shadowLight[3] = {0, 0, 0}; size_t shadowFirst = 0, shadowCount = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        #define NEXT() (i + 1 < argc ? std::string(argv[++i]) : (die("missing value for " + a), std::string()))
        if (a == "--scene") scene = NEXT();
        else if (a == "--assets") g_assetRoot = NEXT();
        else if (a == "--threads") threads = atoi(NEXT().c_str());
        else if (a == "--repeat") repeat = atoi(NEXT().c_str());
        else if (a == "--warmup") warmup = atoi(NEXT().c_str());
        else if (a == "--dump-primary") dumpPrim = NEXT();
        else if (a == "--trace") { traceIn = NEXT(); }
        else if (a == "--hits") traceOut = NEXT();
        else if (a == "--shadow-light") { doShadow = true; for (int k = 0; k < 3; k++) shadowLight[k] = (float)atof(NEXT().c_str());
                                          shadowFirst = (size_t)atoll(NEXT().c_str()); shadowCount = (size_t)atoll(NEXT().c_str()); }
        else if (a == "--shadow-rays") shadowRaysOut = NEXT();
        else if (a == "--shadow-hits") shadowHitsOut = NEXT();
        else if (a == "--render-float") { doFloat = true; floatOut = NEXT(); }
        else if (a == "--render-stock") { stock = true; ppmOut = NEXT(); }
        else if (a == "--dump-meshes") meshDir = NEXT();
        else if (a == "--dump-qbvh") qbvhOut = NEXT();
        else if (a == "--dump-textures") texDir = NEXT();
        else if (a == "--dump-instances") instOut = NEXT();
        else if (a == "--render-gpu") gpuOut = NEXT();
        else if (a == "--gpu-lib") gpuLib = NEXT();
        else if (a == "--gpu-ppm") gpuPpm = NEXT();
        else die("unknown argument " + a);
    }
    if (scene.empty()) die("usage: miro_ref --scene S.miro [--assets DIR] [--threads N] [--dump-primary F] [--trace RAYS --hits F] [--render-float F] [--render-stock F.ppm] [--dump-meshes DIR] [--dump-qbvh F]");
    if (threads < 1) threads = 1; if (threads > 16) threads = 16;   // Ray::counter[128*tid] caps the program at 16 threads (Ray.h:30,74)
    omp_set_num_threads(threads);
    FILE* quiet = freopen("/dev/null", "w", stdout);   // the reference prints per-bucket progress
    (void)quiet;
    double tb0 = omp_get_wtime();
    loadScene(scene);
    double tb1 = omp_get_wtime();
    fprintf(stderr, "{\"event\":\"scene\",\"objects\":%zu,\"qbvh_nodes\":%u,\"qbvh_leaves\":%u,\"build_s\":%.4f}\n",
            g_scene->objects()->size(), QBVH_Node::nodeCount, QBVH_Node::leafCount, tb1 - tb0);
    if (!meshDir.empty()) dumpMeshes(meshDir);
    if (!qbvhOut.empty()) dumpQBVH(qbvhOut);
    if (!texDir.empty()) dumpTextures(texDir);
    if (!instOut.empty()) dumpInstances(instOut);
    if (!dumpPrim.empty()) dumpPrimary(dumpPrim);
    if (!gpuOut.empty()) {
        if (gpuLib.empty()) die("--render-gpu needs --gpu-lib path/to/libmiro_gpu.so");
        renderGPU(gpuOut, gpuLib);
        if (!gpuPpm.empty()) g_image->writePPM((char*)gpuPpm.c_str());
    }
    if (!traceIn.empty()) {
        resetTraceCalls();
        std::vector<RayRec> rays = readVec<RayRec>(traceIn);
        std::vector<RefHit> hits;
        double s = traceBatch(rays, hits, threads, repeat, warmup, !traceOut.empty() || doShadow);
        double mean = g_traceMean;
        if (!traceOut.empty()) writeVec(traceOut, hits);
        unsigned long long n = rays.size();
        fprintf(stderr, "{\"event\":\"trace\",\"rays\":%llu,\"seconds\":%.6f,\"mean_seconds\":%.6f,\"mrays_per_s\":%.4f,\"threads\":%d,\"repeat\":%d,\"warmup\":%d}\n",
                n, s, mean, n / s * 1e-6, threads, repeat, warmup);
        if (doShadow) {
            std::vector<RayRec> srays; std::vector<RefHit> shits;
            shadowRaysFromHits(rays, hits, shadowFirst, shadowCount, shadowLight, srays);
            double ss = traceBatch(srays, shits, threads, repeat, warmup, !shadowHitsOut.empty());
            if (!shadowRaysOut.empty()) writeVec(shadowRaysOut, srays);
            if (!shadowHitsOut.empty()) writeVec(shadowHitsOut, shits);
            fprintf(stderr, "{\"event\":\"shadow\",\"rays\":%zu,\"seconds\":%.6f,\"mean_seconds\":%.6f,\"mrays_per_s\":%.4f,\"threads\":%d}\n",
                    srays.size(), ss, g_traceMean, srays.size() / ss * 1e-6, threads);
        }
    }
    if (doFloat) {
        resetTraceCalls();
        double s = renderFloat(floatOut, threads);
        unsigned long long n = traceCalls();
        fprintf(stderr, "{\"event\":\"render_float\",\"rays\":%llu,\"seconds\":%.6f,\"mrays_per_s\":%.4f,\"threads\":%d,\"width\":%d,\"height\":%d}\n",
                n, s, n / s * 1e-6, threads, g_image->width(), g_image->height());
    }
    if (stock) {
        double best = 1e30; unsigned long long n = 0;
        for (int it = 0; it < repeat; ++it) {
            resetTraceCalls();
            double t0 = omp_get_wtime();
            g_scene->raytraceImage(g_camera, g_image);       // the stock render loop (Scene.cpp:86-217)
            double t1 = omp_get_wtime(); if (t1 - t0 < best) best = t1 - t0;
            n = traceCalls();
        }
        if (!ppmOut.empty() && ppmOut != "-") g_image->writePPM((char*)ppmOut.c_str());
        fprintf(stderr, "{\"event\":\"render_stock\",\"rays\":%llu,\"seconds\":%.6f,\"mrays_per_s\":%.4f,\"threads\":%d,\"width\":%d,\"height\":%d}\n",
                n, best, n / best * 1e-6, threads, g_image->width(), g_image->height());
    }
    return 0;
}
