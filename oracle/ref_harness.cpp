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

// --trace: Scene::trace (Scene.cpp:295) over a caller-supplied ray buffer.
double g_traceMean = 0.0;       // mean seconds over the timed repeats of the last traceRays call
double traceRays(const std::string& in, const std::string& out, int threads, int repeat, int warmup) {
    std::vector<RayRec> rays = readVec<RayRec>(in);
    std::vector<RefHit> hits(rays.size());
    double best = 1e30, sum = 0.0;
    for (int it = -warmup; it < repeat; ++it) {
        double t0 = omp_get_wtime();
        #pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
        for (long i = 0; i < (long)rays.size(); i++) {
            unsigned tid = omp_get_thread_num();
            const RayRec& q = rays[i];
            Ray r(tid, Vector3(q.ox, q.oy, q.oz), Vector3(q.dx, q.dy, q.dz), q.time);
            HitInfo h; h.t = q.tmax;
            bool hit = g_scene->trace(tid, h, r, q.tmin);
            hits[i] = toRefHit(hit, h);
        }
        double t1 = omp_get_wtime();
        if (it >= 0) { if (t1 - t0 < best) best = t1 - t0; sum += t1 - t0; }
    }
    g_traceMean = sum / (repeat > 0 ? repeat : 1);
    if (!out.empty()) writeVec(out, hits);
    return best;
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

}  // namespace

int main(int argc, char** argv) {
    std::string scene, dumpPrim, traceIn, traceOut, floatOut, ppmOut, meshDir, qbvhOut, texDir;
    int threads = 1, repeat = 1, warmup = 0; bool stock = false, doFloat = false;
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
        else if (a == "--render-float") { doFloat = true; floatOut = NEXT(); }
        else if (a == "--render-stock") { stock = true; ppmOut = NEXT(); }
        else if (a == "--dump-meshes") meshDir = NEXT();
        else if (a == "--dump-qbvh") qbvhOut = NEXT();
        else if (a == "--dump-textures") texDir = NEXT();
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
    if (!dumpPrim.empty()) dumpPrimary(dumpPrim);
    if (!traceIn.empty()) {
        resetTraceCalls();
        double s = traceRays(traceIn, traceOut, threads, repeat, warmup);
        unsigned long long n = traceCalls() / (unsigned long long)(repeat + warmup);
        fprintf(stderr, "{\"event\":\"trace\",\"rays\":%llu,\"seconds\":%.6f,\"mean_seconds\":%.6f,\"mrays_per_s\":%.4f,\"threads\":%d,\"repeat\":%d,\"warmup\":%d}\n",
                n, s, g_traceMean, n / s * 1e-6, threads, repeat, warmup);
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
