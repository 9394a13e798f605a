// miro_script.cpp — ".miro" scene scripts: a line-based rendition of the reference's make*Scene()
// functions (src/assignment2.h, src/Assignment3.h, src/main.cpp), one API call per line.  The same
// script is read by oracle/ref_harness.cpp, which drives the reference's own classes with it — so "a
// scene that renders on the CPU renders unchanged on the GPU".
//
//   image W H
//   camera [eye x y z] [lookat x y z] [viewdir x y z] [up x y z] [fov deg] [focus f] [aperture a] [shutter s]
//   scene  [bgcolor r g b] [pathtrace 0|1] [numpaths n] [maxbounces n] [minsubdivs n] [maxsubdivs n]
//          [noise f] [sampleenv 0|1] [envmap TEX exposure] [seed n] [devicebuild 0|1]
//   texture NAME file.{hdr,tga,ppm}
//   material NAME lambert [kd r g b] [ka r g b] [colormap TEX]
//   material NAME blinn   [kd ..] [ka ..] [ks ..] [specexp f] [specamt f] [ior f | ior_i i f] [reflect f] [refract f]
//                         [gloss f] [translucency f] [emit intensity r g b] [colormap TEX] [alphamap TEX] [sampleenv 0|1]
//                         [normalmap TEX] [specularmap TEX] [reflectmap TEX] [refractmap TEX]
//   light point [pos x y z] [power p] [shadows 0|1] [fastshadows 0|1]
//   light rect  [v1 x y z] [v2 x y z] [v3 x y z] [power p] [samples n] [noise t] [shadows 0|1] [fastshadows 0|1]
//   light dome  [tex TEX] [power gain] [samples n] [noise t] [fastshadows 0|1]        (fastshadows 0 = Light::setFastShadows(false))
//   mesh NAME file.obj [ctm m11 m12 ... m44]          (row-major)
//   object MESH MATERIAL                               makeMeshObjs
//   mbobject MESH_T1 MESH_T2 MATERIAL                  makeMBMeshObjs
//   blas NAME MESH MATERIAL [MESH MATERIAL ...]        ProxyObject::setupProxy / setupMultiProxy
//   instance BLAS m11 m12 ... m44                      new ProxyObject(objs, bvh, M)
#include "miro_host.h"
#include <fstream>
#include <sstream>
#include <map>

namespace miro {

static Vector3 read3(std::istringstream& ss) { float x = 0, y = 0, z = 0; ss >> x >> y >> z; return Vector3(x, y, z); }

bool loadSceneScript(const char* file, const char* assetRoot, LoadedScene& out, std::string& error,
                     const std::vector<std::pair<std::string, TriangleMesh*>>* preloaded,
                     const std::vector<std::pair<std::string, RawImage*>>* preloadedImages) {
    std::ifstream in(file);
    if (!in) { error = std::string("cannot open scene script ") + file; return false; }
    const std::string root = assetRoot ? assetRoot : ".";
    auto path = [&](const std::string& p) { return (!p.empty() && p[0] == '/') ? p : root + "/" + p; };
    out.scene.reset(new Scene); out.camera.reset(new Camera); out.image.reset(new Image);
    out.image->resize(512, 512);
    std::map<std::string, TriangleMesh*> meshes;
    std::map<std::string, Material*> materials;
    std::map<std::string, Texture*> textures;
    std::map<std::string, ProxyBLAS*> blases;
    int lineNo = 0;
    std::string line;
    auto fail = [&](const std::string& m) { error = std::string(file) + ":" + std::to_string(lineNo) + ": " + m; return false; };
    while (std::getline(in, line)) {
        ++lineNo;
        size_t h = line.find('#'); if (h != std::string::npos) line = line.substr(0, h);
        std::istringstream ss(line);
        std::string cmd, k;
        if (!(ss >> cmd)) continue;
        if (cmd == "image") { int w = 0, hh = 0; ss >> w >> hh; if (w <= 0 || hh <= 0) return fail("bad image size"); out.image->resize(w, hh); }
        else if (cmd == "camera") {
            while (ss >> k) {
                if (k == "eye") out.camera->setEye(read3(ss));
                else if (k == "lookat") out.camera->setLookAt(read3(ss));
                else if (k == "viewdir") out.camera->setViewDir(read3(ss));
                else if (k == "up") out.camera->setUp(read3(ss));
                else if (k == "fov") { float f; ss >> f; out.camera->setFOV(f); }
                else if (k == "focus") { float f; ss >> f; out.camera->setFocusPlane(f); }
                else if (k == "aperture") { float f; ss >> f; out.camera->setAperture(f); }
                else if (k == "shutter") { float f; ss >> f; out.camera->setShutterSpeed(f); }
                else return fail("camera: unknown key " + k);
            }
        } else if (cmd == "scene") {
            while (ss >> k) {
                if (k == "bgcolor") out.scene->setBGColor(read3(ss));
                else if (k == "pathtrace") { int v; ss >> v; out.scene->setPathTrace(v != 0); }
                else if (k == "numpaths") { int v; ss >> v; out.scene->setNumPaths(v); }
                else if (k == "maxbounces") { int v; ss >> v; out.scene->setMaxBounces(v); }
                else if (k == "minsubdivs") { int v; ss >> v; out.scene->setMinSubdivs(v); }
                else if (k == "maxsubdivs") { int v; ss >> v; out.scene->setMaxSubdivs(v); }
                else if (k == "noise") { float v; ss >> v; out.scene->setNoise(v); }
                else if (k == "sampleenv") { int v; ss >> v; out.scene->setSampleEnv(v != 0); }
                else if (k == "seed") { unsigned long long v; ss >> v; out.scene->setSeed(v); }
                else if (k == "devicebuild") { int v; ss >> v; out.scene->setBuildOnDevice(v != 0); }
                else if (k == "envmap") {
                    std::string t; float e; ss >> t >> e;
                    if (!textures.count(t)) return fail("unknown texture " + t);
                    out.scene->setEnvMap(textures[t]); out.scene->setEnvExposure(e);
                } else return fail("scene: unknown key " + k);
            }
        } else if (cmd == "texture") {
            std::string name, p; ss >> name >> p;
            RawImage* img = nullptr;
            if (preloadedImages) for (auto& pr : *preloadedImages) if (pr.first == name) img = pr.second;
            if (!img) {
                out.images.emplace_back(new RawImage);
                img = out.images.back().get();
                if (!img->loadImage(path(p).c_str())) return fail("cannot load texture " + path(p));
            }
            out.textures.emplace_back(new Texture(img));
            textures[name] = out.textures.back().get();
        } else if (cmd == "material") {
            std::string name, kind; ss >> name >> kind;
            if (kind == "lambert") {
                Lambert* m = new Lambert(Vector3(1.f), Vector3(0.f));
                out.materials.emplace_back(m);
                while (ss >> k) {
                    if (k == "kd") m->setKd(read3(ss));
                    else if (k == "ka") m->setKa(read3(ss));
                    else if (k == "colormap") { std::string t; ss >> t; if (!textures.count(t)) return fail("unknown texture " + t); m->setColorMap(textures[t]); }
                    else return fail("lambert: unknown key " + k);
                }
                materials[name] = m;
            } else if (kind == "blinn") {
                Blinn* m = new Blinn(Vector3(1.f));
                out.materials.emplace_back(m);
                while (ss >> k) {
                    if (k == "kd") m->setKd(read3(ss));
                    else if (k == "ka") m->setKa(read3(ss));
                    else if (k == "ks") m->setKs(read3(ss));
                    else if (k == "specexp") { float f; ss >> f; m->setSpecExp(f); }
                    else if (k == "specamt") { float f; ss >> f; m->setSpecAmt(f); }
                    else if (k == "ior") { float f; ss >> f; m->setIor(f, 0); m->setIor(f, 1); m->setIor(f, 2); }
                    else if (k == "ior_i") { int i; float f; ss >> i >> f; if (i < 0 || i > 2) return fail("ior_i: index must be 0..2"); m->setIor(f, i); }   // Blinn::setIor(ior, i)
                    else if (k == "disperse") { int v; ss >> v; m->m_disperse = v != 0; }
                    else if (k == "reflect") { float f; ss >> f; m->setReflectAmt(f); }
                    else if (k == "refract") { float f; ss >> f; m->setRefractAmt(f); }
                    else if (k == "gloss") { float f; ss >> f; m->setReflectGloss(f); }
                    else if (k == "translucency") { float f; ss >> f; m->setTranslucency(f); }
                    else if (k == "emit") { float i; ss >> i; Vector3 c = read3(ss); m->setLightEmittedIntensity(i); m->setLightEmittedColor(c); }
                    else if (k == "colormap") { std::string t; ss >> t; if (!textures.count(t)) return fail("unknown texture " + t); m->setColorMap(textures[t]); }
                    else if (k == "alphamap") { std::string t; ss >> t; if (!textures.count(t)) return fail("unknown texture " + t); m->setAlphaMap(textures[t]); }
                    else if (k == "normalmap" || k == "specularmap" || k == "reflectmap" || k == "refractmap") {
                        std::string t; ss >> t; if (!textures.count(t)) return fail("unknown texture " + t);
                        if (k == "normalmap") m->setNormalMap(textures[t]); else if (k == "specularmap") m->setSpecularMap(textures[t]);
                        else if (k == "reflectmap") m->setReflectMap(textures[t]); else m->setRefractMap(textures[t]);
                    }
                    else if (k == "sampleenv") { int v; ss >> v; m->setSampleEnv(v != 0); }
                    else return fail("blinn: unknown key " + k);
                }
                materials[name] = m;
            } else return fail("unknown material kind " + kind);
        } else if (cmd == "light") {
            std::string kind; ss >> kind;
            if (kind == "point") {
                PointLight* l = new PointLight; out.lights.emplace_back(l); l->setColor(Vector3(1, 1, 1));
                while (ss >> k) {
                    if (k == "pos") l->setPosition(read3(ss));
                    else if (k == "power") { float f; ss >> f; l->setPower(f); }
                    else if (k == "shadows") { int v; ss >> v; l->setCastShadows(v != 0); }
                    else if (k == "fastshadows") { int v; ss >> v; l->setFastShadows(v != 0); }
                    else return fail("point light: unknown key " + k);
                }
                out.scene->addLight(l);
            } else if (kind == "rect") {
                RectangleLight* l = new RectangleLight; out.lights.emplace_back(l); l->setColor(Vector3(1, 1, 1));
                Vector3 v1, v2, v3; float power = 0.f;
                while (ss >> k) {
                    if (k == "v1") v1 = read3(ss);
                    else if (k == "v2") v2 = read3(ss);
                    else if (k == "v3") v3 = read3(ss);
                    else if (k == "power") ss >> power;
                    else if (k == "samples") { int n; ss >> n; l->setSamples(n); }
                    else if (k == "noise") { float f; ss >> f; l->setNoiseThreshold(f); }
                    else if (k == "shadows") { int v; ss >> v; l->setCastShadows(v != 0); }
                    else if (k == "fastshadows") { int v; ss >> v; l->setFastShadows(v != 0); }
                    else return fail("rect light: unknown key " + k);
                }
                l->setPower(power); l->setVertices(v1, v2, v3);   // call order of src/assignment2.h:404-405
                out.scene->addLight(l);
            } else if (kind == "dome") {
                DomeLight* l = new DomeLight; out.lights.emplace_back(l);
                while (ss >> k) {
                    if (k == "tex") { std::string t; ss >> t; if (!textures.count(t)) return fail("unknown texture " + t); l->setTexture(textures[t]); }
                    else if (k == "power") { float f; ss >> f; l->setPower(f); }
                    else if (k == "samples") { int n; ss >> n; l->setSamples(n); }
                    else if (k == "noise") { float f; ss >> f; l->setNoiseThreshold(f); }
                    else if (k == "fastshadows") { int v; ss >> v; l->setFastShadows(v != 0); }
                    else return fail("dome light: unknown key " + k);
                }
                if (!l->m_lightMap) return fail("dome light without tex");
                out.scene->addLight(l);
            } else return fail("unknown light kind " + kind);
        } else if (cmd == "mesh") {
            std::string name, p; ss >> name >> p;
            Matrix4x4 ctm;
            if (ss >> k) {
                if (k != "ctm") return fail("mesh: expected ctm");
                for (int i = 0; i < 16; ++i) ss >> ctm.m[i];
            }
            TriangleMesh* mesh = nullptr;
            if (preloaded) for (auto& pr : *preloaded) if (pr.first == name) mesh = pr.second;
            if (!mesh) {
                out.meshes.emplace_back(new TriangleMesh);
                mesh = out.meshes.back().get();
                if (!mesh->load(path(p).c_str(), ctm)) return fail("cannot load mesh " + path(p));
            }
            mesh->ordinal = (int)out.meshNames.size();
            out.meshNames.push_back(name);
            meshes[name] = mesh;
        } else if (cmd == "object") {
            std::string m, mat; ss >> m >> mat;
            if (!meshes.count(m)) return fail("unknown mesh " + m);
            if (!materials.count(mat)) return fail("unknown material " + mat);
            makeMeshObjs(out.scene.get(), meshes[m], materials[mat]);
        } else if (cmd == "mbobject") {
            std::string m1, m2, mat; ss >> m1 >> m2 >> mat;
            if (!meshes.count(m1) || !meshes.count(m2)) return fail("unknown mesh");
            if (!materials.count(mat)) return fail("unknown material " + mat);
            makeMBMeshObjs(out.scene.get(), meshes[m1], meshes[m2], materials[mat]);
        } else if (cmd == "blas") {
            std::string name, m, mat; ss >> name;
            std::vector<TriangleMesh*> ms; std::vector<Material*> mats;
            while (ss >> m >> mat) {
                if (!meshes.count(m)) return fail("unknown mesh " + m);
                if (!materials.count(mat)) return fail("unknown material " + mat);
                ms.push_back(meshes[m]); mats.push_back(materials[mat]);
            }
            if (ms.empty()) return fail("blas without meshes");
            out.blases.emplace_back(ProxyBLAS::setupMultiProxy(ms.data(), (int)ms.size(), mats.data()));
            blases[name] = out.blases.back().get();
        } else if (cmd == "instance") {
            std::string name; ss >> name;
            if (!blases.count(name)) return fail("unknown blas " + name);
            Matrix4x4 M; for (int i = 0; i < 16; ++i) ss >> M.m[i];
            addProxyObject(out.scene.get(), blases[name], M);
        } else return fail("unknown command " + cmd);
    }
    return true;
}

}  // namespace miro
