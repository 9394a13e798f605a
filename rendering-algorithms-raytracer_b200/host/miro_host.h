// miro_host.h — the product's C++ host layer: a mirror of the reference's scene-description API
// (same class / method names, argument meaning and defaults) whose Scene::preCalc() flattens the
// scene into a miro_gpu_scene_desc and whose Scene::raytraceImage()/trace() run on the GPU through
// the C ABI of include/miro_gpu.h.  There is no CPU renderer here.
//
//   reference                         here
//   Camera (src/Camera.h)             miro::Camera       same setters, same defaults (src/Camera.cpp:15-27)
//   Scene (src/Scene.h)               miro::Scene        addObject/addLight/setEnvMap/preCalc/raytraceImage/trace
//   Object/MBObject/ProxyObject       miro::Object (POD: mesh, index, material, type, + mesh_t2 / blas+matrix)
//   makeMeshObjs / makeMBMeshObjs     miro::makeMeshObjs / makeMBMeshObjs   (src/main.cpp:22-23)
//   ProxyObject::setupProxy / setupMultiProxy (src/ProxyObject.cpp:131-166)  miro::ProxyBLAS
//   TriangleMesh::load (src/TriangleMeshLoad.cpp)   miro::TriangleMesh::load / setGeometry
//   RawImage / Texture                miro::RawImage / miro::Texture
//   Lambert / Blinn                   miro::Lambert / miro::Blinn
//   PointLight / RectangleLight / DomeLight   same names
//   Image (src/Image.h)               miro::Image        setPixel = clamp + 2.2 gamma LUT, writePPM bottom-up
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include <memory>
#include "miro_math.h"
#include "miro_bvh.h"
#include "../../include/miro_gpu.h"

namespace miro {

// ---- geometry --------------------------------------------------------------------------------
class TriangleMesh {
public:
    struct TupleI3 { uint32_t x, y, z; };
    struct VectorR2 { float x, y; };
    // OBJ subset of the reference loader: v / vt / vn / f with triangles only and v, v/t, v//n, v/t/n
    // corner forms; faces without normals get one flat normal each (src/TriangleMeshLoad.cpp:100-214).
    bool load(const char* file, const Matrix4x4& ctm = Matrix4x4());
    // in-memory alternative (tests, synthetic scenes).  nidx/tidx may be NULL (flat normals / no uvs).
    void setGeometry(const float* vertices, uint32_t nv, const uint32_t* vidx, uint32_t nf,
                     const float* normals = nullptr, uint32_t nn = 0, const uint32_t* nidx = nullptr,
                     const float* uvs = nullptr, uint32_t nt = 0, const uint32_t* tidx = nullptr);
    std::vector<Vector3> m_vertices, m_normals;
    std::vector<VectorR2> m_texCoords;
    std::vector<TupleI3> m_vertexIndices, m_normalIndices, m_texCoordIndices;   // m_texCoordIndices empty: no uvs
    uint32_t m_numTris = 0;
    int ordinal = -1;          // set by Scene when first referenced (reported back in hits)
    // Per-normal-index tangents / bitangents for normal mapping (TriangleMesh::preCalc, src/TriangleMesh.cpp:107-150): every
    // triangle with a non-degenerate uv mapping writes the Gram-Schmidt tangent of its three corners, later triangles
    // overwrite earlier ones.  Entries no triangle writes are ZERO here (the reference leaves them uninitialised).  Empty
    // when the mesh has no texture coordinates.
    void computeTangents(std::vector<Vector3>& tangents, std::vector<Vector3>& bitangents) const;
private:
    void makeFlatNormals();
};

// ---- images / textures -----------------------------------------------------------------------
enum ImageType { RGB, RGBA, GRAYSCALE, HDR };
class RawImage {
public:
    RawImage() {}
    RawImage(int w, int h, const float* data, ImageType t);
    bool loadImage(const char* filename);    // .hdr (RGBE), .ppm (P6), .tga (type 2/3, gamma->linear, BGR swap)
    std::vector<float> m_rawData;
    int m_width = 0, m_height = 0;
    ImageType m_imageType = RGB;
    int channels() const { return m_imageType == GRAYSCALE ? 1 : (m_imageType == RGBA ? 4 : 3); }
};
class Texture {
public:
    explicit Texture(RawImage* image) : m_image(image) {}
    float getWidth() const { return (float)m_image->m_width; }
    float getHeight() const { return (float)m_image->m_height; }
    RawImage* m_image;
    int ordinal = -1;
};

class Image {   // frame buffer: 8-bit gamma-mapped pixels (as the reference) + the float radiance the GPU returned
public:
    Image();
    ~Image();
    Image(const Image&) = delete;
    Image& operator=(const Image&) = delete;
    void resize(int width, int height);
    // page-lock the frame buffers (once per size) so that the frame comes down at PCIe speed: a 1080p float frame is 25 MB, and
    // into ordinary memory the driver stages it at ~8 GB/s — longer than the frame takes to render
    void pin();
    void setPixel(int x, int y, const Vector3& p);     // src/Image.cpp:71-87 (Map: clamp, 2.2 gamma LUT)
    void writePPM(const char* file) const;             // bottom-up flip, src/Image.cpp:132-154
    int width() const { return m_width; }
    int height() const { return m_height; }
    const unsigned char* getCharPixels() const { return m_pixels.data(); }
    unsigned char* charPixels() { return m_pixels.data(); }      // filled by the GPU (Image::setPixel's mapping runs on the device)
    std::vector<float> m_radiance;                     // width*height*3, row 0 = bottom
    static unsigned char Map(float r);
    static const float* linearToGammaF();              // 32769-entry table used by the adaptive cut-off
private:
    void unpin();
    std::vector<unsigned char> m_pixels;
    int m_width = 1, m_height = 1;
    bool m_pinned = false;
};

// ---- materials -------------------------------------------------------------------------------
class Material {
public:
    virtual ~Material() {}
    virtual uint32_t kind() const = 0;
    void setColorMap(Texture* t) { m_colorMap = t; }
    void setAlphaMap(Texture* t) { m_alphaMap = t; }
    void setNormalMap(Texture* t) { m_normalMap = t; }        // src/Material.h:22-25; used by Blinn::shade only (src/Blinn.cpp:120-142)
    void setSpecularMap(Texture* t) { m_specularMap = t; }
    void setReflectMap(Texture* t) { m_reflectMap = t; }
    void setRefractMap(Texture* t) { m_refractMap = t; }
    void setSampleEnv(bool b) { m_sampleEnv = b; }
    void setTranslucency(float t) { m_translucency = t; }
    void setRefractAmt(float r) { m_refractAmt = r; }
    bool m_disperse = false;
    Texture* m_colorMap = nullptr;
    Texture* m_alphaMap = nullptr;
    Texture* m_normalMap = nullptr;
    Texture* m_specularMap = nullptr;
    Texture* m_reflectMap = nullptr;
    Texture* m_refractMap = nullptr;
    bool m_sampleEnv = true;
    float m_translucency = 0.f;
    float m_refractAmt = 0.f;
    int ordinal = -1;
    virtual void fill(miro_gpu_material& m) const = 0;
};
class Lambert : public Material {
public:
    Lambert(const Vector3& kd = Vector3(1.f), const Vector3& ka = Vector3(0.f)) : m_kd(kd), m_ka(ka) {}
    uint32_t kind() const override { return MIRO_GPU_MAT_LAMBERT; }
    void setKd(const Vector3& v) { m_kd = v; }
    void setKa(const Vector3& v) { m_ka = v; }
    void fill(miro_gpu_material& m) const override;
    Vector3 m_kd, m_ka;
};
class Blinn : public Material {
public:
    // same defaults as src/Blinn.h:11-22 / src/Blinn.cpp:15-35 (the ctor zeroes emission)
    Blinn(const Vector3& kd = Vector3(1.f), const Vector3& ka = Vector3(0.f), const Vector3& ks = Vector3(1.f),
          const Vector3& kt = Vector3(0.f), float ior = 1.5f, float specExp = 1.0f, float specAmt = 0.0f,
          float reflectAmt = 0.0f, float refractAmt = 0.0f, float specGloss = 1.0f);
    uint32_t kind() const override { return MIRO_GPU_MAT_BLINN; }
    void setKd(const Vector3& v) { m_kd = v; }
    void setKa(const Vector3& v) { m_ka = v; }
    void setKs(const Vector3& v) { m_ks = v; }
    void setIor(float ior, int i = 0) { m_ior[i] = ior; }
    void setSpecExp(float v) { m_specExp = v; }
    void setSpecAmt(float v) { m_specAmt = v; }
    void setReflectAmt(float v) { m_reflectAmt = v; }
    void setReflectGloss(float v) { m_specGloss = v; }
    void setLightEmittedIntensity(float le) { m_lightEmitted = le; }
    void setLightEmittedColor(const Vector3& le) { m_Le = le; }
    void fill(miro_gpu_material& m) const override;
    Vector3 m_kd, m_ka, m_ks, m_kt;
    float m_ior[3];
    float m_specExp, m_specAmt, m_reflectAmt, m_lightEmitted = 0.f, m_specGloss;
    Vector3 m_Le;
};

// ---- lights ----------------------------------------------------------------------------------
enum lightType_t { RECTANGLE_LIGHT, POINT_LIGHT, DOME_LIGHT };
class Light {
public:
    virtual ~Light() {}
    void setColor(const Vector3& v) { m_color = v; }
    virtual void setPower(float f) { m_power = f; }
    void setSamples(int n) { m_numSamples = n; }
    void setCastShadows(bool c) { m_castShadows = c; }
    void setFastShadows(bool c) { m_fastShadows = c; }   // src/Light.h:24
    void setNoiseThreshold(float t) { m_noiseThreshold = t; }
    virtual void fill(miro_gpu_light& l) const = 0;
    Vector3 m_color;
    float m_power = 0.f;
    int m_numSamples = 1;
    bool m_castShadows = true;
    bool m_fastShadows = true;                       // src/Light.h:16
    float m_noiseThreshold = MIRO_GPU_EPSILON;       // src/Light.h:17
};
class PointLight : public Light {
public:
    void setPosition(const Vector3& v) { m_position = v; }
    void fill(miro_gpu_light& l) const override;
    Vector3 m_position;
};
class RectangleLight : public Light {
public:
    void setVertices(const Vector3& v1, const Vector3& v2, const Vector3& v3) { m_v1 = v1; m_v2 = v2; m_v3 = v3; setPower(m_power); }
    void setPower(float f) override;                 // area-normalised, src/RectangleLight.cpp:14-40
    void fill(miro_gpu_light& l) const override;
    Vector3 m_v1, m_v2, m_v3;
};
class DomeLight : public Light {
public:
    void setTexture(Texture* t) { m_lightMap = t; }
    void setPower(float f) override { m_Gain = f; }  // src/DomeLight.h:52
    void fill(miro_gpu_light& l) const override;
    Texture* m_lightMap = nullptr;
    float m_Gain = 1.f;
};

// ---- objects ---------------------------------------------------------------------------------
enum objectType_t { OBJECT, MB_OBJECT, PROXY_OBJECT };
class ProxyBLAS;
struct Object {          // one per triangle (OBJECT / MB_OBJECT) or one per instance (PROXY_OBJECT)
    const Material* m_material = nullptr;
    TriangleMesh* m_mesh = nullptr;
    uint32_t m_index = 0;
    objectType_t m_objectType = OBJECT;
    TriangleMesh* m_mesh_t2 = nullptr;        // MB_OBJECT: second pose
    ProxyBLAS* m_blas = nullptr;              // PROXY_OBJECT
    Matrix4x4 m_transform;                    // PROXY_OBJECT
};
typedef std::vector<Object> Objects;

class ProxyBLAS {        // the shared geometry + BVH of a ProxyObject family
public:
    static ProxyBLAS* setupProxy(TriangleMesh* mesh, Material* mat);
    static ProxyBLAS* setupMultiProxy(TriangleMesh* mesh[], int numObjs, Material* mat[]);
    Objects m_objects;
    int32_t root_ref = MIRO_GPU_CHILD_EMPTY;  // filled by Scene::preCalc
    bool flattened = false;
};

class Camera {
public:
    Camera();
    void setEye(const Vector3& e) { m_eye = e; }
    void setUp(const Vector3& u) { m_up = u.normalized(); }
    void setViewDir(const Vector3& v) { m_viewDir = v.normalized(); }
    void setLookAt(const Vector3& l) { m_lookAt = l; setViewDir(l - m_eye); }
    void setFOV(float fovDeg) { m_fov = fovDeg; }
    void setFocusPlane(float f) { m_focusPlane = f; }
    void setAperture(float f) { m_aperture = f; }
    void setShutterSpeed(float f) { m_shutterSpeed = f; }
    void fill(miro_gpu_camera& c) const;
    Vector3 m_eye, m_up, m_viewDir, m_lookAt;
    float m_fov, m_focusPlane, m_aperture, m_shutterSpeed;
};

// Flattened scene: owns the arrays a miro_gpu_scene_desc points into.
struct FlatScene {
    std::vector<miro_gpu_node> nodes;
    std::vector<miro_gpu_tri> tris;
    std::vector<miro_gpu_mbtri> mbtris;
    std::vector<miro_gpu_instance> instances;
    std::vector<miro_gpu_prim> prims;
    std::vector<float> normals, tangents, bitangents, uvs, inst_nxf;      // tangents / bitangents: same indexing as normals
    std::vector<miro_gpu_material> materials;
    std::vector<miro_gpu_light> lights;
    std::vector<miro_gpu_texture> textures;
    int32_t root = MIRO_GPU_CHILD_EMPTY;
    int32_t env_map = -1;
    float env_exposure = 1.f;
    float bg[3] = {0, 0, 0};
    BvhStats top_stats;
    miro_gpu_scene_desc desc() const;
};

class Scene {
public:
    Scene();
    ~Scene();
    void addObject(const Object& o) { m_objects.push_back(o); }
    const Objects* objects() const { return &m_objects; }
    void addLight(Light* l) { m_lights.push_back(l); }
    void setEnvMap(Texture* t) { m_envMap = t; }
    void setEnvExposure(float e) { m_envExposure = e; }
    void setBGColor(const Vector3& c) { m_BGColor = c; }
    void setPathTrace(bool pt) { m_pathTrace = pt; }
    void setMinSubdivs(int r) { m_minSubdivs = r; }
    void setMaxSubdivs(int r) { m_maxSubdivs = r; }
    void setMaxBounces(int mb) { m_maxBounces = mb; }
    void setNumPaths(int p) { m_numPaths = p; }
    void setNoise(float n) { m_noiseThreshold = n; }
    void setSampleEnv(bool b) { m_sampleLightFromEnv = b; }
    void setSeed(uint64_t s) { m_seed = s; }
    // Build the acceleration structure on the GPU at attach() instead of on the host in preCalc() (static triangles only;
    // falls back to the host build when the scene has motion-blur objects or instances).
    void setBuildOnDevice(bool b) { m_buildOnDevice = b; }

    // Scene::preCalc (src/Scene.cpp:63-79): build the BVHs and flatten.  Pure host work; no GPU needed.
    // Returns false (see lastError()) when the scene uses something outside the supported scope.
    bool preCalc();
    const FlatScene& flat() const { return m_flat; }

    // GPU side.  attach() creates the device context (one per process/GPU) and uploads the flattened scene.
    bool attach(int device_id = 0);
    // The same over several GPUs of one box (miro_gpu_group_*): the scene is replicated, raytraceImage() deals the frame's buckets
    // (or, with setSampleSharding(true), the paths of every camera sample) to the devices, trace() / traceAny() split their batch.
    bool attachDevices(const int* device_ids, int n);
    void setSampleSharding(bool on) { m_sampleSharding = on; }
    miro_gpu_group* group() const { return m_group; }
    miro_gpu_ctx* context() const { return m_group ? miro_gpu_group_ctx(m_group, 0) : m_ctx; }
    // Scene::raytraceImage (src/Scene.cpp:86-217): float radiance into img->m_radiance and 8-bit pixels via Image::setPixel.
    bool raytraceImage(const Camera* cam, Image* img, int shard_index = 0, int shard_count = 1);
    // Scene::trace (src/Scene.cpp:295-298), batched.
    bool trace(const miro_gpu_ray* rays, size_t n, miro_gpu_hit* hits);
    bool traceAny(const miro_gpu_ray* rays, size_t n, uint32_t* occluded_bits);
    void renderParams(const Image* img, miro_gpu_render_params& p) const;
    const std::string& lastError() const { return m_error; }

    bool m_pathTrace = false;
    int m_numPaths = 1, m_minSubdivs = 1, m_maxSubdivs = 1, m_maxBounces = 10;

    // bookkeeping used by the C API / tests
    std::vector<TriangleMesh*> meshes;     // by ordinal
protected:
    int meshOrdinal(TriangleMesh* m);
    int materialOrdinal(const Material* m);
    int textureOrdinal(Texture* t);
    bool flattenBLAS(ProxyBLAS* b);
    bool appendTriangle(const Object& o, uint32_t& outIndex, float lo[3], float hi[3]);
    Objects m_objects;
    std::vector<Light*> m_lights;
    Vector3 m_BGColor;
    Texture* m_envMap = nullptr;
    float m_envExposure = 1.f;
    float m_noiseThreshold = 0.01f;
    bool m_sampleLightFromEnv = false;
    bool m_buildOnDevice = false;
    uint64_t m_seed = 3163513;             // the reference seeds its MT19937 with this (src/Scene.cpp:24)
    FlatScene m_flat;
    std::vector<const Material*> m_materialList;
    std::vector<Texture*> m_textureList;
    // staging used while flattening: source triangles before leaf ordering
    std::vector<miro_gpu_tri> m_srcTris; std::vector<miro_gpu_prim> m_srcPrims;
    std::vector<miro_gpu_mbtri> m_srcMB; std::vector<miro_gpu_prim> m_srcMBPrims;
    std::vector<miro_gpu_instance> m_srcInst; std::vector<float> m_srcInstNxf;
    std::vector<uint32_t> m_meshNormalBase, m_meshUvBase;
    miro_gpu_ctx* m_ctx = nullptr;
    miro_gpu_group* m_group = nullptr;
    bool m_sampleSharding = false;
    std::string m_error;
};

void makeMeshObjs(Scene* scene, TriangleMesh* mesh, Material* mat);                         // one Object per triangle, reverse order
void makeMBMeshObjs(Scene* scene, TriangleMesh* mesh, TriangleMesh* mesh2, Material* mat);
void addProxyObject(Scene* scene, ProxyBLAS* blas, const Matrix4x4& m);                      // new ProxyObject(objs, bvh, m) + addObject

// ".miro" scene script (shared with oracle/ref_harness.cpp): builds scene/camera/image.
struct LoadedScene {
    std::unique_ptr<Scene> scene; std::unique_ptr<Camera> camera; std::unique_ptr<Image> image;
    std::vector<std::unique_ptr<TriangleMesh>> meshes; std::vector<std::string> meshNames;
    std::vector<std::unique_ptr<Material>> materials; std::vector<std::unique_ptr<Light>> lights;
    std::vector<std::unique_ptr<RawImage>> images; std::vector<std::unique_ptr<Texture>> textures;
    std::vector<std::unique_ptr<ProxyBLAS>> blases;
};
// `preloaded` lets callers supply in-memory meshes by name (used instead of reading the OBJ path).
bool loadSceneScript(const char* file, const char* assetRoot, LoadedScene& out, std::string& error,
                     const std::vector<std::pair<std::string, TriangleMesh*>>* preloaded = nullptr,
                     const std::vector<std::pair<std::string, RawImage*>>* preloadedImages = nullptr);

}  // namespace miro
