// rt.hpp — C++ host-side mirror of the reference's Go `rt` package API for the path-tracing hot path.
//
// The reference's toolchain (Go) is absent from the build image, so the host side above the C-ABI
// (include/rtx_b200.h) is written in C++ with the same names, argument meaning and error behaviour as
// the Go package: scene types (rt/sphere.go, quad.go, triangle.go, plane.go, hittable_list.go,
// primitives.go, transform.go, volume.go), materials/textures (rt/material.go, rt/texture.go), the camera
// builder and presets (rt/camera.go:106-280), NewBVHNodeFromList (rt/bvh.go:64), LoadOBJ
// (rt/obj_loader.go:15), the HDR loader (rt/image_loader.go:152-383) and BucketRenderer /
// ProgressiveRenderer (rt/bucket_renderer.go:54, rt/renderer.go:27).
//
// These types only DESCRIBE a scene. They have no Hit()/Scatter(): all arithmetic of the hot path runs in
// the CUDA library, and an object the flattener does not know is a flatten-time error (no CPU fallback).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtx_b200.h"

namespace rt {

// ---- rt/vec3.go ---------------------------------------------------------------------------------------
struct Vec3 {
    double X = 0, Y = 0, Z = 0;
    Vec3() = default;
    Vec3(double x, double y, double z) : X(x), Y(y), Z(z) {}
    Vec3 Add(const Vec3& u) const { return {X + u.X, Y + u.Y, Z + u.Z}; }
    Vec3 Sub(const Vec3& u) const { return {X - u.X, Y - u.Y, Z - u.Z}; }
    Vec3 Scale(double t) const { return {t * X, t * Y, t * Z}; }
    Vec3 Div(double t) const { return Scale(1 / t); }
    Vec3 Neg() const { return {-X, -Y, -Z}; }
    double Len2() const { return X * X + Y * Y + Z * Z; }
    double Len() const { return std::sqrt(Len2()); }
    Vec3 Unit() const {
        double l = Len();
        if (l == 0) return *this;
        return Div(l);
    }
};
using Point3 = Vec3;
using Color = Vec3;
inline double Dot(const Vec3& a, const Vec3& b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }
inline Vec3 Cross(const Vec3& a, const Vec3& b) {
    return {a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X};
}
constexpr double Pi = 3.1415926535897932385;  // rt/utils.go:11
inline double DegreesToRadians(double degrees) { return degrees * Pi / 180.0; }

// ---- rt/interval.go, rt/aabb.go (bounding boxes: used only to order the reference BVH) ------------------
struct Interval {
    double Min = std::numeric_limits<double>::infinity(), Max = -std::numeric_limits<double>::infinity();
    double Size() const { return Max - Min; }
    Interval Expand(double d) const { return {Min - d, Max + d}; }
    Interval Add(double d) const { return {Min + d, Max + d}; }
};
Interval NewIntervalFromIntervals(const Interval& a, const Interval& b);
struct AABB {
    Interval X, Y, Z;
    void padToMinimums();
    AABB Translate(const Vec3& o) const;
    int LongestAxis() const;
    Vec3 Centroid() const;
};
AABB NewAABBFromIntervals(Interval x, Interval y, Interval z);
AABB NewAABBFromPoints(const Point3& a, const Point3& b);
AABB NewAABBFromBoxes(const AABB& a, const AABB& b);

// ---- rt/texture.go ------------------------------------------------------------------------------------
struct Texture {
    virtual ~Texture() = default;
};
using TexturePtr = std::shared_ptr<Texture>;
struct SolidColor : Texture {
    Color Albedo;
    explicit SolidColor(Color a) : Albedo(a) {}
};
struct CheckerTexture : Texture {
    double invScale;
    TexturePtr even, odd;
    CheckerTexture(double scale, TexturePtr e, TexturePtr o) : invScale(1.0 / scale), even(e), odd(o) {}
};
struct Perlin {  // rt/noise.go:8-28. The reference fills the tables from Go's auto-seeded global source; here from a seeded stream in the same draw order
    Vec3 randvec[256];
    int permX[256], permY[256], permZ[256];
};
struct NoiseTexture : Texture {  // rt/texture.go:19-29
    std::shared_ptr<Perlin> noise;
    double scale;
};
struct ImageLoader {  // rt/image_loader.go:14-120. The reference decodes any format Go's image package knows; without an image
    // decoder in this toolchain, Load reads binary PPM (P6, 8 bit) — tools/make_assets.py converts the reference's earthmap.jpg
    int imageWidth = 0, imageHeight = 0;
    std::vector<double> data;  // 3 per pixel: LinearToGamma(v / 255) as ImageLoader.Load stores it (:62-70)
    bool Load(const std::string& filename);
    int Width() const { return data.empty() ? 0 : imageWidth; }
    int Height() const { return data.empty() ? 0 : imageHeight; }
};
struct ImageTexture : Texture {  // rt/image_texture.go
    std::shared_ptr<ImageLoader> image;
};
TexturePtr NewImageTexture(const std::string& filename);                          // resolved with FindAsset(filename, "images")
TexturePtr NewImageTextureFromImage(std::shared_ptr<ImageLoader> image);
TexturePtr NewNoiseTexture(double scale, uint64_t seed = 0x5EEDull);
TexturePtr NewSolidColor(Color albedo);
TexturePtr NewCheckerTexture(double scale, TexturePtr even, TexturePtr odd);
TexturePtr NewCheckerTextureFromColors(double scale, Color c1, Color c2);

// ---- rt/material.go -----------------------------------------------------------------------------------
struct Material {
    virtual ~Material() = default;
};
using MaterialPtr = std::shared_ptr<Material>;
struct Lambertian : Material {
    TexturePtr tex;
};
struct Metal : Material {
    Color Albedo;
    double Fuzz;
};
struct Dielectric : Material {
    double RefractionIndex;
};
struct DiffuseLight : Material {
    TexturePtr tex;
};
struct Isotropic : Material {
    TexturePtr tex;
};
MaterialPtr NewLambertian(Color albedo);
MaterialPtr NewLambertianTexture(TexturePtr tex);
MaterialPtr NewMetal(Color albedo, double fuzz);
MaterialPtr NewDielectric(double refractionIndex);
MaterialPtr NewDiffuseLight(TexturePtr tex);
MaterialPtr NewDiffuseLightColor(Color emit);
MaterialPtr NewIsotropic(TexturePtr tex);
MaterialPtr NewIsotropicFromColor(Color albedo);

// ---- hittables ----------------------------------------------------------------------------------------
struct Hittable {
    virtual ~Hittable() = default;
    virtual AABB BoundingBox() const = 0;
};
using HittablePtr = std::shared_ptr<Hittable>;

struct Sphere : Hittable {  // rt/sphere.go
    Point3 Center0;
    Vec3 Velocity;
    double Radius;     // max(0, radius)
    double rawRadius;  // as passed to the constructor (the bbox uses it, rt/sphere.go:15-21)
    MaterialPtr Mat;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
};
struct Quad : Hittable {  // rt/quad.go
    Point3 Q;
    Vec3 u, v;
    MaterialPtr mat;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
};
struct Triangle : Hittable {  // rt/triangle.go
    Point3 v0, v1, v2;
    MaterialPtr mat;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
};
struct Plane : Hittable {  // rt/plane.go
    Point3 Point;
    Vec3 Normal;
    MaterialPtr Mat;
    AABB BoundingBox() const override;
};
struct Circle : Hittable {  // rt/circle.go
    Point3 center;
    Vec3 normal;  // unit
    double radius;
    MaterialPtr mat;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
};
struct HittableList : Hittable {  // rt/hittable_list.go
    std::vector<HittablePtr> Objects;
    AABB bbox;
    void Add(HittablePtr o);
    AABB BoundingBox() const override { return bbox; }
};
using HittableListPtr = std::shared_ptr<HittableList>;

struct BVHLeaf : Hittable {  // rt/bvh.go:21
    std::vector<HittablePtr> objects;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
};
struct BVHNode : Hittable {  // rt/bvh.go:13
    HittablePtr left, right;
    AABB bbox;
    // Unexported addition of the drop-in: the source slice in insertion order (root only), so that the
    // flattener can report stable identifiers; the tree itself only fixes the test order.
    std::vector<HittablePtr> src;
    // A mesh root returned by LoadOBJ may be DEFERRED: its box and source slice are set, the pointer tree is not built. These types never
    // traverse (no Hit here): the tree's only product on this side is the test order of its leaves, which the CUDA library derives on the
    // device when the flattener leaves tri_rank NULL (csrc/rtx_rank_gpu.cuh: 3 ms against a 100 ms tree build for 280 K triangles).
    // EnsureBuilt() builds the reference-order tree on demand (SetEagerMeshBVH(true) / RT_EAGER_BVH=1 make LoadOBJ build it at once).
    bool deferred = false;
    // ... and then LoadOBJ does not even create the 280 K Triangle objects: it keeps the faces as flat arrays (what the flattener wants
    // anyway); EnsureBuilt() materialises src from them before it builds the tree.
    struct Soup { std::vector<double> v0, v1, v2; std::shared_ptr<Material> mat; size_t n = 0; };
    std::shared_ptr<Soup> soup;
    void EnsureBuilt();
    AABB BoundingBox() const override { return bbox; }
};
using BVHNodePtr = std::shared_ptr<BVHNode>;
void SetEagerMeshBVH(bool eager);
bool EagerMeshBVH();

struct Translate : Hittable {  // rt/transform.go:78
    HittablePtr Obj;
    Vec3 Offset;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
};
struct RotateY : Hittable {  // rt/transform.go:113
    HittablePtr Obj;
    double SinTheta, CosTheta;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
};
struct Scale : Hittable {  // rt/transform.go:360
    HittablePtr Obj;
    Vec3 Factor, InvFactor;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
};
struct Volume : Hittable {  // rt/volume.go:10
    HittablePtr boundary;
    double negInvDensity;
    MaterialPtr phaseFunction;
    AABB BoundingBox() const override { return boundary->BoundingBox(); }
};

std::shared_ptr<Sphere> NewSphere(Point3 center, double radius, MaterialPtr mat);
std::shared_ptr<Sphere> NewMovingSphere(Point3 c1, Point3 c2, double radius, MaterialPtr mat);
std::shared_ptr<Quad> NewQuad(Point3 Q, Vec3 u, Vec3 v, MaterialPtr mat);
std::shared_ptr<Triangle> NewTriangle(Point3 v0, Point3 v1, Point3 v2, MaterialPtr mat);
std::shared_ptr<Plane> NewPlane(Point3 point, Vec3 normal, MaterialPtr mat);
HittableListPtr NewHittableList();
HittablePtr Box(Point3 a, Point3 b, MaterialPtr mat);  // rt/primitives.go:5
std::shared_ptr<Circle> NewCircle(Point3 center, Vec3 normal, double radius, MaterialPtr mat);  // rt/circle.go:14
HittablePtr Pyramid(Point3 baseCenter, double baseSize, double height, MaterialPtr mat);         // rt/primitives.go:39
std::shared_ptr<Translate> NewTranslate(HittablePtr obj, Vec3 offset);
std::shared_ptr<RotateY> Ry(HittablePtr obj, double angleDegrees);
std::shared_ptr<Scale> NewScale(HittablePtr obj, Vec3 factor);
std::shared_ptr<Scale> NewUniformScale(HittablePtr obj, double factor);
std::shared_ptr<Volume> NewVolume(HittablePtr boundary, double density, TexturePtr tex);
std::shared_ptr<Volume> NewVolumeFromColor(HittablePtr boundary, double density, Color albedo);
BVHNodePtr NewBVHNodeFromList(const HittableListPtr& list);                                  // rt/bvh.go:64
BVHNodePtr NewBVHNode(const std::vector<HittablePtr>& objects, size_t start, size_t end);     // rt/bvh.go:69
BVHNodePtr NewBVHNodeDeferred(std::vector<HittablePtr>&& objects);   // a mesh root without its tree (see BVHNode::deferred)
BVHNodePtr NewMeshRootFromSoup(std::shared_ptr<BVHNode::Soup> soup, int threads);   // the same from flat face arrays (LoadOBJ)

struct Transform {  // rt/transform.go:9-71
    Vec3 scale{1, 1, 1}, rotation{0, 0, 0}, position{0, 0, 0};
    Transform& SetScale(Vec3 s) { scale = s; return *this; }
    Transform& SetUniformScale(double s) { scale = {s, s, s}; return *this; }
    Transform& SetRotation(Vec3 r) { rotation = r; return *this; }
    Transform& SetRotationY(double a) { rotation.Y = a; return *this; }
    Transform& SetPosition(Vec3 p) { position = p; return *this; }
    HittablePtr Apply(HittablePtr obj) const;
};
inline Transform NewTransform() { return Transform{}; }

HittablePtr LoadOBJ(const std::string& filename, MaterialPtr material);  // rt/obj_loader.go:15 (throws on error); rt_obj.cpp
// the text parse of LoadOBJ on `threads` threads (0 = all, 1 = the reference's sequential scan): vertices and 3 vertex indices per triangle, file order
void ParseOBJ(const std::string& filename, int threads, std::vector<Point3>& vertices, std::vector<uint32_t>& tri_indices);
HittablePtr LoadOBJWithTransform(const std::string& filename, MaterialPtr material, const Transform* transform);

// ---- rt/image_loader.go (HDR part) + rt/hdri.go (state only; the distribution is built by the library) --
struct HDRIEnvironment {
    int width = 0, height = 0;
    std::vector<double> rgb;  // 3*w*h, empty = load failed (IsValid() == false)
    double rotation = 0;      // radians
    bool useImportanceSampling = true;
    bool IsValid() const { return !rgb.empty(); }
    void SetRotation(double degrees) { rotation = degrees * M_PI / 180.0; }
    void DisableImportanceSampling() { useImportanceSampling = false; }
};
std::shared_ptr<HDRIEnvironment> NewHDRIEnvironment(const std::string& filename);
bool LoadHDR(const std::string& path, int& w, int& h, std::vector<double>& rgb, std::string* err);
std::string FindAsset(const std::string& filename, const std::string& assetType);  // "" when not found

// ---- rt/camera.go --------------------------------------------------------------------------------------
struct CameraPreset {
    double AspectRatio;
    int ImageWidth, SamplesPerPixel, MaxDepth;
    double Vfov, DefocusAngle, FocusDist;
    Point3 LookFrom, LookAt;
    Vec3 Vup;
    bool FreeCamera = false;
    Vec3 Forward;
    Color Background;
    bool UseSkyGradient = false;
};
CameraPreset QuickPreview();
CameraPreset StandardQuality();
CameraPreset HighQuality();

struct Camera {
    double AspectRatio = 1.0;
    int ImageWidth = 800, ImageHeight = 0, SamplesPerPixel = 10, MaxDepth = 50;
    double Vfov = 90;
    Point3 LookFrom{0, 0, 0}, LookAt{0, 0, -1};
    Vec3 Vup{0, 1, 0};
    double DefocusAngle = 0, FocusDist = 1.0;
    Point3 LookFrom2, LookAt2;
    bool CameraMotion = false, FreeCamera = false;
    Vec3 Forward{0, 0, -1};
    Color Background{0, 0, 0};
    bool UseSkyGradient = false, PhantomHDRI = false;
    std::vector<HittablePtr> Lights;
    std::shared_ptr<HDRIEnvironment> Environment;

    void ApplyPreset(const CameraPreset& p);
    Camera& SetResolution(int width, double aspect) { ImageWidth = width; AspectRatio = aspect; return *this; }
    Camera& SetQuality(int samples, int maxDepth) { SamplesPerPixel = samples; MaxDepth = maxDepth; return *this; }
    Camera& SetPosition(Point3 from, Point3 at, Vec3 vup) { LookFrom = from; LookAt = at; Vup = vup; return *this; }
    Camera& SetLens(double vfov, double defocus, double focus) { Vfov = vfov; DefocusAngle = defocus; FocusDist = focus; return *this; }
    Camera& SetMotion(Point3 from2, Point3 at2) { LookFrom2 = from2; LookAt2 = at2; CameraMotion = true; return *this; }
    Camera& SetVFOV(double v) { Vfov = v; return *this; }
    Camera& SetDefocus(double angle, double focus) { DefocusAngle = angle; FocusDist = focus; return *this; }
    Camera& DisableMotion() { CameraMotion = false; return *this; }
    Camera& EnableFreeCamera(Point3 pos, Vec3 forward, Vec3 vup) { LookFrom = pos; Forward = forward.Unit(); Vup = vup.Unit(); FreeCamera = true; return *this; }
    Camera& SetBackground(Color c) { Background = c; return *this; }
    Camera& EnableSkyGradient(bool e) { UseSkyGradient = e; return *this; }
    Camera& SetEnvironmentMap(const std::string& filename) { Environment = NewHDRIEnvironment(filename); return *this; }
    Camera& SetEnvironmentRotation(double deg) { if (Environment) Environment->SetRotation(deg); return *this; }
    Camera& DisableEnvironmentImportanceSampling() { if (Environment) Environment->DisableImportanceSampling(); return *this; }
    Camera& SetPhantomHDRI(bool p) { PhantomHDRI = p; return *this; }
    Camera& AddLight(HittablePtr l) { Lights.push_back(l); return *this; }
    std::shared_ptr<Camera> Build();  // Initialize(): only ImageHeight is derived on the host (rt/camera.go:299)
    void Initialize();
    void FillDesc(rtx_camera_desc& d) const;
};
using CameraPtr = std::shared_ptr<Camera>;
inline Camera NewCameraBuilder() { return Camera{}; }

// ---- scenes (rt/scenes.go) -----------------------------------------------------------------------------
// RandomScene is seeded (the reference draws from Go's auto-seeded global source, rt/utils.go:18).
struct Scene {
    HittableListPtr world;
    CameraPtr camera;
};
Scene RandomScene(uint64_t seed = 0x5EEDull);                 // rt/scenes.go:30
Scene HDRITestScene(const std::string& hdrPath);              // rt/scenes.go:406
Scene CornellBoxScene();                                      // rt/scenes.go:463
Scene CornellBoxGlossy();                                     // rt/scenes.go:606
Scene CornellBoxLucy(const std::string& objPath);             // rt/scenes.go:714
Scene EarthScene(const std::string& imagePath);                 // rt/scenes.go:210
Scene PerlinSpheresScene(uint64_t seed = 0x5EEDull);           // rt/scenes.go:242
Scene PrimitivesScene();                                       // rt/scenes.go:313
Scene CheckeredSpheresScene();                                 // rt/scenes.go:132
Scene SimpleScene();                                           // rt/scenes.go:172
Scene QuadsScene();                                            // rt/scenes.go:274
Scene GlossyMetalTest();                                       // rt/scenes.go:564
Scene CornellSmoke();                                          // rt/scenes.go:820
Scene LoadSceneByName(const std::string& name, const std::string& assetRoot, uint64_t seed);  // main.go:108

// ---- flattening: pointer graph -> rtx_scene_desc ----------------------------------------------------------
struct FlatScene {
    std::vector<int32_t> tex_type, tex_even, tex_odd;
    std::vector<double> tex_color, tex_inv_scale;
    std::vector<int32_t> mat_type, mat_tex;
    std::vector<double> mat_albedo, mat_fuzz, mat_ior;
    std::vector<double> sph_center, sph_velocity, sph_radius;
    std::vector<int32_t> sph_mat;
    std::vector<double> quad_q, quad_u, quad_v;
    std::vector<int32_t> quad_mat;
    std::vector<double> tri_v0, tri_v1, tri_v2;
    std::vector<int32_t> tri_mat, tri_rank;
    bool have_tri_rank = true;   // false: some mesh tree was deferred, the library derives the canonical ranks on the device
    std::vector<double> plane_point, plane_normal;
    std::vector<int32_t> plane_mat;
    std::vector<double> circle_center, circle_normal, circle_radius;
    std::vector<int32_t> circle_mat;
    std::vector<double> perlin_vec;
    std::vector<int32_t> perlin_perm;
    std::vector<int32_t> image_width, image_height;
    std::vector<int64_t> image_offset;
    std::vector<double> image_rgb;
    std::vector<int32_t> group_kind, group_begin, group_count, list_item_kind, list_item_index;
    std::vector<int32_t> xf_type;
    std::vector<double> xf_a, xf_b;
    std::vector<double> vol_neg_inv_density;
    std::vector<int32_t> vol_mat;
    std::vector<int32_t> entry_geom_kind, entry_geom_index, entry_xf_begin, entry_xf_count, entry_volume, entry_rank;
    std::vector<int32_t> light_quad;
    std::shared_ptr<HDRIEnvironment> env;
    bool world_is_bvh = false;
    rtx_scene_desc Desc() const;  // borrowed pointers into this object
};
// world: *HittableList (linear traversal, Camera.Render) or *BVHNode from NewBVHNodeFromList.
// Throws std::runtime_error for user-defined / unsupported Hittable, Material or Texture types.
std::shared_ptr<FlatScene> Flatten(const HittablePtr& world, const Camera& camera);

// ---- renderers (rt/bucket_renderer.go, rt/renderer.go) -----------------------------------------------------
class BucketRenderer {
public:
    BucketRenderer(CameraPtr camera, HittablePtr world, int bucketSize, int numWorkers, int deviceId = 0);
    ~BucketRenderer();
    // One tick of the pass state machine (rt/bucket_renderer.go:127-164). NEVER blocks on a pass: like the reference, which
    // renders in goroutines, the first call starts pass 0 on a worker thread, later calls poll `passComplete`, publish the
    // finished pass into the framebuffer under the mutex and start the next one. A display loop calls it once per frame.
    // Returns 0, or -1 once a pass failed (LastError()).
    int Update();
    void RenderToCompletion();    // headless helper: Update() until IsCompleted(), sleeping between polls
    bool IsCompleted() const { return completed_; }
    int CurrentPass() const { return currentPass_; }
    double GetRenderDurationSeconds() const { return duration_s_; }
    int SaveImage(const std::string& filename) const;  // binary PPM (P6) or PNG by extension
    std::vector<uint8_t> CopyFramebuffer() const;      // framebuffer.Pix (RGBA8, stride 4*W) under the mutex: safe while a pass runs (Draw, :303)
    const std::vector<uint8_t>& Pix() const { return pix_; }  // unlocked view: only when no pass is running (after IsCompleted())
    int Width() const { return w_; }
    int Height() const { return h_; }
    rtx_ctx* Context() { return ctx_; }
    const std::string& LastError() const { return err_; }
    uint64_t seed = 0x9E3779B97F4A7C15ull;
    // What the reference's Update() does when the last pass ends (rt/bucket_renderer.go:151-155): burn the stats bar into the bottom 30
    // rows and save "image.png". Off by default in this mirror so that tests and benchmarks see the rendered pixels only;
    // a display front-end sets burnStats = true, autoSave = "image.png" to behave exactly like the reference.
    bool burnStats = false;
    std::string autoSave;
    void DrawStatsToFramebuffer();   // rt/bucket_renderer.go:375-407 (black bar, white 7x13 text "WxH | SPP:n | Depth:d | 100.0% | time | Workers: n")

private:
    void renderPass(int pass);   // worker thread: one rtx_render_pass + resolve into back_, then passComplete = true
    void join();
    CameraPtr camera_;
    std::shared_ptr<FlatScene> flat_;
    rtx_ctx* ctx_ = nullptr;
    int w_ = 0, h_ = 0, currentPass_ = 0, totalPasses_ = 3, numWorkers_ = 0;
    bool completed_ = false, renderStarted_ = false;
    double duration_s_ = 0;
    std::chrono::steady_clock::time_point renderStart_;
    std::vector<uint8_t> pix_, back_;   // framebuffer shown / pass being resolved
    mutable std::mutex mu_;             // protects pix_ (rt/bucket_renderer.go:51)
    std::atomic<bool> passComplete_{false};
    std::thread worker_;
    std::string err_;
};
std::shared_ptr<BucketRenderer> NewBucketRenderer(CameraPtr camera, HittablePtr world, int bucketSize, int numWorkers);
// ProgressiveRenderer (rt/renderer.go:27) keeps its API and routes to the same device pass.
std::shared_ptr<BucketRenderer> NewProgressiveRenderer(CameraPtr camera, HittablePtr world);

int WritePNG(const std::string& path, const uint8_t* rgba, int w, int h);
std::string FormatDuration(double seconds);                                                                    // rt/utils.go:50-61
std::string StatsBarText(int width, int height, int spp, int depth, double seconds, int numWorkers);          // rt/bucket_renderer.go:391-398
void DrawStatsBar(uint8_t* rgba, int w, int h, const std::string& text);                                       // rt/bucket_renderer.go:378-406

}  // namespace rt
