// rt.cpp — host-side mirror of the Go `rt` package: constructors, reference-order BVH, loaders, camera.
// See rt.hpp. Reference citations are relative to /root/reference/.
#include <thread>

#include "rt.hpp"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <future>
#include <sstream>
#include <sys/stat.h>

namespace rt {

static const double kInf = std::numeric_limits<double>::infinity();

// ---- rt/interval.go:28-40, rt/aabb.go ------------------------------------------------------------------
Interval NewIntervalFromIntervals(const Interval& a, const Interval& b) {
    double mn = a.Min;
    if (b.Min < a.Min) mn = b.Min;
    double mx = a.Max;
    if (b.Max > a.Max) mx = b.Max;
    return {mn, mx};
}
void AABB::padToMinimums() {  // rt/aabb.go:117-128
    const double delta = 0.0001;
    if (X.Size() < delta) X = X.Expand(delta);
    if (Y.Size() < delta) Y = Y.Expand(delta);
    if (Z.Size() < delta) Z = Z.Expand(delta);
}
AABB NewAABBFromIntervals(Interval x, Interval y, Interval z) {
    AABB b{x, y, z};
    b.padToMinimums();
    return b;
}
AABB NewAABBFromPoints(const Point3& a, const Point3& b) {
    AABB box{{std::fmin(a.X, b.X), std::fmax(a.X, b.X)}, {std::fmin(a.Y, b.Y), std::fmax(a.Y, b.Y)}, {std::fmin(a.Z, b.Z), std::fmax(a.Z, b.Z)}};
    // Go's math.Min/Max propagate NaN, fmin/fmax do not; only the NaN centroid of an infinite Plane reaches here.
    if (std::isnan(a.X) || std::isnan(b.X)) box.X = {NAN, NAN};
    if (std::isnan(a.Y) || std::isnan(b.Y)) box.Y = {NAN, NAN};
    if (std::isnan(a.Z) || std::isnan(b.Z)) box.Z = {NAN, NAN};
    box.padToMinimums();
    return box;
}
AABB NewAABBFromBoxes(const AABB& a, const AABB& b) {
    return {NewIntervalFromIntervals(a.X, b.X), NewIntervalFromIntervals(a.Y, b.Y), NewIntervalFromIntervals(a.Z, b.Z)};
}
AABB AABB::Translate(const Vec3& o) const { return NewAABBFromIntervals(X.Add(o.X), Y.Add(o.Y), Z.Add(o.Z)); }
int AABB::LongestAxis() const {  // rt/aabb.go:139-150
    double xs = X.Size(), ys = Y.Size(), zs = Z.Size();
    if (xs > ys && xs > zs) return 0;
    if (ys > zs) return 1;
    return 2;
}
Vec3 AABB::Centroid() const { return {(X.Min + X.Max) * 0.5, (Y.Min + Y.Max) * 0.5, (Z.Min + Z.Max) * 0.5}; }

// ---- textures / materials --------------------------------------------------------------------------------
TexturePtr NewSolidColor(Color albedo) { return std::make_shared<SolidColor>(albedo); }
bool ImageLoader::Load(const std::string& filename) {  // rt/image_loader.go:44-73 (decode -> LinearToGamma(v / 65535 * 257) = sqrt(v8 / 255))
    FILE* f = std::fopen(filename.c_str(), "rb");
    if (!f) return false;
    char magic[3] = {0, 0, 0};
    int w = 0, h = 0, maxv = 0;
    auto token = [&](int& out) {   // next integer, skipping whitespace and # comments
        int c = std::fgetc(f);
        while (c == ' ' || c == '\n' || c == '\r' || c == '\t' || c == '#') {
            if (c == '#') while (c != '\n' && c != EOF) c = std::fgetc(f);
            else c = std::fgetc(f);
        }
        if (c < '0' || c > '9') return false;
        out = 0;
        while (c >= '0' && c <= '9') { out = out * 10 + (c - '0'); c = std::fgetc(f); }
        return true;   // the single whitespace after the token is consumed
    };
    bool ok = std::fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && magic[1] == '6' && token(w) && token(h) && token(maxv) && w > 0 && h > 0 && maxv == 255;
    std::vector<unsigned char> raw;
    if (ok) {
        raw.resize((size_t)3 * w * h);
        ok = std::fread(raw.data(), 1, raw.size(), f) == raw.size();
    }
    std::fclose(f);
    if (!ok) return false;
    imageWidth = w; imageHeight = h;
    data.resize(raw.size());
    for (size_t i = 0; i < raw.size(); i++) {
        double lin = (double)raw[i] * 257.0 / 65535.0;   // Go's RGBA() of an 8-bit channel is v * 257; / 65535
        data[i] = lin > 0 ? std::sqrt(lin) : 0.0;       // LinearToGamma (rt/utils.go:85-90)
    }
    return true;
}
TexturePtr NewImageTextureFromImage(std::shared_ptr<ImageLoader> image) {
    auto t = std::make_shared<ImageTexture>();
    t->image = image;
    return t;
}
TexturePtr NewImageTexture(const std::string& filename) {  // rt/image_texture.go:11-15; a missing file leaves an empty image (debug colours in the reference)
    auto img = std::make_shared<ImageLoader>();
    std::string path = FindAsset(filename, "images");
    if (!path.empty()) img->Load(path);
    return NewImageTextureFromImage(img);
}
TexturePtr NewNoiseTexture(double scale, uint64_t seed) {  // rt/texture.go:24-29 + NewPerlin rt/noise.go:15-28 (same draw order, seeded SplitMix64)
    uint64_t st = seed;
    auto next = [&st]() {
        uint64_t z = (st += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    auto uniform = [&]() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); };
    auto p = std::make_shared<Perlin>();
    for (int i = 0; i < 256; i++) {  // RandomVec3Range(-1, 1).Unit()
        Vec3 v{-1 + 2 * uniform(), -1 + 2 * uniform(), -1 + 2 * uniform()};
        p->randvec[i] = v.Unit();
    }
    int* perms[3] = {p->permX, p->permY, p->permZ};
    for (int a = 0; a < 3; a++) {  // perlinGeneratePerm + permute :67-78
        for (int i = 0; i < 256; i++) perms[a][i] = i;
        for (int i = 255; i > 0; i--) {
            int target = (int)(uniform() * (double)(i + 1));   // RandomInt(0, i)
            if (target > i) target = i;
            std::swap(perms[a][i], perms[a][target]);
        }
    }
    auto t = std::make_shared<NoiseTexture>();
    t->noise = p; t->scale = scale;
    return t;
}
TexturePtr NewCheckerTexture(double scale, TexturePtr even, TexturePtr odd) { return std::make_shared<CheckerTexture>(scale, even, odd); }
TexturePtr NewCheckerTextureFromColors(double scale, Color c1, Color c2) { return NewCheckerTexture(scale, NewSolidColor(c1), NewSolidColor(c2)); }

MaterialPtr NewLambertianTexture(TexturePtr tex) { auto m = std::make_shared<Lambertian>(); m->tex = tex; return m; }
MaterialPtr NewLambertian(Color albedo) { return NewLambertianTexture(NewSolidColor(albedo)); }
MaterialPtr NewMetal(Color albedo, double fuzz) {
    auto m = std::make_shared<Metal>();
    m->Albedo = albedo;
    m->Fuzz = fuzz > 1 ? 1 : fuzz;  // rt/material.go:92
    return m;
}
MaterialPtr NewDielectric(double ri) { auto m = std::make_shared<Dielectric>(); m->RefractionIndex = ri; return m; }
MaterialPtr NewDiffuseLight(TexturePtr tex) { auto m = std::make_shared<DiffuseLight>(); m->tex = tex; return m; }
MaterialPtr NewDiffuseLightColor(Color emit) { return NewDiffuseLight(NewSolidColor(emit)); }
MaterialPtr NewIsotropic(TexturePtr tex) { auto m = std::make_shared<Isotropic>(); m->tex = tex; return m; }
MaterialPtr NewIsotropicFromColor(Color albedo) { return NewIsotropic(NewSolidColor(albedo)); }

// ---- primitives ---------------------------------------------------------------------------------------
std::shared_ptr<Sphere> NewSphere(Point3 center, double radius, MaterialPtr mat) {  // rt/sphere.go:14
    auto s = std::make_shared<Sphere>();
    Vec3 rvec{radius, radius, radius};
    s->Center0 = center;
    s->Velocity = {0, 0, 0};
    s->Radius = std::fmax(0.0, radius);
    s->rawRadius = radius;
    s->Mat = mat;
    s->bbox = NewAABBFromPoints(center.Sub(rvec), center.Add(rvec));
    return s;
}
std::shared_ptr<Sphere> NewMovingSphere(Point3 c1, Point3 c2, double radius, MaterialPtr mat) {  // rt/sphere.go:24
    auto s = std::make_shared<Sphere>();
    Vec3 rvec{radius, radius, radius};
    s->Center0 = c1;
    s->Velocity = c2.Sub(c1);
    s->Radius = std::fmax(0.0, radius);
    s->rawRadius = radius;
    s->Mat = mat;
    s->bbox = NewAABBFromBoxes(NewAABBFromPoints(c1.Sub(rvec), c1.Add(rvec)), NewAABBFromPoints(c2.Sub(rvec), c2.Add(rvec)));
    return s;
}
std::shared_ptr<Quad> NewQuad(Point3 Q, Vec3 u, Vec3 v, MaterialPtr mat) {  // rt/quad.go:16
    auto q = std::make_shared<Quad>();
    q->Q = Q; q->u = u; q->v = v; q->mat = mat;
    AABB d1 = NewAABBFromPoints(Q, Q.Add(u).Add(v));
    AABB d2 = NewAABBFromPoints(Q.Add(u), Q.Add(v));
    q->bbox = NewAABBFromBoxes(d1, d2);
    return q;
}
std::shared_ptr<Triangle> NewTriangle(Point3 v0, Point3 v1, Point3 v2, MaterialPtr mat) {  // rt/triangle.go:17
    auto t = std::make_shared<Triangle>();
    t->v0 = v0; t->v1 = v1; t->v2 = v2; t->mat = mat;
    Point3 mn{std::fmin(v0.X, std::fmin(v1.X, v2.X)), std::fmin(v0.Y, std::fmin(v1.Y, v2.Y)), std::fmin(v0.Z, std::fmin(v1.Z, v2.Z))};
    Point3 mx{std::fmax(v0.X, std::fmax(v1.X, v2.X)), std::fmax(v0.Y, std::fmax(v1.Y, v2.Y)), std::fmax(v0.Z, std::fmax(v1.Z, v2.Z))};
    t->bbox = NewAABBFromPoints(mn, mx);
    return t;
}
std::shared_ptr<Plane> NewPlane(Point3 point, Vec3 normal, MaterialPtr mat) {  // rt/plane.go:12
    auto p = std::make_shared<Plane>();
    p->Point = point; p->Normal = normal.Unit(); p->Mat = mat;
    return p;
}
AABB Plane::BoundingBox() const { return AABB{{-kInf, kInf}, {-kInf, kInf}, {-kInf, kInf}}; }

HittableListPtr NewHittableList() { return std::make_shared<HittableList>(); }
void HittableList::Add(HittablePtr o) {
    Objects.push_back(o);
    bbox = NewAABBFromBoxes(bbox, o->BoundingBox());
}
std::shared_ptr<Circle> NewCircle(Point3 center, Vec3 normal, double radius, MaterialPtr mat) {  // rt/circle.go:14-31
    auto c = std::make_shared<Circle>();
    c->normal = normal.Unit(); c->center = center; c->radius = radius; c->mat = mat;
    Vec3 rvec{radius, radius, radius};
    c->bbox = NewAABBFromPoints(center.Sub(rvec), center.Add(rvec));
    return c;
}
HittablePtr Pyramid(Point3 baseCenter, double baseSize, double height, MaterialPtr mat) {  // rt/primitives.go:39-71
    auto sides = NewHittableList();
    sides->Add(NewQuad({baseCenter.X - baseSize / 2, baseCenter.Y, baseCenter.Z - baseSize / 2}, {baseSize, 0, 0}, {0, 0, baseSize}, mat));
    Point3 apex{baseCenter.X, baseCenter.Y + height, baseCenter.Z};
    double halfSize = baseSize / 2;
    Point3 corners[4] = {{baseCenter.X + halfSize, baseCenter.Y, baseCenter.Z - halfSize}, {baseCenter.X + halfSize, baseCenter.Y, baseCenter.Z + halfSize},
                         {baseCenter.X - halfSize, baseCenter.Y, baseCenter.Z + halfSize}, {baseCenter.X - halfSize, baseCenter.Y, baseCenter.Z - halfSize}};
    for (int i = 0; i < 4; i++) sides->Add(NewTriangle(corners[i], corners[(i + 1) % 4], apex, mat));
    return sides;
}
HittablePtr Box(Point3 a, Point3 b, MaterialPtr mat) {  // rt/primitives.go:5-37
    auto sides = NewHittableList();
    Point3 mn{std::fmin(a.X, b.X), std::fmin(a.Y, b.Y), std::fmin(a.Z, b.Z)};
    Point3 mx{std::fmax(a.X, b.X), std::fmax(a.Y, b.Y), std::fmax(a.Z, b.Z)};
    Vec3 dx{mx.X - mn.X, 0, 0}, dy{0, mx.Y - mn.Y, 0}, dz{0, 0, mx.Z - mn.Z};
    sides->Add(NewQuad({mn.X, mn.Y, mx.Z}, dx, dy, mat));        // front
    sides->Add(NewQuad({mx.X, mn.Y, mx.Z}, dz.Neg(), dy, mat));  // right
    sides->Add(NewQuad({mx.X, mn.Y, mn.Z}, dx.Neg(), dy, mat));  // back
    sides->Add(NewQuad({mn.X, mn.Y, mn.Z}, dz, dy, mat));        // left
    sides->Add(NewQuad({mn.X, mx.Y, mx.Z}, dx, dz.Neg(), mat));  // top
    sides->Add(NewQuad({mn.X, mn.Y, mn.Z}, dx, dz, mat));        // bottom
    return sides;
}

// ---- transforms (rt/transform.go) -----------------------------------------------------------------------
std::shared_ptr<Translate> NewTranslate(HittablePtr obj, Vec3 offset) {
    auto t = std::make_shared<Translate>();
    t->Obj = obj; t->Offset = offset; t->bbox = obj->BoundingBox().Translate(offset);
    return t;
}
std::shared_ptr<RotateY> Ry(HittablePtr obj, double angle) {  // rt/transform.go:120-157
    auto r = std::make_shared<RotateY>();
    double radians = DegreesToRadians(angle);
    r->Obj = obj; r->SinTheta = std::sin(radians); r->CosTheta = std::cos(radians);
    AABB bb = obj->BoundingBox();
    Point3 mn{kInf, kInf, kInf}, mx{-kInf, -kInf, -kInf};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
                double x = i * bb.X.Max + (1 - i) * bb.X.Min;
                double y = j * bb.Y.Max + (1 - j) * bb.Y.Min;
                double z = k * bb.Z.Max + (1 - k) * bb.Z.Min;
                double nx = r->CosTheta * x + r->SinTheta * z;
                double nz = -r->SinTheta * x + r->CosTheta * z;
                mn.X = std::fmin(mn.X, nx); mx.X = std::fmax(mx.X, nx);
                mn.Y = std::fmin(mn.Y, y);  mx.Y = std::fmax(mx.Y, y);
                mn.Z = std::fmin(mn.Z, nz); mx.Z = std::fmax(mx.Z, nz);
            }
    r->bbox = NewAABBFromPoints(mn, mx);
    return r;
}
std::shared_ptr<Scale> NewScale(HittablePtr obj, Vec3 f) {  // rt/transform.go:367-402
    auto s = std::make_shared<Scale>();
    s->Obj = obj; s->Factor = f; s->InvFactor = {1.0 / f.X, 1.0 / f.Y, 1.0 / f.Z};
    AABB bb = obj->BoundingBox();
    Point3 mn{bb.X.Min * f.X, bb.Y.Min * f.Y, bb.Z.Min * f.Z}, mx{bb.X.Max * f.X, bb.Y.Max * f.Y, bb.Z.Max * f.Z};
    if (mn.X > mx.X) std::swap(mn.X, mx.X);
    if (mn.Y > mx.Y) std::swap(mn.Y, mx.Y);
    if (mn.Z > mx.Z) std::swap(mn.Z, mx.Z);
    s->bbox = NewAABBFromPoints(mn, mx);
    return s;
}
std::shared_ptr<Scale> NewUniformScale(HittablePtr obj, double f) { return NewScale(obj, {f, f, f}); }
std::shared_ptr<Volume> NewVolume(HittablePtr boundary, double density, TexturePtr tex) {
    auto v = std::make_shared<Volume>();
    v->boundary = boundary; v->negInvDensity = -1.0 / density; v->phaseFunction = NewIsotropic(tex);
    return v;
}
std::shared_ptr<Volume> NewVolumeFromColor(HittablePtr boundary, double density, Color albedo) {
    return NewVolume(boundary, density, NewSolidColor(albedo));
}
HittablePtr Transform::Apply(HittablePtr obj) const {  // rt/transform.go:24-46
    HittablePtr result = obj;
    if (scale.X != 1.0 || scale.Y != 1.0 || scale.Z != 1.0) result = NewScale(result, scale);
    if (rotation.X != 0 || rotation.Z != 0) throw std::runtime_error("rt: RotateX/RotateZ are outside the device hot path (SURVEY 8f)");
    if (rotation.Y != 0) result = Ry(result, rotation.Y);
    if (position.X != 0 || position.Y != 0 || position.Z != 0) result = NewTranslate(result, position);
    return result;
}

// ---- reference-order BVH (rt/bvh.go:64-217): median split on the longest centroid axis, leaves <= 4 ---------
namespace {
struct BvhPrim {
    size_t index;
    AABB bbox;
    Vec3 centroid;
};
// Stable insertion/merge sort with a fully defined behaviour for the NaN keys an infinite Plane produces
// (Go uses the unstable sort.Slice, rt/bvh.go:148; the stable order is this repo's canonical choice).
template <class Less>
void stableSort(std::vector<BvhPrim>& a, size_t lo, size_t hi, std::vector<BvhPrim>& tmp, Less less) {
    size_t n = hi - lo;
    if (n <= 12) {
        for (size_t i = lo + 1; i < hi; i++) {
            BvhPrim x = a[i];
            size_t j = i;
            while (j > lo && less(x, a[j - 1])) { a[j] = a[j - 1]; j--; }
            a[j] = x;
        }
        return;
    }
    size_t mid = lo + n / 2;
    stableSort(a, lo, mid, tmp, less);
    stableSort(a, mid, hi, tmp, less);
    size_t i = lo, j = mid, k = lo;
    while (i < mid && j < hi) tmp[k++] = less(a[j], a[i]) ? a[j++] : a[i++];
    while (i < mid) tmp[k++] = a[i++];
    while (j < hi) tmp[k++] = a[j++];
    for (size_t t = lo; t < hi; t++) a[t] = tmp[t];
}
// The two halves of a split are independent (disjoint ranges of `prims` and `tmp`): the top levels of a big mesh run them on
// separate threads (the reference's recursion is sequential; the tree is the same).
BVHNodePtr buildBVHNode(const std::vector<HittablePtr>& objects, std::vector<BvhPrim>& prims, size_t lo, size_t hi, std::vector<BvhPrim>& tmp, int depth = 0) {
    size_t n = hi - lo;
    AABB bounds = prims[lo].bbox;
    AABB cb = NewAABBFromPoints(prims[lo].centroid, prims[lo].centroid);
    for (size_t i = lo + 1; i < hi; i++) {
        bounds = NewAABBFromBoxes(bounds, prims[i].bbox);
        cb = NewAABBFromBoxes(cb, NewAABBFromPoints(prims[i].centroid, prims[i].centroid));
    }
    auto node = std::make_shared<BVHNode>();
    node->bbox = bounds;
    if (n <= 4) {
        auto leaf = std::make_shared<BVHLeaf>();
        leaf->bbox = bounds;
        for (size_t i = lo; i < hi; i++) leaf->objects.push_back(objects[prims[i].index]);
        node->left = leaf;
        node->right = leaf;  // rt/bvh.go:141 — the same leaf on both sides
        return node;
    }
    int axis = cb.LongestAxis();
    stableSort(prims, lo, hi, tmp, [axis](const BvhPrim& a, const BvhPrim& b) {
        return axis == 0 ? a.centroid.X < b.centroid.X : axis == 1 ? a.centroid.Y < b.centroid.Y : a.centroid.Z < b.centroid.Z;
    });
    size_t mid = lo + n / 2;
    if (n >= 16384 && depth < 4) {
        auto left = std::async(std::launch::async, [&, depth] { return buildBVHNode(objects, prims, lo, mid, tmp, depth + 1); });
        node->right = buildBVHNode(objects, prims, mid, hi, tmp, depth + 1);
        node->left = left.get();
    } else {
        node->left = buildBVHNode(objects, prims, lo, mid, tmp, depth + 1);
        node->right = buildBVHNode(objects, prims, mid, hi, tmp, depth + 1);
    }
    return node;
}
}  // namespace

BVHNodePtr NewBVHNode(const std::vector<HittablePtr>& objects, size_t start, size_t end) {
    size_t n = end - start;
    if (n == 0) return std::make_shared<BVHNode>();
    std::vector<BvhPrim> prims(n), tmp(n);
    for (size_t i = 0; i < n; i++) {
        AABB bb = objects[start + i]->BoundingBox();
        prims[i] = {start + i, bb, bb.Centroid()};
    }
    BVHNodePtr root = buildBVHNode(objects, prims, 0, n, tmp);
    root->src.assign(objects.begin() + start, objects.begin() + end);
    return root;
}
BVHNodePtr NewBVHNodeFromList(const HittableListPtr& list) { return NewBVHNode(list->Objects, 0, list->Objects.size()); }

static bool g_eager_mesh_bvh = std::getenv("RT_EAGER_BVH") && std::atoi(std::getenv("RT_EAGER_BVH")) != 0;
void SetEagerMeshBVH(bool eager) { g_eager_mesh_bvh = eager; }
bool EagerMeshBVH() { return g_eager_mesh_bvh; }
// The root of a mesh without its tree: the bounding box is what the world BVH / the flattener need from it.
BVHNodePtr NewBVHNodeDeferred(std::vector<HittablePtr>&& objects) {
    auto root = std::make_shared<BVHNode>();
    if (objects.empty()) return root;
    const size_t n = objects.size();
    const size_t T = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(1, n / 32768));
    std::vector<AABB> part(T);
    std::vector<std::thread> th;
    auto work = [&](size_t t) {
        const size_t lo = n * t / T, hi = n * (t + 1) / T;
        AABB b = objects[lo]->BoundingBox();
        for (size_t i = lo + 1; i < hi; i++) b = NewAABBFromBoxes(b, objects[i]->BoundingBox());
        part[t] = b;
    };
    for (size_t t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    AABB b = part[0];
    for (size_t t = 1; t < T; t++) b = NewAABBFromBoxes(b, part[t]);   // min / max: the union does not depend on the grouping
    root->bbox = b;
    root->src = std::move(objects);
    root->deferred = true;
    return root;
}
BVHNodePtr NewMeshRootFromSoup(std::shared_ptr<BVHNode::Soup> soup, int threads) {
    auto root = std::make_shared<BVHNode>();
    const size_t n = soup->n;
    if (n == 0) return root;
    const size_t T = std::min<size_t>(threads > 0 ? (size_t)threads : std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(1, n / 32768));
    std::vector<AABB> part(T);
    std::vector<std::thread> th;
    auto work = [&](size_t t) {
        AABB b;
        bool first = true;
        for (size_t k = n * t / T, e = n * (t + 1) / T; k < e; k++) {   // the box NewTriangle gives every face (rt/triangle.go:27-37), unioned
            const double *a = &soup->v0[3 * k], *bb = &soup->v1[3 * k], *c = &soup->v2[3 * k];
            AABB tb = NewAABBFromPoints({std::fmin(a[0], std::fmin(bb[0], c[0])), std::fmin(a[1], std::fmin(bb[1], c[1])), std::fmin(a[2], std::fmin(bb[2], c[2]))},
                                        {std::fmax(a[0], std::fmax(bb[0], c[0])), std::fmax(a[1], std::fmax(bb[1], c[1])), std::fmax(a[2], std::fmax(bb[2], c[2]))});
            b = first ? tb : NewAABBFromBoxes(b, tb);
            first = false;
        }
        part[t] = b;
    };
    for (size_t t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    AABB b = part[0];
    for (size_t t = 1; t < T; t++) b = NewAABBFromBoxes(b, part[t]);
    root->bbox = b;
    root->soup = std::move(soup);
    root->deferred = true;
    return root;
}
void BVHNode::EnsureBuilt() {
    if (!deferred) return;
    if (soup && src.empty()) {
        src.resize(soup->n);
        for (size_t k = 0; k < soup->n; k++)
            src[k] = NewTriangle({soup->v0[3 * k], soup->v0[3 * k + 1], soup->v0[3 * k + 2]}, {soup->v1[3 * k], soup->v1[3 * k + 1], soup->v1[3 * k + 2]},
                                 {soup->v2[3 * k], soup->v2[3 * k + 1], soup->v2[3 * k + 2]}, soup->mat);
    }
    BVHNodePtr built = NewBVHNode(src, 0, src.size());
    left = built->left; right = built->right; bbox = built->bbox;
    deferred = false;
}

// rt/obj_loader.go: rt_obj.cpp (parallel text parse)

// ---- rt/image_loader.go:122-383 -----------------------------------------------------------------------------
static bool fileExists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}
std::string FindAsset(const std::string& filename, const std::string& assetType) {
    std::vector<std::string> paths = {filename, assetType + "/" + filename, "assets/" + assetType + "/" + filename,
                                      "../" + assetType + "/" + filename, "../assets/" + assetType + "/" + filename};
    for (auto& p : paths)
        if (fileExists(p)) return p;
    return "";
}
bool LoadHDR(const std::string& path, int& width, int& height, std::vector<double>& rgb, std::string* err) {
    auto fail = [&](const std::string& m) { if (err) *err = m; return false; };
    std::ifstream in(path, std::ios::binary);
    if (!in) return fail("could not open HDR file " + path);
    std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    size_t pos = 0;
    auto readLine = [&](std::string& out) {
        size_t nl = data.find('\n', pos);
        if (nl == std::string::npos) return false;
        out = data.substr(pos, nl - pos);
        pos = nl + 1;
        return true;
    };
    auto trim = [](std::string s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    std::string line;
    if (!readLine(line) || line.rfind("#?", 0) != 0) return fail("not a valid Radiance HDR file (missing #? signature)");
    while (true) {
        if (!readLine(line)) return fail("unexpected end of header");
        if (trim(line).empty()) break;
    }
    if (!readLine(line)) return fail("failed to read resolution");
    std::istringstream rs(trim(line));
    std::string a, b, c, d;
    rs >> a >> b >> c >> d;
    if (a == "-Y" && c == "+X") { height = std::atoi(b.c_str()); width = std::atoi(d.c_str()); }
    else if (a == "+X" && c == "-Y") { width = std::atoi(b.c_str()); height = std::atoi(d.c_str()); }
    else return fail("unsupported resolution format: " + line);
    if (width <= 0 || height <= 0) return fail("invalid resolution");
    rgb.assign((size_t)3 * width * height, 0.0);
    const unsigned char* p = (const unsigned char*)data.data();
    size_t n = data.size();
    std::vector<unsigned char> scan((size_t)4 * width);
    auto put = [&](int y, int x, const unsigned char* q) {  // rgbeToColor, rt/image_loader.go:364-383
        size_t idx = ((size_t)y * width + x) * 3;
        if (q[3] == 0) return;
        double scale = std::ldexp(1.0, (int)q[3] - 128 - 8);
        rgb[idx] = (q[0] + 0.5) * scale; rgb[idx + 1] = (q[1] + 0.5) * scale; rgb[idx + 2] = (q[2] + 0.5) * scale;
    };
    for (int y = 0; y < height; y++) {
        if (pos + 4 > n) return fail("failed to read scanline header");
        const unsigned char* h = p + pos;
        pos += 4;
        if (h[0] == 2 && h[1] == 2) {
            int sw = (h[2] << 8) | h[3];
            if (sw != width) return fail("scanline width mismatch");
            for (int comp = 0; comp < 4; comp++) {
                int x = 0;
                while (x < width) {
                    if (pos >= n) return fail("failed to read RLE code");
                    int code = p[pos++];
                    if (code > 128) {
                        int count = code - 128;
                        if (pos >= n) return fail("failed to read RLE value");
                        unsigned char v = p[pos++];
                        for (int i = 0; i < count && x < width; i++) scan[(size_t)comp * width + x++] = v;
                    } else {
                        for (int i = 0; i < code && x < width; i++) {
                            if (pos >= n) return fail("failed to read raw value");
                            scan[(size_t)comp * width + x++] = p[pos++];
                        }
                    }
                }
            }
            for (int x = 0; x < width; x++) {
                unsigned char q[4] = {scan[x], scan[(size_t)width + x], scan[(size_t)2 * width + x], scan[(size_t)3 * width + x]};
                put(y, x, q);
            }
        } else {  // flat RGBE: the 4 header bytes are the first pixel
            put(y, 0, h);
            for (int x = 1; x < width; x++) {
                if (pos + 4 > n) return fail("failed to read pixel");
                put(y, x, p + pos);
                pos += 4;
            }
        }
    }
    return true;
}
std::shared_ptr<HDRIEnvironment> NewHDRIEnvironment(const std::string& filename) {
    auto env = std::make_shared<HDRIEnvironment>();
    std::string path = FindAsset(filename, "hdri");
    std::string err;
    if (path.empty() || !LoadHDR(path, env->width, env->height, env->rgb, &err)) {
        std::fprintf(stderr, "Warning: Failed to load HDRI '%s' %s\n", filename.c_str(), err.c_str());
        env->rgb.clear();
        env->width = env->height = 0;
    }
    return env;
}

// ---- camera (rt/camera.go:106-169, :277-299) --------------------------------------------------------------
CameraPreset QuickPreview() { return {16.0 / 9.0, 400, 10, 10, 20, 0.0, 10.0, {13, 2, 3}, {0, 0, 0}, {0, 1, 0}, false, {0, 0, 0}, {0.5, 0.7, 1.0}, true}; }
CameraPreset StandardQuality() { return {16.0 / 9.0, 600, 100, 50, 20, 0.6, 10.0, {13, 2, 3}, {0, 0, 0}, {0, 1, 0}, false, {0, 0, 0}, {0.5, 0.7, 1.0}, false}; }
CameraPreset HighQuality() { return {16.0 / 9.0, 1200, 500, 50, 20, 0.6, 10.0, {13, 2, 3}, {0, 0, 0}, {0, 1, 0}, false, {0, 0, 0}, {0.5, 0.7, 1.0}, false}; }
void Camera::ApplyPreset(const CameraPreset& p) {  // UseSkyGradient is NOT copied (rt/camera.go:155-169)
    AspectRatio = p.AspectRatio; ImageWidth = p.ImageWidth; SamplesPerPixel = p.SamplesPerPixel; MaxDepth = p.MaxDepth;
    Vfov = p.Vfov; DefocusAngle = p.DefocusAngle; FocusDist = p.FocusDist; LookFrom = p.LookFrom; LookAt = p.LookAt;
    Vup = p.Vup; FreeCamera = p.FreeCamera; Forward = p.Forward; Background = p.Background;
}
void Camera::Initialize() { ImageHeight = std::max((int)((double)ImageWidth / AspectRatio), 1); }
std::shared_ptr<Camera> Camera::Build() {
    Initialize();
    return std::make_shared<Camera>(*this);
}
void Camera::FillDesc(rtx_camera_desc& d) const {
    std::memset(&d, 0, sizeof(d));
    d.aspect_ratio = AspectRatio; d.image_width = ImageWidth; d.samples_per_pixel = SamplesPerPixel; d.max_depth = MaxDepth;
    d.vfov = Vfov;
    auto put = [](double* o, const Vec3& v) { o[0] = v.X; o[1] = v.Y; o[2] = v.Z; };
    put(d.look_from, LookFrom); put(d.look_at, LookAt); put(d.vup, Vup);
    d.defocus_angle = DefocusAngle; d.focus_dist = FocusDist;
    put(d.look_from2, LookFrom2); put(d.look_at2, LookAt2);
    d.camera_motion = CameraMotion; d.free_camera = FreeCamera;
    put(d.forward, Forward); put(d.background, Background);
    d.use_sky_gradient = UseSkyGradient; d.phantom_hdri = PhantomHDRI;
    d.has_derived = 0;
}

}  // namespace rt
