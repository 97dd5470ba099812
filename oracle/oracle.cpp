// oracle.cpp — TEST INFRASTRUCTURE ONLY. CPU (float64) restatement of byvfx/go-raytracing's per-pixel
// path-tracing hot path, used as the parity checker and as the timed CPU baseline ("port").
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library. The product (go-raytracing_b200/) never links, imports or calls it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer fixtures for this path
//  and its Go toolchain is absent from this image, so this restatement cannot be checked
// against outputs of the reference itself. It is pinned instead by hand-derivable known answers
// (tests/test_oracle_kat.py), by the HDRI total-power figure of the shipped .hdr file and by a coarse
// region-mean comparison with the reference's committed image.png.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference/).
// All arithmetic is IEEE float64 with no FMA contraction (built with -ffp-contract=off), like Go on amd64.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <list>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../include/rtx_b200.h"

namespace orc {

static const double kInf = std::numeric_limits<double>::infinity();

// ---- rt/vec3.go -----------------------------------------------------------------------------------------
struct Vec3 {
    double X = 0, Y = 0, Z = 0;
    Vec3() = default;
    Vec3(double x, double y, double z) : X(x), Y(y), Z(z) {}
    Vec3 Add(const Vec3& u) const { return {X + u.X, Y + u.Y, Z + u.Z}; }
    Vec3 Sub(const Vec3& u) const { return {X - u.X, Y - u.Y, Z - u.Z}; }
    Vec3 Mult(const Vec3& u) const { return {X * u.X, Y * u.Y, Z * u.Z}; }
    Vec3 Scale(double t) const { return {t * X, t * Y, t * Z}; }
    Vec3 Div(double t) const { return Scale(1 / t); }  // rt/vec3.go:27: multiply by the reciprocal
    Vec3 Neg() const { return {-X, -Y, -Z}; }
    double Len2() const { return X * X + Y * Y + Z * Z; }
    double Len() const { return std::sqrt(Len2()); }
    Vec3 Unit() const {  // rt/vec3.go:32-38
        double l = Len();
        if (l == 0) return *this;
        return Div(l);
    }
    bool NearZero() const {  // rt/vec3.go:40-43
        const double s = 1e-8;
        return std::fabs(X) < s && std::fabs(Y) < s && std::fabs(Z) < s;
    }
};
using Point3 = Vec3;
using Color = Vec3;
static inline double Dot(const Vec3& a, const Vec3& b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }
static inline Vec3 Cross(const Vec3& a, const Vec3& b) {
    return {a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X};
}
static inline Vec3 Reflect(const Vec3& v, const Vec3& n) { return v.Sub(n.Scale(2 * Dot(v, n))); }  // rt/vec3.go:106
static inline Vec3 Refract(const Vec3& uv, const Vec3& n, double etaiOverEtat) {                    // rt/vec3.go:110-117
    double cosTheta = std::fmin(Dot(uv.Neg(), n), 1.0);
    Vec3 rOutPerp = uv.Add(n.Scale(cosTheta)).Scale(etaiOverEtat);
    Vec3 rOutParallel = n.Scale(-std::sqrt(std::fabs(1.0 - rOutPerp.Len2())));
    return rOutPerp.Add(rOutParallel);
}

// ---- rt/utils.go ------------------------------------------------------------------------------------------
constexpr double Pi = 3.1415926535897932385;                                              // rt/utils.go:11
static inline double DegreesToRadians(double d) { return d * Pi / 180.0; }                // rt/utils.go:14
static inline double LinearToGamma(double l) { return l > 0 ? std::sqrt(l) : 0; }         // rt/utils.go:85-90

// RNG: the reference calls Go's auto-seeded global math/rand (rt/utils.go:18). Here: one xoshiro256++ stream
// per worker thread, 53-bit doubles in [0,1) like rand.Float64.
struct Rng {
    uint64_t s[4];
    explicit Rng(uint64_t seed = 1) { reseed(seed); }
    void reseed(uint64_t seed) {
        for (int i = 0; i < 4; i++) {
            uint64_t z = (seed += 0x9E3779B97F4A7C15ull);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double Float64() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};
static thread_local Rng t_rng(0x1234);
static inline double RandomDouble() { return t_rng.Float64(); }                                   // rt/utils.go:18
static inline double RandomDoubleRange(double mn, double mx) { return mn + (mx - mn) * RandomDouble(); }  // :22
static inline Vec3 RandomVec3Range(double mn, double mx) {                                        // rt/vec3.go:98-104
    double x = RandomDoubleRange(mn, mx), y = RandomDoubleRange(mn, mx), z = RandomDoubleRange(mn, mx);
    return {x, y, z};
}
static inline Vec3 RandomUnitVector() {  // rt/vec3.go:45-54
    for (;;) {
        Vec3 p = RandomVec3Range(-1, 1);
        double lensq = p.Len2();
        if (1e-160 < lensq && lensq <= 1) return p.Div(std::sqrt(lensq));
    }
}
static inline Vec3 RandomInUnitDisk() {  // rt/vec3.go:66-77
    for (;;) {
        double x = RandomDoubleRange(-1, 1), y = RandomDoubleRange(-1, 1);
        Vec3 p{x, y, 0};
        if (p.Len2() < 1) return p;
    }
}

// ---- counters (rt/profiler.go:65-78). The reference bumps global atomics inside the innermost loops. -------
struct Stats {
    std::atomic<int64_t> RayCount{0}, BVHIntersections{0}, SamplesComputed{0}, PixelsRendered{0};
    std::atomic<int64_t> ShadowQueries{0};  // not in the reference (its ShadowRays counter is never incremented)
};
static Stats g_stats;
static bool g_use_atomics = true;            // "atomics off" variant of the CPU baseline
static bool g_volumes_transparent = false;   // level-1 parity: Volume.Hit draws random numbers (rt/volume.go:66)
#define STAT_ADD(c) do { if (g_use_atomics) g_stats.c.fetch_add(1, std::memory_order_relaxed); } while (0)

// ---- rt/ray.go, rt/interval.go ------------------------------------------------------------------------------
struct Ray {
    Point3 orig;
    Vec3 dir;
    double tm = 0;
    Point3 At(double t) const { return orig.Add(dir.Scale(t)); }
};
struct Interval {
    double Min = kInf, Max = -kInf;  // empty
    double Size() const { return Max - Min; }
    bool Contains(double x) const { return Min <= x && x <= Max; }   // rt/interval.go:56
    bool Surrounds(double x) const { return Min < x && x < Max; }    // rt/interval.go:61
    double Clamp(double x) const { return x < Min ? Min : (x > Max ? Max : x); }  // rt/interval.go:64-72
    Interval Expand(double d) const { return {Min - d, Max + d}; }
    Interval Add(double d) const { return {Min + d, Max + d}; }
};
static const Interval UniverseInterval{-kInf, kInf};
static inline Interval IntervalFromIntervals(const Interval& a, const Interval& b) {  // rt/interval.go:28-40
    double mn = a.Min;
    if (b.Min < a.Min) mn = b.Min;
    double mx = a.Max;
    if (b.Max > a.Max) mx = b.Max;
    return {mn, mx};
}
static inline double goMin(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? NAN : std::fmin(a, b); }  // math.Min
static inline double goMax(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? NAN : std::fmax(a, b); }  // math.Max

// ---- rt/aabb.go ------------------------------------------------------------------------------------------------
struct AABB {
    Interval X, Y, Z;
    void padToMinimums() {  // :117-128
        const double delta = 0.0001;
        if (X.Size() < delta) X = X.Expand(delta);
        if (Y.Size() < delta) Y = Y.Expand(delta);
        if (Z.Size() < delta) Z = Z.Expand(delta);
    }
    bool Hit(const Ray& r, Interval rayT) const {  // :59-116, unrolled per axis exactly as written
        const Point3& o = r.orig;
        const Vec3& d = r.dir;
        double adinv = 1.0 / d.X;
        double t0 = (X.Min - o.X) * adinv, t1 = (X.Max - o.X) * adinv;
        if (adinv < 0) std::swap(t0, t1);
        if (t0 > rayT.Min) rayT.Min = t0;
        if (t1 < rayT.Max) rayT.Max = t1;
        if (rayT.Max <= rayT.Min) return false;
        adinv = 1.0 / d.Y;
        t0 = (Y.Min - o.Y) * adinv; t1 = (Y.Max - o.Y) * adinv;
        if (adinv < 0) std::swap(t0, t1);
        if (t0 > rayT.Min) rayT.Min = t0;
        if (t1 < rayT.Max) rayT.Max = t1;
        if (rayT.Max <= rayT.Min) return false;
        adinv = 1.0 / d.Z;
        t0 = (Z.Min - o.Z) * adinv; t1 = (Z.Max - o.Z) * adinv;
        if (adinv < 0) std::swap(t0, t1);
        if (t0 > rayT.Min) rayT.Min = t0;
        if (t1 < rayT.Max) rayT.Max = t1;
        if (rayT.Max <= rayT.Min) return false;
        return true;
    }
    int LongestAxis() const {  // :139-150
        double xs = X.Size(), ys = Y.Size(), zs = Z.Size();
        if (xs > ys && xs > zs) return 0;
        if (ys > zs) return 1;
        return 2;
    }
    Vec3 Centroid() const { return {(X.Min + X.Max) * 0.5, (Y.Min + Y.Max) * 0.5, (Z.Min + Z.Max) * 0.5}; }  // :153-159
    AABB Translate(const Vec3& off) const;
};
static inline AABB AABBFromIntervals(Interval x, Interval y, Interval z) { AABB b{x, y, z}; b.padToMinimums(); return b; }  // :26-30
static inline AABB AABBFromPoints(const Point3& a, const Point3& b) {                                                        // :32-40
    AABB box{{goMin(a.X, b.X), goMax(a.X, b.X)}, {goMin(a.Y, b.Y), goMax(a.Y, b.Y)}, {goMin(a.Z, b.Z), goMax(a.Z, b.Z)}};
    box.padToMinimums();
    return box;
}
static inline AABB AABBFromBoxes(const AABB& a, const AABB& b) {  // :42-48
    return {IntervalFromIntervals(a.X, b.X), IntervalFromIntervals(a.Y, b.Y), IntervalFromIntervals(a.Z, b.Z)};
}
AABB AABB::Translate(const Vec3& off) const { return AABBFromIntervals(X.Add(off.X), Y.Add(off.Y), Z.Add(off.Z)); }  // :130-136

// ---- rt/texture.go ---------------------------------------------------------------------------------------------------
struct Texture {
    virtual ~Texture() = default;
    virtual Color Value(double u, double v, const Point3& p) const = 0;
};
struct SolidColor : Texture {
    Color Albedo;
    Color Value(double, double, const Point3&) const override { return Albedo; }  // :43-45
};
struct CheckerTexture : Texture {
    double invScale;
    const Texture *even, *odd;
    Color Value(double u, double v, const Point3& p) const override {  // :63-77
        const double epsilon = 1e-4;
        long long xi = (long long)std::floor(invScale * p.X + epsilon);
        long long yi = (long long)std::floor(invScale * p.Y + epsilon);
        long long zi = (long long)std::floor(invScale * p.Z + epsilon);
        bool isEven = (xi + yi + zi) % 2 == 0;
        return isEven ? even->Value(u, v, p) : odd->Value(u, v, p);
    }
};

// ---- rt/noise.go, rt/texture.go:19-29, :81-85 -------------------------------------------------------------------------
struct Perlin {  // the tables come with the scene (the reference draws them from Go's global source, rt/noise.go:15-28)
    Vec3 randvec[256];
    int permX[256], permY[256], permZ[256];
    double Noise(const Point3& pt) const {  // :30-53
        double u = pt.X - std::floor(pt.X), v = pt.Y - std::floor(pt.Y), w = pt.Z - std::floor(pt.Z);
        int i = (int)std::floor(pt.X), j = (int)std::floor(pt.Y), k = (int)std::floor(pt.Z);
        Vec3 c[2][2][2];
        for (int di = 0; di < 2; di++)
            for (int dj = 0; dj < 2; dj++)
                for (int dk = 0; dk < 2; dk++) c[di][dj][dk] = randvec[permX[(i + di) & 255] ^ permY[(j + dj) & 255] ^ permZ[(k + dk) & 255]];
        double accum = 0.0;  // perlinInterp :79-92
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++)
                for (int dk = 0; dk < 2; dk++) {
                    Vec3 weightV{u - (double)a, v - (double)b, w - (double)dk};
                    accum += ((double)a * u + (1 - (double)a) * (1 - u)) * ((double)b * v + (1 - (double)b) * (1 - v)) *
                             ((double)dk * w + (1 - (double)dk) * (1 - w)) * Dot(c[a][b][dk], weightV);
                }
        return accum;
    }
    double Turb(const Point3& pt, int depth) const {  // :55-65
        double accum = 0.0, weight = 1.0;
        Point3 tempPt = pt;
        for (int i = 0; i < depth; i++) {
            accum += weight * Noise(tempPt);
            weight *= 0.5;
            tempPt = tempPt.Scale(2);
        }
        return std::fabs(accum);
    }
};
struct NoiseTexture : Texture {
    const Perlin* noise;
    double scale;
    Color Value(double, double, const Point3& p) const override {  // rt/texture.go:81-85
        double s = scale * p.Z + 10.0 * noise->Turb(p.Scale(scale), 7);
        double turbValue = 0.5 * (1.0 + std::sin(s));
        return Color{1, 1, 1}.Scale(turbValue);
    }
};

// ---- rt/image_texture.go, rt/image_loader.go:97-120 ---------------------------------------------------------------------------------
struct ImageTexture : Texture {
    int width = 0, height = 0;
    const double* data = nullptr;  // ImageLoader.data (3 doubles per pixel, after the load-time sqrt)
    static int clampi(int x, int low, int high) { return x < low ? low : (x < high ? x : high - 1); }
    Color Value(double u, double v, const Point3&) const override {  // rt/image_texture.go:27-43
        if (height <= 0) return Color{0, 1, 1};
        u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        v = 1.0 - (v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v));
        int i = clampi((int)(u * (double)width), 0, width), j = clampi((int)(v * (double)height), 0, height);
        const double* px = data + 3 * ((size_t)j * width + i);
        return Color{px[0], px[1], px[2]};
    }
};

// ---- rt/hittable.go --------------------------------------------------------------------------------------------------
struct Material;
struct HitRecord {
    Point3 P;
    Vec3 Normal;
    const Material* Mat = nullptr;
    double U = 0, V = 0, T = 0;
    bool FrontFace = false;
    int entry = -1, prim = -1;  // bookkeeping of this oracle (level-1 identifiers), not in the reference
    void SetFaceNormal(const Ray& r, const Vec3& outwardNormal) {  // :20-30
        FrontFace = Dot(r.dir, outwardNormal) < 0;
        Normal = FrontFace ? outwardNormal : outwardNormal.Neg();
    }
};
struct Hittable {
    virtual ~Hittable() = default;
    virtual bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const = 0;
    virtual AABB BoundingBox() const = 0;
};

// ---- rt/material.go ------------------------------------------------------------------------------------------------------
struct Material {
    virtual ~Material() = default;
    virtual bool Scatter(const Ray& rIn, const HitRecord* rec, Color* attenuation, Ray* scattered) const = 0;
    virtual Color Emitted(double, double, const Point3&) const { return {0, 0, 0}; }
    virtual double PDF(const Vec3& wi, const Vec3& wo, const Vec3& normal) const = 0;
    virtual bool CanUseNEE() const { return false; }  // MaterialProperties.CanUseNEE
};
struct Lambertian : Material {
    const Texture* tex;
    bool CanUseNEE() const override { return true; }  // :49-55
    bool Scatter(const Ray& rIn, const HitRecord* rec, Color* attenuation, Ray* scattered) const override {  // :57-68
        Vec3 dir = rec->Normal.Add(RandomUnitVector());
        if (dir.NearZero()) dir = rec->Normal;
        *scattered = Ray{rec->P, dir, rIn.tm};
        *attenuation = tex->Value(rec->U, rec->V, rec->P);
        return true;
    }
    double PDF(const Vec3&, const Vec3& wo, const Vec3& normal) const override {  // :70-76
        double c = Dot(normal, wo);
        if (c < 0) return 0;
        return c / M_PI;
    }
};
struct Metal : Material {
    Color Albedo;
    double Fuzz;
    bool Scatter(const Ray& rIn, const HitRecord* rec, Color* attenuation, Ray* scattered) const override {  // :113-119
        Vec3 reflected = Reflect(rIn.dir, rec->Normal);
        reflected = reflected.Unit().Add(RandomUnitVector().Scale(Fuzz));
        *scattered = Ray{rec->P, reflected, rIn.tm};
        *attenuation = Albedo;
        return Dot(scattered->dir, rec->Normal) > 0;
    }
    double PDF(const Vec3& wi, const Vec3& wo, const Vec3& normal) const override {  // :121-136 (dead code on the hot path)
        if (Fuzz == 0) return 0;
        Vec3 reflected = Reflect(wi.Scale(-1), normal);
        double cosAlpha = Dot(reflected, wo);
        if (cosAlpha < 0) return 0;
        double exponent = (1.0 - Fuzz) * 50.0;
        return (exponent + 1) / (2 * M_PI) * std::pow(cosAlpha, exponent);
    }
};
static inline double reflectance(double cosine, double ri) {  // :284-288
    double r0 = (1 - ri) / (1 + ri);
    r0 = r0 * r0;
    return r0 + (1 - r0) * std::pow(1 - cosine, 5);
}
struct Dielectric : Material {
    double RefractionIndex;
    bool Scatter(const Ray& rIn, const HitRecord* rec, Color* attenuation, Ray* scattered) const override {  // :164-188
        *attenuation = {1.0, 1.0, 1.0};
        double ri = rec->FrontFace ? 1.0 / RefractionIndex : RefractionIndex;
        Vec3 unitDirection = rIn.dir.Unit();
        double cosTheta = std::fmin(Dot(unitDirection.Neg(), rec->Normal), 1.0);
        double sinTheta = std::sqrt(1.0 - cosTheta * cosTheta);
        bool cannotRefract = ri * sinTheta > 1.0;
        Vec3 direction;
        if (cannotRefract || reflectance(cosTheta, ri) > RandomDouble()) direction = Reflect(unitDirection, rec->Normal);
        else direction = Refract(unitDirection, rec->Normal, ri);
        *scattered = Ray{rec->P, direction, rIn.tm};
        return true;
    }
    double PDF(const Vec3&, const Vec3&, const Vec3&) const override { return 0; }
};
struct DiffuseLight : Material {
    const Texture* tex;
    bool Scatter(const Ray&, const HitRecord*, Color*, Ray*) const override { return false; }                // :226-228
    Color Emitted(double u, double v, const Point3& p) const override { return tex->Value(u, v, p); }         // :234-236
    double PDF(const Vec3&, const Vec3&, const Vec3&) const override { return 0; }
};
struct Isotropic : Material {
    const Texture* tex;
    bool Scatter(const Ray& rIn, const HitRecord* rec, Color* attenuation, Ray* scattered) const override {  // :266-270
        *scattered = Ray{rec->P, RandomUnitVector(), rIn.tm};
        *attenuation = tex->Value(rec->U, rec->V, rec->P);
        return true;
    }
    double PDF(const Vec3&, const Vec3&, const Vec3&) const override { return 1.0 / (4.0 * M_PI); }
};

// ---- rt/sphere.go ----------------------------------------------------------------------------------------------------------
struct Sphere : Hittable {
    Ray Center;
    double Radius;
    const Material* Mat;
    AABB bbox;
    int id = 0;
    static Sphere* New(Point3 c1, Vec3 velocity, double radius, const Material* m) {  // :14-43
        auto s = new Sphere();
        Vec3 rvec{radius, radius, radius};
        s->Center = Ray{c1, velocity, 0};
        s->Radius = std::fmax(0.0, radius);
        s->Mat = m;
        Point3 c2 = c1.Add(velocity);
        AABB b1 = AABBFromPoints(c1.Sub(rvec), c1.Add(rvec));
        if (velocity.X == 0 && velocity.Y == 0 && velocity.Z == 0) s->bbox = b1;  // NewSphere
        else s->bbox = AABBFromBoxes(b1, AABBFromPoints(c2.Sub(rvec), c2.Add(rvec)));  // NewMovingSphere
        return s;
    }
    AABB BoundingBox() const override { return bbox; }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :63-94
        Point3 sphereCenter = Center.At(r.tm);
        Vec3 oc = sphereCenter.Sub(r.orig);
        double a = r.dir.Len2();
        double h = Dot(r.dir, oc);
        double c = oc.Len2() - Radius * Radius;
        double discriminant = h * h - a * c;
        if (discriminant < 0) return false;
        double sqrtd = std::sqrt(discriminant);
        double root = (h - sqrtd) / a;
        if (!rayT.Surrounds(root)) {
            root = (h + sqrtd) / a;
            if (!rayT.Surrounds(root)) return false;
        }
        rec->T = root;
        rec->P = r.At(rec->T);
        Vec3 outwardNormal = rec->P.Sub(sphereCenter).Div(Radius);
        rec->SetFaceNormal(r, outwardNormal);
        double theta = std::acos(-outwardNormal.Y);                        // getSphereUV :53-59
        double phi = std::atan2(-outwardNormal.Z, outwardNormal.X) + M_PI;
        rec->U = phi / (2 * M_PI);
        rec->V = theta / M_PI;
        rec->Mat = Mat;
        rec->prim = id;
        return true;
    }
};

// ---- rt/quad.go ------------------------------------------------------------------------------------------------------------
struct Quad : Hittable {
    Point3 Q;
    Vec3 u, v, w, normal;
    double D;
    const Material* mat;
    AABB bbox;
    int id = 0;
    static Quad* New(Point3 Q, Vec3 u, Vec3 v, const Material* m) {  // :16-42
        auto q = new Quad();
        q->Q = Q; q->u = u; q->v = v; q->mat = m;
        Vec3 n = Cross(u, v);
        q->normal = n.Unit();
        q->D = Dot(q->normal, Q);
        q->w = n.Scale(1.0 / Dot(n, n));
        AABB d1 = AABBFromPoints(Q, Q.Add(u).Add(v));
        AABB d2 = AABBFromPoints(Q.Add(u), Q.Add(v));
        q->bbox = AABBFromBoxes(d1, d2);
        return q;
    }
    AABB BoundingBox() const override { return bbox; }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :44-84
        double denom = Dot(normal, r.dir);
        if (std::fabs(denom) < 1e-8) return false;
        double t = (D - Dot(normal, r.orig)) / denom;
        if (!rayT.Contains(t)) return false;
        Point3 intersection = r.At(t);
        Vec3 planar = intersection.Sub(Q);
        double alpha = Dot(w, Cross(planar, v));
        double beta = Dot(w, Cross(u, planar));
        Interval unit{0, 1};
        if (!unit.Contains(alpha) || !unit.Contains(beta)) return false;  // isInterior :72-84
        rec->U = alpha; rec->V = beta;
        rec->T = t; rec->P = intersection; rec->Mat = mat;
        rec->SetFaceNormal(r, normal);
        rec->prim = id;
        return true;
    }
    Point3 SamplePoint() const {  // :87-92
        double alpha = RandomDouble(), beta = RandomDouble();
        return Q.Add(u.Scale(alpha)).Add(v.Scale(beta));
    }
    double Area() const { return Cross(u, v).Len(); }  // :95-97
};

// ---- rt/triangle.go ----------------------------------------------------------------------------------------------------------
struct Triangle : Hittable {
    Point3 v0, v1, v2;
    Vec3 normal;
    const Material* mat;
    AABB bbox;
    int id = 0;
    static Triangle* New(Point3 v0, Point3 v1, Point3 v2, const Material* m) {  // :17-51
        auto t = new Triangle();
        Vec3 e1 = v1.Sub(v0), e2 = v2.Sub(v0);
        t->v0 = v0; t->v1 = v1; t->v2 = v2; t->mat = m;
        t->normal = Cross(e1, e2).Unit();
        Point3 mn{goMin(v0.X, goMin(v1.X, v2.X)), goMin(v0.Y, goMin(v1.Y, v2.Y)), goMin(v0.Z, goMin(v1.Z, v2.Z))};
        Point3 mx{goMax(v0.X, goMax(v1.X, v2.X)), goMax(v0.Y, goMax(v1.Y, v2.Y)), goMax(v0.Z, goMax(v1.Z, v2.Z))};
        t->bbox = AABBFromPoints(mn, mx);
        return t;
    }
    AABB BoundingBox() const override { return bbox; }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :57-104 Möller–Trumbore
        Vec3 edge1 = v1.Sub(v0), edge2 = v2.Sub(v0);
        Vec3 h = Cross(r.dir, edge2);
        double a = Dot(edge1, h);
        if (std::fabs(a) < 1e-8) return false;
        double f = 1.0 / a;
        Vec3 s = r.orig.Sub(v0);
        double uu = f * Dot(s, h);
        if (uu < 0.0 || uu > 1.0) return false;
        Vec3 q = Cross(s, edge1);
        double vv = f * Dot(r.dir, q);
        if (vv < 0.0 || uu + vv > 1.0) return false;
        double hitT = f * Dot(edge2, q);
        if (!rayT.Contains(hitT)) return false;
        rec->T = hitT; rec->P = r.At(hitT); rec->Mat = mat;
        rec->SetFaceNormal(r, normal);
        rec->U = uu; rec->V = vv;
        rec->prim = id;
        return true;
    }
};

// ---- rt/plane.go -------------------------------------------------------------------------------------------------------------
struct Plane : Hittable {
    Point3 Point;
    Vec3 Normal;
    const Material* Mat;
    int id = 0;
    AABB BoundingBox() const override { return AABBFromIntervals(UniverseInterval, UniverseInterval, UniverseInterval); }  // :17
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :24-42 (U,V not written)
        double denom = Dot(Normal, r.dir);
        if (std::fabs(denom) < 1e-8) return false;
        double t = Dot(Point.Sub(r.orig), Normal) / denom;
        if (!rayT.Surrounds(t)) return false;
        rec->T = t; rec->P = r.At(t);
        rec->SetFaceNormal(r, Normal);
        rec->Mat = Mat;
        rec->prim = id;
        return true;
    }
};

// ---- rt/circle.go ------------------------------------------------------------------------------------------------------------
struct Circle : Hittable {
    Point3 center;
    Vec3 normal;
    double radius, D;
    const Material* mat;
    AABB bbox;
    int id = 0;
    static Circle* New(Point3 center, Vec3 normal, double radius, const Material* m) {  // :14-31
        auto c = new Circle();
        c->normal = normal.Unit(); c->center = center; c->radius = radius; c->mat = m;
        c->D = Dot(c->normal, center);
        Vec3 rvec{radius, radius, radius};
        c->bbox = AABBFromPoints(center.Sub(rvec), center.Add(rvec));
        return c;
    }
    AABB BoundingBox() const override { return bbox; }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :36-74
        double denom = Dot(normal, r.dir);
        if (std::fabs(denom) < 1e-8) return false;
        double t = (D - Dot(normal, r.orig)) / denom;
        if (!rayT.Contains(t)) return false;
        Point3 intersection = r.At(t);
        double distanceFromCenter = intersection.Sub(center).Len();
        if (distanceFromCenter > radius) return false;
        rec->T = t; rec->P = intersection; rec->Mat = mat;
        rec->SetFaceNormal(r, normal);
        Vec3 u = std::fabs(normal.Y) > 0.9 ? Cross(Vec3{1, 0, 0}, normal).Unit() : Cross(Vec3{0, 1, 0}, normal).Unit();
        Vec3 v = Cross(normal, u);
        Vec3 localPoint = intersection.Sub(center);
        rec->U = (Dot(localPoint, u) / radius + 1.0) * 0.5;
        rec->V = (Dot(localPoint, v) / radius + 1.0) * 0.5;
        rec->prim = id;
        return true;
    }
};

// ---- rt/hittable_list.go -------------------------------------------------------------------------------------------------------
struct HittableList : Hittable {
    std::vector<const Hittable*> Objects;
    AABB bbox;
    bool tagItems = false;  // oracle bookkeeping: report the item index as prim id (Box sides)
    void Add(const Hittable* o) {  // :16-19
        Objects.push_back(o);
        bbox = AABBFromBoxes(bbox, o->BoundingBox());
    }
    AABB BoundingBox() const override { return bbox; }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :31-45
        HitRecord tempRec;
        bool hitAnything = false;
        double closestSoFar = rayT.Max;
        for (size_t k = 0; k < Objects.size(); k++) {
            if (Objects[k]->Hit(r, Interval{rayT.Min, closestSoFar}, &tempRec)) {
                hitAnything = true;
                closestSoFar = tempRec.T;
                if (tagItems) tempRec.prim = (int)k;
                *rec = tempRec;
            }
        }
        return hitAnything;
    }
};

// ---- rt/bvh.go -------------------------------------------------------------------------------------------------------------------
struct BVHLeaf : Hittable {
    std::vector<const Hittable*> objects;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :26-37
        bool hitAnything = false;
        double closest = rayT.Max;
        for (const Hittable* obj : objects) {
            if (obj->Hit(r, Interval{rayT.Min, closest}, rec)) {
                hitAnything = true;
                closest = rec->T;
            }
        }
        return hitAnything;
    }
};
struct BVHNode : Hittable {
    const Hittable *left = nullptr, *right = nullptr;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :219-239
        STAT_ADD(BVHIntersections);
        if (!left) return false;  // empty BVH (:73-75): Go would nil-deref; an empty world simply misses here
        if (!bbox.Hit(r, rayT)) return false;
        bool hitLeft = left->Hit(r, rayT, rec);
        double rightMax = rayT.Max;
        if (hitLeft) rightMax = rec->T;
        bool hitRight = right->Hit(r, Interval{rayT.Min, rightMax}, rec);  // a leaf is stored on both sides (:141)
        return hitLeft || hitRight;
    }
};
struct BvhPrim {
    int index;
    AABB bbox;
    Vec3 centroid;
};
// Go sorts with the unstable sort.Slice (:148); its order on equal keys cannot be reproduced without the Go
// runtime, so this oracle fixes a stable merge sort as canonical. Only exact ties in t can observe it.
template <class Less>
static void stableSort(std::vector<BvhPrim>& a, size_t lo, size_t hi, std::vector<BvhPrim>& tmp, Less less) {
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
static BVHNode* buildBVHNode(const std::vector<const Hittable*>& objects, std::vector<BvhPrim>& prims, size_t lo, size_t hi,
                             std::vector<BvhPrim>& tmp, size_t* nodeCount) {  // :120-217 (goroutine fan-out dropped: same tree)
    size_t n = hi - lo;
    AABB bounds = prims[lo].bbox;
    AABB cb = AABBFromPoints(prims[lo].centroid, prims[lo].centroid);
    for (size_t i = lo + 1; i < hi; i++) {
        bounds = AABBFromBoxes(bounds, prims[i].bbox);
        cb = AABBFromBoxes(cb, AABBFromPoints(prims[i].centroid, prims[i].centroid));
    }
    auto node = new BVHNode();
    node->bbox = bounds;
    if (nodeCount) (*nodeCount)++;
    if (n <= 4) {  // bvhLeafMaxSize :53
        auto leaf = new BVHLeaf();
        leaf->bbox = bounds;
        for (size_t i = lo; i < hi; i++) leaf->objects.push_back(objects[prims[i].index]);
        node->left = leaf;
        node->right = leaf;
        return node;
    }
    int axis = cb.LongestAxis();
    stableSort(prims, lo, hi, tmp, [axis](const BvhPrim& a, const BvhPrim& b) {
        return axis == 0 ? a.centroid.X < b.centroid.X : axis == 1 ? a.centroid.Y < b.centroid.Y : a.centroid.Z < b.centroid.Z;
    });
    size_t mid = lo + n / 2;
    node->left = buildBVHNode(objects, prims, lo, mid, tmp, nodeCount);
    node->right = buildBVHNode(objects, prims, mid, hi, tmp, nodeCount);
    return node;
}
static BVHNode* NewBVHNode(const std::vector<const Hittable*>& objects, size_t* nodeCount = nullptr) {  // :69-118
    size_t n = objects.size();
    if (n == 0) return new BVHNode();
    std::vector<BvhPrim> prims(n), tmp(n);
    for (size_t i = 0; i < n; i++) {
        AABB bb = objects[i]->BoundingBox();
        prims[i] = {(int)i, bb, bb.Centroid()};
    }
    return buildBVHNode(objects, prims, 0, n, tmp, nodeCount);
}

// ---- rt/transform.go -----------------------------------------------------------------------------------------------------------------
struct Translate : Hittable {
    const Hittable* Obj;
    Vec3 Offset;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :93-102
        Ray offsetRay{r.orig.Sub(Offset), r.dir, r.tm};
        if (!Obj->Hit(offsetRay, rayT, rec)) return false;
        rec->P = rec->P.Add(Offset);
        return true;
    }
};
struct RotateY : Hittable {
    const Hittable* Obj;
    double SinTheta, CosTheta;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
    void computeBox() {  // :125-156
        AABB bb = Obj->BoundingBox();
        Point3 mn{kInf, kInf, kInf}, mx{-kInf, -kInf, -kInf};
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 2; k++) {
                    double x = i * bb.X.Max + (1 - i) * bb.X.Min;
                    double y = j * bb.Y.Max + (1 - j) * bb.Y.Min;
                    double z = k * bb.Z.Max + (1 - k) * bb.Z.Min;
                    double nx = CosTheta * x + SinTheta * z;
                    double nz = -SinTheta * x + CosTheta * z;
                    mn.X = goMin(mn.X, nx); mx.X = goMax(mx.X, nx);
                    mn.Y = goMin(mn.Y, y);  mx.Y = goMax(mx.Y, y);
                    mn.Z = goMin(mn.Z, nz); mx.Z = goMax(mx.Z, nz);
                }
        bbox = AABBFromPoints(mn, mx);
    }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :159-187
        Point3 origin = r.orig;
        Vec3 direction = r.dir;
        origin.X = CosTheta * r.orig.X - SinTheta * r.orig.Z;
        origin.Z = SinTheta * r.orig.X + CosTheta * r.orig.Z;
        direction.X = CosTheta * r.dir.X - SinTheta * r.dir.Z;
        direction.Z = SinTheta * r.dir.X + CosTheta * r.dir.Z;
        Ray rotated{origin, direction, r.tm};
        if (!Obj->Hit(rotated, rayT, rec)) return false;
        Point3 p = rec->P;
        p.X = CosTheta * rec->P.X + SinTheta * rec->P.Z;
        p.Z = -SinTheta * rec->P.X + CosTheta * rec->P.Z;
        Vec3 normal = rec->Normal;
        normal.X = CosTheta * rec->Normal.X + SinTheta * rec->Normal.Z;
        normal.Z = -SinTheta * rec->Normal.X + CosTheta * rec->Normal.Z;
        rec->P = p;
        rec->Normal = normal;
        return true;
    }
};
struct Scale : Hittable {
    const Hittable* Obj;
    Vec3 Factor, InvFactor;
    AABB bbox;
    AABB BoundingBox() const override { return bbox; }
    void computeBox() {  // :374-401
        AABB bb = Obj->BoundingBox();
        Point3 mn{bb.X.Min * Factor.X, bb.Y.Min * Factor.Y, bb.Z.Min * Factor.Z}, mx{bb.X.Max * Factor.X, bb.Y.Max * Factor.Y, bb.Z.Max * Factor.Z};
        if (mn.X > mx.X) std::swap(mn.X, mx.X);
        if (mn.Y > mx.Y) std::swap(mn.Y, mx.Y);
        if (mn.Z > mx.Z) std::swap(mn.Z, mx.Z);
        bbox = AABBFromPoints(mn, mx);
    }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :408-440
        Point3 origin{r.orig.X * InvFactor.X, r.orig.Y * InvFactor.Y, r.orig.Z * InvFactor.Z};
        Vec3 direction{r.dir.X * InvFactor.X, r.dir.Y * InvFactor.Y, r.dir.Z * InvFactor.Z};
        Ray scaled{origin, direction, r.tm};
        if (!Obj->Hit(scaled, rayT, rec)) return false;
        rec->P = {rec->P.X * Factor.X, rec->P.Y * Factor.Y, rec->P.Z * Factor.Z};
        Vec3 normal{rec->Normal.X * InvFactor.X, rec->Normal.Y * InvFactor.Y, rec->Normal.Z * InvFactor.Z};
        rec->Normal = normal.Unit();
        return true;
    }
};

// ---- rt/volume.go ------------------------------------------------------------------------------------------------------------------------
struct Volume : Hittable {
    const Hittable* boundary;
    double negInvDensity;
    const Material* phaseFunction;
    AABB BoundingBox() const override { return boundary->BoundingBox(); }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {  // :34-79
        if (g_volumes_transparent) return false;
        HitRecord rec1, rec2;
        if (!boundary->Hit(r, UniverseInterval, &rec1)) return false;
        if (!boundary->Hit(r, Interval{rec1.T + 0.0001, kInf}, &rec2)) return false;
        if (rec1.T < rayT.Min) rec1.T = rayT.Min;
        if (rec2.T > rayT.Max) rec2.T = rayT.Max;
        if (rec1.T >= rec2.T) return false;
        if (rec1.T < 0) rec1.T = 0;
        double rayLength = r.dir.Len();
        double distanceInsideBoundary = (rec2.T - rec1.T) * rayLength;
        double hitDistance = negInvDensity * std::log(RandomDouble());
        if (hitDistance > distanceInsideBoundary) return false;
        rec->T = rec1.T + hitDistance / rayLength;
        rec->P = r.At(rec->T);
        rec->Normal = {1, 0, 0};
        rec->FrontFace = true;
        rec->Mat = phaseFunction;
        rec->prim = 0;
        return true;
    }
};

// Oracle bookkeeping wrapper: records which world entry was hit. Forwards Hit unchanged.
struct Tagged : Hittable {
    const Hittable* obj;
    int entry;
    bool singlePrim;
    AABB BoundingBox() const override { return obj->BoundingBox(); }
    bool Hit(const Ray& r, Interval rayT, HitRecord* rec) const override {
        if (!obj->Hit(r, rayT, rec)) return false;
        rec->entry = entry;
        if (singlePrim) rec->prim = 0;
        return true;
    }
};

// ---- rt/image_loader.go:97-120, :399-436 and rt/hdri.go ------------------------------------------------------------------------------------
static inline int clampi(int x, int low, int high) {  // :112-120 (returns high-1 at or above high)
    if (x < low) return low;
    if (x < high) return x;
    return high - 1;
}
struct HDRIEnvironment {
    int width = 0, height = 0;
    std::vector<Color> data;
    double rotation = 0;
    bool useImportanceSampling = true;
    std::vector<double> pdf, marginalCDF;
    std::vector<std::vector<double>> conditionalCDFs;
    double totalPower = 0;
    bool IsValid() const { return !data.empty(); }
    Color PixelData(int x, int y) const {  // rt/image_loader.go:97-109
        x = clampi(x, 0, width);
        y = clampi(y, 0, height);
        return data[(size_t)y * width + x];
    }
    Color PixelDataBilinear(double u, double v) const {  // rt/image_loader.go:399-436
        double px = u * (double)width - 0.5, py = v * (double)height - 0.5;
        int x0 = (int)std::floor(px), y0 = (int)std::floor(py);
        int x1 = x0 + 1, y1 = y0 + 1;
        double fx = px - (double)x0, fy = py - (double)y0;
        x0 = ((x0 % width) + width) % width;
        x1 = ((x1 % width) + width) % width;
        y0 = clampi(y0, 0, height);
        y1 = clampi(y1, 0, height);
        Color c00 = PixelData(x0, y0), c10 = PixelData(x1, y0), c01 = PixelData(x0, y1), c11 = PixelData(x1, y1);
        Color c0 = c00.Scale(1 - fx).Add(c10.Scale(fx));
        Color c1 = c01.Scale(1 - fx).Add(c11.Scale(fx));
        return c0.Scale(1 - fy).Add(c1.Scale(fy));
    }
    void DirectionToUV(const Vec3& dir, double& u, double& v) const {  // rt/hdri.go:75-94
        Vec3 d = dir.Unit();
        double phi = std::atan2(d.Z, d.X);
        double theta = std::asin(d.Y);
        u = 0.5 + phi / (2 * M_PI);
        v = 0.5 - theta / M_PI;
        u = u + rotation / (2 * M_PI);
        u = u - std::floor(u);
    }
    Vec3 UVToDirection(double u, double v) const {  // rt/hdri.go:97-113
        u = u - rotation / (2 * M_PI);
        u = u - std::floor(u);
        double phi = (u - 0.5) * 2 * M_PI;
        double theta = (0.5 - v) * M_PI;
        double cosTheta = std::cos(theta);
        return {cosTheta * std::cos(phi), std::sin(theta), cosTheta * std::sin(phi)};
    }
    Color Sample(const Vec3& dir) const {  // rt/hdri.go:120-128
        double u, v;
        DirectionToUV(dir, u, v);
        return PixelDataBilinear(u, v);
    }
    void BuildDistribution() {  // rt/hdri.go:145-224
        if (!IsValid()) return;
        int total = width * height;
        pdf.assign(total, 0.0);
        marginalCDF.assign(height + 1, 0.0);
        conditionalCDFs.assign(height, {});
        totalPower = 0;
        std::vector<double> rowSums(height, 0.0);
        for (int y = 0; y < height; y++) {
            double v = ((double)y + 0.5) / (double)height;
            double theta = (0.5 - v) * M_PI;
            double sinTheta = std::cos(theta);
            conditionalCDFs[y].assign(width + 1, 0.0);
            for (int x = 0; x < width; x++) {
                int idx = y * width + x;
                Color c = PixelData(x, y);
                double luminance = 0.2126 * c.X + 0.7152 * c.Y + 0.0722 * c.Z;
                double weight = luminance * sinTheta;
                if (weight < 0) weight = 0;
                pdf[idx] = weight;
                rowSums[y] += weight;
                totalPower += weight;
                conditionalCDFs[y][x + 1] = conditionalCDFs[y][x] + weight;
            }
        }
        for (int y = 0; y < height; y++)
            if (rowSums[y] > 0)
                for (int x = 0; x <= width; x++) conditionalCDFs[y][x] /= rowSums[y];
        marginalCDF[0] = 0;
        for (int y = 0; y < height; y++) marginalCDF[y + 1] = marginalCDF[y] + rowSums[y];
        if (totalPower > 0) {
            for (int y = 0; y <= height; y++) marginalCDF[y] /= totalPower;
            for (auto& p : pdf) p /= totalPower;
        }
    }
    static int searchCDF(const std::vector<double>& cdf, double xi) {  // rt/hdri.go:300-322
        int n = (int)cdf.size() - 1;
        int low = 0, high = n;
        while (low < high) {
            int mid = (low + high) / 2;
            if (cdf[mid + 1] <= xi) low = mid + 1;
            else high = mid;
        }
        if (low >= n) low = n - 1;
        if (low < 0) low = 0;
        return low;
    }
    double PDF(const Vec3& dir) const {  // rt/hdri.go:262-297
        if (!IsValid() || !useImportanceSampling || totalPower == 0) return 1.0 / (4.0 * M_PI);
        double u, v;
        DirectionToUV(dir, u, v);
        int x = (int)(u * (double)width), y = (int)(v * (double)height);
        x = clampi(x, 0, width);
        y = clampi(y, 0, height);
        int idx = y * width + x;
        double theta = (0.5 - v) * M_PI;
        double sinTheta = std::cos(theta);
        if (sinTheta < 1e-10) sinTheta = 1e-10;
        double pdfSolidAngle = pdf[idx] * (double)(width * height) / (2.0 * M_PI * M_PI * sinTheta);
        if (pdfSolidAngle < 1e-10) return 1e-10;
        return pdfSolidAngle;
    }
    // SampleDirection with the two uniforms made explicit (rt/hdri.go:228-259 draws xi1 then xi2).
    void SampleDirectionXi(double xi1, double xi2, Vec3& dir, Color& emission, double& p) const {
        if (!IsValid() || !useImportanceSampling || totalPower == 0) {
            dir = RandomUnitVector();
            emission = Sample(dir);
            p = 1.0 / (4.0 * M_PI);
            return;
        }
        int y = searchCDF(marginalCDF, xi1);
        int x = searchCDF(conditionalCDFs[y], xi2);
        double u = ((double)x + 0.5) / (double)width;
        double v = ((double)y + 0.5) / (double)height;
        dir = UVToDirection(u, v);
        emission = PixelData(x, y);
        p = PDF(dir);
    }
    void SampleDirection(Vec3& dir, Color& emission, double& p) const {
        double xi1 = RandomDouble();
        double xi2 = RandomDouble();
        SampleDirectionXi(xi1, xi2, dir, emission, p);
    }
};

// rt/image_loader.go:165-383 — Radiance .hdr decoder (new-RLE and flat scanlines), (m+0.5)*2^(e-136).
static bool LoadHDR(const std::string& path, int& width, int& height, std::vector<Color>& out, std::string& err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "could not open"; return false; }
    std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    size_t pos = 0;
    auto readLine = [&](std::string& o) {
        size_t nl = data.find('\n', pos);
        if (nl == std::string::npos) return false;
        o = data.substr(pos, nl - pos);
        pos = nl + 1;
        return true;
    };
    auto trim = [](const std::string& s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    std::string line;
    if (!readLine(line) || line.rfind("#?", 0) != 0) { err = "missing #? signature"; return false; }  // :211-213
    for (;;) {
        if (!readLine(line)) { err = "unexpected end of header"; return false; }
        if (trim(line).empty()) break;  // :224-227
    }
    if (!readLine(line)) { err = "no resolution"; return false; }
    std::istringstream rs(trim(line));
    std::string a, b, c, d;
    rs >> a >> b >> c >> d;
    if (a == "-Y" && c == "+X") { height = atoi(b.c_str()); width = atoi(d.c_str()); }       // :247-256
    else if (a == "+X" && c == "-Y") { width = atoi(b.c_str()); height = atoi(d.c_str()); }  // :257-267
    else { err = "unsupported resolution format"; return false; }
    out.assign((size_t)width * height, Color{0, 0, 0});
    const unsigned char* p = (const unsigned char*)data.data();
    size_t n = data.size();
    auto put = [&](int y, int x, const unsigned char* q) {  // rgbeToColor :364-383
        if (q[3] == 0) return;
        double scale = std::ldexp(1.0, (int)q[3] - 128 - 8);
        out[(size_t)y * width + x] = {((double)q[0] + 0.5) * scale, ((double)q[1] + 0.5) * scale, ((double)q[2] + 0.5) * scale};
    };
    std::vector<unsigned char> scan((size_t)4 * width);
    for (int y = 0; y < height; y++) {
        if (pos + 4 > n) { err = "scanline header"; return false; }
        const unsigned char* h = p + pos;
        pos += 4;
        if (h[0] == 2 && h[1] == 2) {  // :285-292
            if (((h[2] << 8) | h[3]) != width) { err = "scanline width mismatch"; return false; }
            for (int comp = 0; comp < 4; comp++) {  // :320-352
                int x = 0;
                while (x < width) {
                    if (pos >= n) { err = "rle"; return false; }
                    int code = p[pos++];
                    if (code > 128) {
                        int count = code - 128;
                        if (pos >= n) { err = "rle"; return false; }
                        unsigned char v = p[pos++];
                        for (int i = 0; i < count && x < width; i++) scan[(size_t)comp * width + x++] = v;
                    } else {
                        for (int i = 0; i < code && x < width; i++) {
                            if (pos >= n) { err = "rle"; return false; }
                            scan[(size_t)comp * width + x++] = p[pos++];
                        }
                    }
                }
            }
            for (int x = 0; x < width; x++) {
                unsigned char q[4] = {scan[x], scan[(size_t)width + x], scan[(size_t)2 * width + x], scan[(size_t)3 * width + x]};
                put(y, x, q);
            }
        } else {  // :294-308
            put(y, 0, h);
            for (int x = 1; x < width; x++) {
                if (pos + 4 > n) { err = "pixel"; return false; }
                put(y, x, p + pos);
                pos += 4;
            }
        }
    }
    return true;
}

// ---- rt/camera.go -----------------------------------------------------------------------------------------------------------------------------------
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
    std::vector<const Hittable*> Lights;
    const HDRIEnvironment* Environment = nullptr;

    double pixelsSamplesScale = 0;
    Point3 center, pixel00Loc;
    Vec3 pixelDeltaU, pixelDeltaV, u, v, w, defocusDiskU, defocusDiskV;
    Ray centerMotion, lookAtMotion;
    double viewportHeight = 0, viewportWidth = 0;

    void Initialize() {  // :286-344
        if (CameraMotion) {
            centerMotion = Ray{LookFrom, LookFrom2.Sub(LookFrom), 0};
            lookAtMotion = Ray{LookAt, LookAt2.Sub(LookAt), 0};
        } else {
            centerMotion = Ray{LookFrom, {0, 0, 0}, 0};
            lookAtMotion = Ray{LookAt, {0, 0, 0}, 0};
        }
        ImageHeight = std::max((int)((double)ImageWidth / AspectRatio), 1);
        pixelsSamplesScale = 1.0 / (double)SamplesPerPixel;
        center = LookFrom;
        double theta = DegreesToRadians(Vfov);
        double h = std::tan(theta / 2);
        viewportHeight = 2 * h * FocusDist;
        viewportWidth = viewportHeight * ((double)ImageWidth / (double)ImageHeight);
        if (FreeCamera) w = Forward.Neg();
        else w = center.Sub(LookAt).Unit();
        u = Cross(Vup, w).Unit();
        v = Cross(w, u);
        Vec3 viewportU = u.Scale(viewportWidth);
        Vec3 viewportV = v.Neg().Scale(viewportHeight);
        pixelDeltaU = viewportU.Div((double)ImageWidth);
        pixelDeltaV = viewportV.Div((double)ImageHeight);
        Point3 upperLeft = center.Sub(w.Scale(FocusDist)).Sub(viewportU.Div(2)).Sub(viewportV.Div(2));
        pixel00Loc = upperLeft.Add(pixelDeltaU.Add(pixelDeltaV).Scale(0.5));
        double defocusRadius = FocusDist * std::tan(DegreesToRadians(DefocusAngle / 2));
        defocusDiskU = u.Scale(defocusRadius);
        defocusDiskV = v.Scale(defocusRadius);
    }
    Point3 defocusDiskAt(const Point3& c, const Vec3& uu, const Vec3& vv, const Vec3& p) const {  // :354-362 with p given
        double defocusRadius = FocusDist * std::tan(DegreesToRadians(DefocusAngle / 2));
        Vec3 dU = uu.Scale(defocusRadius), dV = vv.Scale(defocusRadius);
        return c.Add(dU.Scale(p.X)).Add(dV.Scale(p.Y));
    }
    // GetRay (:368-435) with its random inputs made explicit: offset = sampleSquare(), rayTime, p = unit-disk point.
    Ray GetRayExplicit(int i, int j, const Vec3& offset, double rayTime, const Vec3& diskP) const {
        if (!CameraMotion && !FreeCamera) {
            Point3 pixelSample = pixel00Loc.Add(pixelDeltaU.Scale((double)i + offset.X)).Add(pixelDeltaV.Scale((double)j + offset.Y));
            Point3 rayOrigin = DefocusAngle <= 0 ? center : defocusDiskAt(center, u, v, diskP);
            return Ray{rayOrigin, pixelSample.Sub(rayOrigin), rayTime};
        }
        Point3 currentCenter = centerMotion.At(rayTime);
        Vec3 uu, vv, ww;
        if (FreeCamera) {
            ww = Forward.Neg();
        } else {
            Point3 currentLookAt = lookAtMotion.At(rayTime);
            ww = currentCenter.Sub(currentLookAt).Unit();
        }
        uu = Cross(Vup, ww).Unit();
        vv = Cross(ww, uu);
        Vec3 viewportU = uu.Scale(viewportWidth), viewportV = vv.Neg().Scale(viewportHeight);
        Vec3 dU = viewportU.Div((double)ImageWidth), dV = viewportV.Div((double)ImageHeight);
        Point3 upperLeft = currentCenter.Sub(ww.Scale(FocusDist)).Sub(viewportU.Div(2)).Sub(viewportV.Div(2));
        Point3 p00 = upperLeft.Add(dU.Add(dV).Scale(0.5));
        Point3 pixelSample = p00.Add(dU.Scale((double)i + offset.X)).Add(dV.Scale((double)j + offset.Y));
        Point3 rayOrigin = DefocusAngle <= 0 ? currentCenter : defocusDiskAt(currentCenter, uu, vv, diskP);
        return Ray{rayOrigin, pixelSample.Sub(rayOrigin), rayTime};
    }
    Ray GetRay(int i, int j) const {
        double ox = RandomDouble() - 0.5, oy = RandomDouble() - 0.5;  // sampleSquare :346-352
        double rayTime = RandomDouble();                               // :370
        Vec3 p{0, 0, 0};
        if (DefocusAngle > 0) {
            p = RandomInUnitDisk();  // :355 — drawn and discarded
            p = RandomInUnitDisk();  // :359
        }
        return GetRayExplicit(i, j, {ox, oy, 0}, rayTime, p);
    }
    Color SkyGradient(const Ray& r) const {  // :520-526
        Vec3 ud = r.dir.Unit();
        double a = 0.5 * (ud.Y + 1.0);
        return Color{1.0, 1.0, 1.0}.Scale(1.0 - a).Add(Color{0.5, 0.7, 1.0}.Scale(a));
    }
    Color RayColor(const Ray& r, int depth, const Hittable* world) const {  // :438-441
        STAT_ADD(RayCount);
        return rayColorInternal(r, depth, world, true);
    }
    Color rayColorInternal(const Ray& r, int depth, const Hittable* world, bool allowLightHits) const {  // :443-518
        if (depth <= 0) return {0, 0, 0};
        STAT_ADD(RayCount);
        std::unique_ptr<HitRecord> rec(new HitRecord());  // heap record per call, as :449
        if (!world->Hit(r, Interval{0.001, kInf}, rec.get())) {
            if (Environment && Environment->IsValid()) {
                bool isPrimaryRay = (depth == MaxDepth);
                if (PhantomHDRI && isPrimaryRay) return {0, 0, 0};
                return Environment->Sample(r.dir);
            }
            if (UseSkyGradient) return SkyGradient(r);
            return Background;
        }
        Color attenuation;
        Ray scattered;
        Color colorFromEmission = rec->Mat->Emitted(rec->U, rec->V, rec->P);
        if (!rec->Mat->Scatter(r, rec.get(), &attenuation, &scattered)) {
            if (allowLightHits) return colorFromEmission;
            return {0, 0, 0};
        }
        bool useMIS = rec->Mat->CanUseNEE() && !Lights.empty();  // every shipped material implements both interfaces (:484-489)
        if (!useMIS) {
            Color colorFromScatter = attenuation.Mult(rayColorInternal(scattered, depth - 1, world, true));
            return colorFromEmission.Add(colorFromScatter);
        }
        int lightIdx = (int)(RandomDouble() * (double)Lights.size());
        if (lightIdx >= (int)Lights.size()) lightIdx = (int)Lights.size() - 1;
        Color directLight = sampleLightMIS(rec->P, rec->Normal, r.dir, world, lightIdx, attenuation, rec->Mat);
        Color indirectLight = attenuation.Mult(rayColorInternal(scattered, depth - 1, world, false));
        return colorFromEmission.Add(directLight).Add(indirectLight);
    }
    Color sampleLightMIS(const Point3& hitPoint, const Vec3& hitNormal, const Vec3& rayDirection, const Hittable* world, int lightIdx,
                         const Color& attenuation, const Material* pdfEval) const {  // :538-562
        Color total{0, 0, 0};
        if (Environment && Environment->IsValid() && Environment->useImportanceSampling)
            total = total.Add(sampleHDRILight(hitPoint, hitNormal, rayDirection, world, attenuation, pdfEval));
        if (!Lights.empty() && lightIdx < (int)Lights.size())
            total = total.Add(sampleAreaLight(hitPoint, hitNormal, rayDirection, world, lightIdx, attenuation, pdfEval));
        return total;
    }
    Color sampleHDRILight(const Point3& hitPoint, const Vec3& hitNormal, const Vec3& rayDirection, const Hittable* world,
                          const Color& attenuation, const Material* pdfEval) const {  // :565-607
        Vec3 lightDir;
        Color emission;
        double pdfHDRI;
        Environment->SampleDirection(lightDir, emission, pdfHDRI);
        double cosTheta = Dot(hitNormal, lightDir);
        if (cosTheta <= 0) return {0, 0, 0};
        Ray shadowRay{hitPoint, lightDir, 0};
        std::unique_ptr<HitRecord> shadowRec(new HitRecord());
        STAT_ADD(ShadowQueries);
        if (world->Hit(shadowRay, Interval{0.001, kInf}, shadowRec.get())) return {0, 0, 0};
        Vec3 wi = rayDirection.Neg().Unit();
        double pdfBRDF = pdfEval->PDF(wi, lightDir, hitNormal);
        double weight = pdfHDRI / (pdfHDRI + pdfBRDF);
        Color contribution = emission.Scale(cosTheta / pdfHDRI * weight);
        contribution = contribution.Mult(attenuation);
        const double maxComponent = 20.0;
        contribution.X = std::fmin(contribution.X, maxComponent);
        contribution.Y = std::fmin(contribution.Y, maxComponent);
        contribution.Z = std::fmin(contribution.Z, maxComponent);
        return contribution;
    }
    Color sampleAreaLight(const Point3& hitPoint, const Vec3& hitNormal, const Vec3& rayDirection, const Hittable* world, int lightIdx,
                          const Color& attenuation, const Material* pdfEval) const {  // :610-678
        const Quad* lightQuad = dynamic_cast<const Quad*>(Lights[lightIdx]);
        if (!lightQuad) return {0, 0, 0};
        Point3 lightPoint = lightQuad->SamplePoint();
        Vec3 toLight = lightPoint.Sub(hitPoint);
        double distanceToLight = toLight.Len();
        Vec3 lightDir = toLight.Unit();
        double cosTheta = Dot(hitNormal, lightDir);
        if (cosTheta <= 0) return {0, 0, 0};
        Ray shadowRay{hitPoint, lightDir, 0};
        std::unique_ptr<HitRecord> shadowRec(new HitRecord());
        STAT_ADD(ShadowQueries);
        if (world->Hit(shadowRay, Interval{0.001, distanceToLight - 0.001}, shadowRec.get())) return {0, 0, 0};
        Color emission = lightQuad->mat->Emitted(0, 0, lightPoint);
        double lightArea = lightQuad->Area();
        double cosLightAngle = std::fabs(Dot(lightQuad->normal, lightDir.Neg()));
        if (cosLightAngle < 0.001) return {0, 0, 0};
        double pdfLight = (distanceToLight * distanceToLight) / (cosLightAngle * lightArea);
        Vec3 wi = rayDirection.Neg().Unit();
        double pdfBRDF = pdfEval->PDF(wi, lightDir, hitNormal);
        double weight = pdfLight / (pdfLight + pdfBRDF);
        Color contribution = emission.Scale(cosTheta / pdfLight * weight);
        contribution = contribution.Mult(attenuation).Scale((double)Lights.size());
        const double maxComponent = 20.0;
        contribution.X = std::fmin(contribution.X, maxComponent);
        contribution.Y = std::fmin(contribution.Y, maxComponent);
        contribution.Z = std::fmin(contribution.Z, maxComponent);
        return contribution;
    }
};

// ---- a scene rebuilt from the flat description: the same pointer graph the Go constructors would create ----------------------------------
struct Scene {
    std::vector<std::unique_ptr<Texture>> textures;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<Hittable>> owned;
    std::vector<std::unique_ptr<Perlin>> perlins;
    std::list<std::vector<double>> imageData;
    std::vector<Quad*> quads;
    HittableList* worldList = nullptr;
    const Hittable* world = nullptr;
    HDRIEnvironment env;
    Camera cam;
    size_t bvhNodes = 0;
    template <class T>
    T* keep(T* p) { owned.emplace_back(p); return p; }
};

static Vec3 v3(const double* p, int i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

static Scene* sceneFromDesc(const rtx_scene_desc* d, const rtx_camera_desc* c) {
    auto S = new Scene();
    S->textures.resize(d->n_textures);
    for (int i = 0; i < d->n_textures; i++) {
        if (d->tex_type[i] == RTX_TEX_SOLID) { auto t = new SolidColor(); t->Albedo = v3(d->tex_color, i); S->textures[i].reset(t); }
        else if (d->tex_type[i] == RTX_TEX_NOISE) {
            S->perlins.emplace_back(new Perlin());
            Perlin* pn = S->perlins.back().get();
            const double* pv = d->perlin_vec + (size_t)768 * d->tex_even[i];
            const int32_t* pp = d->perlin_perm + (size_t)768 * d->tex_even[i];
            for (int k = 0; k < 256; k++) { pn->randvec[k] = {pv[3 * k], pv[3 * k + 1], pv[3 * k + 2]}; pn->permX[k] = pp[k]; pn->permY[k] = pp[256 + k]; pn->permZ[k] = pp[512 + k]; }
            auto t = new NoiseTexture();
            t->noise = pn; t->scale = d->tex_inv_scale[i];
            S->textures[i].reset(t);
        } else if (d->tex_type[i] == RTX_TEX_IMAGE) {
            auto t = new ImageTexture();
            const int im = d->tex_even[i];
            t->width = d->image_width[im]; t->height = d->image_height[im];
            const double* src = d->image_rgb + 3 * (size_t)d->image_offset[im];
            S->imageData.emplace_back(src, src + 3 * (size_t)t->width * t->height);   // the descriptor is only borrowed
            t->data = S->imageData.back().data();
            S->textures[i].reset(t);
        } else S->textures[i].reset(new CheckerTexture());
    }
    for (int i = 0; i < d->n_textures; i++)
        if (d->tex_type[i] == RTX_TEX_CHECKER) {
            auto t = static_cast<CheckerTexture*>(S->textures[i].get());
            t->invScale = d->tex_inv_scale[i]; t->even = S->textures[d->tex_even[i]].get(); t->odd = S->textures[d->tex_odd[i]].get();
        }
    for (int i = 0; i < d->n_materials; i++) {
        Material* m = nullptr;
        const Texture* tex = d->mat_tex[i] >= 0 ? S->textures[d->mat_tex[i]].get() : nullptr;
        switch (d->mat_type[i]) {
            case RTX_MAT_LAMBERTIAN: { auto x = new Lambertian(); x->tex = tex; m = x; break; }
            case RTX_MAT_METAL: { auto x = new Metal(); x->Albedo = v3(d->mat_albedo, i); x->Fuzz = d->mat_fuzz[i]; m = x; break; }
            case RTX_MAT_DIELECTRIC: { auto x = new Dielectric(); x->RefractionIndex = d->mat_ior[i]; m = x; break; }
            case RTX_MAT_DIFFUSE_LIGHT: { auto x = new DiffuseLight(); x->tex = tex; m = x; break; }
            default: { auto x = new Isotropic(); x->tex = tex; m = x; break; }
        }
        S->materials.emplace_back(m);
    }
    auto mat = [&](int id) { return S->materials[id].get(); };
    std::vector<Sphere*> spheres(d->n_spheres);
    for (int i = 0; i < d->n_spheres; i++) spheres[i] = S->keep(Sphere::New(v3(d->sph_center, i), v3(d->sph_velocity, i), d->sph_radius[i], mat(d->sph_mat[i])));
    S->quads.resize(d->n_quads);
    for (int i = 0; i < d->n_quads; i++) S->quads[i] = S->keep(Quad::New(v3(d->quad_q, i), v3(d->quad_u, i), v3(d->quad_v, i), mat(d->quad_mat[i])));
    std::vector<Triangle*> tris(d->n_tris);
    for (int i = 0; i < d->n_tris; i++) tris[i] = S->keep(Triangle::New(v3(d->tri_v0, i), v3(d->tri_v1, i), v3(d->tri_v2, i), mat(d->tri_mat[i])));
    std::vector<Plane*> planes(d->n_planes);
    for (int i = 0; i < d->n_planes; i++) {
        auto p = S->keep(new Plane());
        p->Point = v3(d->plane_point, i); p->Normal = v3(d->plane_normal, i); p->Mat = mat(d->plane_mat[i]);
        planes[i] = p;
    }
    std::vector<Circle*> circles(d->n_circles);
    for (int i = 0; i < d->n_circles; i++) circles[i] = S->keep(Circle::New(v3(d->circle_center, i), v3(d->circle_normal, i), d->circle_radius[i], mat(d->circle_mat[i])));
    auto isPrim = [](int kind) { return kind <= RTX_GEOM_PLANE || kind == RTX_GEOM_CIRCLE; };
    auto prim = [&](int kind, int idx) -> const Hittable* {
        switch (kind) {
            case RTX_GEOM_CIRCLE: return circles[idx];
            case RTX_GEOM_SPHERE: return spheres[idx];
            case RTX_GEOM_QUAD: return S->quads[idx];
            case RTX_GEOM_TRIANGLE: return tris[idx];
            default: return planes[idx];
        }
    };
    std::vector<const Hittable*> groups(d->n_groups);
    for (int g = 0; g < d->n_groups; g++) {
        if (d->group_kind[g] == RTX_GEOM_LIST) {
            auto l = S->keep(new HittableList());
            l->tagItems = true;
            for (int k = 0; k < d->group_count[g]; k++) {
                int it = d->group_begin[g] + k;
                l->Add(prim(d->list_item_kind[it], d->list_item_index[it]));
            }
            groups[g] = l;
        } else {
            std::vector<const Hittable*> objs(d->group_count[g]);
            for (int k = 0; k < d->group_count[g]; k++) {
                tris[d->group_begin[g] + k]->id = k;
                objs[k] = tris[d->group_begin[g] + k];
            }
            groups[g] = S->keep(NewBVHNode(objs, &S->bvhNodes));  // rt/obj_loader.go:109 (leaks inner nodes; test infra)
        }
    }
    S->worldList = S->keep(new HittableList());
    for (int e = 0; e < d->n_entries; e++) {
        int kind = d->entry_geom_kind[e];
        const Hittable* h = isPrim(kind) ? prim(kind, d->entry_geom_index[e]) : groups[d->entry_geom_index[e]];
        for (int k = d->entry_xf_count[e] - 1; k >= 0; k--) {  // innermost first
            int x = d->entry_xf_begin[e] + k;
            if (d->xf_type[x] == RTX_XF_TRANSLATE) {
                auto t = S->keep(new Translate());
                t->Obj = h; t->Offset = v3(d->xf_a, x); t->bbox = h->BoundingBox().Translate(t->Offset);
                h = t;
            } else if (d->xf_type[x] == RTX_XF_ROTATE_Y) {
                auto r = S->keep(new RotateY());
                r->Obj = h; r->SinTheta = d->xf_a[3 * x]; r->CosTheta = d->xf_a[3 * x + 1]; r->computeBox();
                h = r;
            } else {
                auto s = S->keep(new Scale());
                s->Obj = h; s->Factor = v3(d->xf_a, x); s->InvFactor = v3(d->xf_b, x); s->computeBox();
                h = s;
            }
        }
        if (d->entry_volume[e] >= 0) {
            auto v = S->keep(new Volume());
            v->boundary = h; v->negInvDensity = d->vol_neg_inv_density[d->entry_volume[e]]; v->phaseFunction = mat(d->vol_mat[d->entry_volume[e]]);
            h = v;
        }
        auto tag = S->keep(new Tagged());
        tag->obj = h; tag->entry = e; tag->singlePrim = isPrim(kind) || d->entry_volume[e] >= 0;
        S->worldList->Add(tag);
    }
    S->world = d->world_is_bvh ? (const Hittable*)S->keep(NewBVHNode(S->worldList->Objects, &S->bvhNodes)) : S->worldList;  // main.go:77
    if (d->env_width > 0 && d->env_rgb) {
        S->env.width = d->env_width; S->env.height = d->env_height;
        S->env.data.resize((size_t)d->env_width * d->env_height);
        for (size_t i = 0; i < S->env.data.size(); i++) S->env.data[i] = {d->env_rgb[3 * i], d->env_rgb[3 * i + 1], d->env_rgb[3 * i + 2]};
        S->env.rotation = d->env_rotation;
        S->env.BuildDistribution();                                     // NewHDRIEnvironment, rt/hdri.go:29-48
        S->env.useImportanceSampling = d->env_importance_sampling != 0;  // DisableImportanceSampling :56
    }
    Camera& cam = S->cam;
    if (c) {
        cam.AspectRatio = c->aspect_ratio; cam.ImageWidth = c->image_width; cam.SamplesPerPixel = c->samples_per_pixel; cam.MaxDepth = c->max_depth;
        cam.Vfov = c->vfov;
        cam.LookFrom = {c->look_from[0], c->look_from[1], c->look_from[2]};
        cam.LookAt = {c->look_at[0], c->look_at[1], c->look_at[2]};
        cam.Vup = {c->vup[0], c->vup[1], c->vup[2]};
        cam.DefocusAngle = c->defocus_angle; cam.FocusDist = c->focus_dist;
        cam.LookFrom2 = {c->look_from2[0], c->look_from2[1], c->look_from2[2]};
        cam.LookAt2 = {c->look_at2[0], c->look_at2[1], c->look_at2[2]};
        cam.CameraMotion = c->camera_motion; cam.FreeCamera = c->free_camera;
        cam.Forward = {c->forward[0], c->forward[1], c->forward[2]};
        cam.Background = {c->background[0], c->background[1], c->background[2]};
        cam.UseSkyGradient = c->use_sky_gradient; cam.PhantomHDRI = c->phantom_hdri;
    }
    for (int i = 0; i < d->n_lights; i++) cam.Lights.push_back(d->light_quad[i] >= 0 ? (const Hittable*)S->quads[d->light_quad[i]] : (const Hittable*)S->worldList);
    if (S->env.IsValid()) cam.Environment = &S->env;
    cam.Initialize();
    return S;
}

// ---- rt/bucket_renderer.go -----------------------------------------------------------------------------------------------------------------------------
struct Bucket { int X, Y, Width, Height; };
static std::vector<Bucket> generateBuckets(int width, int height, int bucketSize) {  // :77-125
    std::vector<Bucket> buckets;
    for (int y = 0; y < height; y += bucketSize)
        for (int x = 0; x < width; x += bucketSize) buckets.push_back({x, y, std::min(bucketSize, width - x), std::min(bucketSize, height - y)});
    int cx = width / 2, cy = height / 2;
    auto dist = [&](const Bucket& b) {
        double dx = (double)(b.X + b.Width / 2 - cx), dy = (double)(b.Y + b.Height / 2 - cy);
        return dx * dx + dy * dy;
    };
    std::stable_sort(buckets.begin(), buckets.end(), [&](const Bucket& a, const Bucket& b) { return dist(a) < dist(b); });
    return buckets;
}
// renderPass + workerMultiPass + renderBucketWithQuality (:170-301). Instead of the 8-bit tile buffer the per-pixel
// sample sum and sum of squares (linear, before scale/gamma) are returned; orc_resolve_rgba8 finishes :275-285.
static void renderPass(const Scene* S, int samplesPerPixel, int maxDepth, int numWorkers, int bucketSize, uint64_t seed, double* sum, double* sumsq) {
    const Camera& cam = S->cam;
    int W = cam.ImageWidth, H = cam.ImageHeight;
    std::vector<Bucket> buckets = generateBuckets(W, H, bucketSize);
    std::atomic<size_t> next{0};  // the buffered channel of :194 as a shared cursor
    auto worker = [&](int workerID) {
        t_rng.reseed(seed * 0x9E3779B97F4A7C15ull + (uint64_t)workerID + 1);
        for (;;) {
            size_t bi = next.fetch_add(1);
            if (bi >= buckets.size()) break;
            const Bucket& b = buckets[bi];
            for (int ly = 0; ly < b.Height; ly++)
                for (int lx = 0; lx < b.Width; lx++) {
                    int gx = b.X + lx, gy = b.Y + ly;
                    Color pixelColor{0, 0, 0}, sq{0, 0, 0};
                    for (int s = 0; s < samplesPerPixel; s++) {
                        Ray ray = cam.GetRay(gx, gy);
                        Color c = cam.RayColor(ray, maxDepth, S->world);
                        pixelColor = pixelColor.Add(c);
                        sq = sq.Add(c.Mult(c));
                        STAT_ADD(SamplesComputed);
                    }
                    size_t o = ((size_t)gy * W + gx) * 3;
                    sum[o] = pixelColor.X; sum[o + 1] = pixelColor.Y; sum[o + 2] = pixelColor.Z;
                    if (sumsq) { sumsq[o] = sq.X; sumsq[o + 1] = sq.Y; sumsq[o + 2] = sq.Z; }
                    STAT_ADD(PixelsRendered);
                }
        }
    };
    std::vector<std::thread> threads;
    for (int i = 0; i < numWorkers; i++) threads.emplace_back(worker, i);
    for (auto& t : threads) t.join();
}

}  // namespace orc

// ==========================================================================================================
// C surface for ctypes
// ==========================================================================================================
using namespace orc;
extern "C" {

struct orc_scene { Scene* S; };

orc_scene* orc_scene_from_desc(const rtx_scene_desc* d, const rtx_camera_desc* c) {
    auto h = new orc_scene();
    h->S = sceneFromDesc(d, c);
    return h;
}
void orc_scene_free(orc_scene* h) {
    if (h) { delete h->S; delete h; }
}
void orc_image_size(orc_scene* h, int32_t* w, int32_t* hh) { *w = h->S->cam.ImageWidth; *hh = h->S->cam.ImageHeight; }
int64_t orc_bvh_nodes(orc_scene* h) { return (int64_t)h->S->bvhNodes; }

// world.Hit(r, Interval{tmin,tmax}, rec) per ray; volumes transparent (level-1 protocol).
void orc_trace_closest(orc_scene* h, const double* rays, int64_t n, double tmin, double tmax, int32_t* entry_id, int32_t* prim_id,
                       double* t, double* normal, uint8_t* front, double* uv, double* p) {
    bool savedT = g_volumes_transparent, savedA = g_use_atomics;
    g_volumes_transparent = true;
    g_use_atomics = false;
    int nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    auto work = [&](int tid) {
        for (int64_t i = tid; i < n; i += nthreads) {
            const double* q = rays + 7 * i;
            Ray r{{q[0], q[1], q[2]}, {q[3], q[4], q[5]}, q[6]};
            HitRecord rec;
            bool hit = h->S->world->Hit(r, Interval{tmin, tmax}, &rec);
            if (entry_id) entry_id[i] = hit ? rec.entry : -1;
            if (prim_id) prim_id[i] = hit ? rec.prim : -1;
            if (t) t[i] = hit ? rec.T : 0;
            if (normal) { normal[3 * i] = hit ? rec.Normal.X : 0; normal[3 * i + 1] = hit ? rec.Normal.Y : 0; normal[3 * i + 2] = hit ? rec.Normal.Z : 0; }
            if (front) front[i] = hit && rec.FrontFace;
            if (uv) { uv[2 * i] = hit ? rec.U : 0; uv[2 * i + 1] = hit ? rec.V : 0; }
            if (p) { p[3 * i] = hit ? rec.P.X : 0; p[3 * i + 1] = hit ? rec.P.Y : 0; p[3 * i + 2] = hit ? rec.P.Z : 0; }
        }
    };
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; i++) th.emplace_back(work, i);
    for (auto& x : th) x.join();
    g_volumes_transparent = savedT;
    g_use_atomics = savedA;
}

void orc_camera_rays(orc_scene* h, const int32_t* ij, const double* sq, const double* disk, const double* tm, int64_t n, double* out) {
    for (int64_t k = 0; k < n; k++) {
        Ray r = h->S->cam.GetRayExplicit(ij[2 * k], ij[2 * k + 1], {sq[2 * k], sq[2 * k + 1], 0}, tm[k], {disk[2 * k], disk[2 * k + 1], 0});
        double* o = out + 7 * k;
        o[0] = r.orig.X; o[1] = r.orig.Y; o[2] = r.orig.Z; o[3] = r.dir.X; o[4] = r.dir.Y; o[5] = r.dir.Z; o[6] = r.tm;
    }
}

// One BucketRenderer pass. counters_out[5] = RayCount, BVHIntersections, SamplesComputed, PixelsRendered, ShadowQueries.
// Returns wall seconds.
double orc_render(orc_scene* h, int32_t spp, int32_t depth, uint64_t seed, int32_t threads, int32_t use_atomics, double* sum, double* sumsq,
                  int64_t* counters_out) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    g_use_atomics = use_atomics != 0;
    g_stats.RayCount = 0; g_stats.BVHIntersections = 0; g_stats.SamplesComputed = 0; g_stats.PixelsRendered = 0; g_stats.ShadowQueries = 0;
    auto t0 = std::chrono::steady_clock::now();
    renderPass(h->S, spp, depth, threads, 32, seed, sum, sumsq);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (counters_out) {
        counters_out[0] = g_stats.RayCount; counters_out[1] = g_stats.BVHIntersections; counters_out[2] = g_stats.SamplesComputed;
        counters_out[3] = g_stats.PixelsRendered; counters_out[4] = g_stats.ShadowQueries;
    }
    g_use_atomics = true;
    return s;
}
int32_t orc_hardware_threads() { return (int32_t)std::max(1u, std::thread::hardware_concurrency()); }

// rt/bucket_renderer.go:275-285: scale, LinearToGamma, clamp [0,0.999], uint8(256*x), A = 255.
void orc_resolve_rgba8(const double* sum, int32_t spp, int32_t W, int32_t H, uint8_t* pix) {
    double scale = 1.0 / (double)spp;
    Interval intensity{0.0, 0.999};
    for (size_t i = 0; i < (size_t)W * H; i++) {
        for (int c = 0; c < 3; c++) pix[4 * i + c] = (uint8_t)(256 * intensity.Clamp(LinearToGamma(sum[3 * i + c] * scale)));
        pix[4 * i + 3] = 255;
    }
}

// HDRI
double orc_hdri_total_power(orc_scene* h) { return h->S->env.totalPower; }
void orc_hdri_sample(orc_scene* h, const double* xi, int64_t n, double* dir, double* emission, double* pdf) {
    for (int64_t i = 0; i < n; i++) {
        Vec3 d;
        Color e;
        double p;
        h->S->env.SampleDirectionXi(xi[2 * i], xi[2 * i + 1], d, e, p);
        dir[3 * i] = d.X; dir[3 * i + 1] = d.Y; dir[3 * i + 2] = d.Z;
        emission[3 * i] = e.X; emission[3 * i + 1] = e.Y; emission[3 * i + 2] = e.Z;
        pdf[i] = p;
    }
}
void orc_hdri_pdf(orc_scene* h, const double* dir, int64_t n, double* pdf) {
    for (int64_t i = 0; i < n; i++) pdf[i] = h->S->env.PDF({dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]});
}
void orc_hdri_lookup(orc_scene* h, const double* dir, int64_t n, double* rgb) {
    for (int64_t i = 0; i < n; i++) {
        Color c = h->S->env.Sample({dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]});
        rgb[3 * i] = c.X; rgb[3 * i + 1] = c.Y; rgb[3 * i + 2] = c.Z;
    }
}
int32_t orc_hdri_search_cdf(const double* cdf, int32_t len, double xi) {
    std::vector<double> v(cdf, cdf + len);
    return HDRIEnvironment::searchCDF(v, xi);
}
// Load a .hdr and build the distribution; rgb_out (3*w*h) may be NULL. Returns total power, < 0 on failure.
double orc_load_hdr(const char* path, int32_t* w, int32_t* hh, double* rgb_out) {
    HDRIEnvironment env;
    std::string err;
    if (!LoadHDR(path, env.width, env.height, env.data, err)) return -1.0;
    *w = env.width; *hh = env.height;
    if (rgb_out)
        for (size_t i = 0; i < env.data.size(); i++) { rgb_out[3 * i] = env.data[i].X; rgb_out[3 * i + 1] = env.data[i].Y; rgb_out[3 * i + 2] = env.data[i].Z; }
    env.BuildDistribution();
    return env.totalPower;
}

// unit-level known-answer hooks
int32_t orc_aabb_hit(const double* box6, const double* ray6, double tmin, double tmax) {
    AABB b{{box6[0], box6[1]}, {box6[2], box6[3]}, {box6[4], box6[5]}};
    Ray r{{ray6[0], ray6[1], ray6[2]}, {ray6[3], ray6[4], ray6[5]}, 0};
    return b.Hit(r, Interval{tmin, tmax});
}
int32_t orc_gamma_byte(double linear) { Interval I{0.0, 0.999}; return (uint8_t)(256 * I.Clamp(LinearToGamma(linear))); }
int32_t orc_image_height(int32_t width, double aspect) { return std::max((int)((double)width / aspect), 1); }  // rt/camera.go:299
double orc_reflectance(double cosine, double ri) { return reflectance(cosine, ri); }
void orc_refract(const double* uv, const double* n, double eta, double* out) {
    Vec3 r = Refract({uv[0], uv[1], uv[2]}, {n[0], n[1], n[2]}, eta);
    out[0] = r.X; out[1] = r.Y; out[2] = r.Z;
}
// Texture.Value of texture `tex` of a scene built from a descriptor (NoiseTexture / ImageTexture KATs).
void orc_texture_value(orc_scene* h, int32_t tex, double u, double v, const double* p, double* out) {
    Color c = h->S->textures[tex]->Value(u, v, {p[0], p[1], p[2]});
    out[0] = c.X; out[1] = c.Y; out[2] = c.Z;
}
void orc_checker(double scale, const double* even, const double* odd, const double* p, double* out) {
    SolidColor e, o;
    e.Albedo = {even[0], even[1], even[2]};
    o.Albedo = {odd[0], odd[1], odd[2]};
    CheckerTexture c;
    c.invScale = 1.0 / scale; c.even = &e; c.odd = &o;
    Color v = c.Value(0, 0, {p[0], p[1], p[2]});
    out[0] = v.X; out[1] = v.Y; out[2] = v.Z;
}
// Single-primitive Hit for interval-convention KATs: kind = RTX_GEOM_*; params: sphere (c,r) 4 / quad (Q,u,v) 9 /
// triangle 9 / plane (p,n) 6. Returns hit flag; out = t, nx,ny,nz, front, u, v.
int32_t orc_prim_hit(int32_t kind, const double* prm, const double* ray7, double tmin, double tmax, double* out) {
    std::unique_ptr<Lambertian> m(new Lambertian());
    std::unique_ptr<Hittable> h;
    if (kind == RTX_GEOM_SPHERE) h.reset(Sphere::New({prm[0], prm[1], prm[2]}, {0, 0, 0}, prm[3], m.get()));
    else if (kind == RTX_GEOM_QUAD) h.reset(Quad::New({prm[0], prm[1], prm[2]}, {prm[3], prm[4], prm[5]}, {prm[6], prm[7], prm[8]}, m.get()));
    else if (kind == RTX_GEOM_TRIANGLE) h.reset(Triangle::New({prm[0], prm[1], prm[2]}, {prm[3], prm[4], prm[5]}, {prm[6], prm[7], prm[8]}, m.get()));
    else {
        auto p = new Plane();
        p->Point = {prm[0], prm[1], prm[2]}; p->Normal = Vec3{prm[3], prm[4], prm[5]}.Unit(); p->Mat = m.get();
        h.reset(p);
    }
    Ray r{{ray7[0], ray7[1], ray7[2]}, {ray7[3], ray7[4], ray7[5]}, ray7[6]};
    HitRecord rec;
    bool hit = h->Hit(r, Interval{tmin, tmax}, &rec);
    if (hit) { out[0] = rec.T; out[1] = rec.Normal.X; out[2] = rec.Normal.Y; out[3] = rec.Normal.Z; out[4] = rec.FrontFace; out[5] = rec.U; out[6] = rec.V; }
    return hit;
}

}  // extern "C"
