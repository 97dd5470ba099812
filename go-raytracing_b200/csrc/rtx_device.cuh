// rtx_device.cuh — device-side data layout and the ray/scene intersection core (sm_100a).
//
// Precision model: the reference computes in float64 with no FMA contraction. Primitive tests,
// instance transforms and hit points are therefore evaluated in float64 in the reference's exact operation
// order (this TU is compiled with --fmad=false), so primitive ids AND t agree bit-for-bit with the CPU
// restatement. Only the BVH box tests run in float32, made conservative (outward-rounded boxes, padded ray
// origin, relative slack on tnear/tfar) so that a box the float64 ray touches is never culled.
// B200 has a half-rate FP64 pipe, which is what makes this affordable.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtx_b200.h"

#define RTX_STACK_SIZE 64   /* checked against the built hierarchy at upload (rtx_api.cu) */
#define RTX_TRI_D 12        /* doubles per triangle record */

// ---- 256-bit global loads / stores (sm_100: LDG.E.256 / STG.E.256) --------------------------------------------------------
// Every lane of a trace warp fetches from a different 128-byte line, and the L1 accepts one line per cycle per SM whatever
// the access width: the cost of fetching a node or a primitive is its NUMBER OF LOAD INSTRUCTIONS. Blackwell's 32-byte
// per-lane accesses halve it (node: 4 instead of 7, triangle: 3 instead of 5, ray: 2 instead of 4). Pointers must be 32-byte aligned.
struct F8 { float4 a, b; };
struct D4 { double x, y, z, w; };
__device__ __forceinline__ F8 ldg256f(const void* p) {   // read-only data (scene)
    F8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
}
__device__ __forceinline__ D4 ldg256d(const void* p) {   // read-only data (scene)
    D4 r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
// The path pool (records, hit records, shadow requests, queue slots) is a STREAM: every byte is written once by one kernel and read once by
// the next, a gigabyte per wavefront iteration. With the default policy it washes through L2 and evicts the scene (ncu: lts hit rate 81 %,
// ~90 B of scene data per ray from DRAM although the whole scene is a quarter of L2). RTX_STREAM_CS marks these accesses cache-streaming
// (ld/st.global.cs: evict-first), so that nodes and triangles stay resident.
#ifndef RTX_STREAM_CS
#define RTX_STREAM_CS 1
#endif
__device__ __forceinline__ D4 ld256d(const void* p) {    // data other kernels of the pass write (path pool)
    D4 r;
#if RTX_STREAM_CS
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
#else
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
#endif
    return r;
}
__device__ __forceinline__ void st256d(void* p, double x, double y, double z, double w) {
#if RTX_STREAM_CS
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x), "d"(y), "d"(z), "d"(w) : "memory");
#else
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x), "d"(y), "d"(z), "d"(w) : "memory");
#endif
}
__device__ __forceinline__ float4 ldrec4(const void* p) {   // a float4 of a pool record
#if RTX_STREAM_CS
    return __ldcs(reinterpret_cast<const float4*>(p));
#else
    return *reinterpret_cast<const float4*>(p);
#endif
}
__device__ __forceinline__ void strec4(void* p, float4 v) {
#if RTX_STREAM_CS
    __stcs(reinterpret_cast<float4*>(p), v);
#else
    *reinterpret_cast<float4*>(p) = v;
#endif
}
__device__ __forceinline__ int ldq(const int* p) {
#if RTX_STREAM_CS
    return __ldcs(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ void stq(int* p, int v) {
#if RTX_STREAM_CS
    __stcs(p, v);
#else
    *p = v;
#endif
}
#define RTX_INF_D (__longlong_as_double(0x7ff0000000000000LL))

struct DEntry {      // 32 B, one per world entry (rt/hittable_list.go:16 insertion order)
    int kind;        // RTX_GEOM_*
    int index;       // primitive index (device arrays), or unused for groups
    int xf_begin, xf_count;  // xf_count bit 16 (RTX_XF_CANON): the chain is [Translate][RotateY][Scale] and S.xf_canon holds it
    int volume;      // -1 or index into volumes
    int rank;        // test-order rank inside the reference BVH (exact-tie resolution only)
    int a;           // LIST: first item   MESH: BLAS root node
    int b;           // LIST: item count   MESH: first triangle of the mesh (device order)
};
struct DXform {      // 64 B
    double a[3];     // TRANSLATE offset | ROTATE_Y (sin, cos, -) | SCALE factor
    double b[3];     // SCALE inverse factor
    int type, pad;
};
struct DVolume { double neg_inv_density; int mat, pad; };
struct DMaterial {   // 40 B
    double fuzz, ior;
    float albedo[3];
    int type, tex, pad;
};
struct DTexture {    // 32 B
    double inv_scale;
    float color[3];
    int type, even, odd;
};

struct DevScene {
    const float4* nodes;  // wide BVH: 8 x float4 (128 B, 128-B aligned) per 4-wide node; TLAS and all BLAS share the array
    int tlas_root;        // -1: no bounded entries
    int n_entries;
    int n_nodes, n_tris_total;   // array sizes (read by the RTX_CHECKED build's index assertions)
    const DEntry* entries;
    const int* unbounded;  // entries tested for every ray (infinite Plane: universe bbox, rt/plane.go:17)
    int n_unbounded;
    // Worlds of a few bounded entries around a mesh (CornellBoxLucy: 6 quads + 10 instances): the top level is a LIST, not a
    // hierarchy. Every ray tests all entry boxes when it enters the ray pool — all lanes of the refill, no NODE rounds — and
    // starts with its entries on the stack, nearest first, each carrying its entry distance (see RTX_TLAS_CODE in rtx_trace.cuh).
    // The boxes travel INSIDE this struct (a __grid_constant__ kernel parameter, i.e. the constant bank): the refill loop indexes them with
    // a warp-uniform counter, so they are uniform-datapath operands instead of 32 global loads per ray.
    float4 tlas_boxes[2 * 16];  // [2 * n_tlas_flat]: (lo.xyz, entry index as int bits) (hi.xyz, -), float32 rounded outward
    int n_tlas_flat;            // 0: the top level is traversed as a hierarchy from tlas_root
    const double* spheres;  // 8 doubles: c0.xyz, vel.xyz, radius, -
    const int* sph_mat;
    const double* quads;    // 16 doubles: Q, u, v, w, normal, D (rt/quad.go:16-33)
    const int* quad_mat;
    const double* tris;     // 12 doubles (96 B, 3 x LDG.256): v0, e1 = v1-v0, e2 = v2-v0, unit normal (rt/triangle.go:19-25)
    const int4* tri_info;   // x: primitive id inside its geometry (face order), y: material, z: rank, w: -
    const float4* tris32;   // 3 x float4 (48 B) per triangle, same order as `tris`: (v0, M_v0) (e1, M_e1) (e2, M_e2) rounded to float32, M = largest
                            // |component| — the record of the conservative float32 pre-test (tri_pretest_reject); NULL: no pre-test
    const double* planes;   // 8 doubles: point, normal, -,-
    const int* plane_mat;
    const double* circles;  // 8 doubles: center, unit normal, radius, D = normal . center (rt/circle.go:14-20)
    const int* circle_mat;
    const double* perlin_vec;  // [n][256][3] (rt/noise.go:9)
    const int* perlin_perm;    // [n][3][256]
    const int4* flat_simple;   // world entries that are a bare primitive (no wrapper, no Volume): (kind, primitive, entry, rank) — the flat kernels' fast path
    const int* flat_complex;   // the other entries (Box lists, wrapped primitives, volumes)
    int n_flat_simple, n_flat_complex;
    const float4* img_rgb;     // ImageLoader.data of every image, back to back
    const int4* img_dim;       // (width, height, first pixel lo, first pixel hi)
    int n_images;              // > 0: hit records carry (u, v)
    const int2* list_items;  // (kind, device index)
    const DXform* xforms;
    const double* xf_canon;  // 12 doubles per entry: offset xyz, sin | cos, inverse scale xyz | scale xyz, has-scale flag (identity values where an op is absent)
    const DVolume* volumes;
    const DMaterial* mats;
    const DTexture* texs;
    const int* light_quads;
    int n_lights;
    // HDRI
    const float4* env_tex;  // w*h linear RGB
    int env_w, env_h, env_is;
    double env_rot;
    const double* env_marg;  // h+1
    const double* env_cond;  // h*(w+1)
    const double* env_pdf;   // w*h (normalised, rt/hdri.go:217-219)
    double env_total;
    // The reference stores every BVH leaf as BOTH children of its node (rt/bvh.go:141), so BVHNode.Hit runs each leaf
    // twice (rt/bvh.go:228-236). Deterministic primitives are unaffected, but Volume.Hit draws a fresh random number
    // per call (rt/volume.go:66): under NewBVHNodeFromList a volume gets two independent free-flight samples per
    // query and the nearer one wins, i.e. the medium is effectively twice as dense. 2 when world_is_bvh, else 1.
    int vol_draws;
};

struct RayD {
    double ox, oy, oz, dx, dy, dz, tm;
};
struct RayF {  // float32 ray for the box tests: t_plane = fma(plane, i, c), with c padded so that t_near only errs low, t_far high
    float ix, iy, iz;     // 1 / d
    float cnx, cny, cnz;  // -(o * i) - E   (near planes)
    float cfx, cfy, cfz;  // -(o * i) + E   (far planes),  E = 2^-21 |o * i|
    int offx, offy, offz; // byte offset of the NEAR plane inside a node's {lo,hi} float4 pair: 0, or 16 when d < 0
};
struct Hit {
    double t;
    int entry;  // -1 = miss
    int kind;   // RTX_GEOM_SPHERE..PLANE / RTX_GEOM_CIRCLE of the primitive that was hit, or RTX_KIND_VOLUME
    int prim;   // device primitive index (spheres/quads/tris/planes arrays)
    int item;   // primitive id inside the entry's geometry (list item / mesh face / 0)
};
#define RTX_KIND_VOLUME 15   /* internal: the hit is a Volume medium (not an RTX_GEOM_* value) */

struct TraceCounters {
    unsigned nodes, tris, spheres, quads, planes;
};

// ---- small float64 helpers in the reference's operation order -------------------------------------------------
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 sub(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 add(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 scale(D3 a, double t) { return d3(t * a.x, t * a.y, t * a.z); }
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // (x+y)+z like rt/vec3.go:79
__device__ __forceinline__ D3 cross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double len2(D3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ D3 unit(D3 a) {  // rt/vec3.go:32-38
    double l = sqrt(len2(a));
    if (l == 0) return a;
    return scale(a, 1 / l);
}
__device__ __forceinline__ D3 ld3(const double* p) { return d3(p[0], p[1], p[2]); }

// ---- instance transforms (rt/transform.go:93-102, :159-187, :408-440): ray into object space, outermost first --
#define RTX_XF_CANON 0x10000
#define RTX_XF_COUNT(e) ((e).xf_count & 0xffff)
__device__ __forceinline__ void xform_ray_chain(const DevScene& S, const DEntry& e, RayD& r) {
    const int n = RTX_XF_COUNT(e);
    for (int k = 0; k < n; k++) {
        const DXform& x = S.xforms[e.xf_begin + k];
        if (x.type == RTX_XF_TRANSLATE) {
            r.ox -= x.a[0]; r.oy -= x.a[1]; r.oz -= x.a[2];
        } else if (x.type == RTX_XF_ROTATE_Y) {
            double s = x.a[0], c = x.a[1];
            double ox = c * r.ox - s * r.oz, oz = s * r.ox + c * r.oz;
            double dx = c * r.dx - s * r.dz, dz = s * r.dx + c * r.dz;
            r.ox = ox; r.oz = oz; r.dx = dx; r.dz = dz;
        } else {
            r.ox *= x.b[0]; r.oy *= x.b[1]; r.oz *= x.b[2];
            r.dx *= x.b[0]; r.dy *= x.b[1]; r.dz *= x.b[2];
        }
    }
}
// The wrappers the shipped scenes build (Transform.Apply, rt/transform.go:24-46; Translate(RotateY(box)), rt/scenes.go:520-538)
// are always Translate outside RotateY outside Scale. For those, one 64-byte record replaces the op loop; absent ops hold
// identity values, for which the float64 arithmetic below is exact (x - 0, 1 * x - 0 * z, x * 1), so the result is
// bit-identical to applying only the ops that are present.
__device__ __forceinline__ void xform_ray(const DevScene& S, int ei, const DEntry& e, RayD& r) {
    if (e.xf_count == 0) return;
    if (e.xf_count & RTX_XF_CANON) {
        const D4 a = ldg256d(S.xf_canon + 12 * (size_t)ei), b = ldg256d(S.xf_canon + 12 * (size_t)ei + 4);  // (offx,offy,offz,sin) (cos,invx,invy,invz)
        const double tx = r.ox - a.x, ty = r.oy - a.y, tz = r.oz - a.z;
        const double sn = a.w, cs = b.x;
        const double ox = cs * tx - sn * tz, oz = sn * tx + cs * tz;
        const double dx = cs * r.dx - sn * r.dz, dz = sn * r.dx + cs * r.dz;
        r.ox = ox * b.y; r.oy = ty * b.z; r.oz = oz * b.w;
        r.dx = dx * b.y; r.dy = r.dy * b.z; r.dz = dz * b.w;
        return;
    }
    xform_ray_chain(S, e, r);
}
// hit point and normal back to world space, innermost first
__device__ __forceinline__ void xform_back_chain(const DevScene& S, const DEntry& e, D3& P, D3& N);
// Canonical chains ([Translate][RotateY][Scale], see xform_ray): one 96-byte record instead of a loop over op records. Absent
// Translate / RotateY hold identity values (x + 0, 1 * x + 0 * z: exact); an absent Scale is skipped, because its
// renormalisation of the normal (rt/transform.go:437) is not an identity in floating point.
__device__ __forceinline__ void xform_back(const DevScene& S, int ei, const DEntry& e, D3& P, D3& N) {
    if (!(e.xf_count & RTX_XF_CANON)) { xform_back_chain(S, e, P, N); return; }
    const D4 a = ldg256d(S.xf_canon + 12 * (size_t)ei), b = ldg256d(S.xf_canon + 12 * (size_t)ei + 4), c = ldg256d(S.xf_canon + 12 * (size_t)ei + 8);
    if (c.w != 0.0) {   // Scale (innermost): rt/transform.go:430-438
        P.x *= c.x; P.y *= c.y; P.z *= c.z;
        N = unit(d3(N.x * b.y, N.y * b.z, N.z * b.w));
    }
    const double sn = a.w, cs = b.x;   // RotateY: rt/transform.go:176-184
    const double px = cs * P.x + sn * P.z, pz = -sn * P.x + cs * P.z;
    const double nx = cs * N.x + sn * N.z, nz = -sn * N.x + cs * N.z;
    P.x = px; P.z = pz; N.x = nx; N.z = nz;
    P.x += a.x; P.y += a.y; P.z += a.z;   // Translate: rt/transform.go:99
}
__device__ __forceinline__ void xform_back_chain(const DevScene& S, const DEntry& e, D3& P, D3& N) {
    for (int k = RTX_XF_COUNT(e) - 1; k >= 0; k--) {
        const DXform& x = S.xforms[e.xf_begin + k];
        if (x.type == RTX_XF_TRANSLATE) {
            P.x += x.a[0]; P.y += x.a[1]; P.z += x.a[2];
        } else if (x.type == RTX_XF_ROTATE_Y) {
            double s = x.a[0], c = x.a[1];
            double px = c * P.x + s * P.z, pz = -s * P.x + c * P.z;
            double nx = c * N.x + s * N.z, nz = -s * N.x + c * N.z;
            P.x = px; P.z = pz; N.x = nx; N.z = nz;
        } else {
            P.x *= x.a[0]; P.y *= x.a[1]; P.z *= x.a[2];
            N = unit(d3(N.x * x.b[0], N.y * x.b[1], N.z * x.b[2]));
        }
    }
}

// ---- primitive tests: return t, or NaN when the reference's Hit would return false before its interval test ----
#define RTX_NAN_D (__longlong_as_double(0x7ff8000000000000LL))

// rt/sphere.go:63-85. Open interval (Surrounds): the root choice depends on the interval, so it is evaluated here.
__device__ __forceinline__ double isect_sphere(const double* s, const RayD& r, double tmin, double tmax) {
    D3 c = d3(s[0] + r.tm * s[3], s[1] + r.tm * s[4], s[2] + r.tm * s[5]);
    D3 o = d3(r.ox, r.oy, r.oz), d = d3(r.dx, r.dy, r.dz);
    D3 oc = sub(c, o);
    double a = len2(d);
    double h = dot(d, oc);
    double cc = len2(oc) - s[6] * s[6];
    double disc = h * h - a * cc;
    if (disc < 0) return RTX_NAN_D;
    double sq = sqrt(disc);
    double root = (h - sq) / a;
    if (!(tmin < root && root < tmax)) {
        root = (h + sq) / a;
        if (!(tmin < root && root < tmax)) return RTX_NAN_D;
    }
    return root;
}
// The same with an optionally closed upper end: a sphere found at exactly the current best t must still reach the tie rule
// (Best::offer) — what Best::tmax_for does with nextafter, without the nextafter.
__device__ __forceinline__ double isect_sphere_incl(const double* s, const RayD& r, double tmin, double tmax, bool incl) {
    D3 c = d3(s[0] + r.tm * s[3], s[1] + r.tm * s[4], s[2] + r.tm * s[5]);
    D3 o = d3(r.ox, r.oy, r.oz), d = d3(r.dx, r.dy, r.dz);
    D3 oc = sub(c, o);
    double a = len2(d);
    double h = dot(d, oc);
    double cc = len2(oc) - s[6] * s[6];
    double disc = h * h - a * cc;
    if (disc < 0) return RTX_NAN_D;
    double sq = sqrt(disc);
    double root = (h - sq) / a;
    if (!(tmin < root && (root < tmax || (incl && root == tmax)))) {
        root = (h + sq) / a;
        if (!(tmin < root && (root < tmax || (incl && root == tmax)))) return RTX_NAN_D;
    }
    return root;
}
// rt/quad.go:44-84 (closed interval is applied by the caller)
__device__ __forceinline__ double isect_quad(const double* q, const RayD& r, double tmin, double tmax, double* uv) {
    const D4 q3 = ldg256d(q + 12);   // normal, D
    D3 n = d3(q3.x, q3.y, q3.z), o = d3(r.ox, r.oy, r.oz), d = d3(r.dx, r.dy, r.dz);
    double denom = dot(n, d);
    if (fabs(denom) < 1e-8) return RTX_NAN_D;
    double t = (q3.w - dot(n, o)) / denom;
    if (!(tmin <= t && t <= tmax)) return RTX_NAN_D;
    const D4 q0 = ldg256d(q), q1 = ldg256d(q + 4), q2 = ldg256d(q + 8);   // Q u | u v | v w
    D3 P = add(o, scale(d, t));
    D3 pl = sub(P, d3(q0.x, q0.y, q0.z));
    D3 w = d3(q2.y, q2.z, q2.w);
    double alpha = dot(w, cross(pl, d3(q1.z, q1.w, q2.x)));
    double beta = dot(w, cross(d3(q0.w, q1.x, q1.y), pl));
    if (!(0.0 <= alpha && alpha <= 1.0) || !(0.0 <= beta && beta <= 1.0)) return RTX_NAN_D;
    if (uv) { uv[0] = alpha; uv[1] = beta; }
    return t;
}
// rt/triangle.go:57-104 Möller–Trumbore; e1/e2 are the same float64 values the reference recomputes per call.
__device__ __forceinline__ double isect_tri(const double* tp, const RayD& r, double* uv) {
    const D4 a0 = ldg256d(tp), a1 = ldg256d(tp + 4), a2 = ldg256d(tp + 8);   // v0 e1.x | e1.yz e2.xy | e2.z n
    D3 v0 = d3(a0.x, a0.y, a0.z), e1 = d3(a0.w, a1.x, a1.y), e2 = d3(a1.z, a1.w, a2.x);
    D3 o = d3(r.ox, r.oy, r.oz), d = d3(r.dx, r.dy, r.dz);
    D3 h = cross(d, e2);
    double a = dot(e1, h);
    if (fabs(a) < 1e-8) return RTX_NAN_D;
    double f = 1.0 / a;
    D3 s = sub(o, v0);
    double u = f * dot(s, h);
    if (u < 0.0 || u > 1.0) return RTX_NAN_D;
    D3 q = cross(s, e1);
    double v = f * dot(d, q);
    if (v < 0.0 || u + v > 1.0) return RTX_NAN_D;
    if (uv) { uv[0] = u; uv[1] = v; }
    return f * dot(e2, q);
}
// Conservative float32 pre-test of a triangle (48-byte float32 record, the float64 record only "when a candidate is
// (re)tested in double"): true = the float64 test of rt/triangle.go:57-104 CERTAINLY rejects this ray for the interval [tmin, tmax], so
// the 96-byte float64 record need not be fetched. The same Moller-Trumbore terms in float32, each compared against its acceptance bound
// with an error margin. Margins (u = 2^-24; M_x = largest |component| of x; Ms = M_o + M_v0 bounds |o - v0|): rounding the float64 inputs
// to float32, the products and the sums give |da| <= 48u Md Me1 Me2, |dU| <= 60u Ms Md Me2, |dV| <= 60u Ms Md Me1, |dT| <= 60u Ms Me1 Me2
// (first order); the margins below use 128u = 2^-17... i.e. K = 64 x 2^-23, more than twice that. A NaN or an infinity anywhere makes
// every comparison false: not rejected, the float64 test decides. |a| inside its margin: undecidable, the float64 test decides (it rejects
// |a| < 1e-8 itself). tmin_f <= tmin and tmax_f >= tmax (rounded outward by the caller).
#define RTX_TRI_PRETEST_K 7.62939453e-6f   /* 64 * 2^-23 */
__device__ __forceinline__ bool tri_pretest_reject(const float4* __restrict__ t32, float ox, float oy, float oz, float dx, float dy, float dz, float Mo, float Md,
                                                   float tmin_f, float tmax_f) {
    const float4 A = __ldg(t32), B = __ldg(t32 + 1), C = __ldg(t32 + 2);
    const float hx = dy * C.z - dz * C.y, hy = dz * C.x - dx * C.z, hz = dx * C.y - dy * C.x;
    const float a = B.x * hx + B.y * hy + B.z * hz;
    const float sx = ox - A.x, sy = oy - A.y, sz = oz - A.z;
    const float U = sx * hx + sy * hy + sz * hz;
    const float qx = sy * B.z - sz * B.y, qy = sz * B.x - sx * B.z, qz = sx * B.y - sy * B.x;
    const float V = dx * qx + dy * qy + dz * qz;
    const float Tt = C.x * qx + C.y * qy + C.z * qz;
    const float Ms = Mo + A.w, k = RTX_TRI_PRETEST_K;
    const float eA = k * Md * B.w * C.w + 1e-30f, eU = k * Md * C.w * Ms + 1e-30f, eV = k * Md * B.w * Ms + 1e-30f, eT = k * B.w * C.w * Ms + 1e-30f;
    const float aa = fabsf(a);
    if (!(aa > eA)) return false;
    const float Us = a > 0.f ? U : -U, Vs = a > 0.f ? V : -V, Ts = a > 0.f ? Tt : -Tt;
    if (Us < -eU) return true;                              // u < 0
    if (Us > aa + eU + eA) return true;                     // u > 1
    if (Vs < -eV) return true;                              // v < 0
    if (Us + Vs > aa + eU + eV + eA) return true;           // u + v > 1
    if (Ts < tmin_f * (aa - eA) - eT) return true;          // t < tmin
    if (Ts > tmax_f * (aa + eA) + eT) return true;          // t > tmax
    return false;
}

// rt/plane.go:24-34 (open interval applied by the caller)
__device__ __forceinline__ double isect_plane(const double* p, const RayD& r) {
    D3 n = ld3(p + 3), o = d3(r.ox, r.oy, r.oz), d = d3(r.dx, r.dy, r.dz);
    double denom = dot(n, d);
    if (fabs(denom) < 1e-8) return RTX_NAN_D;
    return dot(sub(ld3(p), o), n) / denom;
}

// rt/circle.go:33-52 (closed interval: rayT.Contains)
__device__ __forceinline__ double isect_circle(const double* c, const RayD& r, double tmin, double tmax) {
    const D4 c0 = ldg256d(c), c1 = ldg256d(c + 4);   // center.xyz n.x | n.yz radius D
    D3 n = d3(c0.w, c1.x, c1.y), o = d3(r.ox, r.oy, r.oz), d = d3(r.dx, r.dy, r.dz);
    double denom = dot(n, d);
    if (fabs(denom) < 1e-8) return RTX_NAN_D;
    double t = (c1.w - dot(n, o)) / denom;
    if (!(tmin <= t && t <= tmax)) return RTX_NAN_D;
    D3 P = add(o, scale(d, t));
    double dist = sqrt(len2(sub(P, d3(c0.x, c0.y, c0.z))));
    if (dist > c1.z) return RTX_NAN_D;
    return t;
}

__device__ __forceinline__ bool kind_closed(int kind) { return kind == RTX_GEOM_QUAD || kind == RTX_GEOM_TRIANGLE || kind == RTX_GEOM_CIRCLE || kind == RTX_KIND_VOLUME; }

// Generic primitive t with the reference's interval convention against [tmin, tmax]. Out of line (one copy of the
// four float64 tests in each kernel keeps the trace loop inside the instruction cache); everything travels in registers.
__device__ __noinline__ double isect_prim_ool(int kind, const double* p, double ox, double oy, double oz, double dx, double dy, double dz, double tm,
                                              double tmin, double tmax) {
    RayD r; r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz; r.tm = tm;
    double t;
    if (kind == RTX_GEOM_SPHERE) return isect_sphere(p, r, tmin, tmax);
    if (kind == RTX_GEOM_QUAD) return isect_quad(p, r, tmin, tmax, nullptr);
    if (kind == RTX_GEOM_CIRCLE) return isect_circle(p, r, tmin, tmax);
    if (kind == RTX_GEOM_TRIANGLE) {
        t = isect_tri(p, r, nullptr);
        return (tmin <= t && t <= tmax) ? t : RTX_NAN_D;
    }
    t = isect_plane(p, r);
    return (tmin < t && t < tmax) ? t : RTX_NAN_D;
}
__device__ __forceinline__ double isect_prim(const DevScene& S, int kind, int idx, const RayD& r, double tmin, double tmax, TraceCounters* tc) {
    const double* p;
    if (kind == RTX_GEOM_SPHERE) { if (tc) tc->spheres++; p = S.spheres + 8 * (size_t)idx; }
    else if (kind == RTX_GEOM_QUAD) { if (tc) tc->quads++; p = S.quads + 16 * (size_t)idx; }
    else if (kind == RTX_GEOM_TRIANGLE) { if (tc) tc->tris++; p = S.tris + RTX_TRI_D * (size_t)idx; }
    else if (kind == RTX_GEOM_CIRCLE) { if (tc) tc->quads++; p = S.circles + 8 * (size_t)idx; }   // counted with the quads (planar, 64-128 B)
    else { if (tc) tc->planes++; p = S.planes + 8 * (size_t)idx; }
    return isect_prim_ool(kind, p, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, r.tm, tmin, tmax);
}

// ---- float32 conservative ray ---------------------------------------------------------------------------------
// Exact slab parameter: t = (p - o) / d with the float64 ray. Computed: fma(p, i~, c) with i~ = fl(1 / fl(d)) (relative
// error <= 2^-23) and c = -fl(fl(o) * i~) -/+ E. Then |computed - t| <= 2^-21 |t| + 2^-22 |o / d|, so with
// E = 2^-21 |o * i| and the 2^-21 relative slack applied to tnear / tfar in node_test, a box the float64 ray touches
// is never culled. |i| is clamped to 1e30 (a zero direction component would give inf and NaN slab values that never
// cull: an axis-parallel ray would then walk every node its other axes reach). With the clamp, (p - o) * 1e30 keeps the
// sign of the exact slab parameter and dwarfs every finite t, so such rays cull exactly like the reference's 1/0 = inf.
#define RTX_BOX_EPS 4.76837158e-7f  /* 2^-21 */
__device__ __forceinline__ void make_rayf(const RayD& r, RayF& f) {
    const float ox = __double2float_rn(r.ox), oy = __double2float_rn(r.oy), oz = __double2float_rn(r.oz);
    const float dx = __double2float_rn(r.dx), dy = __double2float_rn(r.dy), dz = __double2float_rn(r.dz);
    f.ix = 1.0f / dx; f.iy = 1.0f / dy; f.iz = 1.0f / dz;
    if (!(fabsf(f.ix) <= 1e30f)) f.ix = copysignf(1e30f, dx);
    if (!(fabsf(f.iy) <= 1e30f)) f.iy = copysignf(1e30f, dy);
    if (!(fabsf(f.iz) <= 1e30f)) f.iz = copysignf(1e30f, dz);
    const float px = ox * f.ix, py = oy * f.iy, pz = oz * f.iz;
    const float ex = fmaf(fabsf(px), RTX_BOX_EPS, 1e-30f), ey = fmaf(fabsf(py), RTX_BOX_EPS, 1e-30f), ez = fmaf(fabsf(pz), RTX_BOX_EPS, 1e-30f);
    f.cnx = -px - ex; f.cfx = -px + ex;
    f.cny = -py - ey; f.cfy = -py + ey;
    f.cnz = -pz - ez; f.cfz = -pz + ez;
    f.offx = signbit(dx) ? 16 : 0; f.offy = signbit(dy) ? 16 : 0; f.offz = signbit(dz) ? 16 : 0;
}

// One 4-wide node (128 B: {lox,hix,loy,hiy,loz,hiz} float4 pairs, child int4, pad): conservative entry distances
// (inf = culled) of the 4 children. Three 256-bit loads bring the {lo, hi} pairs, one 128-bit load the child links.
__device__ __forceinline__ void node_test(const float4* __restrict__ nodes, int node, const RayF& f, float tmin, float tmax, float d[4], int c[4]) {
    const char* nb = reinterpret_cast<const char*>(nodes) + (size_t)node * 128;
    const F8 X = ldg256f(nb), Y = ldg256f(nb + 32), Z = ldg256f(nb + 64);   // {lo, hi} of the four children, per axis
    int4 ch;
    asm("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(ch.x), "=r"(ch.y), "=r"(ch.z), "=r"(ch.w) : "l"(nb + 96));
    const bool sx = f.offx != 0, sy = f.offy != 0, sz = f.offz != 0;   // direction negative: the near plane is hi
#define RTX_SEL4(s, a, b) make_float4(s ? a.x : b.x, s ? a.y : b.y, s ? a.z : b.z, s ? a.w : b.w)
    const float4 nx = RTX_SEL4(sx, X.b, X.a), fx = RTX_SEL4(sx, X.a, X.b);
    const float4 ny = RTX_SEL4(sy, Y.b, Y.a), fy = RTX_SEL4(sy, Y.a, Y.b);
    const float4 nz = RTX_SEL4(sz, Z.b, Z.a), fz = RTX_SEL4(sz, Z.a, Z.b);
#undef RTX_SEL4
#define RTX_CHILD(k, comp)                                                                                   \
    {                                                                                                        \
        float tn = fmaxf(fmaxf(fmaf(nx.comp, f.ix, f.cnx), fmaf(ny.comp, f.iy, f.cny)), fmaxf(fmaf(nz.comp, f.iz, f.cnz), tmin)); \
        float tf = fminf(fminf(fmaf(fx.comp, f.ix, f.cfx), fmaf(fy.comp, f.iy, f.cfy)), fminf(fmaf(fz.comp, f.iz, f.cfz), tmax)); \
        tn = fmaf(-fabsf(tn), RTX_BOX_EPS, tn);                                                              \
        tf = fmaf(fabsf(tf), RTX_BOX_EPS, tf);                                                               \
        d[k] = (tn <= tf) ? tn : __int_as_float(0x7f800000);                                                 \
    }
    RTX_CHILD(0, x) RTX_CHILD(1, y) RTX_CHILD(2, z) RTX_CHILD(3, w)
#undef RTX_CHILD
    c[0] = ch.x; c[1] = ch.y; c[2] = ch.z; c[3] = ch.w;
}

// Exact ties in t: the reference keeps whichever primitive its traversal order and interval conventions favour
// (quads/triangles accept t == max, spheres/planes do not; rt/quad.go:53, rt/triangle.go:89, rt/sphere.go:80).
// For two candidates a (tested earlier) and b (tested later) at the same t the survivor is b iff b is closed.
__device__ __forceinline__ bool tie_candidate_wins(int crank_e, int crank_p, int ckind, int brank_e, int brank_p, int bkind) {
    bool cand_later = (crank_e > brank_e) || (crank_e == brank_e && crank_p > brank_p);
    return cand_later ? kind_closed(ckind) : !kind_closed(bkind);
}

// Full hit record (rec.P, rec.Normal against the ray, FrontFace, material) of a finished query.
struct HitInfo {
    D3 P, N;
    int mat;
    bool front;
    double u, v;
};
// Compile-time scene vocabulary of the one-kernel bounce of flat worlds (k_bounce_flat<.., FEAT>): a variant only contains the code of
// the primitive kinds, materials, textures and light samplers its mask names, and rtx_render_pass picks the smallest variant that covers
// the scene (ncu: the all-features kernel is 9904 SASS instructions and stalls 2 cycles per issue on instruction fetch, profiles/r01_k_bounce_flat_hdri.md).
// Pruned branches are unreachable for a covered scene, so every variant computes what RTX_F_ALL computes.
#define RTX_F_QUAD 1u
#define RTX_F_SPHERE 2u
#define RTX_F_PLANE 4u
#define RTX_F_OTHER_PRIM 8u    /* Circle, bare Triangle */
#define RTX_F_COMPLEX 16u      /* Box / Pyramid lists, wrapper chains, Volumes */
#define RTX_F_ENV 32u          /* HDRI environment: lookup at a miss, importance sampling */
#define RTX_F_LIGHTS 64u       /* registered area lights: next-event estimation */
#define RTX_F_TEX_X 128u       /* NoiseTexture, ImageTexture */
#define RTX_F_CAM_SLOW 256u    /* camera motion / free camera */
#define RTX_F_METAL 512u
#define RTX_F_DIELECTRIC 1024u
#define RTX_F_ISOTROPIC 2048u
#define RTX_F_MESH 4096u        /* OBJ meshes: BLAS traversal, the TRI phase, instance entry / exit */
#define RTX_F_XFORM 8192u       /* wrapper chains on mesh instances (wrapped primitives and lists count as RTX_F_COMPLEX) */
#define RTX_F_ALL 0xffffffffu
template <unsigned FEAT = RTX_F_ALL>
__device__ __forceinline__ void finalize_hit(const DevScene& S, const RayD& rw, const Hit& h, bool want_uv, HitInfo& out) {
    DEntry e = S.entries[h.entry];
    out.u = 0; out.v = 0;
    if ((FEAT & RTX_F_COMPLEX) && h.kind == RTX_KIND_VOLUME) {  // rt/volume.go:72-76
        out.P = d3(rw.ox + h.t * rw.dx, rw.oy + h.t * rw.dy, rw.oz + h.t * rw.dz);
        out.N = d3(1, 0, 0);
        out.front = true;
        out.mat = S.volumes[h.prim].mat;
        return;
    }
    RayD r = rw;
    if (FEAT & (RTX_F_COMPLEX | RTX_F_XFORM)) xform_ray(S, h.entry, e, r);
    D3 o = d3(r.ox, r.oy, r.oz), d = d3(r.dx, r.dy, r.dz);
    D3 P = add(o, scale(d, h.t));  // r.At(t), rt/ray.go:21
    D3 n;
    if ((FEAT & RTX_F_SPHERE) && h.kind == RTX_GEOM_SPHERE) {
        const double* s = S.spheres + 8 * (size_t)h.prim;
        D3 c = d3(s[0] + r.tm * s[3], s[1] + r.tm * s[4], s[2] + r.tm * s[5]);
        n = scale(sub(P, c), 1 / s[6]);  // Div(Radius)
        out.mat = S.sph_mat[h.prim];
        if (want_uv) {  // getSphereUV rt/sphere.go:53-59
            const double PI = 3.14159265358979323846;
            double theta = acos(-n.y), phi = atan2(-n.z, n.x) + PI;
            out.u = phi / (2 * PI); out.v = theta / PI;
        }
    } else if ((FEAT & RTX_F_QUAD) && h.kind == RTX_GEOM_QUAD) {
        const double* q = S.quads + 16 * (size_t)h.prim;
        n = ld3(q + 12);
        out.mat = S.quad_mat[h.prim];
        if (want_uv) { double uv[2] = {0, 0}; isect_quad(q, r, -RTX_INF_D, RTX_INF_D, uv); out.u = uv[0]; out.v = uv[1]; }
    } else if ((FEAT & (RTX_F_OTHER_PRIM | RTX_F_COMPLEX | RTX_F_MESH)) && h.kind == RTX_GEOM_TRIANGLE) {
        { const D4 tn = ldg256d(S.tris + RTX_TRI_D * (size_t)h.prim + 8); n = d3(tn.y, tn.z, tn.w); }
        out.mat = S.tri_info[h.prim].y;
        if (want_uv) { double uv[2] = {0, 0}; isect_tri(S.tris + RTX_TRI_D * (size_t)h.prim, r, uv); out.u = uv[0]; out.v = uv[1]; }
    } else if ((FEAT & (RTX_F_OTHER_PRIM | RTX_F_COMPLEX)) && h.kind == RTX_GEOM_CIRCLE) {
        const double* c = S.circles + 8 * (size_t)h.prim;
        n = ld3(c + 3);
        out.mat = S.circle_mat[h.prim];
        if (want_uv) {  // rt/circle.go:59-72
            D3 uu = fabs(n.y) > 0.9 ? unit(cross(d3(1, 0, 0), n)) : unit(cross(d3(0, 1, 0), n));
            D3 vv = cross(n, uu);
            D3 lp = sub(P, ld3(c));
            out.u = (dot(lp, uu) / c[6] + 1.0) * 0.5;
            out.v = (dot(lp, vv) / c[6] + 1.0) * 0.5;
        }
    } else if (FEAT & RTX_F_PLANE) {
        n = ld3(S.planes + 8 * (size_t)h.prim + 3);
        out.mat = S.plane_mat[h.prim];
    } else {
        n = d3(0, 1, 0); out.mat = 0;   // unreachable for a scene the mask covers
    }
    out.front = dot(d, n) < 0;  // SetFaceNormal with the object-space ray (rt/hittable.go:20-30); never recomputed afterwards
    if (!out.front) n = d3(-n.x, -n.y, -n.z);
    if ((FEAT & (RTX_F_COMPLEX | RTX_F_XFORM)) && e.xf_count) xform_back(S, h.entry, e, P, n);
    out.P = P; out.N = n;
}
