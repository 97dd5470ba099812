// rtx_trace.cuh — the scene query world.Hit(r, [tmin, tmax]) as a PERSISTENT, LANE-REFILLING warp program (sm_100a).
//
// Why this shape (profiles/r01_k_extend_baseline.md): with one thread per ray and one launch-sized grid, a warp lives as
// long as its slowest ray — in the Cornell/Lucy scene rays need anything from 3 to 60 node visits, and the first cut ran
// with 4.8-5.5 of 32 lanes active. Here a fixed grid of warps (SM count x resident blocks) pulls rays from a global cursor;
// every lane is a small state machine
//
//     NODE   an internal 4-wide node to test (float32 slabs, conservative)
//     TRI    a BLAS leaf: triangles still to test, one per round (float64 Moller-Trumbore, reference operation order)
//     ENTRY  a TLAS leaf (world entry: primitive / Box list / mesh instance / Volume) or the marker that ends an instance
//     DONE   query finished, result not yet handed back        IDLE   no ray
//
// and each round the WARP runs the one phase that the most lanes are waiting for (three ballots and a compare), so the
// instructions it issues always serve the largest available group of lanes. DONE lanes are retired and refilled together
// (the "retire" phase is just another candidate of the vote), which removes the tail effect.
//
// The traversal stack lives in shared memory, one column per thread (bank-conflict free). Exactness: see rtx_device.cuh —
// float32 box tests only cull, every accepted hit is decided by the float64 primitive tests.
#pragma once
#include "rtx_device.cuh"

#define RTX_TRACE_THREADS 128
#ifndef RTX_TRACE_BLOCKS
#define RTX_TRACE_BLOCKS 4
#endif
/* resident blocks per SM the trace kernels are compiled for (register cap 65536 / (128 * 4) = 128) */
#define RTX_ST_SENTINEL ((int)0x80000000)  /* stack marker: instance finished, back to the TLAS */
#define RTX_ST_DONE ((int)0x80000001)
#define RTX_ST_IDLE ((int)0x80000002)
// node >= 0: internal node index. node in (RTX_ST_IDLE, -1]: leaf, code = ~node
//   while in the TLAS (cur < 0): code = world entry index;  inside an instance (cur >= 0): code = first_tri << 3 | (count - 1).
// rtx_scene_upload keeps first_tri + count < 2^28, so leaf codes never collide with the three specials.

struct VolumeRng {
    uint32_t k0, k1, c0, c1, c2;
    bool transparent;  // level-1 parity protocol: volumes do not intersect
};
__device__ double2 rtx_volume_uniform(const VolumeRng& vr, int entry);  // two uniforms in (0,1), Philox (rtx_kernels.cuh)

// Closest boundary crossing of an entry's geometry in [tmin, tmax] for Volume (rt/volume.go:38-46): HittableList
// semantics (rt/hittable_list.go:31-45) over the list items, or a single primitive.
__device__ __forceinline__ double isect_boundary(const DevScene& S, const DEntry& e, const RayD& ro, double tmin, double tmax, TraceCounters* tc) {
    double closest = tmax;
    bool hit = false;
    if (e.kind == RTX_GEOM_LIST) {
        for (int k = 0; k < e.b; k++) {
            int2 it = S.list_items[e.a + k];
            double t = isect_prim(S, it.x, it.y, ro, tmin, closest, tc);
            if (t == t) { closest = t; hit = true; }
        }
    } else {
        double t = isect_prim(S, e.kind, e.index, ro, tmin, closest, tc);
        if (t == t) { closest = t; hit = true; }
    }
    return hit ? closest : RTX_NAN_D;
}

// Per-lane query state that survives between rounds.
struct Best {
    double t;
    int entry, kind, prim, item;
    int rank_e, rank_p;
    bool have;
    float ft;  // float32 upper bound of t for the box tests
    __device__ __forceinline__ void reset(double tmax) {
        t = tmax; entry = -1; kind = -1; prim = -1; item = -1; rank_e = -1; rank_p = -1; have = false;
        ft = __double2float_ru(tmax);
    }
    // candidate at parameter t (already inside the primitive's own interval convention w.r.t. [tmin, t_best])
    __device__ __forceinline__ void offer(double tc, int e, int erank, int k, int p, int it, int prank) {
        if (!(tc == tc)) return;
        if (tc == t) {
            if (!have) { if (!kind_closed(k)) return; }
            else if (!tie_candidate_wins(erank, prank, k, rank_e, rank_p, kind)) return;
        }
        t = tc; entry = e; kind = k; prim = p; item = it; rank_e = erank; rank_p = prank; have = true;
        ft = __double2float_ru(tc);
    }
    // open-interval primitives (sphere, plane) must still be allowed to tie with an existing best hit
    __device__ __forceinline__ double tmax_for(int k) const { return (kind_closed(k) || !have) ? t : nextafter(t, RTX_INF_D); }
    __device__ __forceinline__ void test_prim(const DevScene& S, int k, int idx, const RayD& r, double tmin, int e, int erank, int it, int prank, TraceCounters* tc) {
        double tt = isect_prim(S, k, idx, r, tmin, tmax_for(k), tc);
        offer(tt, e, erank, k, idx, it, prank);
    }
};

// A Policy supplies the rays and consumes the results:
//   static constexpr bool ANY_HIT;
//   double tmin() const;                                    uniform lower bound of the interval
//   void   load(int job, RayD& r, double& tmax) const;      world-space ray of job `job` (called again when an instance ends)
//   VolumeRng volume_rng(int job) const;
//   void   retire(int job, bool valid, const RayD& r, const Best& b);   warp-collective: every lane calls it, `valid` lanes own a finished query
template <class Policy, bool COUNT>
__device__ __forceinline__ void trace_persistent(const DevScene& S, Policy& P, int* cursor, int njobs, TraceCounters& tc) {
    __shared__ int s_stack[RTX_STACK_SIZE * RTX_TRACE_THREADS];
    int* const stack = s_stack + threadIdx.x;
#define STK(i) stack[(i) * RTX_TRACE_THREADS]
#define POP() do { if (sp > 0) { sp--; node = STK(sp); } else node = RTX_ST_DONE; } while (0)
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const double tmin = P.tmin();
    const float ftmin = __double2float_rd(tmin);
    const float INF = __int_as_float(0x7f800000);
    TraceCounters* const tcp = COUNT ? &tc : nullptr;

    int node = RTX_ST_IDLE, sp = 0, cur = -1, job = -1;
    RayD r;
    RayF f;
    Best B;
    bool exhausted = false;
    int round = 0;
    r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = r.tm = 0;
    f.ix = f.iy = f.iz = f.cnx = f.cny = f.cnz = f.cfx = f.cfy = f.cfz = 0; f.offx = f.offy = f.offz = 0;
    B.reset(0);

#ifdef RTX_DEBUG_LONGRAY
    int dbg_rounds = 0;
#endif
    for (;;) {
#ifdef RTX_DEBUG_LONGRAY
        if (node != RTX_ST_IDLE && node != RTX_ST_DONE && ++dbg_rounds == 20000) {
            double tm_; RayD w; P.load(job, w, tm_);
            printf("[longray] job %d node %d sp %d cur %d world o=(%.17g %.17g %.17g) d=(%.17g %.17g %.17g) tm %.17g | cur o=(%.9g %.9g %.9g) d=(%.9g %.9g %.9g) best %.9g ft %g | f.i=(%g %g %g) cn=(%g %g %g) cf=(%g %g %g)\n",
                   job, node, sp, cur, w.ox, w.oy, w.oz, w.dx, w.dy, w.dz, w.tm, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, B.t, B.ft, f.ix, f.iy, f.iz, f.cnx, f.cny, f.cnz, f.cfx, f.cfy, f.cfz);
        }
        if (node == RTX_ST_IDLE || node == RTX_ST_DONE) dbg_rounds = 0;
#endif
        const bool leaf = node < 0 && node > RTX_ST_IDLE;
        const bool sN = node >= 0;
        const bool sT = leaf && cur >= 0;
        const bool sE = (leaf && cur < 0) || node == RTX_ST_SENTINEL;
        const bool sR = node == RTX_ST_DONE || (node == RTX_ST_IDLE && !exhausted);
        const unsigned mN = __ballot_sync(FULL, sN), mT = __ballot_sync(FULL, sT), mE = __ballot_sync(FULL, sE), mR = __ballot_sync(FULL, sR);
        if ((mN | mT | mE | mR) == 0) break;
        // the phase with the most waiting lanes wins; equal counts are broken by a rotating priority, so that a lane
        // can never be starved by a long-running neighbour that keeps re-entering a "higher" phase
        round++;
        const int cN = (__popc(mN) << 2) | (round & 3), cT = (__popc(mT) << 2) | ((round + 1) & 3), cE = (__popc(mE) << 2) | ((round + 2) & 3),
                  cR = (__popc(mR) << 2) | ((round + 3) & 3);

        if (cN >= cT && cN >= cE && cN >= cR) {
            // ---- NODE: one 4-wide node per lane -------------------------------------------------------------------
            if (sN) {
                float d[4]; int c[4];
                if (COUNT) tc.nodes++;
                node_test(S.nodes, node, f, ftmin, B.ft, d, c);
#define RTX_CSWAP(i, j) if (d[j] < d[i]) { float td = d[i]; d[i] = d[j]; d[j] = td; int tcx = c[i]; c[i] = c[j]; c[j] = tcx; }
                RTX_CSWAP(0, 1) RTX_CSWAP(2, 3) RTX_CSWAP(0, 2) RTX_CSWAP(1, 3) RTX_CSWAP(1, 2)
#undef RTX_CSWAP
                if (d[3] < INF) { STK(sp) = c[3]; sp++; }
                if (d[2] < INF) { STK(sp) = c[2]; sp++; }
                if (d[1] < INF) { STK(sp) = c[1]; sp++; }
                if (d[0] < INF) node = c[0];
                else POP();
            }
        } else if (cT >= cE && cT >= cR) {
            // ---- TRI: one triangle of the pending BLAS leaf per lane ---------------------------------------------
            if (sT) {
                const int code = ~node;
                const int ti = code >> 3, rem = code & 7;
                if (COUNT) tc.tris++;
                const double t = isect_tri(S.tris + 10 * (size_t)ti, r, nullptr);
                if (tmin <= t && t <= B.t) {
                    const int4 info = __ldg(S.tri_info + ti);
                    B.offer(t, cur, S.entries[cur].rank, RTX_GEOM_TRIANGLE, ti, info.x, info.z);
                }
                if (Policy::ANY_HIT && B.have) node = RTX_ST_DONE;
                else if (rem == 0) POP();
                else node = ~(((ti + 1) << 3) | (rem - 1));
            }
        } else if (cE >= cR) {
            // ---- ENTRY: a world entry (TLAS leaf), or the end of an instance -------------------------------------
            if (sE) {
                if (node == RTX_ST_SENTINEL) {
                    double tmax_unused;
                    cur = -1;
                    P.load(job, r, tmax_unused);
                    make_rayf(r, f);
                    POP();
                } else {
                    const int ei = ~node;
                    const DEntry e = S.entries[ei];
                    RayD r2 = r;
                    xform_ray(S, ei, e, r2);
                    bool descend = false;
                    if (e.volume >= 0) {
                        const VolumeRng vr = P.volume_rng(job);
                        if (!vr.transparent) {
                            // rt/volume.go:34-79
                            double t1 = isect_boundary(S, e, r2, -RTX_INF_D, RTX_INF_D, tcp);
                            if (t1 == t1) {
                                double t2 = isect_boundary(S, e, r2, t1 + 0.0001, RTX_INF_D, tcp);
                                if (t2 == t2) {
                                    if (t1 < tmin) t1 = tmin;
                                    if (t2 > B.t) t2 = B.t;
                                    if (t1 < t2) {
                                        if (t1 < 0) t1 = 0;
                                        const double rayLength = sqrt(r.dx * r.dx + r.dy * r.dy + r.dz * r.dz);
                                        const double inside = (t2 - t1) * rayLength;
                                        const double2 uu = rtx_volume_uniform(vr, ei);
                                        const double nid = S.volumes[e.volume].neg_inv_density;
                                        double hd = nid * log(uu.x);
                                        if (S.vol_draws > 1) hd = fmin(hd, nid * log(uu.y));  // leaf visited twice, see DevScene::vol_draws
                                        if (!(hd > inside)) B.offer(t1 + hd / rayLength, ei, e.rank, RTX_KIND_VOLUME, e.volume, 0, 0);
                                    }
                                }
                            }
                        }
                    } else if (e.kind == RTX_GEOM_MESH) {
                        STK(sp) = RTX_ST_SENTINEL; sp++;
                        cur = ei; r = r2;
                        make_rayf(r, f);
                        node = e.a;
                        descend = true;
                    } else if (e.kind == RTX_GEOM_LIST) {
                        for (int k = 0; k < e.b; k++) {
                            const int2 it = S.list_items[e.a + k];
                            B.test_prim(S, it.x, it.y, r2, tmin, ei, e.rank, k, k, tcp);
                        }
                    } else {
                        B.test_prim(S, e.kind, e.index, r2, tmin, ei, e.rank, 0, 0, tcp);
                    }
                    if (!descend) {
                        if (Policy::ANY_HIT && B.have) node = RTX_ST_DONE;
                        else POP();
                    }
                }
            }
        } else {
            // ---- RETIRE + REFILL (warp-collective) ----------------------------------------------------------------
            const bool fin = node == RTX_ST_DONE;
            {
                RayD rw = r;
                double tmax_unused;
                if (fin && cur >= 0) P.load(job, rw, tmax_unused);  // an any-hit query may end inside an instance
                P.retire(job, fin, rw, B);
            }
            if (fin) node = RTX_ST_IDLE;
            if (!exhausted) {
                const unsigned want = __ballot_sync(FULL, node == RTX_ST_IDLE);
                const int cnt = __popc(want);
                int base = 0;
                if (lane == 0) base = atomicAdd(cursor, cnt);
                base = __shfl_sync(FULL, base, 0);
                const int my = base + __popc(want & ((1u << lane) - 1u));
                if (node == RTX_ST_IDLE && my < njobs) {
                    double tmax;
                    job = my;
                    P.load(job, r, tmax);
                    B.reset(tmax);
                    cur = -1; sp = 0;
                    // entries with unbounded geometry (infinite Plane, rt/plane.go:17) are tested for every ray
                    for (int k = 0; k < S.n_unbounded; k++) {
                        const int ei = S.unbounded[k];
                        const DEntry e = S.entries[ei];
                        RayD ro = r;
                        xform_ray(S, ei, e, ro);
                        B.test_prim(S, e.kind, e.index, ro, tmin, ei, e.rank, 0, 0, tcp);
                    }
                    if ((Policy::ANY_HIT && B.have) || S.tlas_root < 0) node = RTX_ST_DONE;
                    else { node = S.tlas_root; make_rayf(r, f); }
                }
                if (base + cnt >= njobs) exhausted = true;
            }
        }
    }
#undef POP
#undef STK
}
