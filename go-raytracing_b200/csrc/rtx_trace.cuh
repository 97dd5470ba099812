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

#ifndef RTX_TRACE_THREADS
#define RTX_TRACE_THREADS 128
#endif
#ifndef RTX_TRACE_SLOTS
#define RTX_TRACE_SLOTS 224   /* ray slots per 128-thread block (shared-memory ray pool, see trace_persistent); multiple of 32, <= 256 (A/B on cornell-lucy: 256: 1370, 224: 1401 Mrays/s: a little more L1) */
#endif
#ifndef RTX_TRACE_BLOCKS
#define RTX_TRACE_BLOCKS 4   /* resident blocks per SM the trace kernels are compiled for (caps registers at 65536 / (128 * blocks)) */
#endif
// The lean kernel variants (RTX_FV_*: compiled for one scene vocabulary) need 76-96 registers instead of 128, so six blocks stay resident
// (24 warps per SM instead of 16) with 160 slots each. cornell-lucy 64 spp: all-features kernels 214.5 ms; lean, 4 blocks x 224 slots 204.6;
// 5 x 192 196.6; 5 x 160 199.1; 6 x 160 194.3 ms. random: 14.7 / 13.9 / 13.4 / 13.3 / 13.0 ms.
#ifndef RTX_TRACE_SLOTS_LEAN
#define RTX_TRACE_SLOTS_LEAN 160
#endif
#ifndef RTX_TRACE_BLOCKS_LEAN
#define RTX_TRACE_BLOCKS_LEAN 6
#endif
#ifndef RTX_TRACE_BLOCKS_LEAN_NT
#define RTX_TRACE_BLOCKS_LEAN_NT 7   /* lean variants whose slots do not carry the ray time (50 words): 7 x 160 slots fit the 228 KB of an SM */
#endif
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
//   void   prefetch(int job) const;                         trace_flat only: hint that `job` is loaded next (may do nothing)
//
// Ray pool. Every lane owns K ray slots whose whole state lives in shared memory (slot = k * 128 + thread: each field
// is an array over slots, so a warp touching "its" slots of one k is bank-conflict free). Registers only hold a ray while
// one phase runs on it. With K > 1 a warp schedules 32 K rays onto 32 lanes: the phase that wins the vote finds a
// ready ray in far more lanes than with one ray per lane. The first RTX_SMEM_STACK stack entries of a slot are in shared
// memory, deeper ones spill to a global scratch column (rare: the stack seldom exceeds a dozen entries).
#ifndef RTX_SMEM_STACK
#define RTX_SMEM_STACK 16
#endif
#ifndef RTX_N_STEPS
#define RTX_N_STEPS 4   /* node levels per NODE round (A/B on cornell-lucy: 1: 1041, 2: 1062, 3: 1090, 4: 1080 Mrays/s; with the flat top level — no TLAS nodes in the loop — 64 spp take 185.4 / 182.4 / 180.9 ms at 2 / 3 / 4) */
#endif
#ifndef RTX_T_STEPS
#define RTX_T_STEPS 4   /* triangles per TRI round (1: 1282, 2: 1303-1332, 4: 1373 Mrays/s) */
#endif
#ifndef RTX_PARTNERS
#define RTX_PARTNERS 0   /* partner columns a lane may claim from when its own column has no ready slot of the voted phase (0..3) */
#endif
#ifndef RTX_SKIP_LAST_SENTINEL
#define RTX_SKIP_LAST_SENTINEL 1
#endif
#ifndef RTX_N_FAST
#define RTX_N_FAST 0   /* lanes with a NODE-ready slot from which the NODE phase is taken without a vote (0 = always vote) */
#endif
#ifndef RTX_ANYHIT_SORT
#define RTX_ANYHIT_SORT 0   /* cornell-lucy k_connect: 2646 (sorted) -> 2689 Mrays/s */
#endif
#ifndef RTX_E_BARE_FAST
#define RTX_E_BARE_FAST 2   /* ENTRY phase fast path: 0 off, 1 bare quads only, 2 every bare primitive */
#endif
#ifndef RTX_E_ROOT_STEP
#define RTX_E_ROOT_STEP 1   /* ENTRY of a mesh instance tests the BLAS root's children before committing to the instance */
#endif
#ifndef RTX_TLAS_FLAT_MAX
#define RTX_TLAS_FLAT_MAX 16   /* bounded world entries up to which a mesh world's top level is a sorted list built at refill (0 = always a hierarchy); at most RTX_SMEM_STACK */
#endif
// Stack code of a top-level entry in flat-TLAS mode: bit 31 (leaf class), bits 30..8 = the upper 23 bits of the float32 entry distance
// (truncated, i.e. rounded DOWN for the positive distances that occur: still a lower bound), bits 7..0 = world entry index. The distance
// is at least tmin > 0, so bits 30..8 are never all zero and the code cannot collide with RTX_ST_SENTINEL / DONE / IDLE. Codes of one
// ray order like their distances when compared as unsigned integers.
#define RTX_TLAS_CODE(tn, ei) ((int)(0x80000000u | ((unsigned)__float_as_int(tn) & 0x7fffff00u) | (unsigned)(ei)))
#define RTX_TLAS_CODE_DIST(code) (__int_as_float((code) & 0x7fffff00))
#define RTX_TLAS_CODE_ENTRY(code) ((code) & 0xff)
#ifndef RTX_EARLY_TICKET
#define RTX_EARLY_TICKET 1
#endif
#ifndef RTX_PREFETCH_DIST
#define RTX_PREFETCH_DIST 8192   /* cornell-lucy 64 spp: 169.0 ms without, 168.3 / 168.4 / 169.7 ms at 8 K / 32 K / 128 K */
#endif
#ifndef RTX_TRI_PRETEST
#define RTX_TRI_PRETEST 1   /* TRI phase: conservative float32 test from the 48-byte record first, the 96-byte float64 record only for survivors
                               (SURVEY Appendix C layout). Bit-exact (the level-1 suite runs with it), and on cornell-lucy it takes the float64 triangle
                               tests from 3.26 to 0.46 per ray (86 % of the tested triangles are rejected in float32). What it buys in time is
                               within the noise of code-layout effects: 64 spp 169.5 ms with, 170.5 ms with the pre-test compiled out of the same
                               two-pass loop, 168.6 ms with the old one-pass loop — the kernel is bound by rounds in flight, not by what a TRI
                               round costs. Kept on: less float64 work and less L2 traffic at the same speed. Option tri_pretest = 0 (at the next
                               rtx_scene_upload) skips the records and the test. */
#endif
#ifndef RTX_SORT_FULL
#define RTX_SORT_FULL 1   /* NODE phase: 1 = the four children of a node fully ordered front to back, 0 = only the nearest in front (3 of the 5 compare-exchanges) */
#endif
#define RTX_PH_N 0
#define RTX_PH_T 1
#define RTX_PH_E 2
#define RTX_PH_R 3
#define RTX_PH_NONE 4   /* parked: idle slot after the job queue ran dry */
// 32-bit words of shared memory per ray slot: float32 ray 9, box offsets | stack pointer 1, ft 1, node / cur / job 3, current-space ray 12 (+ 2 for the ray
// time, which only moving spheres and Volumes read: vocabularies without them do not carry it), best t 2, best ids 6, stack
#define RTX_FEAT_HAS_TIME(FEAT) (((FEAT) & (RTX_F_SPHERE | RTX_F_COMPLEX)) != 0)
#define RTX_SLOT_WORDS_T(TIME) (9 + 1 + 1 + 3 + 12 + ((TIME) ? 2 : 0) + 2 + 6 + RTX_SMEM_STACK)
#define RTX_SLOT_WORDS_OF(FEAT) RTX_SLOT_WORDS_T(RTX_FEAT_HAS_TIME(FEAT))
#define RTX_SLOT_WORDS RTX_SLOT_WORDS_T(true)
// RTX_CHECKED (make checked -> librtx_b200_checked.so): the build that stands in for compute-sanitizer racecheck / memcheck, which is closed
// on the GPU pool this repo is measured on. The slot hand-over protocol of trace_persistent (claim with a shared-memory CAS, work on the
// slot, __threadfence_block, publish with an atomic XOR) and every index the traversal dereferences are asserted at run time; violations are
// COUNTED, per kind, in g_rtx_check and come back through rtx_stats.checked_violations (tests/test_checked_build.py demands zero):
//   0 a slot claimed while another warp owns it      1 a slot published by a warp that does not own it    2 phase nibble not BUSY at publish
//   3 stack pointer out of range                     4 node index out of range                            5 triangle index out of range
//   6 world entry index out of range                 7 job index out of range
#ifdef RTX_CHECKED
__device__ unsigned long long g_rtx_check[8];
#define RTX_CHECK(cond, k) do { if (!(cond)) atomicAdd(&g_rtx_check[k], 1ull); } while (0)
#define RTX_POOL_EXTRA_BYTES (256 + 4 * 256)                            /* column states + flags + one owner word per slot (<= 256 slots) */
#else
#define RTX_CHECK(cond, k) do { } while (0)
#define RTX_POOL_EXTRA_BYTES 256                                        /* column states + flags */
#endif
#define RTX_PH_BUSY 5   /* claimed by a warp for the current round */

template <int NSLOTS, bool TIME = true>
struct TracePool {
    static constexpr int NS = NSLOTS;
    static_assert(NSLOTS % 32 == 0 && NSLOTS <= 256, "slots per block: a multiple of 32, at most 8 per bank column");
    float* f;      // [9][NS]  ix iy iz cnx cny cnz cfx cfy cfz
    int* off;      // [NS]     offx | offy << 8 | offz << 16
    float* ft;     // [NS]
    int *node, *cur, *job;
    unsigned char* spb;  // stack pointer of slot s: byte 3 of off[s] (spb[4 * s + 3]); the NODE phase reads it with the offsets, nobody else pays a word for it
    double* r;     // [7][NS]  current-space ray ox oy oz dx dy dz, and (TIME) the ray time
    double* bt;    // [NS]
    int *be, *bk, *bp, *bi, *bre, *brp;  // best: entry, kind | have << 8, prim, item, rank_e, rank_p
    int* stack;    // [RTX_SMEM_STACK][NS]
    unsigned* col; // [32]  packed phase nibbles of the NS/32 slots of each bank column (block-shared scheduling state)
    int* flags;    // [32]  flags[0]: job queue ran dry
    int* owner;    // RTX_CHECKED only: [NS] 0 = free, else 1 + warp of the block that claimed the slot
    __device__ __forceinline__ explicit TracePool(unsigned char* base) {
        double* d = reinterpret_cast<double*>(base);
        r = d; d += (TIME ? 7 : 6) * NS;
        bt = d; d += NS;
        float* w = reinterpret_cast<float*>(d);
        f = w; w += 9 * NS;
        ft = w; w += NS;
        int* q = reinterpret_cast<int*>(w);
        off = q; q += NS; node = q; q += NS; cur = q; q += NS; job = q; q += NS;
        spb = reinterpret_cast<unsigned char*>(off);
        be = q; q += NS; bk = q; q += NS; bp = q; q += NS; bi = q; q += NS; bre = q; q += NS; brp = q; q += NS;
        stack = q; q += RTX_SMEM_STACK * NS;
        col = reinterpret_cast<unsigned*>(q); q += 32;
        flags = q; q += 32;
        owner = q;
    }
    __device__ __forceinline__ void load_rayf(int s, RayF& x) const {
        x.ix = f[s]; x.iy = f[NS + s]; x.iz = f[2 * NS + s]; x.cnx = f[3 * NS + s]; x.cny = f[4 * NS + s]; x.cnz = f[5 * NS + s];
        x.cfx = f[6 * NS + s]; x.cfy = f[7 * NS + s]; x.cfz = f[8 * NS + s];
        const int o = off[s];
        x.offx = o & 0xff; x.offy = (o >> 8) & 0xff; x.offz = (o >> 16) & 0xff;
    }
    __device__ __forceinline__ void store_rayf(int s, const RayF& x) const {
        f[s] = x.ix; f[NS + s] = x.iy; f[2 * NS + s] = x.iz; f[3 * NS + s] = x.cnx; f[4 * NS + s] = x.cny; f[5 * NS + s] = x.cnz;
        f[6 * NS + s] = x.cfx; f[7 * NS + s] = x.cfy; f[8 * NS + s] = x.cfz;
        off[s] = x.offx | (x.offy << 8) | (x.offz << 16);   // clears the stack-pointer byte: every caller stores sp after the ray
    }
    __device__ __forceinline__ void load_ray(int s, RayD& x) const {
        x.ox = r[s]; x.oy = r[NS + s]; x.oz = r[2 * NS + s]; x.dx = r[3 * NS + s]; x.dy = r[4 * NS + s]; x.dz = r[5 * NS + s]; x.tm = TIME ? r[(TIME ? 6 : 0) * NS + s] : 0.0;
    }
    __device__ __forceinline__ void store_ray(int s, const RayD& x, bool with_time) const {
        r[s] = x.ox; r[NS + s] = x.oy; r[2 * NS + s] = x.oz; r[3 * NS + s] = x.dx; r[4 * NS + s] = x.dy; r[5 * NS + s] = x.dz;
        if (TIME && with_time) r[(TIME ? 6 : 0) * NS + s] = x.tm;
    }
    __device__ __forceinline__ void load_best(int s, Best& b) const {
        b.t = bt[s]; b.ft = ft[s]; b.entry = be[s];
        const int kh = bk[s];
        b.kind = (int)(signed char)(kh & 0xff); b.have = (kh >> 8) & 1;
        b.prim = bp[s]; b.item = bi[s]; b.rank_e = bre[s]; b.rank_p = brp[s];
    }
    __device__ __forceinline__ void store_best(int s, const Best& b) const {
        bt[s] = b.t; ft[s] = b.ft; be[s] = b.entry; bk[s] = (b.kind & 0xff) | (b.have ? 0x100 : 0);
        bp[s] = b.prim; bi[s] = b.item; bre[s] = b.rank_e; brp[s] = b.rank_p;
    }
};

// A world entry that is not a mesh instance — a primitive, a Box list, a Volume — against the world ray `r` (registers
// only): the reference's Hit of that object with the interval [tmin, B.t] (rt/hittable_list.go:31-45, rt/volume.go:34-79,
// the wrappers of rt/transform.go). `vr` is only read for volumes.
__device__ __forceinline__ void entry_core(const DevScene& S, int ei, const DEntry& e, const RayD& r, Best& B, double tmin, const VolumeRng& vr, TraceCounters* tcp) {
    RayD r2 = r;
    xform_ray(S, ei, e, r2);
    if (e.volume >= 0) {
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
    } else if (e.kind == RTX_GEOM_LIST) {
        for (int q = 0; q < e.b; q++) {
            const int2 it = S.list_items[e.a + q];
            B.test_prim(S, it.x, it.y, r2, tmin, ei, e.rank, q, q, tcp);
        }
    } else {
        B.test_prim(S, e.kind, e.index, r2, tmin, ei, e.rank, 0, 0, tcp);
    }
}

// The same for the ray of pool slot `s` of the persistent kernels: the rare, bulky part of the ENTRY phase (float64 sphere /
// quad / plane tests, the Volume free-flight with its log). Measured both ways on B200: inlined 1039 Mrays/s, out of line
// (-DRTX_ENTRY_OOL) 958 on cornell-lucy. State travels through the shared-memory pool; Sp points at the kernel's
// __grid_constant__ parameter.
template <int NSLOTS, bool TIME>
#ifndef RTX_ENTRY_OOL
__device__ __forceinline__
#else
__device__ __noinline__
#endif
bool entry_other(const DevScene* Sp, unsigned char* smem, int s, int ei, double tmin, uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                 uint32_t c2, bool transparent, TraceCounters* tcp) {
    const DevScene& S = *Sp;
    const TracePool<NSLOTS, TIME> T(smem);
    const DEntry e = S.entries[ei];
    RayD r;
    T.load_ray(s, r);
    Best B;
    T.load_best(s, B);
    VolumeRng vr; vr.k0 = k0; vr.k1 = k1; vr.c0 = c0; vr.c1 = c1; vr.c2 = c2; vr.transparent = transparent;
    entry_core(S, ei, e, r, B, tmin, vr, tcp);
    T.store_best(s, B);
    return B.have;
}

// Scenes whose whole world is a handful of primitives (no mesh): the TLAS would be a single leaf, so the query is the
// reference's HittableList.Hit loop itself — one thread per ray, every entry in turn, no pool, no stack, no divergence
// between lanes beyond the tests' own early-outs. Results are identical to the hierarchy's (closest hits do not depend on
// the order of the tests; exact ties are resolved by rank as everywhere).
// a policy whose retire is collective over groups of FOUR warps (shade_commit's quad tickets) says so with QUAD_TICKETS = true
template <class P, class = void> struct policy_quad_tickets { static constexpr bool value = false; };
template <class P> struct policy_quad_tickets<P, decltype((void)P::QUAD_TICKETS)> { static constexpr bool value = P::QUAD_TICKETS; };
template <class Policy, bool COUNT, unsigned FEAT = RTX_F_ALL>
__device__ __forceinline__ void trace_flat(const DevScene& S, Policy& P, int njobs, TraceCounters& tc) {
    TraceCounters* const tcp = COUNT ? &tc : nullptr;
    const double tmin = P.tmin();
    // whole warps (or whole groups of four warps) reach the collective retire
    const int nrounded = policy_quad_tickets<Policy>::value ? (njobs + 127) & ~127 : (njobs + 31) & ~31;
    for (int job = blockIdx.x * blockDim.x + threadIdx.x; job < nrounded; job += gridDim.x * blockDim.x) {
        const bool valid = job < njobs;
        RayD r;
        Best B;
        r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = r.tm = 0;
        B.reset(0);
        if (valid) {
            double tmax;
            P.load(job, r, tmax);
            B.reset(tmax);
            // bare primitives (the usual case: spheres, walls, the ground plane): the test is inlined and selected by a branch
            // every lane takes alike; no wrapper chain, no entry record, no out-of-line dispatch
            for (int k = 0; k < S.n_flat_simple; k++) {
                const int4 fe = __ldg(S.flat_simple + k);   // kind, primitive, entry, rank
                const bool incl = B.have;                   // an open-interval primitive may still tie with the current best (Best::tmax_for)
                double t;
                if ((FEAT & RTX_F_SPHERE) && fe.x == RTX_GEOM_SPHERE) {
                    if (COUNT) tc.spheres++;
                    t = isect_sphere_incl(S.spheres + 8 * (size_t)fe.y, r, tmin, B.t, incl);
                } else if ((FEAT & RTX_F_QUAD) && fe.x == RTX_GEOM_QUAD) {
                    if (COUNT) tc.quads++;
                    t = isect_quad(S.quads + 16 * (size_t)fe.y, r, tmin, B.t, nullptr);
                } else if ((FEAT & RTX_F_PLANE) && fe.x == RTX_GEOM_PLANE) {
                    if (COUNT) tc.planes++;
                    t = isect_plane(S.planes + 8 * (size_t)fe.y, r);
                    if (!(tmin < t && (t < B.t || (incl && t == B.t)))) t = RTX_NAN_D;
                } else if ((FEAT & RTX_F_OTHER_PRIM) && fe.x == RTX_GEOM_CIRCLE) {
                    if (COUNT) tc.quads++;
                    t = isect_circle(S.circles + 8 * (size_t)fe.y, r, tmin, B.t);
                } else if (FEAT & RTX_F_OTHER_PRIM) {
                    if (COUNT) tc.tris++;
                    t = isect_tri(S.tris + RTX_TRI_D * (size_t)fe.y, r, nullptr);
                    if (!(tmin <= t && t <= B.t)) t = RTX_NAN_D;
                } else t = RTX_NAN_D;   // a kind outside the variant's vocabulary: unreachable for a scene the mask covers
                B.offer(t, fe.z, fe.w, fe.x, fe.y, 0, 0);
                if (Policy::ANY_HIT && B.have) break;
            }
            if ((FEAT & RTX_F_COMPLEX) && !(Policy::ANY_HIT && B.have))
                for (int k = 0; k < S.n_flat_complex; k++) {
                    const int ei = S.flat_complex[k];
                    const DEntry e = S.entries[ei];
                    VolumeRng vr = {0, 0, 0, 0, 0, true};
                    if (e.volume >= 0) vr = P.volume_rng(job);
                    entry_core(S, ei, e, r, B, tmin, vr, tcp);
                    if (Policy::ANY_HIT && B.have) break;
                }
        }
#ifndef RTX_FLAT_PREFETCH
#define RTX_FLAT_PREFETCH 0   /* measured: hdri-test 6933 Mpaths/s without, 6672 with; cornell-glossy 3146 / 3118 */
#endif
        // (off) the record of this thread's next job requested before the retire, whose warp-aggregated append waits for an atomic's return
        if (RTX_FLAT_PREFETCH && job + (int)(gridDim.x * blockDim.x) < njobs) P.prefetch(job + (int)(gridDim.x * blockDim.x));
        P.retire(valid ? job : -1, valid, r, B);
    }
}

// The drain of a pass (a few thousand to a few hundred thousand rays per iteration) is latency-bound in the persistent kernel: its rays
// are spread one per lane over the pool and every round serves a handful of lanes, so a launch cannot finish faster than ~60 rounds
// whatever the ray count (measured: 230 us for 100 K rays). For such batches the plain shape wins: ONE THREAD PER RAY, the whole query
// in registers and a local-memory stack, no pool, no votes — a launch then lasts as long as its longest ray (tens of microseconds).
// Same primitive tests, same tie rules (Best::offer), same two-level traversal: hit records are bit-identical to trace_persistent's
// (test_small_batch_kernels_are_result_neutral; closest hits do not depend on the order of the tests).
template <class Policy, bool COUNT, unsigned FEAT = RTX_F_ALL>
__device__ __forceinline__ void trace_simple(const DevScene& S, Policy& P, int njobs, TraceCounters& tc) {
    TraceCounters* const tcp = COUNT ? &tc : nullptr;
    const double tmin = P.tmin();
    const float ftmin = __double2float_rd(tmin);
    const float INF = __int_as_float(0x7f800000);
    const int nrounded = (njobs + 31) & ~31;   // whole warps reach the warp-collective retire
    for (int job = blockIdx.x * blockDim.x + threadIdx.x; job < nrounded; job += gridDim.x * blockDim.x) {
        const bool valid = job < njobs;
        RayD rw;
        Best B;
        rw.ox = rw.oy = rw.oz = rw.dx = rw.dy = rw.dz = rw.tm = 0;
        B.reset(0);
        if (valid) {
            double tmax;
            P.load(job, rw, tmax);
            B.reset(tmax);
            for (int q = 0; q < S.n_unbounded; q++) {   // entries tested for every ray (infinite planes; pre-tested bare primitives)
                const int ei = S.unbounded[q];
                const DEntry e = S.entries[ei];
                RayD ro = rw;
                if (FEAT & (RTX_F_COMPLEX | RTX_F_XFORM)) xform_ray(S, ei, e, ro);
                B.test_prim(S, e.kind, e.index, ro, tmin, ei, e.rank, 0, 0, tcp);
                if (Policy::ANY_HIT && B.have) break;
            }
            int stk[RTX_STACK_SIZE];
            int sp = 0, cur = -1;
            RayD r = rw;     // current-space ray (world, or the object space of instance `cur`)
            RayF f;
            make_rayf(r, f);
            if (S.tlas_root >= 0 && !(Policy::ANY_HIT && B.have)) stk[sp++] = S.tlas_root;
            while (sp > 0) {
                const int node = stk[--sp];
                if (node >= 0) {
                    float d[4]; int ch[4];
                    if (COUNT) tc.nodes++;
                    node_test(S.nodes, node, f, ftmin, B.ft, d, ch);
#define RTX_CSWAP(i, j) if (d[j] < d[i]) { float td = d[i]; d[i] = d[j]; d[j] = td; int tcx = ch[i]; ch[i] = ch[j]; ch[j] = tcx; }
                    if (!Policy::ANY_HIT) { RTX_CSWAP(0, 1) RTX_CSWAP(2, 3) RTX_CSWAP(0, 2) RTX_CSWAP(1, 3) RTX_CSWAP(1, 2) }
#undef RTX_CSWAP
                    if (d[3] < INF) stk[sp++] = ch[3];
                    if (d[2] < INF) stk[sp++] = ch[2];
                    if (d[1] < INF) stk[sp++] = ch[1];
                    if (d[0] < INF) stk[sp++] = ch[0];
                } else if ((FEAT & RTX_F_MESH) && node == RTX_ST_SENTINEL) {   // the instance is exhausted: back to the world ray
                    r = rw; make_rayf(r, f); cur = -1;
                } else if ((FEAT & RTX_F_MESH) && cur >= 0) {                  // a BLAS leaf
                    const int code = ~node;
                    const int first = code >> 3, cnt = (code & 7) + 1;
                    for (int k = 0; k < cnt; k++) {
                        const int ti = first + k;
                        if (COUNT) tc.tris++;
                        const double t = isect_tri(S.tris + RTX_TRI_D * (size_t)ti, r, nullptr);
                        if (tmin <= t && t <= B.t) {
                            const int4 info = __ldg(S.tri_info + ti);
                            B.offer(t, cur, S.entries[cur].rank, RTX_GEOM_TRIANGLE, ti, info.x, info.z);
                        }
                    }
                    if (Policy::ANY_HIT && B.have) break;
                } else {                                                          // a world entry
                    const int ei = ~node;
                    const DEntry e = S.entries[ei];
                    if ((FEAT & RTX_F_MESH) && e.volume < 0 && e.kind == RTX_GEOM_MESH) {
                        RayD r2 = rw;
                        if (FEAT & (RTX_F_XFORM | RTX_F_COMPLEX)) xform_ray(S, ei, e, r2);
                        r = r2; make_rayf(r, f); cur = ei;
                        stk[sp++] = RTX_ST_SENTINEL;
                        stk[sp++] = e.a;
                    } else {
                        VolumeRng vr = {0, 0, 0, 0, 0, true};
                        if ((FEAT & RTX_F_COMPLEX) && e.volume >= 0) vr = P.volume_rng(job);
                        if ((FEAT & RTX_F_COMPLEX) || e.kind != RTX_GEOM_LIST) entry_core(S, ei, e, rw, B, tmin, vr, tcp);
                        if (Policy::ANY_HIT && B.have) break;
                    }
                }
            }
        }
        P.retire(valid ? job : -1, valid, rw, B);
    }
}

template <class Policy, bool COUNT, int NSLOTS, unsigned FEAT = RTX_F_ALL>
__device__ __forceinline__ void trace_persistent(const DevScene& S, Policy& P, int* cursor, int njobs, TraceCounters& tc, int* spill, unsigned char* smem) {
    typedef TracePool<NSLOTS, RTX_FEAT_HAS_TIME(FEAT)> Pool_;
    constexpr int NS = Pool_::NS;
    const Pool_ T(smem);
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const double tmin = P.tmin();
    const float ftmin = __double2float_rd(tmin);
    const float INF = __int_as_float(0x7f800000);
    TraceCounters* const tcp = COUNT ? &tc : nullptr;
    const bool flatTlas = (FEAT & RTX_F_MESH) && RTX_TLAS_FLAT_MAX > 0 && S.n_tlas_flat > 0;   // uniform: the top level is a sorted list on each ray's stack
    const size_t spill_stride = (size_t)gridDim.x * NS;
    int* const spill_col = spill + (size_t)blockIdx.x * NS;

    // Block-shared scheduling. A slot's phase is a nibble of its bank column's state word (column = slot % 32, NS / 32
    // slots per column). Lane l of ANY warp of the block may run a phase on any slot of column l — shared-memory accesses stay
    // conflict-free — and claims it with a compare-and-swap (phase -> BUSY). So each lane chooses among 4K candidates
    // instead of its own K, and the four warps of a block usually run different phases at the same time.
    constexpr int NCOL = NS / 32;
    static_assert(NCOL <= 8, "one 32-bit state word per column holds at most 8 slots");
    constexpr unsigned ALL_PARKED = 0x44444444u;
    const unsigned warp = threadIdx.x >> 5;
    volatile unsigned* const colstate = T.col;
    volatile int* const dry = T.flags;
    if (threadIdx.x < 32) {
        unsigned w0 = ALL_PARKED;
        for (int j = 0; j < NCOL; j++) w0 = (w0 & ~(0xfu << (4 * j))) | ((unsigned)RTX_PH_R << (4 * j));
        T.col[threadIdx.x] = w0;
        T.flags[threadIdx.x] = 0;
    }
    for (int i = threadIdx.x; i < NS; i += RTX_TRACE_THREADS) T.node[i] = RTX_ST_IDLE;
#ifdef RTX_CHECKED
    for (int i = threadIdx.x; i < NS; i += RTX_TRACE_THREADS) T.owner[i] = 0;
#endif
    __syncthreads();
    unsigned round = warp;

#define RTX_PUSH(v) do { if (sp < RTX_SMEM_STACK) T.stack[sp * NS + s] = (v); else spill_col[(size_t)(sp - RTX_SMEM_STACK) * spill_stride + s] = (v); sp++; } while (0)
#define RTX_POP() do { if (sp > 0) { sp--; node = sp < RTX_SMEM_STACK ? T.stack[sp * NS + s] : spill_col[(size_t)(sp - RTX_SMEM_STACK) * spill_stride + s]; } \
                       else node = RTX_ST_DONE; } while (0)
// pop at the top level: in flat-TLAS mode the entries are sorted nearest first and carry their entry distance, so when the next one lies
// beyond the best hit the query is finished (ft: float32 upper bound of the best t, re-read because the phase may just have improved it)
#define RTX_POP_TOP() do { RTX_POP(); if (flatTlas && node != RTX_ST_DONE && RTX_TLAS_CODE_DIST(node) > T.ft[s]) { node = RTX_ST_DONE; sp = 0; } } while (0)
#define RTX_CLASSIFY(nd, inst) ((nd) >= 0 ? RTX_PH_N : (nd) == RTX_ST_DONE ? RTX_PH_R : (nd) == RTX_ST_SENTINEL ? RTX_PH_E : (inst) ? RTX_PH_T : RTX_PH_E)

#define RTX_HASZERO_NIB(x) ((((x) - 0x11111111u) & ~(x)) & 0x88888888u)   /* lowest set bit marks the lowest zero nibble exactly */
    for (;;) {
        // ---- vote: one REDUX over packed per-phase counts of lanes whose column holds a ready slot ------------------------
        unsigned w = colstate[lane];
        int phase;
        // fast path: when at least RTX_N_FAST lanes have a NODE-ready slot the NODE phase runs without a full vote (one ballot
        // instead of four zero-nibble tests, a REDUX and the arg-max). The rarer phases then wait until NODE runs short of
        // lanes, which also lets them collect more lanes per round.
        if (RTX_N_FAST > 0 && __popc(__ballot_sync(FULL, RTX_HASZERO_NIB(w) != 0u)) >= RTX_N_FAST) {
            phase = RTX_PH_N;
            round++;
        } else {
            unsigned present = 0;
#pragma unroll
            for (unsigned X = 0; X < 4; X++) present |= RTX_HASZERO_NIB(w ^ (X * 0x11111111u)) ? (1u << (8 * X)) : 0u;
            const unsigned c = __reduce_add_sync(FULL, present);
            if (c == 0) {
                if (__all_sync(FULL, w == ALL_PARKED)) break;   // every slot of the block is parked: queue dry, all rays retired
                __nanosleep(100);                                // other warps hold the remaining slots (BUSY): wait for them
                continue;
            }
            // the phase with the most ready lanes wins; equal counts are broken by a priority that rotates with the round and
            // differs between the warps of a block, so they spread over the phases and no ray starves
            round++;
            const int cN = ((c & 0xff) << 2) | (round & 3), cT = (((c >> 8) & 0xff) << 2) | ((round + 1) & 3),
                      cE = (((c >> 16) & 0xff) << 2) | ((round + 2) & 3), cR = (((c >> 24) & 0xff) << 2) | ((round + 3) & 3);
            phase = (cN >= cT && cN >= cE && cN >= cR) ? RTX_PH_N : (cT >= cE && cT >= cR) ? RTX_PH_T : (cE >= cR) ? RTX_PH_E : RTX_PH_R;
        }
        // claim one ready slot of my column (search start rotates so that no slot index is favoured); a lane whose own column
        // has none tries its RTX_PARTNERS partner columns (lane ^ 16, ^ 8, ^ 24): at worst a 2-way bank conflict with the
        // partner lane, against an idle lane for the whole round
        int j = -1;
        unsigned col = lane;
        {
            const unsigned rot = (round % NCOL) * 4u;
            const unsigned pat = (unsigned)phase * 0x11111111u;
#pragma unroll
            for (int attempt = 0; attempt <= RTX_PARTNERS; attempt++) {
                if (attempt > 0) {
                    if (j >= 0) break;
                    col = lane ^ (attempt == 1 ? 16u : attempt == 2 ? 8u : 24u);
                    w = colstate[col];
                }
                for (;;) {
                    const unsigned wr = __funnelshift_r(w, w, rot);
                    const unsigned hz = RTX_HASZERO_NIB(wr ^ pat);
                    if (!hz) break;
                    const int jj = (int)((((unsigned)(__ffs(hz) - 1) >> 2) + (rot >> 2)) & 7u);
                    const unsigned neww = w ^ ((unsigned)(phase ^ RTX_PH_BUSY) << (4 * jj));
                    const unsigned old = atomicCAS(const_cast<unsigned*>(colstate) + col, w, neww);
                    if (old == w) { j = jj; break; }
                    w = old;
                }
            }
        }
        const bool mine = j >= 0;
        const int s = (int)col + 32 * (mine ? j : 0);
        if (mine) __threadfence_block();   // see the slot as its previous owner left it
#ifdef RTX_CHECKED
        if (mine) { const int prev = atomicExch(&T.owner[s], 1 + (int)warp); RTX_CHECK(prev == 0, 0); RTX_CHECK(phase == RTX_PH_R || (int)T.spb[4 * s + 3] <= RTX_STACK_SIZE, 3); }   // (an idle slot's stack pointer is not initialised)
#endif
        int newst = -1;  // phase of the claimed slot after this round

        if (phase == RTX_PH_N) {
            // ---- NODE: one 4-wide node per lane -----------------------------------------------------------------------
            if (mine) {
                int node = T.node[s], sp = T.spb[4 * s + 3];
                const bool inst = T.cur[s] >= 0;
                RayF f;
                T.load_rayf(s, f);
                const float ftmax = T.ft[s];
                // up to RTX_N_STEPS levels per round: every extra level saves a vote and a state round trip, at the price of
                // idle lanes once their next node is a leaf
#pragma unroll 1
                for (int step = 0; step < RTX_N_STEPS && node >= 0; step++) {
                    float d[4]; int ch[4];
                    if (COUNT) tc.nodes++;
                    RTX_CHECK(node < S.n_nodes, 4);
                    node_test(S.nodes, node, f, ftmin, ftmax, d, ch);
#define RTX_CSWAP(i, j) if (d[j] < d[i]) { float td = d[i]; d[i] = d[j]; d[j] = td; int tcx = ch[i]; ch[i] = ch[j]; ch[j] = tcx; }
                    // closest-hit queries visit the children front to back; an any-hit query only needs SOME hit, so the order is
                    // irrelevant to the result (RTX_ANYHIT_SORT = 0 drops the sorting network there)
                    if (RTX_ANYHIT_SORT || !Policy::ANY_HIT) { RTX_CSWAP(0, 1) RTX_CSWAP(2, 3) RTX_CSWAP(0, 2) if (RTX_SORT_FULL) { RTX_CSWAP(1, 3) RTX_CSWAP(1, 2) } }
#undef RTX_CSWAP
                    if (sp + 3 <= RTX_SMEM_STACK) {
                        int* const st = T.stack + s;
                        if (d[3] < INF) { st[sp * NS] = ch[3]; sp++; }
                        if (d[2] < INF) { st[sp * NS] = ch[2]; sp++; }
                        if (d[1] < INF) { st[sp * NS] = ch[1]; sp++; }
                        if (d[0] < INF) node = ch[0];
                        else if (sp > 0) { sp--; node = st[sp * NS]; }
                        else node = RTX_ST_DONE;
                    } else {
                        if (d[3] < INF) RTX_PUSH(ch[3]);
                        if (d[2] < INF) RTX_PUSH(ch[2]);
                        if (d[1] < INF) RTX_PUSH(ch[1]);
                        if (d[0] < INF) node = ch[0];
                        else RTX_POP();
                    }
                }
                // flat top level: the instance is exhausted and the nearest top-level entry still waiting lies beyond the best hit (they are
                // sorted, so all of them do): the query is finished without the round that would restore the world ray
                if ((FEAT & RTX_F_MESH) && flatTlas && node == RTX_ST_SENTINEL && (sp == 0 || RTX_TLAS_CODE_DIST(T.stack[(sp - 1) * NS + s]) > ftmax)) node = RTX_ST_DONE;
                T.node[s] = node; T.spb[4 * s + 3] = (unsigned char)sp;
                newst = RTX_CLASSIFY(node, inst);
            }
        } else if ((FEAT & RTX_F_MESH) && phase == RTX_PH_T) {
            // ---- TRI: one triangle of the pending BLAS leaf per lane ------------------------------------------------
            if (mine) {
                int node = T.node[s];
                RayD r;
                T.load_ray(s, r);
                double bt = T.bt[s];
                // float32 copy of the ray for the conservative pre-test (tri_pretest_reject): survivors go on to the float64 test
                const bool pre = RTX_TRI_PRETEST && S.tris32 != nullptr;
                const float ofx = (float)r.ox, ofy = (float)r.oy, ofz = (float)r.oz, dfx = (float)r.dx, dfy = (float)r.dy, dfz = (float)r.dz;
                const float Mo = fmaxf(fabsf(ofx), fmaxf(fabsf(ofy), fabsf(ofz))) * 1.0000002f, Md = fmaxf(fabsf(dfx), fmaxf(fabsf(dfy), fabsf(dfz))) * 1.0000002f;
                float btf = T.ft[s];
                // Up to RTX_T_STEPS triangles of the leaf per round (the ray is loaded once); lanes with shorter leaves idle. Two passes, so that
                // the float64 code runs as rarely as the pre-test allows: with one loop over the triangles the warp would execute the float64
                // test whenever ANY lane's triangle survived — 84 % of the steps at 12 lanes and a 14 % survival rate, i.e. nothing saved.
                //   pass 1 (pre): the float32 pre-test of every triangle of the step window -> a survivor mask per lane;
                //   pass 2: the float64 test of the survivors, in leaf order (the loop runs as often as the lane with the most survivors needs).
                const int code0 = ~node;
                const int ti0 = code0 >> 3, rem0 = code0 & 7;
                const int nwin = min(rem0 + 1, RTX_T_STEPS);   // triangles of this round: ti0 .. ti0 + nwin - 1
                unsigned surv = (1u << nwin) - 1u;
                if (pre) {
                    surv = 0;
#pragma unroll 1
                    for (int k = 0; k < nwin; k++) {
                        RTX_CHECK(ti0 + k >= 0 && ti0 + k < S.n_tris_total, 5);
                        if (!tri_pretest_reject(S.tris32 + 3 * (size_t)(ti0 + k), ofx, ofy, ofz, dfx, dfy, dfz, Mo, Md, ftmin, btf)) surv |= 1u << k;
                    }
                }
                if (COUNT) { tc.tris += nwin; if (pre) tc.spheres += __popc(surv); }   // (mesh worlds without spheres: the float64 confirmations ride in the sphere counter)
                bool have = false;
#pragma unroll 1
                while (surv) {
                    const int k = __ffs(surv) - 1;
                    surv &= surv - 1;
                    const int ti = ti0 + k;
                    RTX_CHECK(ti >= 0 && ti < S.n_tris_total, 5);
                    const double t = isect_tri(S.tris + RTX_TRI_D * (size_t)ti, r, nullptr);
                    if (tmin <= t && t <= bt) {
                        const int4 info = __ldg(S.tri_info + ti);
                        Best B;
                        T.load_best(s, B);
                        const int cur = T.cur[s];
                        B.offer(t, cur, S.entries[cur].rank, RTX_GEOM_TRIANGLE, ti, info.x, info.z);
                        T.store_best(s, B);
                        have = B.have;
                        bt = B.t;
                        if (Policy::ANY_HIT && have) break;
                    }
                }
                if (Policy::ANY_HIT && have) node = RTX_ST_DONE;
                else if (rem0 + 1 == nwin) {   // the leaf is finished
                    int sp = T.spb[4 * s + 3];
                    RTX_POP();
                    if (flatTlas && node == RTX_ST_SENTINEL && (sp == 0 || RTX_TLAS_CODE_DIST(T.stack[(sp - 1) * NS + s]) > T.ft[s])) node = RTX_ST_DONE;   // as in the NODE phase
                    T.spb[4 * s + 3] = (unsigned char)sp;
                } else node = ~(((ti0 + nwin) << 3) | (rem0 - nwin));
                T.node[s] = node;
                newst = RTX_CLASSIFY(node, true);
            }
        } else if (phase == RTX_PH_E) {
            // ---- ENTRY: a world entry (TLAS leaf), or the end of an instance -----------------------------------------
            if (mine) {
                int node = T.node[s], sp = T.spb[4 * s + 3];
                const int job = T.job[s];
                bool in_inst = false;
                if ((FEAT & RTX_F_MESH) && node == RTX_ST_SENTINEL) {
                    RayD r; RayF f;
                    double tmax_unused;
                    P.load(job, r, tmax_unused);
                    make_rayf(r, f);
                    T.store_ray(s, r, false); T.store_rayf(s, f);
                    T.cur[s] = -1;
                    RTX_POP_TOP();
                } else {
                    const int ei = flatTlas ? RTX_TLAS_CODE_ENTRY(node) : ~node;
                    RTX_CHECK(ei >= 0 && ei < S.n_entries, 6);
                    const DEntry e = S.entries[ei];
                    if ((FEAT & RTX_F_MESH) && e.volume < 0 && e.kind == RTX_GEOM_MESH) {
                        RayD r2; RayF f;
                        T.load_ray(s, r2);
                        if (FEAT & (RTX_F_XFORM | RTX_F_COMPLEX)) xform_ray(S, ei, e, r2);
                        make_rayf(r2, f);
#if RTX_E_ROOT_STEP
                        // The instance's world-space box (the TLAS leaf) is loose around a rotated statue: test the BLAS root's four
                        // children here, with the object-space ray already in registers. If none is hit the instance is never
                        // entered — the stored world ray is untouched, no sentinel, no NODE round, no round to come back — and
                        // otherwise the NODE phase starts one level down.
                        float d[4]; int ch[4];
                        if (COUNT) tc.nodes++;
                        node_test(S.nodes, e.a, f, ftmin, T.ft[s], d, ch);
#define RTX_CSWAP(i, j) if (d[j] < d[i]) { float td = d[i]; d[i] = d[j]; d[j] = td; int tcx = ch[i]; ch[i] = ch[j]; ch[j] = tcx; }
                        if (RTX_ANYHIT_SORT || !Policy::ANY_HIT) { RTX_CSWAP(0, 1) RTX_CSWAP(2, 3) RTX_CSWAP(0, 2) RTX_CSWAP(1, 3) RTX_CSWAP(1, 2) }
                        else { RTX_CSWAP(0, 1) RTX_CSWAP(2, 3) RTX_CSWAP(0, 2) }   // the minimum in front: d[0] < INF iff any child is hit
#undef RTX_CSWAP
                        if (!(d[0] < INF)) {
#ifdef RTX_DEBUG_ENTRY_COUNT
                            if (COUNT) tc.planes++;
#endif
                            RTX_POP_TOP();
                        } else {
#ifdef RTX_DEBUG_ENTRY_COUNT
                            if (COUNT) tc.spheres++;
#endif
                            // nothing left in the TLAS: the query ends inside the instance (retire re-reads the world ray), no way back needed
                            if (RTX_SKIP_LAST_SENTINEL == 0 || sp > 0) RTX_PUSH(RTX_ST_SENTINEL);
                            if (d[3] < INF) RTX_PUSH(ch[3]);
                            if (d[2] < INF) RTX_PUSH(ch[2]);
                            if (d[1] < INF) RTX_PUSH(ch[1]);
                            node = ch[0];
                            T.store_ray(s, r2, false); T.store_rayf(s, f);
                            T.cur[s] = ei;
                            in_inst = true;
                        }
#else
                        // nothing left in the TLAS: the query ends inside the instance (retire re-reads the world ray), no way back needed
                        if (RTX_SKIP_LAST_SENTINEL == 0 || sp > 0) RTX_PUSH(RTX_ST_SENTINEL);
                        T.store_ray(s, r2, false); T.store_rayf(s, f);
                        T.cur[s] = ei;
                        node = e.a;
                        in_inst = true;
#endif
                    } else if (RTX_E_BARE_FAST && e.kind != RTX_GEOM_LIST && e.xf_count == 0 && e.volume < 0 && (RTX_E_BARE_FAST > 1 || e.kind == RTX_GEOM_QUAD)) {
                        // a bare primitive (the walls of every Cornell box, the spheres of RandomScene): the test inlined, without the
                        // generic entry machinery (wrapper chain, Volume / list handling, out-of-line dispatch, nextafter);
                        // cornell-lucy k_extend 1572 -> 1809 Mrays/s
                        RayD r;
                        T.load_ray(s, r);
                        Best B;
                        T.load_best(s, B);
                        double t;
                        if ((FEAT & RTX_F_QUAD) && e.kind == RTX_GEOM_QUAD) {
                            if (COUNT) tc.quads++;
                            t = isect_quad(S.quads + 16 * (size_t)e.index, r, tmin, B.t, nullptr);
                        } else if ((FEAT & RTX_F_SPHERE) && e.kind == RTX_GEOM_SPHERE) {
                            if (COUNT) tc.spheres++;
                            t = isect_sphere_incl(S.spheres + 8 * (size_t)e.index, r, tmin, B.t, B.have);
                        } else if ((FEAT & RTX_F_OTHER_PRIM) && e.kind == RTX_GEOM_TRIANGLE) {
                            if (COUNT) tc.tris++;
                            t = isect_tri(S.tris + RTX_TRI_D * (size_t)e.index, r, nullptr);
                            if (!(tmin <= t && t <= B.t)) t = RTX_NAN_D;
                        } else if ((FEAT & RTX_F_OTHER_PRIM) && e.kind == RTX_GEOM_CIRCLE) {
                            if (COUNT) tc.quads++;
                            t = isect_circle(S.circles + 8 * (size_t)e.index, r, tmin, B.t);
                        } else if (FEAT & RTX_F_PLANE) {
                            if (COUNT) tc.planes++;
                            t = isect_plane(S.planes + 8 * (size_t)e.index, r);
                            if (!(tmin < t && (t < B.t || (B.have && t == B.t)))) t = RTX_NAN_D;
                        } else t = RTX_NAN_D;   // a kind outside the variant's vocabulary: unreachable for a scene the mask covers
                        B.offer(t, ei, e.rank, e.kind, e.index, 0, 0);
                        T.store_best(s, B);
                        if (Policy::ANY_HIT && B.have) node = RTX_ST_DONE;
                        else RTX_POP_TOP();
                    } else if (FEAT & RTX_F_COMPLEX) {
                        VolumeRng vr = {0, 0, 0, 0, 0, true};
                        if (e.volume >= 0) vr = P.volume_rng(job);
                        const bool have = entry_other<NSLOTS, RTX_FEAT_HAS_TIME(FEAT)>(&S, smem, s, ei, tmin, vr.k0, vr.k1, vr.c0, vr.c1, vr.c2, vr.transparent, tcp);
                        if (Policy::ANY_HIT && have) node = RTX_ST_DONE;
                        else RTX_POP_TOP();
                    } else RTX_POP_TOP();   // an entry outside the variant's vocabulary: unreachable for a scene the mask covers
                }
                T.node[s] = node; T.spb[4 * s + 3] = (unsigned char)sp;
                newst = RTX_CLASSIFY(node, in_inst);
            }
        } else {
            // ---- RETIRE + REFILL (warp-collective; one slot per lane per round) ---------------------------------------
            const bool fin = mine && T.node[s] == RTX_ST_DONE;
            bool exhausted = __any_sync(FULL, dry[0] != 0);
            // Policy::CONTINUES (the drain of a pass, k_drain): a retiring ray is shaded on the spot and, when its path goes on, the NEXT ray of
            // the path takes over the slot — no ticket, no record round trip; only lanes whose path ended draw a ticket for a new path.
            bool cont_lane = false;
            RayD r_cont;
            unsigned want = 0;
            int cnt = 0, base = 0;
            if constexpr (!Policy::CONTINUES) {
                // The ticket for the rays this round will load is drawn FIRST (RTX_EARLY_TICKET): every claimed slot is refilled, so the count is
                // known, and the atomic's round trip to L2 (~700 cycles) then runs under the retire work below instead of in front of the refill.
                want = __ballot_sync(FULL, mine);
                cnt = __popc(want);
                if (RTX_EARLY_TICKET && !exhausted && lane == 0) base = atomicAdd(cursor, cnt);
            }
            {
                RayD rw;
                Best B;
                int job = -1;
                B.reset(0);
                rw.ox = rw.oy = rw.oz = rw.dx = rw.dy = rw.dz = rw.tm = 0;
                if (fin) {
                    job = T.job[s];
                    T.load_best(s, B);
                    double tmax_unused;
                    if (T.cur[s] >= 0) P.load(job, rw, tmax_unused);  // an any-hit query may end inside an instance
                    else T.load_ray(s, rw);
                }
                if constexpr (Policy::CONTINUES) cont_lane = P.retire_continue(job, fin, rw, B, r_cont);
                else P.retire(job, fin, rw, B);
            }
            if constexpr (Policy::CONTINUES) {
                want = __ballot_sync(FULL, mine && !cont_lane);
                cnt = __popc(want);
                if (!exhausted && cnt > 0 && lane == 0) base = atomicAdd(cursor, cnt);
            }
            const int job_held = (Policy::CONTINUES && cont_lane) ? T.job[s] : -1;
            if (mine) {
                T.node[s] = RTX_ST_IDLE;
                newst = RTX_PH_NONE;
            }
            if (!exhausted || (Policy::CONTINUES && __any_sync(FULL, cont_lane))) {
                if (!exhausted) {
                    if (!Policy::CONTINUES && !RTX_EARLY_TICKET && lane == 0) base = atomicAdd(cursor, cnt);
                    base = __shfl_sync(FULL, base, 0);
                }
                const int my = cont_lane ? job_held : (exhausted ? njobs : base + __popc(want & ((1u << lane) - 1u)));
                RTX_CHECK(!mine || my >= 0, 7);
                // the record a refill RTX_PREFETCH_DIST tickets from now will load (jobs are handed out in order): from DRAM into L2 meanwhile
                if (RTX_PREFETCH_DIST > 0 && mine && !cont_lane && my + RTX_PREFETCH_DIST < njobs) P.prefetch_far(my + RTX_PREFETCH_DIST);
                if (mine && (cont_lane || my < njobs)) {
                    RayF f; Best B;
                    RayD r;
                    double tmax;
                    if (Policy::CONTINUES && cont_lane) { r = r_cont; tmax = RTX_INF_D; }
                    else P.load(my, r, tmax);
                    B.reset(tmax);
                    // entries tested for every ray: unbounded geometry (infinite Plane, rt/plane.go:17) and, with option pretest_bare, the
                    // few bare primitives beside a mesh (the Cornell walls) that rtx_scene_upload kept out of the TLAS
                    for (int q = 0; q < S.n_unbounded; q++) {
                        const int ei = S.unbounded[q];
                        const DEntry e = S.entries[ei];
                        if ((FEAT & RTX_F_PLANE) && e.xf_count == 0 && e.kind == RTX_GEOM_PLANE) {   // the bare ground plane of RandomScene / HDRITestScene: inlined
                            if (COUNT) tc.planes++;
                            double t = isect_plane(S.planes + 8 * (size_t)e.index, r);
                            if (!(tmin < t && (t < B.t || (B.have && t == B.t)))) t = RTX_NAN_D;
                            B.offer(t, ei, e.rank, RTX_GEOM_PLANE, e.index, 0, 0);
                        } else if ((FEAT & RTX_F_QUAD) && e.xf_count == 0 && e.kind == RTX_GEOM_QUAD) {
                            if (COUNT) tc.quads++;
                            const double t = isect_quad(S.quads + 16 * (size_t)e.index, r, tmin, B.t, nullptr);
                            B.offer(t, ei, e.rank, RTX_GEOM_QUAD, e.index, 0, 0);
                            if (Policy::ANY_HIT && B.have) break;
                        } else if (FEAT & RTX_F_COMPLEX) {   // a wrapped Plane (or, with pretest_bare, another bare kind)
                            RayD ro = r;
                            xform_ray(S, ei, e, ro);
                            B.test_prim(S, e.kind, e.index, ro, tmin, ei, e.rank, 0, 0, tcp);
                        }
                    }
                    int node, sp = 0;
                    if ((Policy::ANY_HIT && B.have) || (S.tlas_root < 0 && !flatTlas)) node = RTX_ST_DONE;
                    else if (flatTlas) {
                        // the top level as a list: every entry box against the ray (conservative float32 slabs, as node_test), the hit ones
                        // pushed with their entry distance and kept sorted, nearest on top
                        make_rayf(r, f); T.store_rayf(s, f);
                        const bool sx = f.offx != 0, sy = f.offy != 0, sz = f.offz != 0;
                        const float ftm = B.ft;
                        for (int q = 0; q < S.n_tlas_flat; q++) {
                            const float4 lo = S.tlas_boxes[2 * q], hi = S.tlas_boxes[2 * q + 1];
                            float tn = fmaxf(fmaxf(fmaf(sx ? hi.x : lo.x, f.ix, f.cnx), fmaf(sy ? hi.y : lo.y, f.iy, f.cny)), fmaxf(fmaf(sz ? hi.z : lo.z, f.iz, f.cnz), ftmin));
                            float tf = fminf(fminf(fmaf(sx ? lo.x : hi.x, f.ix, f.cfx), fmaf(sy ? lo.y : hi.y, f.iy, f.cfy)), fminf(fmaf(sz ? lo.z : hi.z, f.iz, f.cfz), ftm));
                            tn = fmaf(-fabsf(tn), RTX_BOX_EPS, tn);
                            tf = fmaf(fabsf(tf), RTX_BOX_EPS, tf);
                            if (COUNT && (q & 3) == 0) tc.nodes++;   // four 32-byte entry boxes = the bytes of one 4-wide node
                            if (tn <= tf) {
                                const int code = RTX_TLAS_CODE(fmaxf(tn, 1e-30f), __float_as_int(lo.w));
                                int k = sp;
                                while (k > 0 && (unsigned)T.stack[(k - 1) * NS + s] < (unsigned)code) { T.stack[k * NS + s] = T.stack[(k - 1) * NS + s]; k--; }
                                T.stack[k * NS + s] = code;
                                sp++;
                            }
                        }
                        if (sp > 0) { sp--; node = T.stack[sp * NS + s]; } else node = RTX_ST_DONE;
                    }
                    else { node = S.tlas_root; make_rayf(r, f); T.store_rayf(s, f); }
                    T.store_ray(s, r, true); T.store_best(s, B);
                    T.node[s] = node; T.spb[4 * s + 3] = (unsigned char)sp; T.cur[s] = -1; T.job[s] = my;
                    newst = RTX_CLASSIFY(node, false);
                }
                if (!exhausted && cnt > 0 && base + cnt >= njobs) {
                    exhausted = true;
                    if (lane == 0) dry[0] = 1;
                }
            }
            // once the queue is dry, idle slots are parked for good; until then they keep asking for a refill
            if (mine && newst == RTX_PH_NONE && !exhausted) newst = RTX_PH_R;
        }
        if (mine) {   // publish: slot state first, then its phase nibble (BUSY -> newst)
#ifdef RTX_CHECKED
            { const int prev = atomicExch(&T.owner[s], 0); RTX_CHECK(prev == 1 + (int)warp, 1); RTX_CHECK(((colstate[col] >> (4 * j)) & 0xfu) == (unsigned)RTX_PH_BUSY, 2); RTX_CHECK(newst >= 0 && newst <= RTX_PH_NONE, 2); }
#endif
            __threadfence_block();
            atomicXor(const_cast<unsigned*>(colstate) + col, (unsigned)(RTX_PH_BUSY ^ newst) << (4 * j));
        }
    }
#undef RTX_HASZERO_NIB
#undef RTX_PUSH
#undef RTX_POP_TOP
#undef RTX_POP
#undef RTX_CLASSIFY
}
