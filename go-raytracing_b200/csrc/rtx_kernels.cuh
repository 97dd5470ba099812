// rtx_kernels.cuh — the wavefront path tracer: generate -> extend -> shade -> connect -> accumulate, plus resolve
// and the batch entry points used by the parity tests. Hand-written CUDA for sm_100a; no tensor cores (nothing on
// this path is a dense contraction). One wavefront iteration advances every in-flight path by one bounce.
//
// Mapping to the reference (paths relative to /root/reference/):
//   k_generate    Camera.GetRay                      rt/camera.go:368-435
//   k_extend      world.Hit(r, [0.001, inf))         rt/camera.go:451 -> rt/bvh.go:219, rt/aabb.go:59, primitives
//   k_shade       rayColorInternal body              rt/camera.go:453-518, rt/material.go, rt/texture.go, rt/hdri.go
//   k_connect     shadow world.Hit of NEE            rt/camera.go:582, :639
//   k_accumulate  pixelColor.Add(...)                rt/bucket_renderer.go:271
//   k_resolve     scale / gamma / clamp / RGBA8      rt/bucket_renderer.go:275-285
#pragma once
#include "rtx_trace.cuh"

// ---- Philox4x32-10 counter RNG: key = run seed, counter = (pixel, global sample, bounce, stream) ------------------
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ double u01(uint32_t x) { return ((double)x + 0.5) * 2.3283064365386963e-10; }   // (0,1)
__device__ __forceinline__ float u01f(uint32_t x) { return ((float)(x >> 9) + 0.5f) * 1.1920929e-7f; }      // [2^-24, 1 - 2^-24]: k + 0.5 is exact for k < 2^23

__device__ double2 rtx_volume_uniform(const VolumeRng& vr, int entry) {
    uint4 r = philox4x32(vr.c0, vr.c1, vr.c2, 64u + (uint32_t)entry, vr.k0, vr.k1);
    return make_double2(u01(r.x), u01(r.y));
}

enum { Q_MISS = 0, Q_LAMBERTIAN = 1, Q_METAL = 2, Q_DIELECTRIC = 3, Q_LIGHT = 4, Q_ISOTROPIC = 5, Q_COUNT = 6 };
enum { STREAM_CAMERA = 0, STREAM_LENS = 1, STREAM_NEE = 2, STREAM_SCATTER = 3 };

struct DevCamera {  // post-Initialize state (rt/camera.go:41-56)
    double center[3], pixel00[3], du[3], dv[3], u[3], v[3], w[3];
    double defocus_radius, defocus_angle, focus_dist, viewport_w, viewport_h;
    double look_from[3], look_vel[3], look_at[3], look_at_vel[3], vup[3], forward[3];
    int camera_motion, free_camera;
    int width, height;
    float background[3];
    int use_sky, phantom, max_depth;
};

struct Ctl {  // device-resident control block of the wavefront loop
    unsigned long long cursor, total;  // next path id / paths of this pass
    unsigned long long gen_base;
    int n_active, n_cont, n_gen, pad0;
    // The two counters k_shade appends through, in ONE word per iteration parity: low half = survivors written to the other record
    // buffer, high half = shadow requests of this iteration (by parity: k_connect of iteration i may still run while iteration i + 1 is
    // generated, extended and shaded). One 64-bit atomic per warp draws both tickets (RTX_PACKED_TICKETS).
    unsigned long long tk[2];
    int n_mat[Q_COUNT];
    int done, pad;
    int cur_extend, cur_connect[2];  // job cursors of the persistent trace kernels (k_connect: by iteration parity)
    // statistics
    unsigned long long ext_rays, shadow_rays, nodes, tris, spheres, quads, planes, iterations;
    // the tail of a pass: iterations after its last camera path was generated (the stream only shrinks), timed on the device
    unsigned long long t_tail_begin, t_end, tail_iterations;
    int drain_bounce_min, drain_bounce_max;   // k_drain: shallowest / deepest bounce index it traced (its span = the per-bounce launches it replaced)
};
__device__ __forceinline__ unsigned long long rtx_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Path state: a STREAM, not a slot pool. The in-flight paths of one wavefront iteration are the records [0, n_active) of
// rec[cur]: the survivors of the previous iteration (written by k_shade in the order it appended them) followed by the
// paths k_generate starts. Ray i of k_extend is record i, its hit lands in hit[i], the material queues hold such indices,
// and k_shade writes the next ray of a surviving path to the next free record of rec[cur ^ 1]. So every kernel reads and
// writes path state front to back (k_shade's reads follow the material queues, i.e. k_extend's retire order: a window of a
// few MB that stays in L2); there is no free list, no slot recycling and no gather over a GB-sized pool.
// Radiance never travels with the path: every contribution (environment / emission at the end of a path, next-event
// estimation when its shadow ray arrives unoccluded) is added where it arises, with float atomics, to the pixel's sum —
// or, when per-sample moments are requested (parity tests), to a per-sample sum that a final pass squares and folds in.
#define RTX_REC_BYTES 96       /* [ox oy oz time][dx dy dz (pixel | sample << 32)][throughput rgb, bounce | allowLightHits << 16]; packed (128-byte stride: hdri-test -10 %) */
#define RTX_HIT_BYTES 64       /* [Px Py Pz t][Nx Ny Nz (material | front << 31)] */
#define RTX_HIT_BYTES_UV 96    /* scenes with image textures: + [u v - -] in float64 — ImageTexture.Value computes int(u * W), int((1 - v) * H) in float64 (rt/image_texture.go:27-43) */
#define RTX_SHADOW_BYTES 96    /* [ox oy oz tmax][dx dy dz (pixel | sample << 32)][contribution rgb, bounce][-] */
struct Pool {
    int capacity;
    char* rec[2];     // path records, ping-pong
    char* hit;        // hit record of job i of the current iteration
    int* q_mat;       // [Q_COUNT * P] material-sorted shading queues (job indices)
    char* shadow;     // [2P] shadow requests of the current iteration, self-contained
    float* target;    // where radiance goes: float4 per pixel (accumulation buffer) or, with moments, float4 per sample of this pass
    int moments;
    int hit_bytes;    // stride of `hit`: RTX_HIT_BYTES, or RTX_HIT_BYTES_UV in scenes with image textures
    uint32_t npix, sample_base;
    __device__ __forceinline__ char* records(int which) const { return which ? rec[1] : rec[0]; }   // (no dynamic indexing of a kernel parameter)
    __device__ __forceinline__ float* contribution_target(uint32_t pixel, uint32_t sample) const {
        return target + 4 * (moments ? (size_t)(sample - sample_base) * npix + pixel : (size_t)pixel);
    }
    __device__ __forceinline__ void contribute(uint32_t pixel, uint32_t sample, float r, float g, float b) const {
        if (r == 0.f && g == 0.f && b == 0.f) return;
        float* t = contribution_target(pixel, sample);
        atomicAdd(t + 0, r); atomicAdd(t + 1, g); atomicAdd(t + 2, b);
    }
};

struct PassParams {
    int spp, max_depth, camera_max_depth;
    uint32_t seed_lo, seed_hi, sample_base;
    int moments, count_stats, pixel_major;
    int shade_direct;   // 1: k_shade walks the rays in STREAM order (the hit record carries its shading queue), 0: through the material-sorted queues
};

// warp-aggregated queue append: one atomic per warp (ballot + popc), returns this lane's position
__device__ __forceinline__ int warp_append(int* counter, bool pred) {
    unsigned mask = __ballot_sync(__activemask(), pred);
    if (!pred) return -1;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1));
}

__host__ __device__ inline int ctl_survivors(const Ctl* c, int par) { return (int)(c->tk[par] & 0xffffffffull); }
__host__ __device__ inline int ctl_shadow(const Ctl* c, int par) { return (int)(c->tk[par] >> 32); }
#ifndef RTX_PACKED_TICKETS
#define RTX_PACKED_TICKETS 1
#endif
// Warp leader: reserve n_surv survivor records and n_sh shadow requests of parity par; returns the two bases.
__device__ __forceinline__ void ctl_draw_tickets(Ctl* ctl, int par, int n_surv, int n_sh, int& base_surv, int& base_sh) {
#if RTX_PACKED_TICKETS
    const unsigned long long old = atomicAdd(&ctl->tk[par], (unsigned long long)(unsigned)n_surv | ((unsigned long long)(unsigned)n_sh << 32));
    base_surv = (int)(old & 0xffffffffull); base_sh = (int)(old >> 32);
#else
    unsigned* w = reinterpret_cast<unsigned*>(&ctl->tk[par]);   // little-endian halves
    base_surv = n_surv ? (int)atomicAdd(w, (unsigned)n_surv) : 0;
    base_sh = n_sh ? (int)atomicAdd(w + 1, (unsigned)n_sh) : 0;
#endif
}

// ---- K0: iteration bookkeeping ---------------------------------------------------------------------------------
__global__ void k_iter_begin(Ctl* ctl, int capacity, int par) {
    if (threadIdx.x != 0) return;
    int n_cont = ctl_survivors(ctl, par ^ 1);   // survivors: records [0, n_cont) of the buffer k_shade of the previous iteration just wrote
    unsigned long long remaining = ctl->total - ctl->cursor;
    int n_gen = (int)min((unsigned long long)(capacity - n_cont), remaining);
    ctl->gen_base = ctl->cursor;
    ctl->cursor += n_gen;
    ctl->n_cont = n_cont;
    ctl->n_gen = n_gen;
    ctl->n_active = n_cont + n_gen;
    ctl->tk[par] = 0;   // this iteration's survivor and shadow counters (the previous iteration's word is still read by its k_connect)
    ctl->cur_extend = 0; ctl->cur_connect[par] = 0;
    for (int i = 0; i < Q_COUNT; i++) ctl->n_mat[i] = 0;
    ctl->done = (n_cont + n_gen == 0);
    ctl->iterations += (n_cont + n_gen != 0);
    if (n_gen == 0) {   // nothing left to generate: the drain
        if (ctl->t_tail_begin == 0) ctl->t_tail_begin = rtx_globaltimer();
        if (n_cont != 0) ctl->tail_iterations++;
        else if (ctl->t_end == 0) ctl->t_end = rtx_globaltimer();
    }
}

// ---- K1: camera ray generation (rt/camera.go:368-435) ----------------------------------------------------------
template <unsigned FEAT = RTX_F_ALL>
__device__ __forceinline__ RayD camera_ray(const DevCamera& C, int i, int j, double offx, double offy, double tm, double px, double py) {
    D3 center, du, dv, uu, vv;
    D3 p00;
    if (!(FEAT & RTX_F_CAM_SLOW) || (!C.camera_motion && !C.free_camera)) {
        center = ld3(C.center); du = ld3(C.du); dv = ld3(C.dv); uu = ld3(C.u); vv = ld3(C.v); p00 = ld3(C.pixel00);
    } else {  // slow path :390-417
        center = add(ld3(C.look_from), scale(ld3(C.look_vel), tm));
        D3 ww;
        if (C.free_camera) ww = d3(-C.forward[0], -C.forward[1], -C.forward[2]);
        else ww = unit(sub(center, add(ld3(C.look_at), scale(ld3(C.look_at_vel), tm))));
        uu = unit(cross(ld3(C.vup), ww));
        vv = cross(ww, uu);
        D3 vpU = scale(uu, C.viewport_w), vpV = scale(d3(-vv.x, -vv.y, -vv.z), C.viewport_h);
        du = scale(vpU, 1 / (double)C.width);
        dv = scale(vpV, 1 / (double)C.height);
        D3 ul = sub(sub(sub(center, scale(ww, C.focus_dist)), scale(vpU, 1 / 2.0)), scale(vpV, 1 / 2.0));
        p00 = add(ul, scale(add(du, dv), 0.5));
    }
    D3 ps = add(add(p00, scale(du, (double)i + offx)), scale(dv, (double)j + offy));
    D3 ro = center;
    if (C.defocus_angle > 0) {  // defocusDiskSample :354-362
        D3 dU = scale(uu, C.defocus_radius), dV = scale(vv, C.defocus_radius);
        ro = add(add(center, scale(dU, px)), scale(dV, py));
    }
    D3 rd = sub(ps, ro);
    RayD r; r.ox = ro.x; r.oy = ro.y; r.oz = ro.z; r.dx = rd.x; r.dy = rd.y; r.dz = rd.z; r.tm = tm;
    return r;
}

// The stream kernels (generate / shade / accumulate) run on a fixed grid (a few blocks per SM) and stride over the work the
// device-side control block announces: the host never learns the queue lengths, and a launch sized for the whole pool
// (16 K blocks) costs ~70 us of block scheduling even when a handful of paths are left.
// camera path `pid` of the pass: its primary ray and its (pixel | sample << 32) identity
template <unsigned FEAT = RTX_F_ALL>
__device__ __forceinline__ RayD generate_path(const DevCamera& C, const PassParams& pp, unsigned long long pid, unsigned long long& pixsample) {
    unsigned npix = (unsigned)C.width * (unsigned)C.height;
    // path order: pixel-major (all samples of a pixel are consecutive paths) keeps the pixels in flight few, so the radiance
    // atomics stay in L2 and neighbouring lanes start with nearly the same ray; sample-major is the other order
    uint32_t sample, pixel;
    if (pp.pixel_major == 1) {
        // TILE order (default): the 32 lanes of a warp are 32 neighbouring pixels at the same sample index, consecutive warps walk the samples
        // of that tile. The pixels in flight stay as few as in pixel-major order, but a warp's radiance atomics go to 32 different words (one
        // 512-byte run of the accumulation buffer) instead of 32 times to the same one, which the L2 serialises.
        const unsigned long long per_tile = 32ull * (unsigned)pp.spp, full = npix / 32u;
        const unsigned long long tile = pid / per_tile;
        if (tile < full) {
            const unsigned within = (unsigned)(pid - tile * per_tile);
            pixel = (uint32_t)tile * 32u + (within & 31u); sample = pp.sample_base + (within >> 5);
        } else {   // the last npix % 32 pixels
            const unsigned rest = npix - (unsigned)full * 32u;
            const unsigned long long q = pid - full * per_tile;
            pixel = (uint32_t)full * 32u + (uint32_t)(q % rest); sample = pp.sample_base + (uint32_t)(q / rest);
        }
    }
    else if (pp.pixel_major) { pixel = (uint32_t)(pid / (unsigned)pp.spp); sample = pp.sample_base + (uint32_t)(pid % (unsigned)pp.spp); }
    else { sample = pp.sample_base + (uint32_t)(pid / npix); pixel = (uint32_t)(pid % npix); }
    int px = pixel % C.width, py = pixel / C.width;
    uint4 r0 = philox4x32(pixel, sample, 0, STREAM_CAMERA, pp.seed_lo, pp.seed_hi);
    double offx = u01(r0.x) - 0.5, offy = u01(r0.y) - 0.5, tm = u01(r0.z);
    double lx = 0, ly = 0;
    if (C.defocus_angle > 0) {  // uniform unit disk (the reference rejection-samples the same distribution, rt/vec3.go:66-77)
        float rr = sqrtf(u01f(r0.w));
        uint4 r1 = philox4x32(pixel, sample, 0, STREAM_LENS, pp.seed_lo, pp.seed_hi);
        float sn, cs;
        sincospif(2.0f * u01f(r1.x), &sn, &cs);
        lx = rr * cs; ly = rr * sn;
    }
    pixsample = ((unsigned long long)sample << 32) | pixel;
    return camera_ray<FEAT>(C, px, py, offx, offy, tm, lx, ly);
}
#define RTX_FRESH_PATH_FLAGS (0 | (1 << 16))   /* bounce 0, light hits allowed */
__global__ void __launch_bounds__(256, 4) k_generate(Ctl* ctl, Pool pool, int cur, DevCamera C, PassParams pp) {
  const int n_gen = ctl->n_gen, n_cont = ctl->n_cont;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_gen; i += gridDim.x * blockDim.x) {
    unsigned long long ps;
    const RayD r = generate_path(C, pp, ctl->gen_base + (unsigned long long)i, ps);
    char* rec = pool.records(cur) + (size_t)(n_cont + i) * RTX_REC_BYTES;   // appended behind the survivors
    st256d(rec, r.ox, r.oy, r.oz, r.tm);
    st256d(rec + 32, r.dx, r.dy, r.dz, __longlong_as_double((long long)ps));
    strec4(rec + 64, make_float4(1.f, 1.f, 1.f, __int_as_float(RTX_FRESH_PATH_FLAGS)));
  }
}

__device__ __forceinline__ void flush_counters(Ctl* ctl, const TraceCounters& tc) {
    // one set of atomics per lane per launch of a persistent kernel; counters are a measurement option
    atomicAdd(&ctl->nodes, (unsigned long long)tc.nodes);
    atomicAdd(&ctl->tris, (unsigned long long)tc.tris);
    atomicAdd(&ctl->spheres, (unsigned long long)tc.spheres);
    atomicAdd(&ctl->quads, (unsigned long long)tc.quads);
    atomicAdd(&ctl->planes, (unsigned long long)tc.planes);
}

__device__ __forceinline__ Hit best_to_hit(const Best& b) {
    Hit h; h.t = b.t; h.entry = b.entry; h.kind = b.kind; h.prim = b.prim; h.item = b.item;
    return h;
}

// the lean kernel variants (rtx_render_pass picks the first whose mask covers the scene and the camera; RTX_F_ALL is the fallback)
#define RTX_FV_LUCY (RTX_F_QUAD | RTX_F_MESH | RTX_F_XFORM | RTX_F_LIGHTS)                       /* CornellBoxLucy: walls, light, mesh instances, Lambertian */
#define RTX_FV_SKY (RTX_F_SPHERE | RTX_F_PLANE | RTX_F_ENV | RTX_F_METAL | RTX_F_DIELECTRIC)   /* RandomScene, HDRITestScene, SimpleScene, CheckeredSpheres */
#define RTX_FV_BOX (RTX_F_QUAD | RTX_F_SPHERE | RTX_F_LIGHTS | RTX_F_METAL | RTX_F_DIELECTRIC)  /* CornellBoxGlossy, QuadsScene */
#define RTX_FV_CORNELL (RTX_F_QUAD | RTX_F_COMPLEX | RTX_F_LIGHTS | RTX_F_ISOTROPIC)             /* CornellBoxScene, CornellSmoke: walls, rotated Box lists, Volumes (flat kernels only) */

extern __shared__ __align__(16) unsigned char rtx_smem[];  // the trace kernels' ray pool (TracePool)
#define RTX_TRACE_SMEM_BYTES ((size_t)RTX_TRACE_SLOTS * RTX_SLOT_WORDS * 4 + RTX_POOL_EXTRA_BYTES)
#define RTX_TRACE_SMEM_BYTES_OF(FEAT) ((size_t)RTX_TRACE_SLOTS_OF(FEAT) * RTX_SLOT_WORDS_OF(FEAT) * 4 + RTX_POOL_EXTRA_BYTES)
#define RTX_TRACE_BLOCKS_OF(FEAT) ((FEAT) == RTX_F_ALL ? RTX_TRACE_BLOCKS : RTX_FEAT_HAS_TIME(FEAT) ? RTX_TRACE_BLOCKS_LEAN : RTX_TRACE_BLOCKS_LEAN_NT)
#define RTX_TRACE_SLOTS_OF(FEAT) ((FEAT) == RTX_F_ALL ? RTX_TRACE_SLOTS : RTX_TRACE_SLOTS_LEAN)

// ---- K2: extend — closest hit of every active path, then binning into material-sorted shading queues -------------
template <bool UV, unsigned FEAT = RTX_F_ALL>
struct ExtendPolicyT {
    static constexpr bool ANY_HIT = false, CONTINUES = false;
    Ctl* ctl; char* hit; int* q_mat; int capacity; const char* rec; const DevScene* S; uint32_t seed_lo, seed_hi; int direct;
    __device__ __forceinline__ double tmin() const { return 0.001; }  // rt/camera.go:451
    __device__ __forceinline__ void prefetch(int job) const {
        const char* q = rec + (size_t)job * RTX_REC_BYTES;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q + 64));
    }
    __device__ __forceinline__ void load(int job, RayD& r, double& tmax) const {
        const char* q = rec + (size_t)job * RTX_REC_BYTES;
        const D4 a = ld256d(q), c = ld256d(q + 32);
        r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tm = a.w; r.dx = c.x; r.dy = c.y; r.dz = c.z;
        tmax = RTX_INF_D;
    }
    __device__ __forceinline__ void prefetch_far(int job) const {   // a record a later refill will load: pulled from DRAM into L2 now
        const char* q = rec + (size_t)job * RTX_REC_BYTES;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 32));
    }
    __device__ __forceinline__ VolumeRng volume_rng(int job) const {
        const char* q = rec + (size_t)job * RTX_REC_BYTES;
        const unsigned long long ps = (unsigned long long)__double_as_longlong(ld256d(q + 32).w);
        const int bounce = __float_as_int(ldrec4(q + 64).w) & 0xffff;
        VolumeRng vr; vr.k0 = seed_lo; vr.k1 = seed_hi; vr.c0 = (uint32_t)ps; vr.c1 = (uint32_t)(ps >> 32); vr.c2 = (uint32_t)bounce * 4u; vr.transparent = false;
        return vr;
    }
    __device__ __forceinline__ void retire(int job, bool valid, const RayD& r, const Best& b) const {
        int q = -1;
        if (valid) {
            char* h = hit + (size_t)job * (UV ? RTX_HIT_BYTES_UV : RTX_HIT_BYTES);
            if (b.entry < 0) {
                q = Q_MISS;
                if (direct) st256d(h + 32, 0.0, 0.0, 0.0, __longlong_as_double((long long)Q_MISS << 28));   // stream-order shading reads the queue from the record
            } else {
                HitInfo hi;
                finalize_hit<FEAT>(*S, r, best_to_hit(b), UV, hi);
                const int mt = S->mats[hi.mat].type;
                q = mt == RTX_MAT_LAMBERTIAN ? Q_LAMBERTIAN : mt == RTX_MAT_METAL ? Q_METAL : mt == RTX_MAT_DIELECTRIC ? Q_DIELECTRIC
                    : mt == RTX_MAT_DIFFUSE_LIGHT ? Q_LIGHT : Q_ISOTROPIC;
                // material (28 bits) | shading queue (3 bits) | front face
                const long long bits = (long long)(unsigned)hi.mat | ((long long)q << 28) | (hi.front ? (1LL << 31) : 0);
                st256d(h, hi.P.x, hi.P.y, hi.P.z, b.t);
                st256d(h + 32, hi.N.x, hi.N.y, hi.N.z, __longlong_as_double(bits));
                if (UV) st256d(h + 64, hi.u, hi.v, 0.0, 0.0);   // scenes with image textures: the hit's (u, v) in float64, like rec.U / rec.V
            }
        }
        if (direct) return;   // stream-order shading: no queues (and no warp-aggregated atomics in the trace kernel's retire round)
        // material-sorted queues: one warp-aggregated append per queue
#pragma unroll
        for (int k = 0; k < Q_COUNT; k++) {
            const int pos = warp_append(&ctl->n_mat[k], q == k);
            if (q == k) stq(q_mat + (size_t)k * capacity + pos, job);
        }
    }
};
typedef ExtendPolicyT<false> ExtendPolicy;

// UV = true: the variant for scenes with image textures (hit records carry (u, v); sphere UVs cost an acos and an atan2 per hit)
// FEAT: the scene vocabulary the variant contains (RTX_F_*, rtx_device.cuh); rtx_render_pass picks the smallest covering one
template <bool COUNT, bool UV = false, unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(RTX_TRACE_THREADS, RTX_TRACE_BLOCKS_OF(FEAT)) k_extend(Ctl* ctl, Pool pool, int cur, const __grid_constant__ DevScene S, PassParams pp, int* spill) {
    ExtendPolicyT<UV, FEAT> P{ctl, pool.hit, pool.q_mat, pool.capacity, pool.records(cur), &S, pp.seed_lo, pp.seed_hi, pp.shade_direct};
    TraceCounters tc = {0, 0, 0, 0, 0};
    const int n = ctl->n_active;
    trace_persistent<ExtendPolicyT<UV, FEAT>, COUNT, RTX_TRACE_SLOTS_OF(FEAT), FEAT>(S, P, &ctl->cur_extend, n, tc, spill, rtx_smem);
    if (COUNT) flush_counters(ctl, tc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->ext_rays, (unsigned long long)n);
}

// small batches (the drain of a pass): one thread per ray, see trace_simple
template <bool UV = false, unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(256) k_extend_simple(Ctl* ctl, Pool pool, int cur, const __grid_constant__ DevScene S, PassParams pp) {
    ExtendPolicyT<UV, FEAT> P{ctl, pool.hit, pool.q_mat, pool.capacity, pool.records(cur), &S, pp.seed_lo, pp.seed_hi, pp.shade_direct};
    TraceCounters tc = {0, 0, 0, 0, 0};
    const int n = ctl->n_active;
    trace_simple<ExtendPolicyT<UV, FEAT>, false, FEAT>(S, P, n, tc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->ext_rays, (unsigned long long)n);
}

template <bool COUNT, bool UV = false>
__global__ void __launch_bounds__(256) k_extend_flat(Ctl* ctl, Pool pool, int cur, const __grid_constant__ DevScene S, PassParams pp) {
    ExtendPolicyT<UV> P{ctl, pool.hit, pool.q_mat, pool.capacity, pool.records(cur), &S, pp.seed_lo, pp.seed_hi, pp.shade_direct};
    TraceCounters tc = {0, 0, 0, 0, 0};
    const int n = ctl->n_active;
    trace_flat<ExtendPolicyT<UV>, COUNT>(S, P, n, tc);
    if (COUNT) flush_counters(ctl, tc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->ext_rays, (unsigned long long)n);
}

// ---- textures (rt/texture.go:43-45, :63-77) ----------------------------------------------------------------------
__device__ __forceinline__ int clampi_img(int x, int high) { return x < 0 ? 0 : (x < high ? x : high - 1); }  // rt/image_loader.go:112-120
// Perlin noise with the caller's tables (rt/noise.go:30-92), float64, reference operation order
__device__ __noinline__ double perlin_noise(const double* vec, const int* perm, double px, double py, double pz) {
    const double fx = floor(px), fy = floor(py), fz = floor(pz);
    const double u = px - fx, v = py - fy, w = pz - fz;
    const int i = (int)fx, j = (int)fy, k = (int)fz;
    double accum = 0.0;
    for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
            for (int dk = 0; dk < 2; dk++) {
                const int idx = perm[(i + di) & 255] ^ perm[256 + ((j + dj) & 255)] ^ perm[512 + ((k + dk) & 255)];
                const double* c = vec + 3 * idx;
                const double wx = u - (double)di, wy = v - (double)dj, wz = w - (double)dk;
                accum += ((double)di * u + (1 - (double)di) * (1 - u)) * ((double)dj * v + (1 - (double)dj) * (1 - v)) *
                         ((double)dk * w + (1 - (double)dk) * (1 - w)) * (c[0] * wx + c[1] * wy + c[2] * wz);
            }
    return accum;
}
template <unsigned FEAT = RTX_F_ALL>
__device__ __forceinline__ float3 tex_value(const DevScene& S, int id, D3 p, double u = 0.0, double v = 0.0) {
    DTexture t = S.texs[id];
    for (int guard = 0; guard < 8 && t.type == RTX_TEX_CHECKER; guard++) {
        long long xi = (long long)floor(t.inv_scale * p.x + 1e-4);
        long long yi = (long long)floor(t.inv_scale * p.y + 1e-4);
        long long zi = (long long)floor(t.inv_scale * p.z + 1e-4);
        bool even = ((xi + yi + zi) % 2) == 0;
        t = S.texs[even ? t.even : t.odd];
    }
    if ((FEAT & RTX_F_TEX_X) && t.type == RTX_TEX_NOISE) {  // NoiseTexture.Value rt/texture.go:81-85: 0.5 (1 + sin(scale z + 10 turb(scale p, 7)))
        const double* vec = S.perlin_vec + (size_t)t.even * 768;
        const int* perm = S.perlin_perm + (size_t)t.even * 768;
        const double sc = t.inv_scale;   // NoiseTexture.scale (not inverted)
        double accum = 0.0, weight = 1.0, qx = sc * p.x, qy = sc * p.y, qz = sc * p.z;   // Perlin.Turb rt/noise.go:55-65
        for (int oct = 0; oct < 7; oct++) {
            accum += weight * perlin_noise(vec, perm, qx, qy, qz);
            weight *= 0.5;
            qx = 2 * qx; qy = 2 * qy; qz = 2 * qz;
        }
        const float tv = (float)(0.5 * (1.0 + sin(sc * p.z + 10.0 * fabs(accum))));
        return make_float3(tv, tv, tv);
    }
    if ((FEAT & RTX_F_TEX_X) && t.type == RTX_TEX_IMAGE) {  // ImageTexture.Value rt/image_texture.go:27-43 + ImageLoader.PixelData rt/image_loader.go:97-120
        const int4 dim = S.img_dim[t.even];
        const double uu = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);                  // Interval{0,1}.Clamp, float64 like the reference
        const double vv = 1.0 - (v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v));           // flip V to image coordinates
        const int i = clampi_img((int)(uu * (double)dim.x), dim.x), j = clampi_img((int)(vv * (double)dim.y), dim.y);
        const size_t first = ((size_t)(unsigned)dim.w << 32) | (unsigned)dim.z;
        const float4 c = __ldg(S.img_rgb + first + (size_t)j * dim.x + i);
        return make_float3(c.x, c.y, c.z);
    }
    return make_float3(t.color[0], t.color[1], t.color[2]);
}

// ---- HDRI (rt/hdri.go) -------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampi_dev(int x, int low, int high) { return x < low ? low : (x < high ? x : high - 1); }  // rt/image_loader.go:112-120
__device__ __forceinline__ void dir_to_uv(const DevScene& S, D3 dir, double& u, double& v) {  // rt/hdri.go:75-94
    const double PI = 3.14159265358979323846;
    D3 d = unit(dir);
    double phi = atan2(d.z, d.x), theta = asin(d.y);
    u = 0.5 + phi / (2 * PI);
    v = 0.5 - theta / PI;
    u = u + S.env_rot / (2 * PI);
    u = u - floor(u);
}
__device__ __forceinline__ float3 env_lookup(const DevScene& S, D3 dir) {  // Sample :120-128 + PixelDataBilinear rt/image_loader.go:399-436
    double u, v;
    dir_to_uv(S, dir, u, v);
    int W = S.env_w, H = S.env_h;
    double px = u * (double)W - 0.5, py = v * (double)H - 0.5;
    int x0 = (int)floor(px), y0 = (int)floor(py);
    int x1 = x0 + 1, y1 = y0 + 1;
    float fx = (float)(px - (double)x0), fy = (float)(py - (double)y0);
    x0 = ((x0 % W) + W) % W; x1 = ((x1 % W) + W) % W;
    y0 = clampi_dev(y0, 0, H); y1 = clampi_dev(y1, 0, H);
    float4 c00 = __ldg(S.env_tex + (size_t)y0 * W + x0), c10 = __ldg(S.env_tex + (size_t)y0 * W + x1);
    float4 c01 = __ldg(S.env_tex + (size_t)y1 * W + x0), c11 = __ldg(S.env_tex + (size_t)y1 * W + x1);
    float3 c0 = make_float3(c00.x * (1 - fx) + c10.x * fx, c00.y * (1 - fx) + c10.y * fx, c00.z * (1 - fx) + c10.z * fx);
    float3 c1 = make_float3(c01.x * (1 - fx) + c11.x * fx, c01.y * (1 - fx) + c11.y * fx, c01.z * (1 - fx) + c11.z * fx);
    return make_float3(c0.x * (1 - fy) + c1.x * fy, c0.y * (1 - fy) + c1.y * fy, c0.z * (1 - fy) + c1.z * fy);
}
__device__ __forceinline__ int search_cdf(const double* cdf, int n, double xi) {  // rt/hdri.go:300-322, n = len(cdf)-1
    int low = 0, high = n;
    while (low < high) {
        int mid = (low + high) / 2;
        if (__ldg(cdf + mid + 1) <= xi) low = mid + 1;
        else high = mid;
    }
    if (low >= n) low = n - 1;
    if (low < 0) low = 0;
    return low;
}
__device__ __forceinline__ double env_pdf(const DevScene& S, D3 dir) {  // rt/hdri.go:262-297
    const double PI = 3.14159265358979323846;
    if (!S.env_is || S.env_total == 0) return 1.0 / (4.0 * PI);
    double u, v;
    dir_to_uv(S, dir, u, v);
    int x = clampi_dev((int)(u * (double)S.env_w), 0, S.env_w), y = clampi_dev((int)(v * (double)S.env_h), 0, S.env_h);
    double theta = (0.5 - v) * PI;
    double sinTheta = cos(theta);
    if (sinTheta < 1e-10) sinTheta = 1e-10;
    double p = S.env_pdf[(size_t)y * S.env_w + x] * (double)(S.env_w * S.env_h) / (2.0 * PI * PI * sinTheta);
    return p < 1e-10 ? 1e-10 : p;
}
__device__ __forceinline__ void env_sample(const DevScene& S, double xi1, double xi2, D3& dir, float3& emission, double& pdf) {  // :228-259
    const double PI = 3.14159265358979323846;
    int y = search_cdf(S.env_marg, S.env_h, xi1);
    int x = search_cdf(S.env_cond + (size_t)y * (S.env_w + 1), S.env_w, xi2);
    double u = ((double)x + 0.5) / (double)S.env_w, v = ((double)y + 0.5) / (double)S.env_h;
    u = u - S.env_rot / (2 * PI);  // UVToDirection :97-113
    u = u - floor(u);
    double phi = (u - 0.5) * 2 * PI, theta = (0.5 - v) * PI;
    double ct = cos(theta);
    dir = d3(ct * cos(phi), sin(theta), ct * sin(phi));
    float4 e = __ldg(S.env_tex + (size_t)y * S.env_w + x);
    emission = make_float3(e.x, e.y, e.z);
    pdf = env_pdf(S, dir);
}

// uniform direction on the unit sphere (the reference rejection-samples the same distribution, rt/vec3.go:45-54)
__device__ __forceinline__ D3 unit_sphere(uint32_t a, uint32_t b) {
    float z = 1.0f - 2.0f * u01f(a);
    float r = sqrtf(fmaxf(0.f, 1.0f - z * z));
    float sn, cs;
    sincospif(2.0f * u01f(b), &sn, &cs);
    return d3((double)(r * cs), (double)(r * sn), (double)z);
}

// ---- K4: shade — one thread per queue element, queues concatenated in material order --------------------------------
#ifndef RTX_SHADE_PREFETCH
#define RTX_SHADE_PREFETCH 1
#endif
#ifndef RTX_SHADE_BLOCKS
#define RTX_SHADE_BLOCKS 2   /* resident 256-thread blocks per SM k_shade is compiled for */
#endif
// What one shaded element carries out of the material code: the path's next record (written only if it survives) and up to two
// shadow requests per Lambertian hit, in fixed registers (no dynamically indexed arrays: those live in local memory).
struct ShadeVars {
    bool cont, has_env, has_area;
    D3 P, nd, env_dir, area_dir;
    double tm, pixbits, area_tmax;
    float4 th;
    int bounce0;
    float3 env_c, area_c;
    __device__ __forceinline__ void reset() {
        cont = has_env = has_area = false;
        P = nd = env_dir = area_dir = d3(0, 0, 0);
        tm = pixbits = area_tmax = 0;
        th = make_float4(0, 0, 0, 0);
        bounce0 = 0;
        env_c = area_c = make_float3(0, 0, 0);
    }
};

// rayColorInternal for one ray whose query is finished (rt/camera.go:443-518): `type` is its shading queue; V.tm / V.pixbits / V.th hold
// the path record's time, pixel | sample and throughput | flags; P_in, N, (hu, hv), mat, front are the hit (unused for Q_MISS).
template <unsigned FEAT = RTX_F_ALL>
__device__ __forceinline__ void shade_element(const DevScene& S, const DevCamera& C, const PassParams& pp, const Pool& pool, const int type, const D3 rd,
                                              const D3 P_in, const D3 N, const double hu, const double hv, const int mat, const bool front, ShadeVars& V) {
    bool& cont = V.cont; bool& has_env = V.has_env; bool& has_area = V.has_area;
    D3& nd = V.nd; D3& env_dir = V.env_dir; D3& area_dir = V.area_dir;
    double& area_tmax = V.area_tmax;
    float4& th = V.th;
    int& bounce0 = V.bounce0;
    float3& env_c = V.env_c; float3& area_c = V.area_c;
    const double pixbits = V.pixbits;
    int flags = __float_as_int(th.w);
    int bounce = flags & 0xffff;
    bounce0 = bounce;
    bool allow = (flags >> 16) & 1;
    const unsigned long long psb = (unsigned long long)__double_as_longlong(pixbits);
    uint2 ps = make_uint2((uint32_t)psb, (uint32_t)(psb >> 32));
    if (type == Q_MISS) {  // rt/camera.go:451-466
        float3 col;
        if ((FEAT & RTX_F_ENV) && S.env_w > 0) {
            bool primary = (bounce == 0) && (pp.max_depth == pp.camera_max_depth);  // depth == c.MaxDepth
            if (C.phantom && primary) col = make_float3(0, 0, 0);
            else col = env_lookup(S, rd);
        } else if (C.use_sky) {  // SkyGradient :520-526
            D3 ud = unit(rd);
            float t = (float)(0.5 * (ud.y + 1.0));
            col = make_float3((1.f - t) + 0.5f * t, (1.f - t) + 0.7f * t, (1.f - t) + 1.0f * t);
        } else col = make_float3(C.background[0], C.background[1], C.background[2]);
        pool.contribute(ps.x, ps.y, th.x * col.x, th.y * col.y, th.z * col.z);
    } else {
        V.P = P_in;
        const D3 P = P_in;
        DMaterial M = S.mats[mat];
        if (type == Q_LIGHT) {  // Scatter == false: rt/camera.go:473-481, rt/material.go:226-236
            if (allow) {
                float3 e = tex_value<FEAT>(S, M.tex, P, hu, hv);
                pool.contribute(ps.x, ps.y, th.x * e.x, th.y * e.y, th.z * e.z);
            }
        } else {
            uint4 rs = philox4x32(ps.x, ps.y, (uint32_t)bounce, STREAM_SCATTER, pp.seed_lo, pp.seed_hi);
            float3 att;
            bool scattered = true, next_allow = true;
            if (type == Q_LAMBERTIAN) {  // rt/material.go:57-68
                nd = add(N, unit_sphere(rs.x, rs.y));
                if (fabs(nd.x) < 1e-8 && fabs(nd.y) < 1e-8 && fabs(nd.z) < 1e-8) nd = N;
                att = tex_value<FEAT>(S, M.tex, P, hu, hv);
                if ((FEAT & RTX_F_LIGHTS) && S.n_lights > 0) {  // useMIS, rt/camera.go:487-517
                    const double PI = 3.14159265358979323846;
                    uint4 rn = philox4x32(ps.x, ps.y, (uint32_t)bounce, STREAM_NEE, pp.seed_lo, pp.seed_hi);
                    int li = (int)(u01(rn.x) * (double)S.n_lights);
                    if (li >= S.n_lights) li = S.n_lights - 1;
                    if ((FEAT & RTX_F_ENV) && S.env_w > 0 && S.env_is && S.env_total != 0) {  // sampleHDRILight :565-607
                        D3 ldir; float3 em; double pdfH;
                        env_sample(S, u01(rs.z), u01(rs.w), ldir, em, pdfH);
                        double cosT = dot(N, ldir);
                        if (cosT > 0) {
                            double pdfB = cosT / PI;  // Lambertian.PDF rt/material.go:70-76
                            double w = pdfH / (pdfH + pdfB);
                            double s = cosT / pdfH * w;
                            float3 cc = make_float3(fminf((float)(em.x * s) * att.x, 20.f), fminf((float)(em.y * s) * att.y, 20.f), fminf((float)(em.z * s) * att.z, 20.f));
                            env_dir = ldir;
                            env_c = make_float3(th.x * cc.x, th.y * cc.y, th.z * cc.z);
                            has_env = true;
                        }
                    }
                    int lq = S.light_quads[li];
                    if (lq >= 0) {  // sampleAreaLight :610-678
                        const double* q = S.quads + 16 * (size_t)lq;
                        D3 lp = add(add(ld3(q), scale(ld3(q + 3), u01(rn.y))), scale(ld3(q + 6), u01(rn.z)));  // SamplePoint rt/quad.go:87-92
                        D3 toL = sub(lp, P);
                        double dist = sqrt(len2(toL));
                        D3 ldir = unit(toL);
                        double cosT = dot(N, ldir);
                        double cosL = fabs(dot(ld3(q + 12), d3(-ldir.x, -ldir.y, -ldir.z)));
                        if (cosT > 0 && !(cosL < 0.001)) {
                            float3 em = tex_value<FEAT>(S, S.mats[S.quad_mat[lq]].tex, lp);  // lightQuad.mat.Emitted(0,0,lightPoint)
                            if (S.mats[S.quad_mat[lq]].type != RTX_MAT_DIFFUSE_LIGHT) em = make_float3(0, 0, 0);
                            double area = sqrt(len2(cross(ld3(q + 3), ld3(q + 6))));
                            double pdfL = (dist * dist) / (cosL * area);
                            double pdfB = cosT / PI;
                            double w = pdfL / (pdfL + pdfB);
                            double s = cosT / pdfL * w;
                            double nl = (double)S.n_lights;
                            float3 cc = make_float3(fminf((float)(em.x * s * att.x * nl), 20.f), fminf((float)(em.y * s * att.y * nl), 20.f),
                                                    fminf((float)(em.z * s * att.z * nl), 20.f));
                            area_dir = ldir; area_tmax = dist - 0.001;
                            area_c = make_float3(th.x * cc.x, th.y * cc.y, th.z * cc.z);
                            has_area = true;
                        }
                    }
                    next_allow = false;  // indirect path must not pick up the light again (:514)
                }
            } else if ((FEAT & RTX_F_METAL) && type == Q_METAL) {  // rt/material.go:113-119
                double dn = dot(rd, N);
                D3 refl = sub(rd, scale(N, 2 * dn));
                nd = add(unit(refl), scale(unit_sphere(rs.x, rs.y), M.fuzz));
                att = make_float3(M.albedo[0], M.albedo[1], M.albedo[2]);
                scattered = dot(nd, N) > 0;
            } else if ((FEAT & RTX_F_DIELECTRIC) && type == Q_DIELECTRIC) {  // rt/material.go:164-188
                att = make_float3(1.f, 1.f, 1.f);
                double ri = front ? 1.0 / M.ior : M.ior;
                D3 ud = unit(rd);
                double cosT = fmin(dot(d3(-ud.x, -ud.y, -ud.z), N), 1.0);
                double sinT = sqrt(1.0 - cosT * cosT);
                bool cannot = ri * sinT > 1.0;
                bool reflect = cannot;
                if (!cannot) {
                    double r0 = (1 - ri) / (1 + ri);
                    r0 = r0 * r0;
                    double om = 1 - cosT;
                    double refl = r0 + (1 - r0) * (om * om * om * om * om);
                    reflect = refl > u01(rs.x);
                }
                if (reflect) nd = sub(ud, scale(N, 2 * dot(ud, N)));
                else {  // Refract rt/vec3.go:110-117
                    D3 perp = scale(add(ud, scale(N, cosT)), ri);
                    D3 par = scale(N, -sqrt(fabs(1.0 - len2(perp))));
                    nd = add(perp, par);
                }
            } else if (FEAT & RTX_F_ISOTROPIC) {  // Q_ISOTROPIC rt/material.go:266-270
                nd = unit_sphere(rs.x, rs.y);
                att = tex_value<FEAT>(S, M.tex, P, hu, hv);
            } else { att = make_float3(0.f, 0.f, 0.f); scattered = false; }   // a material outside the variant's vocabulary: unreachable for a covered scene
            if (!scattered) {
                has_env = has_area = false;  // absorbed: emission of a scattering material is zero
            } else {
                th.x *= att.x; th.y *= att.y; th.z *= att.z;
                bounce++;
                th.w = __int_as_float((bounce & 0xffff) | (next_allow ? (1 << 16) : 0));
                cont = bounce < pp.max_depth;  // rayColorInternal(depth <= 0) returns black (:444-446)
            }
        }
    }
}

// Warp-collective: appends the survivor's next record to the other record buffer and the shadow requests to this iteration's half.
__device__ __forceinline__ void shade_commit(Ctl* ctl, const Pool& pool, const int cur, const ShadeVars& V, unsigned long long* quad = nullptr, int quad_iter = 0) {
    const bool cont = V.cont, has_env = V.has_env, has_area = V.has_area;
    const D3 P = V.P, nd = V.nd, env_dir = V.env_dir, area_dir = V.area_dir;
    const double tm = V.tm, pixbits = V.pixbits, area_tmax = V.area_tmax;
    const float4 th = V.th;
    const int bounce0 = V.bounce0;
    const float3 env_c = V.env_c, area_c = V.area_c;
    // Survivors go to the next free records of the other buffer, shadow requests to this iteration's half, in the order the warps arrive.
    // Both tickets come from ONE 64-bit atomic (ctl_draw_tickets) drawn by one lane per warp — or, in quad mode, per four warps — before
    // any lane uses either: with three dependent warp_append calls this HBM-bound kernel spent its time waiting on atomic returns, and
    // with one atomic per counter and warp it ran at the L2's rate for same-address atomics.
    const unsigned am = __activemask();
    const unsigned mc = __ballot_sync(am, cont), me = __ballot_sync(am, has_env), ma = __ballot_sync(am, has_area);
    const int lane = threadIdx.x & 31, leader = __ffs(am) - 1;
    int bc = 0, bs = 0;
    if (quad) {
        // FOUR warps (threads 128 g .. 128 g + 127 of a 256-thread block, every one of them here in every iteration) draw one ticket pair:
        // the counters of a pass are single words, and k_shade's rate was the rate of same-address atomics in L2 (one per warp: 41 % of
        // its stall samples; packed into one word: -12 % time; one per four warps: see DESIGN.md section 8). Named barrier 1 + g, double-
        // buffered shared words, so two barriers per iteration are enough.
        const int w = threadIdx.x >> 5, g = w >> 2;
        unsigned long long* cnt = quad + 16 * (quad_iter & 1);   // [0..7] per-warp packed counts, [8..9] per-group bases
        if (lane == 0) cnt[w] = (unsigned long long)(unsigned)__popc(mc) | ((unsigned long long)(unsigned)(__popc(me) + __popc(ma)) << 32);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        if ((w & 3) == 0 && lane == 0) {
            const unsigned long long tot = cnt[w] + cnt[w + 1] + cnt[w + 2] + cnt[w + 3];
            cnt[8 + g] = tot ? atomicAdd(&ctl->tk[cur], tot) : 0ull;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        unsigned long long base = cnt[8 + g];
        for (int k = 4 * g; k < w; k++) base += cnt[k];
        bc = (int)(base & 0xffffffffull); bs = (int)(base >> 32);
    } else {
        if (lane == leader && (mc | me | ma)) ctl_draw_tickets(ctl, cur, __popc(mc), __popc(me) + __popc(ma), bc, bs);
        bc = __shfl_sync(am, bc, leader);
        bs = __shfl_sync(am, bs, leader);
    }
    const unsigned below = (1u << lane) - 1u;
    if (cont) {
        char* out = pool.records(cur ^ 1) + (size_t)(bc + __popc(mc & below)) * RTX_REC_BYTES;
        st256d(out, P.x, P.y, P.z, tm);
        st256d(out + 32, nd.x, nd.y, nd.z, pixbits);
        strec4(out + 64, th);
    }
    // shadow requests carry everything k_connect needs (origin = the hit point, where the contribution goes)
    if (has_env) {
        char* q = pool.shadow + (size_t)(bs + __popc(me & below)) * RTX_SHADOW_BYTES;
        st256d(q, P.x, P.y, P.z, RTX_INF_D);
        st256d(q + 32, env_dir.x, env_dir.y, env_dir.z, pixbits);
        strec4(q + 64, make_float4(env_c.x, env_c.y, env_c.z, __int_as_float(bounce0)));
    }
    if (has_area) {
        char* q = pool.shadow + (size_t)(bs + __popc(me) + __popc(ma & below)) * RTX_SHADOW_BYTES;
        st256d(q, P.x, P.y, P.z, area_tmax);
        st256d(q + 32, area_dir.x, area_dir.y, area_dir.z, pixbits);
        strec4(q + 64, make_float4(area_c.x, area_c.y, area_c.z, __int_as_float(bounce0)));
    }
}

// QT < 0: one launch walks all six queues back to back. QT >= 0: the launch shades queue QT only — the material is a compile-time
// constant, every other material's code is gone, and the kernel is compiled for RTX_SHADE_BLOCKS_Q resident blocks (k_shade waits
// on gathers through the queues: ncu long_scoreboard 16 cycles per issue at 16 warps per SM; the one-material kernels fit more warps).
#ifndef RTX_SHADE_BLOCKS_Q
#define RTX_SHADE_BLOCKS_Q 3
#endif
#ifndef RTX_SHADE_BLOCKS_LEAN
#define RTX_SHADE_BLOCKS_LEAN RTX_SHADE_BLOCKS
#endif
#ifndef RTX_SHADE_PIPELINE
#define RTX_SHADE_PIPELINE 1
#endif
#ifndef RTX_SHADE_QUAD_TICKETS
#define RTX_SHADE_QUAD_TICKETS 1
#endif
template <int QT, unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(256, FEAT != RTX_F_ALL ? RTX_SHADE_BLOCKS_LEAN : QT < 0 ? RTX_SHADE_BLOCKS : (QT == Q_LAMBERTIAN ? RTX_SHADE_BLOCKS : RTX_SHADE_BLOCKS_Q)) k_shade(Ctl* ctl, Pool pool, int cur, DevScene S, DevCamera C, PassParams pp) {
  const int n_items = QT < 0 ? ctl->n_active : ctl->n_mat[QT < 0 ? 0 : QT];
  const int n_rounded = (n_items + 31) & ~31;   // whole warps stay together for the queue appends
  int nq[Q_COUNT];
#pragma unroll
  for (int k = 0; k < Q_COUNT; k++) nq[k] = ctl->n_mat[k];
  // (queue, job) of element i of the concatenated queues: prefix over the six queue counts
  auto locate = [&](int i, int& type) {
      if (QT >= 0) {
          type = i < n_items ? QT : -1;
          return type >= 0 ? ldq(pool.q_mat + (size_t)QT * pool.capacity + i) : -1;
      }
      type = -1;
      int idx = i;
#pragma unroll
      for (int k = 0; k < Q_COUNT; k++)
          if (type < 0) {
              if (idx < nq[k]) type = k;
              else idx -= nq[k];
          }
      return type >= 0 ? ldq(pool.q_mat + (size_t)type * pool.capacity + idx) : -1;
  };
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (QT < 0 && pp.shade_direct) {
      // STREAM ORDER (option shade_direct): thread i shades ray i; records and hit records are read front to back, fully coalesced, and the
      // shading queue comes out of the hit record. The material-sorted order costs this HBM-bound kernel its bandwidth (gathers through the
      // queues: 44 % of the copy peak over a pass) and costs the trace kernel an atomic round trip per retire round; what it buys — warps of
      // one material — matters little to a kernel that issues 18 % of the time.
#if RTX_SHADE_PIPELINE
      // Software pipeline: the records of this thread's NEXT element are requested before this element's commit, so that the round trip of
      // the commit's ticket atomics (41 % of the kernel's stall samples) and the latency of those loads (22 %) overlap instead of adding up.
      D4 ro4, rd4, hn, hp, huv;
      float4 th4;
      auto request = [&](int k) {
          const char* rec = pool.records(cur) + (size_t)k * RTX_REC_BYTES;
          const char* hrec = pool.hit + (size_t)k * pool.hit_bytes;
          ro4 = ld256d(rec); rd4 = ld256d(rec + 32); hn = ld256d(hrec + 32); th4 = ldrec4(rec + 64); hp = ld256d(hrec);
          if (S.n_images > 0) huv = ld256d(hrec + 64);
      };
      huv.x = huv.y = huv.z = huv.w = 0.0;
      if (i < n_items) request(i);
#if RTX_SHADE_QUAD_TICKETS
      __shared__ unsigned long long quad_words[32];
      const int n_quad = (n_items + 127) & ~127;   // the four warps of a ticket group make the same number of trips
      for (int it = 0; i < n_quad; i += stride, it++) {
#else
      for (; i < n_rounded; i += stride) {
#endif
          ShadeVars V;
          V.reset();
          if (i < n_items) {
              V.tm = ro4.w; V.pixbits = rd4.w;
              V.th = th4;
              const long long bits = __double_as_longlong(hn.w);
              const int type = (int)((bits >> 28) & 7);
              const D3 P = type != Q_MISS ? d3(hp.x, hp.y, hp.z) : d3(0, 0, 0);
              const double hu = (S.n_images > 0 && type != Q_MISS) ? huv.x : 0.0, hv = (S.n_images > 0 && type != Q_MISS) ? huv.y : 0.0;
              shade_element<FEAT>(S, C, pp, pool, type, d3(rd4.x, rd4.y, rd4.z), P, d3(hn.x, hn.y, hn.z), hu, hv, (int)(bits & 0x0fffffff), (bits >> 31) & 1, V);
          }
          if (i + stride < n_items) request(i + stride);
#if RTX_SHADE_QUAD_TICKETS
          shade_commit(ctl, pool, cur, V, quad_words, it);
#else
          shade_commit(ctl, pool, cur, V);
#endif
      }
#else
      for (; i < n_rounded; i += stride) {
          ShadeVars V;
          V.reset();
          if (i < n_items) {
              const char* rec = pool.records(cur) + (size_t)i * RTX_REC_BYTES;
              const char* hrec = pool.hit + (size_t)i * pool.hit_bytes;
              const D4 ro4 = ld256d(rec), rd4 = ld256d(rec + 32), hn = ld256d(hrec + 32);
              V.tm = ro4.w; V.pixbits = rd4.w;
              V.th = ldrec4(rec + 64);
              const long long bits = __double_as_longlong(hn.w);
              const int type = (int)((bits >> 28) & 7);
              D3 P = d3(0, 0, 0);
              double hu = 0.0, hv = 0.0;
              if (type != Q_MISS) {
                  const D4 hp = ld256d(hrec);
                  P = d3(hp.x, hp.y, hp.z);
                  if (S.n_images > 0) { const D4 huv = ld256d(hrec + 64); hu = huv.x; hv = huv.y; }
              }
              shade_element<FEAT>(S, C, pp, pool, type, d3(rd4.x, rd4.y, rd4.z), P, d3(hn.x, hn.y, hn.z), hu, hv, (int)(bits & 0x0fffffff), (bits >> 31) & 1, V);
          }
          shade_commit(ctl, pool, cur, V);
      }
#endif
      return;
  }
  int type_next = -1;
  int job_next = i < n_rounded ? locate(i, type_next) : -1;
  for (; i < n_rounded; i += stride) {
    const int type = QT >= 0 ? QT : type_next, job_cur = job_next;   // (an element past the end of a one-material queue has job -1)
    // software pipeline: the records of this thread's NEXT element are gathers through the queue (two dependent hops); its
    // queue slot is read now and its record lines are prefetched into L1 while this element is shaded
    job_next = -1; type_next = -1;
    if (RTX_SHADE_PREFETCH && i + stride < n_rounded) {
        job_next = locate(i + stride, type_next);
        if (job_next >= 0) {
            const char* nrec = pool.records(cur) + (size_t)job_next * RTX_REC_BYTES;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nrec));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nrec + 64));
            if (type_next != Q_MISS) asm volatile("prefetch.global.L1 [%0];" ::"l"(pool.hit + (size_t)job_next * pool.hit_bytes));
        }
    } else if (i + stride < n_rounded) job_next = locate(i + stride, type_next);
    bool valid = QT >= 0 ? job_cur >= 0 : type >= 0;
    ShadeVars V;
    V.reset();
    if (valid) {
        const int job = job_cur;
        const char* rec = pool.records(cur) + (size_t)job * RTX_REC_BYTES;
        const D4 ro4 = ld256d(rec), rd4 = ld256d(rec + 32);
        V.tm = ro4.w; V.pixbits = rd4.w;
        V.th = ldrec4(rec + 64);
        D3 P = d3(0, 0, 0), N = d3(0, 0, 0);
        double hu = 0.0, hv = 0.0;   // rec.U, rec.V: carried only in scenes with image textures (96-byte hit records)
        int mat = 0;
        bool front = false;
        if (type != Q_MISS) {
            const char* hrec = pool.hit + (size_t)job * pool.hit_bytes;
            const D4 hp = ld256d(hrec), hn = ld256d(hrec + 32);
            P = d3(hp.x, hp.y, hp.z);
            N = d3(hn.x, hn.y, hn.z);
            if (S.n_images > 0) { const D4 huv = ld256d(hrec + 64); hu = huv.x; hv = huv.y; }
            const long long bits = __double_as_longlong(hn.w);
            mat = (int)(bits & 0x0fffffff);
            front = (bits >> 31) & 1;
        }
        shade_element<FEAT>(S, C, pp, pool, type, d3(rd4.x, rd4.y, rd4.z), P, N, hu, hv, mat, front, V);
    }
    shade_commit(ctl, pool, cur, V);
  }
}

// ---- K2+K4 fused, for the worlds the flat kernels trace (a handful of entries, no mesh) ------------------------------------
// There the query is cheap and every lane runs the same loop, so the hit never has to leave the registers: the thread that traced the ray
// shades it (same shade_element, same Philox counters: the paths are the ones the separate kernels produce) and appends the survivor. The hit
// record (64 B written, 64 B read), the queue slot and the second read of the path record disappear — for these scenes both separate kernels
// are HBM-bound (profiles/r01_k_shade_hdri.md) — at the price of shading without material-sorted warps.
#ifndef RTX_BOUNCE_QUAD_TICKETS
#define RTX_BOUNCE_QUAD_TICKETS 1
#endif
template <bool UV, unsigned FEAT = RTX_F_ALL>
struct BouncePolicyT {
    static constexpr bool ANY_HIT = false, CONTINUES = false;
    // one ticket atomic per four warps (shade_commit; blocks are 256 threads) — for the SKY vocabulary only: hdri-test 64 spp 73.3 -> 67.2 ms (its
    // launch of 16.7 M rays drew 520 K tickets from one word in a millisecond, the L2's rate for one address); where the rays of a warp differ
    // in length the barriers cost more than the atomics did (cornell-glossy 14.1 -> 14.5 ms, earth 1.8 -> 1.9 ms)
    static constexpr bool QUAD_TICKETS = RTX_BOUNCE_QUAD_TICKETS != 0 && FEAT == RTX_FV_SKY;
    Ctl* ctl; Pool pool; int cur; const DevScene* S; const DevCamera* C; PassParams pp;
    int n_cont; unsigned long long gen_base;   // jobs >= n_cont are fresh camera paths: generated here, never written as records
    mutable double pixbits_; mutable float4 th_;   // identity and throughput | flags of the job this thread is working on (load -> retire)
    unsigned long long* quad = nullptr; mutable int trip = 0;   // shared words of the ticket groups, loop trips so far (their double-buffer index)
    __device__ __forceinline__ double tmin() const { return 0.001; }  // rt/camera.go:451
    __device__ __forceinline__ void prefetch(int job) const {
        if (job >= n_cont) return;   // a fresh camera path: generated, not loaded
        const char* q = pool.records(cur) + (size_t)job * RTX_REC_BYTES;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q + 64));
    }
    __device__ __forceinline__ void load(int job, RayD& r, double& tmax) const {
        if (job >= n_cont) {
            unsigned long long ps;
            r = generate_path<FEAT>(*C, pp, gen_base + (unsigned long long)(job - n_cont), ps);
            pixbits_ = __longlong_as_double((long long)ps);
            th_ = make_float4(1.f, 1.f, 1.f, __int_as_float(RTX_FRESH_PATH_FLAGS));
        } else {
            const char* q = pool.records(cur) + (size_t)job * RTX_REC_BYTES;
            const D4 a = ld256d(q), c = ld256d(q + 32);
            r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tm = a.w; r.dx = c.x; r.dy = c.y; r.dz = c.z;
            pixbits_ = c.w;
            th_ = ldrec4(q + 64);
        }
        tmax = RTX_INF_D;
    }
    __device__ __forceinline__ VolumeRng volume_rng(int) const {
        const unsigned long long ps = (unsigned long long)__double_as_longlong(pixbits_);
        const int bounce = __float_as_int(th_.w) & 0xffff;
        VolumeRng vr; vr.k0 = pp.seed_lo; vr.k1 = pp.seed_hi; vr.c0 = (uint32_t)ps; vr.c1 = (uint32_t)(ps >> 32); vr.c2 = (uint32_t)bounce * 4u; vr.transparent = false;
        return vr;
    }
    __device__ __forceinline__ void retire(int, bool valid, const RayD& r, const Best& b) const {
        ShadeVars V;
        V.reset();
        if (valid) {
            V.tm = r.tm; V.pixbits = pixbits_;
            V.th = th_;
            int type = Q_MISS;
            HitInfo hi;
            hi.P = hi.N = d3(0, 0, 0); hi.mat = 0; hi.front = false; hi.u = hi.v = 0;
            if (b.entry >= 0) {
                finalize_hit<FEAT>(*S, r, best_to_hit(b), UV, hi);
                const int mt = S->mats[hi.mat].type;
                type = mt == RTX_MAT_LAMBERTIAN ? Q_LAMBERTIAN : mt == RTX_MAT_METAL ? Q_METAL : mt == RTX_MAT_DIELECTRIC ? Q_DIELECTRIC
                     : mt == RTX_MAT_DIFFUSE_LIGHT ? Q_LIGHT : Q_ISOTROPIC;
            }
            shade_element<FEAT>(*S, *C, pp, pool, type, d3(r.dx, r.dy, r.dz), hi.P, hi.N, UV ? hi.u : 0.0, UV ? hi.v : 0.0, hi.mat, hi.front, V);
        }
        if (QUAD_TICKETS) shade_commit(ctl, pool, cur, V, quad, trip++);
        else shade_commit(ctl, pool, cur, V);
    }
};

#ifndef RTX_BOUNCE_BLOCKS
#define RTX_BOUNCE_BLOCKS 2   /* resident 256-thread blocks per SM k_bounce_flat is compiled for */
#endif
// the lean variants (rtx_render_pass picks the first whose mask covers the scene and the camera; RTX_F_ALL is the fallback)
#ifndef RTX_BOUNCE_BLOCKS_LEAN
#define RTX_BOUNCE_BLOCKS_LEAN 4   /* resident 256-thread blocks of the lean variants (104-120 registers uncapped). hdri-test 64 spp, Mpaths/s: all-features kernel 4147; lean with 2 blocks 6180, 3 blocks (78 registers) 6564-6667, 4 blocks (64 registers, 24 B of spills) 6675; cornell-glossy 2092 / 2480 / 2917 / 3132 */
#endif
template <bool COUNT, bool UV = false, unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(256, (FEAT & RTX_F_COMPLEX) ? RTX_BOUNCE_BLOCKS : RTX_BOUNCE_BLOCKS_LEAN) k_bounce_flat(Ctl* ctl, Pool pool, int cur, const __grid_constant__ DevScene S, const __grid_constant__ DevCamera C, PassParams pp) {
    __shared__ unsigned long long quad_words[32];
    BouncePolicyT<UV, FEAT> P{ctl, pool, cur, &S, &C, pp, ctl->n_cont, ctl->gen_base, 0.0, make_float4(0.f, 0.f, 0.f, 0.f), quad_words, 0};
    TraceCounters tc = {0, 0, 0, 0, 0};
    const int n = ctl->n_active;
    trace_flat<BouncePolicyT<UV, FEAT>, COUNT, FEAT>(S, P, n, tc);
    if (COUNT) flush_counters(ctl, tc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->ext_rays, (unsigned long long)n);
}

// ---- K2+K4 fused for hierarchy worlds (option fuse_tree) ---------------------------------------------------------------------------
// The persistent trace kernel with shading in its RETIRE phase: the lane that retires a finished query shades it (same shade_element,
// same Philox counters, so the paths are the ones k_extend + k_shade produce) and appends the survivor and its shadow requests. The hit
// record (64 B written, 64 B read back), the queue slots and k_shade's gather of the path record disappear, and the HBM-bound shading
// stream runs inside the latency-bound traversal instead of after it. The material code is one out-of-line call (bounce_shade), so the
// traversal loop keeps its registers; what the call needs travels through local memory.
template <bool UV, unsigned FEAT = RTX_F_ALL>
__device__ __noinline__ void bounce_shade(const DevScene* S, const DevCamera* C, const PassParams* pp, const Pool* pool, double pixbits, float4 th, const RayD* rp,
                                          const Best* bp, ShadeVars* Vp) {
    const RayD r = *rp;
    const Best b = *bp;
    ShadeVars V;
    V.reset();
    V.tm = r.tm;
    V.pixbits = pixbits;
    V.th = th;
    int type = Q_MISS;
    HitInfo hi;
    hi.P = hi.N = d3(0, 0, 0); hi.mat = 0; hi.front = false; hi.u = hi.v = 0;
    if (b.entry >= 0) {
        finalize_hit<FEAT>(*S, r, best_to_hit(b), UV, hi);
        const int mt = S->mats[hi.mat].type;
        type = mt == RTX_MAT_LAMBERTIAN ? Q_LAMBERTIAN : mt == RTX_MAT_METAL ? Q_METAL : mt == RTX_MAT_DIELECTRIC ? Q_DIELECTRIC
             : mt == RTX_MAT_DIFFUSE_LIGHT ? Q_LIGHT : Q_ISOTROPIC;
    }
    shade_element<FEAT>(*S, *C, *pp, *pool, type, d3(r.dx, r.dy, r.dz), hi.P, hi.N, UV ? hi.u : 0.0, UV ? hi.v : 0.0, hi.mat, hi.front, V);
    *Vp = V;
}
template <bool UV>
struct BounceTreePolicyT {
    static constexpr bool ANY_HIT = false, CONTINUES = false;
    Ctl* ctl; const Pool* pool; int cur; const DevScene* S; const DevCamera* C; const PassParams* pp; const char* rec;
    __device__ __forceinline__ double tmin() const { return 0.001; }  // rt/camera.go:451
    __device__ __forceinline__ void prefetch_far(int) const {}
    __device__ __forceinline__ void load(int job, RayD& r, double& tmax) const {
        const char* q = rec + (size_t)job * RTX_REC_BYTES;
        const D4 a = ld256d(q), c = ld256d(q + 32);
        r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tm = a.w; r.dx = c.x; r.dy = c.y; r.dz = c.z;
        tmax = RTX_INF_D;
    }
    __device__ __forceinline__ VolumeRng volume_rng(int job) const {
        const char* q = rec + (size_t)job * RTX_REC_BYTES;
        const unsigned long long ps = (unsigned long long)__double_as_longlong(ld256d(q + 32).w);
        const int bounce = __float_as_int(ldrec4(q + 64).w) & 0xffff;
        VolumeRng vr; vr.k0 = pp->seed_lo; vr.k1 = pp->seed_hi; vr.c0 = (uint32_t)ps; vr.c1 = (uint32_t)(ps >> 32); vr.c2 = (uint32_t)bounce * 4u; vr.transparent = false;
        return vr;
    }
    __device__ __forceinline__ void retire(int job, bool valid, const RayD& r, const Best& b) const {
        ShadeVars V;
        if (valid) {
            const char* q = rec + (size_t)job * RTX_REC_BYTES;
            bounce_shade<UV>(S, C, pp, pool, ld256d(q + 32).w, ldrec4(q + 64), &r, &b, &V);
        } else V.reset();
        shade_commit(ctl, *pool, cur, V);
    }
};
template <bool COUNT, bool UV = false>
__global__ void __launch_bounds__(RTX_TRACE_THREADS, RTX_TRACE_BLOCKS) k_bounce(Ctl* ctl, const __grid_constant__ Pool pool, int cur, const __grid_constant__ DevScene S,
                                                                               const __grid_constant__ DevCamera C, const __grid_constant__ PassParams pp, int* spill) {
    BounceTreePolicyT<UV> P{ctl, &pool, cur, &S, &C, &pp, pool.records(cur)};
    TraceCounters tc = {0, 0, 0, 0, 0};
    const int n = ctl->n_active;
    trace_persistent<BounceTreePolicyT<UV>, COUNT, RTX_TRACE_SLOTS>(S, P, &ctl->cur_extend, n, tc, spill, rtx_smem);
    if (COUNT) flush_counters(ctl, tc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->ext_rays, (unsigned long long)n);
}

// ---- the barrier-free drain of a pass (option fuse_drain) ------------------------------------------------------------------------------
// Once a pass has generated its last camera path the stream only shrinks, and the wavefront loop runs mid-size batches (1 M ... 50 K rays) through
// persistent launches that cannot finish faster than the ~60 rounds their slowest lanes need: 7.6 ms per pass at half rate.
// Here ONE persistent launch finishes all remaining bounces. Its jobs are the survivors the last regular iteration left in rec[cur] (a static
// queue); a ray that retires is shaded on the spot (bounce_shade, as k_bounce) and, when its path goes on, the next ray TAKES OVER THE SLOT
// (Policy::CONTINUES in trace_persistent): no barrier between bounces, no queue traffic, and the drain lasts as long as the longest chain of paths
// one slot gets. The path's record is updated in place (an instance exit re-reads the world ray from it). Shadow requests are collected and
// traced by one k_connect launch afterwards. Same shade_element, same Philox counters: the same paths as the iteration loop would trace.
template <bool UV, unsigned FEAT = RTX_F_ALL>
struct DrainPolicyT {
    static constexpr bool ANY_HIT = false, CONTINUES = true;
    Ctl* ctl; const Pool* pool; int cur; int par; const DevScene* S; const DevCamera* C; const PassParams* pp;   // cur: record buffer of the jobs; par: shadow counter
    __device__ __forceinline__ double tmin() const { return 0.001; }
    __device__ __forceinline__ char* record(int job) const { return pool->records(cur) + (size_t)job * RTX_REC_BYTES; }
    // records are rewritten during the launch by whichever warp of the block retires the slot: read past L1
    __device__ __forceinline__ static D4 ld256v(const void* p) {
        D4 r;
        asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
        asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.z), "=d"(r.w) : "l"((const char*)p + 16) : "memory");
        return r;
    }
    __device__ __forceinline__ static float4 ld128v(const void* p) {
        float4 r;
        asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
        return r;
    }
    __device__ __forceinline__ void load(int job, RayD& r, double& tmax) const {
        const char* q = record(job);
        const D4 a = ld256v(q), c = ld256v(q + 32);
        r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tm = a.w; r.dx = c.x; r.dy = c.y; r.dz = c.z;
        tmax = RTX_INF_D;
    }
    __device__ __forceinline__ void prefetch_far(int) const {}
    __device__ __forceinline__ VolumeRng volume_rng(int job) const {
        const char* q = record(job);
        const unsigned long long ps = (unsigned long long)__double_as_longlong(ld256v(q + 32).w);
        const int bounce = __float_as_int(ld128v(q + 64).w) & 0xffff;
        VolumeRng vr; vr.k0 = pp->seed_lo; vr.k1 = pp->seed_hi; vr.c0 = (uint32_t)ps; vr.c1 = (uint32_t)(ps >> 32); vr.c2 = (uint32_t)bounce * 4u; vr.transparent = false;
        return vr;
    }
    // warp-collective; returns whether this lane's path goes on, with its next ray in r_next (and in its record)
    __device__ __forceinline__ bool retire_continue(int job, bool valid, const RayD& r, const Best& b, RayD& r_next) const {
        ShadeVars V;
        if (valid) {
            const char* q = record(job);
            bounce_shade<UV, FEAT>(S, C, pp, pool, ld256v(q + 32).w, ld128v(q + 64), &r, &b, &V);
        } else V.reset();
        const unsigned am = __activemask();
        const unsigned me = __ballot_sync(am, V.has_env), ma = __ballot_sync(am, V.has_area), mv = __ballot_sync(am, valid);
        const int lane = threadIdx.x & 31, leader = __ffs(am) - 1;
        int bs = 0;
        const int bmin = __reduce_min_sync(am, valid ? V.bounce0 : 0x7fffffff), bmax = __reduce_max_sync(am, valid ? V.bounce0 : -1);
        if (lane == leader) {
            if (me | ma) { int unused; ctl_draw_tickets(ctl, par, 0, __popc(me) + __popc(ma), unused, bs); }
            if (mv) {
                atomicAdd(&ctl->ext_rays, (unsigned long long)__popc(mv));
                if (bmin < ctl->drain_bounce_min) atomicMin(&ctl->drain_bounce_min, bmin);
                if (bmax > ctl->drain_bounce_max) atomicMax(&ctl->drain_bounce_max, bmax);
            }
        }
        bs = __shfl_sync(am, bs, leader);
        const unsigned below = (1u << lane) - 1u;
        if (V.has_env) {
            char* q = pool->shadow + (size_t)(bs + __popc(me & below)) * RTX_SHADOW_BYTES;
            st256d(q, V.P.x, V.P.y, V.P.z, RTX_INF_D);
            st256d(q + 32, V.env_dir.x, V.env_dir.y, V.env_dir.z, V.pixbits);
            strec4(q + 64, make_float4(V.env_c.x, V.env_c.y, V.env_c.z, __int_as_float(V.bounce0)));
        }
        if (V.has_area) {
            char* q = pool->shadow + (size_t)(bs + __popc(me) + __popc(ma & below)) * RTX_SHADOW_BYTES;
            st256d(q, V.P.x, V.P.y, V.P.z, V.area_tmax);
            st256d(q + 32, V.area_dir.x, V.area_dir.y, V.area_dir.z, V.pixbits);
            strec4(q + 64, make_float4(V.area_c.x, V.area_c.y, V.area_c.z, __int_as_float(V.bounce0)));
        }
        if (V.cont) {   // the path's record, in place: its next ray, its throughput and bounce count
            char* out = record(job);
            st256d(out, V.P.x, V.P.y, V.P.z, V.tm);
            st256d(out + 32, V.nd.x, V.nd.y, V.nd.z, V.pixbits);
            strec4(out + 64, V.th);
            __threadfence_block();
            r_next.ox = V.P.x; r_next.oy = V.P.y; r_next.oz = V.P.z; r_next.dx = V.nd.x; r_next.dy = V.nd.y; r_next.dz = V.nd.z; r_next.tm = V.tm;
        }
        return V.cont;
    }
};
// resident blocks the lean drain variants are compiled for: shading runs inside the kernel, so they want more registers than the lean trace kernels
#ifndef RTX_DRAIN_BLOCKS_LEAN
#define RTX_DRAIN_BLOCKS_LEAN 5
#endif
#define RTX_DRAIN_BLOCKS_OF(FEAT) ((FEAT) == RTX_F_ALL ? RTX_TRACE_BLOCKS : RTX_DRAIN_BLOCKS_LEAN)
template <bool UV = false, unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(RTX_TRACE_THREADS, RTX_DRAIN_BLOCKS_OF(FEAT)) k_drain(Ctl* ctl, const __grid_constant__ Pool pool, int cur, int par, const __grid_constant__ DevScene S,
                                                                              const __grid_constant__ DevCamera C, const __grid_constant__ PassParams pp, int* spill) {
    DrainPolicyT<UV, FEAT> P{ctl, &pool, cur, par, &S, &C, &pp};
    TraceCounters tc = {0, 0, 0, 0, 0};
    trace_persistent<DrainPolicyT<UV, FEAT>, false, RTX_TRACE_SLOTS_OF(FEAT), FEAT>(S, P, &ctl->cur_extend, ctl->n_active, tc, spill, rtx_smem);
}
// before the drain: job cursor and shadow half; after it (and its k_connect): the books of the pass
__global__ void k_drain_begin(Ctl* ctl, int par, int last_par) {   // last_par: parity of the last per-bounce iteration (its survivors are the drain's jobs)
    if (threadIdx.x != 0) return;
    ctl->n_active = ctl_survivors(ctl, last_par);
    ctl->tk[0] = 0; ctl->tk[1] = 0;
    ctl->cur_extend = 0; ctl->cur_connect[par] = 0;
    ctl->drain_bounce_min = 0x7fffffff; ctl->drain_bounce_max = -1;
    if (ctl->t_tail_begin == 0) ctl->t_tail_begin = rtx_globaltimer();
}
__global__ void k_drain_end(Ctl* ctl) {
    if (threadIdx.x != 0) return;
    ctl->tk[0] = 0; ctl->tk[1] = 0; ctl->n_active = 0; ctl->done = 1;
    // the bounce rounds the launch covered count as the wavefront iterations they replaced
    const int rounds = ctl->drain_bounce_max >= ctl->drain_bounce_min ? ctl->drain_bounce_max - ctl->drain_bounce_min + 1 : 1;
    ctl->iterations += rounds; ctl->tail_iterations += rounds;
    if (ctl->t_end == 0) ctl->t_end = rtx_globaltimer();
}

// ---- K3: connect — shadow rays of next-event estimation (any hit in [0.001, tmax]) ----------------------------------
struct ConnectPolicy {
    static constexpr bool ANY_HIT = true, CONTINUES = false;
    Pool pool; uint32_t seed_lo, seed_hi;
    __device__ __forceinline__ double tmin() const { return 0.001; }  // rt/camera.go:579, :636
    __device__ __forceinline__ void prefetch(int job) const {
        const char* q = pool.shadow + (size_t)job * RTX_SHADOW_BYTES;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q + 64));
    }
    __device__ __forceinline__ void load(int job, RayD& r, double& tmax) const {
        const char* q = pool.shadow + (size_t)job * RTX_SHADOW_BYTES;
        const D4 o0 = ld256d(q), d0 = ld256d(q + 32);
        r.ox = o0.x; r.oy = o0.y; r.oz = o0.z; r.dx = d0.x; r.dy = d0.y; r.dz = d0.z; r.tm = 0;  // NewRay(hitPoint, lightDir, 0)
        tmax = o0.w;
    }
    __device__ __forceinline__ void prefetch_far(int job) const {
        const char* q = pool.shadow + (size_t)job * RTX_SHADOW_BYTES;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 32));
    }
    __device__ __forceinline__ VolumeRng volume_rng(int job) const {
        const char* q = pool.shadow + (size_t)job * RTX_SHADOW_BYTES;
        const unsigned long long ps = (unsigned long long)__double_as_longlong(ld256d(q + 32).w);
        const int bounce = __float_as_int(ldrec4(q + 64).w);   // the bounce whose hit issued the request
        const double tmax = ld256d(q).w;
        VolumeRng vr; vr.k0 = seed_lo; vr.k1 = seed_hi; vr.c0 = (uint32_t)ps; vr.c1 = (uint32_t)(ps >> 32);
        vr.c2 = (uint32_t)bounce * 4u + (tmax == RTX_INF_D ? 2u : 1u); vr.transparent = false;
        return vr;
    }
    __device__ __forceinline__ void retire(int job, bool valid, const RayD&, const Best& b) const {
        if (valid && b.entry < 0) {   // unoccluded: the contribution shade prepared arrives
            const char* q = pool.shadow + (size_t)job * RTX_SHADOW_BYTES;
            const unsigned long long ps = (unsigned long long)__double_as_longlong(ld256d(q + 32).w);
            const float4 cc = ldrec4(q + 64);
            pool.contribute((uint32_t)ps, (uint32_t)(ps >> 32), cc.x, cc.y, cc.z);
        }
    }
};

template <bool COUNT, unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(RTX_TRACE_THREADS, RTX_TRACE_BLOCKS_OF(FEAT)) k_connect(Ctl* ctl, Pool pool, int par, const __grid_constant__ DevScene S, PassParams pp, int* spill) {
    ConnectPolicy P{pool, pp.seed_lo, pp.seed_hi};
    TraceCounters tc = {0, 0, 0, 0, 0};
    const int n = ctl_shadow(ctl, par);
    trace_persistent<ConnectPolicy, COUNT, RTX_TRACE_SLOTS_OF(FEAT), FEAT>(S, P, &ctl->cur_connect[par], n, tc, spill, rtx_smem);
    if (COUNT) flush_counters(ctl, tc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->shadow_rays, (unsigned long long)n);
}

template <unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(256) k_connect_simple(Ctl* ctl, Pool pool, int par, const __grid_constant__ DevScene S, PassParams pp) {
    ConnectPolicy P{pool, pp.seed_lo, pp.seed_hi};
    TraceCounters tc = {0, 0, 0, 0, 0};
    const int n = ctl_shadow(ctl, par);
    trace_simple<ConnectPolicy, false, FEAT>(S, P, n, tc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->shadow_rays, (unsigned long long)n);
}

template <bool COUNT, unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(256) k_connect_flat(Ctl* ctl, Pool pool, int par, const __grid_constant__ DevScene S, PassParams pp) {
    ConnectPolicy P{pool, pp.seed_lo, pp.seed_hi};
    TraceCounters tc = {0, 0, 0, 0, 0};
    const int n = ctl_shadow(ctl, par);
    trace_flat<ConnectPolicy, COUNT, FEAT>(S, P, n, tc);
    if (COUNT) flush_counters(ctl, tc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->shadow_rays, (unsigned long long)n);
}

// ---- scene upload: the float32 pre-test records from the float64 triangle records (tri_pretest_reject, rtx_device.cuh) ----------------------
__global__ void __launch_bounds__(256) k_tris32(const double* tris, int n, float4* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* t = tris + RTX_TRI_D * (size_t)i;   // v0, e1, e2, n
    float4 r[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float x = (float)t[3 * k], y = (float)t[3 * k + 1], z = (float)t[3 * k + 2];
        const double m = fmax(fabs(t[3 * k]), fmax(fabs(t[3 * k + 1]), fabs(t[3 * k + 2])));
        r[k] = make_float4(x, y, z, __double2float_ru(m));
    }
    out[3 * (size_t)i] = r[0]; out[3 * (size_t)i + 1] = r[1]; out[3 * (size_t)i + 2] = r[2];
}

// ---- K6a: end of a pass — every pixel received `spp` samples; with moments, fold the per-sample sums into sum and sum of squares ----
__global__ void __launch_bounds__(256) k_pass_finish(float4* accum, float4* accum_sq, const float4* per_sample, int npix, int spp, int moments) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 a = accum[i];
    if (moments) {
        float4 q = accum_sq[i];
        for (int s = 0; s < spp; s++) {
            const float4 L = per_sample[(size_t)s * npix + i];
            a.x += L.x; a.y += L.y; a.z += L.z;
            q.x += L.x * L.x; q.y += L.y * L.y; q.z += L.z * L.z;
        }
        accum_sq[i] = q;
    }
    a.w += (float)spp;
    accum[i] = a;
}

// ---- K6b: resolve (rt/bucket_renderer.go:275-285, rt/utils.go:85-90) -----------------------------------------------------
__global__ void k_resolve_rgba8(const float4* accum, int npix, double scale, uchar4* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 a = accum[i];
    double c[3] = {(double)a.x * scale, (double)a.y * scale, (double)a.z * scale};
    unsigned char o[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double g = c[k] > 0 ? sqrt(c[k]) : 0;   // LinearToGamma
        g = g < 0.0 ? 0.0 : (g > 0.999 ? 0.999 : g);  // Interval{0,0.999}.Clamp
        o[k] = (unsigned char)(256 * g);
    }
    out[i] = make_uchar4(o[0], o[1], o[2], 255);
}

// ---- batch entry points for the parity tests ----------------------------------------------------------------------------
template <unsigned FEAT = RTX_F_ALL>
struct BatchPolicyT {
    static constexpr bool ANY_HIT = false, CONTINUES = false;
    const DevScene* S; const double* rays; double t0, t1;
    int* entry_id; int* prim_id; double* t; double* normal; unsigned char* front; double* uv; double* p;
    __device__ __forceinline__ double tmin() const { return t0; }
    __device__ __forceinline__ void prefetch(int) const {}
    __device__ __forceinline__ void prefetch_far(int) const {}
    __device__ __forceinline__ void load(int job, RayD& r, double& tmax) const {
        const double* q = rays + 7 * (size_t)job;
        r.ox = q[0]; r.oy = q[1]; r.oz = q[2]; r.dx = q[3]; r.dy = q[4]; r.dz = q[5]; r.tm = q[6];
        tmax = t1;
    }
    __device__ __forceinline__ VolumeRng volume_rng(int) const { VolumeRng vr = {0, 0, 0, 0, 0, true}; return vr; }
    __device__ __forceinline__ void retire(int job, bool valid, const RayD& r, const Best& b) const {
        if (!valid) return;
        const size_t i = (size_t)job;
        const bool hit = b.entry >= 0;
        HitInfo hi;
        if (hit) finalize_hit<FEAT>(*S, r, best_to_hit(b), true, hi);
        if (entry_id) entry_id[i] = hit ? b.entry : -1;
        if (prim_id) prim_id[i] = hit ? b.item : -1;
        if (t) t[i] = hit ? b.t : 0;
        if (normal) { normal[3 * i] = hit ? hi.N.x : 0; normal[3 * i + 1] = hit ? hi.N.y : 0; normal[3 * i + 2] = hit ? hi.N.z : 0; }
        if (front) front[i] = hit && hi.front;
        if (uv) { uv[2 * i] = hit ? hi.u : 0; uv[2 * i + 1] = hit ? hi.v : 0; }
        if (p) { p[3 * i] = hit ? hi.P.x : 0; p[3 * i + 1] = hit ? hi.P.y : 0; p[3 * i + 2] = hit ? hi.P.z : 0; }
    }
};

typedef BatchPolicyT<> BatchPolicy;
// the level-1 parity entry runs the same kernel variant a rendered pass of the scene runs (FEAT), so the bit-exact tests cover the lean code
template <unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(RTX_TRACE_THREADS, RTX_TRACE_BLOCKS_OF(FEAT)) k_trace_closest(const __grid_constant__ DevScene S, const double* rays, int n, double tmin, double tmax, int* cursor, int* spill,
        int* entry_id, int* prim_id, double* t, double* normal, unsigned char* front, double* uv, double* p) {
    BatchPolicyT<FEAT> P{&S, rays, tmin, tmax, entry_id, prim_id, t, normal, front, uv, p};
    TraceCounters tc = {0, 0, 0, 0, 0};
    trace_persistent<BatchPolicyT<FEAT>, false, RTX_TRACE_SLOTS_OF(FEAT), FEAT>(S, P, cursor, n, tc, spill, rtx_smem);
}

template <unsigned FEAT = RTX_F_ALL>
__global__ void __launch_bounds__(256) k_trace_closest_flat(const __grid_constant__ DevScene S, const double* rays, int n, double tmin, double tmax, int* entry_id, int* prim_id,
        double* t, double* normal, unsigned char* front, double* uv, double* p) {
    BatchPolicyT<FEAT> P{&S, rays, tmin, tmax, entry_id, prim_id, t, normal, front, uv, p};
    TraceCounters tc = {0, 0, 0, 0, 0};
    trace_flat<BatchPolicyT<FEAT>, false, FEAT>(S, P, n, tc);
}

__global__ void k_camera_rays(DevCamera C, const int* ij, const double* sq, const double* disk, const double* tm, long long n, double* out) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    RayD r = camera_ray(C, ij[2 * k], ij[2 * k + 1], sq[2 * k], sq[2 * k + 1], tm[k], disk[2 * k], disk[2 * k + 1]);
    double* o = out + 7 * k;
    o[0] = r.ox; o[1] = r.oy; o[2] = r.oz; o[3] = r.dx; o[4] = r.dy; o[5] = r.dz; o[6] = r.tm;
}

__global__ void k_hdri_sample(DevScene S, const double* xi, long long n, double* dir, double* emission, double* pdf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D3 d; float3 e; double p;
    env_sample(S, xi[2 * i], xi[2 * i + 1], d, e, p);
    dir[3 * i] = d.x; dir[3 * i + 1] = d.y; dir[3 * i + 2] = d.z;
    emission[3 * i] = e.x; emission[3 * i + 1] = e.y; emission[3 * i + 2] = e.z;
    pdf[i] = p;
}
__global__ void k_hdri_pdf(DevScene S, const double* dir, long long n, double* pdf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pdf[i] = env_pdf(S, d3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
}
__global__ void k_hdri_lookup(DevScene S, const double* dir, long long n, double* rgb) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 c = env_lookup(S, d3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
    rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
}
