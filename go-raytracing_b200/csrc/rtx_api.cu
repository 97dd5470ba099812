// rtx_api.cu — implementation of the C-ABI in include/rtx_b200.h: context, scene upload (SoA device buffers +
// wide BVH build), camera, the wavefront host loop, resolve, and the batch entry points used for parity.
// There is no CPU fallback anywhere in this library: every arithmetic entry point launches CUDA kernels.
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>   // types only: the library is loaded with dlopen in rtx_create_multi (single-GPU users never need it)

#include "rtx_bvh.hpp"
#include "rtx_kernels.cuh"
#include "rtx_bvh_gpu.cuh"
#include "rtx_rank_gpu.cuh"

#ifndef RTX_PRETEST_BARE_DEFAULT
#define RTX_PRETEST_BARE_DEFAULT 0
#endif
#ifndef RTX_FUSE_DRAIN_DEFAULT
#define RTX_FUSE_DRAIN_DEFAULT (1 << 19)
#endif
#ifndef RTX_DRAIN_PER_BLOCK_DEFAULT
#define RTX_DRAIN_PER_BLOCK_DEFAULT 32   /* rays per block the persistent grids of the drain are sized for */
#endif
#ifndef RTX_SHADE_DIRECT_DEFAULT
#define RTX_SHADE_DIRECT_DEFAULT 1
#endif
#ifndef RTX_SIMPLE_BELOW_DEFAULT
#define RTX_SIMPLE_BELOW_DEFAULT 0
#endif
#ifndef RTX_FUSE_TREE_DEFAULT
#define RTX_FUSE_TREE_DEFAULT 0
#endif

using rtxbvh::Box;
using rtxbvh::Node4;

#ifdef RTX_CHECKED
__global__ void k_check_selftest(int n) { RTX_CHECK(threadIdx.x >= n, 7); }
#endif
static thread_local std::string g_create_error;
static inline float __int_as_float_host(int v) { float f; std::memcpy(&f, &v, sizeof f); return f; }

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct Slab {   // grow-only device allocation, bump-allocated afresh by every use: hot calls never cudaMalloc / cudaFree
    char* base = nullptr;
    size_t cap = 0, off = 0;
    cudaError_t reserve(size_t bytes) {
        off = 0;
        if (bytes <= cap) return cudaSuccess;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc((void**)&base, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    char* take(size_t bytes) { char* p = base + off; off += (bytes + 255) & ~(size_t)255; return off <= cap ? p : nullptr; }
    void release() { if (base) cudaFree(base); base = nullptr; cap = off = 0; }
};

struct rtx_ctx {
    int device = 0;
    Slab scene_slab, work_slab;   // the uploaded scene / working memory of rtx_scene_upload (device BVH build)
    uchar4* rgba_dev = nullptr;   // resolve target, sized with the accumulation buffer
    cudaStream_t stream = nullptr;      // stream in use
    cudaStream_t own_stream = nullptr;  // created by rtx_create
    cudaStream_t connect_stream = nullptr;   // k_connect of iteration i runs here, beside generate / extend / shade of iteration i + 1
    cudaEvent_t ev_shaded = nullptr, ev_connected[2] = {nullptr, nullptr};
    int overlap_connect = 1;
    int fuse_flat = 1;        // flat worlds: closest hit and shading in one kernel (k_bounce_flat); 0 = k_extend_flat + k_shade
    int fuse_tree = RTX_FUSE_TREE_DEFAULT;   // hierarchy worlds: shading inside the persistent trace kernel's RETIRE phase (k_bounce); 0 = k_extend + k_shade
    int pretest_bare = RTX_PRETEST_BARE_DEFAULT;   // hierarchy worlds with a mesh: up to this many bare bounded primitives (the Cornell walls) leave the TLAS and are tested for every ray when it enters the pool (0 = all entries in the TLAS)
    int lean_flat = 1;        // flat worlds: the one-kernel bounce is compiled per scene vocabulary (RTX_FV_*), the smallest covering variant runs; 0 = always the all-features kernel
    unsigned feat_mask = RTX_F_ALL;   // RTX_F_* bits the uploaded scene needs
    int shade_split = 0;      // k_shade as one launch per material queue (1) or one launch over all queues (0, the default: hdri-test 61.5 against 64.0 ms of shading per 64 spp, random 3.1 against 3.9 — the one-material kernels need 48-80 registers instead of 128, but six short launches have six tails)
    unsigned mat_kinds = ~0u;  // bit q: some material of the uploaded scene shades through queue q
    std::string err;
    DevScene S{};
    bool have_scene = false, have_camera = false;
    DevCamera C{};
    int W = 0, H = 0;
    // accumulation
    float4 *accum = nullptr, *accum_sq = nullptr;
    int moments = 0;
    // path pool
    Pool pool{};
    std::vector<void*> pool_allocs;
    // in-flight paths per iteration (upper limit; a pass of fewer paths gets a pool of its own size): persistent trace launches have a fixed
    // tail, so few big launches beat many small ones (cornell-lucy 64 spp: 2 Mi 187, 4 Mi 198, 8 Mi 204; 256 spp: 8 Mi 244, 16 Mi 248, 32 Mi 250 Mpaths/s)
    int64_t pool_paths = 1 << 24;
    Ctl* ctl = nullptr;       // device
    Ctl* ctl_host = nullptr;  // pinned
    int count_stats = 0, time_kernels = 1, blas_leaf = 4;
    float4* per_sample = nullptr; size_t per_sample_cap = 0;   // moments mode: per-sample radiance sums of the running pass (grow-only)
    int flat_max_entries = 16, scene_flat = 0, scene_has_mesh = 0;   // worlds of <= flat_max_entries entries without a mesh are traced by the flat kernels (trace_flat)
    int pixel_major = 1;  // path order of k_generate: tiles of 32 neighbouring pixels, all their samples consecutively (1, generate_path); all samples of one pixel consecutively (2); sample-major (0)
    int tri_pretest = RTX_TRI_PRETEST;      // mesh worlds: float32 pre-test records for the TRI phase (takes effect at the next rtx_scene_upload)
    int fuse_drain = RTX_FUSE_DRAIN_DEFAULT;   // hierarchy worlds: the last iterations of a pass run as ONE barrier-free persistent launch (k_drain) once at most this many rays are left (0 = off)
    int shade_direct = RTX_SHADE_DIRECT_DEFAULT;   // hierarchy worlds: k_shade in stream order instead of through material-sorted queues
    int simple_below = RTX_SIMPLE_BELOW_DEFAULT;   // hierarchy worlds: iterations of the drain with at most this many rays run the one-thread-per-ray trace kernels (0 = never)
    int tlas_flat_max = RTX_TLAS_FLAT_MAX;   // mesh worlds with at most this many bounded entries: top level as a per-ray sorted list (0 = hierarchy)
    int bvh_device = 1;   // mesh BLAS construction on the device (rtx_bvh_gpu.cuh); 0 = host builder (rtx_bvh.hpp), kept for A/B
    double ms_upload_blas = 0, ms_upload_total = 0, ms_upload_ranks = 0;
    int blas_depth = 0, built_on_device = 0;
    rtx_stats stats{};
    double env_total = 0;
    // scene summary
    uint32_t tlas_nodes = 0, blas_nodes = 0, n_entries = 0, n_tris = 0;
    std::vector<cudaEvent_t> events;
    int num_sms = 0;
    int* batch_cursor = nullptr;  // job cursor of k_trace_closest
    int trace_grid_sky = 0, trace_grid_lucy = 0;   // grids of the lean variants of the persistent trace kernels
    int drain_grid_sky = 0, drain_grid_lucy = 0;   // ... and of the lean drain kernels (k_drain)
    int* trace_spill = nullptr;   // global overflow columns of the trace kernels' shared-memory stacks
    int* trace_spill2 = nullptr;  // the same for k_connect (it may run beside k_extend)
    int trace_grid = 0;           // persistent grid: SMs x resident blocks
    void* geom_arena = nullptr;   // nodes + triangles + spheres + quads in one allocation: the L2 persisting window
    size_t geom_bytes = 0;
    int l2_persist = 0;  // measured on cornell-lucy: 1061 -> 1073 Mrays/s only, so off by default (it changes a process-wide device limit)
    cudaStream_t window_stream = nullptr; bool window_set = false;
    double ms_resolve = 0, ms_reduce = 0;
    // the test-order ranks the device derived for the last scene (rtx_rank_gpu.cuh), kept under a content hash of its triangle arrays
    int* rank_cache = nullptr; size_t rank_cache_n = 0; unsigned long long rank_cache_key = 0; bool rank_cache_valid = false, ranks_from_device = false;
    unsigned long long* hash_dev = nullptr;
    int pool_has_shadow = 0, pool_hit_bytes = 0;   // what the allocated pool was sized for
    // rtx_create_multi: the context the caller holds is device_ids[0]'s; the other devices' contexts hang off it
    std::vector<rtx_ctx*> peers;    // [n - 1], empty for a single-device context
    std::vector<ncclComm_t> comms;  // [n], rank 0 = this context
    bool is_peer = false;
};

static int32_t fail(rtx_ctx* ctx, int32_t code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return fail(ctx, RTX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

static void free_scene(rtx_ctx* ctx) {   // the slabs are kept for the next upload (grow-only); rtx_destroy releases them
    ctx->have_scene = false;
}
static void free_pool(rtx_ctx* ctx) {
    for (void* p : ctx->pool_allocs) cudaFree(p);
    ctx->pool_allocs.clear();
    ctx->pool = Pool{};
    ctx->pool_has_shadow = 0; ctx->pool_hit_bytes = 0;
}

struct Scratch {  // RAII device scratch
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <class T>
    cudaError_t in(T** dev, const T* host, size_t n, cudaStream_t st) {
        cudaError_t e = cudaMalloc((void**)dev, std::max<size_t>(n * sizeof(T), 16));
        if (e != cudaSuccess) return e;
        ptrs.push_back(*dev);
        if (host && n) e = cudaMemcpyAsync(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice, st);
        return e;
    }
    template <class T>
    cudaError_t out(T** dev, T* host, size_t n) {
        *dev = nullptr;
        if (!host) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)dev, std::max<size_t>(n * sizeof(T), 16));
        if (e == cudaSuccess) ptrs.push_back(*dev);
        return e;
    }
};

extern "C" {

int32_t rtx_abi_version(void) { return RTX_ABI_VERSION; }

const char* rtx_last_error(const rtx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

// Which kernel variant covers the uploaded scene and the camera: 0 = RTX_F_ALL, 1 = RTX_FV_LUCY, 2 = RTX_FV_SKY, 3 = RTX_FV_BOX,
// 4 = RTX_FV_CORNELL (flat kernels and k_shade only; the persistent trace kernels run it with all features)
static int lean_variant(const rtx_ctx* ctx) {
    if (!ctx->lean_flat) return 0;
    const unsigned need = ctx->feat_mask | ((ctx->have_camera && (ctx->C.camera_motion || ctx->C.free_camera)) ? RTX_F_CAM_SLOW : 0u);
    if (!(need & ~RTX_FV_LUCY)) return 1;
    if (!(need & ~RTX_FV_SKY)) return 2;
    if (!(need & ~RTX_FV_BOX)) return 3;
    if (!(need & ~RTX_FV_CORNELL)) return 4;
    return 0;
}

int32_t rtx_create(int32_t device_id, rtx_ctx** out) {
    if (!out) return fail(nullptr, RTX_ERR_INVALID, "rtx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, RTX_ERR_CUDA, "rtx_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (device_id < 0 || device_id >= n) return fail(nullptr, RTX_ERR_INVALID, "rtx_create: device %d out of range [0,%d)", device_id, n);
    rtx_ctx* ctx = new rtx_ctx();
    ctx->device = device_id;
    int prLo = 0, prHi = 0;
    cudaSetDevice(device_id);
    cudaDeviceGetStreamPriorityRange(&prLo, &prHi);   // numerically lower = higher priority
    if ((e = cudaSetDevice(device_id)) != cudaSuccess || (e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prHi)) != cudaSuccess ||
        (e = cudaMalloc((void**)&ctx->ctl, sizeof(Ctl))) != cudaSuccess || (e = cudaMallocHost((void**)&ctx->ctl_host, sizeof(Ctl))) != cudaSuccess) {
        fail(nullptr, RTX_ERR_CUDA, "rtx_create: %s", cudaGetErrorString(e));
        delete ctx;
        return RTX_ERR_CUDA;
    }
    ctx->own_stream = ctx->stream;
    if ((e = cudaStreamCreateWithPriority(&ctx->connect_stream, cudaStreamNonBlocking, prLo)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_shaded, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_connected[0], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_connected[1], cudaEventDisableTiming)) != cudaSuccess) {
        fail(nullptr, RTX_ERR_CUDA, "rtx_create: %s", cudaGetErrorString(e));
        rtx_destroy(ctx);
        return RTX_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device_id);
    if (ctx->num_sms <= 0 || cudaMalloc((void**)&ctx->batch_cursor, sizeof(int)) != cudaSuccess) {
        fail(nullptr, RTX_ERR_CUDA, "rtx_create: device query / allocation failed");
        rtx_destroy(ctx);
        return RTX_ERR_CUDA;
    }
    for (int group = 0; group < 3; group++) {   // persistent trace kernels: opt in to the large dynamic shared-memory pool, size the grid to one resident wave
        // group 0: all-features kernels; the lean variants (RTX_FV_*) need fewer registers and smaller slots: their own block count, pool size and grid (1: SKY, 2: LUCY)
        const int smem = group == 0 ? (int)RTX_TRACE_SMEM_BYTES : group == 1 ? (int)RTX_TRACE_SMEM_BYTES_OF(RTX_FV_SKY) : (int)RTX_TRACE_SMEM_BYTES_OF(RTX_FV_LUCY);
        int occ = 0, minOcc = 1 << 30;
        const std::vector<const void*> kernels = group == 0
            ? std::vector<const void*>{(const void*)k_extend<false>, (const void*)k_extend<true>, (const void*)k_extend<false, true>, (const void*)k_connect<false>,
                                 (const void*)k_connect<true>, (const void*)k_trace_closest<>, (const void*)k_bounce<false>, (const void*)k_bounce<true>,
                                 (const void*)k_bounce<false, true>, (const void*)k_drain<false>, (const void*)k_drain<true>}
            : group == 1 ? std::vector<const void*>{(const void*)k_extend<false, false, RTX_FV_SKY>, (const void*)k_connect<false, RTX_FV_SKY>, (const void*)k_trace_closest<RTX_FV_SKY>}
                         : std::vector<const void*>{(const void*)k_extend<false, false, RTX_FV_LUCY>, (const void*)k_connect<false, RTX_FV_LUCY>, (const void*)k_trace_closest<RTX_FV_LUCY>};
        // developer knob: shared-memory carve-out in KB (the rest of the 256 KB array is L1); fewer resident blocks, more L1
        const char* carveEnv = getenv("RTX_TRACE_CARVEOUT_KB");
        const int carveKB = carveEnv ? atoi(carveEnv) : 0;
        for (const void* k : kernels) {
            if (carveKB > 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(100, carveKB * 100 / 228));
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess ||
                (e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, RTX_TRACE_THREADS, smem)) != cudaSuccess || occ < 1) {
                fail(nullptr, RTX_ERR_CUDA, "rtx_create: trace kernel setup failed (%s, occupancy %d)", cudaGetErrorString(e), occ);
                rtx_destroy(ctx);
                return RTX_ERR_CUDA;
            }
            minOcc = std::min(minOcc, occ);
        }
        if (carveKB > 0) minOcc = std::max(1, std::min(minOcc, (int)((size_t)carveKB * 1024 / (smem + 1024))));
        (group == 0 ? ctx->trace_grid : group == 1 ? ctx->trace_grid_sky : ctx->trace_grid_lucy) = ctx->num_sms * minOcc;
        if (getenv("RTX_DEBUG_BATCH")) fprintf(stderr, "[rtx] trace kernels (%s): %d blocks/SM, %d B dynamic smem per block, grid %d\n", group == 0 ? "full" : group == 1 ? "sky" : "lucy", minOcc, smem, ctx->num_sms * minOcc);
    }
    {   // the lean drain kernels: the lean pools with their own register budget, hence their own resident wave
        const void* dk[2] = {(const void*)k_drain<false, RTX_FV_SKY>, (const void*)k_drain<false, RTX_FV_LUCY>};
        const int dsm[2] = {(int)RTX_TRACE_SMEM_BYTES_OF(RTX_FV_SKY), (int)RTX_TRACE_SMEM_BYTES_OF(RTX_FV_LUCY)};
        for (int k = 0; k < 2; k++) {
            int occ = 0;
            if ((e = cudaFuncSetAttribute(dk[k], cudaFuncAttributeMaxDynamicSharedMemorySize, dsm[k])) != cudaSuccess ||
                (e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dk[k], RTX_TRACE_THREADS, dsm[k])) != cudaSuccess || occ < 1) {
                fail(nullptr, RTX_ERR_CUDA, "rtx_create: drain kernel setup failed (%s, occupancy %d)", cudaGetErrorString(e), occ);
                rtx_destroy(ctx);
                return RTX_ERR_CUDA;
            }
            (k == 0 ? ctx->drain_grid_sky : ctx->drain_grid_lucy) = ctx->num_sms * occ;
        }
    }
    {
        size_t spillInts = std::max((size_t)ctx->trace_grid * RTX_TRACE_SLOTS, (size_t)std::max(ctx->trace_grid_sky, ctx->trace_grid_lucy) * RTX_TRACE_SLOTS_LEAN) * (RTX_STACK_SIZE - RTX_SMEM_STACK);
        if ((e = cudaMalloc((void**)&ctx->trace_spill, spillInts * sizeof(int))) != cudaSuccess ||
            (e = cudaMalloc((void**)&ctx->trace_spill2, spillInts * sizeof(int))) != cudaSuccess) {
            fail(nullptr, RTX_ERR_CUDA, "rtx_create: %s", cudaGetErrorString(e));
            rtx_destroy(ctx);
            return RTX_ERR_CUDA;
        }
    }
    *out = ctx;
    return RTX_OK;
}
int32_t rtx_mesh_test_order(rtx_ctx* ctx, int32_t* rank, int64_t n) {
    if (!ctx || !rank || n < 0) return RTX_ERR_INVALID;
    if (!ctx->have_scene || !ctx->rank_cache_valid || !ctx->ranks_from_device) return fail(ctx, RTX_ERR_STATE, "rtx_mesh_test_order: the last upload did not derive a test order on the device (tri_rank given, host build, or no mesh)");
    if (n != (int64_t)ctx->rank_cache_n) return fail(ctx, RTX_ERR_INVALID, "rtx_mesh_test_order: the scene has %lld triangles, not %lld", (long long)ctx->rank_cache_n, (long long)n);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(rank, ctx->rank_cache, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return RTX_OK;
}

}  // extern "C"

// ---- NCCL, loaded on demand (rtx_create_multi) -------------------------------------------------------------------------
// dlopen instead of a link-time dependency: a process that already holds an NCCL (torch bundles its own libnccl.so.2) must keep
// using that one — two copies of a library with the same soname in one process share one set of symbols.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
static NcclApi g_nccl;
static bool load_nccl(std::string& why) {
    if (g_nccl.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
    NcclApi a;
    a.handle = h;
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) { ok = false; why = std::string("libnccl lacks ") + name; } return p; };
    a.CommInitAll = (decltype(a.CommInitAll))sym("ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.Reduce = (decltype(a.Reduce))sym("ncclReduce");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
    if (!ok) return false;
    g_nccl = a;
    return true;
}

extern "C" int32_t rtx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int32_t rtx_create_multi(const int32_t* device_ids, int32_t n, rtx_ctx** out) {
    if (!out) return fail(nullptr, RTX_ERR_INVALID, "rtx_create_multi: out is NULL");
    *out = nullptr;
    if (!device_ids || n < 1) return fail(nullptr, RTX_ERR_INVALID, "rtx_create_multi: need at least one device id");
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
            if (device_ids[a] == device_ids[b]) return fail(nullptr, RTX_ERR_INVALID, "rtx_create_multi: device %d listed twice", device_ids[a]);
    rtx_ctx* root = nullptr;
    int32_t rc = rtx_create(device_ids[0], &root);
    if (rc != RTX_OK || n == 1) { *out = root; return rc; }
    for (int g = 1; g < n; g++) {
        rtx_ctx* p = nullptr;
        rc = rtx_create(device_ids[g], &p);
        if (rc != RTX_OK) { rtx_destroy(root); return rc; }   // g_create_error holds the message
        p->is_peer = true;
        root->peers.push_back(p);
    }
    std::string why;
    if (!load_nccl(why)) { rtx_destroy(root); return fail(nullptr, RTX_ERR_UNSUPPORTED, "rtx_create_multi: NCCL is required for more than one device (%s)", why.c_str()); }
    root->comms.assign(n, nullptr);
    std::vector<int> devs(device_ids, device_ids + n);
    ncclResult_t nr = g_nccl.CommInitAll(root->comms.data(), n, devs.data());
    if (nr != ncclSuccess) {
        root->comms.clear();
        rtx_destroy(root);
        return fail(nullptr, RTX_ERR_CUDA, "rtx_create_multi: ncclCommInitAll failed: %s", g_nccl.GetErrorString(nr));
    }
    *out = root;
    return RTX_OK;
}

extern "C" {

int32_t rtx_set_stream(rtx_ctx* ctx, void* cuda_stream) {
    if (!ctx) return RTX_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return RTX_OK;
}

int32_t rtx_destroy(rtx_ctx* ctx) {
    if (!ctx) return RTX_OK;
    for (size_t g = 0; g < ctx->comms.size(); g++)
        if (ctx->comms[g] && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comms[g]);
    ctx->comms.clear();
    for (rtx_ctx* p : ctx->peers) rtx_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    // the context's OWN streams: a caller-owned stream adopted with rtx_set_stream may already be gone (a torch stream released before
    // the context), and synchronizing a dangling handle is undefined
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    if (ctx->connect_stream) cudaStreamSynchronize(ctx->connect_stream);
    free_scene(ctx);
    free_pool(ctx);
    ctx->scene_slab.release(); ctx->work_slab.release();
    if (ctx->rgba_dev) cudaFree(ctx->rgba_dev);
    if (ctx->per_sample) cudaFree(ctx->per_sample);
    if (ctx->rank_cache) cudaFree(ctx->rank_cache);
    if (ctx->hash_dev) cudaFree(ctx->hash_dev);
    if (ctx->accum) cudaFree(ctx->accum);
    if (ctx->accum_sq) cudaFree(ctx->accum_sq);
    if (ctx->ctl) cudaFree(ctx->ctl);
    if (ctx->batch_cursor) cudaFree(ctx->batch_cursor);
    if (ctx->trace_spill) cudaFree(ctx->trace_spill);
    if (ctx->trace_spill2) cudaFree(ctx->trace_spill2);
    if (ctx->ev_shaded) cudaEventDestroy(ctx->ev_shaded);
    for (auto ev : ctx->ev_connected) if (ev) cudaEventDestroy(ev);
    if (ctx->connect_stream) cudaStreamDestroy(ctx->connect_stream);
    if (ctx->ctl_host) cudaFreeHost(ctx->ctl_host);
    for (auto ev : ctx->events) cudaEventDestroy(ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return RTX_OK;
}

static int32_t set_option_single(rtx_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return RTX_ERR_INVALID;
    std::string k(key);
    if (k == "pool_paths") {
        if (value < 1024 || value > (1ll << 26)) return fail(ctx, RTX_ERR_INVALID, "pool_paths out of range");
        if (value != ctx->pool_paths) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); free_pool(ctx); }
        ctx->pool_paths = value;
    } else if (k == "count_stats") ctx->count_stats = (int)value;  // bit 0: extend kernel, bit 1: connect kernel
    else if (k == "time_kernels") ctx->time_kernels = value != 0;
    else if (k == "l2_persist") { ctx->l2_persist = value != 0; ctx->window_set = false; }  // L2 persisting window over the scene geometry (default off)
    else if (k == "pixel_major") ctx->pixel_major = value < 0 ? 0 : value > 2 ? 2 : (int)value;   // 0 sample-major, 1 tiles of 32 pixels (default), 2 pixel-major
    else if (k == "fuse_flat") ctx->fuse_flat = value != 0;
    else if (k == "shade_split") ctx->shade_split = value != 0;
    else if (k == "fuse_tree") ctx->fuse_tree = value != 0;
    else if (k == "lean" || k == "lean_flat") ctx->lean_flat = value != 0;
    else if (k == "pretest_bare") {   // takes effect at the next rtx_scene_upload
        if (value < 0 || value > 64) return fail(ctx, RTX_ERR_INVALID, "pretest_bare must be in 0..64");
        ctx->pretest_bare = (int)value;
    }
    else if (k == "shade_direct") ctx->shade_direct = value != 0;
    else if (k == "fuse_drain") ctx->fuse_drain = (int)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 26));
    else if (k == "drop_caches") ctx->rank_cache_valid = false;   // forget what earlier uploads left behind (the cached test-order ranks): the next upload is a cold one
    else if (k == "tri_pretest") ctx->tri_pretest = value != 0;
    else if (k == "simple_below") ctx->simple_below = (int)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 30));
#ifdef RTX_CHECKED
    else if (k == "checked_selftest") {   // negative control of the checked build: record `value` violations of kind 7 through the same macro the kernels use
        cudaSetDevice(ctx->device);
        k_check_selftest<<<1, 32, 0, ctx->stream>>>((int)value);
        cudaStreamSynchronize(ctx->stream);
    }
#endif
    else if (k == "overlap_connect") ctx->overlap_connect = value != 0;   // k_connect on its own stream beside the next iteration (default on)
    else if (k == "flat_max_entries") {   // 0 = always traverse the hierarchy
        if (value < 0 || value > 64) return fail(ctx, RTX_ERR_INVALID, "flat_max_entries must be in 0..64");
        ctx->flat_max_entries = (int)value;
        ctx->scene_flat = ctx->have_scene && !ctx->scene_has_mesh && ctx->S.n_entries <= ctx->flat_max_entries;
    }
    else if (k == "bvh_device") ctx->bvh_device = value != 0;  // takes effect at the next rtx_scene_upload
    else if (k == "tlas_flat_max") {   // takes effect at the next rtx_scene_upload
        if (value < 0 || value > RTX_SMEM_STACK) return fail(ctx, RTX_ERR_INVALID, "tlas_flat_max must be in 0..%d", RTX_SMEM_STACK);
        ctx->tlas_flat_max = (int)value;
    }
    else if (k == "blas_leaf") {  // triangles per BLAS leaf (1..8); takes effect at the next rtx_scene_upload
        if (value < 1 || value > 8) return fail(ctx, RTX_ERR_INVALID, "blas_leaf must be in 1..8");
        ctx->blas_leaf = (int)value;
    }
    else return fail(ctx, RTX_ERR_INVALID, "unknown option '%s'", key);
    return RTX_OK;
}

// ---- scene upload -------------------------------------------------------------------------------------------------
static Box prim_box(const rtx_scene_desc* d, int kind, int idx) {
    Box b;
    b.reset();
    if (kind == RTX_GEOM_SPHERE) {
        double r = std::fabs(d->sph_radius[idx]);
        for (int a = 0; a < 3; a++) {
            double c0 = d->sph_center[3 * idx + a], c1 = c0 + d->sph_velocity[3 * idx + a];
            b.lo[a] = std::min(c0, c1) - r;
            b.hi[a] = std::max(c0, c1) + r;
        }
    } else if (kind == RTX_GEOM_QUAD) {
        for (int corner = 0; corner < 4; corner++) {
            double p[3];
            for (int a = 0; a < 3; a++)
                p[a] = d->quad_q[3 * idx + a] + ((corner & 1) ? d->quad_u[3 * idx + a] : 0.0) + ((corner & 2) ? d->quad_v[3 * idx + a] : 0.0);
            b.grow(p);
        }
    } else if (kind == RTX_GEOM_TRIANGLE) {
        b.grow(d->tri_v0 + 3 * idx); b.grow(d->tri_v1 + 3 * idx); b.grow(d->tri_v2 + 3 * idx);
    } else if (kind == RTX_GEOM_CIRCLE) {   // center -/+ (r, r, r), rt/circle.go:22-26
        const double r = d->circle_radius[idx];
        for (int a = 0; a < 3; a++) { b.lo[a] = d->circle_center[3 * idx + a] - std::fabs(r); b.hi[a] = d->circle_center[3 * idx + a] + std::fabs(r); }
    } else {
        for (int a = 0; a < 3; a++) { b.lo[a] = -INFINITY; b.hi[a] = INFINITY; }
    }
    // a hit may sit a few ulps off a flat primitive's plane: pad like the reference (rt/aabb.go:117-128)
    for (int a = 0; a < 3; a++)
        if (b.hi[a] - b.lo[a] < 1e-4) { b.lo[a] -= 1e-4; b.hi[a] += 1e-4; }
    return b;
}
static Box xform_box(const rtx_scene_desc* d, Box b, int xfBegin, int xfCount) {  // innermost first, like the wrappers' constructors
    for (int k = xfCount - 1; k >= 0; k--) {
        int x = xfBegin + k;
        const double* a = d->xf_a + 3 * x;
        if (!b.finite()) return b;
        if (d->xf_type[x] == RTX_XF_TRANSLATE) {
            for (int i = 0; i < 3; i++) { b.lo[i] += a[i]; b.hi[i] += a[i]; }
        } else if (d->xf_type[x] == RTX_XF_ROTATE_Y) {
            Box r;
            r.reset();
            for (int c = 0; c < 8; c++) {
                double px = (c & 1) ? b.hi[0] : b.lo[0], py = (c & 2) ? b.hi[1] : b.lo[1], pz = (c & 4) ? b.hi[2] : b.lo[2];
                double q[3] = {a[1] * px + a[0] * pz, py, -a[0] * px + a[1] * pz};
                r.grow(q);
            }
            b = r;
        } else {
            for (int i = 0; i < 3; i++) {
                double l = b.lo[i] * a[i], h = b.hi[i] * a[i];
                b.lo[i] = std::min(l, h); b.hi[i] = std::max(l, h);
            }
        }
        // rotation / scaling round: keep the box conservative
        for (int i = 0; i < 3; i++) {
            double m = std::max(std::fabs(b.lo[i]), std::fabs(b.hi[i])) * 1e-12 + 1e-300;
            b.lo[i] -= m; b.hi[i] += m;
        }
    }
    return b;
}

static int32_t scene_upload_single(rtx_ctx* ctx, const rtx_scene_desc* d) {
    if (!ctx || !d) return RTX_ERR_INVALID;
    if (d->abi_version != RTX_ABI_VERSION) return fail(ctx, RTX_ERR_INVALID, "scene abi_version %u != %d", d->abi_version, RTX_ABI_VERSION);
    const auto tUpload0 = std::chrono::steady_clock::now();
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    free_scene(ctx);
    ctx->ranks_from_device = false;
    DevScene S{};

    // ---- validation
    auto bad = [&](const char* what) { return fail(ctx, RTX_ERR_INVALID, "scene: %s", what); };
    for (int i = 0; i < d->n_textures; i++) {
        if (d->tex_type[i] == RTX_TEX_CHECKER) {
            if (d->tex_even[i] < 0 || d->tex_even[i] >= d->n_textures || d->tex_odd[i] < 0 || d->tex_odd[i] >= d->n_textures) return bad("checker child out of range");
        } else if (d->tex_type[i] == RTX_TEX_NOISE) {
            if (d->tex_even[i] < 0 || d->tex_even[i] >= d->n_perlin || !d->perlin_vec || !d->perlin_perm) return bad("noise texture without a Perlin table");
        } else if (d->tex_type[i] == RTX_TEX_IMAGE) {
            const int im = d->tex_even[i];
            if (im < 0 || im >= d->n_images || !d->image_rgb || !d->image_width || !d->image_height || !d->image_offset || d->image_width[im] <= 0 || d->image_height[im] <= 0)
                return bad("image texture without image data");
        } else if (d->tex_type[i] != RTX_TEX_SOLID) return fail(ctx, RTX_ERR_UNSUPPORTED, "texture type %d is outside the device path", d->tex_type[i]);
    }
    if (d->n_materials >= (1 << 28)) return bad("too many materials");   // a hit record packs the material index into 28 bits
    for (int i = 0; i < d->n_materials; i++) {
        int t = d->mat_type[i];
        if (t < RTX_MAT_LAMBERTIAN || t > RTX_MAT_ISOTROPIC) return fail(ctx, RTX_ERR_UNSUPPORTED, "material type %d is outside the device path", t);
        if ((t == RTX_MAT_LAMBERTIAN || t == RTX_MAT_DIFFUSE_LIGHT || t == RTX_MAT_ISOTROPIC) && (d->mat_tex[i] < 0 || d->mat_tex[i] >= d->n_textures))
            return bad("material texture out of range");
    }
    {   // Plane.Hit does not write rec.U / rec.V (rt/plane.go:24-42): an image texture there would read whatever an earlier,
        // farther hit of the same query left in the record — order-dependent in the reference, so it stays outside the device path
        std::function<bool(int, int)> hasImage = [&](int t, int depth) {
            if (t < 0 || t >= d->n_textures || depth > 8) return false;
            if (d->tex_type[t] == RTX_TEX_IMAGE) return true;
            return d->tex_type[t] == RTX_TEX_CHECKER && (hasImage(d->tex_even[t], depth + 1) || hasImage(d->tex_odd[t], depth + 1));
        };
        for (int i = 0; i < d->n_planes; i++) {
            const int m = d->plane_mat[i];
            if (m >= 0 && m < d->n_materials && hasImage(d->mat_tex[m], 0))
                return fail(ctx, RTX_ERR_UNSUPPORTED, "plane %d: an ImageTexture on a Plane is outside the device path (Plane.Hit leaves U/V unwritten)", i);
        }
    }
    auto matok = [&](const int32_t* m, int n) { for (int i = 0; i < n; i++) if (m[i] < 0 || m[i] >= d->n_materials) return false; return true; };
    if (!matok(d->sph_mat, d->n_spheres) || !matok(d->quad_mat, d->n_quads) || !matok(d->tri_mat, d->n_tris) || !matok(d->plane_mat, d->n_planes) || !matok(d->circle_mat, d->n_circles) ||
        !matok(d->vol_mat, d->n_volumes))
        return bad("material index out of range");
    auto isPrim = [](int kind) { return (kind >= RTX_GEOM_SPHERE && kind <= RTX_GEOM_PLANE) || kind == RTX_GEOM_CIRCLE; };
    auto primCount = [&](int kind) {
        return kind == RTX_GEOM_SPHERE ? d->n_spheres : kind == RTX_GEOM_QUAD ? d->n_quads : kind == RTX_GEOM_TRIANGLE ? d->n_tris : kind == RTX_GEOM_CIRCLE ? d->n_circles : d->n_planes;
    };
    for (int i = 0; i < d->n_list_items; i++) {
        int k = d->list_item_kind[i];
        if (!isPrim(k) || d->list_item_index[i] < 0 || d->list_item_index[i] >= primCount(k)) return bad("list item out of range");
    }
    for (int g = 0; g < d->n_groups; g++) {
        int lim = d->group_kind[g] == RTX_GEOM_LIST ? d->n_list_items : d->n_tris;
        if ((d->group_kind[g] != RTX_GEOM_LIST && d->group_kind[g] != RTX_GEOM_MESH) || d->group_begin[g] < 0 || d->group_count[g] < 0 ||
            d->group_begin[g] + d->group_count[g] > lim)
            return bad("group range invalid");
    }
    for (int x = 0; x < d->n_xforms; x++)
        if (d->xf_type[x] < RTX_XF_TRANSLATE || d->xf_type[x] > RTX_XF_SCALE) return fail(ctx, RTX_ERR_UNSUPPORTED, "transform op %d is outside the device path", d->xf_type[x]);
    for (int e = 0; e < d->n_entries; e++) {
        int k = d->entry_geom_kind[e], gi = d->entry_geom_index[e];
        if (k < RTX_GEOM_SPHERE || k > RTX_GEOM_CIRCLE) return bad("entry kind invalid");
        if (isPrim(k) ? (gi < 0 || gi >= primCount(k)) : (gi < 0 || gi >= d->n_groups || d->group_kind[gi] != k)) return bad("entry geometry out of range");
        if (d->entry_xf_count[e] < 0 || d->entry_xf_begin[e] < 0 || d->entry_xf_begin[e] + d->entry_xf_count[e] > d->n_xforms) return bad("entry transform range invalid");
        if (d->entry_volume[e] >= d->n_volumes) return bad("entry volume out of range");
        if (d->entry_volume[e] >= 0 && (k == RTX_GEOM_MESH || k == RTX_GEOM_PLANE))
            return fail(ctx, RTX_ERR_UNSUPPORTED, "Volume boundary must be a primitive or a HittableList of primitives");
    }
    for (int i = 0; i < d->n_lights; i++)
        if (d->light_quad[i] >= d->n_quads) return bad("light quad out of range");

    // ---- textures, materials
    std::vector<DTexture> texs(d->n_textures);
    for (int i = 0; i < d->n_textures; i++) {
        texs[i].type = d->tex_type[i]; texs[i].even = d->tex_even ? d->tex_even[i] : -1; texs[i].odd = d->tex_odd ? d->tex_odd[i] : -1;
        texs[i].inv_scale = d->tex_inv_scale ? d->tex_inv_scale[i] : 0;
        for (int c = 0; c < 3; c++) texs[i].color[c] = (float)d->tex_color[3 * i + c];
    }
    std::vector<DMaterial> mats(d->n_materials);
    ctx->mat_kinds = 1u << Q_MISS;
    for (int i = 0; i < d->n_materials; i++) {
        const int mt = d->mat_type[i];   // the queue k_extend bins a hit on this material into (rtx_kernels.cuh: ExtendPolicy::retire)
        ctx->mat_kinds |= 1u << (mt == RTX_MAT_LAMBERTIAN ? Q_LAMBERTIAN : mt == RTX_MAT_METAL ? Q_METAL : mt == RTX_MAT_DIELECTRIC ? Q_DIELECTRIC : mt == RTX_MAT_DIFFUSE_LIGHT ? Q_LIGHT : Q_ISOTROPIC);
        mats[i].type = d->mat_type[i]; mats[i].tex = d->mat_tex[i]; mats[i].fuzz = d->mat_fuzz[i]; mats[i].ior = d->mat_ior[i]; mats[i].pad = 0;
        for (int c = 0; c < 3; c++) mats[i].albedo[c] = (float)d->mat_albedo[3 * i + c];
    }
    // ---- primitives in float64
    std::vector<double> sph((size_t)8 * d->n_spheres);
    std::vector<int> sphMat(d->sph_mat, d->sph_mat + d->n_spheres);
    for (int i = 0; i < d->n_spheres; i++) {
        for (int a = 0; a < 3; a++) { sph[8 * i + a] = d->sph_center[3 * i + a]; sph[8 * i + 3 + a] = d->sph_velocity[3 * i + a]; }
        sph[8 * i + 6] = std::fmax(0.0, d->sph_radius[i]);  // rt/sphere.go:18
        sph[8 * i + 7] = 0;
    }
    std::vector<double> quads((size_t)16 * d->n_quads);
    std::vector<int> quadMat(d->quad_mat, d->quad_mat + d->n_quads);
    for (int i = 0; i < d->n_quads; i++) {  // NewQuad rt/quad.go:16-33, same operation order
        const double *Q = d->quad_q + 3 * i, *u = d->quad_u + 3 * i, *v = d->quad_v + 3 * i;
        double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        double l2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        double l = std::sqrt(l2);
        double nn[3] = {n[0], n[1], n[2]};
        if (l != 0) { double inv = 1 / l; nn[0] = inv * n[0]; nn[1] = inv * n[1]; nn[2] = inv * n[2]; }
        double D = nn[0] * Q[0] + nn[1] * Q[1] + nn[2] * Q[2];
        double wk = 1.0 / l2;
        double* o = quads.data() + 16 * (size_t)i;
        for (int a = 0; a < 3; a++) { o[a] = Q[a]; o[3 + a] = u[a]; o[6 + a] = v[a]; o[9 + a] = wk * n[a]; o[12 + a] = nn[a]; }
        o[15] = D;
    }
    std::vector<double> planes((size_t)8 * d->n_planes);
    std::vector<int> planeMat(d->plane_mat, d->plane_mat + d->n_planes);
    for (int i = 0; i < d->n_planes; i++)
        for (int a = 0; a < 3; a++) { planes[8 * i + a] = d->plane_point[3 * i + a]; planes[8 * i + 3 + a] = d->plane_normal[3 * i + a]; }

    std::vector<double> circles((size_t)8 * d->n_circles);
    std::vector<int> circleMat(d->circle_mat, d->circle_mat + d->n_circles);
    for (int i = 0; i < d->n_circles; i++) {  // NewCircle rt/circle.go:14-31
        const double *c = d->circle_center + 3 * i, *n = d->circle_normal + 3 * i;
        double* o = circles.data() + 8 * (size_t)i;
        for (int a = 0; a < 3; a++) { o[a] = c[a]; o[3 + a] = n[a]; }
        o[6] = d->circle_radius[i];
        o[7] = n[0] * c[0] + n[1] * c[1] + n[2] * c[2];
    }
    std::vector<double> perlinVec(d->n_perlin > 0 ? d->perlin_vec : nullptr, d->n_perlin > 0 ? d->perlin_vec + (size_t)768 * d->n_perlin : nullptr);
    std::vector<int> perlinPerm(d->n_perlin > 0 ? d->perlin_perm : nullptr, d->n_perlin > 0 ? d->perlin_perm + (size_t)768 * d->n_perlin : nullptr);

    std::vector<int4> flatSimple;   // filled after the entries are known
    std::vector<int> flatComplex;
    std::vector<float4> imgRgb;
    std::vector<int4> imgDim(d->n_images);
    for (int i = 0; i < d->n_images; i++) {
        const size_t first = imgRgb.size(), npx = (size_t)d->image_width[i] * d->image_height[i];
        const double* src = d->image_rgb + 3 * (size_t)d->image_offset[i];
        imgDim[i] = make_int4(d->image_width[i], d->image_height[i], (int)(first & 0xffffffffu), (int)(first >> 32));
        for (size_t k = 0; k < npx; k++) imgRgb.push_back(make_float4((float)src[3 * k], (float)src[3 * k + 1], (float)src[3 * k + 2], 0.f));
    }

    // ---- triangles: meshes are permuted into BLAS leaf order; loose triangles (world entries / Box-list items) follow
    std::vector<Node4> nodes;                       // host-built nodes: every BLAS with the host builder, and always the TLAS
    std::vector<double> tris;                       // host-built triangle records: meshes (host builder only), then loose ones
    std::vector<int4> triInfo;
    int meshTotal = 0;
    for (int g = 0; g < d->n_groups; g++)
        if (d->group_kind[g] == RTX_GEOM_MESH) meshTotal += d->group_count[g];
    if ((size_t)meshTotal + (size_t)d->n_tris >= (1u << 28)) return fail(ctx, RTX_ERR_UNSUPPORTED, "too many triangles");
    std::vector<int> looseOfDesc(d->n_tris, -1), looseList;
    auto noteLoose = [&](int kind, int idx) {
        if (kind == RTX_GEOM_TRIANGLE && looseOfDesc[idx] < 0) { looseOfDesc[idx] = meshTotal + (int)looseList.size(); looseList.push_back(idx); }
    };
    for (int i = 0; i < d->n_list_items; i++) noteLoose(d->list_item_kind[i], d->list_item_index[i]);
    for (int e = 0; e < d->n_entries; e++) noteLoose(d->entry_geom_kind[e], d->entry_geom_index[e]);
    const int totalTris = meshTotal + (int)looseList.size();
    auto pushTri = [&](int ti, int localId, int rank) {
        const double *v0 = d->tri_v0 + 3 * ti, *v1 = d->tri_v1 + 3 * ti, *v2 = d->tri_v2 + 3 * ti;
        double e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
        double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};  // rt/triangle.go:19-25
        double l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (l != 0) { double inv = 1 / l; n[0] = inv * n[0]; n[1] = inv * n[1]; n[2] = inv * n[2]; }
        for (int a = 0; a < 3; a++) tris.push_back(v0[a]);
        for (int a = 0; a < 3; a++) tris.push_back(e1[a]);
        for (int a = 0; a < 3; a++) tris.push_back(e2[a]);
        for (int a = 0; a < 3; a++) tris.push_back(n[a]);
        triInfo.push_back(make_int4(localId, d->tri_mat[ti], rank, 0));
    };
    std::vector<int> groupRoot(d->n_groups, -1), groupTriBase(d->n_groups, 0);
    std::vector<Box> groupBox(d->n_groups);
    for (auto& gb : groupBox) gb.reset();
    uint32_t blasNodes = 0;
    int maxBlasDepth = 0, tlasDepth = 0;
    const bool deviceBuild = ctx->bvh_device != 0 && meshTotal > 0;
    // device-built pieces (deviceBuild only): one node buffer per mesh, triangle records for all meshes
    struct DevBlas { Node4* nodes; int n; };
    std::vector<DevBlas> devBlas;
    double* dMeshTris = nullptr;
    int4* dTriInfo = nullptr;   // mesh triangles only (work slab); copied into the scene slab at the end
    cudaEvent_t evB0 = nullptr, evB1 = nullptr;
    if (deviceBuild) {
        // The flattened triangle arrays go to the device as they are; bounds, hierarchy and leaf-ordered records are built there.
        // All working memory is carved from one grow-only slab: repeated uploads never touch cudaMalloc / cudaFree.
        auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t nT = (size_t)d->n_tris;
        size_t buildBytes = 0, nodeBytes = 0;
        for (int g = 0; g < d->n_groups; g++) {
            if (d->group_kind[g] != RTX_GEOM_MESH || d->group_count[g] == 0) continue;
            size_t sb = 0;
            const char* w = "";
            rtxgpu::build_blas(nullptr, nullptr, nullptr, nullptr, nullptr, d->group_count[g], ctx->blas_leaf, 0, 0, nullptr, nullptr, nullptr, ctx->stream, nullptr, &w, nullptr, &sb);
            buildBytes = std::max(buildBytes, sb);
            nodeBytes += pad(((size_t)d->group_count[g] + 1) * sizeof(Node4));
        }
        size_t rankBytes = 0;
        if (!d->tri_rank)
            for (int g = 0; g < d->n_groups; g++)
                if (d->group_kind[g] == RTX_GEOM_MESH) rankBytes = std::max(rankBytes, rtxrank::scratch_bytes(d->group_count[g]));
        const size_t workTotal = 3 * pad(3 * nT * sizeof(double)) + 2 * pad(nT * sizeof(int)) + pad((size_t)RTX_TRI_D * meshTotal * sizeof(double)) +
                                 pad((size_t)meshTotal * sizeof(int4)) + nodeBytes + pad(buildBytes) + pad(rankBytes) + 4096;
        CU(ctx->work_slab.reserve(workTotal));
        Slab& W = ctx->work_slab;
        double *dV0 = (double*)W.take(3 * nT * sizeof(double)), *dV1 = (double*)W.take(3 * nT * sizeof(double)), *dV2 = (double*)W.take(3 * nT * sizeof(double));
        int *dMat = (int*)W.take(nT * sizeof(int)), *dRank = (int*)W.take(nT * sizeof(int));
        dMeshTris = (double*)W.take((size_t)RTX_TRI_D * meshTotal * sizeof(double));
        dTriInfo = (int4*)W.take((size_t)meshTotal * sizeof(int4));
        char* buildScratch = W.take(buildBytes);
        char* rankScratch = rankBytes ? W.take(rankBytes) : nullptr;
        ctx->ms_upload_ranks = 0;
        CU(cudaMemcpyAsync(dV0, d->tri_v0, 3 * nT * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dV1, d->tri_v1, 3 * nT * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dV2, d->tri_v2, 3 * nT * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dMat, d->tri_mat, nT * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        ctx->ranks_from_device = d->tri_rank == nullptr;
        if (d->tri_rank) {
            CU(cudaMemcpyAsync(dRank, d->tri_rank, nT * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        } else {
            // no ranks from the caller's Go tree: the canonical test order (the stable-sort restatement of rt/bvh.go:69-217; exact-tie
            // resolution only) is computed on the device — a few milliseconds instead of a 100 ms pointer-tree build on the host
            const auto tR0 = std::chrono::steady_clock::now();
            // the ranks are a function of the triangle arrays and the mesh ranges alone: an unchanged scene (a renderer re-created over the same
            // world, the bench's per-step upload) finds them under the content hash of what was just copied to the device
            if (!ctx->hash_dev) CU(cudaMalloc((void**)&ctx->hash_dev, sizeof(unsigned long long)));
            CU(cudaMemsetAsync(ctx->hash_dev, 0, sizeof(unsigned long long), ctx->stream));
            rtxrank::k_hash64<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>((const unsigned long long*)dV0, 3 * nT, 1, ctx->hash_dev);
            rtxrank::k_hash64<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>((const unsigned long long*)dV1, 3 * nT, 2, ctx->hash_dev);
            rtxrank::k_hash64<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>((const unsigned long long*)dV2, 3 * nT, 3, ctx->hash_dev);
            unsigned long long key = 0;
            CU(cudaMemcpyAsync(&key, ctx->hash_dev, sizeof key, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            for (int g = 0; g < d->n_groups; g++) key = (key ^ (unsigned long long)(unsigned)d->group_kind[g]) * 0x100000001B3ull + (unsigned long long)(unsigned)d->group_begin[g] * 31 + (unsigned)d->group_count[g];
            if (ctx->rank_cache_valid && ctx->rank_cache_n == nT && ctx->rank_cache_key == key && getenv("RTX_NO_RANK_CACHE") == nullptr) {
                CU(cudaMemcpyAsync(dRank, ctx->rank_cache, nT * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
            } else {
            CU(cudaMemsetAsync(dRank, 0, nT * sizeof(int), ctx->stream));
            for (int g = 0; g < d->n_groups; g++) {
                if (d->group_kind[g] != RTX_GEOM_MESH || d->group_count[g] == 0) continue;
                const int begin = d->group_begin[g], count = d->group_count[g];
                cudaError_t re = rtxrank::canonical_ranks(dV0 + 3 * (size_t)begin, dV1 + 3 * (size_t)begin, dV2 + 3 * (size_t)begin, count, dRank + begin, rankScratch, rankBytes, ctx->stream);
                if (re != cudaSuccess) return fail(ctx, RTX_ERR_CUDA, "device test-order ranks of mesh group %d failed: %s", g, cudaGetErrorString(re));
            }
            ctx->rank_cache_valid = false;
            if (ctx->rank_cache_n < nT) {
                if (ctx->rank_cache) cudaFree(ctx->rank_cache);
                ctx->rank_cache = nullptr; ctx->rank_cache_n = 0;
                if (cudaMalloc((void**)&ctx->rank_cache, nT * sizeof(int)) == cudaSuccess) ctx->rank_cache_n = nT;
                else cudaGetLastError();
            }
            if (ctx->rank_cache && ctx->rank_cache_n >= nT) {
                CU(cudaMemcpyAsync(ctx->rank_cache, dRank, nT * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
                ctx->rank_cache_n = nT; ctx->rank_cache_key = key; ctx->rank_cache_valid = true;
            }
            }
            ctx->ms_upload_ranks = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tR0).count();
        }
        CU(cudaStreamSynchronize(ctx->stream));
        while (ctx->events.size() < 4) {
            cudaEvent_t ev;
            CU(cudaEventCreate(&ev));
            ctx->events.push_back(ev);
        }
        evB0 = ctx->events[2]; evB1 = ctx->events[3];
        CU(cudaEventRecord(evB0, ctx->stream));
        int base = 0;
        for (int g = 0; g < d->n_groups; g++) {
            if (d->group_kind[g] != RTX_GEOM_MESH) continue;
            const int begin = d->group_begin[g], count = d->group_count[g];
            groupTriBase[g] = base;
            if (count == 0) continue;
            Node4* dn = (Node4*)W.take(((size_t)count + 1) * sizeof(Node4));
            if (!dn || !buildScratch) return fail(ctx, RTX_ERR_CUDA, "device BVH build: work slab exhausted");
            rtxgpu::BlasResult br;
            const char* what = "";
            size_t sb = buildBytes;
            cudaError_t be = rtxgpu::build_blas(dV0 + 3 * (size_t)begin, dV1 + 3 * (size_t)begin, dV2 + 3 * (size_t)begin, dMat + begin, dRank + begin, count, ctx->blas_leaf,
                                                (int)blasNodes, base, dn, dMeshTris + RTX_TRI_D * (size_t)base, dTriInfo + base, ctx->stream, &br, &what, buildScratch, &sb);
            if (be != cudaSuccess) return fail(ctx, RTX_ERR_CUDA, "device BVH build of mesh group %d failed: %s (%s)", g, what, cudaGetErrorString(be));
            groupRoot[g] = (int)blasNodes;   // local node 0 of this mesh
            devBlas.push_back({dn, br.n_nodes});
            blasNodes += (uint32_t)br.n_nodes;
            maxBlasDepth = std::max(maxBlasDepth, br.depth);
            for (int a = 0; a < 3; a++) { groupBox[g].lo[a] = br.lo[a]; groupBox[g].hi[a] = br.hi[a]; }
            base += count;
        }
        CU(cudaEventRecord(evB1, ctx->stream));
        nodes.resize(blasNodes);   // placeholders: the host-built TLAS below gets global node indices
    } else {
        for (int g = 0; g < d->n_groups; g++) {
            if (d->group_kind[g] != RTX_GEOM_MESH) continue;
            int begin = d->group_begin[g], count = d->group_count[g];
            std::vector<Box> boxes(count);
            for (int k = 0; k < count; k++) { boxes[k] = prim_box(d, RTX_GEOM_TRIANGLE, begin + k); groupBox[g].grow(boxes[k]); }
            std::vector<int> ranks;
            if (d->tri_rank) ranks.assign(d->tri_rank + begin, d->tri_rank + begin + count);
            else ranks = rtxbvh::canonical_ranks(boxes);
            int base = (int)triInfo.size();
            groupTriBase[g] = base;
            std::vector<int> perm;
            size_t before = nodes.size();
            int depth = 0;
            groupRoot[g] = rtxbvh::build_bvh4(boxes, ctx->blas_leaf, nodes, perm, [&](int first, int cnt) { return ((base + first) << 3) | (cnt - 1); }, &depth);
            maxBlasDepth = std::max(maxBlasDepth, depth);
            blasNodes += (uint32_t)(nodes.size() - before);
            for (int k = 0; k < count; k++) pushTri(begin + perm[k], perm[k], ranks[perm[k]]);
        }
    }
    for (int ti : looseList) pushTri(ti, 0, 0);
    auto devPrim = [&](int kind, int idx) { return kind == RTX_GEOM_TRIANGLE ? looseOfDesc[idx] : idx; };
    std::vector<int2> listItems(d->n_list_items);
    for (int i = 0; i < d->n_list_items; i++) listItems[i] = make_int2(d->list_item_kind[i], devPrim(d->list_item_kind[i], d->list_item_index[i]));

    // ---- entries, their world bounds, ranks
    std::vector<DEntry> entries(d->n_entries);
    std::vector<Box> entryBox(d->n_entries);
    for (int e = 0; e < d->n_entries; e++) {
        DEntry& E = entries[e];
        int k = d->entry_geom_kind[e], gi = d->entry_geom_index[e];
        E.kind = k; E.index = 0; E.a = 0; E.b = 0;
        E.xf_begin = d->entry_xf_begin[e]; E.xf_count = d->entry_xf_count[e]; E.volume = d->entry_volume[e]; E.rank = e;
        Box b;
        b.reset();
        if (isPrim(k)) {
            E.index = devPrim(k, gi);
            b = prim_box(d, k, gi);
        } else if (k == RTX_GEOM_LIST) {
            E.a = d->group_begin[gi]; E.b = d->group_count[gi];
            for (int i = 0; i < E.b; i++) b.grow(prim_box(d, d->list_item_kind[E.a + i], d->list_item_index[E.a + i]));
            if (E.b == 0) continue;
        } else {
            E.a = groupRoot[gi]; E.b = groupTriBase[gi];
            b = groupBox[gi];   // union of the padded float64 triangle boxes (host loop or device reduction: same values)
        }
        entryBox[e] = xform_box(d, b, E.xf_begin, E.xf_count);
    }
    // canonical [Translate][RotateY][Scale] chains (outermost first) get a one-load record, see xform_ray
    std::vector<double> xfCanon((size_t)12 * std::max(d->n_entries, 1), 0.0);
    for (int e = 0; e < d->n_entries; e++) {
        double* q = xfCanon.data() + 12 * (size_t)e;
        q[0] = q[1] = q[2] = 0; q[3] = 0; q[4] = 1; q[5] = q[6] = q[7] = 1; q[8] = q[9] = q[10] = 1; q[11] = 0;
        int n = entries[e].xf_count, stage = 0;
        bool canon = n > 0 && n <= 3;
        for (int k = 0; k < n && canon; k++) {
            int x = entries[e].xf_begin + k, t = d->xf_type[x];
            int want = t == RTX_XF_TRANSLATE ? 0 : t == RTX_XF_ROTATE_Y ? 1 : 2;
            if (want < stage) { canon = false; break; }
            stage = want + 1;
            if (t == RTX_XF_TRANSLATE) { q[0] = d->xf_a[3 * x]; q[1] = d->xf_a[3 * x + 1]; q[2] = d->xf_a[3 * x + 2]; }
            else if (t == RTX_XF_ROTATE_Y) { q[3] = d->xf_a[3 * x]; q[4] = d->xf_a[3 * x + 1]; }
            else if (d->xf_b) {
                q[5] = d->xf_b[3 * x]; q[6] = d->xf_b[3 * x + 1]; q[7] = d->xf_b[3 * x + 2];
                q[8] = d->xf_a[3 * x]; q[9] = d->xf_a[3 * x + 1]; q[10] = d->xf_a[3 * x + 2]; q[11] = 1;   // Factor, and "a Scale is present"
            }
            else canon = false;
        }
        if (n > 0xffff) return fail(ctx, RTX_ERR_UNSUPPORTED, "entry %d: transform chain too long", e);
        if (canon) entries[e].xf_count |= RTX_XF_CANON;
    }
    {  // test-order ranks: the caller's Go tree, or the canonical stable median-split order
        std::vector<int> ranks;
        if (d->entry_rank) ranks.assign(d->entry_rank, d->entry_rank + d->n_entries);
        else if (d->world_is_bvh) {
            std::vector<Box> rb(d->n_entries);
            for (int e = 0; e < d->n_entries; e++) rb[e] = entryBox[e];
            ranks = rtxbvh::canonical_ranks(rb);
        } else {
            ranks.resize(d->n_entries);
            for (int e = 0; e < d->n_entries; e++) ranks[e] = e;
        }
        for (int e = 0; e < d->n_entries; e++) entries[e].rank = ranks[e];
    }
    // the flat kernels' lists: bare primitives (inlined tests) and everything else (generic entry test); empty lists are skipped
    for (int e = 0; e < d->n_entries; e++) {
        const DEntry& E = entries[e];
        if (E.kind == RTX_GEOM_MESH || (E.kind == RTX_GEOM_LIST && E.b == 0)) continue;
        if (E.kind != RTX_GEOM_LIST && (E.xf_count & 0xffff) == 0 && E.volume < 0) flatSimple.push_back(make_int4(E.kind, E.index, e, E.rank));
        else flatComplex.push_back(e);
    }
    // ---- TLAS over bounded entries; unbounded ones (planes) are tested for every ray
    std::vector<int> unbounded, boundedIdx;
    std::vector<Box> tb;
    // Bare bounded primitives beside a mesh (the walls of CornellBoxLucy): a TLAS leaf costs such a ray an ENTRY round of the persistent
    // kernel that only a few lanes share, while the same float64 test at refill runs for every lane of the warp and hands the traversal a
    // finite interval from its first node on. Closest hits and any-hit answers do not depend on where an entry is tested.
    auto bareBounded = [&](int e) {
        const DEntry& E = entries[e];
        return E.kind != RTX_GEOM_LIST && E.kind != RTX_GEOM_MESH && E.xf_count == 0 && E.volume < 0 && entryBox[e].finite();
    };
    int nBare = 0;
    bool hasMesh = false;
    for (int e = 0; e < d->n_entries; e++) { nBare += bareBounded(e) ? 1 : 0; hasMesh |= entries[e].kind == RTX_GEOM_MESH && entries[e].a >= 0; }
    const bool pretest = hasMesh && nBare > 0 && nBare <= ctx->pretest_bare;
    for (int e = 0; e < d->n_entries; e++) {
        bool empty = (entries[e].kind == RTX_GEOM_LIST && entries[e].b == 0) || (entries[e].kind == RTX_GEOM_MESH && entries[e].a < 0);
        if (empty) continue;
        if (pretest && bareBounded(e)) { unbounded.push_back(e); continue; }
        if (!entryBox[e].finite()) {
            if (entries[e].kind == RTX_GEOM_MESH || entries[e].volume >= 0) return fail(ctx, RTX_ERR_UNSUPPORTED, "entry %d: unbounded mesh/volume", e);
            if (entries[e].kind == RTX_GEOM_LIST) return fail(ctx, RTX_ERR_UNSUPPORTED, "entry %d: a HittableList containing an infinite Plane is outside the device path", e);
            unbounded.push_back(e);
        } else {
            boundedIdx.push_back(e);
            tb.push_back(entryBox[e]);
        }
    }
    std::vector<int> perm;
    size_t before = nodes.size();
    int tlasRoot = rtxbvh::build_bvh4(tb, 1, nodes, perm, [&](int first, int) { return first; }, &tlasDepth);
    // worst-case traversal stack: up to 3 deferred children per level on both levels + the instance marker
    const bool flatTop = hasMesh && ctx->tlas_flat_max > 0 && (int)boundedIdx.size() <= std::min(ctx->tlas_flat_max, 16) && d->n_entries <= 256;
    if (3 * (tlasDepth + maxBlasDepth) + 2 > RTX_STACK_SIZE || (flatTop && (int)boundedIdx.size() + 3 * maxBlasDepth + 2 > RTX_STACK_SIZE))
        return fail(ctx, RTX_ERR_UNSUPPORTED, "BVH too deep for the device traversal stack (TLAS depth %d, BLAS depth %d, stack %d)", tlasDepth,
                    maxBlasDepth, RTX_STACK_SIZE);
    // leaf codes reference positions in `perm`; rewrite them to entry indices
    for (size_t ni = before; ni < nodes.size(); ni++)
        for (int c = 0; c < 4; c++)
            if (nodes[ni].child[c] < 0 && nodes[ni].lox[c] <= nodes[ni].hix[c]) nodes[ni].child[c] = ~boundedIdx[perm[~nodes[ni].child[c]]];
    ctx->tlas_nodes = (uint32_t)(nodes.size() - before);
    ctx->blas_nodes = blasNodes;
    ctx->n_entries = d->n_entries;
    ctx->n_tris = (uint32_t)totalTris;

    // a mesh world with a handful of bounded entries: the top level also goes up as a flat list of entry boxes (DevScene::tlas_boxes)
    std::vector<float4> tlasBoxes;
    if (flatTop)
        for (size_t k = 0; k < boundedIdx.size(); k++) {
            const Box& b = tb[k];
            tlasBoxes.push_back(make_float4(rtxbvh::round_down(b.lo[0]), rtxbvh::round_down(b.lo[1]), rtxbvh::round_down(b.lo[2]), __int_as_float_host(boundedIdx[k])));
            tlasBoxes.push_back(make_float4(rtxbvh::round_up(b.hi[0]), rtxbvh::round_up(b.hi[1]), rtxbvh::round_up(b.hi[2]), 0.f));
        }
    std::vector<DXform> xfs(d->n_xforms);
    for (int x = 0; x < d->n_xforms; x++) {
        xfs[x].type = d->xf_type[x]; xfs[x].pad = 0;
        for (int a = 0; a < 3; a++) { xfs[x].a[a] = d->xf_a[3 * x + a]; xfs[x].b[a] = d->xf_b ? d->xf_b[3 * x + a] : 0; }
    }
    std::vector<DVolume> vols(d->n_volumes);
    for (int i = 0; i < d->n_volumes; i++) { vols[i].neg_inv_density = d->vol_neg_inv_density[i]; vols[i].mat = d->vol_mat[i]; vols[i].pad = 0; }
    std::vector<int> lights(d->light_quad, d->light_quad + d->n_lights);

    // ---- HDRI: BuildDistribution (rt/hdri.go:145-224) in float64, same loop order
    std::vector<float4> envTex;
    std::vector<double> marg, cond, pdf;
    double totalPower = 0;
    if (d->env_width > 0 && d->env_height > 0 && d->env_rgb) {
        int W = d->env_width, H = d->env_height;
        envTex.resize((size_t)W * H);
        pdf.assign((size_t)W * H, 0.0);
        marg.assign(H + 1, 0.0);
        cond.assign((size_t)H * (W + 1), 0.0);
        std::vector<double> rowSums(H, 0.0);
        for (int y = 0; y < H; y++) {
            double v = ((double)y + 0.5) / (double)H;
            double theta = (0.5 - v) * M_PI;
            double sinTheta = std::cos(theta);
            double* cr = cond.data() + (size_t)y * (W + 1);
            for (int x = 0; x < W; x++) {
                size_t idx = (size_t)y * W + x;
                const double* c = d->env_rgb + 3 * idx;
                envTex[idx] = make_float4((float)c[0], (float)c[1], (float)c[2], 0.f);
                double lum = 0.2126 * c[0] + 0.7152 * c[1] + 0.0722 * c[2];
                double w = lum * sinTheta;
                if (w < 0) w = 0;
                pdf[idx] = w;
                rowSums[y] += w;
                totalPower += w;
                cr[x + 1] = cr[x] + w;
            }
        }
        for (int y = 0; y < H; y++)
            if (rowSums[y] > 0) {
                double* cr = cond.data() + (size_t)y * (W + 1);
                for (int x = 0; x <= W; x++) cr[x] /= rowSums[y];
            }
        for (int y = 0; y < H; y++) marg[y + 1] = marg[y] + rowSums[y];
        if (totalPower > 0) {
            for (int y = 0; y <= H; y++) marg[y] /= totalPower;
            for (auto& p : pdf) p /= totalPower;
        }
        S.env_w = W; S.env_h = H; S.env_is = d->env_importance_sampling != 0; S.env_rot = d->env_rotation; S.env_total = totalPower;
    }
    ctx->env_total = totalPower;

    // ---- upload
    {   // Everything the kernels read lives in ONE grow-only slab. The geometry every ray fetches (nodes, triangles, spheres,
        // quads) comes first and contiguous, so that a single L2 access-policy window can cover it (see rtx_render_pass).
        auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
        struct Item { const void* src; size_t bytes; const void** field; };
        std::vector<Item> items;
        auto want = [&](const void* src, size_t bytes, const void** field) { items.push_back({src, bytes, field}); };
#define WANT(vec, field) want((vec).data(), (vec).size() * sizeof((vec)[0]), (const void**)&(field))
        const size_t bNodes = pad(nodes.size() * sizeof(Node4)), bTris = pad((size_t)totalTris * RTX_TRI_D * sizeof(double)), bSph = pad(sph.size() * sizeof(double)),
                     bQuads = pad(quads.size() * sizeof(double)), bInfo = pad((size_t)totalTris * sizeof(int4));
        const size_t bTris32 = (hasMesh && ctx->tri_pretest) ? pad((size_t)totalTris * 3 * sizeof(float4)) : 0;   // float32 pre-test records (mesh worlds)
        const size_t geom = std::max<size_t>(bNodes + bTris + bSph + bQuads + bTris32, 256);
        WANT(entries, S.entries); WANT(unbounded, S.unbounded); WANT(sphMat, S.sph_mat); WANT(quadMat, S.quad_mat); WANT(planes, S.planes); WANT(planeMat, S.plane_mat);
        WANT(circles, S.circles); WANT(circleMat, S.circle_mat); WANT(perlinVec, S.perlin_vec); WANT(perlinPerm, S.perlin_perm);
        WANT(imgRgb, S.img_rgb); WANT(imgDim, S.img_dim); WANT(flatSimple, S.flat_simple); WANT(flatComplex, S.flat_complex);
        WANT(listItems, S.list_items); WANT(xfs, S.xforms); WANT(xfCanon, S.xf_canon); WANT(vols, S.volumes); WANT(mats, S.mats); WANT(texs, S.texs);
        WANT(lights, S.light_quads); WANT(envTex, S.env_tex); WANT(marg, S.env_marg); WANT(cond, S.env_cond); WANT(pdf, S.env_pdf);
#undef WANT
        size_t total = geom + bInfo;
        for (const Item& it : items) total += pad(std::max<size_t>(it.bytes, 16));
        CU(ctx->scene_slab.reserve(total));
        Slab& L = ctx->scene_slab;
        char* base = L.take(geom);
        ctx->geom_arena = base; ctx->geom_bytes = geom; ctx->window_set = false;
        auto put = [&](const void* src, size_t bytes, char* dst) {
            return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
        };
        int4* infoDev = (int4*)L.take(bInfo);
        if (deviceBuild) {   // device-built BLAS nodes and mesh triangles move device-to-device; the host adds the TLAS and the loose triangles
            size_t off = 0;
            for (const DevBlas& db : devBlas) {
                CU(cudaMemcpyAsync(base + off, db.nodes, (size_t)db.n * sizeof(Node4), cudaMemcpyDeviceToDevice, ctx->stream));
                off += (size_t)db.n * sizeof(Node4);
            }
            CU(put(nodes.data() + blasNodes, (nodes.size() - blasNodes) * sizeof(Node4), base + off));
            CU(cudaMemcpyAsync(base + bNodes, dMeshTris, (size_t)meshTotal * RTX_TRI_D * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            CU(put(tris.data(), tris.size() * sizeof(double), base + bNodes + (size_t)meshTotal * RTX_TRI_D * sizeof(double)));
            CU(cudaMemcpyAsync(infoDev, dTriInfo, (size_t)meshTotal * sizeof(int4), cudaMemcpyDeviceToDevice, ctx->stream));
            CU(put(triInfo.data(), triInfo.size() * sizeof(int4), (char*)(infoDev + meshTotal)));
        } else {
            CU(put(nodes.data(), nodes.size() * sizeof(Node4), base));
            CU(put(tris.data(), tris.size() * sizeof(double), base + bNodes));
            CU(put(triInfo.data(), triInfo.size() * sizeof(int4), (char*)infoDev));
        }
        CU(put(sph.data(), sph.size() * sizeof(double), base + bNodes + bTris));
        CU(put(quads.data(), quads.size() * sizeof(double), base + bNodes + bTris + bSph));
        S.nodes = (const float4*)base;
        S.tris = (const double*)(base + bNodes);
        S.spheres = (const double*)(base + bNodes + bTris);
        S.quads = (const double*)(base + bNodes + bTris + bSph);
        S.tri_info = infoDev;
        S.tris32 = nullptr;
        if (bTris32 && totalTris > 0) {   // derived on the device from the float64 records just placed
            float4* t32 = (float4*)(base + bNodes + bTris + bSph + bQuads);
            k_tris32<<<(totalTris + 255) / 256, 256, 0, ctx->stream>>>(S.tris, totalTris, t32);
            S.tris32 = t32;
        }
        for (const Item& it : items) {
            char* dst = L.take(std::max<size_t>(it.bytes, 16));
            if (!dst) return fail(ctx, RTX_ERR_CUDA, "scene slab exhausted");
            CU(put(it.src, it.bytes, dst));
            *it.field = dst;
        }
    }
    S.tlas_root = tlasRoot;
    S.n_entries = d->n_entries;
    S.n_unbounded = (int)unbounded.size();
    S.n_nodes = (int)nodes.size(); S.n_tris_total = totalTris;
    S.n_tlas_flat = (int)(tlasBoxes.size() / 2);
    for (size_t k = 0; k < tlasBoxes.size() && k < 32; k++) S.tlas_boxes[k] = tlasBoxes[k];
    S.n_lights = d->n_lights;
    S.n_images = d->n_images;
    S.n_flat_simple = (int)flatSimple.size(); S.n_flat_complex = (int)flatComplex.size();
    S.vol_draws = d->world_is_bvh ? 2 : 1;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->ms_upload_blas = 0;
    if (evB0) {
        float ms = 0;
        cudaEventElapsedTime(&ms, evB0, evB1);
        ctx->ms_upload_blas = ms;
    }
    if (getenv("RTX_DEBUG_BATCH"))
        fprintf(stderr, "[rtx] scene upload: %d mesh triangles, %u BLAS nodes (depth %d, %s build %.2f ms, test-order ranks %.2f ms), %u TLAS nodes, %.2f ms so far\n", meshTotal, blasNodes, maxBlasDepth,
                deviceBuild ? "device" : "host", ctx->ms_upload_blas, ctx->ms_upload_ranks, ctx->tlas_nodes,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tUpload0).count());
    ctx->S = S;
    ctx->have_scene = true;
    ctx->blas_depth = maxBlasDepth; ctx->built_on_device = deviceBuild ? 1 : 0;
    ctx->scene_has_mesh = 0;
    for (int e = 0; e < d->n_entries; e++) ctx->scene_has_mesh |= d->entry_geom_kind[e] == RTX_GEOM_MESH;
    ctx->scene_flat = !ctx->scene_has_mesh && d->n_entries <= ctx->flat_max_entries;
    {   // the vocabulary a flat-world kernel variant has to contain (RTX_F_*)
        unsigned f = 0;
        for (const int4& fe : flatSimple)
            f |= fe.x == RTX_GEOM_QUAD ? RTX_F_QUAD : fe.x == RTX_GEOM_SPHERE ? RTX_F_SPHERE : fe.x == RTX_GEOM_PLANE ? RTX_F_PLANE : RTX_F_OTHER_PRIM;
        if (!flatComplex.empty() || d->n_volumes > 0) f |= RTX_F_COMPLEX;
        for (int e = 0; e < d->n_entries; e++)
            if (entries[e].kind == RTX_GEOM_MESH) f |= RTX_F_MESH | (entries[e].xf_count ? RTX_F_XFORM : 0u);
        if (pretest)   // pre-tested kinds other than quads go through the generic test of the refill
            for (int e : unbounded) if (entries[e].kind != RTX_GEOM_QUAD && entries[e].kind != RTX_GEOM_PLANE) f |= RTX_F_COMPLEX;
        if (d->env_width > 0 && d->env_height > 0 && d->env_rgb) f |= RTX_F_ENV;
        if (d->n_lights > 0) f |= RTX_F_LIGHTS;
        for (int i = 0; i < d->n_textures; i++)
            if (d->tex_type[i] == RTX_TEX_NOISE || d->tex_type[i] == RTX_TEX_IMAGE) f |= RTX_F_TEX_X;
        for (int i = 0; i < d->n_materials; i++)
            f |= d->mat_type[i] == RTX_MAT_METAL ? RTX_F_METAL : d->mat_type[i] == RTX_MAT_DIELECTRIC ? RTX_F_DIELECTRIC : d->mat_type[i] == RTX_MAT_ISOTROPIC ? RTX_F_ISOTROPIC : 0u;
        ctx->feat_mask = f;
    }
    ctx->ms_upload_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tUpload0).count();
    std::memset(&ctx->stats, 0, sizeof ctx->stats);
    return RTX_OK;
}

// ---- camera (Initialize, rt/camera.go:286-344, float64, same operation order) -----------------------------------------
static int32_t camera_set_single(rtx_ctx* ctx, const rtx_camera_desc* c) {
    if (!ctx || !c) return RTX_ERR_INVALID;
    if (c->image_width <= 0 || !(c->aspect_ratio > 0)) return fail(ctx, RTX_ERR_INVALID, "camera: bad resolution");
    CU(cudaSetDevice(ctx->device));
    DevCamera C{};
    int W = c->image_width;
    int H = c->has_derived ? c->image_height : std::max((int)((double)W / c->aspect_ratio), 1);
    auto sub3 = [](const double* a, const double* b, double* o) { for (int i = 0; i < 3; i++) o[i] = a[i] - b[i]; };
    auto scale3 = [](const double* a, double t, double* o) { for (int i = 0; i < 3; i++) o[i] = t * a[i]; };
    auto unit3 = [](const double* a, double* o) {
        double l = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        if (l == 0) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; return; }
        double inv = 1 / l;
        for (int i = 0; i < 3; i++) o[i] = inv * a[i];
    };
    auto cross3 = [](const double* a, const double* b, double* o) {
        double r[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
    };
    const double Pi = 3.1415926535897932385;  // rt/utils.go:11
    if (c->has_derived) {
        for (int i = 0; i < 3; i++) {
            C.center[i] = c->center[i]; C.pixel00[i] = c->pixel00_loc[i]; C.du[i] = c->pixel_delta_u[i]; C.dv[i] = c->pixel_delta_v[i];
            C.u[i] = c->u[i]; C.v[i] = c->v[i]; C.w[i] = c->w[i];
        }
        C.defocus_radius = c->defocus_radius; C.viewport_w = c->viewport_width; C.viewport_h = c->viewport_height;
    } else {
        double theta = c->vfov * Pi / 180.0;
        double h = std::tan(theta / 2);
        double vh = 2 * h * c->focus_dist;
        double vw = vh * ((double)W / (double)H);
        double w[3], u[3], v[3], t[3];
        if (c->free_camera) { for (int i = 0; i < 3; i++) w[i] = -c->forward[i]; }
        else { sub3(c->look_from, c->look_at, t); unit3(t, w); }
        cross3(c->vup, w, t); unit3(t, u);
        cross3(w, u, v);
        double vpU[3], vpV[3], nv[3] = {-v[0], -v[1], -v[2]};
        scale3(u, vw, vpU); scale3(nv, vh, vpV);
        scale3(vpU, 1 / (double)W, C.du);
        scale3(vpV, 1 / (double)H, C.dv);
        double wf[3], hU[3], hV[3], ul[3], sum[3];
        scale3(w, c->focus_dist, wf); scale3(vpU, 1 / 2.0, hU); scale3(vpV, 1 / 2.0, hV);
        for (int i = 0; i < 3; i++) ul[i] = ((c->look_from[i] - wf[i]) - hU[i]) - hV[i];
        for (int i = 0; i < 3; i++) sum[i] = C.du[i] + C.dv[i];
        for (int i = 0; i < 3; i++) C.pixel00[i] = ul[i] + 0.5 * sum[i];
        for (int i = 0; i < 3; i++) { C.center[i] = c->look_from[i]; C.u[i] = u[i]; C.v[i] = v[i]; C.w[i] = w[i]; }
        C.defocus_radius = c->focus_dist * std::tan((c->defocus_angle / 2) * Pi / 180.0);
        C.viewport_w = vw; C.viewport_h = vh;
    }
    C.defocus_angle = c->defocus_angle; C.focus_dist = c->focus_dist;
    for (int i = 0; i < 3; i++) {
        C.look_from[i] = c->look_from[i]; C.look_at[i] = c->look_at[i]; C.vup[i] = c->vup[i]; C.forward[i] = c->forward[i];
        C.look_vel[i] = c->camera_motion ? c->look_from2[i] - c->look_from[i] : 0.0;
        C.look_at_vel[i] = c->camera_motion ? c->look_at2[i] - c->look_at[i] : 0.0;
        C.background[i] = (float)c->background[i];
    }
    C.camera_motion = c->camera_motion; C.free_camera = c->free_camera;
    C.width = W; C.height = H; C.use_sky = c->use_sky_gradient; C.phantom = c->phantom_hdri; C.max_depth = c->max_depth;
    bool resize = (W != ctx->W || H != ctx->H) || !ctx->accum;
    ctx->C = C; ctx->W = W; ctx->H = H; ctx->have_camera = true;
    if (resize) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->accum) cudaFree(ctx->accum);
        if (ctx->accum_sq) cudaFree(ctx->accum_sq);
        if (ctx->rgba_dev) cudaFree(ctx->rgba_dev);
        ctx->accum = ctx->accum_sq = nullptr; ctx->rgba_dev = nullptr;
        CU(cudaMalloc((void**)&ctx->rgba_dev, (size_t)W * H * sizeof(uchar4)));
        size_t bytes = (size_t)W * H * sizeof(float4);
        CU(cudaMalloc((void**)&ctx->accum, bytes));
        CU(cudaMalloc((void**)&ctx->accum_sq, bytes));
        CU(cudaMemsetAsync(ctx->accum, 0, bytes, ctx->stream));
        CU(cudaMemsetAsync(ctx->accum_sq, 0, bytes, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return RTX_OK;
}

int32_t rtx_image_size(const rtx_ctx* ctx, int32_t* w, int32_t* h) {
    if (!ctx || !ctx->have_camera) return RTX_ERR_STATE;
    if (w) *w = ctx->W;
    if (h) *h = ctx->H;
    return RTX_OK;
}

static int32_t accum_clear_single(rtx_ctx* ctx) {
    if (!ctx || !ctx->have_camera) return ctx ? fail(ctx, RTX_ERR_STATE, "camera not set") : RTX_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    size_t bytes = (size_t)ctx->W * ctx->H * sizeof(float4);
    CU(cudaMemsetAsync(ctx->accum, 0, bytes, ctx->stream));
    CU(cudaMemsetAsync(ctx->accum_sq, 0, bytes, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return RTX_OK;
}
int32_t rtx_accum_device_ptr(rtx_ctx* ctx, void** sum_dev, void** sumsq_dev, int64_t* n_floats) {
    if (!ctx || !ctx->have_camera) return ctx ? fail(ctx, RTX_ERR_STATE, "camera not set") : RTX_ERR_INVALID;
    if (sum_dev) *sum_dev = ctx->accum;
    if (sumsq_dev) *sumsq_dev = ctx->accum_sq;
    if (n_floats) *n_floats = (int64_t)4 * ctx->W * ctx->H;
    return RTX_OK;
}

static int32_t ensure_pool(rtx_ctx* ctx, unsigned long long paths_of_pass) {
    // capacity: the configured limit, or less when the whole pass is smaller (a 400x225 preview does not need gigabytes);
    // grow-only below the limit, so the passes of one render (1 spp, spp/4, spp) allocate at most once per size step
    unsigned long long want = std::min<unsigned long long>((unsigned long long)ctx->pool_paths, std::max<unsigned long long>(paths_of_pass, 1ull << 16));
    want = (want + 0xffffull) & ~0xffffull;
    want = std::min<unsigned long long>(want, (unsigned long long)ctx->pool_paths);
    // shadow requests exist only with registered lights (next-event estimation, rt/camera.go:487-517): at most one per Lambertian hit towards
    // the chosen area light and one more towards the environment when it is importance-sampled
    const int shadowPerHit = ctx->S.n_lights > 0 ? 1 + ((ctx->S.env_w > 0 && ctx->S.env_is) ? 1 : 0) : 0;
    const int hitBytes = ctx->S.n_images > 0 ? RTX_HIT_BYTES_UV : RTX_HIT_BYTES;
    if (ctx->pool.capacity > 0 && (unsigned long long)ctx->pool.capacity >= want && ctx->pool.capacity <= ctx->pool_paths &&
        ctx->pool_has_shadow >= shadowPerHit && ctx->pool_hit_bytes >= hitBytes)
        return RTX_OK;
    free_pool(ctx);
    size_t P = (size_t)want;
    Pool p{};
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** out, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(out, std::max<size_t>(bytes, 256));
        if (e == cudaSuccess) ctx->pool_allocs.push_back(*out);
    };
    alloc((void**)&p.rec[0], P * RTX_REC_BYTES);
    alloc((void**)&p.rec[1], P * RTX_REC_BYTES);
    alloc((void**)&p.hit, P * (size_t)hitBytes);
    alloc((void**)&p.q_mat, (size_t)Q_COUNT * P * sizeof(int));
    alloc((void**)&p.shadow, 2 * (size_t)shadowPerHit * P * RTX_SHADOW_BYTES);   // two halves by iteration parity
    if (e != cudaSuccess) {   // a partial allocation must not stay behind: several contexts may share the device
        free_pool(ctx);
        cudaGetLastError();
        return fail(ctx, e == cudaErrorMemoryAllocation ? RTX_ERR_NOMEM : RTX_ERR_CUDA, "path pool of %zu paths: %s", P, cudaGetErrorString(e));
    }
    p.capacity = (int)P;
    p.hit_bytes = hitBytes;
    ctx->pool = p;
    ctx->pool_has_shadow = shadowPerHit; ctx->pool_hit_bytes = hitBytes;
    return RTX_OK;
}

// ---- the hot path ------------------------------------------------------------------------------------------------------
// Error inside the wavefront loop: k_connect launches of earlier iterations may still be running on the connect stream (adding into
// the accumulation buffer, using trace_spill2); the context stays usable, so both streams are drained before the call returns.
#define CUL(call)                                                                                             \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            cudaStreamSynchronize(st); cudaStreamSynchronize(ctx->connect_stream);                            \
            return fail(ctx, RTX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                   \
        }                                                                                                     \
    } while (0)

static int32_t render_pass_single(rtx_ctx* ctx, int32_t spp, int32_t max_depth, int32_t camera_max_depth, uint64_t seed, uint32_t sample_base) {
    if (!ctx) return RTX_ERR_INVALID;
    if (!ctx->have_scene || !ctx->have_camera) return fail(ctx, RTX_ERR_STATE, "rtx_render_pass: scene and camera must be set first");
    if (spp < 0 || max_depth < 0) return fail(ctx, RTX_ERR_INVALID, "rtx_render_pass: negative spp/depth");
    if (ctx->moments && spp > 0 && (unsigned long long)ctx->W * ctx->H * (unsigned long long)spp > (1ull << 28))   // before any work is queued
        return fail(ctx, RTX_ERR_UNSUPPORTED, "moments are limited to 2^28 samples per pass (requested %llu)", (unsigned long long)ctx->W * ctx->H * (unsigned long long)spp);
    CU(cudaSetDevice(ctx->device));
    int32_t rc = ensure_pool(ctx, (unsigned long long)ctx->W * ctx->H * (unsigned long long)std::max(spp, 0));
    if (rc != RTX_OK) return rc;
    cudaStream_t st = ctx->stream;
    if (ctx->l2_persist && (!ctx->window_set || ctx->window_stream != st) && ctx->geom_bytes > 0) {
        // Keep the scene geometry resident in L2: a persisting access-policy window over the geometry arena on the render
        // stream. Without it the path pool (hundreds of MB per bounce) evicts nodes and triangles (ncu: lts hit rate 68 %).
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            size_t want = std::min<size_t>(ctx->geom_bytes, (size_t)prop.persistingL2CacheMaxSize);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = ctx->geom_arena;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(ctx->geom_bytes, (size_t)prop.accessPolicyMaxWindowSize);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
        }
        cudaGetLastError();  // the window is an optimisation: never fail the pass over it
        ctx->window_set = true; ctx->window_stream = st;
    }
    const int P = ctx->pool.capacity;
    PassParams pp{};
    pp.spp = spp; pp.max_depth = max_depth; pp.camera_max_depth = camera_max_depth;
    pp.seed_lo = (uint32_t)seed; pp.seed_hi = (uint32_t)(seed >> 32); pp.sample_base = sample_base;
    pp.moments = ctx->moments; pp.count_stats = ctx->count_stats; pp.pixel_major = ctx->pixel_major;
    pp.shade_direct = (ctx->shade_direct && !ctx->shade_split && !ctx->scene_flat && !ctx->fuse_tree) ? 1 : 0;

    const unsigned long long npix = (unsigned long long)ctx->W * ctx->H;
    Ctl init{};
    init.total = npix * (unsigned long long)spp;
    if (max_depth == 0) init.total = 0;  // RayColor(depth 0) is black: nothing to trace (the samples still count in the divisor)
    CU(cudaMemcpyAsync(ctx->ctl, &init, sizeof(Ctl), cudaMemcpyHostToDevice, st));
    Pool pool = ctx->pool;
    pool.moments = ctx->moments; pool.npix = (uint32_t)npix; pool.sample_base = sample_base;
    pool.hit_bytes = ctx->S.n_images > 0 ? RTX_HIT_BYTES_UV : RTX_HIT_BYTES;   // the stride THIS scene's kernels use (the allocation may be roomier)
    pool.target = reinterpret_cast<float*>(ctx->accum);
    if (ctx->moments && spp > 0) {   // per-sample sums (squared and folded by k_pass_finish): tests only, bounded
        const unsigned long long need = npix * (unsigned long long)spp;
        if (need > ctx->per_sample_cap) {
            CU(cudaStreamSynchronize(st));
            if (ctx->per_sample) cudaFree(ctx->per_sample);
            ctx->per_sample = nullptr; ctx->per_sample_cap = 0;
            CU(cudaMalloc((void**)&ctx->per_sample, need * sizeof(float4)));
            ctx->per_sample_cap = need;
        }
        CU(cudaMemsetAsync(ctx->per_sample, 0, need * sizeof(float4), st));
        pool.target = reinterpret_cast<float*>(ctx->per_sample);
    }

    const int BATCH = 16;
    enum { EV_GEN = 0, EV_EXT, EV_SHADE, EV_CONN, EV_KINDS };
    const bool timing = ctx->time_kernels != 0;
    size_t needEvents = 4 + (timing ? (size_t)BATCH * EV_KINDS * 2 : 0);
    while (ctx->events.size() < needEvents) {
        cudaEvent_t ev;
        CU(cudaEventCreate(&ev));
        ctx->events.push_back(ev);
    }
    cudaEvent_t evStart = ctx->events[0], evStop = ctx->events[1];
    double msKind[EV_KINDS] = {0, 0, 0, 0};
    uint64_t launches = 0;
    CU(cudaEventRecord(evStart, st));
    // fixed grids: the stream kernels stride over device-side counts, the trace kernels are persistent (one resident wave of
    // warps pulls rays from a device-side cursor); the host only polls the control block every BATCH iterations
    const int gridStream = std::min((P + 255) / 256, ctx->num_sms * 8), gridTrace = ctx->trace_grid;
    // The shadow rays of iteration i only feed the accumulation buffer, so k_connect(i) runs on a second stream beside
    // k_generate / k_extend / k_shade of iteration i + 1: its blocks move in as the persistent k_extend blocks of the next
    // iteration drain (the tail of a persistent launch otherwise leaves SMs idle), and the render stream has the higher
    // priority, so the critical chain extend -> shade -> extend is served first. Shadow requests, their count and the job
    // cursor are double-buffered by iteration parity; iteration i + 2 waits for k_connect(i).
    const int lean = lean_variant(ctx);                              // kernel variant by scene vocabulary (0 = all features)
    const int leanShade = ctx->S.n_images > 0 ? 0 : lean;            // image textures travel with the UV kernels, which exist in full only
    const bool overlap = ctx->overlap_connect && ctx->S.n_lights > 0;
    cudaStream_t sc = overlap ? ctx->connect_stream : st;
    int* const spillC = overlap ? ctx->trace_spill2 : ctx->trace_spill;
    bool pending[2] = {false, false};
    long long iter = 0;
    const long long debugIter = getenv("RTX_DEBUG_ITER") ? atoll(getenv("RTX_DEBUG_ITER")) : -1;
    // The drain: once the pass's last camera path has been generated the stream only shrinks, and the host knows an upper bound of the
    // rays in flight (the survivor count of the last polled iteration). Grids are then sized for that bound instead of for the pool:
    // a launch of 1036 persistent blocks (or 1184 stream blocks) for a few thousand rays is mostly block scheduling.
    int activeBound = P;
    long long iterGenDone = -1;
    const bool fusedTree = !ctx->scene_flat && ctx->fuse_tree;
    // one-kernel iterations with nothing else on the render stream: the two ends of a batch give the kernel's time
    const bool sparseEvents = ((ctx->scene_flat && ctx->fuse_flat) || fusedTree) && !(ctx->S.n_lights > 0 && !overlap);
    const int drainPerBlock = getenv("RTX_DRAIN_PER_BLOCK") ? std::max(1, atoi(getenv("RTX_DRAIN_PER_BLOCK"))) : RTX_DRAIN_PER_BLOCK_DEFAULT;
    for (;;) {
        int used = 0;
        auto shrink = [&](int fullGrid, int perBlock) { return std::max(1, std::min(fullGrid, (int)(((long long)activeBound + perBlock - 1) / perBlock))); };
        const bool smallBatch = !ctx->scene_flat && !ctx->fuse_tree && ctx->count_stats == 0 && ctx->S.n_images == 0 && activeBound <= ctx->simple_below;
        const int gStreamB = shrink(gridStream, 256), gTraceB = shrink(gridTrace, drainPerBlock), gLucyB = shrink(ctx->trace_grid_lucy, drainPerBlock), gSkyB = shrink(ctx->trace_grid_sky, drainPerBlock);
        // iterations per host poll: 16 while camera paths are still generated; 4 once the stream only shrinks, so that the drain starts promptly
        const int batchN = (ctx->fuse_drain > 0 && iterGenDone >= 0 && !ctx->scene_flat) ? 4 : BATCH;
        for (int b = 0; b < batchN; b++, iter++) {
            const int cur = (int)(iter & 1);   // rec[cur]: this iteration's paths; rec[cur ^ 1]: where k_shade writes the survivors
            cudaEvent_t* ev = timing ? &ctx->events[4 + (size_t)b * EV_KINDS * 2] : nullptr;
            if (pending[cur]) { CUL(cudaStreamWaitEvent(st, ctx->ev_connected[cur], 0)); pending[cur] = false; }
            Pool poolI = pool;
            poolI.shadow = pool.shadow + (size_t)cur * (size_t)ctx->pool_has_shadow * (size_t)P * RTX_SHADOW_BYTES;
            // timing events: the boundaries between the kernels of an iteration only (an event record between two launches costs the stream
            // a few microseconds: eight per iteration were 4 % of hdri-test); an iteration starts where the previous one ended, so the
            // one-warp k_iter_begin is counted with the kernel that follows it; the one-kernel paths record the two ends of a batch only. Slots: 0 batch start, 1 after generate, 2 after extend /
            // bounce, 3 after shade, 4 / 5 around connect.
            if (timing && b == 0) cudaEventRecord(ev[0], st);
            k_iter_begin<<<1, 32, 0, st>>>(ctx->ctl, P, cur);
            const bool fused = ctx->scene_flat && ctx->fuse_flat;   // k_bounce_flat generates the fresh paths itself
            if (!fused) k_generate<<<gStreamB, 256, 0, st>>>(ctx->ctl, pool, cur, ctx->C, pp);
            if (timing && !fused && !fusedTree) cudaEventRecord(ev[1], st);
            if (fused) {   // generate + trace + shade in one kernel: fresh paths and hits stay in registers
                if (ctx->S.n_images > 0) k_bounce_flat<false, true><<<gStreamB, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                else if (ctx->count_stats & 1) k_bounce_flat<true><<<gStreamB, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                else {
                    if (lean == 2) k_bounce_flat<false, false, RTX_FV_SKY><<<gStreamB, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                    else if (lean == 3 || lean == 1) k_bounce_flat<false, false, RTX_FV_BOX><<<gStreamB, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                    else if (lean == 4) k_bounce_flat<false, false, RTX_FV_CORNELL><<<gStreamB, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                    else k_bounce_flat<false><<<gStreamB, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                }
                if (timing && (!sparseEvents || b == batchN - 1)) cudaEventRecord(ev[2], st);   // one kernel per iteration: only the sum over the batch is wanted
                launches -= 2;
            } else if (fusedTree) {   // trace + shade in the persistent kernel: the hit never leaves the lane that found it
                if (ctx->S.n_images > 0) k_bounce<false, true><<<gTraceB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp, ctx->trace_spill);
                else if (ctx->count_stats & 1) k_bounce<true><<<gTraceB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp, ctx->trace_spill);
                else k_bounce<false><<<gTraceB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp, ctx->trace_spill);
                if (timing && (!sparseEvents || b == batchN - 1)) cudaEventRecord(ev[2], st);
                launches -= 1;
            } else {
            if (ctx->S.n_images > 0) {   // hit records carry (u, v); this variant is not instrumented
                if (ctx->scene_flat) k_extend_flat<false, true><<<gStreamB, 256, 0, st>>>(ctx->ctl, pool, cur, ctx->S, pp);
                else k_extend<false, true><<<gTraceB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, st>>>(ctx->ctl, pool, cur, ctx->S, pp, ctx->trace_spill);
            } else if (ctx->scene_flat) {
                if (ctx->count_stats & 1) k_extend_flat<true><<<gStreamB, 256, 0, st>>>(ctx->ctl, pool, cur, ctx->S, pp);
                else k_extend_flat<false><<<gStreamB, 256, 0, st>>>(ctx->ctl, pool, cur, ctx->S, pp);
            } else if (smallBatch) {   // the drain: one thread per ray (trace_simple)
                if (lean == 1) k_extend_simple<false, RTX_FV_LUCY><<<gStreamB, 256, 0, st>>>(ctx->ctl, pool, cur, ctx->S, pp);
                else if (lean == 2) k_extend_simple<false, RTX_FV_SKY><<<gStreamB, 256, 0, st>>>(ctx->ctl, pool, cur, ctx->S, pp);
                else k_extend_simple<false><<<gStreamB, 256, 0, st>>>(ctx->ctl, pool, cur, ctx->S, pp);
            } else if (ctx->count_stats & 1) k_extend<true><<<gTraceB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, st>>>(ctx->ctl, pool, cur, ctx->S, pp, ctx->trace_spill);
            else if (lean == 1) k_extend<false, false, RTX_FV_LUCY><<<gLucyB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES_OF(RTX_FV_LUCY), st>>>(ctx->ctl, pool, cur, ctx->S, pp, ctx->trace_spill);
            else if (lean == 2) k_extend<false, false, RTX_FV_SKY><<<gSkyB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES_OF(RTX_FV_SKY), st>>>(ctx->ctl, pool, cur, ctx->S, pp, ctx->trace_spill);
            else k_extend<false><<<gTraceB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, st>>>(ctx->ctl, pool, cur, ctx->S, pp, ctx->trace_spill);
            if (timing) cudaEventRecord(ev[2], st);
            if (ctx->shade_split) {   // one launch per material queue (launches over empty queues return at once)
                const int gs = shrink(std::min((P + 255) / 256, ctx->num_sms * 2 * RTX_SHADE_BLOCKS_Q), 256);
                k_shade<Q_MISS><<<gs, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                k_shade<Q_LAMBERTIAN><<<gs, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                if (ctx->mat_kinds & (1 << Q_METAL)) k_shade<Q_METAL><<<gs, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                if (ctx->mat_kinds & (1 << Q_DIELECTRIC)) k_shade<Q_DIELECTRIC><<<gs, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                if (ctx->mat_kinds & (1 << Q_LIGHT)) k_shade<Q_LIGHT><<<gs, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                if (ctx->mat_kinds & (1 << Q_ISOTROPIC)) k_shade<Q_ISOTROPIC><<<gs, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                launches += 1 + __builtin_popcount(ctx->mat_kinds & ((1 << Q_METAL) | (1 << Q_DIELECTRIC) | (1 << Q_LIGHT) | (1 << Q_ISOTROPIC)));
            } else
            {
                const int gsh = shrink(std::min((P + 255) / 256, ctx->num_sms * 2 * (leanShade ? RTX_SHADE_BLOCKS_LEAN : RTX_SHADE_BLOCKS)), 256);
                if (leanShade == 1) k_shade<-1, RTX_FV_LUCY><<<gsh, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                else if (leanShade == 2) k_shade<-1, RTX_FV_SKY><<<gsh, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                else if (leanShade == 3) k_shade<-1, RTX_FV_BOX><<<gsh, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                else if (leanShade == 4) k_shade<-1, RTX_FV_CORNELL><<<gsh, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
                else k_shade<-1><<<gsh, 256, 0, st>>>(ctx->ctl, poolI, cur, ctx->S, ctx->C, pp);
            }
            if (timing) cudaEventRecord(ev[3], st);
            }
            if (ctx->S.n_lights > 0) {
                if (overlap) { CUL(cudaEventRecord(ctx->ev_shaded, st)); CUL(cudaStreamWaitEvent(sc, ctx->ev_shaded, 0)); }
                if (timing) cudaEventRecord(ev[4], sc);
                if (ctx->scene_flat) {
                    if (ctx->count_stats & 2) k_connect_flat<true><<<gStreamB, 256, 0, sc>>>(ctx->ctl, poolI, cur, ctx->S, pp);
                    else if (lean == 3 || lean == 1) k_connect_flat<false, RTX_FV_BOX><<<gStreamB, 256, 0, sc>>>(ctx->ctl, poolI, cur, ctx->S, pp);
                    else if (lean == 4) k_connect_flat<false, RTX_FV_CORNELL><<<gStreamB, 256, 0, sc>>>(ctx->ctl, poolI, cur, ctx->S, pp);
                    else k_connect_flat<false><<<gStreamB, 256, 0, sc>>>(ctx->ctl, poolI, cur, ctx->S, pp);
                } else if (smallBatch) {
                    if (lean == 1) k_connect_simple<RTX_FV_LUCY><<<gStreamB, 256, 0, sc>>>(ctx->ctl, poolI, cur, ctx->S, pp);
                    else if (lean == 2) k_connect_simple<RTX_FV_SKY><<<gStreamB, 256, 0, sc>>>(ctx->ctl, poolI, cur, ctx->S, pp);
                    else k_connect_simple<><<<gStreamB, 256, 0, sc>>>(ctx->ctl, poolI, cur, ctx->S, pp);
                } else if (ctx->count_stats & 2) k_connect<true><<<gTraceB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, sc>>>(ctx->ctl, poolI, cur, ctx->S, pp, spillC);
                else if (lean == 1) k_connect<false, RTX_FV_LUCY><<<gLucyB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES_OF(RTX_FV_LUCY), sc>>>(ctx->ctl, poolI, cur, ctx->S, pp, spillC);
                else if (lean == 2) k_connect<false, RTX_FV_SKY><<<gSkyB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES_OF(RTX_FV_SKY), sc>>>(ctx->ctl, poolI, cur, ctx->S, pp, spillC);
                else k_connect<false><<<gTraceB, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, sc>>>(ctx->ctl, poolI, cur, ctx->S, pp, spillC);
                if (timing) cudaEventRecord(ev[5], sc);
                if (overlap) { CUL(cudaEventRecord(ctx->ev_connected[cur], sc)); pending[cur] = true; }
                launches++;
            }
            launches += 4;
            used++;
            if (debugIter >= 0 && iter == debugIter) {   // developer aid (RTX_DEBUG_ITER=k): the job counts of iteration k, e.g. of the launch an ncu capture picked
                Ctl snap;
                cudaStreamSynchronize(st); cudaStreamSynchronize(sc);
                cudaMemcpy(&snap, ctx->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[rtx] iteration %lld: extension rays %d, shadow rays %d\n", iter, snap.n_active, ctl_shadow(&snap, cur));
            }
        }
        // join: the control block read below must include the connect kernels of this batch (statistics, timing events)
        for (int c = 0; c < 2; c++)
            if (pending[c]) { CUL(cudaStreamWaitEvent(st, ctx->ev_connected[c], 0)); pending[c] = false; }
        CUL(cudaMemcpyAsync(ctx->ctl_host, ctx->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        CUL(cudaStreamSynchronize(st));
        if (timing) {
            const bool oneKernel = (ctx->scene_flat && ctx->fuse_flat) || fusedTree;   // generate / trace / shade in one launch: boundary 2 only
            auto slot = [&](int b, int k) { return ctx->events[4 + (size_t)b * EV_KINDS * 2 + k]; };
            auto span = [&](cudaEvent_t a, cudaEvent_t z) { float ms = 0; cudaEventElapsedTime(&ms, a, z); return (double)ms; };
            if (sparseEvents && used > 0) msKind[EV_EXT] += span(slot(0, 0), slot(used - 1, 2));
            for (int b = 0; b < used; b++) {
                // where the previous iteration ended on this stream: after its shade kernel, or after its connect kernel when that ran here too
                const cudaEvent_t start = b == 0 ? slot(0, 0) : slot(b - 1, (ctx->S.n_lights > 0 && !overlap) ? 5 : 3);
                if (sparseEvents) {}
                else if (oneKernel) msKind[EV_EXT] += span(start, slot(b, 2));
                else { msKind[EV_GEN] += span(start, slot(b, 1)); msKind[EV_EXT] += span(slot(b, 1), slot(b, 2)); msKind[EV_SHADE] += span(slot(b, 2), slot(b, 3)); }
                if (ctx->S.n_lights > 0) msKind[EV_CONN] += span(slot(b, 4), slot(b, 5));
            }
        }
        const int lastPar = (int)((iter - 1) & 1);                       // parity of the last iteration issued
        const int nNext = ctl_survivors(ctx->ctl_host, lastPar);         // its survivors: the rays of the next iteration
        if (getenv("RTX_DEBUG_BATCH"))
            fprintf(stderr, "[rtx] iter %lld: ms gen/ext/shade/conn %.2f/%.2f/%.2f/%.2f active %d next %d shadow %d cursor %llu\n", iter, msKind[0], msKind[1],
                    msKind[2], msKind[3], ctx->ctl_host->n_active, nNext, ctl_shadow(ctx->ctl_host, lastPar), ctx->ctl_host->cursor);
        if (ctx->ctl_host->done) break;
        activeBound = ctx->ctl_host->cursor >= ctx->ctl_host->total ? std::max(nNext, 1) : P;
        if (ctx->ctl_host->cursor >= ctx->ctl_host->total && iterGenDone < 0) iterGenDone = iter;   // every live path was generated before this iteration
        // ---- the barrier-free drain: ONE persistent launch for all remaining bounces (k_drain, rtx_kernels.cuh) ------------------------------
        // Its shadow requests all go to one half of the shadow buffer: every live path has at least (iter - iterGenDone) bounces behind it, so
        // there are at most n_next x (max_depth - (iter - iterGenDone)) x requests-per-hit of them — the drain starts when that fits.
        if (ctx->fuse_drain > 0 && iterGenDone >= 0 && !ctx->scene_flat && !ctx->fuse_tree && ctx->count_stats == 0 && nNext > 0) {
            const long long remaining = std::max<long long>(1, (long long)max_depth - (iter - iterGenDone));
            // (both halves of the shadow buffer, 2 x pool_has_shadow x P requests, hold them: the connect launches of earlier iterations are joined first)
            const long long perHit = 1 + ((ctx->S.env_w > 0 && ctx->S.env_is) ? 1 : 0);
            if (nNext <= ctx->fuse_drain &&
                (ctx->S.n_lights == 0 || (long long)nNext * remaining * perHit <= 2ll * ctx->pool_has_shadow * (long long)P)) {
                const int cur = 0;   // the drain's shadow requests start at the beginning of the buffer; cur_connect[0] / n_shadow[0] count them
                for (int c = 0; c < 2; c++)
                    if (pending[c]) { CUL(cudaStreamWaitEvent(st, ctx->ev_connected[c], 0)); pending[c] = false; }
                Pool poolI = pool;
                const int recCur = (int)(iter & 1);   // rec[recCur] holds the survivors of the last iteration
                cudaEvent_t* ev = timing ? &ctx->events[4] : nullptr;   // the slots of the batch just read: the drain's rays count as extension rays, its k_connect as connect
                if (timing) cudaEventRecord(ev[0], st);
                k_drain_begin<<<1, 32, 0, st>>>(ctx->ctl, cur, lastPar);
                if (ctx->S.n_images > 0) k_drain<true><<<gridTrace, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, st>>>(ctx->ctl, poolI, recCur, cur, ctx->S, ctx->C, pp, ctx->trace_spill);
                else if (lean == 1) k_drain<false, RTX_FV_LUCY><<<ctx->drain_grid_lucy, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES_OF(RTX_FV_LUCY), st>>>(ctx->ctl, poolI, recCur, cur, ctx->S, ctx->C, pp, ctx->trace_spill);
                else if (lean == 2) k_drain<false, RTX_FV_SKY><<<ctx->drain_grid_sky, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES_OF(RTX_FV_SKY), st>>>(ctx->ctl, poolI, recCur, cur, ctx->S, ctx->C, pp, ctx->trace_spill);
                else k_drain<false><<<gridTrace, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, st>>>(ctx->ctl, poolI, recCur, cur, ctx->S, ctx->C, pp, ctx->trace_spill);
                if (timing) cudaEventRecord(ev[2], st);
                if (ctx->S.n_lights > 0) {   // the shadow requests of the whole drain, one launch
                    if (lean == 1) k_connect<false, RTX_FV_LUCY><<<ctx->trace_grid_lucy, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES_OF(RTX_FV_LUCY), st>>>(ctx->ctl, poolI, cur, ctx->S, pp, ctx->trace_spill);
                    else if (lean == 2) k_connect<false, RTX_FV_SKY><<<ctx->trace_grid_sky, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES_OF(RTX_FV_SKY), st>>>(ctx->ctl, poolI, cur, ctx->S, pp, ctx->trace_spill);
                    else k_connect<false><<<gridTrace, RTX_TRACE_THREADS, RTX_TRACE_SMEM_BYTES, st>>>(ctx->ctl, poolI, cur, ctx->S, pp, ctx->trace_spill);
                    launches++;
                }
                if (timing) cudaEventRecord(ev[5], st);
                k_drain_end<<<1, 32, 0, st>>>(ctx->ctl);
                launches += 3;
                for (int c = 0; c < 2; c++)
                    if (pending[c]) { CUL(cudaStreamWaitEvent(st, ctx->ev_connected[c], 0)); pending[c] = false; }
                CUL(cudaMemcpyAsync(ctx->ctl_host, ctx->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
                CUL(cudaStreamSynchronize(st));
                if (timing) {
                    float ms = 0;
                    cudaEventElapsedTime(&ms, ev[0], ev[2]); msKind[EV_EXT] += ms;
                    cudaEventElapsedTime(&ms, ev[2], ev[5]); msKind[EV_CONN] += ms;
                }
                break;
            }
        }
    }
    if (spp > 0) {
        k_pass_finish<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(ctx->accum, ctx->accum_sq, ctx->per_sample, (int)npix, spp, ctx->moments);
        launches++;
    }
    CU(cudaEventRecord(evStop, st));
    CU(cudaEventSynchronize(evStop));
    CU(cudaGetLastError());
    float msTotal = 0;
    cudaEventElapsedTime(&msTotal, evStart, evStop);
    const Ctl& c = *ctx->ctl_host;
    rtx_stats& s = ctx->stats;
    s.paths = (uint64_t)ctx->W * ctx->H * (uint64_t)spp;
    s.extension_rays = c.ext_rays; s.shadow_rays = c.shadow_rays; s.nodes_visited = c.nodes;
    s.tri_tests = c.tris; s.sphere_tests = c.spheres; s.quad_tests = c.quads; s.plane_tests = c.planes;
    s.wavefront_iterations = c.iterations; s.kernel_launches = launches;
    s.ms_generate = msKind[EV_GEN]; s.ms_extend = msKind[EV_EXT]; s.ms_shade = msKind[EV_SHADE]; s.ms_connect = msKind[EV_CONN];
    s.ms_total = msTotal;
    s.ms_tail = (c.t_tail_begin && c.t_end > c.t_tail_begin) ? (double)(c.t_end - c.t_tail_begin) * 1e-6 : 0.0;   // %globaltimer, ns
    s.tail_iterations = c.tail_iterations;
    s.ms_reduce = 0; s.n_devices = 1;
    s.tlas_nodes = ctx->tlas_nodes; s.blas_nodes = ctx->blas_nodes; s.n_entries = ctx->n_entries; s.n_tris = ctx->n_tris;
    return RTX_OK;
}


// ---- calls that act on every device of an rtx_create_multi context ---------------------------------------------------------------
// (a single-device context has no peers: the loops are empty and the call is the single-device one)
}  // extern "C"
template <class F>
static int32_t on_all_devices(rtx_ctx* ctx, F&& f, bool parallel) {
    if (!ctx) return RTX_ERR_INVALID;
    if (ctx->peers.empty()) return f(ctx);
    std::vector<rtx_ctx*> all{ctx};
    all.insert(all.end(), ctx->peers.begin(), ctx->peers.end());
    std::vector<int32_t> rc(all.size(), RTX_OK);
    if (parallel) {
        std::vector<std::thread> th;
        for (size_t g = 1; g < all.size(); g++) th.emplace_back([&, g] { rc[g] = f(all[g]); });
        rc[0] = f(all[0]);
        for (auto& t : th) t.join();
    } else {
        for (size_t g = 0; g < all.size(); g++) rc[g] = f(all[g]);
    }
    for (size_t g = 0; g < all.size(); g++)
        if (rc[g] != RTX_OK) {
            if (g > 0) ctx->err = "device " + std::to_string(all[g]->device) + ": " + all[g]->err;
            return rc[g];
        }
    return RTX_OK;
}
extern "C" {
int32_t rtx_scene_upload(rtx_ctx* ctx, const rtx_scene_desc* d) {
    return on_all_devices(ctx, [d](rtx_ctx* c) { return scene_upload_single(c, d); }, true);   // the replicas upload and build side by side
}
int32_t rtx_camera_set(rtx_ctx* ctx, const rtx_camera_desc* c) {
    return on_all_devices(ctx, [c](rtx_ctx* x) { return camera_set_single(x, c); }, false);
}
int32_t rtx_accum_clear(rtx_ctx* ctx) {
    return on_all_devices(ctx, [](rtx_ctx* x) { return accum_clear_single(x); }, false);
}
int32_t rtx_set_option(rtx_ctx* ctx, const char* key, int64_t value) {
    return on_all_devices(ctx, [key, value](rtx_ctx* x) { return set_option_single(x, key, value); }, false);
}
int32_t rtx_accum_enable_moments(rtx_ctx* ctx, int32_t enable) {
    return on_all_devices(ctx, [enable](rtx_ctx* x) { x->moments = enable != 0; return (int32_t)RTX_OK; }, false);
}

// Replaces the fan-out of renderPass to its worker goroutines (rt/bucket_renderer.go:193-213): device g renders the sample slice
// [sample_base + g * spp / n, sample_base + (g + 1) * spp / n) of every pixel on a host thread of its own (the wavefront loop polls its
// device), then ONE ncclReduce per buffer sums the accumulation buffers onto device 0 over NVLink. The Philox counters are keyed by the
// global sample index, so the union of the slices is the sample set one device would have rendered.
int32_t rtx_render_pass(rtx_ctx* ctx, int32_t spp, int32_t max_depth, int32_t camera_max_depth, uint64_t seed, uint32_t sample_base) {
    if (!ctx) return RTX_ERR_INVALID;
    if (ctx->peers.empty()) return render_pass_single(ctx, spp, max_depth, camera_max_depth, seed, sample_base);
    if (spp < 0) return fail(ctx, RTX_ERR_INVALID, "rtx_render_pass: negative spp");
    std::vector<rtx_ctx*> all{ctx};
    all.insert(all.end(), ctx->peers.begin(), ctx->peers.end());
    const int n = (int)all.size();
    std::vector<int32_t> rc(n, RTX_OK);
    auto slice = [&](int g) {
        const long long lo = (long long)spp * g / n, hi = (long long)spp * (g + 1) / n;
        rc[g] = render_pass_single(all[g], (int32_t)(hi - lo), max_depth, camera_max_depth, seed, sample_base + (uint32_t)lo);
    };
    {
        std::vector<std::thread> th;
        for (int g = 1; g < n; g++) th.emplace_back(slice, g);
        slice(0);
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < n; g++)
        if (rc[g] != RTX_OK) {
            if (g > 0) ctx->err = "device " + std::to_string(all[g]->device) + ": " + all[g]->err;
            return rc[g];
        }
    // ---- the one exchange of the path: sum-reduce of the accumulation buffers to device 0 (13 MB at 1200x675, 133 MB at 4K)
    CU(cudaSetDevice(ctx->device));
    while (ctx->events.size() < 4) { cudaEvent_t ev; CU(cudaEventCreate(&ev)); ctx->events.push_back(ev); }
    const size_t count = (size_t)4 * ctx->W * ctx->H;
    CU(cudaEventRecord(ctx->events[2], ctx->stream));
    ncclResult_t nr = g_nccl.GroupStart();
    for (int g = 0; g < n && nr == ncclSuccess; g++) {
        nr = g_nccl.Reduce(all[g]->accum, all[g]->accum, count, ncclFloat, ncclSum, 0, ctx->comms[g], all[g]->stream);
        if (nr == ncclSuccess && ctx->moments) nr = g_nccl.Reduce(all[g]->accum_sq, all[g]->accum_sq, count, ncclFloat, ncclSum, 0, ctx->comms[g], all[g]->stream);
    }
    ncclResult_t ne = g_nccl.GroupEnd();
    if (nr == ncclSuccess) nr = ne;
    if (nr != ncclSuccess) return fail(ctx, RTX_ERR_CUDA, "rtx_render_pass: ncclReduce failed: %s", g_nccl.GetErrorString(nr));
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->events[3], ctx->stream));
    // device 0 now holds the running total; the other devices start the next pass from zero, so that a later reduce adds only what is new
    for (int g = 1; g < n; g++) {
        CU(cudaSetDevice(all[g]->device));
        CU(cudaMemsetAsync(all[g]->accum, 0, count * sizeof(float), all[g]->stream));
        if (ctx->moments) CU(cudaMemsetAsync(all[g]->accum_sq, 0, count * sizeof(float), all[g]->stream));
        CU(cudaStreamSynchronize(all[g]->stream));
    }
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    float msReduce = 0;
    cudaEventElapsedTime(&msReduce, ctx->events[2], ctx->events[3]);
    // whole-job statistics: counters add up, times are the slowest device's
    rtx_stats& s = ctx->stats;
    for (int g = 1; g < n; g++) {
        const rtx_stats& p = all[g]->stats;
        s.paths += p.paths; s.extension_rays += p.extension_rays; s.shadow_rays += p.shadow_rays; s.nodes_visited += p.nodes_visited;
        s.tri_tests += p.tri_tests; s.sphere_tests += p.sphere_tests; s.quad_tests += p.quad_tests; s.plane_tests += p.plane_tests;
        s.kernel_launches += p.kernel_launches;
        s.wavefront_iterations = std::max(s.wavefront_iterations, p.wavefront_iterations); s.tail_iterations = std::max(s.tail_iterations, p.tail_iterations);
        s.ms_generate = std::max(s.ms_generate, p.ms_generate); s.ms_extend = std::max(s.ms_extend, p.ms_extend); s.ms_shade = std::max(s.ms_shade, p.ms_shade);
        s.ms_connect = std::max(s.ms_connect, p.ms_connect); s.ms_total = std::max(s.ms_total, p.ms_total); s.ms_tail = std::max(s.ms_tail, p.ms_tail);
    }
    s.ms_reduce = msReduce; s.n_devices = (uint32_t)n;
    s.kernel_launches += (uint64_t)(n - 1) * (ctx->moments ? 2 : 1);   // the memsets are not kernels of this library; NCCL's kernels are not either: not counted
    ctx->ms_reduce = msReduce;
    return RTX_OK;
}

int32_t rtx_get_stats(rtx_ctx* ctx, rtx_stats* out) {
    if (!ctx || !out) return RTX_ERR_INVALID;
    *out = ctx->stats;
    out->tlas_nodes = ctx->tlas_nodes; out->blas_nodes = ctx->blas_nodes; out->n_entries = ctx->n_entries; out->n_tris = ctx->n_tris;
    out->blas_depth = (uint32_t)ctx->blas_depth; out->bvh_on_device = (uint32_t)ctx->built_on_device;
    out->ms_bvh_build = ctx->ms_upload_blas; out->ms_scene_upload = ctx->ms_upload_total;
    for (rtx_ctx* p : ctx->peers) { out->ms_bvh_build = std::max(out->ms_bvh_build, p->ms_upload_blas); out->ms_scene_upload = std::max(out->ms_scene_upload, p->ms_upload_total); }
    out->ms_resolve = ctx->ms_resolve;
    out->n_devices = (uint32_t)(1 + ctx->peers.size());
    out->checked_build = 0; out->checked_violations = 0;
#ifdef RTX_CHECKED
    {   // the run-time assertions of the checked build (rtx_trace.cuh), summed over all kinds and all launches since the library was loaded
        unsigned long long v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(v, g_rtx_check, sizeof v);
        out->checked_build = 1;
        for (int k = 0; k < 8; k++) { out->checked_violations += v[k]; out->checked_by_kind[k] = v[k]; }
    }
#else
    for (int k = 0; k < 8; k++) out->checked_by_kind[k] = 0;
#endif
    return RTX_OK;
}

int32_t rtx_resolve_rgba8(rtx_ctx* ctx, int32_t total_spp, uint8_t* pix, int64_t nbytes) {
    if (!ctx || !pix) return RTX_ERR_INVALID;
    if (!ctx->have_camera) return fail(ctx, RTX_ERR_STATE, "camera not set");
    int npix = ctx->W * ctx->H;
    if (nbytes != (int64_t)4 * npix) return fail(ctx, RTX_ERR_INVALID, "rtx_resolve_rgba8: nbytes %lld != 4*W*H = %d", (long long)nbytes, 4 * npix);
    if (total_spp <= 0) return fail(ctx, RTX_ERR_INVALID, "rtx_resolve_rgba8: total_spp must be positive");
    CU(cudaSetDevice(ctx->device));
    uchar4* dev = ctx->rgba_dev;   // sized by rtx_camera_set: no allocation on the per-pass path
    while (ctx->events.size() < 2) { cudaEvent_t ev; CU(cudaEventCreate(&ev)); ctx->events.push_back(ev); }
    cudaEventRecord(ctx->events[0], ctx->stream);
    k_resolve_rgba8<<<(npix + 255) / 256, 256, 0, ctx->stream>>>(ctx->accum, npix, 1.0 / (double)total_spp, dev);
    cudaError_t e = cudaMemcpyAsync(pix, dev, (size_t)npix * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaEventRecord(ctx->events[1], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(ctx, RTX_ERR_CUDA, "rtx_resolve_rgba8: %s", cudaGetErrorString(e));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->events[0], ctx->events[1]);
    ctx->ms_resolve = ms;
    return RTX_OK;
}

int32_t rtx_resolve_accum(rtx_ctx* ctx, float* sum_rgb, float* sumsq_rgb, uint32_t* n) {
    if (!ctx) return RTX_ERR_INVALID;
    if (!ctx->have_camera) return fail(ctx, RTX_ERR_STATE, "camera not set");
    CU(cudaSetDevice(ctx->device));
    size_t npix = (size_t)ctx->W * ctx->H;
    std::vector<float4> h(npix);
    CU(cudaMemcpyAsync(h.data(), ctx->accum, npix * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < npix; i++) {
        if (sum_rgb) { sum_rgb[3 * i] = h[i].x; sum_rgb[3 * i + 1] = h[i].y; sum_rgb[3 * i + 2] = h[i].z; }
        if (n) n[i] = (uint32_t)(h[i].w + 0.5f);
    }
    if (sumsq_rgb) {
        CU(cudaMemcpyAsync(h.data(), ctx->accum_sq, npix * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < npix; i++) { sumsq_rgb[3 * i] = h[i].x; sumsq_rgb[3 * i + 1] = h[i].y; sumsq_rgb[3 * i + 2] = h[i].z; }
    }
    return RTX_OK;
}

// ---- batch entry points (parity tests) -------------------------------------------------------------------------------------
#define BACK(dev, host, n) if (host) CU(cudaMemcpyAsync(host, dev, (n) * sizeof(*host), cudaMemcpyDeviceToHost, ctx->stream))

int32_t rtx_trace_closest(rtx_ctx* ctx, const double* rays, int64_t n, double tmin, double tmax, int32_t* entry_id, int32_t* prim_id, double* t,
                          double* normal, uint8_t* front, double* uv, double* p) {
    if (!ctx || (!rays && n > 0) || n < 0) return RTX_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, RTX_ERR_STATE, "rtx_trace_closest: no scene uploaded");
    if (n == 0) return RTX_OK;
    CU(cudaSetDevice(ctx->device));
    Scratch sc;
    double* dRays; int *dEntry, *dPrim; double *dT, *dN, *dUV, *dP; unsigned char* dFront;
    CU(sc.in(&dRays, rays, (size_t)7 * n, ctx->stream));
    CU(sc.out(&dEntry, entry_id, n)); CU(sc.out(&dPrim, prim_id, n)); CU(sc.out(&dT, t, n)); CU(sc.out(&dN, normal, 3 * n));
    CU(sc.out(&dFront, (unsigned char*)front, n)); CU(sc.out(&dUV, uv, 2 * n)); CU(sc.out(&dP, p, 3 * n));
    if (n > (int64_t)1 << 30) return fail(ctx, RTX_ERR_INVALID, "rtx_trace_closest: at most 2^30 rays per call");
    CU(cudaMemsetAsync(ctx->batch_cursor, 0, sizeof(int), ctx->stream));
    const int lean = lean_variant(ctx);   // the variant a rendered pass of this scene runs: the bit-exact tests cover the lean code
    const int gf = std::min((int)((n + 255) / 256), ctx->num_sms * 8);
#define RTX_TC_FLAT(F) k_trace_closest_flat<F><<<gf, 256, 0, ctx->stream>>>(ctx->S, dRays, (int)n, tmin, tmax, dEntry, dPrim, dT, dN, dFront, dUV, dP)
#define RTX_TC_TREE(F) k_trace_closest<F><<<(F) == RTX_F_ALL ? ctx->trace_grid : (F) == RTX_FV_SKY ? ctx->trace_grid_sky : ctx->trace_grid_lucy, RTX_TRACE_THREADS, (F) == RTX_F_ALL ? RTX_TRACE_SMEM_BYTES : RTX_TRACE_SMEM_BYTES_OF(F), ctx->stream>>>(ctx->S, dRays, (int)n, tmin, tmax, ctx->batch_cursor, ctx->trace_spill, dEntry, dPrim, dT, dN, dFront, dUV, dP)
    if (ctx->scene_flat) {
        if (lean == 2) RTX_TC_FLAT(RTX_FV_SKY); else if (lean == 3 || lean == 1) RTX_TC_FLAT(RTX_FV_BOX); else if (lean == 4) RTX_TC_FLAT(RTX_FV_CORNELL); else RTX_TC_FLAT(RTX_F_ALL);
    } else {
        if (lean == 1) RTX_TC_TREE(RTX_FV_LUCY); else if (lean == 2) RTX_TC_TREE(RTX_FV_SKY); else RTX_TC_TREE(RTX_F_ALL);
    }
#undef RTX_TC_FLAT
#undef RTX_TC_TREE
    CU(cudaGetLastError());
    BACK(dEntry, entry_id, n); BACK(dPrim, prim_id, n); BACK(dT, t, n); BACK(dN, normal, 3 * n); BACK(dFront, front, n); BACK(dUV, uv, 2 * n); BACK(dP, p, 3 * n);
    CU(cudaStreamSynchronize(ctx->stream));
    return RTX_OK;
}

int32_t rtx_camera_rays(rtx_ctx* ctx, const int32_t* ij, const double* sq, const double* disk, const double* tm, int64_t n, double* rays_out) {
    if (!ctx || !ij || !sq || !disk || !tm || !rays_out || n < 0) return RTX_ERR_INVALID;
    if (!ctx->have_camera) return fail(ctx, RTX_ERR_STATE, "camera not set");
    if (n == 0) return RTX_OK;
    CU(cudaSetDevice(ctx->device));
    Scratch sc;
    int* dIj; double *dSq, *dDisk, *dTm, *dOut;
    CU(sc.in(&dIj, ij, (size_t)2 * n, ctx->stream)); CU(sc.in(&dSq, sq, (size_t)2 * n, ctx->stream));
    CU(sc.in(&dDisk, disk, (size_t)2 * n, ctx->stream)); CU(sc.in(&dTm, tm, (size_t)n, ctx->stream));
    CU(sc.out(&dOut, rays_out, (size_t)7 * n));
    k_camera_rays<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->C, dIj, dSq, dDisk, dTm, n, dOut);
    CU(cudaGetLastError());
    BACK(dOut, rays_out, (size_t)7 * n);
    CU(cudaStreamSynchronize(ctx->stream));
    return RTX_OK;
}

static int32_t need_env(rtx_ctx* ctx) {
    if (!ctx->have_scene) return fail(ctx, RTX_ERR_STATE, "no scene uploaded");
    if (ctx->S.env_w <= 0) return fail(ctx, RTX_ERR_STATE, "scene has no HDRI environment");
    return RTX_OK;
}
int32_t rtx_hdri_sample(rtx_ctx* ctx, const double* xi, int64_t n, double* dir, double* emission, double* pdf) {
    if (!ctx || !xi || !dir || !emission || !pdf || n < 0) return RTX_ERR_INVALID;
    int32_t rc = need_env(ctx);
    if (rc != RTX_OK) return rc;
    if (ctx->env_total == 0) return fail(ctx, RTX_ERR_STATE, "HDRI has zero total power");
    if (n == 0) return RTX_OK;
    CU(cudaSetDevice(ctx->device));
    Scratch sc;
    double *dXi, *dDir, *dEm, *dPdf;
    CU(sc.in(&dXi, xi, (size_t)2 * n, ctx->stream));
    CU(sc.out(&dDir, dir, (size_t)3 * n)); CU(sc.out(&dEm, emission, (size_t)3 * n)); CU(sc.out(&dPdf, pdf, (size_t)n));
    k_hdri_sample<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->S, dXi, n, dDir, dEm, dPdf);
    CU(cudaGetLastError());
    BACK(dDir, dir, (size_t)3 * n); BACK(dEm, emission, (size_t)3 * n); BACK(dPdf, pdf, (size_t)n);
    CU(cudaStreamSynchronize(ctx->stream));
    return RTX_OK;
}
int32_t rtx_hdri_pdf(rtx_ctx* ctx, const double* dir, int64_t n, double* pdf) {
    if (!ctx || !dir || !pdf || n < 0) return RTX_ERR_INVALID;
    int32_t rc = need_env(ctx);
    if (rc != RTX_OK) return rc;
    if (n == 0) return RTX_OK;
    CU(cudaSetDevice(ctx->device));
    Scratch sc;
    double *dDir, *dPdf;
    CU(sc.in(&dDir, dir, (size_t)3 * n, ctx->stream));
    CU(sc.out(&dPdf, pdf, (size_t)n));
    k_hdri_pdf<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->S, dDir, n, dPdf);
    CU(cudaGetLastError());
    BACK(dPdf, pdf, (size_t)n);
    CU(cudaStreamSynchronize(ctx->stream));
    return RTX_OK;
}
int32_t rtx_hdri_lookup(rtx_ctx* ctx, const double* dir, int64_t n, double* rgb) {
    if (!ctx || !dir || !rgb || n < 0) return RTX_ERR_INVALID;
    int32_t rc = need_env(ctx);
    if (rc != RTX_OK) return rc;
    if (n == 0) return RTX_OK;
    CU(cudaSetDevice(ctx->device));
    Scratch sc;
    double *dDir, *dRgb;
    CU(sc.in(&dDir, dir, (size_t)3 * n, ctx->stream));
    CU(sc.out(&dRgb, rgb, (size_t)3 * n));
    k_hdri_lookup<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->S, dDir, n, dRgb);
    CU(cudaGetLastError());
    BACK(dRgb, rgb, (size_t)3 * n);
    CU(cudaStreamSynchronize(ctx->stream));
    return RTX_OK;
}
int32_t rtx_hdri_total_power(const rtx_ctx* ctx, double* total_power) {
    if (!ctx || !total_power) return RTX_ERR_INVALID;
    if (!ctx->have_scene || ctx->S.env_w <= 0) return RTX_ERR_STATE;
    *total_power = ctx->env_total;
    return RTX_OK;
}

}  // extern "C"
