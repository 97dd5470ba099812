// rtx_rank_gpu.cuh — the reference's BVH TEST ORDER of a triangle mesh, computed on the device (sm_100a).
//
// Why it exists: exact ties in t between two triangles of one mesh are resolved the way the reference's traversal would resolve
// them — the triangle its BVH tests later wins (closed interval, rt/triangle.go:89) — so every mesh triangle carries its rank in
// the DFS leaf order of NewBVHNode (rt/bvh.go:69-217: median split of the centroid-sorted range along the longest axis of the
// centroid bounds, leaves of <= 4). Until round 2 the host mirror built that pointer tree on every load just to read the ranks off
// it (103 ms for 280 K triangles, rt_obj.cpp); the library's own fallback (rtxbvh::canonical_ranks) re-did it on one host thread.
// The order itself needs no tree:
//
//   * the split positions depend on the COUNT only (mid = n / 2), so the segments of every level are known up front;
//   * per level: centroid bounds of every segment (segmented min / max), the axis rule of AABB.LongestAxis (rt/aabb.go:139-150)
//     with the 1e-4 padding NewAABBFromPoints gives every centroid box (rt/aabb.go:32-40), one STABLE segmented sort by the
//     centroid coordinate on that axis (Go's sort.Slice is unstable, so the order among equal keys is unknowable; stable is
//     the canonical choice of DESIGN.md section 3, the one the host restatement and the oracle make), and a gather;
//   * a triangle's rank is its final position.
//
// cub::DeviceSegmentedSort does the heavy lifting (library code, off the per-ray hot path). 280 K triangles:
// 17 levels, a few milliseconds.
#pragma once
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <vector>

namespace rtxrank {

__global__ void k_centroids(const double* v0, const double* v1, const double* v2, int n, double* cx, double* cy, double* cz, int* idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double c[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const double p = v0[3 * (size_t)i + a], q = v1[3 * (size_t)i + a], r = v2[3 * (size_t)i + a];
        double lo = fmin(p, fmin(q, r)), hi = fmax(p, fmax(q, r));
        if (hi - lo < 1e-4) { lo -= 1e-4; hi += 1e-4; }   // padToMinimums (rt/aabb.go:117-128), as prim_box in rtx_api.cu
        c[a] = (lo + hi) * 0.5;                            // AABB.Centroid (rt/aabb.go:153-159)
    }
    cx[i] = c[0]; cy[i] = c[1]; cz[i] = c[2]; idx[i] = i;
}

// Axis of every active segment: centroid bounds (min / max over the segment, GROUP threads per segment: a block for the few big segments
// of the top levels, a warp for the many small ones further down), each centroid box padded by 1e-4 per side before the union
// (NewAABBFromPoints -> padToMinimums, rt/aabb.go:32-40), then AABB.LongestAxis (rt/aabb.go:139-150).
template <int GROUP>
__global__ void k_axis(const double* cx, const double* cy, const double* cz, const int* seg_begin, const int* seg_end, int nseg, int* axis) {
    const int s = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) / GROUP), lane = threadIdx.x % GROUP;
    __shared__ double red[6][8];
    double mn[3] = {1e308, 1e308, 1e308}, mx[3] = {-1e308, -1e308, -1e308};
    if (s < nseg)
        for (int i = seg_begin[s] + lane; i < seg_end[s]; i += GROUP) {
            const double c[3] = {cx[i], cy[i], cz[i]};
#pragma unroll
            for (int a = 0; a < 3; a++) { mn[a] = fmin(mn[a], c[a]); mx[a] = fmax(mx[a], c[a]); }
        }
#pragma unroll
    for (int a = 0; a < 3; a++)
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    if (GROUP > 32) {   // one block per segment: combine the warps
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) for (int a = 0; a < 3; a++) { red[a][w] = mn[a]; red[3 + a][w] = mx[a]; }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int a = 0; a < 3; a++)
                for (int k = 1; k < GROUP / 32; k++) { mn[a] = fmin(mn[a], red[a][k]); mx[a] = fmax(mx[a], red[3 + a][k]); }
    }
    if (s < nseg && lane == 0) {
        double sz[3];
#pragma unroll
        for (int a = 0; a < 3; a++) sz[a] = (mx[a] + 0.0001) - (mn[a] - 0.0001);
        axis[s] = (sz[0] > sz[1] && sz[0] > sz[2]) ? 0 : (sz[1] > sz[2] ? 1 : 2);
    }
}
// The few big segments of the top levels: PARTS blocks per segment, combined through 64-bit integer atomics on an order-preserving
// image of the doubles (one block per segment took 0.46 ms for the 280 K-triangle root).
#define RTX_RANK_WIDE_SEGS 16
#define RTX_RANK_WIDE_PARTS 64
__device__ inline unsigned long long ord_bits(double d) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ inline double ord_value(unsigned long long u) { return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u)); }
__global__ void k_axis_init(unsigned long long* mm, int nseg) {
    const int i = threadIdx.x;
    if (i < nseg * 6) mm[i] = (i % 6) < 3 ? ~0ull : 0ull;
}
__global__ void k_axis_part(const double* cx, const double* cy, const double* cz, const int* seg_begin, const int* seg_end, unsigned long long* mm) {
    const int s = blockIdx.x, b = seg_begin[s], m = seg_end[s] - b;
    const int lo = b + (int)((long long)m * blockIdx.y / gridDim.y), hi = b + (int)((long long)m * (blockIdx.y + 1) / gridDim.y);
    __shared__ double red[6][8];
    double mn[3] = {1e308, 1e308, 1e308}, mx[3] = {-1e308, -1e308, -1e308};
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double c[3] = {cx[i], cy[i], cz[i]};
#pragma unroll
        for (int a = 0; a < 3; a++) { mn[a] = fmin(mn[a], c[a]); mx[a] = fmax(mx[a], c[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) for (int a = 0; a < 3; a++) { red[a][w] = mn[a]; red[3 + a][w] = mx[a]; }
    __syncthreads();
    if (threadIdx.x == 0 && hi > lo) {
        for (int a = 0; a < 3; a++)
            for (int k = 1; k < (int)blockDim.x / 32; k++) { mn[a] = fmin(mn[a], red[a][k]); mx[a] = fmax(mx[a], red[3 + a][k]); }
        for (int a = 0; a < 3; a++) { atomicMin(&mm[s * 6 + a], ord_bits(mn[a])); atomicMax(&mm[s * 6 + 3 + a], ord_bits(mx[a])); }
    }
}
__global__ void k_axis_finish(const unsigned long long* mm, int nseg, int* axis) {
    const int s = threadIdx.x;
    if (s >= nseg) return;
    double sz[3];
#pragma unroll
    for (int a = 0; a < 3; a++) sz[a] = (ord_value(mm[s * 6 + 3 + a]) + 0.0001) - (ord_value(mm[s * 6 + a]) - 0.0001);
    axis[s] = (sz[0] > sz[1] && sz[0] > sz[2]) ? 0 : (sz[1] > sz[2] ? 1 : 2);
}
__global__ void k_keys(const double* cx, const double* cy, const double* cz, const int* seg_begin, const int* axis, int nseg, int n_total, const int* seg_of, double* key, int* pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    pos[i] = i;
    const int s = seg_of[i];
    if (s < 0) { key[i] = 0.0; return; }   // a finished leaf range: not part of any segment of this level
    const int a = axis[s];
    key[i] = a == 0 ? cx[i] : a == 1 ? cy[i] : cz[i];
}
__global__ void k_mark_segments(const int* seg_begin, const int* seg_end, int nseg, int* seg_of) {
    const int s = blockIdx.x;
    for (int i = seg_begin[s] + blockIdx.y * blockDim.x + threadIdx.x; i < seg_end[s]; i += blockDim.x * gridDim.y) seg_of[i] = s;
}
__global__ void k_gather(const int* pos, int n, const double* cx, const double* cy, const double* cz, const int* idx, double* ox, double* oy, double* oz, int* oidx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = pos[i];
    ox[i] = cx[p]; oy[i] = cy[p]; oz[i] = cz[p]; oidx[i] = idx[p];
}
__global__ void k_ranks(const int* idx, int n, int* rank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rank[idx[i]] = i;
}

// Content hash of a device array of 64-bit words (order-independent sum of position-mixed words): the key under which rtx_scene_upload
// remembers the test order of a mesh it has ranked before, so that re-uploading an unchanged scene does not sort it again.
__global__ void k_hash64(const unsigned long long* w, size_t n, unsigned long long salt, unsigned long long* out) {
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long z = w[i] ^ ((i + salt) * 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        acc += z ^ (z >> 31);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// Scratch bytes for a mesh of n triangles (an upper bound; the caller carves it from its work slab).
inline size_t scratch_bytes(int n) {
    const size_t pad = 256;
    size_t b = 0;
    b += 2 * 3 * ((size_t)n * sizeof(double) + pad);       // centroids, ping-pong
    b += 2 * ((size_t)n * sizeof(int) + pad);               // triangle index, ping-pong
    b += 2 * ((size_t)n * sizeof(double) + pad);            // keys in / out
    b += 3 * ((size_t)n * sizeof(int) + pad);               // pos in / out, seg_of
    b += 2 * ((size_t)n * sizeof(int) + pad);               // segment begin / end (at most n / 2 segments per level)
    b += (size_t)n * sizeof(int) + pad;                     // axis per segment
    b += RTX_RANK_WIDE_SEGS * 6 * sizeof(unsigned long long) + pad;   // centroid bounds of the wide top-level segments
    b += (size_t)n * 48 + (64u << 20);                      // cub temporary storage (segmented sort keeps copies of keys and values)
    return b;
}

// rank[t] = position of triangle t in the reference's test order. All pointers are device pointers; v0 / v1 / v2 are [3 n] doubles.
// Runs on `st`; returns after the work is queued (the host-side segment lists of every level are copied with cudaMemcpyAsync from
// vectors that live until the final synchronize inside this function).
inline cudaError_t canonical_ranks(const double* v0, const double* v1, const double* v2, int n, int* rank, char* scratch, size_t scratch_size, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    char* sp = scratch;
    auto take = [&](size_t bytes) { char* p = sp; sp += (bytes + 255) & ~(size_t)255; return p; };
    double* c[2][3]; int* idx[2];
    for (int b = 0; b < 2; b++) { for (int a = 0; a < 3; a++) c[b][a] = (double*)take((size_t)n * sizeof(double)); idx[b] = (int*)take((size_t)n * sizeof(int)); }
    double *keyIn = (double*)take((size_t)n * sizeof(double)), *keyOut = (double*)take((size_t)n * sizeof(double));
    int *posIn = (int*)take((size_t)n * sizeof(int)), *posOut = (int*)take((size_t)n * sizeof(int)), *segOf = (int*)take((size_t)n * sizeof(int));
    int *dBegin = (int*)take((size_t)n * sizeof(int)), *dEnd = (int*)take((size_t)n * sizeof(int));
    int* axis = (int*)take((size_t)(n / 2 + 1) * sizeof(int));
    unsigned long long* mm = (unsigned long long*)take(RTX_RANK_WIDE_SEGS * 6 * sizeof(unsigned long long));
    char* cubTemp = sp;
    if (sp > scratch + scratch_size) return cudaErrorMemoryAllocation;
    const size_t cubBytesAvail = (size_t)(scratch + scratch_size - sp);
    const int T = 256, G = (n + T - 1) / T;
    k_centroids<<<G, T, 0, st>>>(v0, v1, v2, n, c[0][0], c[0][1], c[0][2], idx[0]);
    // the segments of every level follow from n alone: [lo, hi) -> [lo, lo + m / 2) and [lo + m / 2, hi) while m > 4
    std::vector<std::vector<int>> keepAlive;
    std::vector<int> begin{0}, end{n};
    int cur = 0;
    cudaError_t e = cudaSuccess;
    while (true) {
        std::vector<int> ab, ae;
        for (size_t s = 0; s < begin.size(); s++)
            if (end[s] - begin[s] > 4) { ab.push_back(begin[s]); ae.push_back(end[s]); }
        const int nseg = (int)ab.size();
        if (nseg == 0) break;
        keepAlive.push_back(ab); keepAlive.push_back(ae);
        if ((e = cudaMemcpyAsync(dBegin, keepAlive[keepAlive.size() - 2].data(), (size_t)nseg * sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(dEnd, keepAlive.back().data(), (size_t)nseg * sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(segOf, 0xff, (size_t)n * sizeof(int), st)) != cudaSuccess) return e;
        const bool wide = nseg <= RTX_RANK_WIDE_SEGS;   // the top levels: a handful of segments of tens of thousands of triangles each
        k_mark_segments<<<dim3(nseg, wide ? RTX_RANK_WIDE_PARTS : 1), 128, 0, st>>>(dBegin, dEnd, nseg, segOf);
        if (wide) {
            k_axis_init<<<1, RTX_RANK_WIDE_SEGS * 6, 0, st>>>(mm, nseg);
            k_axis_part<<<dim3(nseg, RTX_RANK_WIDE_PARTS), 256, 0, st>>>(c[cur][0], c[cur][1], c[cur][2], dBegin, dEnd, mm);
            k_axis_finish<<<1, RTX_RANK_WIDE_SEGS, 0, st>>>(mm, nseg, axis);
        } else if (nseg <= 2048) k_axis<256><<<nseg, 256, 0, st>>>(c[cur][0], c[cur][1], c[cur][2], dBegin, dEnd, nseg, axis);
        else k_axis<32><<<(nseg + 7) / 8, 256, 0, st>>>(c[cur][0], c[cur][1], c[cur][2], dBegin, dEnd, nseg, axis);
        k_keys<<<G, T, 0, st>>>(c[cur][0], c[cur][1], c[cur][2], dBegin, axis, nseg, n, segOf, keyIn, posIn);
        // elements outside the level's segments (finished leaf ranges) keep their place: posOut starts as the identity
        if ((e = cudaMemcpyAsync(posOut, posIn, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
        size_t tb = cubBytesAvail;
        if (wide) {
            // cub's segmented sort gives a segment of this size to ONE block (3.8 ms for the root of 280 K triangles); the device-wide
            // radix sort is stable as well and orders doubles the same way (-0.0 = +0.0), one call per segment
            for (int s = 0; s < nseg; s++) {
                const int b = ab[s], m = ae[s] - ab[s];
                tb = cubBytesAvail;
                if ((e = cub::DeviceRadixSort::SortPairs(cubTemp, tb, keyIn + b, keyOut + b, posIn + b, posOut + b, m, 0, 64, st)) != cudaSuccess) return e;
            }
        } else if ((e = cub::DeviceSegmentedSort::StableSortPairs(cubTemp, tb, keyIn, keyOut, posIn, posOut, n, nseg, dBegin, dEnd, st)) != cudaSuccess) return e;
        k_gather<<<G, T, 0, st>>>(posOut, n, c[cur][0], c[cur][1], c[cur][2], idx[cur], c[cur ^ 1][0], c[cur ^ 1][1], c[cur ^ 1][2], idx[cur ^ 1]);
        cur ^= 1;
        std::vector<int> nb, ne;
        for (size_t s = 0; s < begin.size(); s++) {
            const int m = end[s] - begin[s];
            if (m > 4) { nb.push_back(begin[s]); ne.push_back(begin[s] + m / 2); nb.push_back(begin[s] + m / 2); ne.push_back(end[s]); }
        }
        begin.swap(nb); end.swap(ne);
    }
    k_ranks<<<G, T, 0, st>>>(idx[cur], n, rank);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return cudaStreamSynchronize(st);   // the per-level segment lists above are host vectors: keep them alive until the copies ran
}

}  // namespace rtxrank
