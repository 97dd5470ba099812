// rtx_bvh_gpu.cuh — construction of a mesh's 4-wide BVH (BLAS) ON THE DEVICE (sm_100a), replacing the host build for
// triangle meshes (the reference's counterpart is NewBVHNodeFromList, rt/bvh.go:64-217, which main.go:71-79
// times as "BVH Construction").
//
// Closest-hit results do not depend on the hierarchy (rtx_bvh.hpp), so the builder is free to choose it. It is the same
// algorithm as the host builder — top-down binned SAH (16 bins, all three axes), collapsed into 4-wide nodes by repeatedly
// opening the child with the largest surface area — organised level-synchronously so that every step is a data-parallel
// pass over the triangles or over the nodes of one level:
//
//   k_prim_boxes      float64 triangle bounds (padded like rt/aabb.go:117-128), outward-rounded float32 copies, mesh bounds
//   per binary level  k_bin      every triangle of a splitting node drops its box into 3 x 16 bins   (min/max/count atomics: order-free)
//                     k_split    one thread per node sweeps the bins, picks (axis, plane), creates the two children
//                     k_flags + exclusive scan + k_scatter    stable partition of every node's triangle range
//   per wide level    k_collapse_count + exclusive scan + k_collapse_emit    binary tree -> Node4 records, breadth-first numbering
//   k_tri_emit        triangle records (v0, e1, e2, unit normal / info) in leaf order, float64, reference operation order
//
// Determinism: bins are filled with min / max / integer-add atomics (order-independent), partitions and wide-node
// numbering use exclusive scans, so the tree TOPOLOGY, the leaf contents and the node layout are reproducible; only the
// numbering of the temporary binary nodes (atomicAdd allocation) varies, and nothing depends on it.
// Binning uses each node's BOX extent as the bin domain (the centroids lie inside it), which removes the centroid-bounds
// pass; the host builder bins over the centroid extent, so the two trees differ slightly — both are valid.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "rtx_bvh.hpp"

namespace rtxgpu {

constexpr int NB = 16;                 // SAH bins per axis
constexpr int BIN_WORDS = 7;           // lo.xyz, hi.xyz (order-preserving uint encoding), count
constexpr int NODE_BIN_WORDS = 3 * NB * BIN_WORDS;

__device__ __forceinline__ unsigned f2o(float f) { unsigned u = __float_as_uint(f); return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u); }   // order-preserving
__device__ __forceinline__ float o2f(unsigned o) { return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xffffffffu)); }
__device__ __forceinline__ unsigned long long d2o(double d) { unsigned long long u = (unsigned long long)__double_as_longlong(d); return u ^ ((u >> 63) ? ~0ull : 0x8000000000000000ull); }
__host__ __device__ __forceinline__ double o2d_bits(unsigned long long o) {
    unsigned long long u = o ^ ((o >> 63) ? 0x8000000000000000ull : ~0ull);
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
}
#define RTX_O_PINF 0xff800000u   /* f2o(+inf) */
#define RTX_O_NINF 0x007fffffu   /* f2o(-inf) */

struct BNodes {      // temporary binary tree, SoA
    int* first; int* count; int* left;   // left < 0: leaf; right = left + 1
    float4* b0; float2* b1;              // (lo.x, lo.y, lo.z, hi.x) (hi.y, hi.z)
    int* axis; int* plane;               // split of an internal node: triangles with bin <= plane go left; axis < 0: by position (first count/2)
    int* nleft;
};
struct Ctr { int nodes; int active; int depth_pad; int pad; unsigned long long mesh_lo[3], mesh_hi[3]; };

// ---- exclusive scan of ints (n + 1 outputs: out[n] = total) ---------------------------------------------------------------
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_BLOCK = SCAN_T * SCAN_ITEMS;
__device__ __forceinline__ int block_exclusive_scan(int v, int* total_out) {   // 256 threads
    __shared__ int warp_sums[SCAN_T / 32];
    __shared__ int total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < SCAN_T / 32 ? warp_sums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int d = 1; d < SCAN_T / 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += t; }
        if (lane < SCAN_T / 32) warp_sums[lane] = winc - w;
        if (lane == SCAN_T / 32 - 1) total = winc;
    }
    __syncthreads();
    const int r = inc - v + warp_sums[wid];
    *total_out = total;
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_local(const int* in, int* out, int n, int* sums) {
    const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    int total;
    int pre = block_exclusive_scan(s, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = pre; pre += v[k]; }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_sums(int* sums, int nb, int* total_out) {   // one block; nb block sums -> exclusive, in place
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += SCAN_T) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? sums[i] : 0;
        int total;
        const int pre = block_exclusive_scan(v, &total);
        const int carry = carry_s;
        if (i < nb) sums[i] = carry + pre;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_add(int* out, int n, const int* sums) {
    const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    const int add = sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) out[base + k] += add;
}
// out must hold n + 1 ints, sums ceil(n / 2048) ints
static inline void exclusive_scan(const int* in, int* out, int n, int* sums, cudaStream_t st) {
    const int nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    k_scan_local<<<nb, SCAN_T, 0, st>>>(in, out, n, sums);
    k_scan_sums<<<1, SCAN_T, 0, st>>>(sums, nb, out + n);
    k_scan_add<<<nb, SCAN_T, 0, st>>>(out, n, sums);
}

// ---- triangle bounds ----------------------------------------------------------------------------------------------------------
// Float64 bounds exactly as the host path computes them (min / max of the vertices, flat boxes padded by 1e-4 per side:
// rt/aabb.go:117-128), then float32 copies rounded outward. The mesh bounds are reduced in float64 (order-preserving
// 64-bit atomics), so the TLAS sees the same box the host loop produced.
__global__ void __launch_bounds__(256) k_prim_boxes(const double* v0, const double* v1, const double* v2, int n, float4* b0, float2* b1, int* idx, int* node_of,
                                                   Ctr* ctr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (i < n) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const double x = v0[3 * (size_t)i + a], y = v1[3 * (size_t)i + a], z = v2[3 * (size_t)i + a];
            lo[a] = fmin(fmin(x, y), z);
            hi[a] = fmax(fmax(x, y), z);
            if (hi[a] - lo[a] < 1e-4) { lo[a] -= 1e-4; hi[a] += 1e-4; }
        }
        b0[i] = make_float4(__double2float_rd(lo[0]), __double2float_rd(lo[1]), __double2float_rd(lo[2]), __double2float_ru(hi[0]));
        b1[i] = make_float2(__double2float_ru(hi[1]), __double2float_ru(hi[2]));
        idx[i] = i;
        node_of[i] = 0;
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double l = lo[a], h = hi[a];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { l = fmin(l, __shfl_xor_sync(0xffffffffu, l, d)); h = fmax(h, __shfl_xor_sync(0xffffffffu, h, d)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(&ctr->mesh_lo[a], d2o(l)); atomicMax(&ctr->mesh_hi[a], d2o(h)); }
    }
}
__global__ void k_root_init(BNodes N, Ctr* ctr, int n, int maxLeaf) {
    N.first[0] = 0; N.count[0] = n; N.left[0] = -1;
    double lo[3], hi[3];
    for (int a = 0; a < 3; a++) { lo[a] = o2d_bits(ctr->mesh_lo[a]); hi[a] = o2d_bits(ctr->mesh_hi[a]); }
    N.b0[0] = make_float4(__double2float_rd(lo[0]), __double2float_rd(lo[1]), __double2float_rd(lo[2]), __double2float_ru(hi[0]));
    N.b1[0] = make_float2(__double2float_ru(hi[1]), __double2float_ru(hi[2]));
    ctr->nodes = 1;
    ctr->active = n > maxLeaf ? 1 : 0;
}

// ---- one binary level ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bins_clear(unsigned* bins, int nslots) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nslots * NODE_BIN_WORDS) return;
    const int w = (int)(i % BIN_WORDS);
    bins[i] = w < 3 ? RTX_O_PINF : w < 6 ? RTX_O_NINF : 0u;
}
__device__ __forceinline__ int bin_of(float c, float lo, float k) {
    int b = (int)((c - lo) * k);
    return b < 0 ? 0 : (b > NB - 1 ? NB - 1 : b);
}
// Every triangle of a splitting node drops its box into one bin per axis. On the top levels a block's 256 consecutive
// positions belong to one node: the block then bins into shared memory and flushes 336 words with one global atomic each,
// instead of 256 x 21 global atomics onto the same few addresses (k_bin was 3/4 of the build time before: 5.3 of 7 ms for
// 280 K triangles). Blocks that straddle nodes (the deep levels, where contention is low anyway) go to global memory directly.
__global__ void __launch_bounds__(256) k_bin(const float4* b0, const float2* b1, const int* node_of, int n, BNodes N, int lb, unsigned* bins) {
    __shared__ unsigned sbin[NODE_BIN_WORDS];
    __shared__ int s_node;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int nd = p < n ? node_of[p] : -2;          // -1: finished leaf, -2: past the end
    if (threadIdx.x == 0) s_node = nd;
    for (int i = threadIdx.x; i < NODE_BIN_WORDS; i += blockDim.x) { const int w = i % BIN_WORDS; sbin[i] = w < 3 ? RTX_O_PINF : w < 6 ? RTX_O_NINF : 0u; }
    __syncthreads();
    const int nd0 = s_node;
    const bool uniform = __syncthreads_and(nd == nd0 || nd == -2) && nd0 >= 0;
    if (nd >= 0) {
        const float4 a = b0[p];
        const float2 b = b1[p];
        const float4 n0 = N.b0[nd];
        const float2 n1 = N.b1[nd];
        const float plo[3] = {a.x, a.y, a.z}, phi[3] = {a.w, b.x, b.y};
        const float nlo[3] = {n0.x, n0.y, n0.z}, nhi[3] = {n0.w, n1.x, n1.y};
        unsigned* nb = uniform ? sbin : bins + (size_t)(nd - lb) * NODE_BIN_WORDS;
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            const float ext = nhi[ax] - nlo[ax];
            if (!(ext > 0.f)) continue;
            const int bi = bin_of(0.5f * (plo[ax] + phi[ax]), nlo[ax], (float)NB / ext);
            unsigned* q = nb + (ax * NB + bi) * BIN_WORDS;
            atomicMin(q + 0, f2o(plo[0])); atomicMin(q + 1, f2o(plo[1])); atomicMin(q + 2, f2o(plo[2]));
            atomicMax(q + 3, f2o(phi[0])); atomicMax(q + 4, f2o(phi[1])); atomicMax(q + 5, f2o(phi[2]));
            atomicAdd(q + 6, 1u);
        }
    }
    if (!uniform) return;   // (block-uniform)
    __syncthreads();
    unsigned* gb = bins + (size_t)(nd0 - lb) * NODE_BIN_WORDS;
    for (int i = threadIdx.x; i < NODE_BIN_WORDS; i += blockDim.x) {
        const int w = i % BIN_WORDS;
        if (sbin[i - w + 6] == 0u) continue;   // empty bin
        if (w < 3) atomicMin(gb + i, sbin[i]);
        else if (w < 6) atomicMax(gb + i, sbin[i]);
        else atomicAdd(gb + i, sbin[i]);
    }
}
struct FBox {
    float lo[3], hi[3];
    __device__ __forceinline__ void reset() { lo[0] = lo[1] = lo[2] = INFINITY; hi[0] = hi[1] = hi[2] = -INFINITY; }
    __device__ __forceinline__ void grow(const FBox& o) {
#pragma unroll
        for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], o.lo[a]); hi[a] = fmaxf(hi[a], o.hi[a]); }
    }
    __device__ __forceinline__ float area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};
__device__ __forceinline__ void load_bin(const unsigned* q, FBox& b, int& cnt) {
#pragma unroll
    for (int a = 0; a < 3; a++) { b.lo[a] = o2f(q[a]); b.hi[a] = o2f(q[3 + a]); }
    cnt = (int)q[6];
}
// One thread per node of the level: sweep the bins, choose the cheapest plane, create the children.
__global__ void __launch_bounds__(128) k_split(BNodes N, int lb, int le, const unsigned* bins, int maxLeaf, Ctr* ctr) {
    const int nd = lb + blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= le) return;
    const int count = N.count[nd];
    if (count <= maxLeaf) return;   // a leaf
    const unsigned* nb = bins + (size_t)(nd - lb) * NODE_BIN_WORDS;
    int bestAxis = -1, bestPlane = -1, bestLeft = 0;
    float bestCost = INFINITY;
    FBox bestL, bestR;
    bestL.reset(); bestR.reset();
    for (int ax = 0; ax < 3; ax++) {
        float rightArea[NB];
        int rightCnt[NB];
        FBox acc;
        acc.reset();
        int c = 0;
        for (int b = NB - 1; b > 0; b--) {
            FBox bb; int bc;
            load_bin(nb + (ax * NB + b) * BIN_WORDS, bb, bc);
            acc.grow(bb); c += bc;
            rightArea[b] = acc.area(); rightCnt[b] = c;
        }
        acc.reset(); c = 0;
        for (int b = 0; b < NB - 1; b++) {
            FBox bb; int bc;
            load_bin(nb + (ax * NB + b) * BIN_WORDS, bb, bc);
            acc.grow(bb); c += bc;
            if (c == 0 || rightCnt[b + 1] == 0) continue;
            const float cost = acc.area() * (float)c + rightArea[b + 1] * (float)rightCnt[b + 1];
            if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestPlane = b; bestLeft = c; bestL = acc; }
        }
    }
    const float4 p0 = N.b0[nd];
    const float2 p1 = N.b1[nd];
    if (bestAxis >= 0) {   // right box of the chosen plane
        for (int b = bestPlane + 1; b < NB; b++) {
            FBox bb; int bc;
            load_bin(nb + (bestAxis * NB + b) * BIN_WORDS, bb, bc);
            bestR.grow(bb);
        }
    } else {               // every centroid in one bin on all axes: split by position, children keep the parent's box
        bestLeft = count / 2;
        bestL.lo[0] = p0.x; bestL.lo[1] = p0.y; bestL.lo[2] = p0.z; bestL.hi[0] = p0.w; bestL.hi[1] = p1.x; bestL.hi[2] = p1.y;
        bestR = bestL;
    }
    const int c0 = atomicAdd(&ctr->nodes, 2);
    const int first = N.first[nd];
    N.left[nd] = c0; N.axis[nd] = bestAxis; N.plane[nd] = bestPlane; N.nleft[nd] = bestLeft;
    N.first[c0] = first; N.count[c0] = bestLeft; N.left[c0] = -1;
    N.first[c0 + 1] = first + bestLeft; N.count[c0 + 1] = count - bestLeft; N.left[c0 + 1] = -1;
    N.b0[c0] = make_float4(bestL.lo[0], bestL.lo[1], bestL.lo[2], bestL.hi[0]); N.b1[c0] = make_float2(bestL.hi[1], bestL.hi[2]);
    N.b0[c0 + 1] = make_float4(bestR.lo[0], bestR.lo[1], bestR.lo[2], bestR.hi[0]); N.b1[c0 + 1] = make_float2(bestR.hi[1], bestR.hi[2]);
    const int act = (bestLeft > maxLeaf) + (count - bestLeft > maxLeaf);
    if (act) atomicAdd(&ctr->active, act);
}
__device__ __forceinline__ bool goes_left(const float4& a, const float2& b, const BNodes& N, int nd, int p) {
    const int ax = N.axis[nd];
    if (ax < 0) return p - N.first[nd] < N.nleft[nd];
    const float4 n0 = N.b0[nd];
    const float2 n1 = N.b1[nd];
    const float lo = ax == 0 ? n0.x : ax == 1 ? n0.y : n0.z, hi = ax == 0 ? n0.w : ax == 1 ? n1.x : n1.y;
    const float plo = ax == 0 ? a.x : ax == 1 ? a.y : a.z, phi = ax == 0 ? a.w : ax == 1 ? b.x : b.y;
    return bin_of(0.5f * (plo + phi), lo, (float)NB / (hi - lo)) <= N.plane[nd];
}
__global__ void __launch_bounds__(256) k_flags(const float4* b0, const float2* b1, const int* node_of, int n, BNodes N, int* flag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int nd = node_of[p];
    flag[p] = (nd >= 0 && goes_left(b0[p], b1[p], N, nd, p)) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_scatter(const float4* b0, const float2* b1, const int* idx, const int* node_of, int n, BNodes N, const int* flag,
                                                const int* scan, int maxLeaf, float4* b0o, float2* b1o, int* idxo, int* node_ofo) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int nd = node_of[p];
    int q = p, child = -1;
    if (nd >= 0) {
        const int first = N.first[nd], nl = N.nleft[nd];
        const int before = scan[p] - scan[first];   // triangles of this node in front of p that go left
        const bool left = flag[p] != 0;
        q = left ? first + before : first + nl + (p - first - before);
        const int c = N.left[nd] + (left ? 0 : 1);
        child = N.count[c] > maxLeaf ? c : -1;
    }
    b0o[q] = b0[p]; b1o[q] = b1[p]; idxo[q] = idx[p]; node_ofo[q] = child;
}

// ---- binary -> 4-wide ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float node_area(const BNodes& N, int b) {
    const float4 a = N.b0[b];
    const float2 c = N.b1[b];
    const float dx = a.w - a.x, dy = c.x - a.y, dz = c.y - a.z;
    if (dx < 0 || dy < 0 || dz < 0) return 0.f;
    return 2.f * (dx * dy + dy * dz + dz * dx);
}
__global__ void __launch_bounds__(128) k_collapse_count(BNodes N, const int* item_b, int nitems, int4* kids_out, int* cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    const int b = item_b[i];
    int kids[4] = {-1, -1, -1, -1}, nk = 0;
    if (N.left[b] < 0) kids[nk++] = b;   // a single leaf as root
    else { kids[nk++] = N.left[b]; kids[nk++] = N.left[b] + 1; }
    while (nk < 4) {   // open the internal child with the largest surface area
        int pick = -1;
        float best = -1.f;
        for (int k = 0; k < nk; k++)
            if (N.left[kids[k]] >= 0) {
                const float a = node_area(N, kids[k]);
                if (a > best) { best = a; pick = k; }
            }
        if (pick < 0) break;
        const int k = kids[pick];
        kids[pick] = N.left[k];
        kids[nk++] = N.left[k] + 1;
    }
    int c = 0;
    for (int k = 0; k < nk; k++) c += N.left[kids[k]] >= 0;
    kids_out[i] = make_int4(kids[0], kids[1], kids[2], kids[3]);
    cnt[i] = c;
}
__global__ void __launch_bounds__(128) k_collapse_emit(BNodes N, const int* item_b, const int* item_o, int nitems, const int4* kids_in, const int* off, int node_base,
                                                      int next_local, int tri_base, rtxbvh::Node4* out, int* next_b, int* next_o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    const int4 kv = kids_in[i];
    const int kids[4] = {kv.x, kv.y, kv.z, kv.w};
    rtxbvh::Node4 node;
    int slot = off[i];
    for (int k = 0; k < 4; k++) {
        node.pad[k] = 0;
        const int b = kids[k];
        if (b < 0) {
            node.lox[k] = node.loy[k] = node.loz[k] = INFINITY;
            node.hix[k] = node.hiy[k] = node.hiz[k] = -INFINITY;
            node.child[k] = -1;
            continue;
        }
        const float4 a = N.b0[b];
        const float2 c = N.b1[b];
        node.lox[k] = a.x; node.loy[k] = a.y; node.loz[k] = a.z; node.hix[k] = a.w; node.hiy[k] = c.x; node.hiz[k] = c.y;
        if (N.left[b] < 0) {
            node.child[k] = ~(((tri_base + N.first[b]) << 3) | (N.count[b] - 1));
        } else {
            const int local = next_local + slot;
            node.child[k] = node_base + local;
            next_b[slot] = b; next_o[slot] = local;
            slot++;
        }
    }
    out[item_o[i]] = node;
}

// ---- triangle records in leaf order (rt/triangle.go:17-25: e1, e2, unit normal; same float64 operation order as the host path) ----
__global__ void __launch_bounds__(256) k_tri_emit(const double* v0, const double* v1, const double* v2, const int* mat, const int* rank, const int* idx, int n,
                                                 double* tris, int4* info) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int s = idx[p];
    const double* a = v0 + 3 * (size_t)s; const double* b = v1 + 3 * (size_t)s; const double* c = v2 + 3 * (size_t)s;
    const double a0 = a[0], a1 = a[1], a2 = a[2];
    const double e1x = b[0] - a0, e1y = b[1] - a1, e1z = b[2] - a2, e2x = c[0] - a0, e2y = c[1] - a1, e2z = c[2] - a2;
    double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
    const double l = sqrt(nx * nx + ny * ny + nz * nz);
    if (l != 0) { const double inv = 1 / l; nx = inv * nx; ny = inv * ny; nz = inv * nz; }
    double2* t = reinterpret_cast<double2*>(tris + 12 * (size_t)p);   // RTX_TRI_D doubles: v0, e1, e2, n
    t[0] = make_double2(a0, a1); t[1] = make_double2(a2, e1x); t[2] = make_double2(e1y, e1z); t[3] = make_double2(e2x, e2y); t[4] = make_double2(e2z, nx);
    t[5] = make_double2(ny, nz);
    info[p] = make_int4(s, mat[s], rank[s], 0);
}

struct BlasResult {
    int n_nodes = 0;     // wide nodes written to `nodes_out` (local numbering; child links already carry node_base)
    int depth = 0;       // wide levels
    int binary_levels = 0;
    double lo[3], hi[3]; // float64 bounds of the mesh (union of the padded triangle boxes)
};

// Builds the BLAS of one mesh. Device inputs: v0/v1/v2 [3n] float64, mat/rank [n]. Device outputs: nodes_out [>= n] Node4 (local
// index 0 = root), tris/info for the n triangles in leaf order (written at the pointers given: the caller passes
// base + tri_base offsets). Leaf codes carry tri_base, internal links node_base. Returns cudaSuccess or the failing call's error.
// The working memory comes from the caller (no cudaMalloc / cudaFree here: they serialise against the whole driver): call once
// with scratch == nullptr to get the size in *scratch_bytes, then with a buffer of at least that size.
static inline cudaError_t build_blas(const double* v0, const double* v1, const double* v2, const int* mat, const int* rank, int n, int maxLeaf, int node_base,
                                     int tri_base, rtxbvh::Node4* nodes_out, double* tris, int4* info, cudaStream_t st, BlasResult* res,
                                     const char** what, char* scratch, size_t* scratch_bytes) {
#define RTX_G(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { *what = #call; return e_; } } while (0)
    *what = "";
    const size_t N2 = (size_t)2 * n + 2;                               // binary nodes
    const int maxActive = n / (maxLeaf + 1) + 2;                       // nodes of one level that still split
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t oB0a = carve(n * sizeof(float4)), oB0b = carve(n * sizeof(float4)), oB1a = carve(n * sizeof(float2)), oB1b = carve(n * sizeof(float2));
    const size_t oIdxA = carve(n * sizeof(int)), oIdxB = carve(n * sizeof(int)), oNofA = carve(n * sizeof(int)), oNofB = carve(n * sizeof(int));
    const size_t oFlag = carve((size_t)(n + 1) * sizeof(int)), oScan = carve((size_t)(n + 1) * sizeof(int)), oSums = carve(((size_t)n / SCAN_BLOCK + 2) * sizeof(int));
    const size_t oNf = carve(N2 * sizeof(int)), oNc = carve(N2 * sizeof(int)), oNl = carve(N2 * sizeof(int)), oNa = carve(N2 * sizeof(int)), oNp = carve(N2 * sizeof(int)),
                 oNn = carve(N2 * sizeof(int)), oN0 = carve(N2 * sizeof(float4)), oN1 = carve(N2 * sizeof(float2));
    const size_t oBins = carve(((size_t)2 * maxActive + 2) * NODE_BIN_WORDS * sizeof(unsigned));   // a level = the children of <= maxActive splitting nodes
    const size_t oItA = carve((size_t)n * sizeof(int)), oItB = carve((size_t)n * sizeof(int)), oIoA = carve((size_t)n * sizeof(int)), oIoB = carve((size_t)n * sizeof(int));
    const size_t oKids = carve((size_t)n * sizeof(int4)), oCnt = carve((size_t)(n + 1) * sizeof(int)), oCoff = carve((size_t)(n + 1) * sizeof(int));
    const size_t oCtr = carve(sizeof(Ctr));
    if (!scratch) { *scratch_bytes = off; return cudaSuccess; }   // sizing call
    if (*scratch_bytes < off) { *what = "build scratch too small"; return cudaErrorInvalidValue; }
    float4* b0[2] = {(float4*)(scratch + oB0a), (float4*)(scratch + oB0b)};
    float2* b1[2] = {(float2*)(scratch + oB1a), (float2*)(scratch + oB1b)};
    int* idx[2] = {(int*)(scratch + oIdxA), (int*)(scratch + oIdxB)};
    int* nof[2] = {(int*)(scratch + oNofA), (int*)(scratch + oNofB)};
    int *flag = (int*)(scratch + oFlag), *scan = (int*)(scratch + oScan), *sums = (int*)(scratch + oSums);
    BNodes N{(int*)(scratch + oNf), (int*)(scratch + oNc), (int*)(scratch + oNl), (float4*)(scratch + oN0), (float2*)(scratch + oN1), (int*)(scratch + oNa),
             (int*)(scratch + oNp), (int*)(scratch + oNn)};
    unsigned* bins = (unsigned*)(scratch + oBins);
    int* itb[2] = {(int*)(scratch + oItA), (int*)(scratch + oItB)};
    int* ito[2] = {(int*)(scratch + oIoA), (int*)(scratch + oIoB)};
    int4* kids = (int4*)(scratch + oKids);
    int *cnt = (int*)(scratch + oCnt), *coff = (int*)(scratch + oCoff);
    Ctr* ctr = (Ctr*)(scratch + oCtr);
    Ctr h{};
    for (int a = 0; a < 3; a++) { h.mesh_lo[a] = ~0ull; h.mesh_hi[a] = 0ull; }
    RTX_G(cudaMemcpyAsync(ctr, &h, sizeof h, cudaMemcpyHostToDevice, st));
    const int gridN = (n + 255) / 256;
    k_prim_boxes<<<gridN, 256, 0, st>>>(v0, v1, v2, n, b0[0], b1[0], idx[0], nof[0], ctr);
    k_root_init<<<1, 1, 0, st>>>(N, ctr, n, maxLeaf);
    RTX_G(cudaMemcpyAsync(&h, ctr, sizeof h, cudaMemcpyDeviceToHost, st));
    RTX_G(cudaStreamSynchronize(st));
    for (int a = 0; a < 3; a++) { res->lo[a] = o2d_bits(h.mesh_lo[a]); res->hi[a] = o2d_bits(h.mesh_hi[a]); }
    int cur = 0, lb = 0, le = 1, levels = 0;
    while (h.active > 0) {
        if (++levels > 128) { *what = "binary BVH deeper than 128 levels"; return cudaErrorUnknown; }
        const int nlev = le - lb;
        if (nlev > maxActive * 2 + 2) { *what = "level wider than the bin scratch"; return cudaErrorUnknown; }
        // bins are indexed by (node - lb); only splitting nodes use theirs
        const size_t words = (size_t)nlev * NODE_BIN_WORDS;
        k_bins_clear<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(bins, nlev);
        k_bin<<<gridN, 256, 0, st>>>(b0[cur], b1[cur], nof[cur], n, N, lb, bins);
        RTX_G(cudaMemsetAsync(&ctr->active, 0, sizeof(int), st));
        k_split<<<(nlev + 127) / 128, 128, 0, st>>>(N, lb, le, bins, maxLeaf, ctr);
        k_flags<<<gridN, 256, 0, st>>>(b0[cur], b1[cur], nof[cur], n, N, flag);
        exclusive_scan(flag, scan, n, sums, st);
        k_scatter<<<gridN, 256, 0, st>>>(b0[cur], b1[cur], idx[cur], nof[cur], n, N, flag, scan, maxLeaf, b0[cur ^ 1], b1[cur ^ 1], idx[cur ^ 1], nof[cur ^ 1]);
        RTX_G(cudaMemcpyAsync(&h, ctr, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        RTX_G(cudaStreamSynchronize(st));
        cur ^= 1;
        lb = le; le = h.nodes;
    }
    res->binary_levels = levels;
    // collapse, breadth-first
    int nitems = 1, total = 1, it = 0, depth = 0;
    const int zero = 0;
    RTX_G(cudaMemcpyAsync(itb[0], &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    RTX_G(cudaMemcpyAsync(ito[0], &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    while (nitems > 0) {
        depth++;
        k_collapse_count<<<(nitems + 127) / 128, 128, 0, st>>>(N, itb[it], nitems, kids, cnt);
        exclusive_scan(cnt, coff, nitems, sums, st);
        k_collapse_emit<<<(nitems + 127) / 128, 128, 0, st>>>(N, itb[it], ito[it], nitems, kids, coff, node_base, total, tri_base, nodes_out, itb[it ^ 1], ito[it ^ 1]);
        int next = 0;
        RTX_G(cudaMemcpyAsync(&next, coff + nitems, sizeof(int), cudaMemcpyDeviceToHost, st));
        RTX_G(cudaStreamSynchronize(st));
        total += next;
        nitems = next;
        it ^= 1;
    }
    res->n_nodes = total; res->depth = depth;
    k_tri_emit<<<gridN, 256, 0, st>>>(v0, v1, v2, mat, rank, idx[cur], n, tris, info);
    RTX_G(cudaGetLastError());
    RTX_G(cudaStreamSynchronize(st));
    return cudaSuccess;
#undef RTX_G
}

}  // namespace rtxgpu
