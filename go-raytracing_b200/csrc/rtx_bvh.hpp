// rtx_bvh.hpp — host-side construction of the 4-wide, 128-byte-aligned BVH the device traverses, and of the
// canonical "reference test order" ranks used only to resolve exact ties in t.
//
// The reference builds a binary median-split BVH (rt/bvh.go:120-217). Closest-hit results do not depend on the
// hierarchy, so the device BVH is built for traversal speed instead: binned-SAH binary tree collapsed into
// 4-wide nodes whose child boxes are float32, rounded OUTWARD from the float64 primitive bounds.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace rtxbvh {

struct Box {
    double lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = std::numeric_limits<double>::infinity(); hi[a] = -lo[a]; } }
    void grow(const Box& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const double* p) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    double area() const {
        double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0;
        return 2 * (dx * dy + dy * dz + dz * dx);
    }
    bool finite() const {
        for (int a = 0; a < 3; a++) if (!std::isfinite(lo[a]) || !std::isfinite(hi[a])) return false;
        return true;
    }
};

struct Node4 {  // 128 bytes; matches the 8 x float4 the device loads
    float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];  // {lo,hi} pairs per axis: the device picks near/far by a 16-byte offset
    int32_t child[4];  // >= 0 internal node index (global); < 0 leaf code (~child); empty: box inverted
    int32_t pad[4];
};
static_assert(sizeof(Node4) == 128, "Node4 must be 128 bytes");

inline float round_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
inline float round_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}

struct Builder {
    const std::vector<Box>& boxes;
    std::vector<int> idx;  // permutation of primitives; leaves reference contiguous ranges of it
    int maxLeaf;
    struct N2 { Box box; int left = -1, right = -1, first = 0, count = 0; };
    std::vector<N2> n2;

    Builder(const std::vector<Box>& b, int maxLeafSize) : boxes(b), maxLeaf(maxLeafSize) {
        idx.resize(b.size());
        for (size_t i = 0; i < b.size(); i++) idx[i] = (int)i;
        n2.reserve(b.size() * 2 + 1);
    }
    int build(int first, int count) {
        int me = (int)n2.size();
        n2.emplace_back();
        Box bb, cb;
        bb.reset(); cb.reset();
        for (int i = first; i < first + count; i++) {
            const Box& b = boxes[idx[i]];
            bb.grow(b);
            double c[3] = {0.5 * (b.lo[0] + b.hi[0]), 0.5 * (b.lo[1] + b.hi[1]), 0.5 * (b.lo[2] + b.hi[2])};
            cb.grow(c);
        }
        n2[me].box = bb; n2[me].first = first; n2[me].count = count;
        if (count <= maxLeaf) return me;
        // binned SAH over the axis of largest centroid extent (16 bins), all three axes evaluated
        const int NB = 16;
        int bestAxis = -1, bestBin = -1;
        double bestCost = std::numeric_limits<double>::infinity();
        for (int axis = 0; axis < 3; axis++) {
            double lo = cb.lo[axis], ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0)) continue;
            Box bins[NB];
            int cnt[NB] = {0};
            for (int b = 0; b < NB; b++) bins[b].reset();
            double k = NB / ext;
            for (int i = first; i < first + count; i++) {
                const Box& b = boxes[idx[i]];
                int bi = (int)((0.5 * (b.lo[axis] + b.hi[axis]) - lo) * k);
                bi = std::min(std::max(bi, 0), NB - 1);
                bins[bi].grow(b);
                cnt[bi]++;
            }
            double rightArea[NB];
            int rightCnt[NB];
            Box acc; acc.reset();
            int c = 0;
            for (int b = NB - 1; b > 0; b--) { acc.grow(bins[b]); c += cnt[b]; rightArea[b] = acc.area(); rightCnt[b] = c; }
            acc.reset(); c = 0;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bins[b]); c += cnt[b];
                if (c == 0 || rightCnt[b + 1] == 0) continue;
                double cost = acc.area() * c + rightArea[b + 1] * rightCnt[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
            }
        }
        int mid;
        if (bestAxis < 0) {
            mid = first + count / 2;  // all centroids coincide: split by count
        } else {
            double lo = cb.lo[bestAxis], k = NB / (cb.hi[bestAxis] - cb.lo[bestAxis]);
            auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](int p) {
                const Box& b = boxes[p];
                int bi = (int)((0.5 * (b.lo[bestAxis] + b.hi[bestAxis]) - lo) * k);
                bi = std::min(std::max(bi, 0), NB - 1);
                return bi <= bestBin;
            });
            mid = (int)(it - idx.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        int l = build(first, mid - first);
        int r = build(mid, first + count - mid);
        n2[me].left = l; n2[me].right = r;
        return me;
    }
};

// Builds a 4-wide BVH over `boxes`. Leaf code for a leaf covering permuted range [first, first+count):
//   leafCode(first, count) -> non-negative int; the node stores ~code. Appends nodes to `out` (global indices).
// Returns the root node index (global) or -1 for an empty input. `perm` receives the primitive permutation.
template <class LeafCode>
int build_bvh4(const std::vector<Box>& boxes, int maxLeaf, std::vector<Node4>& out, std::vector<int>& perm, LeafCode leafCode, int* maxDepth = nullptr) {
    if (maxDepth) *maxDepth = 0;
    if (boxes.empty()) { perm.clear(); return -1; }
    Builder B(boxes, maxLeaf);
    int root2 = B.build(0, (int)boxes.size());
    perm = B.idx;
    struct Work { int n2; int outIndex; int depth; };
    int rootOut = (int)out.size();
    out.emplace_back();
    std::vector<Work> stack;
    stack.push_back({root2, rootOut, 1});
    while (!stack.empty()) {
        Work w = stack.back();
        stack.pop_back();
        if (maxDepth && w.depth > *maxDepth) *maxDepth = w.depth;
        int kids[4];
        int nk = 0;
        const auto& n = B.n2[w.n2];
        if (n.left < 0) { kids[nk++] = w.n2; }  // a single leaf as root
        else { kids[nk++] = n.left; kids[nk++] = n.right; }
        while (nk < 4) {  // expand the internal child with the largest surface area
            int pick = -1;
            double best = -1;
            for (int i = 0; i < nk; i++)
                if (B.n2[kids[i]].left >= 0) {
                    double a = B.n2[kids[i]].box.area();
                    if (a > best) { best = a; pick = i; }
                }
            if (pick < 0) break;
            int k = kids[pick];
            kids[pick] = B.n2[k].left;
            kids[nk++] = B.n2[k].right;
        }
        Node4 node;
        for (int i = 0; i < 4; i++) {
            if (i < nk) {
                const Box& b = B.n2[kids[i]].box;
                node.lox[i] = round_down(b.lo[0]); node.loy[i] = round_down(b.lo[1]); node.loz[i] = round_down(b.lo[2]);
                node.hix[i] = round_up(b.hi[0]); node.hiy[i] = round_up(b.hi[1]); node.hiz[i] = round_up(b.hi[2]);
                if (B.n2[kids[i]].left < 0) {
                    node.child[i] = ~leafCode(B.n2[kids[i]].first, B.n2[kids[i]].count);
                } else {
                    int oi = (int)out.size();
                    out.emplace_back();
                    node.child[i] = oi;
                    stack.push_back({kids[i], oi, w.depth + 1});
                }
            } else {
                float inf = std::numeric_limits<float>::infinity();
                node.lox[i] = node.loy[i] = node.loz[i] = inf;
                node.hix[i] = node.hiy[i] = node.hiz[i] = -inf;
                node.child[i] = -1;
            }
            node.pad[i] = 0;
        }
        out[w.outIndex] = node;
    }
    return rootOut;
}

// ---- canonical reference test order (rt/bvh.go:120-217 with a stable sort; rt/aabb.go:117-159 conventions) ----------
// Used when the caller does not pass the ranks of its own Go-built tree. rank[i] = position of primitive i in the
// depth-first leaf order of the reference's binary BVH.
struct RefPrim { int index; Box box; double c[3]; };
template <class Less>
inline void stable_sort_ref(std::vector<RefPrim>& a, size_t lo, size_t hi, std::vector<RefPrim>& tmp, Less less) {
    size_t n = hi - lo;
    if (n <= 12) {
        for (size_t i = lo + 1; i < hi; i++) {
            RefPrim x = a[i];
            size_t j = i;
            while (j > lo && less(x, a[j - 1])) { a[j] = a[j - 1]; j--; }
            a[j] = x;
        }
        return;
    }
    size_t mid = lo + n / 2;
    stable_sort_ref(a, lo, mid, tmp, less);
    stable_sort_ref(a, mid, hi, tmp, less);
    size_t i = lo, j = mid, k = lo;
    while (i < mid && j < hi) tmp[k++] = less(a[j], a[i]) ? a[j++] : a[i++];
    while (i < mid) tmp[k++] = a[i++];
    while (j < hi) tmp[k++] = a[j++];
    for (size_t t = lo; t < hi; t++) a[t] = tmp[t];
}
inline void ref_order_rec(std::vector<RefPrim>& p, size_t lo, size_t hi, std::vector<RefPrim>& tmp, std::vector<int>& rank, int& next) {
    size_t n = hi - lo;
    if (n <= 4) {
        for (size_t i = lo; i < hi; i++) rank[p[i].index] = next++;
        return;
    }
    // centroid bounds with the reference's NaN behaviour (an infinite Plane has a NaN centroid, rt/aabb.go:153-159):
    // comparisons against NaN are false, so a NaN first element poisons the interval exactly as in Go.
    double cmin[3], cmax[3];
    for (int a = 0; a < 3; a++) { cmin[a] = p[lo].c[a]; cmax[a] = p[lo].c[a]; }
    for (size_t i = lo + 1; i < hi; i++)
        for (int a = 0; a < 3; a++) {
            if (p[i].c[a] < cmin[a]) cmin[a] = p[i].c[a];
            if (p[i].c[a] > cmax[a]) cmax[a] = p[i].c[a];
        }
    double sz[3];
    // every centroid point box is padded by 1e-4 per side (NewAABBFromPoints -> padToMinimums) before the union
    for (int a = 0; a < 3; a++) sz[a] = (cmax[a] + 0.0001) - (cmin[a] - 0.0001);
    int axis = (sz[0] > sz[1] && sz[0] > sz[2]) ? 0 : (sz[1] > sz[2] ? 1 : 2);
    stable_sort_ref(p, lo, hi, tmp, [axis](const RefPrim& a, const RefPrim& b) { return a.c[axis] < b.c[axis]; });
    size_t mid = lo + n / 2;
    ref_order_rec(p, lo, mid, tmp, rank, next);
    ref_order_rec(p, mid, hi, tmp, rank, next);
}
inline std::vector<int> canonical_ranks(const std::vector<Box>& boxes) {
    std::vector<int> rank(boxes.size(), 0);
    if (boxes.empty()) return rank;
    std::vector<RefPrim> p(boxes.size()), tmp(boxes.size());
    for (size_t i = 0; i < boxes.size(); i++) {
        p[i].index = (int)i; p[i].box = boxes[i];
        for (int a = 0; a < 3; a++) p[i].c[a] = (boxes[i].lo[a] + boxes[i].hi[a]) * 0.5;
    }
    int next = 0;
    ref_order_rec(p, 0, p.size(), tmp, rank, next);
    return rank;
}

}  // namespace rtxbvh
