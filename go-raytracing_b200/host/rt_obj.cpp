// rt_obj.cpp — LoadOBJ (rt/obj_loader.go:15-113) with a parallel text parse.
//
// The reference scans the file line by line with bufio.Scanner on one goroutine. Here the file is read in one piece and cut
// at line boundaries into one chunk per thread. Phase 1 (parallel): every chunk is tokenised into its own vertex list and
// face list; a face keeps its indices as written plus the number of vertices its own chunk had seen before it, so that the
// two things that depend on file order — a negative index ("from the end", :69-72) and the bounds check against the
// vertices read SO FAR (:86-90) — can be resolved once the chunks' vertex counts are prefix-summed. Phase 2 (parallel):
// faces become triangles in file order (fan triangulation :79-97). The result — vertices, triangles, their order, and the
// first error by line number — is the one the sequential scan produces; threads = 1 is that scan.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

#include "rt.hpp"

namespace rt {
namespace {

struct ObjFace {
    uint32_t first, count;   // indices [first, first + count) of the chunk's index array
    uint32_t vseen;          // vertices of this chunk before the face
    uint32_t line;           // line number inside the chunk (1-based)
};
struct ObjChunk {
    const char *begin = nullptr, *end = nullptr;
    std::vector<Point3> vertices;
    std::vector<long> indices;
    std::vector<ObjFace> faces;
    size_t lines = 0, tris = 0;
    size_t err_line = 0;      // first syntax error of the chunk (line inside the chunk), 0 = none
    std::string err;
    // after the prefix sums
    size_t vbase = 0, lbase = 0, tbase = 0;
    bool active = false;      // lies before the first syntax error
};

inline bool isBlank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// strconv.ParseFloat(tok, 64): the whole token must be a number. Plain decimals with at most 15 significant digits and a
// power of ten of at most 22 take Clinger's fast path — the digits as an integer below 2^53 and 10^k are both exact doubles,
// so ONE IEEE multiplication or division is the correctly rounded result, the same value strtod and Go return; everything
// else (more digits, large exponents, inf / nan / hex) goes to strtod.
inline bool parseDouble(const char* b, const char* e, double& out) {
    static const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    {
        const char* p = b;
        bool neg = false;
        if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
        uint64_t m = 0;
        int digits = 0, frac = 0, any = 0;
        for (; p < e && *p >= '0' && *p <= '9'; p++, any++) { m = m * 10 + (uint64_t)(*p - '0'); digits += (m != 0); if (digits > 15) goto slow; }
        if (p < e && *p == '.') {
            p++;
            for (; p < e && *p >= '0' && *p <= '9'; p++, any++, frac++) { m = m * 10 + (uint64_t)(*p - '0'); digits += (m != 0); if (digits > 15) goto slow; }
        }
        if (!any) goto slow;
        int ex = 0;
        if (p < e && (*p == 'e' || *p == 'E')) {
            p++;
            bool eneg = false;
            if (p < e && (*p == '-' || *p == '+')) { eneg = *p == '-'; p++; }
            if (p >= e) goto slow;
            for (; p < e && *p >= '0' && *p <= '9'; p++) { ex = ex * 10 + (*p - '0'); if (ex > 400) goto slow; }
            if (eneg) ex = -ex;
        }
        if (p != e) goto slow;
        ex -= frac;
        if (ex < -22 || ex > 22) goto slow;
        double v = (double)m;
        v = ex < 0 ? v / P10[-ex] : v * P10[ex];
        out = neg ? -v : v;
        return true;
    }
slow:
    char buf[64];
    size_t n = (size_t)(e - b);
    if (n == 0 || n >= sizeof buf) return false;
    std::memcpy(buf, b, n);
    buf[n] = 0;
    char* end;
    out = std::strtod(buf, &end);
    return end == buf + n;
}
// strconv.Atoi of the part before the first '/': optional sign, then digits only
inline bool parseIndex(const char* b, const char* e, long& out) {
    const char* s = b;
    while (s < e && *s != '/') s++;
    e = s;
    if (b == e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') { neg = *b == '-'; b++; }
    if (b == e) return false;
    long v = 0;
    for (; b < e; b++) {
        if (*b < '0' || *b > '9') return false;
        v = v * 10 + (*b - '0');
        if (v > (1l << 40)) return false;
    }
    out = neg ? -v : v;
    return true;
}

void parseChunk(ObjChunk& c) {
    const char* p = c.begin;
    const char* tok[2][64];   // begin / end of up to 64 fields; longer faces fall back to a vector
    std::vector<std::pair<const char*, const char*>> big;
    while (p < c.end) {
        const char* eol = (const char*)std::memchr(p, '\n', (size_t)(c.end - p));
        if (!eol) eol = c.end;
        c.lines++;
        const char* q = p;
        while (q < eol && isBlank(*q)) q++;
        if (q < eol && *q != '#') {
            // strings.Fields
            int n = 0;
            big.clear();
            const char* s = q;
            while (s < eol) {
                while (s < eol && isBlank(*s)) s++;
                if (s >= eol) break;
                const char* b = s;
                while (s < eol && !isBlank(*s)) s++;
                if (n < 64) { tok[0][n] = b; tok[1][n] = s; }
                else big.emplace_back(b, s);
                n++;
            }
            auto B = [&](int i) { return i < 64 ? tok[0][i] : big[(size_t)i - 64].first; };
            auto E = [&](int i) { return i < 64 ? tok[1][i] : big[(size_t)i - 64].second; };
            if (n > 0 && E(0) - B(0) == 1) {
                if (*B(0) == 'v') {
                    double x, y, z;
                    if (n < 4) { c.err_line = c.lines; c.err = "invalid vertex at line "; return; }
                    if (!parseDouble(B(1), E(1), x) || !parseDouble(B(2), E(2), y) || !parseDouble(B(3), E(3), z)) {
                        c.err_line = c.lines; c.err = "invalid vertex coordinates at line "; return;
                    }
                    c.vertices.push_back({x, y, z});
                } else if (*B(0) == 'f' && n >= 4) {
                    ObjFace f;
                    f.first = (uint32_t)c.indices.size(); f.count = (uint32_t)(n - 1); f.vseen = (uint32_t)c.vertices.size(); f.line = (uint32_t)c.lines;
                    for (int i = 1; i < n; i++) {
                        long idx;
                        if (!parseIndex(B(i), E(i), idx)) { c.err_line = c.lines; c.err = "invalid face index at line "; return; }
                        c.indices.push_back(idx);
                    }
                    c.faces.push_back(f);
                    c.tris += (size_t)(n - 3);
                }
            }
        }
        p = eol + 1;
    }
}

}  // namespace

// Parses `filename`; fills the vertices and, per triangle, three vertex indices (file order, fan triangulation). Throws the
// reference's error (first by line number) on malformed input.
void ParseOBJ(const std::string& filename, int threads, std::vector<Point3>& vertices, std::vector<uint32_t>& tri_indices) {
    const auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (std::getenv("RT_DEBUG_TIMING")) std::fprintf(stderr, "[rt] ParseOBJ %s at %.1f ms\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count() * 1e3);
    };
    FILE* f = std::fopen(filename.c_str(), "rb");
    if (!f) throw std::runtime_error("failed to open OBJ file: " + filename);
    std::fseek(f, 0, SEEK_END);
    long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> text((size_t)std::max(size, 0l));
    size_t got = size > 0 ? std::fread(text.data(), 1, (size_t)size, f) : 0;
    std::fclose(f);
    if ((long)got != size) throw std::runtime_error("error reading OBJ file: " + filename);
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, text.size() / (256 << 10)));   // at least 256 KB of text per thread
    std::vector<ObjChunk> chunks((size_t)threads);
    const char* base = text.data();
    const char* end = base + text.size();
    const char* cut = base;
    for (int t = 0; t < threads; t++) {
        chunks[(size_t)t].begin = cut;
        const char* next = t + 1 == threads ? end : base + text.size() * (size_t)(t + 1) / (size_t)threads;
        if (next < cut) next = cut;
        if (t + 1 < threads) {   // move the cut to just behind the next newline
            const char* nl = (const char*)std::memchr(next, '\n', (size_t)(end - next));
            next = nl ? nl + 1 : end;
        }
        chunks[(size_t)t].end = next;
        cut = next;
    }
    auto runAll = [&](auto fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(fn, t);
        fn(0);
        for (auto& x : th) x.join();
    };
    lap("file read");
    runAll([&](int t) { parseChunk(chunks[(size_t)t]); });
    lap("chunks tokenised");

    // prefix sums; a syntax error ends the scan where the reference's would end
    size_t nv = 0, nl = 0, nt = 0, err_line = 0;
    std::string err;
    for (auto& c : chunks) {
        c.vbase = nv; c.lbase = nl; c.tbase = nt; c.active = true;
        nv += c.vertices.size(); nl += c.lines; nt += c.tris;
        if (c.err_line) { err_line = c.lbase + c.err_line; err = c.err; break; }
    }
    vertices.resize(nv);
    tri_indices.assign(3 * nt, 0);
    std::vector<size_t> bound_err((size_t)threads, 0);   // first out-of-bounds face of each chunk (global line number)
    runAll([&](int t) {
        ObjChunk& c = chunks[(size_t)t];
        if (!c.active) return;
        std::copy(c.vertices.begin(), c.vertices.end(), vertices.begin() + (long)c.vbase);
        size_t k = c.tbase;
        for (const ObjFace& fc : c.faces) {
            const long seen = (long)(c.vbase + fc.vseen);   // len(vertices) when the reference reaches this line
            const long* ix = c.indices.data() + fc.first;
            auto at = [&](uint32_t i) { long v = ix[i]; if (v < 0) v = seen + v + 1; return v - 1; };
            const long i0 = at(0);
            for (uint32_t i = 1; i + 1 < fc.count; i++, k++) {
                const long i1 = at(i), i2 = at(i + 1);
                if (i0 < 0 || i0 >= seen || i1 < 0 || i1 >= seen || i2 < 0 || i2 >= seen) {
                    bound_err[(size_t)t] = c.lbase + fc.line;
                    return;
                }
                tri_indices[3 * k] = (uint32_t)i0; tri_indices[3 * k + 1] = (uint32_t)i1; tri_indices[3 * k + 2] = (uint32_t)i2;
            }
        }
    });
    lap("faces resolved");
    for (size_t t = 0; t < bound_err.size(); t++)
        if (bound_err[t] && (!err_line || bound_err[t] < err_line)) { err_line = bound_err[t]; err = "vertex index out of bounds at line "; }
    if (err_line) throw std::runtime_error(err + std::to_string(err_line));
}

HittablePtr LoadOBJ(const std::string& filename, MaterialPtr material) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<Point3> vertices;
    std::vector<uint32_t> idx;
    const char* env = std::getenv("RT_OBJ_THREADS");
    const int threads = env ? std::atoi(env) : 0;
    ParseOBJ(filename, threads, vertices, idx);
    const size_t nt = idx.size() / 3;
    const int T = (int)std::min<size_t>(threads > 0 ? (size_t)threads : std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(1, nt / 16384));
    HittablePtr root;
    std::chrono::steady_clock::time_point t1;
    if (EagerMeshBVH()) {
        std::vector<HittablePtr> triangles(nt);
        {   // NewTriangle per face (rt/obj_loader.go:92-96), in file order; the objects are independent
            std::vector<std::thread> th;
            auto work = [&](int t) {
                for (size_t k = nt * (size_t)t / (size_t)T, e = nt * (size_t)(t + 1) / (size_t)T; k < e; k++)
                    triangles[k] = NewTriangle(vertices[idx[3 * k]], vertices[idx[3 * k + 1]], vertices[idx[3 * k + 2]], material);
            };
            for (int t = 1; t < T; t++) th.emplace_back(work, t);
            work(0);
            for (auto& x : th) x.join();
        }
        t1 = std::chrono::steady_clock::now();
        root = NewBVHNode(triangles, 0, triangles.size());   // rt/obj_loader.go:109
    } else {
        // rt/obj_loader.go:92-109 creates a Triangle per face and builds the mesh BVH here. The device path needs neither the objects nor the
        // tree — only the faces (as flat arrays: what the flattener emits anyway) and the tree's leaf ORDER, which the library derives on the
        // GPU. Both are deferred (BVHNode::EnsureBuilt, RT_EAGER_BVH=1): 130 of the 160 ms a cold load of the 280 K-triangle mesh took.
        auto soup = std::make_shared<BVHNode::Soup>();
        soup->n = nt; soup->mat = material;
        soup->v0.resize(3 * nt); soup->v1.resize(3 * nt); soup->v2.resize(3 * nt);
        std::vector<std::thread> th;
        auto work = [&](int t) {
            for (size_t k = nt * (size_t)t / (size_t)T, e = nt * (size_t)(t + 1) / (size_t)T; k < e; k++) {
                const Point3 &a = vertices[idx[3 * k]], &b = vertices[idx[3 * k + 1]], &c = vertices[idx[3 * k + 2]];
                soup->v0[3 * k] = a.X; soup->v0[3 * k + 1] = a.Y; soup->v0[3 * k + 2] = a.Z;
                soup->v1[3 * k] = b.X; soup->v1[3 * k + 1] = b.Y; soup->v1[3 * k + 2] = b.Z;
                soup->v2[3 * k] = c.X; soup->v2[3 * k + 1] = c.Y; soup->v2[3 * k + 2] = c.Z;
            }
        };
        for (int t = 1; t < T; t++) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        t1 = std::chrono::steady_clock::now();
        root = NewMeshRootFromSoup(soup, threads);
    }
    if (std::getenv("RT_DEBUG_TIMING"))
        std::fprintf(stderr, "[rt] LoadOBJ %s: parse %.3f s, NewBVHNode %.3f s (%zu vertices, %zu triangles)\n", filename.c_str(),
                     std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count(),
                     vertices.size(), nt);
    return root;
}
HittablePtr LoadOBJWithTransform(const std::string& filename, MaterialPtr material, const Transform* transform) {
    HittablePtr mesh = LoadOBJ(filename, material);
    if (transform) return transform->Apply(mesh);
    return mesh;
}

}  // namespace rt
