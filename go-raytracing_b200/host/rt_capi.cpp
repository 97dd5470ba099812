// rt_capi.cpp — a small C surface over the C++ host mirror so that tests/bench (Python, ctypes) can obtain the
// flattened description of the named scenes and drive the BucketRenderer mirror. Not part of the drop-in
// boundary (that is include/rtx_b200.h); it only exposes host logic that a Go caller has natively.
#include <cstring>

#include <chrono>
#include <thread>

#include "rt.hpp"

using namespace rt;

struct rth_scene {
    Scene scene;
    HittablePtr world;  // HittableList or BVHNode
    std::shared_ptr<FlatScene> flat;
    rtx_scene_desc desc;
    rtx_camera_desc cam;
};

static thread_local std::string g_err;

extern "C" {

const char* rth_last_error() { return g_err.c_str(); }

// name: random | cornell | cornell-glossy | cornell-lucy | hdri-test (main.go:108-152). use_bvh = wrap the world in
// NewBVHNodeFromList as main.go:77 does. width<=0 keeps the scene's own resolution / quality.
rth_scene* rth_scene_named(const char* name, const char* asset_root, uint64_t seed, int32_t use_bvh,
                           int32_t width, double aspect, int32_t spp, int32_t depth) {
    try {
        auto s = new rth_scene();
        s->scene = LoadSceneByName(name, asset_root ? asset_root : ".", seed);
        Camera& c = *s->scene.camera;
        if (width > 0) c.SetResolution(width, aspect);
        if (spp > 0) c.SamplesPerPixel = spp;
        if (depth > 0) c.MaxDepth = depth;
        c.Initialize();
        s->world = use_bvh ? std::static_pointer_cast<Hittable>(NewBVHNodeFromList(s->scene.world)) : std::static_pointer_cast<Hittable>(s->scene.world);
        s->flat = Flatten(s->world, c);
        s->desc = s->flat->Desc();
        c.FillDesc(s->cam);
        return s;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void rth_scene_free(rth_scene* s) { delete s; }
void rth_set_eager_mesh_bvh(int32_t eager) { SetEagerMeshBVH(eager != 0); }   // LoadOBJ builds the reference-order mesh tree at once (default: deferred)
const rtx_scene_desc* rth_scene_desc(rth_scene* s) { return &s->desc; }
const rtx_camera_desc* rth_camera_desc(rth_scene* s) { return &s->cam; }
int32_t rth_image_height(rth_scene* s) { return s->scene.camera->ImageHeight; }

// ImageHeight rule of Camera.Initialize (rt/camera.go:299) for KAT tests.
int32_t rth_image_height_for(int32_t width, double aspect) {
    Camera c;
    c.SetResolution(width, aspect);
    c.Initialize();
    return c.ImageHeight;
}

// Full BucketRenderer run (3 passes, rt/bucket_renderer.go:175-191) of a named scene; pix = RGBA8 4*W*H.
int32_t rth_bucket_render(rth_scene* s, uint64_t seed, uint8_t* pix, int64_t nbytes, double* seconds, const char* save_path) {
    try {
        BucketRenderer r(s->scene.camera, s->world, 32, 1);
        r.seed = seed;
        r.RenderToCompletion();
        if ((int64_t)r.Pix().size() != nbytes) { g_err = "pix size mismatch"; return RTX_ERR_INVALID; }
        std::memcpy(pix, r.Pix().data(), r.Pix().size());
        if (seconds) *seconds = r.GetRenderDurationSeconds();
        if (save_path && *save_path && r.SaveImage(save_path) != 0) { g_err = "SaveImage failed"; return RTX_ERR_INVALID; }
        return RTX_OK;
    } catch (const std::exception& e) {
        g_err = e.what();
        return RTX_ERR_INVALID;
    }
}

// The display-loop use of the renderer (main.go's ebiten game loop calls Update / Draw every frame): ticks Update() without ever
// blocking on a pass and copies the framebuffer each tick, like Draw does. Reports how many ticks returned while a pass was
// still running and how many different finished passes were observed in the framebuffer on the way (the preview arrives first).
int32_t rth_bucket_render_progressive(rth_scene* s, uint64_t seed, uint8_t* pix, int64_t nbytes, int32_t* ticks_while_rendering, int32_t* frames_seen) {
    try {
        BucketRenderer r(s->scene.camera, s->world, 32, 1);
        r.seed = seed;
        int ticks = 0, frames = 0;
        uint64_t last = 0;
        while (!r.IsCompleted()) {
            if (r.Update() != 0) { g_err = r.LastError(); return RTX_ERR_INVALID; }
            if (!r.IsCompleted()) ticks++;
            std::vector<uint8_t> fb = r.CopyFramebuffer();
            uint64_t hsh = 1469598103934665603ull;
            for (size_t i = 0; i < fb.size(); i += 97) hsh = (hsh ^ fb[i]) * 1099511628211ull;
            bool blank = true;
            for (size_t i = 0; i < fb.size() && blank; i += 4) blank = fb[i] == 0 && fb[i + 1] == 0 && fb[i + 2] == 0 && fb[i + 3] == 0;
            if (!blank && hsh != last) { frames++; last = hsh; }
            std::this_thread::sleep_for(std::chrono::microseconds(500));
        }
        if ((int64_t)r.Pix().size() != nbytes) { g_err = "pix size mismatch"; return RTX_ERR_INVALID; }
        std::memcpy(pix, r.Pix().data(), r.Pix().size());
        if (ticks_while_rendering) *ticks_while_rendering = ticks;
        if (frames_seen) *frames_seen = frames;
        return RTX_OK;
    } catch (const std::exception& e) {
        g_err = e.what();
        return RTX_ERR_INVALID;
    }
}

int32_t rth_load_hdr(const char* path, int32_t* w, int32_t* h, double* rgb_out, int64_t capacity) {
    int ww, hh;
    std::vector<double> rgb;
    std::string err;
    if (!LoadHDR(path, ww, hh, rgb, &err)) { g_err = err; return RTX_ERR_INVALID; }
    *w = ww; *h = hh;
    if (rgb_out) {
        if ((int64_t)rgb.size() > capacity) { g_err = "capacity"; return RTX_ERR_INVALID; }
        std::memcpy(rgb_out, rgb.data(), rgb.size() * sizeof(double));
    }
    return RTX_OK;
}

// ParseOBJ for tests and timing: counts always, arrays when the capacities suffice. seconds = wall time of the parse alone.
int32_t rth_parse_obj(const char* path, int32_t threads, int64_t* n_vertices, int64_t* n_triangles, double* vertices_out, int64_t vcap,
                      uint32_t* indices_out, int64_t tcap, double* seconds) {
    try {
        std::vector<Point3> v;
        std::vector<uint32_t> idx;
        const auto t0 = std::chrono::steady_clock::now();
        ParseOBJ(path, threads, v, idx);
        if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *n_vertices = (int64_t)v.size(); *n_triangles = (int64_t)idx.size() / 3;
        if (vertices_out && vcap >= (int64_t)v.size())
            for (size_t i = 0; i < v.size(); i++) { vertices_out[3 * i] = v[i].X; vertices_out[3 * i + 1] = v[i].Y; vertices_out[3 * i + 2] = v[i].Z; }
        if (indices_out && tcap >= (int64_t)idx.size() / 3) std::memcpy(indices_out, idx.data(), idx.size() * sizeof(uint32_t));
        return RTX_OK;
    } catch (const std::exception& e) {
        g_err = e.what();
        return RTX_ERR_INVALID;
    }
}

// drawStatsToFramebuffer on a caller-owned RGBA8 image; text_out (>= 256 bytes, may be NULL) receives the stats line.
int32_t rth_stats_bar(uint8_t* rgba, int32_t w, int32_t h, int32_t spp, int32_t depth, double seconds, int32_t workers, char* text_out) {
    const std::string t = StatsBarText(w, h, spp, depth, seconds, workers);
    DrawStatsBar(rgba, w, h, t);
    if (text_out) std::snprintf(text_out, 256, "%s", t.c_str());
    return RTX_OK;
}

int32_t rth_write_png(const char* path, const uint8_t* rgba, int32_t w, int32_t h) { return WritePNG(path, rgba, w, h); }

}  // extern "C"
