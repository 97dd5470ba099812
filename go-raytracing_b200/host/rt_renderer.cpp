// rt_renderer.cpp — BucketRenderer / ProgressiveRenderer mirrors (rt/bucket_renderer.go, rt/renderer.go).
//
// The constructor, the 3-pass state machine of Update() — non-blocking, a worker thread per pass like the reference's
// goroutines — and SaveImage/IsCompleted/GetRenderDuration keep the reference's behaviour; the body of renderPass (rt/bucket_renderer.go:170-214: worker goroutines, 32x32 tiles,
// GetRay/RayColor per sample, gamma + RGBA8 pack) is ONE call into the CUDA library per pass.
// bucketSize / numWorkers are accepted and ignored (the device schedules its own work).
// The stats bar the reference burns into the saved image (drawStatsToFramebuffer, rt/bucket_renderer.go:375-407) is
// DrawStatsToFramebuffer() below; the live ebiten overlay of Draw() (:303-373) belongs to the display and stays in Go.
#include <chrono>
#include <cstdio>
#include <cstring>

#include "rt.hpp"

namespace rt {

static void check(rtx_ctx* ctx, int32_t rc, const char* what) {
    if (rc != RTX_OK) throw std::runtime_error(std::string(what) + ": " + rtx_last_error(ctx));
}

BucketRenderer::BucketRenderer(CameraPtr camera, HittablePtr world, int /*bucketSize*/, int numWorkers, int deviceId) : camera_(camera), numWorkers_(numWorkers) {
    flat_ = Flatten(world, *camera);
    int32_t rc = rtx_create(deviceId, &ctx_);
    if (rc != RTX_OK) throw std::runtime_error(std::string("rtx_create: ") + rtx_last_error(nullptr));
    rtx_scene_desc sd = flat_->Desc();
    check(ctx_, rtx_scene_upload(ctx_, &sd), "rtx_scene_upload");
    rtx_camera_desc cd;
    camera->FillDesc(cd);
    check(ctx_, rtx_camera_set(ctx_, &cd), "rtx_camera_set");
    int32_t w, h;
    check(ctx_, rtx_image_size(ctx_, &w, &h), "rtx_image_size");
    w_ = w; h_ = h;
    pix_.assign((size_t)4 * w * h, 0);  // image.NewRGBA (rt/bucket_renderer.go:55)
    back_.assign((size_t)4 * w * h, 0);
}
BucketRenderer::~BucketRenderer() {
    join();
    if (ctx_) rtx_destroy(ctx_);
}
void BucketRenderer::join() {
    if (worker_.joinable()) worker_.join();
}

void BucketRenderer::renderPass(int pass) {  // rt/bucket_renderer.go:170-214: the pass schedule; the worker pool is ONE device call
    int spp, depth;
    switch (pass) {
        case 0: spp = 1; depth = 3; break;
        case 1: spp = std::max(1, camera_->SamplesPerPixel / 4); depth = std::max(3, camera_->MaxDepth / 2); break;
        default: spp = camera_->SamplesPerPixel; depth = camera_->MaxDepth; break;
    }
    try {
        check(ctx_, rtx_accum_clear(ctx_), "rtx_accum_clear");  // each pass overwrites the framebuffer (:291-300)
        check(ctx_, rtx_render_pass(ctx_, spp, depth, camera_->MaxDepth, seed + (uint64_t)pass, 0), "rtx_render_pass");
        check(ctx_, rtx_resolve_rgba8(ctx_, spp, back_.data(), (int64_t)back_.size()), "rtx_resolve_rgba8");
        std::lock_guard<std::mutex> lk(mu_);
        pix_.swap(back_);   // the finished pass becomes visible to Draw / CopyFramebuffer at once
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(mu_);
        err_ = e.what();
    }
    passComplete_.store(true);
}

int BucketRenderer::Update() {  // rt/bucket_renderer.go:127-164
    if (completed_) return err_.empty() ? 0 : -1;
    if (!renderStarted_) {   // first tick: start pass 0 in the background (go r.renderMultiPass())
        renderStarted_ = true;
        renderStart_ = std::chrono::steady_clock::now();
        passComplete_.store(false);
        worker_ = std::thread(&BucketRenderer::renderPass, this, currentPass_);
        return 0;
    }
    if (passComplete_.load() && currentPass_ < totalPasses_) {
        join();
        passComplete_.store(false);
        currentPass_++;
        bool failed;
        { std::lock_guard<std::mutex> lk(mu_); failed = !err_.empty(); }
        if (currentPass_ < totalPasses_ && !failed) {
            worker_ = std::thread(&BucketRenderer::renderPass, this, currentPass_);
        } else {
            completed_ = true;
            duration_s_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - renderStart_).count();
            if (!failed) {   // rt/bucket_renderer.go:151-155: the bar is burnt in, then the image is saved
                if (burnStats) DrawStatsToFramebuffer();
                if (!autoSave.empty()) SaveImage(autoSave);
            }
            return failed ? -1 : 0;
        }
    }
    return 0;
}
void BucketRenderer::RenderToCompletion() {
    while (!completed_) {
        Update();
        if (!completed_) std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    if (!err_.empty()) throw std::runtime_error(err_);
}
std::vector<uint8_t> BucketRenderer::CopyFramebuffer() const {
    std::lock_guard<std::mutex> lk(mu_);
    return pix_;
}

std::shared_ptr<BucketRenderer> NewBucketRenderer(CameraPtr camera, HittablePtr world, int bucketSize, int numWorkers) {
    return std::make_shared<BucketRenderer>(camera, world, bucketSize, numWorkers);
}
std::shared_ptr<BucketRenderer> NewProgressiveRenderer(CameraPtr camera, HittablePtr world) {
    return std::make_shared<BucketRenderer>(camera, world, 0, 0);
}

// ---- stats bar (rt/bucket_renderer.go:375-407, rt/utils.go:50-61) -------------------------------------------
std::string FormatDuration(double seconds) {
    // time.Duration is whole nanoseconds; Hours()/Minutes()/Seconds() are float divisions truncated by int()
    const long long total = (long long)seconds;
    const int hours = (int)(total / 3600), minutes = (int)((total / 60) % 60), secs = (int)(total % 60);
    char buf[64];
    if (hours > 0) std::snprintf(buf, sizeof buf, "%dh %dm %ds", hours, minutes, secs);
    else if (minutes > 0) std::snprintf(buf, sizeof buf, "%dm %ds", minutes, secs);
    else std::snprintf(buf, sizeof buf, "%.2fs", seconds);
    return buf;
}
std::string StatsBarText(int width, int height, int spp, int depth, double seconds, int numWorkers) {
    char buf[256];
    std::snprintf(buf, sizeof buf, "%dx%d | SPP:%d | Depth:%d | 100.0%% | %s | Workers: %d", width, height, spp, depth,
                  FormatDuration(seconds).c_str(), numWorkers);
    return buf;
}
// basicfont.Face7x13 (the X11 misc-fixed 7x13 face, public domain): one byte per row, bit 7 = leftmost pixel, 13 rows per
// 7-pixel-wide cell. Only the characters the stats line can contain; the rows of all but '7', '9' and 'm' are pinned by the bar of
// the reference's committed image.png (tests/golden/image_png_stats_bar.json), those three follow the font's BDF source.
struct Glyph7x13 { char c; uint8_t rows[13]; };
static const Glyph7x13 kGlyphs[] = {
    {'%', {0x00, 0x00, 0x44, 0xA4, 0x48, 0x10, 0x10, 0x20, 0x48, 0x94, 0x88, 0x00, 0x00}},
    {'.', {0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x10, 0x38, 0x10, 0x00}},
    {'0', {0x00, 0x00, 0x30, 0x48, 0x84, 0x84, 0x84, 0x84, 0x84, 0x48, 0x30, 0x00, 0x00}},
    {'1', {0x00, 0x00, 0x10, 0x30, 0x50, 0x10, 0x10, 0x10, 0x10, 0x10, 0x7C, 0x00, 0x00}},
    {'2', {0x00, 0x00, 0x78, 0x84, 0x84, 0x04, 0x08, 0x30, 0x40, 0x80, 0xFC, 0x00, 0x00}},
    {'3', {0x00, 0x00, 0xFC, 0x04, 0x08, 0x10, 0x38, 0x04, 0x04, 0x84, 0x78, 0x00, 0x00}},
    {'4', {0x00, 0x00, 0x08, 0x18, 0x28, 0x48, 0x88, 0x88, 0xFC, 0x08, 0x08, 0x00, 0x00}},
    {'5', {0x00, 0x00, 0xFC, 0x80, 0x80, 0xB8, 0xC4, 0x04, 0x04, 0x84, 0x78, 0x00, 0x00}},
    {'6', {0x00, 0x00, 0x38, 0x40, 0x80, 0x80, 0xB8, 0xC4, 0x84, 0x84, 0x78, 0x00, 0x00}},
    {'7', {0x00, 0x00, 0xFC, 0x04, 0x08, 0x08, 0x10, 0x10, 0x20, 0x20, 0x20, 0x00, 0x00}},
    {'8', {0x00, 0x00, 0x78, 0x84, 0x84, 0x84, 0x78, 0x84, 0x84, 0x84, 0x78, 0x00, 0x00}},
    {'9', {0x00, 0x00, 0x78, 0x84, 0x84, 0x8C, 0x74, 0x04, 0x04, 0x08, 0x70, 0x00, 0x00}},
    {':', {0x00, 0x00, 0x00, 0x00, 0x10, 0x38, 0x10, 0x00, 0x00, 0x10, 0x38, 0x10, 0x00}},
    {'D', {0x00, 0x00, 0xF8, 0x44, 0x44, 0x44, 0x44, 0x44, 0x44, 0x44, 0xF8, 0x00, 0x00}},
    {'P', {0x00, 0x00, 0xF8, 0x84, 0x84, 0x84, 0xF8, 0x80, 0x80, 0x80, 0x80, 0x00, 0x00}},
    {'S', {0x00, 0x00, 0x78, 0x84, 0x80, 0x80, 0x78, 0x04, 0x04, 0x84, 0x78, 0x00, 0x00}},
    {'W', {0x00, 0x00, 0x84, 0x84, 0x84, 0x84, 0xB4, 0xB4, 0xCC, 0xCC, 0x84, 0x00, 0x00}},
    {'e', {0x00, 0x00, 0x00, 0x00, 0x00, 0x78, 0x84, 0xFC, 0x80, 0x84, 0x78, 0x00, 0x00}},
    {'h', {0x00, 0x00, 0x80, 0x80, 0x80, 0xB8, 0xC4, 0x84, 0x84, 0x84, 0x84, 0x00, 0x00}},
    {'k', {0x00, 0x00, 0x80, 0x80, 0x80, 0x88, 0x90, 0xE0, 0x90, 0x88, 0x84, 0x00, 0x00}},
    {'m', {0x00, 0x00, 0x00, 0x00, 0x00, 0xEC, 0x92, 0x92, 0x92, 0x92, 0x82, 0x00, 0x00}},
    {'o', {0x00, 0x00, 0x00, 0x00, 0x00, 0x78, 0x84, 0x84, 0x84, 0x84, 0x78, 0x00, 0x00}},
    {'p', {0x00, 0x00, 0x00, 0x00, 0x00, 0xB8, 0xC4, 0x84, 0xC4, 0xB8, 0x80, 0x80, 0x80}},
    {'r', {0x00, 0x00, 0x00, 0x00, 0x00, 0xB8, 0x44, 0x40, 0x40, 0x40, 0x40, 0x00, 0x00}},
    {'s', {0x00, 0x00, 0x00, 0x00, 0x00, 0x78, 0x84, 0x60, 0x18, 0x84, 0x78, 0x00, 0x00}},
    {'t', {0x00, 0x00, 0x00, 0x40, 0x40, 0xF0, 0x40, 0x40, 0x40, 0x44, 0x38, 0x00, 0x00}},
    {'x', {0x00, 0x00, 0x00, 0x00, 0x00, 0x84, 0x48, 0x30, 0x30, 0x48, 0x84, 0x00, 0x00}},
    {'|', {0x00, 0x00, 0x10, 0x10, 0x10, 0x10, 0x10, 0x10, 0x10, 0x10, 0x10, 0x00, 0x00}},
};
void DrawStatsBar(uint8_t* pix, int w, int h, const std::string& text) {
    const int barHeight = 30, barY = h - barHeight;
    auto set = [&](int x, int y, uint8_t v) {   // image.RGBA.Set ignores points outside the image
        if (x < 0 || y < 0 || x >= w || y >= h) return;
        uint8_t* p = pix + 4 * ((size_t)y * w + x);
        p[0] = p[1] = p[2] = v; p[3] = 255;
    };
    for (int y = barY; y < h; y++)
        for (int x = 0; x < w; x++) set(x, y, 0);
    // text.Draw at (15, barY + 10): the origin is the top-left corner of the first 7x13 cell, white, no anti-aliasing
    for (size_t i = 0; i < text.size(); i++) {
        const Glyph7x13* g = nullptr;
        for (const Glyph7x13& k : kGlyphs) if (k.c == text[i]) { g = &k; break; }
        if (!g) continue;   // ' ' and anything the format string cannot produce
        for (int ry = 0; ry < 13; ry++)
            for (int rx = 0; rx < 7; rx++)
                if (g->rows[ry] & (0x80 >> rx)) set(15 + 7 * (int)i + rx, barY + 10 + ry, 255);
    }
}
void BucketRenderer::DrawStatsToFramebuffer() {
    std::lock_guard<std::mutex> lk(mu_);
    DrawStatsBar(pix_.data(), w_, h_, StatsBarText(w_, h_, camera_->SamplesPerPixel, camera_->MaxDepth, duration_s_, numWorkers_));
}

// ---- PNG (stored deflate blocks; rt/bucket_renderer.go:417-438 uses image/png) ------------------------------
static uint32_t crcTable[256];
static void crcInit() {
    static bool done = false;
    if (done) return;
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crcTable[n] = c;
    }
    done = true;
}
static uint32_t crcUpdate(uint32_t c, const uint8_t* p, size_t n) {
    for (size_t i = 0; i < n; i++) c = crcTable[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c;
}
static void chunk(FILE* f, const char* type, const std::vector<uint8_t>& data) {
    uint8_t len[4] = {(uint8_t)(data.size() >> 24), (uint8_t)(data.size() >> 16), (uint8_t)(data.size() >> 8), (uint8_t)data.size()};
    std::fwrite(len, 1, 4, f);
    std::fwrite(type, 1, 4, f);
    if (!data.empty()) std::fwrite(data.data(), 1, data.size(), f);
    uint32_t c = crcUpdate(0xFFFFFFFFu, (const uint8_t*)type, 4);
    c = crcUpdate(c, data.data(), data.size()) ^ 0xFFFFFFFFu;
    uint8_t cb[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
    std::fwrite(cb, 1, 4, f);
}
int WritePNG(const std::string& path, const uint8_t* rgba, int w, int h) {
    crcInit();
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return -1;
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::fwrite(sig, 1, 8, f);
    std::vector<uint8_t> ihdr = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                                 (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 6, 0, 0, 0};
    chunk(f, "IHDR", ihdr);
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * (4 * w + 1));
    for (int y = 0; y < h; y++) {
        raw.push_back(0);
        raw.insert(raw.end(), rgba + (size_t)y * 4 * w, rgba + (size_t)(y + 1) * 4 * w);
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (uint8_t v : raw) { a = (a + v) % 65521; b = (b + a) % 65521; }
    size_t pos = 0;
    while (pos < raw.size()) {
        size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)n); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)~n); z.push_back((uint8_t)(~n >> 8));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        pos += n;
    }
    uint32_t adler = (b << 16) | a;
    z.push_back((uint8_t)(adler >> 24)); z.push_back((uint8_t)(adler >> 16)); z.push_back((uint8_t)(adler >> 8)); z.push_back((uint8_t)adler);
    chunk(f, "IDAT", z);
    chunk(f, "IEND", {});
    std::fclose(f);
    return 0;
}
int BucketRenderer::SaveImage(const std::string& filename) const {
    if (filename.size() > 4 && filename.substr(filename.size() - 4) == ".ppm") {
        FILE* f = std::fopen(filename.c_str(), "wb");
        if (!f) return -1;
        std::fprintf(f, "P6\n%d %d\n255\n", w_, h_);
        for (size_t i = 0; i < (size_t)w_ * h_; i++) std::fwrite(&pix_[4 * i], 1, 3, f);
        std::fclose(f);
        return 0;
    }
    return WritePNG(filename, pix_.data(), w_, h_);
}

}  // namespace rt
