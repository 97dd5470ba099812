// rt_scenes.cpp — the five configured scenes of BASELINE.json and the other scene functions of rt/scenes.go whose vocabulary
// the device path covers (CheckeredSpheres, Simple, PerlinSpheres, Quads, Primitives, GlossyMetalTest, CornellSmoke), mirrored
// from the reference. EarthScene reads a PPM conversion of the reference's earthmap.jpg (no JPEG decoder in this toolchain).
// RandomScene draws from a seeded SplitMix64 stream in the reference's draw order (the reference uses
// Go's auto-seeded global source, rt/utils.go:18, so its geometry differs run to run).
#include <sys/stat.h>

#include "rt.hpp"

namespace rt {

namespace {
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double RandomDouble() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1), 53 bits
    double RandomDoubleRange(double mn, double mx) { return mn + (mx - mn) * RandomDouble(); }
};

void addCornellWalls(HittableListPtr world, MaterialPtr green, MaterialPtr red, MaterialPtr white) {
    world->Add(NewQuad({555, 0, 0}, {0, 555, 0}, {0, 0, 555}, green));
    world->Add(NewQuad({0, 0, 0}, {0, 555, 0}, {0, 0, 555}, red));
    world->Add(NewQuad({0, 0, 0}, {555, 0, 0}, {0, 0, 555}, white));
    world->Add(NewQuad({555, 555, 555}, {-555, 0, 0}, {0, 0, -555}, white));
    world->Add(NewQuad({0, 0, 555}, {555, 0, 0}, {0, 555, 0}, white));
}
}  // namespace

Scene RandomScene(uint64_t seed) {  // rt/scenes.go:30-130 with DefaultSceneConfig (:13-28)
    SplitMix64 rng(seed);
    auto world = NewHittableList();
    auto groundChecker = NewCheckerTextureFromColors(0.32, {0.5, 0.5, 0.5}, {0.9, 0.9, 0.9});
    world->Add(NewPlane({0, 0, -1}, {0, 1, 0}, NewLambertianTexture(groundChecker)));
    const double lambertT = 0.3, metalT = 0.3 + lambertT, dielT = 0.3 + metalT;
    for (int a = -10; a < 10; a++) {
        for (int b = -10; b < 10; b++) {
            double chooseMat = rng.RandomDouble();
            double cx = (double)a + 0.9 * rng.RandomDouble();
            double cz = (double)b + 0.9 * rng.RandomDouble();
            Point3 center{cx, 0.2, cz};
            if (center.Sub({4, 0.2, 0}).Len() > 0.9) {
                if (chooseMat < lambertT) {
                    double r0 = rng.RandomDouble(), r1 = rng.RandomDouble(), g0 = rng.RandomDouble(), g1 = rng.RandomDouble(),
                           b0 = rng.RandomDouble(), b1 = rng.RandomDouble();
                    Color albedo{r0 * r1, g0 * g1, b0 * b1};
                    Point3 center2 = center.Add({0, rng.RandomDoubleRange(0, 0.5), 0});
                    world->Add(NewMovingSphere(center, center2, 0.2, NewLambertian(albedo)));
                } else if (chooseMat < metalT) {
                    double r = 0.5 + rng.RandomDouble() * 0.5, g = 0.5 + rng.RandomDouble() * 0.5, bl = 0.5 + rng.RandomDouble() * 0.5;
                    double fuzz = rng.RandomDouble() * 0.5;
                    world->Add(NewSphere(center, 0.2, NewMetal({r, g, bl}, fuzz)));
                } else if (chooseMat < dielT) {
                    world->Add(NewSphere(center, 0.2, NewDielectric(1.5)));
                }
            }
        }
    }
    world->Add(NewSphere({0, 1, 0}, 1.0, NewDielectric(1.5)));
    world->Add(NewSphere({-4, 1, 0}, 1.0, NewLambertian({0.4, 0.2, 0.1})));
    world->Add(NewSphere({4, 1, 0}, 1.0, NewMetal({0.7, 0.6, 0.5}, 0.0)));
    auto cam = NewCameraBuilder()
                   .SetResolution(1200, 16.0 / 9.0)
                   .SetQuality(500, 50)
                   .SetPosition({13, 2, 3}, {0, 0, 0}, {0, 1, 0})
                   .SetLens(20, 0.6, 10.0)
                   .EnableSkyGradient(true)
                   .Build();
    return {world, cam};
}

Scene HDRITestScene(const std::string& hdrPath) {  // rt/scenes.go:406-458
    auto world = NewHittableList();
    auto glass = NewDielectric(1.5);
    auto mirror = NewMetal({1.0, 1.0, 1.0}, 0.0);
    auto gold = NewMetal({1.0, 0.84, 0.0}, 0.1);
    auto ground = NewLambertianTexture(NewCheckerTextureFromColors(0.5, {0.1, 0.1, 0.1}, {0.9, 0.9, 0.9}));
    world->Add(NewPlane({0, 0, 0}, {0, 1, 0}, ground));
    world->Add(NewSphere({0, 1, 0}, 1.0, glass));
    world->Add(NewSphere({-2.5, 1, 0}, 1.0, mirror));
    world->Add(NewSphere({2.5, 1, 0}, 1.0, gold));
    world->Add(NewSphere({-1.2, 0.4, 2}, 0.4, glass));
    world->Add(NewSphere({1.2, 0.4, 2}, 0.4, glass));
    auto cam = NewCameraBuilder()
                   .SetResolution(800, 16.0 / 9.0)
                   .SetQuality(200, 20)
                   .SetPosition({0, 2.5, 8}, {0, 1, 0}, {0, 1, 0})
                   .SetLens(40, 0, 10)
                   .SetEnvironmentMap(hdrPath)
                   .SetEnvironmentRotation(0)
                   .SetPhantomHDRI(true)
                   .Build();
    return {world, cam};
}

Scene CornellBoxScene() {  // rt/scenes.go:463-562
    auto world = NewHittableList();
    auto white = NewLambertian({0.73, 0.73, 0.73});
    auto red = NewLambertian({0.65, 0.05, 0.05});
    auto green = NewLambertian({0.12, 0.45, 0.15});
    auto light = NewDiffuseLight(NewSolidColor({3, 3, 3}));
    auto areaLight = NewQuad({213, 554, 227}, {130, 0, 0}, {0, 0, 105}, light);
    world->Add(areaLight);
    addCornellWalls(world, green, red, white);
    auto box1 = Box({0, 0, 0}, {165, 330, 165}, white);
    world->Add(NewTransform().SetScale({1, 1, 1}).SetRotationY(15).SetPosition({265, 0, 295}).Apply(box1));
    auto box2 = Box({0, 0, 0}, {165, 165, 165}, white);
    world->Add(NewTransform().SetScale({1, 1, 1}).SetRotationY(-18).SetPosition({130, 0, 65}).Apply(box2));
    auto fogBoundary = Box({0, 0, 0}, {555, 555, 555}, white);
    world->Add(NewVolumeFromColor(fogBoundary, 0.001, {1, 1, 1}));
    auto cam = NewCameraBuilder()
                   .SetResolution(600, 1.0)
                   .SetQuality(500, 5)
                   .SetPosition({278, 278, -800}, {278, 278, 0}, {0, 1, 0})
                   .SetLens(40, 0, 10)
                   .SetBackground({0, 0, 0})
                   .AddLight(areaLight)
                   .Build();
    return {world, cam};
}

Scene CornellBoxGlossy() {  // rt/scenes.go:606-711
    auto world = NewHittableList();
    auto white = NewLambertian({0.73, 0.73, 0.73});
    auto red = NewLambertian({0.65, 0.05, 0.05});
    auto green = NewLambertian({0.12, 0.45, 0.15});
    auto goldShiny = NewMetal({1.0, 0.84, 0.0}, 0.05);
    auto goldBrushed = NewMetal({1.0, 0.84, 0.0}, 0.15);
    auto silverRough = NewMetal({0.95, 0.95, 0.98}, 0.25);
    auto glass = NewDielectric(1.5);
    auto light = NewDiffuseLightColor({15, 15, 15});
    addCornellWalls(world, green, red, white);
    auto areaLight = NewQuad({213, 554, 227}, {130, 0, 0}, {0, 0, 105}, light);
    world->Add(areaLight);
    world->Add(NewSphere({150, 100, 400}, 100, goldShiny));
    world->Add(NewSphere({278, 100, 400}, 100, goldBrushed));
    world->Add(NewSphere({410, 100, 400}, 100, silverRough));
    world->Add(NewSphere({278, 130, 180}, 130, glass));
    auto cam = NewCameraBuilder()
                   .SetResolution(600, 1.0)
                   .SetQuality(200, 5)
                   .SetPosition({278, 278, -800}, {278, 200, 200}, {0, 1, 0})
                   .SetLens(40, 0, 10)
                   .SetBackground({0, 0, 0})
                   .AddLight(areaLight)
                   .Build();
    return {world, cam};
}

Scene CornellBoxLucy(const std::string& objPath) {  // rt/scenes.go:714-817
    auto world = NewHittableList();
    auto white = NewLambertian({0.73, 0.73, 0.73});
    auto red = NewLambertian({0.65, 0.05, 0.05});
    auto green = NewLambertian({0.12, 0.45, 0.15});
    auto light = NewDiffuseLight(NewSolidColor({15, 15, 15}));
    auto areaLight = NewQuad({213, 554, 227}, {130, 0, 0}, {0, 0, 105}, light);
    world->Add(areaLight);
    addCornellWalls(world, green, red, white);
    auto lucyMat = NewLambertian({0.9, 0.9, 0.9});
    const double scale = 0.15;
    HittablePtr lucyMesh = LoadOBJ(objPath, lucyMat);  // throws, where the Go code panics (rt/scenes.go:771-773)
    struct Inst { Vec3 pos; double rot; };
    const Inst positions[] = {{{150, 0, 150}, 45}, {{400, 0, 150}, 315}, {{150, 0, 400}, 135}, {{400, 0, 400}, 225}, {{278, 0, 278}, 0},
                              {{100, 0, 278}, 90}, {{450, 0, 278}, 270}, {{278, 0, 100}, 180}, {{278, 0, 450}, 0}, {{200, 0, 350}, 60}};
    for (const Inst& inst : positions)
        world->Add(NewTransform().SetScale({scale, scale, scale}).SetRotationY(inst.rot).SetPosition(inst.pos).Apply(lucyMesh));
    auto cam = NewCameraBuilder()
                   .SetResolution(600, 1.0)
                   .SetQuality(50, 5)
                   .SetPosition({278, 278, -800}, {278, 278, 0}, {0, 1, 0})
                   .SetLens(40, 0, 10)
                   .SetBackground({0, 0, 0})
                   .AddLight(areaLight)
                   .Build();
    return {world, cam};
}

// ---- the remaining scene functions the device vocabulary covers ------------------------------------------------
Scene EarthScene(const std::string& imagePath) {  // rt/scenes.go:210-240
    auto world = NewHittableList();
    auto img = std::make_shared<ImageLoader>();
    if (!img->Load(imagePath)) throw std::runtime_error("EarthScene: cannot load " + imagePath + " (binary PPM; tools/make_assets.py writes it from the reference's earthmap.jpg)");
    world->Add(NewSphere({0, 0, 0}, 2, NewLambertianTexture(NewImageTextureFromImage(img))));
    auto cam = NewCameraBuilder().SetResolution(800, 16.0 / 9.0).SetQuality(100, 50).SetPosition({0, 0, 12}, {0, 0, 0}, {0, 1, 0}).SetLens(20, 0, 10)
                   .EnableSkyGradient(true).Build();
    return {world, cam};
}

Scene PerlinSpheresScene(uint64_t seed) {  // rt/scenes.go:242-272 (the Perlin tables come from a seeded stream, see NewNoiseTexture)
    auto world = NewHittableList();
    auto perl = NewLambertianTexture(NewNoiseTexture(4.0, seed));
    world->Add(NewSphere({0, 2, 0}, 2, perl));
    world->Add(NewPlane({0, 0, -1}, {0, 1, 0}, perl));
    auto cam = NewCameraBuilder().SetResolution(600, 16.0 / 9.0).SetQuality(100, 50).SetPosition({13, 2, -10}, {0, 1.5, 0}, {0, 1, 0}).SetLens(20, 0, 10)
                   .EnableSkyGradient(true).Build();
    return {world, cam};
}

Scene PrimitivesScene() {  // rt/scenes.go:313-404: plane, Circle, Pyramid, glass sphere, Box, area light, mirror sphere
    auto world = NewHittableList();
    auto red = NewLambertian({0.8, 0.1, 0.1});
    auto green = NewLambertian({0.1, 0.8, 0.1});
    auto blue = NewLambertian({0.1, 0.1, 0.8});
    auto metal = NewMetal({1.0, 1.0, 1.0}, 0);
    auto lightMat = NewDiffuseLight(NewSolidColor({2, 2, 2}));
    auto checker = NewLambertianTexture(NewCheckerTextureFromColors(1.0, {0.0, 0.0, 0.0}, {0.9, 0.9, 0.9}));
    world->Add(NewPlane({0, -1, 0}, {0, 1, 0}, checker));
    world->Add(NewCircle({-5, 0, 0}, {0, 1, 0}, 0.9, red));
    world->Add(Pyramid({-2.5, -1, 0}, 1.4, 1.8, green));
    world->Add(NewSphere({0, 0.6, 0}, 0.8, NewDielectric(1.5)));
    const double cubeX = 2.5, cubeSize = 1.0;
    world->Add(Box({cubeX - cubeSize / 2, -1, -cubeSize / 2}, {cubeX + cubeSize / 2, -1 + cubeSize, cubeSize / 2}, blue));
    auto areaLight = NewQuad({-2, 5, -2}, {4, 0, 0}, {0, 0, 4}, lightMat);
    world->Add(areaLight);
    world->Add(NewSphere({5, 0.6, 0}, 0.8, metal));
    auto cam = NewCameraBuilder().SetResolution(800, 16.0 / 9.0).SetQuality(300, 25).SetPosition({0, 2, 10}, {0, 0, 0}, {0, 1, 0}).SetLens(45, 0, 10)
                   .SetBackground({0, 0, 0}).EnableSkyGradient(true).AddLight(areaLight).Build();
    return {world, cam};
}

Scene CheckeredSpheresScene() {  // rt/scenes.go:132-170
    auto world = NewHittableList();
    auto checker = NewLambertianTexture(NewCheckerTextureFromColors(0.32, {0.2, 0.3, 0.1}, {0.9, 0.9, 0.9}));
    world->Add(NewSphere({0, -10, 0}, 10, checker));
    world->Add(NewSphere({0, 10, 0}, 10, checker));
    auto cam = NewCameraBuilder().SetResolution(600, 16.0 / 9.0).SetQuality(100, 50).SetPosition({13, 2, 3}, {0, 0, 0}, {0, 1, 0}).SetLens(20, 0, 10)
                   .EnableSkyGradient(true).Build();
    return {world, cam};
}

Scene SimpleScene() {  // rt/scenes.go:172-209 (a hollow glass sphere: the inner bubble has ior 1/1.5)
    auto world = NewHittableList();
    world->Add(NewPlane({0, -0.5, -1}, {0, 1, 0}, NewLambertian({0.8, 0.8, 0.0})));
    world->Add(NewSphere({0, 0, -1}, 0.5, NewLambertian({0.1, 0.2, 0.5})));
    world->Add(NewSphere({-1, 0, -1}, 0.5, NewDielectric(1.5)));
    world->Add(NewSphere({-1, 0, -1}, 0.4, NewDielectric(1.0 / 1.5)));
    world->Add(NewSphere({1, 0, -1}, 0.5, NewMetal({0.8, 0.6, 0.2}, 0.0)));
    auto cam = NewCameraBuilder().SetResolution(400, 16.0 / 9.0).SetQuality(100, 50).SetPosition({0, 0, 2}, {0, 0, -1}, {0, 1, 0}).SetLens(90, 0, 10)
                   .EnableSkyGradient(true).Build();
    return {world, cam};
}

Scene QuadsScene() {  // rt/scenes.go:274-311
    auto world = NewHittableList();
    world->Add(NewQuad({-3, -2, 5}, {0, 0, -4}, {0, 4, 0}, NewLambertian({1.0, 0.2, 0.2})));
    world->Add(NewQuad({-2, -2, 0}, {4, 0, 0}, {0, 4, 0}, NewLambertian({0.2, 1.0, 0.2})));
    world->Add(NewQuad({3, -2, 1}, {0, 0, 4}, {0, 4, 0}, NewLambertian({0.2, 0.2, 1.0})));
    world->Add(NewQuad({-2, 3, 1}, {4, 0, 0}, {0, 0, 4}, NewLambertian({1.0, 0.5, 0.0})));
    world->Add(NewQuad({-2, -3, 5}, {4, 0, 0}, {0, 0, -4}, NewLambertian({0.2, 0.8, 0.8})));
    auto cam = NewCameraBuilder().SetResolution(400, 1.0).SetQuality(100, 50).SetPosition({0, 0, 9}, {0, 0, 0}, {0, 1, 0}).SetLens(80, 0, 10)
                   .EnableSkyGradient(true).Build();
    return {world, cam};
}

Scene GlossyMetalTest() {  // rt/scenes.go:564-604
    auto world = NewHittableList();
    world->Add(NewPlane({0, 0, 0}, {0, 1, 0}, NewLambertian({0.5, 0.5, 0.5})));
    world->Add(NewSphere({-2.5, 1, 0}, 1.0, NewMetal({0.8, 0.6, 0.2}, 0.0)));
    world->Add(NewSphere({0, 1, 0}, 1.0, NewMetal({0.8, 0.6, 0.2}, 0.2)));
    world->Add(NewSphere({2.5, 1, 0}, 1.0, NewMetal({0.8, 0.6, 0.2}, 0.5)));
    auto areaLight = NewQuad({-2, 5, -2}, {4, 0, 0}, {0, 0, 4}, NewDiffuseLightColor({4, 4, 4}));
    world->Add(areaLight);
    auto cam = NewCameraBuilder().SetResolution(640, 16.0 / 9.0).SetQuality(100, 10).SetPosition({0, 2, 10}, {0, 1, 0}, {0, 1, 0}).SetLens(40, 0, 10)
                   .SetBackground({0, 0, 0}).AddLight(areaLight).Build();
    return {world, cam};
}

Scene CornellSmoke() {  // rt/scenes.go:820-925: two rotated boxes of participating medium (black and white smoke)
    auto world = NewHittableList();
    auto white = NewLambertian({0.73, 0.73, 0.73});
    auto red = NewLambertian({0.65, 0.05, 0.05});
    auto green = NewLambertian({0.12, 0.45, 0.15});
    auto areaLight = NewQuad({113, 554, 127}, {330, 0, 0}, {0, 0, 305}, NewDiffuseLight(NewSolidColor({3, 3, 3})));
    world->Add(areaLight);
    addCornellWalls(world, green, red, white);
    auto box1 = NewTransform().SetRotationY(15).SetPosition({265, 0, 295}).Apply(Box({0, 0, 0}, {165, 330, 165}, white));
    world->Add(NewVolumeFromColor(box1, 0.01, {0, 0, 0}));
    auto box2 = NewTransform().SetRotationY(-18).SetPosition({130, 0, 65}).Apply(Box({0, 0, 0}, {165, 165, 165}, white));
    world->Add(NewVolumeFromColor(box2, 0.01, {1, 1, 1}));
    auto cam = NewCameraBuilder().SetResolution(600, 1.0).SetQuality(150, 5).SetPosition({278, 278, -800}, {278, 278, 0}, {0, 1, 0}).SetLens(40, 0, 10)
                   .SetBackground({0, 0, 0}).AddLight(areaLight).Build();
    return {world, cam};
}

static bool bigFile(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0 && st.st_size > 1024;  // the shipped lucy_low.obj is a 133-byte Git-LFS pointer
}

Scene LoadSceneByName(const std::string& nameIn, const std::string& assetRoot, uint64_t seed) {  // main.go:108-152
    std::string name;
    for (char c : nameIn) name.push_back((char)std::tolower((unsigned char)c));
    std::string root = assetRoot.empty() ? std::string(".") : assetRoot;
    if (name == "random" || name == "randomscene") return RandomScene(seed);
    if (name == "cornell" || name == "cornell-box") return CornellBoxScene();
    if (name == "cornell-glossy") return CornellBoxGlossy();
    if (name == "checkered" || name == "checker" || name == "checkered-spheres") return CheckeredSpheresScene();
    if (name == "simple" || name == "simple-scene") return SimpleScene();
    if (name == "quads" || name == "quads-scene") return QuadsScene();
    if (name == "cornell-smoke" || name == "cornell-fog") return CornellSmoke();
    if (name == "glossy-metal" || name == "glossy-metal-test") return GlossyMetalTest();
    if (name == "perlin" || name == "perlin-spheres") return PerlinSpheresScene(seed);
    if (name == "primitives" || name == "primitives-scene") return PrimitivesScene();
    if (name == "earth" || name == "earth-scene") {
        std::string real = root + "/assets/images/earthmap.ppm", standin = root + "/assets/images/synthetic_earth.ppm";
        return EarthScene(bigFile(real) ? real : standin);
    }
    if (name == "cornell-lucy") {
        std::string real = root + "/assets/models/lucy_low.obj", standin = root + "/assets/models/lucy_standin.obj";
        return CornellBoxLucy(bigFile(real) ? real : standin);
    }
    if (name == "hdri" || name == "hdri-test" || name == "hdr") {
        std::string real = root + "/assets/hdri/abandoned_hall_01_1k.hdr", standin = root + "/assets/hdri/synthetic_hall_1k.hdr";
        return HDRITestScene(bigFile(real) ? real : standin);
    }
    throw std::runtime_error("unknown scene: " + nameIn);
}

}  // namespace rt
