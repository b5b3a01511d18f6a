// Tests of the C++ host API (include/cornelis/*.hpp), written after the reference's own Catch2 tests
// (reference tests/test_Camera.cpp, test_Math.cpp, test_Tiles.cpp, test_FrameBuffer.cpp, test_Color.cpp,
// test_SceneDescription.cpp) with a tiny local CHECK macro.  `cpu` runs what needs no GPU; `gpu` additionally drives
// RenderSession::render() end to end through the C-ABI.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include <cornelis/Camera.hpp>
#include <cornelis/Color.hpp>
#include <cornelis/FrameBuffer.hpp>
#include <cornelis/Math.hpp>
#include <cornelis/Render.hpp>
#include <cornelis/SceneDescription.hpp>
#include <cornelis/Tiles.hpp>

using namespace cornelis;

static int failures = 0;
#define CHECK(cond)                                                                                                    \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                              \
            failures++;                                                                                                \
        }                                                                                                              \
    } while (0)
#define CHECK_THROWS(expr)                                                                                             \
    do {                                                                                                               \
        bool threw = false;                                                                                            \
        try {                                                                                                          \
            (void)(expr);                                                                                              \
        } catch (std::exception const &) {                                                                             \
            threw = true;                                                                                              \
        }                                                                                                              \
        if (!threw) {                                                                                                  \
            std::printf("FAILED %s:%d: %s did not throw\n", __FILE__, __LINE__, #expr);                                \
            failures++;                                                                                                \
        }                                                                                                              \
    } while (0)
static bool near(double a, double b, double tol) { return std::fabs(a - b) <= tol; }

static void testCamera() { // reference tests/test_Camera.cpp
    CHECK_THROWS(horizontalFov35mm(0.0f));
    CHECK_THROWS(horizontalFov35mm(-1.0f));
    CHECK(near(horizontalFov35mm(50), 0.691111, 0.001));
    CHECK(near(horizontalFov35mm(75), 0.47109, 0.001));
    PerspectiveCamera cam;
    auto ray = cam(0.5f, 0.5f);
    CHECK(ray.eye() == V3(0));
    CHECK(ray.dir() == V3(0, 0, 1));
    cam = PerspectiveCamera::lookAt(V3(0), V3(1, 0, 0), 1.0f, 1.0f);
    ray = cam(0.0f, 0.5f);
    CHECK(ray.eye() == V3(0));
    V3 expect(1.0f, 0, 0.4794255386f);
    expect.normalize();
    CHECK(ray.dir() == expect);
    cam = PerspectiveCamera::lookAt(V3(0, 0, 2), V3(0, 0, 0), 1.0f, 1.0f);
    ray = cam(0.5f, 0.5f);
    CHECK(ray.eye() == V3(0, 0, 2));
    CHECK(ray.dir() == V3(0, 0, -1));
}

static void testMath() { // reference tests/test_Math.cpp:105-143
    float3 a{1.0f, 2.0f, 3.0f}, b{-1.0f, 2.0f, -2.0f};
    CHECK(dot(a, b) == -1.0f + 4.0f - 6.0f);
    CHECK(mag2(a) == 14.0f);
    CHECK(rayT(float3{-1, 0, 1}, float3{1, 0, 1}, 1.0f) == float3(0, 0, 2));
    CHECK(near(mag2(normalize(float3{2, 2, 1})), 1.0, 1e-3));
    CHECK(cross(float3{1, 0, 0}, float3{0, 1, 0}) == float3(0, 0, 1));
    CHECK(cross(float3{0, 1, 0}, float3{1, 0, 0}) == float3(0, 0, -1));
    CHECK(cross(float3{1, 1, 0}, float3{0, 1, 1}) == float3(1, -1, 1));
    CHECK(normalize(float3{1e-6f, 0, 0}) == float3(0.0f)); // below RayEpsilon collapses to zero
    Basis basis = constructBasis(float3{0, 0, 1});
    CHECK(basis.T == float3(1, 0, 0) && basis.B == float3(0, -1, 0));
    PixelRect r(PixelCoord{3, 9}, PixelCoord{1, 2});
    CHECK(r.min().i == 1 && r.min().j == 2 && r.max().i == 3 && r.max().j == 9 && r.width() == 3 && r.height() == 8);
    CHECK_THROWS(PixelRect(0, 4));
}

static void testTiles() { // reference tests/test_Tiles.cpp
    CHECK_THROWS((FrameTiling{PixelRect{0, 1}, PixelRect{16, 16}}));
    CHECK_THROWS((FrameTiling{PixelRect{1, 0}, PixelRect{16, 16}}));
    CHECK_THROWS((FrameTiling{PixelRect{5, 5}, PixelRect{0, 16}}));
    CHECK_THROWS((FrameTiling{PixelRect{5, 5}, PixelRect{16, 0}}));
    FrameTiling tiling{PixelRect{32, 9}, PixelRect{16, 3}};
    CHECK(tiling.size() == 2 * 3);
    for (std::size_t i = 0; i != tiling.size(); i++) {
        int x = static_cast<int>(i % 2), y = static_cast<int>(i / 2);
        CHECK(tiling[i].tileNumber == i);
        CHECK(tiling[i].bounds == PixelRect(PixelCoord{x * 16, y * 3}, PixelCoord{(x + 1) * 16 - 1, (y + 1) * 3 - 1}));
    }
    // frames that are not a multiple of the tile: every pixel exactly once (the reference's spill case, fixed)
    FrameTiling spill{PixelRect{1920, 1080}, PixelRect{32, 32}};
    CHECK(spill.size() == 60 * 34);
    std::size_t covered = 0;
    for (auto const &t : spill) {
        covered += static_cast<std::size_t>(t.bounds.area());
        CHECK(t.bounds.max().i < 1920 && t.bounds.max().j < 1080);
    }
    CHECK(covered == 1920u * 1080u);
    CHECK(spill[spill.size() - 1].bounds == PixelRect(PixelCoord{1888, 1056}, PixelCoord{1919, 1079}));
}

static void testFrameBufferAndColor() { // reference tests/test_FrameBuffer.cpp, test_Color.cpp
    RGBFrameBuffer fb(PixelRect(128, 64));
    CHECK_THROWS(RGBFrameBuffer(PixelRect(128, 0)));
    CHECK_THROWS(RGBFrameBuffer(PixelRect(0, 128)));
    CHECK(fb.aspect() == 2.0);
    CHECK(fb(0, 0)(0) == 0.0f && fb(0, 0)(1) == 0.0f && fb(0, 0)(2) == 0.0f);
    fb(0, 0) = RGB::red();
    CHECK(fb(0, 0)(0) == 1.0f && fb(0, 0)(1) == 0.0f);
    fb(5, 2) = RGB::green();
    CHECK(fb.data()[2 * 128 + 5](1) == 1.0f); // (i, j) -> j * width + i
    CHECK(quantizeTo8bit(1.0) == 255 && quantizeTo8bit(0.0) == 0 && quantizeTo8bit(0.5) == 128);
    CHECK(quantizeTo8bit(5.0) == 255 && quantizeTo8bit(-5.0) == 0);
    auto q = quantizeTo8bit(SRGB{{5.0f, 1.0f, 0.0f}});
    CHECK(q[0] == 255 && q[1] == 255 && q[2] == 0);
    RGB rgb(1.0f, -2.0f, 3.0f);
    RGB sum = rgb + rgb;
    CHECK(sum(0) == 2.0f && sum(1) == -4.0f && sum(2) == 6.0f);
    RGB half = RGB(1.0f, -2.0f, 4.0f) * 0.5f;
    CHECK(half(0) == 0.5f && half(1) == -1.0f && half(2) == 2.0f);
    SRGB s = toSRGB(RGB::black());
    CHECK(s(0) == 0.0f && s(1) == 0.0f && s(2) == 0.0f);
    s = toSRGB(RGB{0.5f, 0.5f, 0.5f});
    CHECK(near(s(0), 0.7353, 0.01));
    s = toSRGB(RGB{1.0f, 1.0f, 1.0f});
    CHECK(near(s(0), 1.0, 0.01));
}

static SceneDescription smallScene() {
    SceneDescription d;
    PerspectiveCameraDescription cam;
    cam.origin = V3(0, 275, -1100);
    cam.lookAt = V3(0, 275, 0);
    cam.aspect = 1.0f;
    cam.horizontalFov = 0.7f;
    d.setCamera(cam);
    MaterialDescription white;
    white.albedo = RGB(.73f, .73f, .73f);
    auto w = d.addMaterial(white);
    MaterialDescription light;
    light.albedo = RGB::black();
    light.emissive = RGB(15, 15, 15);
    auto l = d.addMaterial(light);
    PlaneDescription floor;
    floor.extents = V3(555, 555, 0);
    floor.material = w;
    d.addPlane(floor);
    PlaneDescription back;
    back.normal = V3(0, 0, -1);
    back.point = V3(0, 275, 275);
    back.extents = V3(555, 555, 0);
    d.addPlane(back); // no material: uses the default material 0
    SphereDescription lamp;
    lamp.center = V3(0, 495, 0);
    lamp.radius = 60;
    lamp.material = l;
    d.addSphere(lamp);
    return d;
}

static void testSceneDescription() { // reference tests/test_SceneDescription.cpp
    SceneDescription d;
    CHECK(d.materials().size() == 1);
    CHECK(d.materials()[0] == MaterialDescription{});
    MaterialDescription m;
    m.roughness = 0.7f;
    CHECK(d.addMaterial(m) == 1);
    CHECK(d.materials().size() == 2 && d.materials()[1] == m);
    SphereDescription s;
    s.radius = 2.0f;
    CHECK(d.addSphere(s) == 0 && d.spheres().size() == 1 && d.spheres()[0] == s && !d.spheres()[0].material);
    PlaneDescription p;
    CHECK(d.addPlane(p) == 0 && d.planes().size() == 1 && d.planes()[0] == p);
    CHECK(RenderOptions{}.samplesAA == 256 && RenderOptions{}.width == 512 && RenderOptions{}.seed == 19791102u);
}

static void testNoDevice() {
    // Without a CUDA device the host layer must refuse: there is no CPU rendering path behind RenderSession.
    bool threw = false;
    try {
        RenderOptions o;
        o.saveImage = false;
        RenderSession session(smallScene(), o);
    } catch (RenderError const &e) {
        threw = true;
        CHECK(std::strstr(e.what(), "no CPU path") != nullptr);
    }
    CHECK(threw);
}

static void testRenderSession(int devices) {
    RenderOptions o;
    o.samplesAA = 64;
    o.width = 96;
    o.height = 64;
    o.saveImage = true;
    o.outputPath = "/tmp/cornelis_host_test.png";
    o.devices = devices;
    RenderSession session(smallScene(), o);
    std::atomic<int> calls{0}, done{0};
    session.render([&](RenderProgress const &p, RenderStatus const &st) {
        calls++;
        if (st == RenderStatus::Done) {
            done++;
            CHECK(p.samplesDone == p.samplesTotal && p.samplesTotal == 96u * 64u * 64u);
        }
        return RenderCommand::Continue;
    });
    CHECK(calls >= 1 && done == 1);
    auto const &fb = session.frameBuffer();
    CHECK(fb.width() == 96 && fb.height() == 64);
    double sum = 0;
    bool finite = true;
    for (auto const &px : fb)
        for (int c = 0; c < 3; c++) {
            sum += px(c);
            finite = finite && std::isfinite(px(c));
        }
    CHECK(finite && sum > 0.0);
    // the lamp is in the upper half of the image (j = 0 is the top row)
    double top = 0, bottom = 0;
    for (int j = 0; j < 64; j++)
        for (int i = 0; i < 96; i++)
            (j < 32 ? top : bottom) += fb(i, j)(0);
    CHECK(top > bottom);
    CHECK(session.statistics().pixelSamples == 96u * 64u * 64u && session.statistics().rays > session.statistics().pixelSamples);
    std::FILE *png = std::fopen(o.outputPath.c_str(), "rb");
    CHECK(png != nullptr);
    if (png) {
        unsigned char sig[8];
        CHECK(std::fread(sig, 1, 8, png) == 8 && sig[1] == 'P' && sig[2] == 'N' && sig[3] == 'G');
        std::fclose(png);
    }
    // same scene, same seed -> same estimate up to fp32 summation order; different device counts agree too
    RenderSession again(smallScene(), o);
    again.render();
    double diff = 0;
    for (int j = 0; j < 64; j++)
        for (int i = 0; i < 96; i++)
            diff = std::max(diff, static_cast<double>(std::fabs(again.frameBuffer()(i, j)(0) - fb(i, j)(0))));
    CHECK(diff < 1e-4);

    // abort from the callback
    RenderOptions big = o;
    big.samplesAA = 4096;
    big.width = 256;
    big.height = 256;
    big.poolPaths = 65536;
    big.saveImage = false;
    RenderSession abortable(smallScene(), big);
    int aborted = 0;
    abortable.render([&](RenderProgress const &, RenderStatus const &st) {
        if (st == RenderStatus::Aborted)
            aborted++;
        return RenderCommand::Abort;
    });
    CHECK(aborted == 1);

    // progressive: slices of the global sample range (n, n, 2n, 4n, ...), an image after every slice, and at the end
    // exactly the paths of the one-shot render (random numbers are keyed by the global sample index)
    RenderOptions prog = o;
    prog.progressive = true;
    prog.saveImage = false;
    RenderSession progressive(smallScene(), prog);
    int running = 0, finished = 0;
    std::uint64_t lastDone = 0;
    bool monotone = true, imageEachSlice = true;
    progressive.render([&](RenderProgress const &p, RenderStatus const &st) {
        monotone = monotone && p.samplesDone > lastDone && p.samplesDone <= p.samplesTotal;
        lastDone = p.samplesDone;
        if (st == RenderStatus::Running) {
            running++;
            double energy = 0;
            for (auto const &px : progressive.frameBuffer())
                energy += px(0) + px(1) + px(2);
            imageEachSlice = imageEachSlice && std::isfinite(energy) && energy > 0.0;
        }
        if (st == RenderStatus::Done)
            finished++;
        return RenderCommand::Continue;
    });
    CHECK(monotone && imageEachSlice && finished == 1 && running >= 3);
    CHECK(progressive.statistics().samplesPerPixel == 64 && progressive.statistics().slices == static_cast<unsigned>(running) + 1u);
    CHECK(progressive.statistics().pixelSamples == 96u * 64u * 64u && progressive.statistics().rays == session.statistics().rays);
    diff = 0;
    for (int j = 0; j < 64; j++)
        for (int i = 0; i < 96; i++)
            diff = std::max(diff, static_cast<double>(std::fabs(progressive.frameBuffer()(i, j)(0) - fb(i, j)(0))));
    CHECK(diff < 1e-4);

    // time budget: samplesAA is only the upper limit; the image holds samplesPerPixel samples and is an unbiased
    // estimate of the same picture
    RenderOptions timed = o;
    timed.samplesAA = 1 << 23; // far more than any number of B200s renders of this frame in the budget (2 GPUs reach 2^20)
    timed.timeBudgetSeconds = 0.15;
    timed.saveImage = false;
    timed.dropNonFinite = true; // half a million spp meet the reference's NaN (|w.z| > 1 in Oren-Nayar) a few times
    RenderSession budgeted(smallScene(), timed);
    auto const t0 = std::chrono::steady_clock::now();
    budgeted.render();
    double const took = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    auto const &bs = budgeted.statistics();
    CHECK(bs.samplesPerPixel >= devices && bs.samplesPerPixel < (1 << 23) && bs.slices >= 2);
    CHECK(bs.pixelSamples == 96ull * 64ull * static_cast<unsigned long long>(bs.samplesPerPixel));
    CHECK(took < 1.0);
    double sumBudgeted = 0;
    for (auto const &px : budgeted.frameBuffer())
        sumBudgeted += px(0) + px(1) + px(2);
    double sumOneShot = 0;
    for (auto const &px : fb)
        sumOneShot += px(0) + px(1) + px(2);
    CHECK(std::fabs(sumBudgeted / sumOneShot - 1.0) < 0.15); // the 64-spp one-shot image is the noisy side
    std::printf("time budget 0.15 s: %d spp in %u slices, %.3f s\n", bs.samplesPerPixel, bs.slices, took);

    // samplesAA <= 0: message and silent return, as the reference (Render.cpp:310-313)
    RenderOptions zero = o;
    zero.samplesAA = 0;
    zero.saveImage = false;
    RenderSession noop(smallScene(), zero);
    noop.render();
}

// `png <path> <width> <height>`: saveImage (display transform, quantisation, PNG encoding — no GPU involved) on a
// deterministic image; tests/test_host_api.py decodes the file and checks every pixel.
static int writeTestPng(char const *path, int width, int height) {
    RGBFrameBuffer fb(PixelRect(width, height));
    for (int j = 0; j < height; j++)
        for (int i = 0; i < width; i++)
            fb(i, j) = RGB(static_cast<float>(i) / static_cast<float>(width), static_cast<float>(j) / static_cast<float>(height),
                           static_cast<float>((i * 7 + j * 13) % 32) / 16.0f); // the third channel exceeds 1 and is noisy
    saveImage(fb, path);
    return 0;
}

int main(int argc, char **argv) {
    std::string const mode = argc > 1 ? argv[1] : "cpu";
    if (mode == "png" && argc > 4)
        return writeTestPng(argv[2], std::atoi(argv[3]), std::atoi(argv[4]));
    testCamera();
    testMath();
    testTiles();
    testFrameBufferAndColor();
    testSceneDescription();
    if (mode == "cpu-nodevice")
        testNoDevice();
    if (mode == "gpu") {
        testRenderSession(1);
        int devices = argc > 2 ? std::atoi(argv[2]) : 1;
        if (devices > 1)
            testRenderSession(devices);
    }
    std::printf("%s: %d failure(s)\n", mode.c_str(), failures);
    return failures ? 1 : 0;
}
