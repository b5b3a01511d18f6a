// RenderSession on top of the C-ABI (include/cornelis_cuda.h).
//
// Where the reference's RenderSession::render builds a FrameTiling, seeds one PRNG per tile and runs integrateTile on
// TBB workers (reference src/Render.cpp:302-363), this one hands the whole frame to the GPU wavefront: one scene handle
// per device, each device renders a contiguous range of the global sample indices of every pixel, the accumulators
// are summed on device 0 and resolved into the RGBFrameBuffer.  No CPU rendering path exists here.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cornelis/Render.hpp>
#include <cornelis_cuda.h>

namespace cornelis {

bool writePngRgb8(std::string const &path, int width, int height, unsigned char const *rgb);

namespace {

[[noreturn]] void raise(int code, char const *where) {
    throw RenderError(code, std::string(where) + ": " + cornelis_cuda_last_error());
}

struct DeviceScene {
    cornelis_cuda_scene *handle = nullptr;
    ~DeviceScene() {
        if (handle)
            cornelis_cuda_scene_destroy(handle);
    }
};

} // namespace

struct RenderSession::State {
    State(SceneDescription const &sc, RenderOptions opts)
        : sceneDescr(sc), options(std::move(opts)), fb(PixelRect(std::max(options.width, 1), std::max(options.height, 1))) {
        CORNELIS_EXPECTS(options.width > 0 && options.height > 0, "Frame cannot be a line or the empty rectangle.");
        CORNELIS_EXPECTS(options.devices > 0, "At least one device is needed.");
        int available = 0;
        if (int rc = cornelis_cuda_device_count(&available))
            raise(rc, "RenderSession");
        if (options.devices > available)
            throw RenderError(CORNELIS_ERR_INVALID_ARGUMENT, "RenderSession: more devices requested than this box has");
        upload();
    }

    // SceneDescription -> PODs of the C-ABI (the flattening SceneData does in reference src/Scene.cpp:5-53).
    void upload() {
        auto const cam = sceneDescr.camera();
        cornelis_camera_desc c{};
        for (int k = 0; k < 3; k++) {
            c.origin[k] = cam.origin[k];
            c.look_at[k] = cam.lookAt[k];
        }
        c.aspect = cam.aspect;
        c.horizontal_fov = cam.horizontalFov;
        std::vector<cornelis_material_desc> mats;
        for (auto const &m : sceneDescr.materials()) {
            cornelis_material_desc d{};
            for (int k = 0; k < 3; k++) {
                d.albedo[k] = m.albedo(k);
                d.emissive[k] = m.emissive(k);
                d.reflection_tint[k] = m.reflectionTint(k);
            }
            d.roughness = m.roughness;
            d.ior = m.ior;
            mats.push_back(d);
        }
        std::vector<cornelis_sphere_desc> spheres;
        for (auto const &s : sceneDescr.spheres()) {
            cornelis_sphere_desc d{};
            for (int k = 0; k < 3; k++)
                d.center[k] = s.center[k];
            d.radius = s.radius;
            d.material = s.material ? static_cast<int32_t>(*s.material) : -1;
            spheres.push_back(d);
        }
        std::vector<cornelis_plane_desc> planes;
        for (auto const &p : sceneDescr.planes()) {
            cornelis_plane_desc d{};
            for (int k = 0; k < 3; k++) {
                d.normal[k] = p.normal[k];
                d.point[k] = p.point[k];
                d.extents[k] = p.extents[k];
            }
            d.material = p.material ? static_cast<int32_t>(*p.material) : -1;
            planes.push_back(d);
        }
        devices.resize(static_cast<std::size_t>(options.devices));
        for (int dev = 0; dev < options.devices; dev++) {
            if (int rc = cornelis_cuda_scene_create(dev, &c, spheres.data(), spheres.size(), planes.data(), planes.size(),
                                                    mats.data(), mats.size(), &devices[static_cast<std::size_t>(dev)].handle))
                raise(rc, "cornelis_cuda_scene_create");
            if (options.acceleration != Acceleration::Auto)
                if (int rc = cornelis_cuda_scene_set_acceleration(devices[static_cast<std::size_t>(dev)].handle,
                                                                  static_cast<int>(options.acceleration)))
                    raise(rc, "cornelis_cuda_scene_set_acceleration");
        }
    }

    SceneDescription sceneDescr;
    RenderOptions options;
    RGBFrameBuffer fb;
    RenderStatistics stats;
    std::vector<DeviceScene> devices;
};

RenderSession::RenderSession(SceneDescription const &sc, RenderOptions options)
    : me_{std::make_unique<State>(sc, std::move(options))} {}

RenderSession::~RenderSession() = default;
RenderSession::RenderSession(RenderSession &&) noexcept = default;
RenderSession &RenderSession::operator=(RenderSession &&) noexcept = default;

void RenderSession::render() {
    render([](RenderProgress const &, RenderStatus const &) { return RenderCommand::Continue; });
}

RGBFrameBuffer const &RenderSession::frameBuffer() const { return me_->fb; }
RenderStatistics const &RenderSession::statistics() const { return me_->stats; }

namespace {
struct ProgressRelay {
    RenderSession::ProgressCallback *callback;
    std::vector<std::atomic<std::uint64_t>> *done;
    std::size_t device;
    std::uint64_t total;
    std::uint64_t doneBefore; // pixel-samples of earlier slices
    std::atomic<bool> *abort;
};

int relay(void *user, std::uint64_t samplesDone, std::uint64_t) {
    auto *r = static_cast<ProgressRelay *>(user);
    (*r->done)[r->device].store(samplesDone);
    RenderProgress p;
    p.samplesDone = r->doneBefore;
    for (auto const &d : *r->done)
        p.samplesDone += d.load();
    p.samplesTotal = r->total;
    if ((*r->callback)(p, RenderStatus::Running) != RenderCommand::Continue)
        r->abort->store(true);
    return r->abort->load() ? 1 : 0;
}
} // namespace

void RenderSession::render(ProgressCallback onProgress) {
    State &s = *me_;
    RenderOptions const &o = s.options;
    if (o.samplesAA <= 0) { // reference src/Render.cpp:310-313: message and silent return
        std::printf("AA Samples must be > 0 (not %d).\n", o.samplesAA);
        return;
    }
    std::size_t const n = s.devices.size();
    std::uint64_t const npixels = static_cast<std::uint64_t>(o.width) * static_cast<std::uint64_t>(o.height);
    std::uint64_t const total = npixels * static_cast<std::uint64_t>(o.samplesAA);
    bool const progressive = o.progressive || o.timeBudgetSeconds > 0.0;
    std::vector<std::atomic<std::uint64_t>> done(n);
    std::atomic<bool> abort{false};
    std::vector<int> rc(n, 0);
    std::vector<std::string> errors(n);
    std::vector<cornelis_render_stats> stats(n);
    s.stats = RenderStatistics{};
    std::uint64_t doneBefore = 0; // pixel-samples of the slices already finished

    // Renders the global sample indices [first, first + count) of every pixel, split over the devices as contiguous
    // sub-ranges (the counter-based RNG makes the union identical for any device count), and leaves the sum of ALL
    // samples rendered so far in device 0's accumulators: device 0 keeps its accumulators from slice to slice, the
    // other devices start every slice from zero and are added to it.  Returns false if the render was aborted.
    auto renderRange = [&](std::int32_t first, std::int32_t count, bool keep) -> bool {
        for (auto &d : done)
            d.store(0);
        auto work = [&](std::size_t dev) {
            // device 0 takes the LAST sub-range, which is never empty: the sum of all devices lives on device 0
            std::int64_t const part = static_cast<std::int64_t>(n - 1 - dev), parts = static_cast<std::int64_t>(n);
            std::int32_t const lo = first + static_cast<std::int32_t>(static_cast<std::int64_t>(count) * part / parts);
            std::int32_t const hi = first + static_cast<std::int32_t>(static_cast<std::int64_t>(count) * (part + 1) / parts);
            stats[dev] = cornelis_render_stats{};
            if (hi == lo) {
                rc[dev] = -1; // nothing to do on this device
                return;
            }
            cornelis_render_params p{};
            p.width = o.width;
            p.height = o.height;
            p.samples = o.samplesAA;
            p.first_sample = lo;
            p.sample_count = hi - lo;
            p.max_depth = o.maxDepth;
            p.seed = o.seed;
            p.flags = (o.dropNonFinite ? CORNELIS_RENDER_DROP_NONFINITE : 0u) | (keep && dev == 0 ? CORNELIS_RENDER_KEEP : 0u);
            p.pool_paths = o.poolPaths;
            ProgressRelay r{&onProgress, &done, dev, total, doneBefore, &abort};
            // a progressive render reports after every slice instead of from inside the slices
            rc[dev] = cornelis_cuda_render_accumulate(s.devices[dev].handle, &p, progressive ? nullptr : relay, &r, &stats[dev]);
            if (rc[dev])
                errors[dev] = cornelis_cuda_last_error();
        };
        if (n == 1) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (std::size_t dev = 0; dev < n; dev++)
                pool.emplace_back(work, dev);
            for (auto &t : pool)
                t.join();
        }
        bool aborted = false;
        for (std::size_t dev = 0; dev < n; dev++) {
            if (rc[dev] == CORNELIS_ERR_ABORTED)
                aborted = true;
            else if (rc[dev] > 0) {
                RenderProgress failed;
                failed.samplesDone = doneBefore;
                failed.samplesTotal = total;
                onProgress(failed, RenderStatus::Failed);
                throw RenderError(rc[dev], "cornelis_cuda_render_accumulate: " + errors[dev]);
            }
        }
        std::vector<cornelis_cuda_scene *> rendered;
        double slowest = 0.0;
        for (std::size_t dev = 0; dev < n; dev++) {
            if (rc[dev] == -1)
                continue;
            rendered.push_back(s.devices[dev].handle);
            s.stats.pixelSamples += stats[dev].pixel_samples;
            s.stats.rays += stats[dev].rays;
            s.stats.passes += stats[dev].iterations;
            s.stats.kernelLaunches += stats[dev].kernel_launches;
            s.stats.maxDepth = std::max(s.stats.maxDepth, stats[dev].max_depth);
            slowest = std::max(slowest, 1e-3 * static_cast<double>(stats[dev].gpu_ms));
        }
        s.stats.gpuSeconds += slowest;
        s.stats.slices += 1;
        // count >= 1 puts device 0 first in `rendered`: the sum lands where the next slice keeps accumulating
        if (rendered.size() > 1)
            if (int e = cornelis_cuda_reduce_framebuffers(rendered.data(), static_cast<int>(rendered.size())))
                raise(e, "cornelis_cuda_reduce_framebuffers");
        doneBefore = s.stats.pixelSamples;
        return !aborted;
    };
    // color = sum * (1 / samples) into the RGBFrameBuffer: RGB is three packed floats, the host_rgb layout of the C-ABI
    auto resolve = [&](std::int32_t samples) {
        if (int e = cornelis_cuda_resolve(s.devices.front().handle, samples, reinterpret_cast<float *>(s.fb.data()), nullptr))
            raise(e, "cornelis_cuda_resolve");
        s.stats.samplesPerPixel = samples;
    };

    bool aborted = false;
    if (!progressive) {
        aborted = !renderRange(0, o.samplesAA, false);
        resolve(o.samplesAA);
    } else {
        // Slices grow geometrically (n, n, 2n, 4n, ... samples per pixel for n devices) so that the first image arrives
        // after one sample per device and the per-slice overhead stays small; with a time budget the next slice is
        // shrunk to what, at the rate measured so far, still ends inside the budget.
        auto const start = std::chrono::steady_clock::now();
        auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count(); };
        std::int32_t rendered = 0;
        std::int32_t slice = static_cast<std::int32_t>(std::min<std::size_t>(n, static_cast<std::size_t>(o.samplesAA)));
        while (rendered < o.samplesAA) {
            slice = std::min(slice, o.samplesAA - rendered);
            if (o.timeBudgetSeconds > 0.0 && rendered > 0) {
                double const perSample = elapsed() / static_cast<double>(rendered);
                double const left = o.timeBudgetSeconds - elapsed();
                if (left <= 0.0)
                    break;
                if (perSample * slice > left) { // shrink the last slice to what fits; stop when not even n samples do
                    slice = static_cast<std::int32_t>(left / perSample);
                    if (slice < static_cast<std::int32_t>(n))
                        break;
                }
            }
            aborted = !renderRange(rendered, slice, rendered > 0);
            if (aborted)
                break; // the slice is incomplete: the frame buffer keeps the last complete estimate
            rendered += slice;
            resolve(rendered);
            if (rendered < o.samplesAA) {
                RenderProgress p;
                p.samplesDone = doneBefore;
                p.samplesTotal = total;
                if (onProgress(p, RenderStatus::Running) != RenderCommand::Continue) {
                    aborted = true;
                    break;
                }
            }
            slice = rendered; // n, n, 2n, 4n, ...: every slice doubles the samples in the image
        }
        if (rendered == 0 && !aborted) { // a budget too small for a single slice still gets one sample per device
            aborted = !renderRange(0, slice, false);
            if (!aborted)
                resolve(slice);
        }
    }

    RenderProgress final;
    final.samplesTotal = total;
    final.samplesDone = s.stats.pixelSamples;
    onProgress(final, aborted ? RenderStatus::Aborted : RenderStatus::Done);
    if (o.saveImage)
        saveImage(s.fb, o.outputPath);
}

void saveImage(RGBFrameBuffer const &fb, std::string const &path) {
    SRGBFrameBuffer display(PixelRect(fb.width(), fb.height()));
    std::transform(fb.begin(), fb.end(), display.begin(), toSRGB);
    auto const bytes = quantizeTo8bit(display);
    if (!writePngRgb8(path, bytes.width(), bytes.height(), reinterpret_cast<unsigned char const *>(bytes.data())))
        throw std::runtime_error("saveImage: cannot write " + path);
}

} // namespace cornelis
