// Render options.  The reference has a single field, samplesAA (include/cornelis/RenderOptions.hpp:6-16); everything
// else it hard-codes — 512x512 (Render.cpp:307), seed 19791102 (PRNG.hpp:12), no depth limit (Render.cpp:237), the
// output name (Render.cpp:264).  Those literals are the defaults here, so RenderOptions{} and
// RenderOptions{samples} behave like the reference.
#pragma once

#include <cstdint>
#include <string>

namespace cornelis {

// How the intersect stage finds the closest sphere (include/cornelis_cuda.h CORNELIS_ACCEL_*): the reference scans
// every sphere for every ray (Render.cpp:115-123); the uniform grid returns the same hit, bit for bit, from the
// spheres near the ray.  Auto picks the grid for scenes with many spheres.
enum class Acceleration : std::int32_t { Auto = 0, None = 1, Grid = 2 };

struct RenderOptions {
    static constexpr std::int32_t DefaultSamplesAA = 1 << 8;

    // Samples per pixel for anti-aliasing — the main quality control (a Monte-Carlo path tracer's noise level).
    std::int32_t samplesAA = DefaultSamplesAA;

    // ---- extensions (defaults = the reference's literals) ----
    std::int32_t width = 512;
    std::int32_t height = 512;
    std::uint64_t seed = 19791102;
    std::int32_t maxDepth = 0;       // 0 = unlimited; otherwise a path stops after this many bounces
    std::int32_t devices = 1;        // GPUs of this box to shard the samples over
    std::int32_t poolPaths = 0;      // paths in flight per GPU (0 = library default)
    Acceleration acceleration = Acceleration::Auto;
    // Progressive rendering (the reference's milestone "progressive/time-budgeted mode", README.md:33, on top of
    // sample-range rendering): the samples are rendered in growing slices of the global sample index range, the
    // frame buffer holds the estimate of all samples so far after every slice, and the progress callback runs after
    // every slice (it may read RenderSession::frameBuffer()).  With a time budget the render stops when the next slice
    // would overrun it — samplesAA is then the upper limit — and RenderStatistics::samplesPerPixel tells how many
    // samples the image holds.  Because random numbers are keyed by the global sample index, a progressive render that
    // runs to samplesAA traces exactly the paths of the one-shot render.
    bool progressive = false;
    double timeBudgetSeconds = 0.0;  // > 0 implies progressive
    bool dropNonFinite = false;      // skip NaN/inf path contributions (the reference lets them through)
    bool saveImage = true;           // write the PNG at the end of render(), as the reference does
    std::string outputPath = "cornelisrender2.png";
};

} // namespace cornelis
