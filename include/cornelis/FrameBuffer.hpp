// Row-major frame buffer: pixel (i, j) lives at j * width + i, j = 0 is the TOP row
// (reference include/cornelis/FrameBuffer.hpp:18-107).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <vector>

#include <cornelis/Color.hpp>
#include <cornelis/Expects.hpp>
#include <cornelis/Math.hpp>

namespace cornelis {

template <typename TPixel>
class FrameBuffer {
  public:
    using value_type = TPixel;
    using container_type = std::vector<value_type>;
    using iterator = typename container_type::iterator;
    using const_iterator = typename container_type::const_iterator;

    // Throws ExpectationException for a zero-sized dimension (through PixelRect).
    explicit FrameBuffer(PixelRect dims)
        : dims_(dims), pixels_(static_cast<std::size_t>(dims.width()) * static_cast<std::size_t>(dims.height())) {}

    value_type const &operator()(PixelCoord::value_type i, PixelCoord::value_type j) const noexcept {
        return pixels_[static_cast<std::size_t>(j) * width() + i];
    }
    value_type &operator()(PixelCoord::value_type i, PixelCoord::value_type j) noexcept {
        return pixels_[static_cast<std::size_t>(j) * width() + i];
    }
    value_type const &operator()(PixelCoord const &c) const noexcept { return (*this)(c.i, c.j); }
    value_type &operator()(PixelCoord const &c) noexcept { return (*this)(c.i, c.j); }

    double aspect() const noexcept { return static_cast<double>(width()) / height(); }
    PixelCoord::value_type width() const noexcept { return dims_.width(); }
    PixelCoord::value_type height() const noexcept { return dims_.height(); }

    iterator begin() noexcept { return pixels_.begin(); }
    iterator end() noexcept { return pixels_.end(); }
    const_iterator begin() const noexcept { return pixels_.begin(); }
    const_iterator end() const noexcept { return pixels_.end(); }
    value_type const *data() const noexcept { return pixels_.data(); }
    value_type *data() noexcept { return pixels_.data(); }

  private:
    PixelRect dims_;
    container_type pixels_;
};

using RGBFrameBuffer = FrameBuffer<RGB>;
using SRGBFrameBuffer = FrameBuffer<SRGB>;

// round(255 v), saturated to [0, 255] (reference FrameBuffer.hpp:91-95).
inline std::uint8_t quantizeTo8bit(double v) {
    v = std::round(255.0 * v);
    return static_cast<std::uint8_t>(std::clamp(v, 0.0, 255.0));
}

inline std::array<std::uint8_t, 3> quantizeTo8bit(SRGB const &v) {
    return {quantizeTo8bit(v(0)), quantizeTo8bit(v(1)), quantizeTo8bit(v(2))};
}

inline FrameBuffer<std::array<std::uint8_t, 3>> quantizeTo8bit(SRGBFrameBuffer const &fb) {
    FrameBuffer<std::array<std::uint8_t, 3>> out(PixelRect(fb.width(), fb.height()));
    std::transform(fb.begin(), fb.end(), out.begin(), [](SRGB const &v) { return quantizeTo8bit(v); });
    return out;
}

} // namespace cornelis
