// RGB arithmetic and the display transform (reference src/Color.cpp).
#include <algorithm>
#include <cmath>

#include <cornelis/Color.hpp>

namespace cornelis {

RGB &RGB::operator+=(RGB const &o) noexcept {
    for (int c = 0; c < 3; c++)
        values[c] += o.values[c];
    return *this;
}

RGB &RGB::operator*=(RGB const &o) noexcept {
    for (int c = 0; c < 3; c++)
        values[c] *= o.values[c];
    return *this;
}

RGB RGB::operator/(float s) const noexcept { return {values[0] / s, values[1] / s, values[2] / s}; }

RGB RGB::clamp(float lo, float hi) const noexcept {
    return {std::clamp(values[0], lo, hi), std::clamp(values[1], lo, hi), std::clamp(values[2], lo, hi)};
}

namespace {
// Reference Color.cpp:64-78: slope 12.95 (sic) on the linear toe, double-precision pow above it.
float encodeChannel(float x) {
    float const a = 0.055f;
    if (static_cast<double>(x) <= 0.0031308)
        return x * 12.95f;
    return static_cast<float>((1 + a) * std::pow(static_cast<double>(x), static_cast<double>(1.0f / 2.4f)) - a);
}
} // namespace

SRGB toSRGB(RGB const &rgb) { return SRGB{{encodeChannel(rgb(0)), encodeChannel(rgb(1)), encodeChannel(rgb(2))}}; }

} // namespace cornelis
