// Linear RGB triplets and the display transform (reference include/cornelis/Color.hpp, src/Color.cpp).
#pragma once

#include <cstddef>

namespace cornelis {

// Non-linear (display) sRGB.
struct SRGB {
    float values[3];
    float const &operator()(std::size_t c) const noexcept { return values[c]; }
    float &operator()(std::size_t c) noexcept { return values[c]; }
};

// Linear sRGB; three packed floats — the pixel layout of RGBFrameBuffer and of the C-ABI's host_rgb output.
struct RGB {
    float values[3];

    constexpr RGB() : values{0.0f, 0.0f, 0.0f} {}
    constexpr RGB(float r, float g, float b) : values{r, g, b} {}

    float const &operator()(std::size_t c) const noexcept { return values[c]; }
    float &operator()(std::size_t c) noexcept { return values[c]; }

    static constexpr RGB black() { return {}; }
    static constexpr RGB red() { return {1.0f, 0.0f, 0.0f}; }
    static constexpr RGB green() { return {0.0f, 1.0f, 0.0f}; }
    static constexpr RGB blue() { return {0.0f, 0.0f, 1.0f}; }

    RGB &operator+=(RGB const &o) noexcept;
    RGB &operator*=(RGB const &o) noexcept; // component-wise
    RGB operator/(float s) const noexcept;
    RGB clamp(float lo, float hi) const noexcept;

    bool operator==(RGB const &o) const noexcept {
        return values[0] == o.values[0] && values[1] == o.values[1] && values[2] == o.values[2];
    }
    bool operator!=(RGB const &o) const noexcept { return !(*this == o); }
};
static_assert(sizeof(RGB) == 12, "RGB must stay three packed floats");

inline RGB operator+(RGB const &a, RGB const &b) noexcept { return {a(0) + b(0), a(1) + b(1), a(2) + b(2)}; }
inline RGB operator-(RGB const &a, RGB const &b) noexcept { return {a(0) - b(0), a(1) - b(1), a(2) - b(2)}; }
inline RGB operator*(RGB const &a, RGB const &b) noexcept { return {a(0) * b(0), a(1) * b(1), a(2) * b(2)}; }
inline RGB operator*(RGB const &a, float s) noexcept { return {a(0) * s, a(1) * s, a(2) * s}; }
inline RGB operator*(float s, RGB const &a) noexcept { return {s * a(0), s * a(1), s * a(2)}; }

// The sRGB transfer function as the reference writes it (Color.cpp:64-80): linear slope 12.95 below 0.0031308,
// 1.055 * x^(1/2.4) - 0.055 above, the power evaluated in double.
SRGB toSRGB(RGB const &);

} // namespace cornelis
