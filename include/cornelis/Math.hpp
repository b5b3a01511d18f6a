// Host-side math types of the cornelis API: V3 (the reference aliases nanovdb::Vec3<float>, Math.hpp:14), Ray,
// float3 helpers, PixelCoord / PixelRect.  The arithmetic follows the reference's operation order so host-computed
// values (camera, plane bases) carry the same bits as the reference's (Math.hpp:278-292, 380-434; NanoVDB.h:911-953).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <ostream>

#include <cornelis/Expects.hpp>

namespace cornelis {

inline constexpr float RayEpsilon = 0.00005f;   // reference Math.hpp:20
inline constexpr float Pi = 3.14159265359f;     // reference Math.hpp:25

inline bool isAlmostZero(float v) { return std::fabs(v) < RayEpsilon; }

// Three floats with the subset of nanovdb::Vec3<float> the cornelis API uses.
class V3 {
  public:
    V3() : v_{0.0f, 0.0f, 0.0f} {}
    explicit V3(float s) : v_{s, s, s} {}
    V3(float x, float y, float z) : v_{x, y, z} {}

    float const &operator[](int i) const { return v_[i]; }
    float &operator[](int i) { return v_[i]; }
    bool operator==(V3 const &o) const { return v_[0] == o.v_[0] && v_[1] == o.v_[1] && v_[2] == o.v_[2]; }
    bool operator!=(V3 const &o) const { return !(*this == o); }

    V3 operator-() const { return {-v_[0], -v_[1], -v_[2]}; }
    V3 operator+(V3 const &o) const { return {v_[0] + o.v_[0], v_[1] + o.v_[1], v_[2] + o.v_[2]}; }
    V3 operator-(V3 const &o) const { return {v_[0] - o.v_[0], v_[1] - o.v_[1], v_[2] - o.v_[2]}; }
    V3 operator*(float s) const { return {s * v_[0], s * v_[1], s * v_[2]}; }
    V3 &operator*=(float s) {
        v_[0] *= s, v_[1] *= s, v_[2] *= s;
        return *this;
    }
    float dot(V3 const &o) const { return v_[0] * o.v_[0] + v_[1] * o.v_[1] + v_[2] * o.v_[2]; }
    V3 cross(V3 const &o) const {
        return {v_[1] * o.v_[2] - v_[2] * o.v_[1], v_[2] * o.v_[0] - v_[0] * o.v_[2], v_[0] * o.v_[1] - v_[1] * o.v_[0]};
    }
    float lengthSqr() const { return v_[0] * v_[0] + v_[1] * v_[1] + v_[2] * v_[2]; }
    float length() const { return std::sqrt(lengthSqr()); }
    // multiply by the rounded reciprocal of the length; no small-length guard (NanoVDB.h:948-953)
    V3 &normalize() { return (*this) *= 1.0f / length(); }

  private:
    float v_[3];
};

inline V3 operator*(float s, V3 const &v) { return v * s; }
inline std::ostream &operator<<(std::ostream &s, V3 const &v) { return s << "(" << v[0] << ", " << v[1] << ", " << v[2] << ")"; }

class Ray {
  public:
    Ray() : eye_(0.0f), dir_(1.0f, 0.0f, 0.0f) {}
    Ray(V3 const &eye, V3 const &dir) : eye_(eye), dir_(dir) {}
    V3 const &eye() const { return eye_; }
    V3 const &dir() const { return dir_; }
    V3 operator()(float t) const { return eye_ + dir_ * t; }

  private:
    V3 eye_, dir_;
};

// The reference's float3 (Math.hpp:157-207) as the render loop uses it: dot / mag2 / cross / normalize with the
// RayEpsilon guard (Math.hpp:392-398), and constructBasis (Math.hpp:424-434).
struct float3 {
    float values[3];
    float3() : values{0.0f, 0.0f, 0.0f} {}
    explicit float3(float a) : values{a, a, a} {}
    float3(float x, float y, float z) : values{x, y, z} {}
    float &operator()(std::size_t i) { return values[i]; }
    float const &operator()(std::size_t i) const { return values[i]; }
    bool operator==(float3 const &o) const { return values[0] == o.values[0] && values[1] == o.values[1] && values[2] == o.values[2]; }
    bool operator!=(float3 const &o) const { return !(*this == o); }
};

inline float3 operator+(float3 const &a, float3 const &b) { return {a(0) + b(0), a(1) + b(1), a(2) + b(2)}; }
inline float3 operator-(float3 const &a, float3 const &b) { return {a(0) - b(0), a(1) - b(1), a(2) - b(2)}; }
inline float3 operator-(float3 const &a) { return {-a(0), -a(1), -a(2)}; }
inline float3 operator*(float3 const &a, float3 const &b) { return {a(0) * b(0), a(1) * b(1), a(2) * b(2)}; }
inline float3 operator*(float3 const &a, float s) { return {a(0) * s, a(1) * s, a(2) * s}; }
inline float3 operator*(float s, float3 const &a) { return {s * a(0), s * a(1), s * a(2)}; }
inline float dot(float3 const &a, float3 const &b) { return a(0) * b(0) + a(1) * b(1) + a(2) * b(2); }
inline float mag2(float3 const &a) { return dot(a, a); }
inline float3 rayT(float3 const &o, float3 const &d, float t) { return o + d * float3{t, t, t}; }
inline float3 cross(float3 const &a, float3 const &b) {
    return {a(1) * b(2) - a(2) * b(1), a(2) * b(0) - a(0) * b(2), a(0) * b(1) - a(1) * b(0)};
}
inline float3 normalize(float3 const &v) {
    float len = std::sqrt(mag2(v));
    if (isAlmostZero(len))
        return float3{0.0f};
    float s = 1.0f / len;
    return v * float3{s};
}

struct Basis {
    float3 N, T, B;
};

inline Basis constructBasis(float3 const &N) {
    float3 helper(0.0f, 1.0f, 0.0f);
    if (std::fabs(N(1)) > 0.95)
        helper = float3(0.0f, 0.0f, 1.0f);
    Basis b;
    b.N = N;
    b.T = normalize(cross(helper, N));
    b.B = cross(b.T, N);
    return b;
}

struct PixelCoord {
    using value_type = std::int32_t;
    value_type i, j;
};

// Inclusive pixel rectangle (reference Math.hpp:226-264): cannot be empty.
class PixelRect {
  public:
    using element_type = PixelCoord::value_type;

    PixelRect() : lo_{0, 0}, hi_{0, 0} {}
    explicit PixelRect(PixelCoord dims) : PixelRect(PixelCoord{0, 0}, PixelCoord{dims.i - 1, dims.j - 1}) {
        CORNELIS_EXPECTS(dims.i != 0 && dims.j != 0, "PixelRect cannot represent lines or the empty rectangle.");
    }
    PixelRect(element_type w, element_type h) : PixelRect(PixelCoord{w, h}) {}
    PixelRect(PixelCoord a, PixelCoord b)
        : lo_{std::min(a.i, b.i), std::min(a.j, b.j)}, hi_{std::max(a.i, b.i), std::max(a.j, b.j)} {}

    element_type width() const noexcept { return hi_.i - lo_.i + 1; }
    element_type height() const noexcept { return hi_.j - lo_.j + 1; }
    element_type area() const noexcept { return width() * height(); }
    PixelCoord const &min() const noexcept { return lo_; }
    PixelCoord const &max() const noexcept { return hi_; }
    bool operator==(PixelRect const &o) const noexcept {
        return lo_.i == o.lo_.i && lo_.j == o.lo_.j && hi_.i == o.hi_.i && hi_.j == o.hi_.j;
    }
    bool operator!=(PixelRect const &o) const noexcept { return !(*this == o); }

  private:
    PixelCoord lo_, hi_;
};

inline std::ostream &operator<<(std::ostream &s, PixelRect const &r) {
    return s << "PixelRect{{" << r.min().i << ", " << r.min().j << "}, {" << r.max().i << ", " << r.max().j << "}}";
}

} // namespace cornelis
