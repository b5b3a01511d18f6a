// A minimal read-only view over contiguous elements.  The reference aliases std::span (include/cornelis/Span.hpp);
// this host layer is C++17, so it carries its own.
#pragma once

#include <cstddef>
#include <vector>

namespace cornelis {

template <typename T>
class span {
  public:
    using element_type = T;
    using iterator = T *;

    constexpr span() noexcept = default;
    constexpr span(T *first, std::size_t count) noexcept : data_(first), size_(count) {}
    template <typename U, typename A>
    span(std::vector<U, A> const &v) noexcept : data_(v.data()), size_(v.size()) {}
    template <typename U, typename A>
    span(std::vector<U, A> &v) noexcept : data_(v.data()), size_(v.size()) {}

    constexpr T *data() const noexcept { return data_; }
    constexpr std::size_t size() const noexcept { return size_; }
    constexpr bool empty() const noexcept { return size_ == 0; }
    constexpr T &operator[](std::size_t k) const noexcept { return data_[k]; }
    constexpr iterator begin() const noexcept { return data_; }
    constexpr iterator end() const noexcept { return data_ + size_; }

  private:
    T *data_ = nullptr;
    std::size_t size_ = 0;
};

} // namespace cornelis
