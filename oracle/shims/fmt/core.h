// TEST INFRASTRUCTURE — stand-in for fmt 8.1.0; the reference only formats a thread name with it
// (reference src/Render.cpp:340): no arithmetic.
#pragma once
#include <string>
namespace fmt {
template <typename... Args>
inline std::string format(char const *pattern, Args &&...) {
    return std::string(pattern);
}
} // namespace fmt
