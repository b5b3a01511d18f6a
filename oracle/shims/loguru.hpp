// TEST INFRASTRUCTURE — no-op stand-in for loguru 2.1.0 as used by the reference's Render.cpp
// (reference src/Render.cpp:315-361): logging only, no arithmetic.
#pragma once
#define LOG_F(...) ((void)0)
#define LOG_SCOPE_F(...) ((void)0)
namespace loguru {
inline void set_thread_name(char const *) {}
} // namespace loguru
