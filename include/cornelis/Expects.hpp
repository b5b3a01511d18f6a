// Precondition checking for the host API (mirrors reference include/cornelis/Expects.hpp:7-21: a failed
// precondition throws cornelis::ExpectationException).
#pragma once

#include <stdexcept>

namespace cornelis {

class ExpectationException : public std::runtime_error {
  public:
    explicit ExpectationException(char const *what) : std::runtime_error(what) {}
};

inline void expects(bool holds, char const *message) {
    if (!holds)
        throw ExpectationException(message);
}

} // namespace cornelis

#define CORNELIS_EXPECTS(pred, msg) ::cornelis::expects(static_cast<bool>(pred), msg)
