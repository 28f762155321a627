// Shim for <capcode/capcode.hpp> (the capcode-cpp package is an external dependency of
// tokenmonster-cpp that is not vendored in the reference).  TEST INFRASTRUCTURE ONLY.
// Lets oracle/_ref compile the reference's own C++ runtime unmodified; the level-2
// capcode itself is the oracle's restatement of javascript/tokenmonster.js:900-1065
// (tm_oracle.c).  Level 1 (NoCapcode) has no specification anywhere in the reference.
#pragma once
#include <cstdint>
#include <span>
#include <stdexcept>
#include <vector>

#include "tm_oracle.h"

namespace capcode {

using Bytes = std::vector<std::uint8_t>;

inline Bytes encode(std::span<const std::uint8_t> in) {
  std::uint8_t* out = nullptr;
  std::size_t n = 0;
  if (tmo_capcode_encode(in.data(), in.size(), &out, &n) != TMO_OK)
    throw std::runtime_error("capcode shim: non-ASCII input is not supported");
  Bytes b(out, out + n);
  tmo_free(out);
  return b;
}

inline Bytes no_capcode_encode(std::span<const std::uint8_t>) {
  throw std::runtime_error("capcode shim: level 1 (NoCapcode) is unspecified");
}

class Decoder {
 public:
  Bytes decode(const Bytes& in) {
    Bytes out(in.size() + 1);
    out.resize(tmo_capcode_decode(&st_, in.data(), in.size(), out.data()));
    return out;
  }
  Bytes no_capcode_decode(const Bytes&) {
    throw std::runtime_error("capcode shim: level 1 (NoCapcode) is unspecified");
  }

 private:
  tmo_capcode_decoder st_{0, 0, 0, 0};
};

inline Bytes decode(const Bytes& in) { return Decoder().decode(in); }
inline Bytes no_capcode_decode(const Bytes& in) { return Decoder().no_capcode_decode(in); }

}  // namespace capcode
