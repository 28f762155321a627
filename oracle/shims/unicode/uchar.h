// Shim for the ICU headers tokenmonster-cpp includes (ICU is not installed in this image).
// TEST INFRASTRUCTURE ONLY.  The reference only enters ICU for non-ASCII input
// (tokenmonster.cpp:191-198, :214-221 are ASCII early-outs); these stubs throw there.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
typedef int UErrorCode;
typedef std::int32_t UChar32;
#define U_ZERO_ERROR 0
#define U_FAILURE(x) ((x) > 0)
#define U_NON_SPACING_MARK 6
inline int u_charType(UChar32) { throw std::runtime_error("ICU shim: non-ASCII unsupported"); }
namespace icu {
struct StringPiece {
  StringPiece(const char*, int) {}
};
struct Locale {
  static Locale getRoot() { return {}; }
};
struct UnicodeString {
  static UnicodeString fromUTF8(StringPiece) { throw std::runtime_error("ICU shim: non-ASCII unsupported"); }
  void toLower(Locale) {}
  template <class S> void toUTF8String(S&) const {}
};
struct Normalizer2 {
  static const Normalizer2* getNFDInstance(UErrorCode&) { throw std::runtime_error("ICU shim: non-ASCII unsupported"); }
  void normalize(const UnicodeString&, UnicodeString&, UErrorCode&) const {}
};
}  // namespace icu
