// Stand-in for ICU's Normalizer2: identity NFC. Only the text front-end (out of scope) touches it.
#pragma once
#include <string>
typedef int UErrorCode;
#define U_ZERO_ERROR 0
#define U_FAILURE(x) ((x) > 0)
namespace icu {
  class UnicodeString {
  public:
    static UnicodeString fromUTF8(const std::string& s) { UnicodeString u; u._s = s; return u; }
    std::string& toUTF8String(std::string& out) const { out += _s; return out; }
    std::string _s;
  };
  class Normalizer2 {
  public:
    static const Normalizer2* getNFCInstance(UErrorCode&) { static Normalizer2 n; return &n; }
    UnicodeString normalize(const UnicodeString& s, UErrorCode&) const { return s; }
  };
}
