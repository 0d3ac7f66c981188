#pragma once
#include <string>
namespace boost { namespace algorithm {
  template <class Seq>
  std::string join(const Seq& parts, const char* sep) {
    std::string out;
    bool first = true;
    for (const auto& p : parts) {
      if (!first) out += sep;
      out += p;
      first = false;
    }
    return out;
  }
} }
