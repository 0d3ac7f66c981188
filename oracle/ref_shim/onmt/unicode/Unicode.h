// Stand-in for OpenNMT Tokenizer's unicode helpers (ASCII-only behaviour; text front-end is out of scope).
#pragma once
namespace onmt { namespace unicode {
  typedef unsigned int code_point_t;
  inline code_point_t utf8_to_cp(const unsigned char* s, unsigned int& l) { l = 1; return s ? *s : 0; }
  inline bool is_number(code_point_t c) { return c >= '0' && c <= '9'; }
  inline bool is_letter(code_point_t c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 128; }
} }
