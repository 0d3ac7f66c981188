// Stand-in for onmt::Tokenizer: whitespace split, no features. The match(Tokens)/add_tm(Tokens)
// path that the oracle drives never calls it; it only has to satisfy the reference's declarations.
#pragma once
#include <sstream>
#include <string>
#include <vector>
namespace onmt {
  class Tokenizer {
  public:
    enum class Mode { Conservative, Aggressive, Char, Space, None };
    enum Flags {
      None = 0, CaseFeature = 1 << 0, JoinerAnnotate = 1 << 1, JoinerNew = 1 << 2,
      SpacerAnnotate = 1 << 6, SpacerNew = 1 << 7, NoSubstitution = 1 << 9,
      SegmentAlphabetChange = 1 << 12, SupportPriorJoiners = 1 << 13
    };
    static inline const std::string joiner_marker = "\xEF\xBF\xAD";
    static inline const std::string spacer_marker = "\xE2\x96\x81";
    static inline const std::string ph_marker_open = "\xEF\xBD\x9F";
    static inline const std::string ph_marker_close = "\xEF\xBD\xA0";
    Tokenizer(Mode, int, const std::string&) {}
    void add_alphabet_to_segment(const std::string&) {}
    static bool is_placeholder(const std::string& t) { return t.find(ph_marker_open) != std::string::npos; }
    void tokenize(const std::string& text, std::vector<std::string>& words,
                  std::vector<std::vector<std::string>>& features) const {
      std::istringstream is(text);
      std::string w;
      while (is >> w) words.push_back(w);
      features.clear();
    }
    std::string detokenize(const std::vector<std::string>& words,
                           const std::vector<std::vector<std::string>>&) const {
      std::string out;
      for (size_t i = 0; i < words.size(); i++) { if (i) out += " "; out += words[i]; }
      return out;
    }
  };
}
