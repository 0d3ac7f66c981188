// fm_cli.cc -- streaming command-line driver on top of the C++ adapter, with the flag names, defaults
// and output format of the reference's FuzzyMatch-cli (cli/src/FuzzyMatch-cli.cc:195-233, 292-343,
// 395-470) for `-a match`:
//
//   fm_cli -c corpus.tsv [-f 0.8] [-n 5] [--ml 3] [--mr 0.3] [-P] [-I 0] [--insert-cost 1] ... < queries > out
//
// The corpus has one sentence per line, optionally "source<TAB>target" (--add-target appends "=target"
// to the 1-based line number used as id, --add-target-no-index uses the target as id, like import_tm,
// :32-79). Each input line yields "score<TAB>id[<TAB>score<TAB>id...]" (scores printed like
// boost::lexical_cast<std::string>(float): 9 significant digits), an empty line when nothing matches,
// and "NMATCH\t<found>\t/\t<total>" goes to stderr (:227-231, 453).
//
// Differences: the text must already be tokenised and normalised (tokens are split on white space; the
// OpenNMT tokenizer / ICU front-end is out of scope, so there are no penalty tokens), and instead of
// -N worker threads the input stream is cut into windows of --batch lines that go through the GPU in
// one launch sequence each (the reference's thread pool is the thing replaced, :112-193).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "fuzzy_match_b200.hh"

static fuzzy::Tokens split_ws(const std::string& line) {
  std::istringstream is(line);
  fuzzy::Tokens t;
  std::string w;
  while (is >> w) t.push_back(w);
  return t;
}

int main(int argc, char** argv) {
  std::string corpus, action = "match", reduce = "mean";
  float fuzzy_thr = 0.8f, mr = 0.3f, idf = 0.f, ins = 1.f, del = 1.f, rep = 1.f, contrast = 0.f;
  int nmatch = 5, ml = 3, buffer = -1, device = 0;
  size_t max_tokens = fuzzy::DEFAULT_MAX_TOKENS_IN_PATTERN, batch = 8192;
  bool no_perfect = false, add_target = false, add_target_no_index = false;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto val = [&]() -> const char* {
      if (i + 1 >= argc) { std::cerr << "ERROR: missing value for " << a << std::endl; std::exit(1); }
      return argv[++i];
    };
    if (a == "-c" || a == "--corpus") corpus = val();
    else if (a == "-a" || a == "--action") action = val();
    else if (a == "-f" || a == "--fuzzy") fuzzy_thr = (float)atof(val());
    else if (a == "-n" || a == "--nmatch") nmatch = atoi(val());
    else if (a == "--ml") ml = atoi(val());
    else if (a == "--mr") mr = (float)atof(val());
    else if (a == "-P" || a == "--no-perfect") no_perfect = true;
    else if (a == "-I" || a == "--idf-penalty") idf = (float)atof(val());
    else if (a == "--insert-cost") ins = (float)atof(val());
    else if (a == "--delete-cost") del = (float)atof(val());
    else if (a == "--replace-cost") rep = (float)atof(val());
    else if (a == "--contrast") contrast = (float)atof(val());
    else if (a == "--contrast-reduce") reduce = val();
    else if (a == "--contrast-buffer") buffer = atoi(val());
    else if (a == "--max-tokens-in-pattern") max_tokens = (size_t)atol(val());
    else if (a == "--add-target") add_target = true;
    else if (a == "--add-target-no-index") add_target_no_index = true;
    else if (a == "-N" || a == "--nthreads") val();  // accepted for compatibility; batching replaces threads
    else if (a == "-p" || a == "--penalty-tokens") { if (std::string(val()) != "none") std::cerr << "WARNING: penalty tokens need the tokenizer front-end; ignored" << std::endl; }
    else if (a == "--batch") batch = (size_t)atol(val());
    else if (a == "--device") device = atoi(val());
    else if (a == "-h" || a == "--help") { std::cout << "see the header of fuzzy_match_b200/cpp/fm_cli.cc" << std::endl; return 0; }
    else { std::cerr << "ERROR: unknown option " << a << std::endl; return 1; }
  }
  if (corpus.empty()) { std::cerr << "ERROR: index file or corpus needs to be provided" << std::endl; return 3; }
  if (action != "match") { std::cerr << "ERROR: only -a match is supported" << std::endl; return 1; }
  try {
    fuzzy::FuzzyMatch fm(fuzzy::FuzzyMatch::pt_none, max_tokens, device);
    std::ifstream in(corpus);
    if (!in) { std::cerr << "ERROR: cannot open " << corpus << std::endl; return 2; }
    std::string line;
    int count = 0;
    while (std::getline(in, line)) {  // import_tm, cli/src/FuzzyMatch-cli.cc:58-75
      std::string tgt;
      const size_t pos = line.find('\t');
      if (pos != std::string::npos) { tgt = line.substr(pos + 1); line.erase(pos); }
      count++;
      std::string id = std::to_string(count);
      if (add_target) id += "=" + tgt;
      if (add_target_no_index) id = tgt;
      const fuzzy::Tokens toks = split_ws(line);
      if (toks.empty()) std::cerr << "WARNING: cannot index empty segment: " << line << " (" << id << ")" << std::endl;
      else fm.add_tm(id, toks, /*sort=*/false);
    }
    fm.sort();
    const fuzzy::EditCosts costs(ins, del, rep);
    const auto red = reduce == "max" ? fuzzy::ContrastReduce::MAX : fuzzy::ContrastReduce::MEAN;
    long found = 0, total = 0;
    std::vector<fuzzy::Tokens> window;
    auto flush = [&]() {
      if (window.empty()) return;
      std::vector<std::vector<fuzzy::FuzzyMatch::Match>> out;
      fm.match_batch(window, fuzzy_thr, (unsigned)nmatch, out, ml, mr, idf, costs, contrast, red, buffer, no_perfect);
      for (const auto& matches : out) {
        std::string text;
        for (const auto& m : matches) {
          char score[32];
          std::snprintf(score, sizeof score, "%.9g", (double)m.score);  // boost::lexical_cast<std::string>(float)
          if (!text.empty()) text += "\t";
          text += std::string(score) + "\t" + m.id;
        }
        std::cout << text << "\n";
        found += !matches.empty();
        total++;
      }
      window.clear();
    };
    while (std::getline(std::cin, line)) {
      window.push_back(split_ws(line));
      if (window.size() >= batch) flush();
    }
    flush();
    std::cout.flush();
    std::cerr << "NMATCH\t" << found << "\t/\t" << total << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "ERROR: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
