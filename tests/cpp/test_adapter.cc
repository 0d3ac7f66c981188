// The reference's tokenizer-free gtest cases (test/test.cc) rewritten against the C++ adapter with
// plain asserts: same calls, same expectations. Built by __graft_entry__.build(), run on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <sstream>

#include "../../fuzzy_match_b200/cpp/fuzzy_match_b200.hh"

#define EXPECT(cond)                                                                 \
  do {                                                                               \
    if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
  } while (0)
#define EXPECT_NEAR(a, b, eps) EXPECT(std::fabs((double)(a) - (double)(b)) <= (eps))

static int failures = 0;
static fuzzy::Tokens split(const std::string& s) {
  std::istringstream is(s);
  fuzzy::Tokens t;
  std::string w;
  while (is >> w) t.push_back(w);
  return t;
}
static void add(fuzzy::FuzzyMatch& fm, const std::string& s) { fm.add_tm("", split(s), false); }

static void small_sentence_matches() {  // test/test.cc:223-262
  fuzzy::FuzzyMatch fm;
  add(fm, "single"); add(fm, "two words"); add(fm, "three kind words");
  fm.sort();
  const char* pats[] = {"single", "two words", "three kind words"};
  for (unsigned i = 0; i < 3; i++) {
    std::vector<fuzzy::FuzzyMatch::Match> matches;
    fm.match(split(pats[i]), 1, 1, matches, 3);
    EXPECT(matches.size() == 1);
    if (matches.size() == 1) EXPECT(matches[0].s_id == i);
  }
}
static void max_tokens_in_pattern() {  // test/test.cc:273-303
  fuzzy::FuzzyMatch fm(fuzzy::FuzzyMatch::pt_none, 2);
  add(fm, "single"); add(fm, "two words"); add(fm, "three kind words");
  fm.sort();
  EXPECT(fm.max_tokens_in_pattern() == 2);
  std::vector<fuzzy::FuzzyMatch::Match> matches;
  EXPECT(!fm.match({"three", "kind", "words"}, 1, 1, matches, 3));
  EXPECT(matches.size() == 0);
  fm.match({"two", "words"}, 1, 1, matches, 2);
  EXPECT(matches.size() == 1);
}
static void lcs_cost() {  // test/test.cc:337-375
  fuzzy::FuzzyMatch fm(fuzzy::FuzzyMatch::pt_none, 300);
  add(fm, "a b c"); add(fm, "a b c d e x x x"); add(fm, "x x a b c d e f x x x x x");
  fm.sort();
  std::vector<fuzzy::FuzzyMatch::Match> matches;
  fm.match({"a", "b", "c", "d", "e", "f"}, 0, 10, matches, 3, 0.5, 0, fuzzy::EditCosts(1, 0, 1));
  EXPECT(matches.size() == 3);
  if (matches.size() == 3) {
    EXPECT(matches[0].s_id == 2); EXPECT_NEAR(matches[0].score, 1.f, 1e-3);
    EXPECT(matches[1].s_id == 1); EXPECT_NEAR(matches[1].score, 5 / 6.f, 1e-3);
    EXPECT(matches[2].s_id == 0); EXPECT_NEAR(matches[2].score, 1 / 2.f, 1e-3);
    EXPECT(matches[0].length == 13 && matches[0].s[2] == matches[2].s[0]);  // Match::s borrows the TM sentence
  }
}
static void pre_reject() {  // test/test.cc:377-418
  fuzzy::FuzzyMatch fm;
  add(fm, "a b c d e"); add(fm, "a b c d e f"); add(fm, "a b c d e f g");
  fm.sort();
  std::vector<fuzzy::FuzzyMatch::Match> m1, m2;
  fm.match({"a", "b", "c"}, 0.5, 10, m1, 0, 0, 0, fuzzy::EditCosts(1, 1, 1));
  EXPECT(m1.size() == 2);
  fm.match(split("a b c d e f g h i j k l"), 0.5, 10, m2, 0, 0, 0, fuzzy::EditCosts(1, 1, 1));
  EXPECT(m2.size() == 2);
}
static void idf_weight() {  // test/test.cc:420-507
  {
    fuzzy::FuzzyMatch fm;
    add(fm, "a b c"); add(fm, "a b d"); add(fm, "d d d d d"); add(fm, "d e"); add(fm, "c");
    fm.sort();
    std::vector<fuzzy::FuzzyMatch::Match> m;
    fm.match(split("a b c d"), 0., 10, m, 0, 0, 1, fuzzy::EditCosts(1, 0, 1));
    EXPECT(m.size() == 2);
    if (m.size() == 2) {
      EXPECT(m[0].s_id == 0 && m[1].s_id == 1);
      EXPECT_NEAR(m[0].score, 0.6706515, 1e-4); EXPECT_NEAR(m[1].score, 0.6076691, 1e-4);
    }
  }
  for (int unit = 0; unit < 2; unit++) {
    fuzzy::FuzzyMatch fm;
    add(fm, "a b c e"); add(fm, "a b e d"); add(fm, "d d d d d"); add(fm, "d e"); add(fm, "c");
    fm.sort();
    std::vector<fuzzy::FuzzyMatch::Match> m;
    fm.match(split("a b c d"), 0., 10, m, 0, 0, 1, unit ? fuzzy::EditCosts(1, 1, 1) : fuzzy::EditCosts(1, 0, 1));
    EXPECT(m.size() == 2);
    if (m.size() == 2) {
      EXPECT(m[0].s_id == 0 && m[1].s_id == 1);
      EXPECT_NEAR(m[0].score, 0.6706515, 1e-4); EXPECT_NEAR(m[1].score, 0.6076691, 1e-4);
    }
  }
}
static void contrastive() {  // test/test.cc:509-632
  for (int mx = 0; mx < 2; mx++) {
    fuzzy::FuzzyMatch fm;
    add(fm, "a b c d"); add(fm, "b c d"); add(fm, "d e f");
    fm.sort();
    std::vector<fuzzy::FuzzyMatch::Match> m;
    fm.match(split("a b c d e f"), 0, 10, m, 0, 0, 0, fuzzy::EditCosts(1, 1, 1), 1., mx ? fuzzy::ContrastReduce::MAX : fuzzy::ContrastReduce::MEAN);
    EXPECT(m.size() == 3);
    if (m.size() == 3) {
      EXPECT(m[0].s_id == 0); EXPECT_NEAR(m[0].score - m[0].penalty, 2 / 3.f, 1e-3);
      EXPECT(m[1].s_id == 2); EXPECT_NEAR(m[1].score - m[1].penalty, 1 / 2.f, 1e-3);
      EXPECT(m[2].s_id == 1); EXPECT_NEAR(m[2].score - m[2].penalty, mx ? -1 / 4.f : 1 / 8.f, 1e-3);
    }
  }
  fuzzy::FuzzyMatch fm;
  add(fm, "a b c d e"); add(fm, "b c d e"); add(fm, "c d e f"); add(fm, "d e f g"); add(fm, "h i j");
  fm.sort();
  std::vector<fuzzy::FuzzyMatch::Match> m;
  fm.match(split("a b c d e f g h i j"), 0, 3, m, 0, 0, 0, fuzzy::EditCosts(1, 0, 1), 1., fuzzy::ContrastReduce::MAX, 10);
  EXPECT(m.size() == 3);
  if (m.size() == 3) EXPECT(m[0].s_id == 0 && m[1].s_id == 3 && m[2].s_id == 4);
}
static void batch_and_append() {
  fuzzy::FuzzyMatch fm;
  add(fm, "a b c d e"); add(fm, "a b c d e f"); add(fm, "x y z");
  fm.sort();
  std::vector<std::vector<fuzzy::FuzzyMatch::Match>> out;
  fm.match_batch({split("a b c d e"), split("x y z"), split("q q q"), {}}, 0.5, 2, out);
  EXPECT(out.size() == 4 && out[0].size() == 2 && out[1].size() == 1 && out[2].empty() && out[3].empty());
  if (out[0].size() == 2) EXPECT(out[0][0].s_id == 0 && out[0][0].score == 1.f && out[0][0].id == "");
  // match() appends to what is already there and stops at number_of_matches entries IN TOTAL
  // (reference src/fuzzy_match.cc:670-679: `matches.size() < number_of_matches`)
  std::vector<fuzzy::FuzzyMatch::Match> m(1);
  EXPECT(fm.match(split("x y z"), 0.5, 1, m));
  EXPECT(m.size() == 1);
  EXPECT(fm.match(split("x y z"), 0.5, 2, m));
  EXPECT(m.size() == 2 && m[1].s_id == 2);
  EXPECT(fm.match(split("a b c d e"), 0.5, 0, m));  // 0 = all
  EXPECT(m.size() == 4);
}

static void contrastive_into_nonempty_vector() {
  // The contrastive rerank penalises the candidates against entries that are in `matches` before the call
  // (reference src/fuzzy_match.cc:634-652). Expected values: the live reference on the same two calls
  // (oracle/ref_driver.cc fmref_match_batch_twice).
  fuzzy::FuzzyMatch fm;
  add(fm, "a b c d e"); add(fm, "a b c d e f"); add(fm, "x y z"); add(fm, "a b c x y");
  fm.sort();
  std::vector<fuzzy::FuzzyMatch::Match> m;
  EXPECT(fm.match(split("a b c d e"), 0.9f, 1, m));
  EXPECT(m.size() == 1 && m[0].s_id == 0);
  std::vector<fuzzy::FuzzyMatch::Match> m3 = m;
  EXPECT(fm.match(split("a b c d"), 0.3f, 0, m, 2, 0, 0, fuzzy::EditCosts(), 0.5f));
  EXPECT(m.size() == 4);
  if (m.size() == 4) {
    EXPECT(m[1].s_id == 0); EXPECT_NEAR(m[1].score, 0.8f, 1e-6); EXPECT_NEAR(m[1].penalty, 1.0f, 1e-6);
    EXPECT(m[2].s_id == 3); EXPECT_NEAR(m[2].score, 0.6f, 1e-6); EXPECT_NEAR(m[2].penalty, 0.6f, 1e-6);  // ahead of the better-scoring s1
    EXPECT(m[3].s_id == 1); EXPECT_NEAR(m[3].score, 0.6666f, 1e-6); EXPECT_NEAR(m[3].penalty, 0.7222f, 1e-6);
  }
  EXPECT(fm.match(split("a b c d"), 0.3f, 3, m3, 2, 0, 0, fuzzy::EditCosts(), 0.5f, fuzzy::ContrastReduce::MAX));
  EXPECT(m3.size() == 3);  // the earlier entry counts against number_of_matches
  if (m3.size() == 3) { EXPECT(m3[1].s_id == 0 && m3[2].s_id == 3); EXPECT_NEAR(m3[2].penalty, 0.6f, 1e-6); }
}

static void sentence_api() {
  // Real tokens, case class and penalty tokens through add_tm(id, Sentence, Tokens) / match(Sentence, ...).
  // Expected scores come from the unmodified reference on the same inputs (tests/golden/make_golden.py
  // conventions): s0 differs from q0 by a case-class real form (+1) and a "," itok (+1): 0.98.
  fuzzy::FuzzyMatch fm;
  fuzzy::Sentence s0({"Lower-a", "b", "c"});
  s0.set_itok(1, ",");
  fm.add_tm("s0", s0, {"a", "b", "c"}, false);
  fuzzy::Sentence s1({"a", "B2", "c", "d"});
  s1.set_itok(4, ".");
  fm.add_tm("s1", s1, {"a", "b", "c", "d"}, false);
  fm.sort();
  {
    std::vector<fuzzy::FuzzyMatch::Match> m;
    fm.match(fuzzy::Sentence(fuzzy::Tokens{"a", "b", "c"}), {"a", "b", "c"}, 0.f, 4, false, m, 2, 0.f);
    EXPECT(m.size() == 2);
    if (m.size() == 2) {
      EXPECT(m[0].s_id == 0 && m[0].id == "s0"); EXPECT_NEAR(m[0].score, 0.98f, 1e-6);
      EXPECT(m[1].s_id == 1); EXPECT_NEAR(m[1].score, 0.72f, 1e-6);
    }
  }
  {
    fuzzy::Sentence q1({"a", "b", "c", "d"});
    q1.set_itok(2, " ");
    std::vector<fuzzy::FuzzyMatch::Match> m;
    fm.match(q1, {"a", "b", "c", "d"}, 0.f, 4, false, m, 2, 0.f);
    EXPECT(m.size() == 2);
    if (m.size() == 2) {
      EXPECT(m[0].s_id == 1); EXPECT_NEAR(m[0].score, 0.96f, 1e-6);
      EXPECT(m[1].s_id == 0); EXPECT_NEAR(m[1].score, 0.72f, 1e-6);
    }
  }
}

static void subsequence() {  // no gtest in the reference calls subsequence(): answers taken from the live reference (oracle/ref_driver.cc)
  fuzzy::FuzzyMatch fm;
  fm.add_tm("id0", split("the quick brown fox jumps over the lazy dog"), false);
  fm.add_tm("id1", split("a quick brown dog"), false);
  fm.add_tm("id2", split("lazy dog sleeps all day"), false);
  fm.add_tm("id3", split("the quick brown fox"), false);
  fm.sort();
  {
    std::vector<fuzzy::FuzzyMatch::Match> m;
    EXPECT(fm.subsequence(split("my quick brown fox sleeps all day"), 1, false, m));
    EXPECT(m.size() == 1);
    if (m.size() == 1) { EXPECT(m[0].s_id == 3); EXPECT_NEAR(m[0].score, 0.4285f, 1e-6); EXPECT(m[0].max_subseq == 3); EXPECT(m[0].id == "id3\tquick brown fox"); }
  }
  {
    std::vector<fuzzy::FuzzyMatch::Match> m;
    EXPECT(fm.subsequence(split("the quick brown fox"), 2, true, m));  // the perfect match is skipped
    if (m.size() == 1) { EXPECT(m[0].s_id == 0); EXPECT_NEAR(m[0].score, 0.4444f, 1e-6); EXPECT(m[0].max_subseq == 4); EXPECT(m[0].id == "id0\tthe quick brown fox"); }
    m.clear();
    EXPECT(fm.subsequence(split("the quick brown fox"), 2, false, m));
    if (m.size() == 1) { EXPECT(m[0].s_id == 3); EXPECT_NEAR(m[0].score, 1.0f, 1e-6); }
  }
  {
    std::vector<fuzzy::FuzzyMatch::Match> m;
    EXPECT(fm.subsequence(split("lazy dog sleeps"), 3, false, m, 2, 0.f, true));
    if (m.size() == 1) { EXPECT(m[0].s_id == 2); EXPECT_NEAR(m[0].score, 0.6f, 1e-6); EXPECT(m[0].max_subseq == 3); }
    m.clear();
    EXPECT(!fm.subsequence(split("nothing here matches"), 1, false, m));
    EXPECT(m.empty());
  }
}

int main() {
  subsequence();
  small_sentence_matches();
  max_tokens_in_pattern();
  lcs_cost();
  pre_reject();
  idf_weight();
  contrastive();
  batch_and_append();
  contrastive_into_nonempty_vector();
  sentence_api();
  std::printf(failures ? "%d FAILURES\n" : "all adapter tests passed\n", failures);
  return failures ? 1 : 0;
}
