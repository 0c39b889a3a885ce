// gen_mt_pairs.cpp -- regenerates the reference's seeded test inputs (cudaSmithM.cu:200-213):
// std::mt19937_64 rng(12345); std::uniform_int_distribution<int>(0,3); all NUM_EXAMPLES pairs are
// drawn up front, seq1[i] before seq2[i].  The draw order depends on libstdc++'s distribution
// algorithm, so the OUTPUT of this program is frozen into tests/golden/mt12345_L*.txt; this
// source is kept only to document how the fixture was made (tests/golden/make_golden.py runs it).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
int main(int argc, char** argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 32;
  const int NUM = 10;
  std::mt19937_64 rng(12345);
  std::uniform_int_distribution<int> d(0, 3);
  const char nt[4] = {'A', 'C', 'G', 'T'};
  std::vector<std::string> a(NUM, std::string(L, 'A')), b(NUM, std::string(L, 'A'));
  for (int ex = 0; ex < NUM; ++ex)
    for (int i = 0; i < L; ++i) { a[ex][i] = nt[d(rng)]; b[ex][i] = nt[d(rng)]; }
  for (int ex = 0; ex < NUM; ++ex) printf("%s %s\n", a[ex].c_str(), b[ex].c_str());
  return 0;
}
