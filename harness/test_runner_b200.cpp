// test_runner_b200.cpp -- the reference's CPU+GPU harness (TestFileWithGPU.cpp:46-205), re-done for libswb200.
//
// Same job: random ACGT pairs of each length, call every score function, say whether they agree, report
// timings.  What changes:
//   * inputs are seeded (splitmix-style counter generator, same stream as concurrentproject_b200/rng.py and the
//     oracle) instead of rand()/srand(time) (TestFileWithGPU.cpp:25-36), so a run can be reproduced;
//   * the success predicate really is "all scores equal" in both modes (the reference's mode-2 test is
//     mis-parenthesised, TestFileWithGPU.cpp:144);
//   * GPU time is reported as GCUPS with the kernel-only time (CUDA events, swb200_last_run) next to the
//     end-to-end wall time of the call, after one warm-up call (the reference's first GPU call pays context
//     creation inside its chrono bracket, TestFileWithGPU.cpp:81-94).
// Built by harness/Makefile.  With -DWITH_REFERENCE_CPU (oracle/Makefile target `harness_ref`, needs the reference
// sources) the reference's own CPU functions are called too, exactly like the reference harness does.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/algoGPU.h"
#include "../include/swb200.h"
#ifdef WITH_REFERENCE_CPU
#include "algoCPU.h"
#endif

static uint64_t mix64(uint64_t seed, uint64_t stream, uint64_t index) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index + 1) + 0xD1B54A32D192ED03ull * stream;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void random_sequence(uint64_t seed, uint64_t stream, int n, unsigned char* out) {
  static const char nt[4] = {'A', 'C', 'G', 'T'};
  for (int k = 0; k < n; ++k) out[k] = (unsigned char)nt[(mix64(seed, stream, (uint64_t)(k >> 5)) >> (2 * (k & 31))) & 3];
}

template <class F>
static double time_ms(F f, int* score) {
  auto t0 = std::chrono::high_resolution_clock::now();
  *score = f();
  auto t1 = std::chrono::high_resolution_clock::now();
  return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

static bool generate_test(int N, int num_tests, int mode, uint64_t seed) {
  bool success = true;
  double sum_ms[5] = {0, 0, 0, 0, 0}, sum_kernel_ms = 0;
  std::vector<unsigned char> s1(N > 0 ? N : 1), s2(N > 0 ? N : 1);
  for (int it = 0; it < num_tests; ++it) {
    random_sequence(seed, 2 * (uint64_t)it, N, s1.data());
    random_sequence(seed, 2 * (uint64_t)it + 1, N, s2.data());
    int sc[5];
    double ms[5];
    ms[0] = time_ms([&] { return SequentialSmithWatermanScoreGPU(s1.data(), s2.data(), N, N); }, &sc[0]);
    ms[1] = time_ms([&] { return SmithWatermanLazyGPU(s1.data(), s2.data(), N, N); }, &sc[1]);
    ms[2] = time_ms([&] { return SmithWatermanScoreCUDA(s1.data(), s2.data(), N, N); }, &sc[2]);
    ms[3] = time_ms([&] { return SmithDiagonalGPU(s1.data(), s2.data(), N, N); }, &sc[3]);
    ms[4] = time_ms([&] {
      int s = -1;
      swb200_options o = {};
      o.lanes = 32;   // the 32-bit kernel as an independent second opinion
      return swb200_score_ex(s1.data(), N, s2.data(), N, nullptr, &o, &s) == SWB200_OK ? s : -1;
    }, &sc[4]);
    swb200_run_info info;
    swb200_last_run(nullptr, &info);
    bool ok = sc[0] == sc[1] && sc[1] == sc[2] && sc[2] == sc[3] && sc[3] == sc[4];
#ifdef WITH_REFERENCE_CPU
    int cpu = SmithWatermanScore(s1.data(), s2.data(), N, N);
    int lazy = LazySmith(s1.data(), s2.data(), N, N);
    int thr = ParallelLazySmith_threads(s1.data(), s2.data(), N, N);
    ok = ok && cpu == sc[0] && lazy == cpu && thr == cpu;
#endif
    success = success && ok;
    for (int k = 0; k < 5; ++k) sum_ms[k] += ms[k];
    sum_kernel_ms += info.engine_ms;
    if (mode == 1)
      printf("TEST %d: score=%d %s  SimpleGPU %.3f ms  LazyGPU %.3f ms  SmithCuda %.3f ms  Diagonal %.3f ms  s32 %.3f ms\n", it,
             sc[0], ok ? "SUCCESS" : "ERROR", ms[0], ms[1], ms[2], ms[3], ms[4]);
  }
  const double cells = (double)N * N;
  printf("LENGTH: %d, NUMBER OF TESTS: %d\nSuccess: %d\n", N, num_tests, success ? 1 : 0);
  const char* names[5] = {"SimpleGPU", "LazySmithGPU", "SmithCuda", "SmithDiagonalGPU", "swb200 s32"};
  for (int k = 0; k < 5; ++k)
    printf("  %-18s: %9.3f ms/pair end to end  (%8.2f GCUPS)\n", names[k], sum_ms[k] / num_tests,
           cells / (sum_ms[k] / num_tests) / 1e6);
  printf("  %-18s: %9.3f ms/pair kernel only (%8.2f GCUPS, last variant)\n\n", "wavefront kernel", sum_kernel_ms / num_tests,
         cells / (sum_kernel_ms / num_tests) / 1e6);
  return success;
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 2;       // 1 = per-test lines, 2 = averages (TestFileWithGPU.cpp:176-183 asks on stdin)
  uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 8;   // the reference's unused default_random_engine(8)
  int num_tests = 10;                            // TestFileWithGPU.cpp:188
  std::vector<int> lengths = {1, 50, 100, 500, 1000, 1500, 2000, 2500, 3000, 3500, 4000, 4500, 5000};   // :190-192
  for (int k = 3; k < argc; ++k) { if (k == 3) lengths.clear(); lengths.push_back(atoi(argv[k])); }
  // warm-up: context creation and first-launch costs stay out of the timings
  unsigned char w1[64], w2[64];
  random_sequence(1, 0, 64, w1); random_sequence(1, 1, 64, w2);
  SmithWatermanScoreCUDA(w1, w2, 64, 64);
  bool all = true;
  for (int N : lengths) {
    printf("\n============================\nRunning tests for sequence length: %d\n============================\n", N);
    all = generate_test(N, num_tests, mode, seed) && all;
  }
  return all ? 0 : 1;
}
