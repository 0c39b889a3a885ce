// ref_shim.cpp -- C-linkage doorway onto the UNMODIFIED reference CPU functions.
//
// TEST INFRASTRUCTURE ONLY.  This file contains no algorithm: it includes the reference's own
// header (algoCPU.h:4-12, found with -I/root/reference at build time) and forwards to the
// reference's own objects, which oracle/Makefile compiles where they lie under /root/reference
// (main.cpp, lazySmith.cpp, lazySmith_parallel_threads.cpp) into oracle/_ref/libref.so.
// ctypes cannot call C++-mangled names portably, hence the extern "C" forwarders.
#include <vector>
#include <thread>
#include "algoCPU.h"

extern "C" {

int ref_SmithWatermanScore(const unsigned char* seq1, const unsigned char* seq2, int n, int m) {
  return SmithWatermanScore(const_cast<unsigned char*>(seq1), const_cast<unsigned char*>(seq2), n, m);
}

int ref_LazySmith(const unsigned char* seq1, const unsigned char* seq2, int n, int m) {
  return LazySmith(const_cast<unsigned char*>(seq1), const_cast<unsigned char*>(seq2), n, m);
}

// default arguments exactly as the harness calls it (TestFileWithGPU.cpp:76)
int ref_ParallelLazySmith_threads(const unsigned char* seq1, const unsigned char* seq2, int n, int m) {
  return ParallelLazySmith_threads(const_cast<unsigned char*>(seq1), const_cast<unsigned char*>(seq2), n, m);
}

// Run `count` independent reference calls, one std::thread per worker ("pairs across cores",
// SURVEY.md 8d): the fair multi-core figure for batch workloads.  fn: 0 = SmithWatermanScore,
// 1 = LazySmith, 2 = ParallelLazySmith_threads.
void ref_batch(int fn, const unsigned char* seq1_all, const long long* off1, const int* len1,
               const unsigned char* seq2_all, const long long* off2, const int* len2,
               long long count, int workers, int* scores) {
  if (workers < 1) workers = 1;
  std::vector<std::thread> pool;
  for (int w = 0; w < workers; ++w) {
    pool.emplace_back([=]() {
      for (long long k = w; k < count; k += workers) {
        unsigned char* a = const_cast<unsigned char*>(seq1_all + off1[k]);
        unsigned char* b = const_cast<unsigned char*>(seq2_all + off2[k]);
        int s;
        if (fn == 0) s = SmithWatermanScore(a, b, len1[k], len2[k]);
        else if (fn == 1) s = LazySmith(a, b, len1[k], len2[k]);
        else s = ParallelLazySmith_threads(a, b, len1[k], len2[k]);
        scores[k] = s;
      }
    });
  }
  for (auto& t : pool) t.join();
}

int ref_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
