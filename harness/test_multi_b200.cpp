// test_multi_b200.cpp -- the reference's caller pattern (a single-threaded C++ loop over host buffers,
// TestFileWithGPU.cpp:57-94) driving 1, 2, 4, 8 B200s through the C ABI alone: no Python, no torch, no NCCL.
//
//   test_multi_b200 pair.bin batch.bin out_prefix [max_gpus] [reps]
//     pair.bin   int64 n, int64 m, n bytes seq1, m bytes seq2         (one long pair -> in-process ring of G GPUs)
//     batch.bin  int64 npairs, int64 len1, int64 len2, npairs*len1 bytes, npairs*len2 bytes
//                                                                      (fixed-length batch -> G contiguous shards)
// For G = 1, 2, 4, 8 (<= max_gpus and the devices present): swb200_set_devices(G), then the SAME host-buffer calls
// (swb200_score_ex, swb200_score_batch, swb200_score_banded_batch when len1 == len2).  Prints one JSON line per G and
// writes the batch scores to <out_prefix>_g<G>.bin; exits 1 if any G disagrees with G = 1.  tests/test_gpu_multi.py
// checks the printed scores against the oracle.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/swb200.h"

static std::vector<unsigned char> slurp(const char* path) {
  std::vector<unsigned char> v;
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  unsigned char buf[1 << 16];
  size_t k;
  while ((k = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + k);
  fclose(f);
  return v;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define CHECK(call)                                                                      \
  do {                                                                                   \
    const int rc_ = (call);                                                              \
    if (rc_ != SWB200_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, swb200_last_error()); return 3; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: test_multi_b200 pair.bin batch.bin out_prefix [max_gpus] [reps]\n"); return 2; }
  const std::vector<unsigned char> pf = slurp(argv[1]), bf = slurp(argv[2]);
  const std::string prefix = argv[3];
  const int max_gpus = argc > 4 ? atoi(argv[4]) : 8, reps = argc > 5 ? atoi(argv[5]) : 1;
  long long n, m, npairs, len1, len2;
  memcpy(&n, pf.data(), 8); memcpy(&m, pf.data() + 8, 8);
  const unsigned char* seq1 = pf.data() + 16;
  const unsigned char* seq2 = seq1 + n;
  memcpy(&npairs, bf.data(), 8); memcpy(&len1, bf.data() + 8, 8); memcpy(&len2, bf.data() + 16, 8);
  const unsigned char* b1 = bf.data() + 24;
  const unsigned char* b2 = b1 + npairs * len1;
  std::vector<long long> off1((size_t)npairs), off2((size_t)npairs);
  std::vector<int> l1((size_t)npairs, (int)len1), l2((size_t)npairs, (int)len2);
  for (long long k = 0; k < npairs; ++k) { off1[(size_t)k] = k * len1; off2[(size_t)k] = k * len2; }

  CHECK(swb200_configure("ring_min_cells", "1"));        // every pair this program scores goes over the ring when G > 1
  const int present = swb200_device_count();
  if (present < 1) { fprintf(stderr, "no CUDA device\n"); return 3; }
  int ref_pair = -1;
  std::vector<int> ref_batch, ref_banded;
  bool ok = true;
  for (int G = 1; G <= max_gpus && G <= present; G *= 2) {
    CHECK(swb200_set_devices(G));
    int score = -1;
    double pair_ms = 1e30, batch_ms = 1e30, banded_ms = 1e30;
    swb200_run_info info;
    memset(&info, 0, sizeof info);
    for (int r = 0; r <= reps; ++r) {                     // r = 0: warm-up (context, rings, peer mappings)
      const double t0 = now_ms();
      CHECK(swb200_score_ex(seq1, n, seq2, m, nullptr, nullptr, &score));
      if (r > 0 || reps == 0) pair_ms = std::min(pair_ms, now_ms() - t0);
    }
    CHECK(swb200_last_run(nullptr, &info));
    std::vector<int> scores((size_t)npairs, -1), banded((size_t)npairs, -1);
    for (int r = 0; r <= reps; ++r) {
      const double t0 = now_ms();
      CHECK(swb200_score_batch(b1, off1.data(), l1.data(), b2, off2.data(), l2.data(), npairs, nullptr, nullptr, scores.data()));
      if (r > 0 || reps == 0) batch_ms = std::min(batch_ms, now_ms() - t0);
    }
    const bool do_banded = len1 == len2;
    if (do_banded)
      for (int r = 0; r <= reps; ++r) {
        const double t0 = now_ms();
        CHECK(swb200_score_banded_batch(b1, off1.data(), l1.data(), b2, off2.data(), l2.data(), npairs, -32, 31, nullptr, nullptr,
                                        banded.data()));
        if (r > 0 || reps == 0) banded_ms = std::min(banded_ms, now_ms() - t0);
      }
    long long checksum = 0, bchecksum = 0;
    for (long long k = 0; k < npairs; ++k) { checksum += (long long)scores[(size_t)k] * (k % 1000003 + 1); bchecksum += (long long)banded[(size_t)k] * (k % 1000003 + 1); }
    if (G == 1) { ref_pair = score; ref_batch = scores; ref_banded = banded; }
    const bool same = score == ref_pair && scores == ref_batch && banded == ref_banded;
    ok = ok && same;
    printf("{\"gpus\": %d, \"devices_in_use\": %d, \"pair_score\": %d, \"pair_ms\": %.3f, \"pair_kernel_ms\": %.3f, \"pair_gcups\": %.1f, "
           "\"two_sided\": %d, \"rebased\": %d, \"rows\": %d, \"config\": %d, \"warps\": %d, "
           "\"batch_pairs\": %lld, \"batch_checksum\": %lld, \"batch_ms\": %.3f, \"banded_checksum\": %lld, \"banded_ms\": %.3f, "
           "\"same_as_one_gpu\": %s}\n",
           G, swb200_get_devices(), score, pair_ms, info.engine_ms, (double)n * (double)m / (pair_ms * 1e6), info.two_sided, info.rebased,
           info.rows, info.config, info.warps, npairs, checksum, batch_ms, do_banded ? bchecksum : 0LL, do_banded ? banded_ms : 0.0,
           same ? "true" : "false");
    fflush(stdout);
    const std::string out = prefix + "_g" + std::to_string(G) + ".bin";
    if (FILE* f = fopen(out.c_str(), "wb")) {
      fwrite(scores.data(), sizeof(int), (size_t)npairs, f);
      if (do_banded) fwrite(banded.data(), sizeof(int), (size_t)npairs, f);
      fclose(f);
    }
  }
  CHECK(swb200_set_devices(1));
  return ok ? 0 : 1;
}
