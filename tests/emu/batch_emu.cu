// batch_emu.cu -- runs the product's batch kernel body (csrc/swb_batch.cuh) and banded kernel body
// (csrc/swb_banded.cuh) on CPU threads, one pthread per lane (see engine_emu.cu).  TEST TOOL ONLY.
//
// usage: batch_emu pairs.bin KIND R MODE G [match mismatch gap_init gap_ext [band_lo]]
//   KIND = batch | banded;  pairs.bin: int32 npairs, then per pair int32 len1, int32 len2, len1 bytes, len2 bytes.
//   batch: the shorter sequence of a pair is striped (as launch_pack_batch does); banded: seq1 = columns, seq2 = rows.
//   prints "scores s0 s1 ..."
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>
#include "../../concurrentproject_b200/csrc/swb_batch.cuh"
#include "../../concurrentproject_b200/csrc/swb_banded.cuh"

using namespace swb;

#if !SWB_DEVICE_CODE

static void pack(const std::vector<uint8_t>& s, uint64_t* words) {
  for (size_t k = 0; k < s.size(); ++k) words[k >> 5] |= (uint64_t)((s[k] >> 1) & 3) << (2 * (k & 31));
}

typedef void (*batch_fn)(const BatchParams*, WarpShared*, int, long long, long long, BatchWarpSmem*);
template <int R, int MODE, int G>
static void run_batch_lane(const BatchParams* P, WarpShared* ws, int lane, long long wid, long long nw, BatchWarpSmem* sm) {
  WarpCtx w{lane, ws};
  batch_warp<R, MODE, G>(*P, w, wid, nw, sm);
}
static batch_fn pick_batch(int R, int mode, int G) {
#define C(RR, GG) if (R == RR && G == GG) return mode ? run_batch_lane<RR, 1, GG> : run_batch_lane<RR, 0, GG>;
  C(2, 8) C(4, 8) C(10, 8) C(4, 16) C(2, 32)
#undef C
  return nullptr;
}
template <int MODE>
static void run_banded_lane(const BandedParams* P, WarpShared* ws, int lane, long long wid, long long nw, BandedWarpSmem* sm) {
  WarpCtx w{lane, ws};
  banded_warp<MODE>(*P, w, wid, nw, sm);
}

template <int MODE>
static void run_banded8_lane(const BandedParams* P, WarpShared* ws, int lane, long long wid, long long nw, BandedWarpSmem8* sm) {
  WarpCtx w{lane, ws};
  banded_warp8<MODE>(*P, w, wid, nw, sm);
}

template <int MODE>
static void run_banded4_lane(const BandedParams* P, WarpShared* ws, int lane, long long wid, long long nw, BandedWarpSmemS<8>* sm) {
  WarpCtx w{lane, ws};
  banded_warpS<MODE, 8>(*P, w, wid, nw, sm);
}

template <int MODE>
static void run_banded2_lane(const BandedParams* P, WarpShared* ws, int lane, long long wid, long long nw, BandedWarpSmemS<16>* sm) {
  WarpCtx w{lane, ws};
  banded_warpS<MODE, 16>(*P, w, wid, nw, sm);
}

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: batch_emu pairs.bin batch|banded R MODE G [ma mi gi ge [band_lo]]\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  const bool banded = !strcmp(argv[2], "banded");
  const int R = atoi(argv[3]), mode = atoi(argv[4]), G = atoi(argv[5]);
  const int ma = argc > 6 ? atoi(argv[6]) : 1, mi = argc > 7 ? atoi(argv[7]) : -1, gi = argc > 8 ? atoi(argv[8]) : 1,
            ge = argc > 9 ? atoi(argv[9]) : 1, band_lo = argc > 10 ? atoi(argv[10]) : -32;
  int32_t np = 0;
  if (fread(&np, 4, 1, f) != 1) return 2;
  std::vector<std::vector<uint8_t>> s1(np), s2(np);
  int max_q = 0, max_t = 0;
  for (int k = 0; k < np; ++k) {
    int32_t l[2];
    if (fread(l, 4, 2, f) != 2) return 2;
    s1[k].resize(l[0]); s2[k].resize(l[1]);
    if (l[0] && fread(s1[k].data(), 1, l[0], f) != (size_t)l[0]) return 2;
    if (l[1] && fread(s2[k].data(), 1, l[1], f) != (size_t)l[1]) return 2;
    if (!banded && s1[k].size() > s2[k].size()) std::swap(s1[k], s2[k]);      // Q = the shorter one
    max_q = std::max(max_q, (int)s1[k].size()); max_t = std::max(max_t, (int)s2[k].size());
  }
  fclose(f);
  const long long qs = std::max(1, (max_q + 31) / 32) + 2, ts = std::max(1, (max_t + 31) / 32) + 2;
  std::vector<uint64_t> qw((size_t)np * qs + 4, 0), tw((size_t)np * ts + 4, 0);
  std::vector<int> ql(np), tl(np), scores(np, -1);
  for (int k = 0; k < np; ++k) { pack(s1[k], &qw[(size_t)k * qs]); pack(s2[k], &tw[(size_t)k * ts]); ql[k] = (int)s1[k].size(); tl[k] = (int)s2[k].size(); }
  const int W = 2;                                     // two emulated warps share the pairs
  std::vector<WarpShared> ws(W);
  for (auto& x : ws) pthread_barrier_init(&x.bar, nullptr, 32);
  std::vector<std::thread> th;
  if (!banded) {
    batch_fn fn = pick_batch(R, mode, G);
    if (!fn) { fprintf(stderr, "unsupported R/G\n"); return 2; }
    BatchParams P{};
    P.q_words = qw.data(); P.t_words = tw.data(); P.q_len = ql.data(); P.t_len = tl.data(); P.q_stride = qs; P.t_stride = ts;
    P.npairs = np; P.scores = scores.data(); P.match = ma; P.mismatch = mi; P.gap_init = gi; P.gap_ext = ge;
    std::vector<BatchWarpSmem> sm(W);
    for (int w = 0; w < W; ++w) for (int l = 0; l < 32; ++l) th.emplace_back(fn, &P, &ws[w], l, (long long)w, (long long)W, &sm[w]);
    for (auto& x : th) x.join();
  } else {
    BandedParams P{};
    P.a_words = qw.data(); P.b_words = tw.data(); P.a_len = ql.data(); P.b_len = tl.data(); P.a_stride = qs; P.b_stride = ts;
    P.npairs = np; P.band_lo = band_lo; P.scores = scores.data(); P.match = ma; P.mismatch = mi; P.gap_init = gi; P.gap_ext = ge;
    if (G == 16) {                                     // the 16-threads-per-pair layout
      std::vector<BandedWarpSmem> sm(W);
      auto fn = mode ? run_banded_lane<1> : run_banded_lane<0>;
      for (int w = 0; w < W; ++w) for (int l = 0; l < 32; ++l) th.emplace_back(fn, &P, &ws[w], l, (long long)w, (long long)W, &sm[w]);
      for (auto& x : th) x.join();
    } else if (G == 2) {                               // two threads per pair, sixteen register sets each
      std::vector<BandedWarpSmemS<16>> sm(W);
      auto fn = mode ? run_banded2_lane<1> : run_banded2_lane<0>;
      for (int w = 0; w < W; ++w) for (int l = 0; l < 32; ++l) th.emplace_back(fn, &P, &ws[w], l, (long long)w, (long long)W, &sm[w]);
      for (auto& x : th) x.join();
    } else if (G == 4) {                               // four threads per pair, eight register sets each
      std::vector<BandedWarpSmemS<8>> sm(W);
      auto fn = mode ? run_banded4_lane<1> : run_banded4_lane<0>;
      for (int w = 0; w < W; ++w) for (int l = 0; l < 32; ++l) th.emplace_back(fn, &P, &ws[w], l, (long long)w, (long long)W, &sm[w]);
      for (auto& x : th) x.join();
    } else {                                           // G = 8: eight threads per pair, four register sets each
      std::vector<BandedWarpSmem8> sm(W);
      auto fn = mode ? run_banded8_lane<1> : run_banded8_lane<0>;
      for (int w = 0; w < W; ++w) for (int l = 0; l < 32; ++l) th.emplace_back(fn, &P, &ws[w], l, (long long)w, (long long)W, &sm[w]);
      for (auto& x : th) x.join();
    }
  }
  printf("scores");
  for (int k = 0; k < np; ++k) printf(" %d", scores[k]);
  printf("\n");
  return 0;
}
#else
int main() { return 0; }
#endif
