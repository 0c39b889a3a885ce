// engine_emu.cu -- runs the product's wavefront engine body (csrc/swb_engine.cuh) on CPU threads.
//
// TEST TOOL ONLY.  There is no GPU in the build container, so the engine body is written
// __host__ __device__ and this program executes it with one pthread per lane, a barrier at every
// warp-synchronous point and real concurrency between warps (and between emulated GPUs), which
// exercises the boundary hand-off protocol (tags, ring laps, back-pressure, ring wrap-around,
// multi-GPU ring) exactly as the kernels do.  tests/test_engine_emu.py compares the printed score
// with the oracle.  Nothing here is linked into libswb200.so.
//
// usage: engine_emu qfile tfile R MODE SLACK warps_per_gpu gpus epoch [match mismatch gap_init gap_ext [link_len]]
//        (qfile/tfile: raw ASCII bytes over {A,C,G,T})
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>
#include "../../concurrentproject_b200/csrc/swb_engine.cuh"

using namespace swb;

#if !SWB_DEVICE_CODE   // host pass only; the device pass of nvcc sees an empty translation unit

static std::vector<uint8_t> read_file(const char* path) {
  std::vector<uint8_t> v;
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  uint8_t buf[65536]; size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + n);
  fclose(f);
  return v;
}

// SLACK 1 = launch config 1, SLACK 0 = launch config 2 (the product pairs SLACK 0 / 8 warps per CTA with the
// long-chain row loop); EMU_SHORT=0/1 overrides the row-loop flavour.
static int g_short = -1, g_hs = 0;
template <int R, int MODE, int SLACK>
static void run_lane(const EngineParams* P, WarpShared* ws, int lane, int lw, WarpSmem* sm) {
  WarpCtx w{lane, ws};
  const bool sh = g_short < 0 ? SLACK == 1 : g_short != 0;
  if constexpr (MODE != 2 && MODE != 6 && MODE != 8) {
    if (g_hs) {            // EMU_HS=1: slack for the hand-off inside a thread as well (launch config 4)
      if (MODE >= 3) engine_warp_s16<R, (MODE >= 3 ? MODE - 3 : 0), SLACK, true, true, 1>(*P, w, lw, sm);
      else engine_warp_s16<R, (MODE < 2 ? MODE : 0), SLACK, false, true, 1>(*P, w, lw, sm);
      return;
    }
  }
  if (MODE == 2) { if (sh) engine_warp_s32<R, SLACK, false, true>(*P, w, lw, sm); else engine_warp_s32<R, SLACK, false, false>(*P, w, lw, sm); }
  else if (MODE == 6) { if (sh) engine_warp_s32<R, SLACK, false, true, true>(*P, w, lw, sm); else engine_warp_s32<R, SLACK, false, false, true>(*P, w, lw, sm); }
  else if (MODE == 8) { if (sh) engine_warp_s32<R, SLACK, false, true, true, true>(*P, w, lw, sm); else engine_warp_s32<R, SLACK, false, false, true, true>(*P, w, lw, sm); }
  else if (MODE >= 3) {
    if (sh) engine_warp_s16<R, (MODE >= 3 ? MODE - 3 : 0), SLACK, true, true>(*P, w, lw, sm);
    else engine_warp_s16<R, (MODE >= 3 ? MODE - 3 : 0), SLACK, true, false>(*P, w, lw, sm);
  } else {
    if (sh) engine_warp_s16<R, (MODE < 2 ? MODE : 0), SLACK, false, true>(*P, w, lw, sm);
    else engine_warp_s16<R, (MODE < 2 ? MODE : 0), SLACK, false, false>(*P, w, lw, sm);
  }
}

typedef void (*lane_fn)(const EngineParams*, WarpShared*, int, int, WarpSmem*);

template <int R>
static lane_fn pick2(int mode, int slack) {
  if (mode == 0) return slack ? run_lane<R, 0, 1> : run_lane<R, 0, 0>;
  if (mode == 1) return slack ? run_lane<R, 1, 1> : run_lane<R, 1, 0>;
  if (mode == 3) return slack ? run_lane<R, 3, 1> : run_lane<R, 3, 0>;
  if (mode == 4) return slack ? run_lane<R, 4, 1> : run_lane<R, 4, 0>;
  if (mode == 6) return slack ? run_lane<R, 6, 1> : run_lane<R, 6, 0>;      // 32-bit lanes + end cell
  if (mode == 8) return slack ? run_lane<R, 8, 1> : run_lane<R, 8, 0>;      // anchored recurrence + position of the maximum
  return slack ? run_lane<R, 2, 1> : run_lane<R, 2, 0>;
}
static lane_fn pick(int R, int mode, int slack) {
  switch (R) {
    case 1: return pick2<1>(mode, slack);
    case 2: return pick2<2>(mode, slack);
    case 3: return pick2<3>(mode, slack);
    case 4: return pick2<4>(mode, slack);
    case 8: return pick2<8>(mode, slack);
    default: return nullptr;
  }
}

int main(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage: engine_emu qfile tfile R MODE SLACK warps gpus epoch [ma mi gi ge [link_len]]\n"); return 2; }
  std::vector<uint8_t> qa = read_file(argv[1]), ta = read_file(argv[2]);
  long long LQ = (long long)qa.size(), LT = (long long)ta.size();
  int R = atoi(argv[3]), mode = atoi(argv[4]), slack = atoi(argv[5]), W = atoi(argv[6]), G = atoi(argv[7]);
  unsigned epoch = (unsigned)atoi(argv[8]);
  int ma = argc > 9 ? atoi(argv[9]) : 1, mi = argc > 10 ? atoi(argv[10]) : -1, gi = argc > 11 ? atoi(argv[11]) : 1,
      ge = argc > 12 ? atoi(argv[12]) : 1;
  long long link_len = argc > 13 ? atoll(argv[13]) : 4096;
  if (getenv("EMU_SHORT")) g_short = atoi(getenv("EMU_SHORT"));
  if (getenv("EMU_HS")) g_hs = atoi(getenv("EMU_HS"));
  lane_fn fn = pick(R, mode, slack);
  if (!fn) { fprintf(stderr, "unsupported R\n"); return 2; }

  // ACGT -> 2-bit codes exactly as the product's encode kernel does: (c >> 1) & 3
  std::vector<uint8_t> q(LQ + 1);
  for (long long k = 0; k < LQ; ++k) q[k] = qa[k] == 'N' ? (uint8_t)4 : (uint8_t)((qa[k] >> 1) & 3);   // 'N' = a row that matches nothing
  std::vector<uint64_t> t((LT + 31) / 32 + 1, 0);
  for (long long k = 0; k < LT; ++k) t[k >> 5] |= (uint64_t)((ta[k] >> 1) & 3) << (2 * (k & 31));

  const int rpb = rows_per_band(R, mode);
  const int NB = (int)((LQ + rpb - 1) / rpb);
  const long long skew = (mode == 2 || mode == 6 || mode == 8) ? 31 * (1 + slack) : 31 * (2 + slack + g_hs) + 1 + g_hs;
  const int align = (mode == 2 || mode == 6 || mode == 8) ? kChunk : kBlock;
  long long nsteps = ((LT + skew + align - 1) / align) * align;
  long long ext_len = 1; int ext_shift = 0;
  while (ext_len < nsteps + kChunk) { ext_len <<= 1; ++ext_shift; }
  int link_shift = 0; while ((1LL << link_shift) < link_len) ++link_shift;

  int result[2] = {0, 0};
  std::vector<int> cand((size_t)3 * (NB + 1), 0);        // mode 6: {H, T position, Q row} per band
  std::vector<std::vector<uint2>> links(G), ext(G);
  std::vector<std::vector<unsigned long long>> progress(G);
  std::vector<EngineParams> P(G);
  for (int g = 0; g < G; ++g) {
    links[g].assign((size_t)(W > 1 ? W - 1 : 1) * 2 * link_len, make_uint2(0, 0));
    ext[g].assign((size_t)2 * ext_len, make_uint2(0, 0));      // ext_in of GPU g
    progress[g].assign(W + 1, 0);
  }
  for (int g = 0; g < G; ++g) {
    EngineParams& p = P[g];
    p.q_codes = q.data(); p.t_packed = t.data(); p.LQ = LQ; p.LT = LT; p.NB = NB;
    p.ring_total = G * W; p.ring_offset = g * W; p.warps_local = W;
    p.links = links[g].data(); p.link_mask = (unsigned)(link_len - 1); p.link_shift = link_shift;
    p.progress = progress[g].data();
    p.ext_in = ext[g].data(); p.ext_out = ext[(g + 1) % G].data();
    p.ext_mask = (unsigned)(ext_len - 1); p.ext_shift = ext_shift;
    p.tag_base = epoch << 26; p.ext_tag_base = (epoch << 26) | 0x5Au; p.result = result;
    p.match = ma; p.mismatch = mi; p.gap_init = gi; p.gap_ext = ge;
    p.spin_limit = 200000000LL;
    p.cand = cand.data();
  }
  // optional: the last band's bottom boundary row (what the two-sided sweep combines), dumped to $EMU_FINAL
  std::vector<uint2> final_row((size_t)4 * ext_len, make_uint2(0, 0));
  if (getenv("EMU_FINAL")) for (int g = 0; g < G; ++g) { P[g].final_out = final_row.data(); P[g].final_mask = (unsigned)(ext_len - 1); }
  std::vector<WarpShared> ws((size_t)G * W);
  std::vector<WarpSmem> sm((size_t)G * W);
  for (auto& x : ws) pthread_barrier_init(&x.bar, nullptr, 32);
  std::vector<std::thread> th;
  for (int g = 0; g < G; ++g)
    for (int w = 0; w < W; ++w)
      for (int l = 0; l < 32; ++l)
        th.emplace_back(fn, &P[g], &ws[(size_t)g * W + w], l, w, &sm[(size_t)g * W + w]);
  for (auto& x : th) x.join();
  printf("score=%d status=%d bands=%d nsteps=%lld", result[0], result[1], NB, nsteps);
  if (mode == 6 || mode == 8) {   // the host's reduction: best H, then smallest T position, then smallest Q row
    int h = 0, pp = 0x7fffffff, rr = 0x7fffffff;
    for (int b = 0; b < NB; ++b) {
      const int ch = cand[3 * b], cp = cand[3 * b + 1], cr = cand[3 * b + 2];
      if (ch > h || (ch == h && (cp < pp || (cp == pp && cr < rr)))) { h = ch; pp = cp; rr = cr; }
    }
    printf(" endh=%d endpos=%d endrow=%d", h, pp, rr);
  }
  printf("\n");
  if (getenv("EMU_FINAL")) {
    FILE* f = fopen(getenv("EMU_FINAL"), "wb");
    long long hdr[2] = {LT, skew};
    fwrite(hdr, sizeof hdr, 1, f); fwrite(final_row.data(), sizeof(uint2), final_row.size(), f); fclose(f);
  }
  return 0;
}
#endif
