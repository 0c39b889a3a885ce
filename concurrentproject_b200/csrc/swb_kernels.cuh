// swb_kernels.cuh -- __global__ wrappers of the wavefront engine and their lookup table.
#pragma once
#include "swb_engine.cuh"

namespace swb {

// config 1: 4 warps per CTA (one per SM scheduler), boundary value consumed one step late (SLACK 1)
// config 2: 8 warps per CTA (two per scheduler), consumed in the same step (SLACK 0)
// config 3: 4 warps per CTA, consumed in the same step (shortest pipeline; the SHFL latency is covered by
//           the independent E / substitution work once R is large enough)
// config 4: config 1 plus the same slack for the hand-off inside a thread (HS 1: the hi sub-lane runs two positions
//           behind the lo sub-lane, which takes the row chain out of the per-step recurrence); packed 16-bit modes only
// config 5: config 1 with the boundary chunk fetched by one cp.async.bulk + mbarrier (TMA) instead of 32 lanes' loads;
//           packed 16-bit modes only
// config 6: 12 warps per CTA (three per scheduler), slack step, short-chain row loop, at most 10 rows per
//           sub-lane (170 registers per thread, 36 KB of shared memory): for pairs with far more bands than warps, where
//           the schedulers and not the pipeline depth are the limit; packed 16-bit modes only
// config 7: the CTA-chained engine of swb_chain.cuh (four consecutive bands per CTA handed over through shared memory,
//           a helper warp for tables and the L2 link); plain packed 16-bit modes, one band per warp, one GPU
constexpr int kNumConfigs = 6;      // launch configs of sw_engine_kernel; 7 is a kernel of its own
SWB_HD int config_wpc(int config) { return config == 2 ? 8 : (config == 6 ? 12 : 4); }
SWB_HD int config_slack(int config) { return (config == 1 || config == 4 || config == 5 || config == 6 || config == 7) ? 1 : 0; }
SWB_HD int config_hs(int config) { return config == 4 ? 1 : 0; }
// steps by which the last sub-lane of a band trails the first (what a band adds to the sweep; entry slot - T position)
SWB_HD int config_skew(int config, int mode) {
  const int slack = config_slack(config), hs = config_hs(config);
  return mode_is_s32(mode) ? 31 * (1 + slack) : 31 * (2 + slack + hs) + 1 + hs;
}

constexpr int kRowChoices[] = {1, 2, 3, 4, 6, 8, 10, 12, 14, 16};
constexpr int kNumRowChoices = 10;

// mode 0: s16x2 affine, 1: s16x2 linear (gap_init == gap_ext), 2: s32 affine,
// mode 3 / 4: modes 0 / 1 with re-based lanes (scores beyond the s16 range at the packed rate)
const void* engine_kernel_mode0(int R, int config);
const void* engine_kernel_mode1(int R, int config);
const void* engine_kernel_mode2(int R, int config);
const void* engine_kernel_mode3(int R, int config);
const void* engine_kernel_mode4(int R, int config);
const void* engine_kernel_mode5(int R, int config);   // s32, any byte alphabet (compare instead of table look-up)
const void* engine_kernel_mode6(int R, int config);   // mode 2 + position of the maximum (swb200_score_end)
const void* engine_kernel_mode7(int R, int config);   // mode 5 + position of the maximum
const void* engine_kernel_mode8(int R, int config);   // mode 6 with the anchored recurrence (start cell, swb200_score_span)
const void* engine_kernel_mode9(int R, int config);   // mode 7 with the anchored recurrence
const void* engine_kernel_mode10(int R, int config);  // anchored recurrence + traceback directions (swb200_align); R = 8 only
const void* engine_kernel_mode11(int R, int config);  // the same for any byte alphabet

// One launch can carry two independent sub-problems (two-sided sweep): warps [0, split) run `a`, the rest run `b`,
// each as its own ring with its own buffers.  split == 0: everything runs `a`.
struct EngineLaunch {
  EngineParams a, b;
  int split;
};

#ifdef __CUDACC__
template <int R, int MODE, int SLACK, int WPC, int HS = 0, bool TMA = false>
__global__ void __launch_bounds__(WPC * 32, 1) sw_engine_kernel(const __grid_constant__ EngineLaunch L) {
  __shared__ typename std::conditional<TMA, WarpSmemTma, WarpSmem>::type sm[WPC];
  WarpCtx w{(int)(threadIdx.x & 31)};
  const int wi = (int)(threadIdx.x >> 5);
  const int lw_all = (int)blockIdx.x * WPC + wi;
  const bool second = L.split > 0 && lw_all >= L.split;
  const EngineParams& P = second ? L.b : L.a;
  const int lw = second ? lw_all - L.split : lw_all;
#ifdef SWB_FORCE_LONG_CHAIN                       // measurement builds only (bench/sweep.py A/B)
  constexpr bool SHORT = false;
#else
  constexpr bool SHORT = WPC != 8;     // one warp per scheduler (and config 6): shortest dependency chain; two: fewest instructions
#endif
  if constexpr (MODE == 2) engine_warp_s32<R, SLACK, false, SHORT>(P, w, lw, &sm[wi]);
  else if constexpr (MODE == 5) engine_warp_s32<R, SLACK, true, SHORT>(P, w, lw, &sm[wi]);
  else if constexpr (MODE == 6) engine_warp_s32<R, SLACK, false, SHORT, true>(P, w, lw, &sm[wi]);
  else if constexpr (MODE == 7) engine_warp_s32<R, SLACK, true, SHORT, true>(P, w, lw, &sm[wi]);
  else if constexpr (MODE == 8) engine_warp_s32<R, SLACK, false, SHORT, true, true>(P, w, lw, &sm[wi]);
  else if constexpr (MODE == 9) engine_warp_s32<R, SLACK, true, SHORT, true, true>(P, w, lw, &sm[wi]);
  else if constexpr (MODE == 10) engine_warp_s32<R, SLACK, false, false, false, true, true>(P, w, lw, &sm[wi]);
  else if constexpr (MODE == 11) engine_warp_s32<R, SLACK, true, false, false, true, true>(P, w, lw, &sm[wi]);
  else if constexpr (MODE >= 3) engine_warp_s16<R, MODE - 3, SLACK, true, SHORT, HS, TMA>(P, w, lw, &sm[wi]);
  else engine_warp_s16<R, MODE, SLACK, false, SHORT, HS, TMA>(P, w, lw, &sm[wi]);
}

template <int MODE>
static const void* engine_kernel_lookup(int R, int config) {
  constexpr bool S16 = MODE <= 1 || MODE == 3 || MODE == 4;
  if (config < 1 || config > kNumConfigs || (config > 3 && !S16)) return nullptr;
#define SWB_CASE(RR)                                                             \
  case RR:                                                                       \
    if constexpr (S16) {                                                         \
      if (config == 4) return (const void*)sw_engine_kernel<RR, MODE, 1, 4, 1>;  \
      if (config == 5) return (const void*)sw_engine_kernel<RR, MODE, 1, 4, 0, true>;  \
      if constexpr (RR <= 10) { if (config == 6) return (const void*)sw_engine_kernel<RR, MODE, 1, 12>; }  \
      if (config == 6) return nullptr;                                           \
    }                                                                            \
    return config == 1 ? (const void*)sw_engine_kernel<RR, MODE, 1, 4>           \
         : config == 2 ? (const void*)sw_engine_kernel<RR, MODE, 0, 8>           \
                       : (const void*)sw_engine_kernel<RR, MODE, 0, 4>;
  if constexpr (MODE == 10 || MODE == 11) {      // one shape only: 8 rows per lane = one 32-bit direction word per step
    if (R != 8 || config == 2) return nullptr;
    return config == 1 ? (const void*)sw_engine_kernel<8, MODE, 1, 4> : (const void*)sw_engine_kernel<8, MODE, 0, 4>;
  } else {
    switch (R) {
      SWB_CASE(1) SWB_CASE(2) SWB_CASE(3) SWB_CASE(4) SWB_CASE(6) SWB_CASE(8) SWB_CASE(10) SWB_CASE(12) SWB_CASE(14) SWB_CASE(16)
      default: return nullptr;
    }
  }
#undef SWB_CASE
}
#endif

}  // namespace swb
