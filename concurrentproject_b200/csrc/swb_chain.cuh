// swb_chain.cuh -- the CTA-chained flavour of the pair engine (launch config 7), for pairs whose sweep is dominated by
// the pipeline fill (BASELINE config 2: 100 000 x 100 000, one band per warp, one warp per scheduler).
//
// Same recurrence, same striping and the same step as engine_warp_s16 (swb_engine.cuh: packed 16-bit lanes, slack step,
// short-chain row loop); what changes is everything AROUND the step, which ncu's source page showed to be a third of the
// sweep (profiles/r02_cfg2_base_stalls.json: per-chunk code 15 %, boundary waits 13 %):
//
//   * a CTA is four compute warps that own four CONSECUTIVE bands, plus one helper warp (two in the affine kernels).  The bottom boundary row of
//     warp w goes straight into the shared-memory inbox of warp w+1: lane 31 keeps the values of a group of steps in
//     registers and stores them after the group (predicated 16-byte stores, a CTA fence, then a step counter the
//     consumer polls).  Three of four hand-offs never leave the SM.  Warp 3 (and the last band of a side) stores
//     tagged 8-byte entries to a full-length link in L2, inside the step, as sw_engine_kernel does.
//   * the helper warp does what the compute warps' chunk prologue did: it expands the streamed sequence into the
//     CTA's ring of 4-byte substitution tables (one ring for all four warps), polls the L2 link of the CTA above,
//     validates the tagged entries and stages them in warp 0's inbox.  It shares a scheduler with one compute warp,
//     fills issue slots that warp leaves empty and sleeps when it has nothing to do -- and still slows that warp by 2-3
//     cycles per step, so WHICH warp it sits next to matters (the slowest band paces the chain): the CTA has eight warp
//     slots, the helper(s) are picked among warps 4-7 per kernel flavour (see SWB_CHAIN_HELPER_WARP below), the rest exit.
//   * what is left in a compute warp between two groups of 32 (linear, rows <= 3) or 16 steps: ONE conditional branch
//     (the loop back-edge, on a counter and a back-pressure word read a few steps earlier), one predicated store of its
//     own progress, two address computations.  A single warp per scheduler pays 20-30 cycles per branch.
//   Measured on BASELINE config 2 (profiles/r02_chain_experiments.txt, r02_chain_fit.txt): 44.0 cycles per step and 150
//   steps between the starts of consecutive bands (pair engine: 56.6 and 159): 4.08 -> 3.11 ms.
//
// Scope: plain packed 16-bit lanes (engine modes 0 and 1), one band per warp (bands <= 4 * CTAs), one GPU, with or
// without the two-sided sweep.  Everything else (re-based lanes, several rounds, the GPU ring, 32-bit lanes) stays
// with sw_engine_kernel; run_once() falls back by itself.  Replaces, like the rest of the engine, the reference's
// per-anti-diagonal launches (cudaSmithM.cu:87-189, SmithDiagonalGPU.cu:40-67).
#pragma once
#include "swb_engine.cuh"
#if defined(SWB_CHAIN_PROF) || defined(SWB_CHAIN_CHECK)
#include <cstdio>
#endif

namespace swb {

constexpr int kChainConfig = 7;
#ifdef SWB_CHAIN_CHECK_BREAK            // negative control of the protocol checker: every compute warp starts a group this many
constexpr int kChainCheckBreak = SWB_CHAIN_CHECK_BREAK;    // entries before its inputs are complete -- the checker must report it
#else
constexpr int kChainCheckBreak = 0;
#endif
constexpr int kChainGMax = 32;       // steps between two looks at the hand-off counters: 32 for the smallest step loops, else 16
#ifdef SWB_CHAIN_G
SWB_HD constexpr int chain_group(int, int) { return SWB_CHAIN_G; }
#else
SWB_HD constexpr int chain_group(int mode, int R) { return (mode == 1 && R <= 3) ? 32 : 16; }   // measured on cfg2 (profiles/r02_chain_*.txt)
#endif
constexpr int kChainInb = 512;       // entries of a compute warp's inbox ring
constexpr int kChainTab = 2048;      // entries of the CTA's table ring (kept twice)
// Helper warps.  A helper shares its scheduler with compute warp (its index % 4) and slows that warp's steps by the issue
// slots it takes -- and the slowest band paces the whole chain (measured per warp on cfg2: 40.3 cycles per step next to the
// helper, 37.5 elsewhere).  Warps 4-7 that are no helper exit at once.  Measured placements (profiles/r02_chain_experiments.txt):
// linear-gap kernels: ONE helper as warp 7 -- it sits next to warp 3, whose global sink makes it the fastest of the four
// (cfg2 3.106 -> 3.032 ms); affine kernels: the work split over TWO helpers on two schedulers (tables: warp 5; L2 link and
// warp 0's counter: warp 6) that sleep 300 ns when idle (4.535 -> 4.436 ms).  Two helpers cost the linear kernel more in
// polling than they spread.
#ifndef SWB_CHAIN_HELPERS                      // helpers of the linear-gap kernels
#define SWB_CHAIN_HELPERS 1
#endif
#ifndef SWB_CHAIN_HELPERS_AFF
#define SWB_CHAIN_HELPERS_AFF 2
#endif
#ifndef SWB_CHAIN_HELPER_WARP                  // the boundary helper (with one helper: the only helper), linear / affine
#define SWB_CHAIN_HELPER_WARP 7
#endif
#ifndef SWB_CHAIN_HELPER_WARP_AFF
#define SWB_CHAIN_HELPER_WARP_AFF 6
#endif
#ifndef SWB_CHAIN_TABLE_WARP                   // the table helper (two helpers only)
#define SWB_CHAIN_TABLE_WARP 5
#endif
#ifndef SWB_CHAIN_TABLE_WARP_AFF
#define SWB_CHAIN_TABLE_WARP_AFF 5
#endif
#ifndef SWB_CHAIN_HELPER_SLEEP                 // ns an idle helper sleeps between two looks, linear / affine
#define SWB_CHAIN_HELPER_SLEEP 100
#endif
#ifndef SWB_CHAIN_HELPER_SLEEP_AFF
#define SWB_CHAIN_HELPER_SLEEP_AFF 300
#endif
constexpr int kChainThreads = 256;
constexpr int kChainSK = 3;          // T positions between neighbouring lanes (two sub-lanes + the slack step)
constexpr int kChainSkew = 31 * kChainSK + 1;

struct ChainParams {
  const uint8_t* q_codes;
  const uint64_t* t_packed;
  long long LQ;
  int LT;
  int NB;                // bands of this side; CTA c owns bands 4c .. 4c+3
  uint2* links;          // link c = bottom boundary row of CTA c's last band: links + c * link_stride, entry index = step
  long long link_stride;
  uint2* final_out;      // optional: bottom boundary row of the side's LAST band, entry index = step (two-sided sweep)
  uint32_t tag;          // what a valid link entry of this launch carries in .y
  int* result;           // [0] best score, [1] status bits
  int match, mismatch, gap_init, gap_ext;
  long long spin_limit;
};

struct ChainLaunch {
  ChainParams a, b;
  int split;             // CTAs [0, split) run side a, the rest side b
};

#ifdef __CUDACC__

struct ChainSmem {
  uint32_t tab[2 * kChainTab];            // substitution tables per T position, filled by the helper warp
  uint32_t inbox[4][kChainInb + kChainGMax + 4]; // per compute warp: boundary values from the band above, slot = producer step mod ring;
                                          // the first group is kept twice (a reader's group of slots may run over the end)
  int cnt[4];                             // producer steps completed for warp w's inbox (w = 0: staged by the helper, table included)
  int done[4];                            // compute warp w has finished every step before this one
  int abort;
  uint32_t tagw;                          // P.tag, read back through shared memory so that it lives in a register (see chain_compute)
  int never;                              // 0x3fffffff: the back-pressure word of a warp whose sink is not a shared-memory inbox
  int tab_pub;                            // two helpers: tables are ready up to this T position (table helper -> boundary helper)
#ifdef SWB_CHAIN_CHECK
  // Protocol checker (bench/chain_variants.sh "-DSWB_CHAIN_CHECK"; compute-sanitizer is closed on this GPU pool): next to
  // every inbox slot the producer step it holds, next to every table slot its T position.  Every read of the step loop
  // checks that it found the entry it expects -- a read ahead of the producer finds an older step, a producer that
  // overwrites an unread slot a newer one.  Mismatches are counted in result[11] and the first few printed.
  unsigned short inbox_step[4][kChainInb + kChainGMax + 4];     // low 16 bits (the rings are far shorter than 65536 entries;
  unsigned short tab_pos[2 * kChainTab];                        //  static shared memory is full otherwise)
#endif
};

__device__ __forceinline__ int chain_ld(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void chain_st(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// The boundary values of a group of steps leave lane 31 together, after the group: stores inside the step sequence
// would pin every shared-memory load behind them (ptxas cannot tell the inbox from the table ring) and cost 9 cycles
// per step in its schedule.  16-byte stores of bare values, predicated, not branched: only lane 31 stores, and a
// divergent region between two groups costs a single warp dearly.
__device__ __forceinline__ void chain_st4_shared(bool on, uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n\t}"
               ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"((uint32_t)on));
}
__device__ __forceinline__ void st_entry_gpu(uint2* p, uint32_t value, uint32_t tag) {
  asm volatile("st.global.relaxed.gpu.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(value), "r"(tag));
}
// counters through 32-bit shared addresses computed once per band (a generic-to-shared conversion inside the loop is an
// S2R + LEA per use, and the S2R sat in the single warp's way at every loop back-edge); no compiler barrier around the
// load: its value is only compared at the end of the group
__device__ __forceinline__ int chain_ld_nb_s(uint32_t addr) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void chain_st_if_s(bool on, uint32_t addr, int v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.volatile.shared.s32 [%0], %1;\n\t}"
               ::"r"(addr), "r"(v), "r"((uint32_t)on) : "memory");
}
// shared-memory word at a 32-bit shared address (no generic-to-shared conversion inside the loops; ptxas folds the
// constant part of the address into the instruction)
__device__ __forceinline__ uint32_t chain_lds(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// orders lane 31's value stores before its counter store (both to shared memory)
__device__ __forceinline__ void chain_fence() {
#ifndef SWB_CHAIN_NO_FENCE
  asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
}

// Out-of-line wait of a compute warp between two groups of steps (every lane reads the same words, so the loop is
// warp-uniform): until the input counter has reached `need` and the band below has consumed up to `bp_need`.
// Returns the input counter, or -1 when the wait budget is spent or another warp of the CTA gave up.
static __device__ __noinline__ int chain_slow(const int* cnt_in, int need, const int* bp_word, int bp_need, long long budget,
                                              ChainSmem* sm, int* result) {
  for (;;) {
    const int v = chain_ld(cnt_in);
    const int b = chain_ld(bp_word);
    if (v >= need && b >= bp_need) return v;
    if (chain_ld(&sm->abort) || --budget < 0) {
      chain_st(&sm->abort, 1);
      atomicOr(result + 1, STATUS_SPIN_TIMEOUT);
      return -1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
//  helper warp
// ---------------------------------------------------------------------------------------------------------------
template <int ROLE, int SLEEP>   // ROLE 0: tables and boundary (the only helper), 1: tables, 2: boundary + warp 0's counter; SLEEP: ns when idle
static __device__ __noinline__ void chain_helper(const ChainParams& P, ChainSmem* sm, int c, int lane, int nact) {
  const int LT = P.LT;
  const int nsteps = ((LT + kChainSkew + kChainGMax - 1) / kChainGMax) * kChainGMax;
  const int tab_end = ((nsteps + 64 + 31) / 32) * 32;              // pad tables beyond LT as far as any lane looks
  const bool has_src = c > 0;
  const uint2* lin = has_src ? P.links + (size_t)(c - 1) * (size_t)P.link_stride : nullptr;
  const uint32_t tag = P.tag;
  const uint32_t padb = (uint32_t)(P.mismatch + P.gap_init) & 0xFFu;
  const uint32_t padw = padb * 0x01010101u;
  const uint32_t flip = padb ^ ((uint32_t)(P.match + P.gap_init) & 0xFFu);
  const uint64_t* tp = P.t_packed;
  int tab_pos = 0;                                                  // ROLE 2: what the table helper has published
  int bnd_pos = (has_src && ROLE != 1) ? 0 : 0x3fffffff;
  uint64_t tw = (ROLE != 2 && LT > 0) ? ld_early_u64(tp) : 0ull;     // packed word of positions [tab_pos, tab_pos + 32)
  long long budget = P.spin_limit;
  int idle = 0;
  while (tab_pos < tab_end || bnd_pos < LT) {
    bool progress = false;
    const int d0 = chain_ld(&sm->done[0]);
    if (ROLE == 2) {
      const int tpub = chain_ld(&sm->tab_pub);
      if (tpub != tab_pos) { tab_pos = tpub; progress = true; }
    }
    // ---- substitution tables: at most 768 positions ahead of warp 0, never over entries the last warp still reads
    if (ROLE != 2 && tab_pos < tab_end && tab_pos + 32 <= chain_ld(&sm->done[nact - 1]) - 128 + kChainTab && (tab_pos < d0 + 768 || (ROLE == 0 && bnd_pos >= LT))) {
      const int q = tab_pos + lane;
      uint32_t code = 4;
      if (q < LT) code = (uint32_t)(tw >> (2 * lane)) & 3u;
      const uint32_t w = table_word(code, padw, flip);
      sm->tab[q & (kChainTab - 1)] = w;
      sm->tab[(q & (kChainTab - 1)) + kChainTab] = w;
#ifdef SWB_CHAIN_CHECK
      sm->tab_pos[q & (kChainTab - 1)] = (unsigned short)q;
      sm->tab_pos[(q & (kChainTab - 1)) + kChainTab] = (unsigned short)q;
#endif
      tab_pos += 32;
      tw = tab_pos < LT ? ld_early_u64(tp + (tab_pos >> 5)) : 0ull;
      progress = true;
    }
    // ---- boundary row of the CTA above: L2 link -> warp 0's inbox (positions beyond LT do not exist and count as present)
    if (ROLE != 1 && bnd_pos < LT && bnd_pos + 32 <= d0 + kChainInb - 64) {
      const int p = bnd_pos + lane;
      uint2 e = make_uint2(0u, tag);
      if (p < LT) e = ld_entry(lin + p + kChainSkew);
      const unsigned okm = __ballot_sync(0xffffffffu, e.y == tag);
      int n = okm == 0xffffffffu ? 32 : __ffs((int)~okm) - 1;
#ifdef SWB_CHAIN_HELPER_MINFWD                                        // experiment: forward only batches of at least this many entries
      if (n < SWB_CHAIN_HELPER_MINFWD && bnd_pos + n < LT) n = 0;
#endif
      if (lane < n && p < LT) {
        const int slot = (p + kChainSkew) & (kChainInb - 1);
        sm->inbox[0][slot] = e.x;
        if (slot < kChainGMax) sm->inbox[0][slot + kChainInb] = e.x;
#ifdef SWB_CHAIN_CHECK
        sm->inbox_step[0][slot] = (unsigned short)(p + kChainSkew);
        if (slot < kChainGMax) sm->inbox_step[0][slot + kChainInb] = (unsigned short)(p + kChainSkew);
#endif
      }
      if (n > 0) { bnd_pos += n; progress = true; }
    }
    if (progress) {
      __syncwarp();
      __threadfence_block();
      if (ROLE == 1) { if (lane == 0) chain_st(&sm->tab_pub, tab_pos); }
      else if (lane == 0) chain_st(&sm->cnt[0], ((bnd_pos >= LT || tab_pos < bnd_pos) ? tab_pos : bnd_pos) + kChainSkew);
      budget = P.spin_limit;
      idle = 0;
#ifdef SWB_CHAIN_HELPER_NAP                                           // experiment: rest after every piece of work as well
      __nanosleep(SWB_CHAIN_HELPER_NAP);
#endif
    } else {
      __nanosleep(SLEEP);                          // an idle helper leaves the scheduler to the compute warp it shares it with
      bool give_up = --budget < 0 || chain_ld(&sm->abort);
      if ((++idle & 255) == 0) give_up = give_up || (ld_flag(P.result + 1) & STATUS_SPIN_TIMEOUT);
      if (give_up) {
        chain_st(&sm->abort, 1);
        if (lane == 0) atomicOr(P.result + 1, STATUS_SPIN_TIMEOUT);
        return;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
//  compute warp: one band, all of T
// ---------------------------------------------------------------------------------------------------------------
template <int R, int MODE>
__device__ __forceinline__ void chain_compute(const ChainParams& P, ChainSmem* sm, int c, int w, int lane, int nact) {
  constexpr int SK = kChainSK, SKEW = kChainSkew, G = chain_group(MODE, R);
  const int band = 4 * c + w;
  const int LT = P.LT;
  const int nsteps = ((LT + SKEW + G - 1) / G) * G;
  const bool last_lane = lane == 31;
  const int src_lane = (lane + 31) & 31;
  const bool is_final = band == P.NB - 1 && P.final_out != nullptr;
  const bool has_sink = band + 1 < P.NB || is_final;
  const bool smem_sink = !is_final && band + 1 < P.NB && w < 3;
  const bool emit = has_sink && last_lane;
  uint2* const out_global = is_final ? P.final_out : P.links + (size_t)c * (size_t)P.link_stride;   // entry index = step
  const uint32_t nopen = pack2(-P.gap_init), next = pack2(-P.gap_ext);
  const uint32_t fnext = pack2(-(P.gap_ext < P.gap_init ? P.gap_ext : P.gap_init));
  const uint32_t padw = ((uint32_t)(P.mismatch + P.gap_init) & 0xFFu) * 0x01010101u;
  const long long spin_limit = P.spin_limit;
  int* const result = P.result;
  // P is picked at run time, so ptxas re-reads P.tag from the constant bank through a register index at every store
  // (LDC in the step loop) unless the value is opaque to it: it comes back from shared memory
  const uint32_t tag_r = (uint32_t)chain_ld(reinterpret_cast<const int*>(&sm->tagw));
  int* const cnt_in = &sm->cnt[w];
  int* const cnt_out = &sm->cnt[(w + 1) & 3];
  int* const done_me = &sm->done[w];
  const int* const done_next = &sm->done[(w + 1) & 3];

  uint32_t sel[R];
  {
    const long long row_lo = (long long)band * (64LL * R) + (long long)(2 * lane) * R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long a = row_lo + r, b = row_lo + R + r;
      sel[r] = mk_sel16(a < P.LQ ? P.q_codes[a] : 4u, b < P.LQ ? P.q_codes[b] : 4u);
    }
  }
  uint32_t Ho[R], E[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { Ho[r] = nopen; E[r] = nopen; }
  uint32_t Fbot = nopen, up_prev = nopen, xsend = nopen, yold = nopen, Thi = padw;
  uint32_t best0 = 0, best1 = 0;

  // Between two groups of steps a compute warp does ONE compare-and-branch (to the out-of-line wait): the input counter
  // and the back-pressure word the next group needs are read a few steps before the end of the current group.  (A single warp per
  // scheduler pays 20-30 cycles for every branch; the first version of this loop had six per group and spent 260
  // cycles there.)  The flavour of the sink is a template parameter of the whole loop, not a branch per group.
  const int limit = w == 0 ? 0x3fffffff : LT + SKEW;      // warps 1-3: the band above stops at LT + SKEW; warp 0 also waits
                                                          // for the helper's pad tables beyond LT, which its counter covers
  const int* const bp_word = smem_sink ? done_next : &sm->never;
#ifdef SWB_CHAIN_PROF
  long long pr_wait = 0, pr_bp = 0, pr_waits = 0; unsigned pr_am_and = 0xffffffffu; int pr_am_n = 0;
#endif
  int have = chain_slow(cnt_in, (G + 1 + SKEW) < limit ? (G + 1 + SKEW) : limit, bp_word, 0, spin_limit, sm, result);
  __syncwarp();
#ifdef SWB_CHAIN_PROF
  const long long pr_t0 = clock64();
  unsigned long long pr_gt0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pr_gt0));
#endif
  // boundary value of T position 0 for lane 0 (nothing was shuffled before step 0)
  if (have >= 0 && lane == 0) yold = sm->inbox[w][SKEW & (kChainInb - 1)];

  uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(&sm->tab[0]);
  uint32_t inb_s = (uint32_t)__cvta_generic_to_shared(&sm->inbox[w][0]);
  uint32_t out_s = (uint32_t)__cvta_generic_to_shared(&sm->inbox[(w + 1) & 3][0]);
  asm volatile("" : "+r"(tab_s), "+r"(inb_s), "+r"(out_s));          // computed once, not once per group
  uint32_t cnt_in_s = (uint32_t)__cvta_generic_to_shared(cnt_in), cnt_out_s = (uint32_t)__cvta_generic_to_shared(cnt_out);
  uint32_t done_me_s = (uint32_t)__cvta_generic_to_shared(done_me), bp_s = (uint32_t)__cvta_generic_to_shared(bp_word);
  asm volatile("" : "+r"(cnt_in_s), "+r"(cnt_out_s), "+r"(done_me_s), "+r"(bp_s));
  auto sweep = [&](auto GS_) {
    // GS (global sink: warp 3 / a side's last band) stores its {value, tag} entry inside the step, as sw_engine_kernel
    // does -- global stores cannot alias the shared-memory loads.  Shared sinks collect the group's values in
    // registers and store them after the group: a shared store inside the step sequence would pin every later
    // table / inbox load behind it (ptxas cannot tell the rings apart) -- 9 cycles per step in its schedule.
    constexpr bool GS = decltype(GS_)::value;
    int i0 = 0;
    for (;;) {
      bool go_on;
      do {      // ---- one group of G steps per trip; ONE branch per group when nothing has to be waited for
        chain_st_if_s(lane == 0, done_me_s, i0);
        const uint32_t tabp = tab_s + 4u * ((i0 - SK * lane) & (kChainTab - 1));
        const uint32_t inbp = inb_s + 4u * ((i0 + SKEW + 1) & (kChainInb - 1));  // lane 31 ships T position i + 1 = producer step i + 1 + SKEW
        uint32_t xq[GS ? 1 : G];
        uint2* const outg = out_global + i0;
        int bp = 0x3fffffff;
        // (the group's first table word and boundary value are read HERE, after the branch that decided this group's inputs
        //  are complete: ptxas moves plain shared loads across the volatile counter load freely, so a read at the end of
        //  the previous group -- "validated" by a counter value read earlier in program order -- returned stale entries)
        uint32_t Tnext = chain_lds(tabp);
        uint32_t xnext = chain_lds(inbp);
#ifdef SWB_CHAIN_PROF
        const long long tg0 = clock64();
#endif
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const uint32_t Tlo = Tnext;
          const uint32_t xin = xnext;
#ifdef SWB_CHAIN_CHECK
          {
            const int pos = i0 + k - SK * lane;                       // the T position whose table word is in Tlo
            const int got_pos = sm->tab_pos[((i0 - SK * lane) & (kChainTab - 1)) + k];
            if (pos >= 0 && got_pos != (pos & 0xffff) && atomicAdd(result + 11, 1) < 8)
              printf("chaincheck TABLE cta %d warp %d lane %d step %d: expected position %d, slot holds %d\n", c, w, lane, i0 + k, pos, got_pos);
            const int pstep = i0 + k + 1 + SKEW;                      // the producer step whose value is in xin
            const int got_step = sm->inbox_step[w][((i0 + SKEW + 1) & (kChainInb - 1)) + k];
            if (last_lane && band > 0 && pstep < LT + SKEW && got_step != (pstep & 0xffff) && atomicAdd(result + 11, 1) < 8)
              printf("chaincheck INBOX cta %d warp %d step %d: expected producer step %d, slot holds %d\n", c, w, i0 + k, pstep, got_step);
          }
#endif
#ifdef SWB_CHAIN_X_NOLDS                                             // timing experiment only (wrong scores)
          Tnext = Tnext * 5u + 1u;
          xnext = xnext ^ Tnext;
#else
          if (k + 1 < G) {
            Tnext = chain_lds(tabp + 4u * (k + 1));
            xnext = chain_lds(inbp + 4u * (k + 1));
          }
#endif
#ifndef SWB_CHAIN_PF
#define SWB_CHAIN_PF 8
#endif
          if (k == G - SWB_CHAIN_PF) {                               // what the NEXT group needs, read a few steps early
            have = chain_ld_nb_s(cnt_in_s);
            if (!GS) bp = chain_ld_nb_s(bp_s);
          }
#ifdef SWB_CHAIN_PROF
          if (k == 3 || k == 12) { const unsigned am = __activemask(); pr_am_and &= am; pr_am_n += (am != 0xffffffffu); }
#endif
          const uint32_t xs = last_lane ? xin : xsend;
#ifdef SWB_CHAIN_X_NOSHFL                                            // timing experiment only (wrong scores)
          const uint32_t ynew = xs;
#else
          const uint32_t ynew = __shfl_sync(0xffffffffu, xs, src_lane);
#endif
          const uint32_t yuse = yold;
          yold = ynew;
          uint32_t upHo, F;
          if (MODE == 0) {
            upHo = prmt(yuse, Ho[R - 1], 0x5410u);
            F = prmt(yuse, Fbot, 0x5432u);
          } else {
            upHo = prmt(yuse, Ho[R - 1], 0x5432u);
            F = 0;
          }
          uint32_t diag = up_prev;
          up_prev = upHo;
          uint32_t X = upHo, hprev = 0;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const uint32_t s = prmt(Tlo, Thi, sel[r]);
            const uint32_t old = Ho[r];
            uint32_t h;
            if (MODE == 0) {
              E[r] = addmax16x2(E[r], next, old);
              const uint32_t m = addmaxrelu16x2(diag, s, E[r]);
              F = addmax16x2(F, r == 0 ? next : fnext, X);
              X = add16x2(m, nopen);
              h = max16x2(m, F);
            } else {
              const uint32_t t = addmax16x2(diag, s, old);
              h = r == 0 ? maxrelu16x2(t, X) : addmaxrelu16x2(hprev, nopen, t);
            }
            hprev = h;
            Ho[r] = add16x2(h, nopen);
            diag = old;
            if (r & 1) best1 = max16x2(best1, h); else best0 = max16x2(best0, h);
          }
          Fbot = F;
          xsend = (MODE == 0) ? prmt(Ho[R - 1], Fbot, 0x7632u) : Ho[R - 1];
          Thi = Tlo;
          if (GS) { if (emit) st_entry_gpu(outg + k, xsend, tag_r); }
          else xq[k] = xsend;
        }
        if (!GS) {
          const uint32_t o = out_s + 4u * (i0 & (kChainInb - 1));
#pragma unroll
          for (int k = 0; k < G; k += 4) chain_st4_shared(emit, o + 4u * k, xq[k], xq[k + 1], xq[k + 2], xq[k + 3]);
          // the ring's first group is kept twice (predicated: a branch here would cost more than four idle stores)
          const bool twice = emit && (i0 & (kChainInb - 1)) == 0;
#pragma unroll
          for (int k = 0; k < G; k += 4) chain_st4_shared(twice, o + 4u * (kChainInb + k), xq[k], xq[k + 1], xq[k + 2], xq[k + 3]);
#ifdef SWB_CHAIN_CHECK
          if (emit) {
            unsigned short* const st = &sm->inbox_step[(w + 1) & 3][i0 & (kChainInb - 1)];
            for (int k = 0; k < G; ++k) { st[k] = (unsigned short)(i0 + k); if ((i0 & (kChainInb - 1)) == 0) st[kChainInb + k] = (unsigned short)(i0 + k); }
          }
#endif
          chain_fence();
          chain_st_if_s(emit, cnt_out_s, i0 + G);
        }
#ifdef SWB_CHAIN_PROF
        pr_bp += clock64() - tg0;
#endif
        // the next group: lane 31 reads producer steps up to i0 + G + SKEW, and (shared sink) the band below must have
        // read what this warp's stores of that group overwrite
        i0 += G;
        const int need_full = i0 + G + 1 + SKEW - kChainCheckBreak;
        const int need = need_full < limit ? need_full : limit;
        go_on = i0 < nsteps && have >= need && (GS || bp >= i0 + G - (kChainInb - 64));
      } while (go_on);
      if (i0 >= nsteps) break;
      {
#ifdef SWB_CHAIN_PROF
        const long long t0 = clock64();
#endif
        const int need_full = i0 + G + 1 + SKEW - kChainCheckBreak;
        have = chain_slow(cnt_in, need_full < limit ? need_full : limit, bp_word, GS ? 0 : i0 + G - (kChainInb - 64), spin_limit, sm, result);
        __syncwarp();
#ifdef SWB_CHAIN_PROF
        pr_wait += clock64() - t0; pr_waits += 1;
#endif
        if (have < 0) break;
      }
    }
  };
  if (have >= 0) {
    if (smem_sink || !has_sink) sweep(std::false_type{}); else sweep(std::true_type{});
  }
#ifdef SWB_CHAIN_PROF
  if (lane == 0 && c >= 40 && c <= 42)
    printf("chainstart cta %d warp %d band %d start_ns %llu\n", c, w, band, pr_gt0);
  if ((lane == 0 || lane == 31) && (c == 10 || c == 40))
    printf("chainprof cta %d warp %d band %d: %.1f cyc/step over %d steps, waits %lld (%.1f cyc/step), group body %.1f cyc/step, activemask and %08x partial %d\n", c, w, band,
           (double)(clock64() - pr_t0) / nsteps, nsteps, pr_waits, (double)pr_wait / nsteps, (double)pr_bp / nsteps, pr_am_and, pr_am_n);
#endif
#ifdef SWB_CHAIN_CHECK
  if (lane == 0 && have >= 0) atomicAdd(result + 10, nsteps >> 10);   // reads checked by this warp, in units of 1024 steps (33 reads per step)
#endif
  if (lane == 0) chain_st(done_me, 0x3fffffff);
  const int m = __reduce_max_sync(0xffffffffu, hi_half_max(max16x2(best0, best1)));
  if (lane == 0) {
    atomicMax(P.result, m);
    if (m > 32767 - P.match - 1) atomicOr(P.result + 1, STATUS_S16_OVERFLOW);
  }
}

template <int R, int MODE>
__global__ void __launch_bounds__(kChainThreads, 1) sw_chain_kernel(const __grid_constant__ ChainLaunch L) {
  __shared__ ChainSmem sm;
  const bool second = (int)blockIdx.x >= L.split;
  const ChainParams& P = second ? L.b : L.a;
  const int c = (int)blockIdx.x - (second ? L.split : 0);
  const int wi = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  const int left = P.NB - 4 * c;
  const int nact = left < 4 ? left : 4;
  if (nact <= 0) return;                                            // (the host never launches such a CTA)
  {
    const uint32_t padw = ((uint32_t)(P.mismatch + P.gap_init) & 0xFFu) * 0x01010101u;
    for (int i = (int)threadIdx.x; i < 2 * kChainTab; i += (int)blockDim.x) sm.tab[i] = padw;
    const uint32_t nopen = pack2(-P.gap_init);
    // every inbox starts as a zero border: the first band of a side keeps reading it; the others read entries of
    // T positions beyond LT before (or without) their producer writing them -- harmless as long as what they find is
    // not above a real score (stale real entries are not, uninitialised shared memory is)
    uint32_t* const ib = &sm.inbox[0][0];
    for (int i = (int)threadIdx.x; i < 4 * (kChainInb + kChainGMax + 4); i += (int)blockDim.x) ib[i] = nopen;
#ifdef SWB_CHAIN_CHECK
    for (int i = (int)threadIdx.x; i < 4 * (kChainInb + kChainGMax + 4); i += (int)blockDim.x) (&sm.inbox_step[0][0])[i] = 0xffff;
    for (int i = (int)threadIdx.x; i < 2 * kChainTab; i += (int)blockDim.x) sm.tab_pos[i] = 0xffff;
#endif
    if (threadIdx.x < 4) { sm.cnt[threadIdx.x] = 0; sm.done[threadIdx.x] = 0; }
    if (threadIdx.x == 0) { sm.abort = 0; sm.tagw = P.tag; sm.never = 0x3fffffff; sm.tab_pub = 0; }
  }
  __syncthreads();
  constexpr int HB = MODE == 1 ? SWB_CHAIN_HELPER_WARP : SWB_CHAIN_HELPER_WARP_AFF;
  constexpr int HT = MODE == 1 ? SWB_CHAIN_TABLE_WARP : SWB_CHAIN_TABLE_WARP_AFF;
  constexpr int NH = MODE == 1 ? SWB_CHAIN_HELPERS : SWB_CHAIN_HELPERS_AFF;
  constexpr int SL = MODE == 1 ? SWB_CHAIN_HELPER_SLEEP : SWB_CHAIN_HELPER_SLEEP_AFF;
  static_assert(HB >= 4 && HB < 8 && HT >= 4 && HT < 8 && (NH == 1 || HB != HT), "helper warps are warps 4-7");
  if (NH == 1) {
    if (wi == HB) chain_helper<0, SL>(P, &sm, c, lane, nact);
  } else {
    if (wi == HT) chain_helper<1, SL>(P, &sm, c, lane, nact);
    if (wi == HB) chain_helper<2, SL>(P, &sm, c, lane, nact);
  }
  if (wi < nact) chain_compute<R, MODE>(P, &sm, c, wi, lane, nact);
}

#endif  // __CUDACC__

const void* chain_kernel(int mode, int R);      // null when there is no chain kernel for this (mode, rows)

}  // namespace swb
