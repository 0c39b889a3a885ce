// swb_engine.cuh -- the wavefront DP engine: score-only Gotoh local alignment of one pair.
//
// Replaces the reference's per-anti-diagonal kernels (simpleGPU.cu:78-107 DPMatrices,
// cudaLazy.cu:21-56 sw_kernel_diag, cudaSmithM.cu:87-126 kernel_compute_diagonal,
// SmithDiagonalGPU.cu:40-67 smithWatermanDiagonal) and their host loops.  Same recurrence as
// main.cpp:57-63, nothing else in common.
//
// Geometry.  One sequence ("Q") is striped across lanes, the other ("T") is streamed.  A *band* is
// the block of Q rows one warp owns while it sweeps all of T.  With packed 16-bit lanes a thread
// carries two *sub-lanes* (lo/hi half of every register) that own R consecutive Q rows each, so a
// band is 64*R rows; with 32-bit lanes a thread is one sub-lane and a band is 32*R rows.  Sub-lane
// v+1 runs one T position behind sub-lane v, so the anti-diagonal dependency of the recurrence
// moves between lanes through one __shfl_sync per step and never touches memory.  H/E stay in
// registers for the whole sweep; the only state that leaves a warp is the band's bottom boundary
// row (H-open and F per T position), handed to the warp that owns the next band through 8-byte
// {value,tag} words in L2 (or in a peer GPU's memory for the last warp of a GPU).  All warps of
// all GPUs form one ring; band b is processed by ring slot b mod ring_total, so consecutive bands
// run concurrently, one boundary-latency apart: a pipelined wavefront with no host involvement.
//
// Arithmetic per cell vector (s16x2: two cells): PRMT (substitution score from a 4-byte table of
// the streamed symbol), VIADD.16x2 (diag+s), VIADDMNMX.S16x2 (E), VIADDMNMX.S16x2 (F),
// VIMNMX3.S16x2.RELU (H), VIADD.16x2 (H-open), 1/2 VIMNMX (running best).  When gap_init ==
// gap_ext the E/F registers are provably redundant (E = H_left-g, F = H_up-g) and MODE 1 drops them.
#pragma once
#include "swb_device.cuh"

namespace swb {

constexpr int kChunk = 32;     // steps between boundary polls / table refills
constexpr int kTabRing = 256;  // per-warp ring of substitution tables (one per T position), kept twice
constexpr int kInbox = 64;     // per-warp ring of validated top-boundary values

enum : int { STATUS_S16_OVERFLOW = 1, STATUS_SPIN_TIMEOUT = 2, STATUS_BAD_SYMBOL = 4 };

struct EngineParams {
  const uint8_t* q_codes;       // LQ codes in {0,1,2,3}; >= 4 never matches
  const uint64_t* t_packed;     // LT 2-bit codes, 32 per 64-bit word, position p at bits 2*(p%32)
  long long LQ, LT;
  int NB;                       // bands = ceil(LQ / rows_per_band)
  int ring_total;               // warps in the whole ring (all GPUs)
  int ring_offset;              // global index of this GPU's first warp
  int warps_local;              // warps launched on this GPU
  uint2* links;                 // (warps_local-1) rings, 2*link_len entries each
  unsigned link_mask;           // link_len-1 (link_len is a power of two, >= 4096)
  int link_shift;               // log2(link_len)
  unsigned long long* progress; // [warps_local] consumer progress words for ring back-pressure
  const uint2* ext_in;          // stream consumed by local warp 0 (full length: 2*ext_len entries)
  uint2* ext_out;               // stream produced by the last local warp (peer memory on multi-GPU)
  unsigned ext_mask;
  int ext_shift;
  uint32_t tag_base;            // epoch << 26
  int* result;                  // [0] best score (atomicMax), [1] status bits (atomicOr)
  int match, mismatch, gap_init, gap_ext;
  long long spin_limit;         // polls before a waiting warp gives up (sets STATUS_SPIN_TIMEOUT)
};

struct WarpSmem {
  uint32_t tab[2 * kTabRing];
  uint32_t inbox[4 * kInbox];   // plane 0: H-open (or packed H-open|F), plane 1: F (s32); each ring kept twice
};

SWB_HD uint32_t pack2(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }

// PRMT selector producing the packed s16x2 substitution word of one row vector: low half from table
// register A (codes of the lo sub-lane's T symbol), high half from table register B.
SWB_HD uint32_t mk_sel16(uint32_t code_lo, uint32_t code_hi) {
  const uint32_t lo = code_lo < 4 ? (code_lo | ((code_lo | 8u) << 4)) : 0x88u;           // byte, then its sign
  const uint32_t hi = code_hi < 4 ? ((code_hi + 4u) | ((code_hi + 12u) << 4)) : 0xCCu;
  return lo | (hi << 8);
}
SWB_HD uint32_t mk_sel32(uint32_t code) {
  return code < 4 ? (code | ((code | 8u) * 0x1110u)) : 0x8888u;
}

// 4-byte table of (substitution score + gap_init) against T symbol c; c >= 4 = matches nothing.
SWB_HD uint32_t table_word(uint32_t c, uint32_t padw, uint32_t flip) { return c < 4 ? (padw ^ (flip << (8 * c))) : padw; }

struct Waiter {
  long long budget;
  bool aborted;
};

// Wait until the entry for producer step j of band `tagband` is present; returns its value.
SWB_HD uint32_t wait_entry(const EngineParams& P, const uint2* slot, uint32_t want_tag, Waiter& wt) {
  uint2 e = ld_entry(slot);
  while (e.y != want_tag && !wt.aborted) {
    if (--wt.budget < 0 || (((wt.budget & 1023) == 0) && (ld_flag(P.result + 1) & STATUS_SPIN_TIMEOUT))) {
      atomic_or_i32(P.result + 1, STATUS_SPIN_TIMEOUT);
      wt.aborted = true;
      break;
    }
    spin_pause();
    e = ld_entry(slot);
  }
  return e.x;
}

// =================================================================================================
//  Packed 16-bit engine.  MODE 0: affine gaps.  MODE 1: gap_init == gap_ext (E/F eliminated).
//  SLACK 1: a shuffled boundary value is consumed one step after it was sent (hides SHFL latency
//  when a scheduler has a single warp); SLACK 0: consumed in the same step (shorter pipeline).
// =================================================================================================
template <int R, int MODE, int SLACK>
SWB_HD void engine_warp_s16(const EngineParams& P, const WarpCtx& w, int lw, WarpSmem* sm) {
  constexpr int SK = 2 + SLACK;        // T positions between neighbouring lanes
  constexpr int SKEW = 31 * SK + 1;    // lane 31's hi sub-lane trails lane 0's lo sub-lane by this
  const int lane = w.lane;
  const bool last_lane = lane == 31;
  const int src_lane = (lane + 31) & 31;
  const uint32_t nopen = pack2(-P.gap_init), next = pack2(-P.gap_ext);
  const uint32_t padb = (uint32_t)(P.mismatch + P.gap_init) & 0xFFu;
  const uint32_t padw = padb * 0x01010101u;
  const uint32_t flip = padb ^ ((uint32_t)(P.match + P.gap_init) & 0xFFu);
  const long long LT = P.LT;
  const long long nsteps = ((LT + SKEW + kChunk - 1) / kChunk) * kChunk;
  uint32_t best0 = 0, best1 = 0;
  Waiter wt{P.spin_limit, false};

  for (long long band = P.ring_offset + lw; band < P.NB; band += P.ring_total) {
    const bool zero_src = band == 0;
    const bool has_sink = band + 1 < P.NB;
    const bool emit = has_sink && last_lane;
    const bool first_local = lw == 0, last_local = lw == P.warps_local - 1;
    const uint2* in = first_local ? P.ext_in : P.links + (size_t)(lw - 1) * 2 * ((size_t)P.link_mask + 1);
    const unsigned in_mask = first_local ? P.ext_mask : P.link_mask;
    const int in_shift = first_local ? P.ext_shift : P.link_shift;
    uint2* out = last_local ? P.ext_out : P.links + (size_t)lw * 2 * ((size_t)P.link_mask + 1);
    const unsigned out_mask = last_local ? P.ext_mask : P.link_mask;
    const int out_shift = last_local ? P.ext_shift : P.link_shift;
    const uint32_t in_tag = P.tag_base | ((uint32_t)band << 8);           // written by band-1 as (band-1)+1
    const uint32_t out_tag = P.tag_base | ((uint32_t)(band + 1) << 8);
    unsigned long long* my_progress = P.progress + lw;
    const unsigned long long* sink_progress = P.progress + lw + 1;       // only read when !last_local

    // ---- per-band constants: PRMT selectors of this thread's 2*R rows
    uint32_t sel[R];
    {
      const long long row_lo = band * (64LL * R) + (long long)(2 * lane) * R;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long long a = row_lo + r, b = row_lo + R + r;
        const uint32_t ca = a < P.LQ ? P.q_codes[a] : 4u;
        const uint32_t cb = b < P.LQ ? P.q_codes[b] : 4u;
        sel[r] = mk_sel16(ca, cb);
      }
    }
    // ---- state: H-open, E per row vector; zero boundary = H 0, E/F any value <= 0
    uint32_t Ho[R], E[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { Ho[r] = nopen; E[r] = nopen; }
    uint32_t Fbot = nopen, up_prev = nopen, xsend = nopen, yold = nopen, Thi = padw;

    // ---- table ring: everything pad, then T positions [0, 32)
    w.sync();
#pragma unroll
    for (int k = 0; k < 2 * kTabRing / 32; ++k) sm->tab[k * 32 + lane] = padw;
    w.sync();
    {
      const long long q = lane;
      uint32_t c = 4;
      if (q < LT) c = (uint32_t)(P.t_packed[q >> 5] >> (2 * (q & 31))) & 3u;
      const uint32_t tw = table_word(c, padw, flip);
      sm->tab[q & (kTabRing - 1)] = tw;
      sm->tab[(q & (kTabRing - 1)) + kTabRing] = tw;
    }

    for (long long i0 = 0; i0 < nsteps; i0 += kChunk) {
      // (a) substitution tables for T positions [i0+32, i0+64)
      {
        const long long q = i0 + kChunk + lane;
        uint32_t c = 4;
        if (q < LT) c = (uint32_t)(P.t_packed[q >> 5] >> (2 * (q & 31))) & 3u;
        const uint32_t tw = table_word(c, padw, flip);
        sm->tab[q & (kTabRing - 1)] = tw;
        sm->tab[(q & (kTabRing - 1)) + kTabRing] = tw;
      }
      // (b) top boundary for lane 0's positions [i0+SLACK, i0+SLACK+32): wait for the producer
      {
        const long long q = i0 + SLACK + lane;
        uint32_t v = nopen;
        if (!zero_src && q < LT) {
          const long long j = q + SKEW;                                   // producer step that emitted q
          v = wait_entry(P, in + (j & in_mask), in_tag | ((uint32_t)(j >> in_shift) & 0xFFu), wt);
        }
        sm->inbox[q & (kInbox - 1)] = v;
        sm->inbox[(q & (kInbox - 1)) + kInbox] = v;
        if (lane == 0 && (i0 & 255) == 0)
          st_progress(my_progress, ((unsigned long long)(band + 1) << 32) | (unsigned long long)(i0 + SLACK + kChunk));
      }
      // (c) ring back-pressure: never overwrite an entry the consumer has not read yet
      if (has_sink && !last_local && (i0 & 1023) == 0 && i0 + 1024 > (long long)out_mask + 1) {
        const unsigned long long need =
            ((unsigned long long)(band + 2) << 32) | (unsigned long long)(i0 + 1024 - ((long long)out_mask + 1));
        while (!wt.aborted && ld_progress(sink_progress) < need) {
          if (--wt.budget < 0) { atomic_or_i32(P.result + 1, STATUS_SPIN_TIMEOUT); wt.aborted = true; }
          spin_pause();
        }
      }
      w.sync();

      const uint32_t* tabp = sm->tab + ((i0 - (long long)SK * lane) & (kTabRing - 1));
      const uint32_t* inbp = sm->inbox + ((i0 + SLACK) & (kInbox - 1));
      uint2* outp = out + (i0 & out_mask);
      const uint32_t otag = out_tag | ((uint32_t)(i0 >> out_shift) & 0xFFu);

#pragma unroll 4
      for (int k = 0; k < kChunk; ++k) {
        const uint32_t Tlo = tabp[k];
        const uint32_t xin = inbp[k];
        const uint32_t xs = last_lane ? xin : xsend;
        const uint32_t ynew = w.shfl(xs, src_lane);
        const uint32_t yuse = SLACK ? yold : ynew;
        yold = ynew;
        uint32_t upHo, F;
        if (MODE == 0) {
          upHo = prmt(yuse, Ho[R - 1], 0x5410u);   // lo <- neighbour's bottom (H-open), hi <- own lo bottom
          F = prmt(yuse, Fbot, 0x5432u);           // lo <- neighbour's bottom F,      hi <- own lo bottom F
        } else {
          upHo = prmt(yuse, Ho[R - 1], 0x5432u);   // linear mode ships the whole H-open word
          F = 0;
        }
        uint32_t diag = up_prev, Hup = upHo;
        up_prev = upHo;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint32_t s = prmt(Tlo, Thi, sel[r]);
          const uint32_t d = add16x2(diag, s);
          const uint32_t old = Ho[r];
          uint32_t h;
          if (MODE == 0) {
            E[r] = addmax16x2(E[r], next, old);
            F = addmax16x2(F, next, Hup);
            h = max3relu16x2(d, E[r], F);
          } else {
            h = max3relu16x2(d, old, Hup);
          }
          Ho[r] = add16x2(h, nopen);
          Hup = Ho[r];
          diag = old;
          if (r & 1) best1 = max16x2(best1, h); else best0 = max16x2(best0, h);
        }
        Fbot = F;
        xsend = (MODE == 0) ? prmt(Ho[R - 1], Fbot, 0x7632u) : Ho[R - 1];
        Thi = Tlo;
        // slot = producer step (always in range); steps whose T position is outside [0,LT) are never read
        if (emit) st_entry(outp + k, xsend, otag);
      }
    }
  }

  // ---- running best: halves -> int, warp max, one atomic
  const uint32_t b = max16x2(best0, best1);
  int bi = (int)(short)(b & 0xFFFFu), bj = (int)(short)(b >> 16);
  int m = w.reduce_max(bi > bj ? bi : bj);
  if (lane == 0) {
    atomic_max_i32(P.result, m);
    if (m > 32767 - P.match - 1) atomic_or_i32(P.result + 1, STATUS_S16_OVERFLOW);
  }
}

// =================================================================================================
//  32-bit engine (scores beyond the s16 range): one sub-lane per thread, band = 32*R rows.
//  Boundary entries come in pairs: slot 2j = H-open, slot 2j+1 = F.
// =================================================================================================
template <int R, int SLACK>
SWB_HD void engine_warp_s32(const EngineParams& P, const WarpCtx& w, int lw, WarpSmem* sm) {
  constexpr int SK = 1 + SLACK;
  constexpr int SKEW = 31 * SK;
  const int lane = w.lane;
  const bool last_lane = lane == 31;
  const int src_lane = (lane + 31) & 31;
  const int nopen = -P.gap_init, next = -P.gap_ext;
  const uint32_t padb = (uint32_t)(P.mismatch + P.gap_init) & 0xFFu;
  const uint32_t padw = padb * 0x01010101u;
  const uint32_t flip = padb ^ ((uint32_t)(P.match + P.gap_init) & 0xFFu);
  const long long LT = P.LT;
  const long long nsteps = ((LT + SKEW + kChunk - 1) / kChunk) * kChunk;
  int best0 = 0, best1 = 0;
  Waiter wt{P.spin_limit, false};

  for (long long band = P.ring_offset + lw; band < P.NB; band += P.ring_total) {
    const bool zero_src = band == 0;
    const bool has_sink = band + 1 < P.NB;
    const bool emit = has_sink && last_lane;
    const bool first_local = lw == 0, last_local = lw == P.warps_local - 1;
    const uint2* in = first_local ? P.ext_in : P.links + (size_t)(lw - 1) * 2 * ((size_t)P.link_mask + 1);
    const unsigned in_mask = first_local ? P.ext_mask : P.link_mask;
    const int in_shift = first_local ? P.ext_shift : P.link_shift;
    uint2* out = last_local ? P.ext_out : P.links + (size_t)lw * 2 * ((size_t)P.link_mask + 1);
    const unsigned out_mask = last_local ? P.ext_mask : P.link_mask;
    const int out_shift = last_local ? P.ext_shift : P.link_shift;
    const uint32_t in_tag = P.tag_base | ((uint32_t)band << 8);
    const uint32_t out_tag = P.tag_base | ((uint32_t)(band + 1) << 8);
    unsigned long long* my_progress = P.progress + lw;
    const unsigned long long* sink_progress = P.progress + lw + 1;

    uint32_t sel[R];
    {
      const long long row0 = band * (32LL * R) + (long long)lane * R;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long long a = row0 + r;
        sel[r] = mk_sel32(a < P.LQ ? P.q_codes[a] : 4u);
      }
    }
    int Ho[R], E[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { Ho[r] = nopen; E[r] = nopen; }
    int up_prev = nopen, xsH = nopen, xsF = nopen, yoldH = nopen, yoldF = nopen;

    w.sync();
#pragma unroll
    for (int k = 0; k < 2 * kTabRing / 32; ++k) sm->tab[k * 32 + lane] = padw;
    w.sync();
    {
      const long long q = lane;
      uint32_t c = 4;
      if (q < LT) c = (uint32_t)(P.t_packed[q >> 5] >> (2 * (q & 31))) & 3u;
      const uint32_t tw = table_word(c, padw, flip);
      sm->tab[q & (kTabRing - 1)] = tw;
      sm->tab[(q & (kTabRing - 1)) + kTabRing] = tw;
    }

    for (long long i0 = 0; i0 < nsteps; i0 += kChunk) {
      {
        const long long q = i0 + kChunk + lane;
        uint32_t c = 4;
        if (q < LT) c = (uint32_t)(P.t_packed[q >> 5] >> (2 * (q & 31))) & 3u;
        const uint32_t tw = table_word(c, padw, flip);
        sm->tab[q & (kTabRing - 1)] = tw;
        sm->tab[(q & (kTabRing - 1)) + kTabRing] = tw;
      }
      {
        const long long q = i0 + SLACK + lane;
        uint32_t vH = (uint32_t)nopen, vF = (uint32_t)nopen;
        if (!zero_src && q < LT) {
          const long long j = q + SKEW;
          const uint32_t tg = in_tag | ((uint32_t)(j >> in_shift) & 0xFFu);
          vH = wait_entry(P, in + 2 * (j & in_mask), tg, wt);
          vF = wait_entry(P, in + 2 * (j & in_mask) + 1, tg, wt);
        }
        sm->inbox[q & (kInbox - 1)] = vH;
        sm->inbox[(q & (kInbox - 1)) + kInbox] = vH;
        sm->inbox[(q & (kInbox - 1)) + 2 * kInbox] = vF;
        sm->inbox[(q & (kInbox - 1)) + 3 * kInbox] = vF;
        if (lane == 0 && (i0 & 255) == 0)
          st_progress(my_progress, ((unsigned long long)(band + 1) << 32) | (unsigned long long)(i0 + SLACK + kChunk));
      }
      if (has_sink && !last_local && (i0 & 1023) == 0 && i0 + 1024 > (long long)out_mask + 1) {
        const unsigned long long need =
            ((unsigned long long)(band + 2) << 32) | (unsigned long long)(i0 + 1024 - ((long long)out_mask + 1));
        while (!wt.aborted && ld_progress(sink_progress) < need) {
          if (--wt.budget < 0) { atomic_or_i32(P.result + 1, STATUS_SPIN_TIMEOUT); wt.aborted = true; }
          spin_pause();
        }
      }
      w.sync();

      const uint32_t* tabp = sm->tab + ((i0 - (long long)SK * lane) & (kTabRing - 1));
      const uint32_t* inbp = sm->inbox + ((i0 + SLACK) & (kInbox - 1));
      uint2* outp = out + 2 * (i0 & out_mask);
      const uint32_t otag = out_tag | ((uint32_t)(i0 >> out_shift) & 0xFFu);

#pragma unroll 4
      for (int k = 0; k < kChunk; ++k) {
        const uint32_t Tw = tabp[k];
        const int xinH = (int)inbp[k], xinF = (int)inbp[k + 2 * kInbox];
        const int ynewH = (int)w.shfl((uint32_t)(last_lane ? xinH : xsH), src_lane);
        const int ynewF = (int)w.shfl((uint32_t)(last_lane ? xinF : xsF), src_lane);
        const int upHo = SLACK ? yoldH : ynewH;
        int F = SLACK ? yoldF : ynewF;
        yoldH = ynewH; yoldF = ynewF;
        int diag = up_prev, Hup = upHo;
        up_prev = upHo;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int s = (int)prmt(Tw, 0u, sel[r]);
          const int d = diag + s;
          const int old = Ho[r];
          E[r] = addmax32(E[r], next, old);
          F = addmax32(F, next, Hup);
          const int h = max3relu32(d, E[r], F);
          Ho[r] = h + nopen;
          Hup = Ho[r];
          diag = old;
          if (r & 1) best1 = best1 > h ? best1 : h; else best0 = best0 > h ? best0 : h;
        }
        xsH = Ho[R - 1];
        xsF = F;
        if (emit) {
          st_entry(outp + 2 * k, (uint32_t)xsH, otag);
          st_entry(outp + 2 * k + 1, (uint32_t)xsF, otag);
        }
      }
    }
  }
  int m = w.reduce_max(best0 > best1 ? best0 : best1);
  if (lane == 0) atomic_max_i32(P.result, m);
}

// rows of Q one band covers
SWB_HD int rows_per_band(int R, int mode) { return (mode == 2 ? 32 : 64) * R; }

}  // namespace swb
