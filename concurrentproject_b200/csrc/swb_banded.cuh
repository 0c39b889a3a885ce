// swb_banded.cuh -- banded Gotoh local alignment for batches of long pairs (BASELINE config 5).
//
// Semantics (oracle: oracle_gotoh_banded): cell (i,j), 1-based, is in band iff
// band_lo <= j-i <= band_lo+63; out-of-band cells are H=E=F=0 and excluded from the max, i.e. main.cpp:57-63
// evaluated on in-band cells only.  The reference has no banded mode.
//
// Geometry.  The 64 diagonals k = j-i of a pair are spread over 16 threads, 4 adjacent diagonals each:
// the two with the parity of band_lo in one packed register set ("A"), the other two in a second ("B").
// The sweep goes over anti-diagonals t = i+j; on one anti-diagonal only the diagonals with k = t (mod 2)
// have a cell, so steps alternate A, B, A, B and every packed instruction does two useful cells.  A cell's
// left neighbour is diagonal k-1 and its upper neighbour diagonal k+1, both on the previous anti-diagonal --
// i.e. always in the OTHER register set of the same thread, except the outermost halves, which travel by one
// __shfl_sync per step.  The diagonal neighbour is the cell's own register two steps ago.  Two pairs per warp.
// Substitution scores: one PRMT per cell vector from two shared-memory rings filled cooperatively per pair,
// a 4-byte score table per column symbol (as in the pair engine) and a PRMT selector per row pair.
#pragma once
#include "swb_engine.cuh"

namespace swb {

constexpr int kBandRing = 128;     // ring entries (tables per column / selectors per row), kept twice
constexpr int kBandThreads = 16;   // threads per pair
constexpr int kBandWidth = 64;     // diagonals per pair

struct BandedParams {
  const uint64_t* a_words;   // seq1 (columns j), 2-bit packed, a_stride words per pair
  const uint64_t* b_words;   // seq2 (rows i)
  const int* a_len;
  const int* b_len;
  long long a_stride, b_stride;
  long long npairs;
  int band_lo;               // band is band_lo <= j-i <= band_lo+63
  int* scores;
  int match, mismatch, gap_init, gap_ext;
};

struct BandedWarpSmem {
  uint32_t tab[2][2 * kBandRing];   // [pair in warp][ring]: score table of column x at slot x & 127
  uint32_t sel[2][2 * kBandRing];   // selector of rows (y, y-1) at slot y & 127
};

SWB_HD uint32_t packed_code(const uint64_t* words, int pos, int len) {
  return (pos >= 0 && pos < len) ? (uint32_t)(words[pos >> 5] >> (2 * (pos & 31))) & 3u : 4u;
}

template <int MODE>
SWB_HD void banded_warp(const BandedParams& P, const WarpCtx& w, long long warp_id, long long num_warps, BandedWarpSmem* sm) {
  const int lane = w.lane;
  const int u = lane & 15, grp = lane >> 4;
  const int src_prev = grp * 16 + ((u + 15) & 15);     // A-step: value comes from thread u-1 (rotating)
  const int src_next = grp * 16 + ((u + 1) & 15);      // B-step: value comes from thread u+1 (rotating)
  const uint32_t nopen = pack2(-P.gap_init), next = pack2(-P.gap_ext);
  const uint32_t padb = (uint32_t)(P.mismatch + P.gap_init) & 0xFFu;
  const uint32_t padw = padb * 0x01010101u;
  const uint32_t flip = padb ^ ((uint32_t)(P.match + P.gap_init) & 0xFFu);
  const uint32_t padsel = mk_sel16(4u, 4u);
  uint32_t* tab = sm->tab[grp];
  uint32_t* sel = sm->sel[grp];
  const long long ngroups = (P.npairs + 1) / 2;
  // first anti-diagonal handled by register set A: the largest t <= 2 with t = band_lo (mod 2)
  const int tA0 = 2 - ((2 - P.band_lo) & 1);
  const int I0 = (tA0 - P.band_lo) / 2, J0 = (tA0 + P.band_lo) / 2;      // exact: both numerators are even

  for (long long pg = warp_id; pg < ngroups; pg += num_warps) {
    const long long pair = pg * 2 + grp;
    const bool valid = pair < P.npairs;
    const int n = valid ? P.a_len[pair] : 0, m = valid ? P.b_len[pair] : 0;
    const uint64_t* aw = P.a_words + (valid ? pair : 0) * P.a_stride;
    const uint64_t* bw = P.b_words + (valid ? pair : 0) * P.b_stride;
    const int NH = (n > 0 && m > 0) ? (n + m - tA0) / 2 + 1 : 0;         // double steps (one A + one B anti-diagonal)
    const int maxNH = w.reduce_max(NH);
    const int nchunks = (maxNH + kChunk - 1) / kChunk;

    // state of the 4 diagonals: H-open, E, F of register sets A and B; all cells start as border (H 0, E/F <= 0)
    uint32_t HoA = nopen, EA = nopen, FA = nopen, HoB = nopen, EB = nopen, FB = nopen;
    uint32_t best = 0;
    // 0-based sequence positions of this thread's A.lo cell at double step h: row y = Y0 + h, column x = X0 + h
    const int Y0 = I0 - 2 * u - 1, X0 = J0 + 2 * u - 1;

    // ---- rings: everything pad, then selectors for rows [Yc-32, Yc) and tables for columns [Xc, Xc+32) of chunk 0
    const int Yc0 = I0 - 1, Xc0 = J0 - 1;                                  // thread 0's positions at h = 0
    w.sync();
    for (int k = u; k < 2 * kBandRing; k += 16) { tab[k] = padw; sel[k] = padsel; }
    w.sync();
    for (int k = u; k < kChunk; k += 16) {
      const int y = Yc0 - kChunk + k;
      const uint32_t sv = mk_sel16(packed_code(bw, y, m), packed_code(bw, y - 1, m));
      sel[y & (kBandRing - 1)] = sv; sel[(y & (kBandRing - 1)) + kBandRing] = sv;
      const int x = Xc0 + k;
      const uint32_t tv = table_word(packed_code(aw, x, n), padw, flip);
      tab[x & (kBandRing - 1)] = tv; tab[(x & (kBandRing - 1)) + kBandRing] = tv;
    }

    for (int c = 0; c < nchunks; ++c) {
      // ---- this chunk reads rows [Yc-30, Yc+32) and columns [Xc, Xc+64): add rows [Yc, Yc+32), columns [Xc+32, Xc+64)
      const int Yc = Yc0 + c * kChunk, Xc = Xc0 + c * kChunk;
      for (int k = u; k < kChunk; k += 16) {
        const int y = Yc + k;
        const uint32_t sv = mk_sel16(packed_code(bw, y, m), packed_code(bw, y - 1, m));
        sel[y & (kBandRing - 1)] = sv; sel[(y & (kBandRing - 1)) + kBandRing] = sv;
        const int x = Xc + kChunk + k;
        const uint32_t tv = table_word(packed_code(aw, x, n), padw, flip);
        tab[x & (kBandRing - 1)] = tv; tab[(x & (kBandRing - 1)) + kBandRing] = tv;
      }
      w.sync();
      const uint32_t* selp = sel + ((Y0 + c * kChunk) & (kBandRing - 1));
      const uint32_t* tabp = tab + ((X0 + c * kChunk) & (kBandRing - 1));
      uint32_t T0 = tabp[0], T1 = tabp[1];
#pragma unroll 4
      for (int h = 0; h < kChunk; ++h) {
        const uint32_t sv = selp[h];
        const uint32_t T2 = tabp[h + 2];
        const uint32_t sA = prmt(T0, T1, sv);        // A.lo: row y, column x ; A.hi: row y-1, column x+1
        const uint32_t sB = prmt(T1, T2, sv);        // B.lo: row y, column x+1 ; B.hi: row y-1, column x+2
        T0 = T1; T1 = T2;
        // ---------------- anti-diagonal of set A: left neighbours are in B (lo: thread u-1's B.hi), upper = B as is
        {
          uint32_t leftHo, h2;
          if (MODE == 0) {
            const uint32_t xs = (u == 15) ? nopen : prmt(HoB, EB, 0x7632u);       // (H-open, E) of B.hi
            const uint32_t rv = w.shfl(xs, src_prev);
            leftHo = prmt(rv, HoB, 0x5410u);
            const uint32_t leftE = prmt(rv, EB, 0x5432u);
            const uint32_t E = addmax16x2(leftE, next, leftHo);
            const uint32_t F = addmax16x2(FB, next, HoB);
            const uint32_t d = add16x2(HoA, sA);
            h2 = max3relu16x2(d, E, F);
            EA = E; FA = F;
          } else {
            const uint32_t xs = (u == 15) ? nopen : HoB;
            const uint32_t rv = w.shfl(xs, src_prev);
            leftHo = prmt(rv, HoB, 0x5432u);
            h2 = max16x2(addmaxrelu16x2(HoA, sA, leftHo), HoB);
          }
          HoA = add16x2(h2, nopen);
          best = max16x2(best, h2);
        }
        // ---------------- anti-diagonal of set B: left = A as is, upper neighbours are in A (hi: thread u+1's A.lo)
        {
          uint32_t upHo, h2;
          if (MODE == 0) {
            const uint32_t xs = (u == 0) ? nopen : prmt(HoA, FA, 0x5410u);        // (H-open, F) of A.lo
            const uint32_t rv = w.shfl(xs, src_next);
            upHo = prmt(HoA, rv, 0x5432u);
            const uint32_t upF = prmt(FA, rv, 0x7632u);
            const uint32_t E = addmax16x2(EA, next, HoA);
            const uint32_t F = addmax16x2(upF, next, upHo);
            const uint32_t d = add16x2(HoB, sB);
            h2 = max3relu16x2(d, E, F);
            EB = E; FB = F;
          } else {
            const uint32_t xs = (u == 0) ? nopen : HoA;
            const uint32_t rv = w.shfl(xs, src_next);
            upHo = prmt(HoA, rv, 0x5432u);
            h2 = max16x2(addmaxrelu16x2(HoB, sB, HoA), upHo);
          }
          HoB = add16x2(h2, nopen);
          best = max16x2(best, h2);
        }
      }
    }
    int mx = (int)(short)(best & 0xFFFFu), mh = (int)(short)(best >> 16);
    mx = mx > mh ? mx : mh;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
      const int o = (int)w.shfl((uint32_t)mx, lane ^ d);
      mx = mx > o ? mx : o;
    }
    if (valid && u == 0) P.scores[pair] = mx;
  }
}

#ifdef __CUDACC__
template <int MODE>
__global__ void __launch_bounds__(256, 2) sw_banded_kernel(const __grid_constant__ BandedParams P) {
  __shared__ BandedWarpSmem sm[8];
  WarpCtx w{(int)(threadIdx.x & 31)};
  const int wi = (int)(threadIdx.x >> 5);
  banded_warp<MODE>(P, w, (long long)blockIdx.x * 8 + wi, (long long)gridDim.x * 8, &sm[wi]);
}
#endif

const void* banded_kernel(int mode);

}  // namespace swb
