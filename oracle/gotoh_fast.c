/*
 * gotoh_fast.c -- the SAME recurrence as oracle_gotoh_rolling / main.cpp:57-63, arranged so that a CPU can
 * finish the pairs the reference cannot reach (BASELINE config 3: 4 000 000 x 4 000 000 = 1.6e13 cells,
 * ~46 h for the reference's LazySmith).
 *
 * TEST INFRASTRUCTURE ONLY (see gotoh_oracle.h).  It exists to PIN golden scores under tests/golden/ for sizes
 * where only analytic checks existed; tests/test_oracle.py proves it equal to oracle_gotoh_rolling (and through
 * it to the unmodified reference) on every size both can run, ragged tile edges included.
 *
 * How: the DP matrix is cut into TB x TB tiles; the tiles of one tile-anti-diagonal are independent, so 16 of
 * them are computed at once, ONE TILE PER SIMD LANE (int32 lanes, 512-bit vectors when the host has them; the
 * code is plain GCC vector extensions and compiles for any target).  Every lane runs exactly the scalar loop of
 * tile_run() in gotoh_oracle.c -- same order, same int32 arithmetic -- on its own tile, so there is nothing to
 * prove about lane interaction: lanes never talk to each other.  Groups of 16 tiles are spread over threads.
 * Tiles hanging over the matrix edge are padded with symbols that match nothing; a padded cell can only extend
 * an alignment by mismatches and gaps (score never above its last real cell) and only feeds other padded cells,
 * and the running maximum is masked to real cells anyway.
 */
#include "gotoh_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FL 16           /* tiles per SIMD group (lanes) */
#define FTB 256         /* tile edge */

typedef int v16 __attribute__((vector_size(FL * sizeof(int))));

static inline __attribute__((always_inline)) v16 vmax(v16 a, v16 b) {
  const v16 m = a > b;              /* -1 where a > b */
  return (a & m) | (b & ~m);
}
static inline __attribute__((always_inline)) v16 vsplat(int x) {
  v16 r;
  for (int l = 0; l < FL; ++l) r[l] = x;
  return r;
}

typedef struct {
  v16 s1[FTB];          /* column symbols of the 16 tiles (as ints; -1 = beyond the edge) */
  v16 hr[FTB + 1];      /* H of the row above / of the last row, index jj+1; [0] unused */
  v16 fr[FTB + 1];
} group_buf;

/* one group: lanes l = 0..cnt-1 are tiles (bi[l], d - bi[l]) */
__attribute__((target_clones("avx512f", "avx2", "default")))
static int group_run(const unsigned char* seq1, const unsigned char* seq2, int n, int m, int d, const int* bi, int cnt,
                     int* Hrow, int* Frow, int* Hcol, int* Ecol, int* corner, const oracle_params* p, group_buf* g) {
  int i0[FL], j0[FL];
  for (int l = 0; l < FL; ++l) {
    const int b = bi[l < cnt ? l : 0];              /* idle lanes recompute lane 0's tile; results discarded */
    i0[l] = b * FTB;
    j0[l] = (d - b) * FTB;
  }
  v16 diag0, next_corner;
  for (int l = 0; l < FL; ++l) {
    diag0[l] = corner[i0[l] / FTB];
    const int jl = j0[l] + FTB < n ? j0[l] + FTB : n;
    next_corner[l] = Hrow[jl];
    for (int jj = 0; jj < FTB; ++jj) {
      const int j = j0[l] + jj;
      g->s1[jj][l] = j < n ? (int)seq1[j] : -1;
      g->hr[jj + 1][l] = j < n ? Hrow[j + 1] : 0;
      g->fr[jj + 1][l] = j < n ? Frow[j + 1] : 0;
    }
  }
  const v16 ge = vsplat(p->gap_ext), gi = vsplat(p->gap_init), zero = vsplat(0);
  const v16 vmis = vsplat(p->mismatch), vdelta = vsplat(p->match - p->mismatch);
  v16 best = zero;
  for (int ii = 0; ii < FTB; ++ii) {
    v16 b, e, hleft, rowok;
    for (int l = 0; l < FL; ++l) {
      const int i = i0[l] + ii;
      b[l] = i < m ? (int)seq2[i] : -2;
      e[l] = i < m ? Ecol[i + 1] : 0;
      hleft[l] = i < m ? Hcol[i + 1] : 0;
      rowok[l] = i < m ? -1 : 0;
    }
    v16 hdiag = diag0;
    diag0 = hleft;
    for (int jj = 0; jj < FTB; ++jj) {
      e = vmax(e - ge, hleft - gi);                                 /* main.cpp:57 */
      const v16 f = vmax(g->fr[jj + 1] - ge, g->hr[jj + 1] - gi);   /* main.cpp:58 */
      const v16 eq = g->s1[jj] == b;
      v16 h = hdiag + vmis + (vdelta & eq);                         /* main.cpp:28-33, 62 */
      h = vmax(h, e);
      h = vmax(h, f);
      h = vmax(h, zero);                                            /* main.cpp:61-63 */
      hdiag = g->hr[jj + 1];
      g->hr[jj + 1] = h;
      g->fr[jj + 1] = f;
      hleft = h;
      const v16 colok = g->s1[jj] >= zero;
      best = vmax(best, h & rowok & colok);
    }
    for (int l = 0; l < cnt; ++l) {
      const int i = i0[l] + ii;
      if (i < m) { Hcol[i + 1] = hleft[l]; Ecol[i + 1] = e[l]; }
    }
  }
  int out = 0;
  for (int l = 0; l < cnt; ++l) {
    for (int jj = 0; jj < FTB; ++jj) {
      const int j = j0[l] + jj;
      if (j < n) { Hrow[j + 1] = g->hr[jj + 1][l]; Frow[j + 1] = g->fr[jj + 1][l]; }
    }
    corner[i0[l] / FTB] = next_corner[l];
    if (best[l] > out) out = best[l];
  }
  return out;
}

int oracle_gotoh_fast(const unsigned char* seq1, const unsigned char* seq2, int n, int m, const oracle_params* p,
                      int threads) {
  if (n < 0 || m < 0) return -1;
  if (n == 0 || m == 0) return 0;
  const int nbj = (n + FTB - 1) / FTB, nbi = (m + FTB - 1) / FTB;
  int* Hrow = (int*)calloc((size_t)n + 1, sizeof(int));
  int* Frow = (int*)calloc((size_t)n + 1, sizeof(int));
  int* Hcol = (int*)calloc((size_t)m + 1, sizeof(int));
  int* Ecol = (int*)calloc((size_t)m + 1, sizeof(int));
  int* corner = (int*)calloc((size_t)nbi, sizeof(int));
  if (!Hrow || !Frow || !Hcol || !Ecol || !corner) { free(Hrow); free(Frow); free(Hcol); free(Ecol); free(corner); return -1; }
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  int best = 0;
#pragma omp parallel num_threads(threads) reduction(max : best)
  {
    group_buf* g = (group_buf*)aligned_alloc(64, sizeof(group_buf));
    for (long long d = 0; d < (long long)nbi + nbj - 1; ++d) {
      const int lo = d - (nbj - 1) > 0 ? (int)(d - (nbj - 1)) : 0;
      const int hi = d < nbi - 1 ? (int)d : nbi - 1;
      const int ngroups = (hi - lo + 1 + FL - 1) / FL;
#pragma omp for schedule(dynamic, 1)
      for (int gidx = 0; gidx < ngroups; ++gidx) {
        int bi[FL];
        int cnt = 0;
        for (int l = 0; l < FL; ++l) {
          const int b = lo + gidx * FL + l;
          if (b <= hi) bi[cnt++] = b;
        }
        const int r = group_run(seq1, seq2, n, m, (int)d, bi, cnt, Hrow, Frow, Hcol, Ecol, corner, p, g);
        if (r > best) best = r;
      }   /* implicit barrier: the next tile diagonal reads what this one wrote */
    }
    free(g);
  }
  free(Hrow); free(Frow); free(Hcol); free(Ecol); free(corner);
  return best;
}
