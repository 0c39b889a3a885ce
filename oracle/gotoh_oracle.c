/*
 * gotoh_oracle.c -- CPU restatement of the reference's score-only Smith-Waterman/Gotoh path.
 *
 * TEST INFRASTRUCTURE ONLY (see gotoh_oracle.h).  Parity status: PINNED against the
 * reference's golden vectors and against oracle/_ref/libref.so (tests/test_oracle.py).
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC).
 */
#include "gotoh_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int imax(int a, int b) { return a > b ? a : b; }

/* main.cpp:28-33 (score), lazySmith.cpp:11-13 (score2): equality on raw bytes. */
static inline int sub_score(unsigned char a, unsigned char b, const oracle_params* p) {
  return a == b ? p->match : p->mismatch;
}

/* ---- main.cpp:40-90 ------------------------------------------------------------------ */
int oracle_gotoh_full(const unsigned char* seq1, const unsigned char* seq2, int n, int m,
                      const oracle_params* p) {
  if (n < 0 || m < 0) return -1;
  const size_t W = (size_t)n + 1, Hh = (size_t)m + 1;
  int* E = (int*)calloc(W * Hh, sizeof(int));
  int* F = (int*)calloc(W * Hh, sizeof(int));
  int* H = (int*)calloc(W * Hh, sizeof(int));
  if (!E || !F || !H) { free(E); free(F); free(H); return -1; }
  /* borders are zero (main.cpp:43-52) -- calloc did that */
  for (int i = 1; i <= m; ++i) {
    for (int j = 1; j <= n; ++j) {
      const size_t c = (size_t)i * W + j;
      E[c] = imax(E[c - 1] - p->gap_ext, H[c - 1] - p->gap_init);      /* main.cpp:57 */
      F[c] = imax(F[c - W] - p->gap_ext, H[c - W] - p->gap_init);      /* main.cpp:58 */
      int t1 = imax(F[c], E[c]);                                        /* main.cpp:61 */
      int t2 = imax(0, H[c - W - 1] + sub_score(seq1[j - 1], seq2[i - 1], p)); /* :62 */
      H[c] = imax(t1, t2);                                              /* main.cpp:63 */
    }
  }
  int best = 0;                                                         /* main.cpp:82-87 */
  for (size_t c = 0; c < W * Hh; ++c) best = imax(best, H[c]);
  free(E); free(F); free(H);
  return best;
}

/* ---- lazySmith.cpp:27-41 (pass 1 only = exact Gotoh with rolling rows) ------------------ */
int oracle_gotoh_rolling(const unsigned char* seq1, const unsigned char* seq2, int n, int m,
                         const oracle_params* p) {
  if (n < 0 || m < 0) return -1;
  int* Hrow = (int*)calloc((size_t)n + 1, sizeof(int));   /* H[i-1][*] then H[i][*] in place */
  int* Frow = (int*)calloc((size_t)n + 1, sizeof(int));
  if (!Hrow || !Frow) { free(Hrow); free(Frow); return -1; }
  const int ge = p->gap_ext, gi = p->gap_init;
  int best = 0;
  for (int i = 1; i <= m; ++i) {
    const unsigned char b = seq2[i - 1];
    int e = 0, hleft = 0, hdiag = 0;   /* E[i][0], H[i][0], H[i-1][0] */
    for (int j = 1; j <= n; ++j) {
      e = imax(e - ge, hleft - gi);
      const int f = imax(Frow[j] - ge, Hrow[j] - gi);
      int h = hdiag + (seq1[j - 1] == b ? p->match : p->mismatch);
      if (e > h) h = e;
      if (f > h) h = f;
      if (h < 0) h = 0;
      hdiag = Hrow[j];
      Hrow[j] = h;
      Frow[j] = f;
      hleft = h;
      if (h > best) best = h;
    }
  }
  free(Hrow); free(Frow);
  return best;
}

/* main.cpp:57-63 with rolling rows, remembering where the maximum sits (rule in the header). */
int oracle_gotoh_end(const unsigned char* seq1, const unsigned char* seq2, int n, int m, const oracle_params* p,
                     int* i_end, int* j_end) {
  *i_end = 0; *j_end = 0;
  if (n < 0 || m < 0) return -1;
  int* Hrow = (int*)calloc((size_t)n + 1, sizeof(int));
  int* Frow = (int*)calloc((size_t)n + 1, sizeof(int));
  if (!Hrow || !Frow) { free(Hrow); free(Frow); return -1; }
  const int ge = p->gap_ext, gi = p->gap_init;
  int best = 0, bi = 0, bj = 0;
  for (int i = 1; i <= m; ++i) {
    const unsigned char b = seq2[i - 1];
    int e = 0, hleft = 0, hdiag = 0;
    for (int j = 1; j <= n; ++j) {
      e = imax(e - ge, hleft - gi);
      const int f = imax(Frow[j] - ge, Hrow[j] - gi);
      int h = hdiag + (seq1[j - 1] == b ? p->match : p->mismatch);
      if (e > h) h = e;
      if (f > h) h = f;
      if (h < 0) h = 0;
      hdiag = Hrow[j];
      Hrow[j] = h;
      Frow[j] = f;
      hleft = h;
      if (h > best || (h == best && h > 0 && (j < bj || (j == bj && i < bi)))) { best = h; bi = i; bj = j; }
    }
  }
  free(Hrow); free(Frow);
  *i_end = bi; *j_end = bj;
  return best;
}

/* Anchored variant of main.cpp:57-63 (every alignment starts at cell (0,0); no clamp at 0; the borders carry the
 * gap costs as the recurrence itself generates them): score and position of the maximum, same tie rule.
 * On the reversed prefixes ending at an alignment's end cell this finds the alignment's start. */
int oracle_gotoh_anchored_end(const unsigned char* seq1, const unsigned char* seq2, int n, int m, const oracle_params* p,
                              int* i_end, int* j_end) {
  *i_end = 0; *j_end = 0;
  if (n < 0 || m < 0) return -1;
  const long long NEGINF = -(1LL << 40);
  long long* Hrow = (long long*)calloc((size_t)n + 1, sizeof(long long));
  long long* Frow = (long long*)calloc((size_t)n + 1, sizeof(long long));
  if (!Hrow || !Frow) { free(Hrow); free(Frow); return -1; }
  const int ge = p->gap_ext, gi = p->gap_init;
  Hrow[0] = 0; Frow[0] = NEGINF;
  { long long e = NEGINF; for (int j = 1; j <= n; ++j) { e = (e - ge > Hrow[j - 1] - gi) ? e - ge : Hrow[j - 1] - gi; Hrow[j] = e; Frow[j] = NEGINF; } }
  long long best = 0; int bi = 0, bj = 0;
  for (int i = 1; i <= m; ++i) {
    const unsigned char b = seq2[i - 1];
    long long hdiag = Hrow[0];
    const long long f0 = (Frow[0] - ge > Hrow[0] - gi) ? Frow[0] - ge : Hrow[0] - gi;   /* column 0: only F */
    Hrow[0] = f0; Frow[0] = f0;
    long long e = NEGINF, hleft = Hrow[0];
    for (int j = 1; j <= n; ++j) {
      e = (e - ge > hleft - gi) ? e - ge : hleft - gi;
      const long long f = (Frow[j] - ge > Hrow[j] - gi) ? Frow[j] - ge : Hrow[j] - gi;
      long long h = hdiag + (seq1[j - 1] == b ? p->match : p->mismatch);
      if (e > h) h = e;
      if (f > h) h = f;
      hdiag = Hrow[j];
      Hrow[j] = h;
      Frow[j] = f;
      hleft = h;
      if (h > best || (h == best && h > 0 && (j < bj || (j == bj && i < bi)))) { best = h; bi = i; bj = j; }
    }
  }
  free(Hrow); free(Frow);
  *i_end = bi; *j_end = bj;
  return (int)best;
}

/* Score, end cell and start cell of the best local alignment: end cell by the rule of oracle_gotoh_end; start cell =
 * the cell where the anchored pass over the reversed prefixes (seq1[0..j_end), seq2[0..i_end)) reaches the score
 * first (same rule in reversed coordinates: the start closest to the end in seq1, then in seq2).  1-based,
 * both inclusive; all zero for score 0. */
int oracle_gotoh_span(const unsigned char* seq1, const unsigned char* seq2, int n, int m, const oracle_params* p,
                      int* i_start, int* j_start, int* i_end, int* j_end) {
  *i_start = *j_start = 0;
  const int s = oracle_gotoh_end(seq1, seq2, n, m, p, i_end, j_end);
  if (s <= 0) return s;
  unsigned char* r1 = (unsigned char*)malloc((size_t)*j_end + 1);
  unsigned char* r2 = (unsigned char*)malloc((size_t)*i_end + 1);
  if (!r1 || !r2) { free(r1); free(r2); return -1; }
  for (int k = 0; k < *j_end; ++k) r1[k] = seq1[*j_end - 1 - k];
  for (int k = 0; k < *i_end; ++k) r2[k] = seq2[*i_end - 1 - k];
  int ri = 0, rj = 0;
  const int s2 = oracle_gotoh_anchored_end(r1, r2, *j_end, *i_end, p, &ri, &rj);
  free(r1); free(r2);
  if (s2 != s) return -2;                      /* cannot happen: the best alignment ends at (i_end, j_end) */
  *i_start = *i_end - ri + 1;
  *j_start = *j_end - rj + 1;
  return s;
}

int oracle_gotoh_last_row(const unsigned char* seq1, const unsigned char* seq2, int n, int m,
                          const oracle_params* p, int* Hrow_out, int* Frow_out) {
  if (n < 0 || m < 0) return -1;
  int* Hrow = (int*)calloc((size_t)n + 1, sizeof(int));
  int* Frow = (int*)calloc((size_t)n + 1, sizeof(int));
  if (!Hrow || !Frow) { free(Hrow); free(Frow); return -1; }
  const int ge = p->gap_ext, gi = p->gap_init;
  int best = 0;
  for (int i = 1; i <= m; ++i) {
    const unsigned char b = seq2[i - 1];
    int e = 0, hleft = 0, hdiag = 0;
    for (int j = 1; j <= n; ++j) {
      e = imax(e - ge, hleft - gi);
      const int f = imax(Frow[j] - ge, Hrow[j] - gi);
      int h = hdiag + (seq1[j - 1] == b ? p->match : p->mismatch);
      if (e > h) h = e;
      if (f > h) h = f;
      if (h < 0) h = 0;
      hdiag = Hrow[j];
      Hrow[j] = h;
      Frow[j] = f;
      hleft = h;
      if (h > best) best = h;
    }
  }
  memcpy(Hrow_out, Hrow, ((size_t)n + 1) * sizeof(int));
  memcpy(Frow_out, Frow, ((size_t)n + 1) * sizeof(int));
  free(Hrow); free(Frow);
  return best;
}

/* ---- lazySmith.cpp:15-69 including the lazy-F pass -------------------------------------- */
int oracle_lazy_smith(const unsigned char* seq1, const unsigned char* seq2, int n, int m,
                      const oracle_params* p) {
  if (n < 0 || m < 0) return -1;
  const size_t W = (size_t)n + 1;
  int* Hp = (int*)calloc(W, sizeof(int));
  int* Hc = (int*)calloc(W, sizeof(int));
  int* E = (int*)calloc(W, sizeof(int));
  int* F = (int*)calloc(W, sizeof(int));
  if (!Hp || !Hc || !E || !F) { free(Hp); free(Hc); free(E); free(F); return -1; }
  const int ge = p->gap_ext, gi = p->gap_init;
  int best = 0;
  for (int i = 1; i <= m; ++i) {
    E[0] = 0; F[0] = 0; Hc[0] = 0;                                   /* :23-25 */
    for (int j = 1; j <= n; ++j) {                                   /* :27-41 */
      E[j] = imax(E[j - 1] - ge, Hc[j - 1] - gi);
      int ht = Hp[j - 1] + sub_score(seq1[j - 1], seq2[i - 1], p);
      F[j] = imax(F[j] - ge, Hp[j] - gi);
      int h = ht;
      if (E[j] > h) h = E[j];
      if (F[j] > h) h = F[j];
      if (h < 0) h = 0;
      Hc[j] = h;
      if (h > best) best = h;
    }
    for (int j = 1; j <= n; ++j) {                                   /* :43-62 */
      int fij = Hc[j] - gi;
      if (fij > F[j]) {
        F[j] = fij;
        if (F[j] > Hc[j]) { Hc[j] = F[j]; if (Hc[j] > best) best = Hc[j]; }
        for (int k = j + 1; k <= n; ++k) {
          int nf = F[k - 1] - ge;
          if (nf <= 0) break;
          if (nf <= F[k]) break;
          F[k] = nf;
          if (F[k] > Hc[k]) { Hc[k] = F[k]; if (Hc[k] > best) best = Hc[k]; }
        }
      }
    }
    int* t = Hp; Hp = Hc; Hc = t;                                    /* :64-65 */
    memset(Hc, 0, W * sizeof(int));
  }
  free(Hp); free(Hc); free(E); free(F);
  return best;
}

/* ---- SmithDiagonalGPU.cu:40-67 recurrence (linear gap, H only) --------------------------- */
int oracle_linear_gap(const unsigned char* seq1, const unsigned char* seq2, int n, int m,
                      const oracle_params* p) {
  if (n < 0 || m < 0) return -1;
  int* Hrow = (int*)calloc((size_t)n + 1, sizeof(int));
  if (!Hrow) return -1;
  int best = 0;
  for (int i = 1; i <= m; ++i) {
    int hleft = 0, hdiag = 0;
    for (int j = 1; j <= n; ++j) {
      int h = hdiag + sub_score(seq1[j - 1], seq2[i - 1], p);
      h = imax(h, Hrow[j] - p->gap_init);
      h = imax(h, hleft - p->gap_init);
      h = imax(h, 0);
      hdiag = Hrow[j];
      Hrow[j] = h;
      hleft = h;
      best = imax(best, h);
    }
  }
  free(Hrow);
  return best;
}

/* ---- banded Gotoh: main.cpp:57-63 on in-band cells only ---------------------------------- */
int oracle_gotoh_banded(const unsigned char* seq1, const unsigned char* seq2, int n, int m,
                        int band_lo, int band_hi, const oracle_params* p, int64_t* cells_out) {
  if (n < 0 || m < 0 || band_lo > band_hi) return -1;
  /* Rolling rows indexed by column j.  The band moves right by one column per row, so an
   * entry that is outside the band of row i-1 has either never been written (still 0 from
   * calloc) or is never read again; out-of-band cells therefore read as H=E=F=0. */
  int* Hrow = (int*)calloc((size_t)n + 2, sizeof(int));
  int* Frow = (int*)calloc((size_t)n + 2, sizeof(int));
  if (!Hrow || !Frow) { free(Hrow); free(Frow); return -1; }
  const int ge = p->gap_ext, gi = p->gap_init;
  int best = 0;
  int64_t cells = 0;
  for (int i = 1; i <= m; ++i) {
    long jlo = (long)i + band_lo, jhi = (long)i + band_hi;
    if (jlo < 1) jlo = 1;
    if (jhi > n) jhi = n;
    if (jlo > jhi) continue;
    int e = 0, hleft = 0;                       /* (i, jlo-1) is out of band or border: 0 */
    int hdiag = (jlo > 1) ? Hrow[jlo - 1] : 0;  /* (i-1, jlo-1): first in-band cell of row i-1, or border */
    for (long j = jlo; j <= jhi; ++j) {
      e = imax(e - ge, hleft - gi);                       /* main.cpp:57 */
      const int f = imax(Frow[j] - ge, Hrow[j] - gi);     /* main.cpp:58 */
      int h = hdiag + sub_score(seq1[j - 1], seq2[i - 1], p);
      if (e > h) h = e;
      if (f > h) h = f;
      if (h < 0) h = 0;                                   /* main.cpp:61-63 */
      hdiag = Hrow[j];
      Hrow[j] = h;
      Frow[j] = f;
      hleft = h;
      if (h > best) best = h;
      ++cells;
    }
  }
  free(Hrow); free(Frow);
  if (cells_out) *cells_out = cells;
  return best;
}

/* ---- tile-blocked multi-threaded exact Gotoh ---------------------------------------------- */
typedef struct { int best; } tile_out;

static int tile_run(const unsigned char* seq1, const unsigned char* seq2, int j0, int j1, int i0, int i1,
                    int* Hrow, int* Frow, int* Hcol, int* Ecol, int* corner, const oracle_params* p) {
  /* rows i0..i1-1 (0-based), columns j0..j1-1 (0-based).
   * Hrow/Frow[j+1]: H/F of the row above the tile on entry, of the tile's last row on exit.
   * Hcol/Ecol[i+1]: H/E of the column left of the tile on entry, of the tile's last column on exit.
   * *corner: H[i0][j0] (1-based border indices) on entry; H[i0][j1] on exit (for the tile to the right). */
  const int ge = p->gap_ext, gi = p->gap_init;
  int best = 0;
  int diag0 = *corner;
  const int next_corner = Hrow[j1];   /* H above the tile at its last column */
  for (int i = i0; i < i1; ++i) {
    const unsigned char b = seq2[i];
    int e = Ecol[i + 1], hleft = Hcol[i + 1];
    int hdiag = diag0;
    diag0 = hleft;                    /* H[i][j0-1] is the diagonal of row i+1 */
    for (int j = j0; j < j1; ++j) {
      e = imax(e - ge, hleft - gi);
      const int f = imax(Frow[j + 1] - ge, Hrow[j + 1] - gi);
      int h = hdiag + (seq1[j] == b ? p->match : p->mismatch);
      if (e > h) h = e;
      if (f > h) h = f;
      if (h < 0) h = 0;
      hdiag = Hrow[j + 1];
      Hrow[j + 1] = h;
      Frow[j + 1] = f;
      hleft = h;
      if (h > best) best = h;
    }
    Hcol[i + 1] = hleft;
    Ecol[i + 1] = e;
  }
  *corner = next_corner;
  return best;
}

int oracle_gotoh_mt(const unsigned char* seq1, const unsigned char* seq2, int n, int m,
                    const oracle_params* p, int threads) {
  if (n < 0 || m < 0) return -1;
  if (n == 0 || m == 0) return 0;
  const int TB = 1024;
  const int nbj = (n + TB - 1) / TB, nbi = (m + TB - 1) / TB;
  int* Hrow = (int*)calloc((size_t)n + 1, sizeof(int));
  int* Frow = (int*)calloc((size_t)n + 1, sizeof(int));
  int* Hcol = (int*)calloc((size_t)m + 1, sizeof(int));
  int* Ecol = (int*)calloc((size_t)m + 1, sizeof(int));
  int* corner = (int*)calloc((size_t)nbi, sizeof(int));
  if (!Hrow || !Frow || !Hcol || !Ecol || !corner) { free(Hrow); free(Frow); free(Hcol); free(Ecol); free(corner); return -1; }
  int best = 0;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  for (int d = 0; d < nbi + nbj - 1; ++d) {
    int lo = d - (nbj - 1); if (lo < 0) lo = 0;
    int hi = d; if (hi > nbi - 1) hi = nbi - 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(max : best)
    for (int bi = lo; bi <= hi; ++bi) {
      const int bj = d - bi;
      const int i0 = bi * TB, i1 = (i0 + TB < m) ? i0 + TB : m;
      const int j0 = bj * TB, j1 = (j0 + TB < n) ? j0 + TB : n;
      int b = tile_run(seq1, seq2, j0, j1, i0, i1, Hrow, Frow, Hcol, Ecol, &corner[bi], p);
      if (b > best) best = b;
    }
  }
  free(Hrow); free(Frow); free(Hcol); free(Ecol); free(corner);
  return best;
}

void oracle_gotoh_batch(const unsigned char* seq1_all, const int64_t* off1, const int* len1,
                        const unsigned char* seq2_all, const int64_t* off2, const int* len2,
                        int64_t npairs, const oracle_params* p, int threads, int* scores_out) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
  for (int64_t k = 0; k < npairs; ++k)
    scores_out[k] = oracle_gotoh_rolling(seq1_all + off1[k], seq2_all + off2[k], len1[k], len2[k], p);
}

void oracle_gotoh_banded_batch(const unsigned char* seq1_all, const int64_t* off1, const int* len1,
                               const unsigned char* seq2_all, const int64_t* off2, const int* len2,
                               int64_t npairs, int band_lo, int band_hi, const oracle_params* p,
                               int threads, int* scores_out) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
  for (int64_t k = 0; k < npairs; ++k)
    scores_out[k] = oracle_gotoh_banded(seq1_all + off1[k], seq2_all + off2[k], len1[k], len2[k],
                                        band_lo, band_hi, p, NULL);
}

/* ---- portable counter-based generator ----------------------------------------------------- */
uint64_t oracle_mix64(uint64_t seed, uint64_t stream, uint64_t index) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index + 1) + 0xD1B54A32D192ED03ull * stream;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void oracle_random_acgt(uint64_t seed, uint64_t stream, int64_t length, unsigned char* out) {
  static const char nt[4] = {'A', 'C', 'G', 'T'};
  for (int64_t k = 0; k < length; ++k) {
    /* one 64-bit word yields 32 symbols: word index k/32, 2 bits at position 2*(k%32) */
    uint64_t w = oracle_mix64(seed, stream, (uint64_t)(k >> 5));
    out[k] = (unsigned char)nt[(w >> (2 * (k & 31))) & 3];
  }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
