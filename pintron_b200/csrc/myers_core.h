// myers_core.h — bit-parallel GLOBAL unit-cost edit distance (Myers 1999 in Hyyro's block form), one job per thread.
//
// Used by k_myers.cu for the two scalar-result DPs of est-fact: edit_distance (reference src/refine.c:51,
// src/compute-alignments.c:235) and K_band_edit_distance (src/compute-alignments.c:319-453).  Both compare literal
// bytes (no N wildcard), so the value is the plain Levenshtein distance D[m][n] with D[i][0] = i, D[0][j] = j.
//
// The pattern (the SHORTER string, m <= 64 * MY_MAXW) is laid along the bits; column j of the DP is held as the
// vertical +1 / -1 delta vectors (Pv, Mv).  Blocks of 64 rows are chained through the horizontal delta that leaves a
// block (hout) and enters the next one; the top of the matrix feeds +1 (row 0 grows by one per column: global
// alignment).  Bits above row m in the last block never influence lower bits (carries and shifts only move up), so
// the score is followed at bit (m - 1) of the last block without any padding.
//
// Alphabet: the match vectors Peq are kept for 10 symbols (ACGT, acgt, N, n); a string holding any other byte is
// reported as MY_UNSUPPORTED and goes to the generic wavefront kernel.  Plain C++: tests compile this header with g++.
#pragma once
#include <stdint.h>

#ifndef MY_HD
#ifdef __CUDACC__
#define MY_HD __host__ __device__ __forceinline__
#else
#define MY_HD inline
#endif
#endif

#define MY_MAXW 5            /* widest variant: up to 320 pattern rows */
#define MY_NSYM 10
#define MY_UNSUPPORTED 0xffffffffu

MY_HD int my_sym_switch(uint8_t c) {
  switch (c) {
    case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
    case 'a': return 4; case 'c': return 5; case 'g': return 6; case 't': return 7;
    case 'N': return 8; case 'n': return 9;
    default: return -1;
  }
}
// The kernel replaces these two: MY_SYM by a 256-entry table in shared memory (a switch diverges inside a warp),
// MY_LOAD4 by aligned word loads + funnel shift (one load per four letters instead of four byte loads).
#ifndef MY_SYM
#define MY_SYM(c) my_sym_switch(c)
#endif
#ifndef MY_LOAD4
#include <string.h>
MY_HD uint32_t my_load4_host(const uint8_t *p) { uint32_t w; memcpy(&w, p, 4); return w; }     /* may read 3 bytes past the string */
#define MY_LOAD4(p) my_load4_host(p)
#endif

// Peq storage is addressed through a stride so that a CUDA block can interleave its threads in shared memory
// (word (sym, w) of this thread lives at peq[(sym * MAXW + w) * stride]).  MAXW = blocks of 64 rows this instance
// can hold (m <= 64 * MAXW is the caller's business).  Both strings must be readable up to 3 bytes past their end.
template <int MAXW>
MY_HD uint32_t my_edit_distance(const uint8_t *pat, int m, const uint8_t *txt, int n, unsigned long long *peq, int stride) {
  if (m == 0) return (uint32_t)n;
  const int W = (m + 63) >> 6;
  for (int s = 0; s < MY_NSYM; ++s)
    for (int w = 0; w < W; ++w) peq[(s * MAXW + w) * stride] = 0ull;
  for (int i0 = 0; i0 < m; i0 += 4) {
    const uint32_t w4 = MY_LOAD4(pat + i0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + q;
      if (i < m) {
        const int s = MY_SYM((uint8_t)(w4 >> (8 * q)));
        if (s < 0) return MY_UNSUPPORTED;
        peq[(s * MAXW + (i >> 6)) * stride] |= 1ull << (i & 63);
      }
    }
  }
  unsigned long long Pv[MAXW], Mv[MAXW];
  for (int w = 0; w < MAXW; ++w) { Pv[w] = ~0ull; Mv[w] = 0ull; }
  const unsigned long long top = 1ull << ((m - 1) & 63);
  uint32_t score = (uint32_t)m;
  for (int j0 = 0; j0 < n; j0 += 4) {
    const uint32_t w4 = MY_LOAD4(txt + j0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (j0 + q < n) {
        const int s = MY_SYM((uint8_t)(w4 >> (8 * q)));
        if (s < 0) return MY_UNSUPPORTED;
        int hin = 1;                                      // D[0][j] - D[0][j-1] = +1
#pragma unroll
        for (int w = 0; w < MAXW; ++w) {
          if (w < W) {
            unsigned long long Eq = peq[(s * MAXW + w) * stride];
            const unsigned long long pv = Pv[w], mv = Mv[w];
            const unsigned long long Xv = Eq | mv;
            if (hin < 0) Eq |= 1ull;
            const unsigned long long Xh = (((Eq & pv) + pv) ^ pv) | Eq;
            unsigned long long Ph = mv | ~(Xh | pv);
            unsigned long long Mh = pv & Xh;
            const unsigned long long hb = (w == W - 1) ? top : (1ull << 63);
            const int hout = (Ph & hb) ? 1 : ((Mh & hb) ? -1 : 0);
            Ph <<= 1; Mh <<= 1;
            if (hin < 0) Mh |= 1ull; else if (hin > 0) Ph |= 1ull;
            Pv[w] = Mh | ~(Xv | Ph);
            Mv[w] = Ph & Xv;
            hin = hout;
          }
        }
        score += (uint32_t)hin;                           // hout of the last block: the step of row m in this column
      }
    }
  }
  return score;
}
