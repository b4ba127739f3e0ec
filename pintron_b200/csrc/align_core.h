// align_core.h — compute_alignment (reference src/compute-alignments.c:39-207: ComputeAlignMatrix + TracebackAlignment)
// as ONE JOB PER THREAD, bit-parallel, with the traceback rebuilt from stored delta vectors.
//
// The wavefront form (k_dp.cu: one warp per job, one diagonal per step through shared memory) spends a __syncwarp and
// three shared-memory round trips per diagonal on matrices that are mostly a few hundred cells (est-fact aligns exon
// pieces of 10-60 nt by the ten thousand): INT-ALU fraction 0.006 in round 2.  Here the EST (rows) lies along the bits of
// 64-bit words and a genome column costs ~20 integer instructions per 64 rows (Myers 1999 / Hyyro 2003, the chaining of
// myers_core.h), with the reference's N wildcard folded into the match vectors: an EST 'N' matches every column (its bit
// is set in every Peq word), a genome 'N' column matches every row (its Peq word is all ones).
//
// Traceback.  The reference stores one direction per cell with the tie-break "diagonal, then up only if strictly
// cheaper, then left only if strictly cheaper" (compute-alignments.c:114-136), i.e. for cell (i, j):
//   diagonal  iff  D[i-1][j-1] + (match ? 0 : 1) == D[i][j]   iff  match or D[i][j] != D[i-1][j-1]
//   else up   iff  D[i-1][j] + 1 == D[i][j]                    iff  the vertical delta of column j at row i is +1
//   else left.
// Both bits are by-products of the column step: D0 = Xh | Mv says D[i][j] == D[i-1][j-1] (Hyyro's diagonal-zero vector;
// at the first row of a chained block the "hin < 0" adjustment makes exactly this bit true), and the new Pv is the vertical
// +1 vector.  So two words per (column, 64-row block) — DG = Eq | ~D0 and Pv — replace the reference's byte per cell: 16
// bytes instead of 64, and no matrix of scores at all.  Border cells carry the pure-gap direction, as in the reference.
//
// Storage is addressed through a stride so that the threads of a CUDA grid interleave their entries (entry q of a thread
// = two words at tb[2 * q * tb_stride]); the host tests use stride 1 and other strides.  Plain C++: tests compile this
// header with g++ (tests/cpu_backend/align_fuzz.cpp) against the oracle port.
#pragma once
#include "myers_core.h"

#define MY_NOSPACE 0xfffffffeu

// pat = EST (rows, n <= 64 * MAXW), txt = genome piece (columns).  ops gets one byte per alignment column, left to right:
// 0 = EST char over genome char, 1 = EST char over '-', 2 = '-' over genome char (include/pintron_cuda.h); room for n + m.
// Returns the score D[n][m]; MY_UNSUPPORTED for bytes outside ACGTacgtNn; MY_NOSPACE when m * ceil(n / 64) > tb_cap entries.
template <int MAXW>
MY_HD uint32_t my_align(const uint8_t *pat, int n, const uint8_t *txt, int m, unsigned long long *peq, int stride,
                        unsigned long long *tb, long long tb_stride, long long tb_cap, uint8_t *ops, int *n_ops) {
  const int W = (n + 63) >> 6;
  if ((long long)m * (long long)W > tb_cap) return MY_NOSPACE;
  for (int s = 0; s < MY_NSYM; ++s)
    for (int w = 0; w < W; ++w) peq[(s * MAXW + w) * stride] = 0ull;
  for (int i0 = 0; i0 < n; i0 += 4) {
    const uint32_t w4 = MY_LOAD4(pat + i0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + q;
      if (i < n) {
        const int s = MY_SYM((uint8_t)(w4 >> (8 * q)));
        if (s < 0) return MY_UNSUPPORTED;
        peq[(s * MAXW + (i >> 6)) * stride] |= 1ull << (i & 63);
      }
    }
  }
  for (int w = 0; w < W; ++w) {                            // the wildcard: EST N/n rows match every symbol, genome N/n columns match every row
    const unsigned long long nrows = peq[(8 * MAXW + w) * stride] | peq[(9 * MAXW + w) * stride];
    for (int s = 0; s < 8; ++s) peq[(s * MAXW + w) * stride] |= nrows;
    peq[(8 * MAXW + w) * stride] = ~0ull;
    peq[(9 * MAXW + w) * stride] = ~0ull;
  }
  unsigned long long Pv[MAXW], Mv[MAXW];
  for (int w = 0; w < MAXW; ++w) { Pv[w] = ~0ull; Mv[w] = 0ull; }
  const unsigned long long top = n > 0 ? 1ull << ((n - 1) & 63) : 0ull;
  uint32_t score = (uint32_t)n;
  long long q_at = 0;
  for (int j0 = 0; j0 < m && n > 0; j0 += 4) {
    const uint32_t w4 = MY_LOAD4(txt + j0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (j0 + q < m) {
        const int s = MY_SYM((uint8_t)(w4 >> (8 * q)));
        if (s < 0) return MY_UNSUPPORTED;
        int hin = 1;                                      // D[0][j] - D[0][j-1] = +1
#pragma unroll
        for (int w = 0; w < MAXW; ++w) {
          if (w < W) {
            const unsigned long long eq_true = peq[(s * MAXW + w) * stride];
            unsigned long long Eq = eq_true;
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
            const unsigned long long pvn = Mh | ~(Xv | Ph);
            Pv[w] = pvn;
            Mv[w] = Ph & Xv;
            hin = hout;
            tb[2 * q_at * tb_stride] = eq_true | ~(Xh | mv);          // DG: the reference takes the diagonal here
            tb[(2 * q_at + 1) * tb_stride] = pvn;                      // else "up" where the vertical delta is +1
            ++q_at;
          }
        }
        score += (uint32_t)hin;
      }
    }
  }
  if (n == 0) score = (uint32_t)m;
  // traceback, last column first (TracebackAlignment, compute-alignments.c:149-207), then reversed in place
  int i = n, j = m, k = 0;
  while (i > 0 && j > 0) {
    const int w = (i - 1) >> 6, b = (i - 1) & 63;
    const long long e = (long long)(j - 1) * W + w;
    uint8_t d;
    if ((tb[2 * e * tb_stride] >> b) & 1ull) { d = 0; --i; --j; }
    else if ((tb[(2 * e + 1) * tb_stride] >> b) & 1ull) { d = 1; --i; }
    else { d = 2; --j; }
    ops[k++] = d;
  }
  while (i > 0) { ops[k++] = 1; --i; }
  while (j > 0) { ops[k++] = 2; --j; }
  for (int a = 0, z = k - 1; a < z; ++a, --z) { const uint8_t t = ops[a]; ops[a] = ops[z]; ops[z] = t; }
  *n_ops = k;
  return score;
}
