/* classify.c — intron typing by position-weight matrices and the DUST-like exon complexity score.
 *
 * classify_intron restates the part of reference src/classify-intron.c that decides the TYPE of an intron
 * (classify_genomic_intron_start_end :95-229, reached from src/factorization-refinement.c:612-629): only the
 * branch-point scan and the two 5' scores enter that decision, so only those six matrices are loaded
 * (pwm_data.h).  Scoring is MatInspector-like and in double, in the reference's operation order
 * (GetMatInspectorScoreOfaMotif :620-663, GetCVectorForPWM :1498, GetMAXVectorForPWM :1520); it stays on the host
 * (SURVEY.md §8(a) row 22).  dust_score restates src/exon-complexity.c:50-131.
 */
#include "ef.h"
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include "pwm_data.h"

enum { M_BPS9 = 0, M_BPS10, M_5GTAG_U12, M_5ATAC_U12, M_5GTAG_U2, M_5GCAG_U2, M_COUNT };

typedef struct pwm { int len; double w[4][14], cv[14], mx[14]; } pwm;
static pwm PW[M_COUNT];
static pthread_once_t pw_once = PTHREAD_ONCE_INIT;

static void pw_init(void) {
  for (int k = 0; k < M_COUNT; ++k) {
    pwm *P = &PW[k];
    P->len = PWM_DEFS[k].len;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < P->len; ++j) P->w[i][j] = PWM_DEFS[k].w[i][j] + 0.00001f;      /* float literal, as published */
    for (int j = 0; j < P->len; ++j) {
      double cvj = 0;
      for (int i = 0; i < 4; ++i) cvj += P->w[i][j] * log(P->w[i][j]);
      cvj += log(5.0f);
      cvj *= (100.0f / log(5.0f));
      P->cv[j] = cvj;
      double mxj = 0.0f;
      for (int i = 0; i < 4; ++i) if (P->w[i][j] > mxj) mxj = P->w[i][j];
      P->mx[j] = mxj;
    }
  }
}

/* s: `len` readable bytes (shorter pieces are padded by the caller with NUL, which scores like the reference's
 * out-of-alphabet case is not reproducible: we give it row A) */
static double motif_score(const char *s, int slen, const pwm *P) {
  double den = 0.0f, num = 0.0f;
  for (int i = 0; i < P->len; ++i) {
    const char ch = i < slen ? s[i] : 0;
    int idx = 0;
    if (ch == 'C' || ch == 'c') idx = 1;
    else if (ch == 'G' || ch == 'g') idx = 2;
    else if (ch == 'T' || ch == 't') idx = 3;
    num += P->cv[i] * P->w[idx][i];
    den += P->cv[i] * P->mx[i];
  }
  return num / den;
}

static double score_at(const char *g, int glen, int index, const pwm *P) {       /* real_substring(index, len) then score */
  int length = P->len;
  if (index < 0) { length += index; index = 0; }
  if (length < 0) length = 0;
  const int avail = index < glen ? MIN2(length, glen - index) : 0;
  return motif_score(g + MIN2(index, glen), avail, P);
}

/* SearchBPSinIntronSequenceWithMathInspector :575-618: best 12-mer (last one wins ties) with its start inside
 * [len-range_end, len-range_start] of the intron */
static int search_bps(const char *intron, int ilen, const pwm *P, double *score, int range_start, int range_end) {
  *score = 0.0f;
  if (ilen < range_start) return -1;
  int start_w = ilen - range_end;
  const int end_w = ilen - range_start;
  if (start_w < 0) start_w = 0;
  int start_bps = -1;
  bool first = true;
  for (int i = start_w; i <= end_w; ++i) {
    const double s = score_at(intron, ilen, i, P);
    if (first || s >= *score) { *score = s; start_bps = i; first = false; }
  }
  return start_bps;
}

static int good_bps(const char *intron, int ilen, int range_start, int range_end) {   /* ExistsGoodBPS... :535-573 */
  if (range_end > ilen) return -1;
  double s9 = 0.0f, s10 = 0.0f;
  const int b9 = search_bps(intron, ilen, &PW[M_BPS9], &s9, range_start, range_end);
  const int b10 = search_bps(intron, ilen, &PW[M_BPS10], &s10, range_start, range_end);
  if (s9 > s10) return s9 > 0.75f ? b9 : -1;
  return s10 > 0.75f ? b10 : -1;
}

static bool two(const char *p, int len, const char *lo, const char *upc) {
  return len == 2 && ((p[0] == lo[0] && p[1] == lo[1]) || (p[0] == upc[0] && p[1] == upc[1]));
}

/* The type of an intron depends on (genome, start, end) only, and the ESTs of a locus keep asking about the same
 * few introns: a fixed-size shared memo (one 64-bit word per entry: start | end | type; racing writers store the
 * same value) takes the PWM arithmetic off the per-EST path.  One genome per process (main.c). */
#define MEMO_BITS 18
static _Atomic uint64_t memo[1u << MEMO_BITS];
static char classify_intron_compute(const char *gen, int glen, int start, int end);

char classify_intron(const char *gen, int glen, int start, int end) {
  if (start < 0 || end < 0 || start >= (1 << 30) || end >= (1 << 30)) return classify_intron_compute(gen, glen, start, end);
  const uint64_t key = ((uint64_t)(uint32_t)start << 33) | ((uint64_t)(uint32_t)end << 2) | 0x8000000000000000ull;
  uint64_t h = key * 0x9E3779B97F4A7C15ull;
  for (int probe = 0; probe < 8; ++probe) {
    const uint32_t slot = (uint32_t)((h >> (64 - MEMO_BITS)) + (uint32_t)probe) & ((1u << MEMO_BITS) - 1u);
    const uint64_t v = atomic_load_explicit(&memo[slot], memory_order_relaxed);
    if ((v & ~3ull) == key) return (char)(v & 3u);
    if (v == 0) {
      const char t = classify_intron_compute(gen, glen, start, end);
      atomic_store_explicit(&memo[slot], key | (uint64_t)(t & 3), memory_order_relaxed);
      return t;
    }
  }
  return classify_intron_compute(gen, glen, start, end);
}

/* 0 = U12, 1 = U2, 2 = not determined (include/classify-intron.h:55-57) */
static char classify_intron_compute(const char *gen, int glen, int start, int end) {
  pthread_once(&pw_once, pw_init);
  /* the intron as real_substring(start, end-start+1) sees it */
  int s = start, length = end - start + 1;
  if (s < 0) { length += s; s = 0; }
  if (length < 0) length = 0;
  const int ilen = s < glen ? MIN2(length, glen - s) : 0;
  const char *intron = gen + MIN2(s, glen);
  const int bps = good_bps(intron, ilen, 14, 30);
  const char *p5 = intron, *p3 = intron + (ilen >= 2 ? ilen - 2 : 0);
  const int l5 = MIN2(2, ilen), l3 = ilen >= 2 ? 2 : ilen;
  const bool ag3 = two(p3, l3, "ag", "AG");
  double u12_5, u2_5;
  bool canonical = false;
  const double gtag_u12 = score_at(gen, glen, start - 3, &PW[M_5GTAG_U12]);
  if (two(p5, l5, "gt", "GT") && ag3) {
    canonical = true;
    u12_5 = gtag_u12;
    u2_5 = score_at(gen, glen, start - 3, &PW[M_5GTAG_U2]);
  } else if (two(p5, l5, "gc", "GC") && ag3) {
    canonical = true;
    u2_5 = score_at(gen, glen, start - 3, &PW[M_5GCAG_U2]);
    u12_5 = gtag_u12;
    const double alt = score_at(gen, glen, start - 3, &PW[M_5ATAC_U12]);
    if (alt > u12_5) u12_5 = alt;
  } else if (two(p5, l5, "at", "AT") && two(p3, l3, "ac", "AC")) {
    u12_5 = score_at(gen, glen, start - 3, &PW[M_5ATAC_U12]);
    u2_5 = score_at(gen, glen, start - 3, &PW[M_5GTAG_U2]);
    const double alt = score_at(gen, glen, start - 3, &PW[M_5GCAG_U2]);
    if (alt > u2_5) u2_5 = alt;
  } else {
    u12_5 = gtag_u12;
    double alt = score_at(gen, glen, start - 3, &PW[M_5ATAC_U12]);
    if (alt > u12_5) u12_5 = alt;
    u2_5 = score_at(gen, glen, start - 3, &PW[M_5GTAG_U2]);
    alt = score_at(gen, glen, start - 3, &PW[M_5GCAG_U2]);
    if (alt > u2_5) u2_5 = alt;
  }
  if (bps != -1) return u12_5 > u2_5 ? 0 : 1;
  if (canonical) return 1;
  if (u12_5 - u2_5 > 0.25 && u12_5 >= 0.75) return 0;
  return 2;
}

/* ---- DUST-like score (exon-complexity.c:50-131) ---------------------------------------------------------------- */
static int nt2(char c) {
  switch (c) { case 'a': case 'A': return 0; case 'c': case 'C': return 1; case 'g': case 'G': return 2; case 't': case 'T': return 3; }
  return -1;
}

double dust_score(const char *s, int len) {
  if (len <= 2) return 0.0;
  int freq[17] = {0}, running = 0;
  for (int i = 0; i < len - 1; ++i) {
    const int a = nt2(s[i]), b = nt2(s[i + 1]);
    const int idx = (a < 0 || b < 0) ? 16 : a * 4 + b;
    running += freq[idx];
    ++freq[idx];
  }
  const double dust = (10.0 * (double)running) / ((double)((size_t)len - 2));
  return dust / (double)(size_t)len;
}
