// TEST INFRASTRUCTURE.  pintron_b200/csrc/align_core.h (compute_alignment as the CUDA kernel k_align_bp runs it: one job per
// thread, bit-parallel, traceback from stored delta vectors) compiled for the host and fuzzed against the oracle port's
// full-matrix compute_alignment (oracle/port/dp_port.c po_align, itself pinned to the reference): score AND every alignment
// column must agree, N wildcards and case included.  Prints "ok N".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../pintron_b200/csrc/align_core.h"
extern "C" int po_align(const char *est, int n, const char *gen, int m, uint8_t *ops, int *ops_len);

static unsigned long long rs = 0x2545F4914F6CDD1Dull;
static unsigned rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (unsigned)(rs >> 11); }

int main(int argc, char **argv) {
  const int cases = argc > 1 ? atoi(argv[1]) : 20000;
  static const char *alpha[4] = {"ACGT", "ACGTacgtNn", "AC", "ACGTN"};
  std::vector<unsigned long long> peq(MY_NSYM * MY_MAXW * 3);
  int unsupported = 0, nospace = 0;
  for (int c = 0; c < cases; ++c) {
    const char *al = alpha[c % 4];
    const int na = (int)strlen(al);
    int n = (c % 5 == 0) ? (int)(rnd() % 321) : (int)(rnd() % 90);          // EST rows: up to 5 blocks
    int m = (c % 7 == 0) ? (int)(rnd() % 30) : n + (int)(rnd() % 40) - (n > 20 ? 10 : 0);
    if (m < 0) m = 0;
    if (c % 211 == 0) n = 0;
    if (c % 223 == 0) m = 0;
    std::vector<char> p(n + 8), t(m + 8);      // the core reads whole 4-byte words
    for (int i = 0; i < n; ++i) p[i] = al[rnd() % na];
    if (c & 1) { for (int j = 0; j < m; ++j) t[j] = al[rnd() % na]; }
    else {       // genome piece = mutated copy of the EST (what est-fact aligns)
      int j = 0;
      for (int i = 0; i < n && j < m; ++i) { unsigned r = rnd() % 100; if (r < 4) continue; if (r < 8 && j + 1 < m) t[j++] = al[rnd() % na]; t[j++] = (r < 12) ? al[rnd() % na] : p[i]; }
      while (j < m) t[j++] = al[rnd() % na];
    }
    if (c % 97 == 0 && m > 0) t[rnd() % m] = '*';
    if (c % 89 == 0 && n > 0) p[rnd() % n] = '#';
    const int stride = 1 + (c % 3);
    const long long tbs = 1 + (c % 4);
    const int W = (n + 63) >> 6;
    long long cap = (long long)m * W + (c % 13 == 0 ? -1 : 3);
    if (cap < 0) cap = 0;
    std::vector<unsigned long long> tb((size_t)(2 * (cap + 1) * tbs + 8));
    std::vector<uint8_t> ops(n + m + 8), want_ops(n + m + 8);
    int k = -1;
    const uint32_t got = (n <= 64 && (c & 2)) ? my_align<1>((const uint8_t *)p.data(), n, (const uint8_t *)t.data(), m, peq.data(), stride, tb.data(), tbs, cap, ops.data(), &k)
                                              : my_align<MY_MAXW>((const uint8_t *)p.data(), n, (const uint8_t *)t.data(), m, peq.data(), stride, tb.data(), tbs, cap, ops.data(), &k);
    if ((long long)m * W > cap) { if (got != MY_NOSPACE) { fprintf(stderr, "case %d: missing space not reported\n", c); return 1; } ++nospace; continue; }
    bool has_other = false;
    for (int i = 0; i < n; ++i) has_other |= my_sym_switch((uint8_t)p[i]) < 0;
    for (int j = 0; j < m && n > 0; ++j) has_other |= my_sym_switch((uint8_t)t[j]) < 0;
    if (has_other) { if (got != MY_UNSUPPORTED) { fprintf(stderr, "case %d: unsupported byte not reported\n", c); return 1; } ++unsupported; continue; }
    int wk = 0;
    const int want = po_align(p.data(), n, t.data(), m, want_ops.data(), &wk);
    if ((int)got != want || k != wk || memcmp(ops.data(), want_ops.data(), (size_t)wk) != 0) {
      fprintf(stderr, "MISMATCH case %d: n %d m %d got %u want %d, ops %d vs %d\n", c, n, m, got, want, k, wk);
      return 1;
    }
  }
  printf("ok %d (unsupported %d, no space %d)\n", cases, unsupported, nospace);
  return 0;
}
