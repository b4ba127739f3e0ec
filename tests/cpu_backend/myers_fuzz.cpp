// TEST INFRASTRUCTURE.  pintron_b200/csrc/myers_core.h (the bit-parallel edit distance the CUDA kernel k_myers runs per
// thread) compiled for the host and fuzzed against the oracle port's full-matrix edit distance (oracle/port/dp_port.c
// po_edit, itself pinned to the reference).  Prints "ok N".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../pintron_b200/csrc/myers_core.h"
extern "C" unsigned po_edit(const char *s1, int l1, const char *s2, int l2);

static unsigned long long rs = 0x9E3779B97F4A7C15ull;
static unsigned rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (unsigned)(rs >> 11); }

int main(int argc, char **argv) {
  const int cases = argc > 1 ? atoi(argv[1]) : 20000;
  static const char *alpha[3] = {"ACGT", "ACGTacgtNn", "AC"};
  std::vector<unsigned long long> peq(MY_NSYM * MY_MAXW * 3);
  int unsupported = 0;
  for (int c = 0; c < cases; ++c) {
    const char *al = alpha[c % 3];
    const int na = (int)strlen(al);
    int m = (c % 5 == 0) ? (int)(rnd() % 321) : (int)(rnd() % 90);          // pattern: up to 5 blocks
    int n = m + (int)(rnd() % 40);
    std::vector<char> p(m + 8), t(n + 8);      // the core reads whole 4-byte words
    for (int i = 0; i < m; ++i) p[i] = al[rnd() % na];
    // text = mutated copy of the pattern (realistic: small distances) or random
    if (c & 1) { for (int j = 0; j < n; ++j) t[j] = al[rnd() % na]; }
    else {
      int j = 0;
      for (int i = 0; i < m && j < n; ++i) { unsigned r = rnd() % 100; if (r < 4) continue; if (r < 8 && j + 1 < n) t[j++] = al[rnd() % na]; t[j++] = (r < 12) ? al[rnd() % na] : p[i]; }
      while (j < n) t[j++] = al[rnd() % na];
    }
    if (c % 97 == 0 && n > 0) t[rnd() % n] = '*';                             // a byte outside the 10 symbols
    const int stride = 1 + (c % 3);                                            // the strided Peq layout of the kernel
    const uint32_t got = (m <= 64 && (c & 2)) ? my_edit_distance<1>((const uint8_t *)p.data(), m, (const uint8_t *)t.data(), n, peq.data(), stride)
                                              : my_edit_distance<MY_MAXW>((const uint8_t *)p.data(), m, (const uint8_t *)t.data(), n, peq.data(), stride);
    bool has_other = false;
    for (int i = 0; i < m; ++i) has_other |= my_sym_switch((uint8_t)p[i]) < 0;
    for (int j = 0; j < n; ++j) has_other |= my_sym_switch((uint8_t)t[j]) < 0;
    if (has_other && m > 0) { if (got != MY_UNSUPPORTED) { fprintf(stderr, "case %d: unsupported byte not reported\n", c); return 1; } ++unsupported; continue; }
    const unsigned want = po_edit(p.data(), m, t.data(), n);
    if (got != want) { fprintf(stderr, "MISMATCH case %d: m %d n %d got %u want %u\n", c, m, n, got, want); return 1; }
  }
  printf("ok %d (unsupported %d)\n", cases, unsupported);
  return 0;
}
