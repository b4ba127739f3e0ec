/* TEST INFRASTRUCTURE.  ef_small_exon_scan (pintron_b200/host/refine_fact.c: one memmem pass per offstart) against a
 * literal restatement of the reference's nested strstr loops (src/factorization-refinement.c:770-834: NUL-patched
 * copies of the EST middle and of the intron, strstr per (offstart, offend), "first strictly longer" update), on
 * random genomes with planted copies of the EST middle (some with N / lower-case bytes), real classify_intron
 * included; both search paths (6-mer index, memmem).  Prints "ok N hits H". */
#include "ef.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static size_t min3(size_t a, size_t b, size_t c) { size_t t = a; if (t > b) t = b; if (t > c) t = c; return t; }

static void scan_literal(const char *g, int glen_all, const char *e, size_t estart, size_t elen, size_t allgstart, size_t allglen,
                         size_t f1slen, size_t f2plen, size_t MINI, size_t out[7]) {
  const size_t LB = 6, MPB = 6;
  char *efact = calloc(elen + 1, 1), *allgfact = calloc(allglen + 1, 1);
  memcpy(efact, e + estart, elen);
  memcpy(allgfact, g + allgstart, allglen);
  size_t best = 0, ecut1 = 0, ecut2 = 0, a1 = 0, a2 = 0, b1 = 0, b2 = 0;
  const size_t max_offstart = min3(f1slen + 1 - MPB, elen + 1 - LB, allglen + 1 - (2 * MINI) - LB);
  for (size_t offstart = 0; offstart < max_offstart; ++offstart) {
    const size_t max_offend = min3(f2plen + 1 - MPB, elen + 1 - offstart - LB, allglen + 1 - (2 * MINI) - LB - offstart);
    for (size_t offend = 0; offend < max_offend; ++offend) {
      const char ce = efact[elen - offend];
      efact[elen - offend] = 0;
      const char cg = allgfact[allglen - offend - MINI];
      allgfact[allglen - offend - MINI] = 0;
      char *occ = allgfact + offstart + MINI;
      while ((occ = strstr(occ, efact + offstart))) {
        const size_t i1start = allgstart + offstart, i1end = allgstart + (size_t)(occ - allgfact) - 1;
        const size_t i2start = i1end + 1 + elen - offstart - offend, i2end = allgstart + allglen - offend - 1;
        const char t1 = classify_intron(g, glen_all, (int)i1start, (int)i1end), t2 = classify_intron(g, glen_all, (int)i2start, (int)i2end);
        if (t1 != 2 && t2 != 2) {
          const size_t sl = elen - offstart - offend;
          if (sl > best) { best = sl; ecut1 = estart + offstart; ecut2 = ecut1 + sl; a1 = i1start; a2 = i1end + 1; b1 = i2start; b2 = i2end + 1; }
        }
        ++occ;
      }
      efact[elen - offend] = ce;
      allgfact[allglen - offend - MINI] = cg;
    }
  }
  out[0] = best; out[1] = ecut1; out[2] = ecut2; out[3] = a1; out[4] = a2; out[5] = b1; out[6] = b2;
  free(efact); free(allgfact);
}

static unsigned long long rs = 88172645463325252ull;
static unsigned rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (unsigned)(rs >> 11); }

int main(int argc, char **argv) {
  const int cases = argc > 1 ? atoi(argv[1]) : 3000;
  int hits = 0;
  for (int c = 0; c < cases; ++c) {
    const int alpha = (c % 3 == 0) ? 2 : 4;                      /* a two-letter genome makes occurrences dense */
    const size_t glen = 600 + rnd() % 3000, elenall = 120;
    char *g = malloc(glen + 64), *e = malloc(elenall + 64);
    for (size_t i = 0; i < glen + 64; ++i) g[i] = "ACGT"[rnd() % alpha];
    for (size_t i = 0; i < elenall + 64; ++i) e[i] = "ACGT"[rnd() % alpha];
    const size_t MINI = (c % 5 == 0) ? 4 : 40;
    const size_t elen = 6 + rnd() % 34, estart = rnd() % 40;
    const size_t allgstart = 50 + rnd() % 100;
    size_t allglen = 2 * MINI + 6 + rnd() % (glen - allgstart - 2 * MINI - 60);
    const size_t f1slen = 6 + rnd() % 18, f2plen = 6 + rnd() % 18;
    /* plant canonical splice sites and copies of (parts of) the EST middle so that some trims occur and classify */
    for (int k = 0; k < 6; ++k) {
      const size_t os = rnd() % 4, oe = rnd() % 4;
      if (elen < 6 + os + oe) continue;
      const size_t sl = elen - os - oe;
      if (allglen < 2 * MINI + sl + 8) continue;
      const size_t q = allgstart + MINI + os + rnd() % (allglen - 2 * MINI - sl - os + 1);
      memcpy(g + q, e + estart + os, sl);
      if (k & 1) { g[q - 2] = 'A'; g[q - 1] = 'G'; g[q + sl] = 'G'; g[q + sl + 1] = 'T'; }
    }
    if (c % 7 == 3) for (int k = 0; k < 12; ++k) { g[rnd() % glen] = "Nacgt"[rnd() % 5]; e[rnd() % elenall] = "Nacgt"[rnd() % 5]; }
    if (c & 1) { memcpy(g + allgstart, "GT", 2); memcpy(g + allgstart + 1, "GT", 2); memcpy(g + allgstart + allglen - 2, "AG", 2); }
    size_t r1[7], r2[7];
    ef_small_exon_index_build(c % 4 != 1 ? g : NULL, c % 4 != 1 ? glen : 0);        /* 3 of 4 cases through the 6-mer index, the rest through memmem */
    scan_literal(g, (int)glen, e, estart, elen, allgstart, allglen, f1slen, f2plen, MINI, r1);
    ef_small_exon_scan(g, (int)glen, e, estart, elen, allgstart, allglen, f1slen, f2plen, MINI, r2);
    if (memcmp(r1, r2, sizeof r1)) {
      fprintf(stderr, "MISMATCH case %d: literal %zu %zu %zu %zu %zu %zu %zu  new %zu %zu %zu %zu %zu %zu %zu\n", c, r1[0], r1[1], r1[2], r1[3],
              r1[4], r1[5], r1[6], r2[0], r2[1], r2[2], r2[3], r2[4], r2[5], r2[6]);
      return 1;
    }
    if (r1[0] >= 6) ++hits;
    free(g); free(e);
  }
  printf("ok %d hits %d\n", cases, hits);
  return 0;
}
