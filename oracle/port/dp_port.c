/* TEST INFRASTRUCTURE — the CPU oracle ("port") for the est-fact DP path.
 *
 * Plain-C restatement of the reference's integer DP routines, written from their behaviour (full
 * matrices, row-major, the reference's tie-breaks) and pinned against the compiled reference
 * (oracle/_ref/libref_dp.so, see tests/golden/make_dp_golden.py and tests/test_oracle_port.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (est-fact + libpintron_cuda.so) never does.
 *
 * Alignment "ops" are one byte per alignment column, in left-to-right order:
 *   0 = EST char over genome char, 1 = EST char over '-', 2 = '-' over genome char.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

static inline int is_n(char c) { return c == 'n' || c == 'N'; }
static inline int wild_eq(char a, char b) { return a == b || is_n(a) || is_n(b); }
#define MIN2(a, b) ((a) < (b) ? (a) : (b))

/* ---- DP A: global unit-cost alignment with N wildcards and traceback --------------------------
 * reference src/compute-alignments.c:39-207 (compute_alignment, ComputeAlignMatrix, TracebackAlignment).
 * Tie-break: diagonal, then "up" (EST char vs '-') only if strictly cheaper, then "left" ('-' vs genome
 * char) only if strictly cheaper than the current best (compute-alignments.c:114-136). */
int po_align(const char *est, int n, const char *gen, int m, uint8_t *ops, int *ops_len) {
  size_t W = (size_t)m + 1;
  uint32_t *prev = malloc(W * sizeof *prev), *cur = malloc(W * sizeof *cur);
  uint8_t *dir = calloc((size_t)(n + 1) * W, 1);
  for (int j = 0; j <= m; ++j) prev[j] = (uint32_t)j;
  for (int i = 1; i <= n; ++i) {
    cur[0] = (uint32_t)i;
    for (int j = 1; j <= m; ++j) {
      uint32_t best = prev[j - 1] + (wild_eq(est[i - 1], gen[j - 1]) ? 0u : 1u);
      uint8_t d = 0;
      if (best > prev[j] + 1) { best = prev[j] + 1; d = 1; }
      if (best > cur[j - 1] + 1) { best = cur[j - 1] + 1; d = 2; }
      cur[j] = best;
      dir[(size_t)i * W + j] = d;
    }
    uint32_t *t = prev; prev = cur; cur = t;
  }
  int score = (int)prev[m];
  /* traceback (compute-alignments.c:149-207): borders are pure gap runs */
  int i = n, j = m, k = 0;
  while (i > 0 && j > 0) {
    uint8_t d = dir[(size_t)i * W + j];
    ops[k++] = d;
    if (d == 0) { --i; --j; } else if (d == 1) --i; else --j;
  }
  while (i > 0) { ops[k++] = 1; --i; }
  while (j > 0) { ops[k++] = 2; --j; }
  for (int a = 0, b = k - 1; a < b; ++a, --b) { uint8_t t = ops[a]; ops[a] = ops[b]; ops[b] = t; }
  *ops_len = k;
  free(prev); free(cur); free(dir);
  return score;
}

/* ---- plain edit distance (no wildcard, case-sensitive), last cell ------------------------------
 * reference src/refine.c:51-83 (edit_distance) and src/compute-alignments.c:210-244
 * (edit_distance_matrix / compute_edit_distance) compute the same value. */
static uint32_t *ed_matrix(const char *rows, int nr, const char *cols, int nc) {
  size_t W = (size_t)nc + 1;
  uint32_t *M = malloc((size_t)(nr + 1) * W * sizeof *M);
  for (int j = 0; j <= nc; ++j) M[j] = (uint32_t)j;
  for (int i = 1; i <= nr; ++i) {
    M[i * W] = (uint32_t)i;
    for (int j = 1; j <= nc; ++j) {
      uint32_t v = M[(i - 1) * W + j - 1] + (rows[i - 1] != cols[j - 1]);
      v = MIN2(v, M[(i - 1) * W + j] + 1);
      v = MIN2(v, M[i * W + j - 1] + 1);
      M[i * W + j] = v;
    }
  }
  return M;
}

unsigned po_edit(const char *s1, int l1, const char *s2, int l2) {
  uint32_t *M = ed_matrix(s1, l1, s2, l2);
  unsigned r = M[(size_t)(l1 + 1) * (l2 + 1) - 1];
  free(M);
  return r;
}

/* ---- DP B: thresholded edit distance inside the band |c-r| <= k ---------------------------------
 * reference src/compute-alignments.c:319-453 (K_band_edit_distance): equal strings -> 0/true;
 * k==0 -> 1/false; longer string becomes the column string; n-m>k -> n-m/false; 2k+1>=n -> exact
 * distance; otherwise a band of diagonals [-k,k] with nothing outside it. */
int po_kband(const char *a, int la, const char *b, int lb, unsigned k, unsigned *edit) {
  if (la == lb && memcmp(a, b, (size_t)la) == 0) { *edit = 0; return 1; }
  if (k == 0) { *edit = 1; return 0; }
  const char *s1 = a, *s2 = b; int n = la, m = lb;
  if (n < m) { s1 = b; s2 = a; n = lb; m = la; }
  if ((unsigned)(n - m) > k) { *edit = (unsigned)(n - m); return 0; }
  if (2 * (size_t)k + 1 >= (size_t)n) { *edit = po_edit(s1, n, s2, m); return *edit <= k; }
  /* rows r over s2 (1..m), columns c over s1 (1..n); cell exists iff |c-r|<=k */
  const uint32_t INF = 0x3fffffffu;
  size_t W = 2 * (size_t)k + 1;
  uint32_t *prev = malloc((W + 2) * sizeof *prev), *cur = malloc((W + 2) * sizeof *cur);
  /* index d = c - r + k + 1 in [1, W]; slots 0 and W+1 are sentinels */
  for (size_t d = 0; d < W + 2; ++d) prev[d] = INF;
  for (unsigned c = 0; c <= k; ++c) prev[c + k + 1] = c;   /* row 0 */
  for (int r = 1; r <= m; ++r) {
    for (size_t d = 0; d < W + 2; ++d) cur[d] = INF;
    for (size_t d = 1; d <= W; ++d) {
      long c = (long)d - (long)k - 1 + r;
      if (c < 0 || c > n) continue;
      if (c == 0) { cur[d] = (uint32_t)r; continue; }
      uint32_t v = prev[d] + (s1[c - 1] != s2[r - 1]);   /* diagonal keeps d */
      v = MIN2(v, cur[d - 1] + 1);                        /* left: (r, c-1) */
      v = MIN2(v, prev[d + 1] + 1);                       /* up:   (r-1, c) */
      cur[d] = v;
    }
    uint32_t *t = prev; prev = cur; cur = t;
  }
  uint32_t res = prev[(size_t)(n - m) + k + 1];
  free(prev); free(cur);
  *edit = res;
  return res <= k;
}

/* ---- Burset splice-site dinucleotide frequencies -------------------------------------------------
 * reference src/refine-intron.c:376-556 (getBursetFrequency; upper-cases its arguments) and :362-374
 * (getBursetFrequency_adaptor). */
static const struct { char d[3], a[3]; int f; } BURSET[] = {
  {"AA","AG",1},{"AA","AT",1},{"AA","GT",1},{"AC","CC",1},{"AG","AC",1},{"AG","AG",5},{"AG","CT",2},
  {"AG","GC",1},{"AG","TG",2},{"AT","AA",1},{"AT","AC",8},{"AT","AG",7},{"AT","AT",2},{"AT","GC",1},
  {"AT","GT",1},{"CA","AG",1},{"CA","TT",1},{"CC","AG",2},{"CG","AG",1},{"CG","CA",1},{"CT","AC",2},
  {"CT","CA",1},{"GA","AG",8},{"GA","GT",1},{"GA","TC",1},{"GA","TG",1},{"GC","AG",126},{"GC","GG",1},
  {"GC","TA",1},{"GG","AC",1},{"GG","AG",11},{"GG","CA",1},{"GG","GA",2},{"GG","TC",2},{"GT","AG",200},
  {"GT","AC",4},{"GT","AT",2},{"GT","CA",9},{"GT","CG",4},{"GT","CT",3},{"GT","GC",1},{"GT","GG",10},
  {"GT","GT",1},{"GT","TA",7},{"GT","TC",2},{"GT","TG",8},{"GT","TT",2},{"TA","AG",6},{"TA","CG",1},
  {"TA","TC",1},{"TC","AG",1},{"TC","GG",1},{"TG","AC",1},{"TG","AG",7},{"TG","GG",2},{"TT","AG",5},
  {"TT","AT",1},{"TT","GG",1},
};

int po_burset(char d0, char d1, char a0, char a1) {
  /* the reference compares NUL-terminated 2-char strings: an embedded NUL shortens the string */
  char d[3] = { (char)toupper((unsigned char)d0), d0 ? (char)toupper((unsigned char)d1) : 0, 0 };
  char a[3] = { (char)toupper((unsigned char)a0), a0 ? (char)toupper((unsigned char)a1) : 0, 0 };
  for (size_t i = 0; i < sizeof BURSET / sizeof BURSET[0]; ++i)
    if (strcmp(d, BURSET[i].d) == 0 && strcmp(a, BURSET[i].a) == 0) return BURSET[i].f;
  return 0;
}

static int burset_adaptor(const char *t, size_t cut1, size_t cut2) {
  if (cut2 < 2) return 0;
  return po_burset(t[cut1], t[cut1 + 1], t[cut2 - 2], t[cut2 - 1]);
}

/* ---- DP C: splice-border placement ---------------------------------------------------------------
 * reference src/refine.c:106-190 (general_refine_borders): prefix DP of p against t[0..t_win) and of
 * reversed p against reversed t, per-row minimum with FIRST argmin, then the split of p minimising
 * the sum, ties broken by the larger Burset frequency (strict), earliest split otherwise.
 * NB: t is read at t[off_t1], t[off_t1+1] (may touch t[len_t] = the caller's byte after t).
 * out = { off_p, off_t1, off_t2 (= len_t - suffix offset), edit distance }. */
int po_borders(const char *p, int len_p, int min_cut, int max_cut, const char *t, int len_t, unsigned max_errs,
               int out[4]) {
  int t_win = (int)MIN2((size_t)len_p + max_errs, (size_t)len_t);
  char *rt = malloc((size_t)len_t + 1), *rp = malloc((size_t)len_p + 1);
  for (int i = 0; i < len_t; ++i) rt[len_t - 1 - i] = t[i];
  for (int i = 0; i < len_p; ++i) rp[len_p - 1 - i] = p[i];
  uint32_t *Mp = ed_matrix(p, len_p, t, t_win);    /* rows over p, columns over t */
  uint32_t *Ms = ed_matrix(rp, len_p, rt, t_win);
  size_t W = (size_t)t_win + 1;
  uint32_t *mn[2], *pos[2];
  for (int s = 0; s < 2; ++s) {
    const uint32_t *M = s ? Ms : Mp;
    mn[s] = malloc(((size_t)len_p + 1) * sizeof(uint32_t));
    pos[s] = malloc(((size_t)len_p + 1) * sizeof(uint32_t));
    mn[s][0] = 0; pos[s][0] = 0;
    for (int i = 1; i <= len_p; ++i) {
      uint32_t best = M[i * W], bj = 0;
      for (int j = 1; j <= t_win; ++j)
        if (best > M[i * W + j]) { best = M[i * W + j]; bj = (uint32_t)j; }
      mn[s][i] = best; pos[s][i] = bj;
    }
  }
  int off_p = min_cut;
  size_t off_t1 = pos[0][min_cut], off_t2 = pos[1][len_p - min_cut];
  uint32_t best = mn[0][min_cut] + mn[1][len_p - min_cut];
  int best_freq = burset_adaptor(t, off_t1, (size_t)len_t - off_t2);
  for (int i = min_cut + 1; i <= max_cut; ++i) {
    int freq = burset_adaptor(t, pos[0][i], (size_t)len_t - pos[1][len_p - i]);
    uint32_t c = mn[0][i] + mn[1][len_p - i];
    if (best > c || (best == c && freq > best_freq)) {
      best = c; off_p = i; off_t1 = pos[0][i]; off_t2 = pos[1][len_p - i]; best_freq = freq;
    }
  }
  out[0] = off_p; out[1] = (int)off_t1; out[2] = len_t - (int)off_t2; out[3] = (int)best;
  free(rt); free(rp); free(Mp); free(Ms);
  for (int s = 0; s < 2; ++s) { free(mn[s]); free(pos[s]); }
  return best <= max_errs;
}

/* ---- DP D: three-state intron gap alignment with traceback ---------------------------------------
 * reference src/refine-intron.c:560-890 (compute_gap_alignment, ComputeGapAlignMatrix,
 * TracebackGapAlignment).  States L (before the intron), G (inside: free genome gap), R (after; the
 * last EST row also gets free trailing genome gaps).  Match +1 (N wildcard), mismatch/indel -1.
 * Directions: 0 diag, 1 up, 2 left, 3 = jump to the previous state while moving left (the
 * reference's -2).  pos = { factor_cut, intron_start, intron_end, intron_start_on_align,
 * intron_end_on_align }, all 0 unless the traceback meets the corresponding jump
 * (gap_alignment_create zero-initialises them).  Returns the number of alignment columns. */
int po_gap(const char *est, int n, const char *gen, int m, uint8_t *ops, int pos[5]) {
  size_t W = (size_t)m + 1, cells = (size_t)(n + 1) * W;
  int32_t *L = calloc(cells, sizeof *L), *G = calloc(cells, sizeof *G), *R = calloc(cells, sizeof *R);
  uint8_t *dL = calloc(cells, 1), *dG = calloc(cells, 1), *dR = calloc(cells, 1);
  for (int i = 1; i <= n; ++i)
    for (int j = 1; j <= m; ++j) {
      size_t c = (size_t)i * W + j, up = c - W, lf = c - 1, dg = c - W - 1;
      int s = wild_eq(est[i - 1], gen[j - 1]) ? 1 : -1;
      int32_t v = L[dg] + s; uint8_t d = 0;
      if (v < L[up] - 1) { v = L[up] - 1; d = 1; }
      if (v < L[lf] - 1) { v = L[lf] - 1; d = 2; }
      L[c] = v; dL[c] = d;
      v = G[lf]; d = 2;
      if (v < L[lf]) { v = L[lf]; d = 3; }
      G[c] = v; dG[c] = d;
      v = R[dg] + s; d = 0;
      int32_t hgap = (i != n) ? R[lf] - 1 : R[lf];
      if (v < hgap) { v = hgap; d = 2; }
      if (v < G[lf]) { v = G[lf]; d = 3; }
      if (v < R[up] - 1) { v = R[up] - 1; d = 1; }
      R[c] = v; dR[c] = d;
    }
  size_t last = cells - 1;
  int state;   /* 2 = R, 1 = G, 0 = L; preference R >= G >= L (refine-intron.c:808-819) */
  if (R[last] >= G[last]) state = (R[last] >= L[last]) ? 2 : 0;
  else state = (G[last] >= L[last]) ? 1 : 0;
  memset(pos, 0, 5 * sizeof(int));
  /* iterative traceback, recording columns right-to-left; *_on_align fixed up after the reversal */
  int i = n, j = m, k = 0, k_end = -1, k_start = -1;
  while (i > 0 || j > 0) {
    if (i > 0 && j > 0) {
      size_t c = (size_t)i * W + j;
      uint8_t d = state == 2 ? dR[c] : state == 1 ? dG[c] : dL[c];
      if (d == 0) { ops[k++] = 0; --i; --j; }
      else if (d == 1) { ops[k++] = 1; --i; }
      else {
        if (d == 3) {
          if (state == 2) { pos[2] = j - 1; pos[0] = i; k_end = k; }
          else { pos[1] = j - 1; k_start = k; }
          --state;
        }
        ops[k++] = 2; --j;
      }
    } else if (i > 0) { ops[k++] = 1; --i; }
    else { ops[k++] = 2; --j; }
  }
  for (int a = 0, b = k - 1; a < b; ++a, --b) { uint8_t t = ops[a]; ops[a] = ops[b]; ops[b] = t; }
  if (k_end >= 0) pos[4] = k - 1 - k_end;
  if (k_start >= 0) pos[3] = k - 1 - k_start;
  free(L); free(G); free(R); free(dL); free(dG); free(dR);
  return k;
}

/* ---- DP E: affix recovery -------------------------------------------------------------------------
 * reference src/factorization-refinement.c:1134-1172 (find_longest_affix) over
 * src/compute-alignments.c:210 (edit_distance_matrix): among cells (ecut,gcut) >= (1,1) whose last
 * characters are equal and whose weight 2*D/(ecut+gcut) is <= 0.17, the LAST one in row-major order
 * with minimal weight (the scan accepts "<= best so far"). */
int po_affix(const char *est, int estl, const char *gen, int genl, int *ecut_out, int *gcut_out) {
  uint32_t *M = ed_matrix(est, estl, gen, genl);
  size_t W = (size_t)genl + 1;
  int valid = 0, be = 0, bg = 0; double best = 1.0;
  for (int e = 1; e <= estl; ++e)
    for (int g = 1; g <= genl; ++g) {
      double w = 2.0 * ((double)M[e * W + g]) / (double)((size_t)e + (size_t)g);
      if (est[e - 1] == gen[g - 1] && w <= 0.17 && w <= best) { be = e; bg = g; best = w; valid = 1; }
    }
  free(M);
  if (valid) { *ecut_out = be; *gcut_out = bg; }
  return valid;
}

/* ---- best suffix / prefix cut ---------------------------------------------------------------------
 * reference src/compute-alignments.c:246-316 (compute_best_suffix_cut / compute_best_prefix_cut):
 * minimum over the last column (LAST argmin, ">=") and last row (LAST argmin) of the matrix, the row
 * wins only if strictly smaller. */
unsigned po_suffix_cut(const char *s1, int l1, const char *s2, int l2, int *c1, int *c2) {
  if (l1 == l2 && memcmp(s1, s2, (size_t)l1) == 0) { *c1 = l1; *c2 = l2; return 0; }
  uint32_t *M = ed_matrix(s1, l1, s2, l2);
  size_t W = (size_t)l2 + 1;
  uint32_t corner = M[(size_t)l1 * W + l2], mincol = corner, minrow = corner;
  int colpos = l1, rowpos = l2;
  for (int i = 0; i < l1; ++i) if (mincol >= M[i * W + l2]) { mincol = M[i * W + l2]; colpos = i; }
  for (int j = 0; j < l2; ++j) if (minrow >= M[(size_t)l1 * W + j]) { minrow = M[(size_t)l1 * W + j]; rowpos = j; }
  unsigned ed;
  if (minrow < mincol) { *c1 = l1; *c2 = rowpos; ed = minrow; } else { *c1 = colpos; *c2 = l2; ed = mincol; }
  free(M);
  return ed;
}

unsigned po_prefix_cut(const char *s1, int l1, const char *s2, int l2, int *c1, int *c2) {
  if (l1 == l2 && memcmp(s1, s2, (size_t)l1) == 0) { *c1 = 0; *c2 = 0; return 0; }
  char *r1 = malloc((size_t)l1 + 1), *r2 = malloc((size_t)l2 + 1);
  for (int i = 0; i < l1; ++i) r1[l1 - 1 - i] = s1[i];
  for (int i = 0; i < l2; ++i) r2[l2 - 1 - i] = s2[i];
  unsigned ed = po_suffix_cut(r1, l1, r2, l2, c1, c2);
  *c1 = l1 - *c1; *c2 = l2 - *c2;
  free(r1); free(r2);
  return ed;
}

/* ---- DP F: longest common substring with N wildcards ----------------------------------------------
 * reference src/factorization-refinement.c:255-315 (find_longest_common_factor_dp, built with
 * Ns_ALWAYS_MATCH_FOR_LCS :74): first strictly longer run in (i1 outer, i2 inner) order.  The
 * reference's "swap if l2>l1" branch recurses and then falls through to recompute with the original
 * argument order (:260-262), so the un-swapped scan is the result. */
void po_lcs(const char *s1, long l1, const char *s2, long l2, long *occ1, long *occ2, long *len) {
  long *prev = calloc((size_t)l2 + 1, sizeof *prev), *cur = calloc((size_t)l2 + 1, sizeof *cur);
  long bo1 = 0, bo2 = 0, bl = 0;
  for (long i1 = 0; i1 < l1; ++i1) {
    cur[0] = 0;
    for (long i2 = 0; i2 < l2; ++i2) {
      cur[i2 + 1] = wild_eq(s1[i1], s2[i2]) ? prev[i2] + 1 : 0;
      if (bl < cur[i2 + 1]) { bl = cur[i2 + 1]; bo1 = i1 + 1 - bl; bo2 = i2 + 1 - bl; }
    }
    long *t = prev; prev = cur; cur = t;
  }
  *occ1 = bo1; *occ2 = bo2; *len = bl;
  free(prev); free(cur);
}

/* ---- maximal-pairing discovery (the vertex set of the MEG) -----------------------------------------
 * reference src/max-emb-graph.c:217-380 (build_vertex_set) over the augmented suffix tree
 * (src/aug_suffix_tree.c:151-264); restated as SURVEY.md Appendix A: per EST position p the
 * left-maximal occurrences (t==0 or p==0 or T[t-1]!=P[p-1]) with LCP >= mfl, thresholded at
 * max(floor(D*rate), mfl), sorted by t, filter A inside one p, filter B across adjacent p.
 * Brute force O(|P|*|T|) with a first-character pre-check: for test sizes only.
 * out = (p,t,l) triples; returns their number, or -(needed) if cap is too small. */
typedef struct { int p, t, l; } ptl;

long po_seed(const char *T, long G, const char *P, long n, int mfl, double rate, int *out, long cap) {
  ptl **V = calloc((size_t)n + 1, sizeof *V);
  long *cnt = calloc((size_t)n + 1, sizeof *cnt);
  for (long p = 0; p < n; ++p) {
    long c = 0, capp = 16; ptl *v = malloc((size_t)capp * sizeof *v);
    long D = 0;
    for (long t = 0; t < G; ++t) {
      if (T[t] != P[p]) continue;
      if (!(t == 0 || p == 0 || T[t - 1] != P[p - 1])) continue;
      long l = 0;
      while (p + l < n && t + l < G && P[p + l] == T[t + l]) ++l;
      if (l < mfl) continue;
      if (c == capp) { capp *= 2; v = realloc(v, (size_t)capp * sizeof *v); }
      v[c].p = (int)p; v[c].t = (int)t; v[c].l = (int)l; ++c;
      if (l > D) D = l;
    }
    if (c) {
      size_t thr = (size_t)((double)D * rate);
      if (thr < (size_t)mfl) thr = (size_t)mfl;
      long k = 0;
      for (long i = 0; i < c; ++i) if ((size_t)v[i].l >= thr) v[k++] = v[i];
      c = k;
      /* filter A, judged against the unfiltered thresholded list */
      char *drop = calloc((size_t)c + 1, 1);
      for (long j = 1; j < c; ++j)
        for (long i = 0; i < j && !drop[j]; ++i)
          if ((v[j].t > v[i].t && v[j].t + v[j].l <= v[i].t + v[i].l) || (v[j].t == v[i].t + 1 && v[j].l == v[i].l))
            drop[j] = 1;
      k = 0;
      for (long i = 0; i < c; ++i) if (!drop[i]) v[k++] = v[i];
      c = k;
      free(drop);
    }
    V[p] = v; cnt[p] = c;
  }
  /* filter B: descending p, each list judged against the (not yet B-filtered) list of p-1 */
  for (long p = n - 2; p >= 0; --p) {
    long k = 0;
    for (long x = 0; x < cnt[p + 1]; ++x) {
      int dropit = 0;
      for (long y = 0; y < cnt[p] && !dropit; ++y)
        if (V[p][y].t == V[p + 1][x].t && V[p][y].l >= V[p + 1][x].l) dropit = 1;
      if (!dropit) V[p + 1][k++] = V[p + 1][x];
    }
    cnt[p + 1] = k;
  }
  long tot = 0;
  for (long p = 0; p < n; ++p) {
    for (long i = 0; i < cnt[p]; ++i, ++tot)
      if (tot < cap) { out[3 * tot] = V[p][i].p; out[3 * tot + 1] = V[p][i].t; out[3 * tot + 2] = V[p][i].l; }
    free(V[p]);
  }
  free(V); free(cnt);
  return tot <= cap ? tot : -tot;
}
