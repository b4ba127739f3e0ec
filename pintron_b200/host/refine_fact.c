#define _GNU_SOURCE
/* refine_fact.c — the post-factorization refinement of one EST (reference src/factorization-refinement.c).
 *
 * refine_EST_factorizations :1269-1305 = remove_invalid_factorizations :120, remove_duplicated_factorizations :174,
 * recover_lost_prefixes_and_suffixes :1175 (find_longest_affix :1134 -> PC_OP_AFFIX), remove_false_small_exons :1095
 * (analyze_possibly_small_exon :958 -> PC_OP_EDIT x3 + PC_OP_BORDERS), search_for_new_small_exons :873
 * (search_small_exon_at_prefix :498 -> PC_OP_LCS over the genome prefix + PC_OP_BORDERS; search_small_exon :639 ->
 * PC_OP_EDIT, PC_OP_LCS and the strstr / PWM scan, which stays on the host), clean_factorizations :909; followed, as
 * in src/compute-est-fact.c:169-175, by remove_factorizations_with_very_small_exons :84 and a second de-duplication.
 * Lengths are size_t where the reference's are: several guards rely on unsigned wrap-around.
 */
#include "ef.h"

#define UB_VERY_SMALL_EXON 2
#define LB_SMALL_EXON 6
#define UB_SMALL_EXON 23
#define UB_MED_EXON 100
#define AFFIXES_LENGTH 5
#define MAX_ERROR_RATE 0.17
#define MIN_PERFECT_BORDER 6
#define MAX_ERRORS_AS_SMALL 2

static void remove_invalid(ef_fzlist *L) {
  for (int k = 0; k < L->n;) {
    const ef_fz *z = L->v[k];
    bool bad = false;
    for (int i = 0; i < z->n && !bad; ++i) {
      const ef_factor *f = &z->f[i];
      bad = f->es > f->ee || f->gs > f->ge;
      if (!bad && i > 0) bad = z->f[i - 1].ee >= f->es || z->f[i - 1].ge >= f->gs;
    }
    if (bad) fzl_remove(L, k); else ++k;
  }
}

/* the reference's hash pre-filter never hides a true duplicate, so this is the plain pairwise check: a
 * factorization equal to an earlier surviving one is dropped */
static void remove_duplicates(ef_fzlist *L) {
  for (int k = 0; k < L->n;) {
    bool dup = false;
    for (int q = 0; q < k && !dup; ++q) {
      const ef_fz *a = L->v[k], *b = L->v[q];
      dup = a->n == b->n && memcmp(a->f, b->f, sizeof(ef_factor) * (size_t)a->n) == 0;
    }
    if (dup) fzl_remove(L, k); else ++k;
  }
}

static void remove_very_small_exons(ef_fzlist *L) {
  for (int k = 0; k < L->n;) {
    const ef_fz *z = L->v[k];
    bool small = false;
    for (int i = 0; i < z->n && !small; ++i) small = z->f[i].ee + 1 - z->f[i].es <= UB_VERY_SMALL_EXON;
    if (small) fzl_remove(L, k); else ++k;
  }
}

static char *rev_copy(ef_task *T, const char *s, size_t n) {
  char *r = ar_alloc(&T->ar, n + 1);
  for (size_t i = 0; i < n; ++i) r[i] = s[n - 1 - i];
  return r;
}

/* one AFFIX job per end that still mismatches on its border; all ends of all factorizations in one batch */
static void recover_lost_affixes(ef_task *T, const ef_seq *est, ef_fzlist *L) {
  const char *e = est->seq, *g = T->gen->seq;
  const size_t totelen = (size_t)est->len, totglen = (size_t)T->gen->len;
  int *hp = ar_alloc(&T->ar, sizeof(int) * (size_t)(L->n + 1)), *hs = ar_alloc(&T->ar, sizeof(int) * (size_t)(L->n + 1));
  for (int k = 0; k < L->n; ++k) {
    ef_fz *z = L->v[k];
    hp[k] = hs[k] = -1;
    const ef_factor *ff = &z->f[0], *fl = &z->f[z->n - 1];
    if (ff->es > 0 && ff->gs > 0) {
      const size_t flen = (size_t)MIN2(ff->es, ff->gs);
      const size_t elen = (size_t)MIN2(ff->es, (int)((1.0 + MAX_ERROR_RATE) * flen));
      const size_t glen = (size_t)MIN2(ff->gs, (int)((1.0 + MAX_ERROR_RATE) * flen));
      char *ef = rev_copy(T, e + ff->es - elen, elen), *gf = rev_copy(T, g + ff->gs - glen, glen);
      if (ef[0] != gf[0]) hp[k] = dp_push(PC_OP_AFFIX, S_(ef, (int)elen), S_(gf, (int)glen), 0, 0, 0, 0);
    }
    if (totelen - (size_t)fl->ee > 1 && totglen - (size_t)fl->ge > 1) {
      const size_t flen = MIN2(totelen - (size_t)fl->ee - 1, totglen - (size_t)fl->ge - 1);
      const size_t elen = MIN2(totelen - (size_t)fl->ee - 1, (size_t)((int)(1.0 + MAX_ERROR_RATE)) * flen);   /* the cast binds first */
      const size_t glen = MIN2(totglen - (size_t)fl->ge - 1, (size_t)((int)(1.0 + MAX_ERROR_RATE)) * flen);
      if (e[fl->ee] != g[fl->ge]) hs[k] = dp_push(PC_OP_AFFIX, S_(e + fl->ee, (int)elen), S_(g + fl->ge, (int)glen), 0, 0, 0, 0);
    }
  }
  dp_wait();
  for (int k = 0; k < L->n; ++k) {
    ef_fz *z = L->v[k];
    if (hp[k] >= 0 && dp_res(hp[k])[1]) { z->f[0].es -= dp_res(hp[k])[2]; z->f[0].gs -= dp_res(hp[k])[3]; }
    if (hs[k] >= 0 && dp_res(hs[k])[1]) { z->f[z->n - 1].ee += dp_res(hs[k])[2]; z->f[z->n - 1].ge += dp_res(hs[k])[3]; }
  }
}

/* compute_edit_distance (compute-alignments.c:235): plain edit distance, no wildcard */
static int edit_push(const char *a, size_t la, const char *b, size_t lb) {
  return dp_push(PC_OP_EDIT, S_(a, (int)la), S_(b, (int)lb), 0, 0, 0, 0);
}

/* analyze_possibly_small_exon: try to drop internal exon `c` of z by re-placing it across ONE intron; true if removed */
static bool analyze_small_exon(ef_task *T, const ef_seq *est, ef_fz *z, int c) {
  if (c == 0 || c == z->n - 1) return false;
  const char *e = est->seq, *g = T->gen->seq;
  ef_factor *prev = &z->f[c - 1], *cur = &z->f[c], *next = &z->f[c + 1];
  const size_t elen = (size_t)(cur->ee + 1 - cur->es), glen = (size_t)(cur->ge + 1 - cur->gs);
  if (elen > UB_MED_EXON) return false;
  const size_t estart = (size_t)MAX2(prev->es + 1, prev->ee + 1 - AFFIXES_LENGTH), eend = (size_t)MIN2(next->ee, next->es + AFFIXES_LENGTH);
  const size_t epreflen = (size_t)prev->ee + 1 - estart, esufflen = eend - (size_t)next->es, allelen = eend - estart;
  const char *allefact = e + estart;
  const size_t gstart = (size_t)MAX2(prev->gs + 1, prev->ge + 1 - AFFIXES_LENGTH), gend = (size_t)MIN2(next->ge, next->gs + AFFIXES_LENGTH);
  const size_t gpreflen = (size_t)prev->ge + 1 - gstart, gsufflen = gend - (size_t)next->gs, allglen = gend - gstart;
  const char *allgfact = g + gstart;
  const int h0 = edit_push(e + cur->es, elen, g + cur->gs, glen);
  const int h1 = edit_push(allefact, epreflen, allgfact, gpreflen);
  const int h2 = edit_push(allefact - esufflen, esufflen, allgfact - gsufflen, gsufflen);   /* sic: the bytes BEFORE the window */
  dp_wait();
  const size_t orig = (size_t)dp_res(h0)[1] + (size_t)dp_res(h1)[1] + (size_t)dp_res(h2)[1];
  int off_p, off_t1, off_t2;
  unsigned new_ed;
  if (!dp_borders(allefact, (int)allelen, 0, (int)allelen, allgfact, (int)allglen, (unsigned)orig, &off_p, &off_t1, &off_t2, &new_ed))
    return false;
  const double prev_avg = (burset_adaptor(g, (size_t)prev->ge + 1, (size_t)cur->gs) + burset_adaptor(g, (size_t)cur->ge + 1, (size_t)next->gs)) / 2.0;
  const double new_freq = burset_adaptor(g, gstart + (size_t)off_t1, gend - allglen + (size_t)off_t2);
  if (!(new_freq >= prev_avg)) return false;
  prev->ee = (int)(estart + (size_t)off_p - 1);
  next->es = (int)(eend + (size_t)off_p - allelen);
  prev->ge = (int)(gstart + (size_t)off_t1 - 1);
  next->gs = (int)(gend + (size_t)off_t2 - allglen);
  fz_remove(z, c);
  return true;
}

typedef struct par_fz { const ef_seq *est; ef_fzlist *L; } par_fz;      /* loops over factorizations: each iteration touches its own only */
static void false_small_exons_of(ef_task *T, int k, void *user) {
  const par_fz *P = user;
  ef_fz *z = P->L->v[k];
  /* after a removal the exon before the removed one is examined again against the same successor */
  for (int c = 0; c < z->n;) {
    if (analyze_small_exon(T, P->est, z, c)) --c; else ++c;
  }
}
static void remove_false_small_exons(ef_task *T, const ef_seq *est, ef_fzlist *L) {
  par_fz P = {est, L};
  dp_parallel_for(T, L->n, false_small_exons_of, &P);
}

static bool canonical_intron(const char *g, size_t s, size_t e) {
  return (g[s] == 'G' && g[s + 1] == 'T' && g[e - 1] == 'A' && g[e] == 'G') || (g[s] == 'g' && g[s + 1] == 't' && g[e - 1] == 'a' && g[e] == 'g');
}

/* search_small_exon_at_prefix: the unaligned EST prefix (<= 46 nt) against the WHOLE genome prefix before exon 0 */
static void small_exon_at_prefix(ef_task *T, const ef_seq *est, ef_fz *z) {
  const ef_config *c = T->cfg;
  const char *e = est->seq, *g = T->gen->seq;
  ef_factor *p1 = &z->f[0];
  const size_t e1len = (size_t)(p1->ee + 1 - p1->es), g1len = (size_t)(p1->ge + 1 - p1->gs);
  if (e1len + (size_t)p1->es < LB_SMALL_EXON + UB_SMALL_EXON) return;
  const size_t eplen = (size_t)MIN2(MIN2(p1->es, p1->gs), 2 * UB_SMALL_EXON);
  const char *epfact = e + p1->es - eplen;
  const size_t e1plen = MIN2(MIN2(e1len, g1len), (size_t)UB_SMALL_EXON);
  long pg_, pe_, cflen_;
  {
    ef_str s1 = {g, p1->gs, true, 0, false};
    const int h = dp_push(PC_OP_LCS, S_(epfact, (int)eplen), s1, 0, 0, 0, 0);
    dp_wait();
    cflen_ = dp_res(h)[1]; pg_ = dp_res(h)[2]; pe_ = dp_res(h)[3];
  }
  const size_t pg = (size_t)pg_, pe = (size_t)pe_, cflen = (size_t)cflen_;
  if (cflen < LB_SMALL_EXON) return;
  const int hed = edit_push(e + p1->es, e1plen, g + p1->gs, e1plen);
  dp_wait();
  const unsigned edp = (unsigned)dp_res(hed)[1];
  const size_t allelen = (size_t)MIN2(p1->ee + 1, p1->es + UB_SMALL_EXON) - pe;
  const size_t allglen = (size_t)MIN2(p1->ge + 1, p1->gs + UB_SMALL_EXON) - pg;
  int off_p, off_t1, off_t2;
  unsigned new_ed;
  const bool ok = dp_borders(e + pe, (int)allelen, LB_SMALL_EXON, (int)(allelen - LB_SMALL_EXON), g + pg, (int)allglen, edp,
                             &off_p, &off_t1, &off_t2, &new_ed);
  if (!ok) return;
  if (off_t2 - off_t1 < c->min_intron_length) return;
  if (!canonical_intron(g, pg + (size_t)off_t1, pg + (size_t)off_t2 - 1)) return;
  if ((size_t)off_p - pe < LB_SMALL_EXON) return;
  ef_factor nw = {(int)pe, (int)(pe + (size_t)off_p - 1), (int)pg, (int)(pg + (size_t)off_t1 - 1)};
  p1->es = (int)(pe + (size_t)off_p);
  p1->gs = (int)(pg + (size_t)off_t2);
  fz_insert(T, z, 0, nw);
}

static size_t min3z(size_t a, size_t b, size_t c) { size_t t = a; if (t > b) t = b; if (t > c) t = c; return t; }

/* ---- 6-mer position index of the genome (host side) -------------------------------------------------------------
 * Every trim searched by ef_small_exon_scan is at least LB_SMALL_EXON = 6 letters long, so its occurrences are among
 * the positions of its first 6-mer.  Built once per genome (counting sort: positions ascending inside a bucket);
 * windows with a byte outside upper-case ACGT are not indexed, and a trim that starts with such a byte is searched
 * with memmem instead. */
static struct { const char *g; size_t len; uint32_t *start, *pos; } g_kidx;
static inline int nt_code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }
static inline int kmer6_code(const char *s) {
  int code = 0;
  for (int i = 0; i < 6; ++i) { const int c = nt_code(s[i]); if (c < 0) return -1; code = code * 4 + c; }
  return code;
}

void ef_small_exon_index_build(const char *g, size_t len) {
  free(g_kidx.start); free(g_kidx.pos);
  g_kidx.start = g_kidx.pos = NULL;
  g_kidx.g = g; g_kidx.len = len;
  if (!g) return;                                             /* (NULL, 0) drops the index */
  g_kidx.start = calloc(4097 + 1, sizeof(uint32_t));
  g_kidx.pos = malloc(sizeof(uint32_t) * (len ? len : 1));
  if (!g_kidx.start || !g_kidx.pos) { fprintf(stderr, "* FATAL est-fact: out of memory\n"); exit(1); }
  if (len < 6) return;
  for (int sweep = 0; sweep < 2; ++sweep) {
    int code = 0, valid = 0;                                /* rolling: valid = letters of the current run of ACGT */
    for (size_t i = 0; i < len; ++i) {
      const int c = nt_code(g[i]);
      if (c < 0) { valid = 0; code = 0; continue; }
      code = ((code << 2) | c) & 4095;
      if (++valid >= 6) {
        if (sweep == 0) ++g_kidx.start[code + 1];
        else g_kidx.pos[g_kidx.start[code]++] = (uint32_t)(i - 5);
      }
    }
    if (sweep == 0) for (int k = 0; k < 4096; ++k) g_kidx.start[k + 1] += g_kidx.start[k];
    else { for (int k = 4096; k > 0; --k) g_kidx.start[k] = g_kidx.start[k - 1]; g_kidx.start[0] = 0; }   /* undo the cursor advance */
  }
}

/* The scan of search_small_exon (factorization-refinement.c:770-834): the longest trim (offstart, offend) of the EST
 * middle e[estart, estart+elen) that occurs inside g[allgstart, allgstart+allglen) leaving two introns that both
 * classify.  out = {max_sexon, ecut1, ecut2, gcut1_1, gcut1_2, gcut2_1, gcut2_2}.
 * The reference runs strstr for every (offstart, offend) and keeps the first strictly longer hit.  Same result, fewer
 * scans: an occurrence at q of the trim (offstart, offend), length sl = elen - offstart - offend, must satisfy
 * offstart + MINI <= q <= allglen - MINI - elen + offstart (offend cancels out of the right limit), so for one
 * offstart every longer trim occurs where the shortest one does: the occurrences of the shortest trim (from the 6-mer
 * index, else one memmem pass) find all of them, and trims that cannot beat the current best are skipped
 * (classify_intron is pure). */
typedef struct se_state {
  const char *g, *pat, *allgfact;
  int glen_all;
  size_t elen, offstart, allgstart, allglen, sl_min, sl_max, max_sexon, best_here, b_q, b_offend;
} se_state;

static inline void se_occurrence(se_state *S, size_t q) {
  size_t ext = S->sl_min;
  while (ext < S->sl_max && S->pat[ext] == S->allgfact[q + ext]) ++ext;
  /* trims that occur at q: sl_min <= sl <= ext, longest (= smallest offend) first; only sl > best so far matter:
   * for equal sl an earlier q wins, and earlier offstarts already hold max_sexon */
  const size_t floor_sl = MAX2(S->max_sexon, S->best_here);
  for (size_t sl = ext; sl >= S->sl_min && sl > floor_sl; --sl) {
    const size_t offend = S->elen - S->offstart - sl;
    const size_t i1start = S->allgstart + S->offstart, i1end = S->allgstart + q - 1;
    const size_t i2start = i1end + 1 + sl, i2end = S->allgstart + S->allglen - offend - 1;
    const char t1 = classify_intron(S->g, S->glen_all, (int)i1start, (int)i1end), t2 = classify_intron(S->g, S->glen_all, (int)i2start, (int)i2end);
    if (t1 != 2 && t2 != 2) { S->best_here = sl; S->b_q = q; S->b_offend = offend; break; }
  }
}

void ef_small_exon_scan(const char *g, int glen_all, const char *e, size_t estart, size_t elen, size_t allgstart, size_t allglen,
                        size_t f1slen, size_t f2plen, size_t MINI, size_t out[7]) {
  const char *efact = e + estart, *allgfact = g + allgstart;
  size_t max_sexon = 0, ecut1 = 0, ecut2 = 0, gcut1_1 = 0, gcut1_2 = 0, gcut2_1 = 0, gcut2_2 = 0;
  const size_t max_offstart = min3z(f1slen + 1 - MIN_PERFECT_BORDER, elen + 1 - LB_SMALL_EXON, allglen + 1 - (2 * MINI) - LB_SMALL_EXON);
  const bool have_index = g_kidx.g == g && g_kidx.start != NULL;
  for (size_t offstart = 0; offstart < max_offstart; ++offstart) {
    const size_t max_offend = min3z(f2plen + 1 - MIN_PERFECT_BORDER, elen + 1 - offstart - LB_SMALL_EXON,
                                    allglen + 1 - (2 * MINI) - LB_SMALL_EXON - offstart);
    const size_t sl_max = elen - offstart, sl_min = elen - offstart - (max_offend - 1);
    if (sl_max <= max_sexon) continue;
    if (allglen + offstart < MINI + elen) continue;                 /* no room for any occurrence */
    const size_t qmin = offstart + MINI, qmax = allglen - MINI - elen + offstart;
    if (qmax < qmin) continue;
    se_state S = {g, efact + offstart, allgfact, glen_all, elen, offstart, allgstart, allglen, sl_min, sl_max, max_sexon, 0, 0, 0};
    const int code = have_index ? kmer6_code(S.pat) : -1;
    if (code >= 0) {
      /* occurrences of the shortest trim, ascending: bucket positions inside [allgstart + qmin, allgstart + qmax] */
      const uint32_t *lo = g_kidx.pos + g_kidx.start[code], *hi = g_kidx.pos + g_kidx.start[code + 1];
      const size_t pmin = allgstart + qmin, pmax = allgstart + qmax;
      while (lo < hi) { const uint32_t *mid = lo + (hi - lo) / 2; if (*mid < pmin) lo = mid + 1; else hi = mid; }
      for (const uint32_t *it = lo, *end = g_kidx.pos + g_kidx.start[code + 1]; it < end && *it <= pmax; ++it)
        if (sl_min == 6 || memcmp(g + *it + 6, S.pat + 6, sl_min - 6) == 0) se_occurrence(&S, (size_t)*it - allgstart);
    } else {
      const char *scan = allgfact + qmin, *scan_end = allgfact + qmax + sl_min;      /* last window ends here */
      while (scan + sl_min <= scan_end) {
        const char *occ = memmem(scan, (size_t)(scan_end - scan), S.pat, sl_min);
        if (!occ) break;
        se_occurrence(&S, (size_t)(occ - allgfact));
        scan = occ + 1;
      }
    }
    if (S.best_here > max_sexon) {
      const size_t sl = S.best_here, i1start = allgstart + offstart, i1end = allgstart + S.b_q - 1;
      const size_t i2start = i1end + 1 + sl, i2end = allgstart + allglen - S.b_offend - 1;
      max_sexon = sl; ecut1 = estart + offstart; ecut2 = estart + offstart + sl;
      gcut1_1 = i1start; gcut1_2 = i1end + 1; gcut2_1 = i2start; gcut2_2 = i2end + 1;
    }
  }
  out[0] = max_sexon; out[1] = ecut1; out[2] = ecut2; out[3] = gcut1_1; out[4] = gcut1_2; out[5] = gcut2_1; out[6] = gcut2_2;
}

/* search_small_exon between exons i and i+1 of z; returns true when a new exon was inserted at i+1 */
static bool small_exon_between(ef_task *T, const ef_seq *est, ef_fz *z, int i) {
  const ef_config *c = T->cfg;
  const char *e = est->seq, *g = T->gen->seq;
  const int glen_all = T->gen->len;
  ef_factor *p1 = &z->f[i], *p2 = &z->f[i + 1];
  const size_t e1len = (size_t)(p1->ee + 1 - p1->es), g1len = (size_t)(p1->ge + 1 - p1->gs);
  const size_t e2len = (size_t)(p2->ee + 1 - p2->es), g2len = (size_t)(p2->ge + 1 - p2->gs);
  if (e1len + e2len < LB_SMALL_EXON + 2 * UB_SMALL_EXON) return false;
  const size_t e1slen = MIN2(MIN2(e1len, g1len), (size_t)UB_SMALL_EXON), g1slen = e1slen;
  const size_t e1sstart = (size_t)p1->ee + 1 - e1slen, g1sstart = (size_t)p1->ge + 1 - g1slen;
  const char *e1sfact = e + e1sstart, *g1sfact = g + g1sstart;
  const size_t e2plen = MIN2(MIN2(e2len, g2len), (size_t)UB_SMALL_EXON), g2plen = e2plen;
  const size_t e2pstart = (size_t)p2->es, g2pstart = (size_t)p2->gs;
  const char *e2pfact = e + e2pstart, *g2pfact = g + g2pstart;
  const int hs = edit_push(e1sfact, e1slen, g1sfact, g1slen), hp = edit_push(e2pfact, e2plen, g2pfact, g2plen);
  dp_wait();
  const size_t sed = (size_t)dp_res(hs)[1], ped = (size_t)dp_res(hp)[1];
  bool go = false;
  const char orig_type = classify_intron(g, glen_all, p1->ge + 1, p2->gs - 1);
  if (sed + ped > MAX_ERRORS_AS_SMALL) go = true;
  if (orig_type == 2) go = true;
  if (!go) return false;
  /* the longest common factors of the two borders are needed only from here on, and only for a border with errors */
  int hl1 = -1, hl2 = -1;
  if (sed > 0) hl1 = dp_push(PC_OP_LCS, S_(g1sfact, (int)g1slen), S_(e1sfact, (int)e1slen), 0, 0, 0, 0);
  if (ped > 0) hl2 = dp_push(PC_OP_LCS, S_(g2pfact, (int)g2plen), S_(e2pfact, (int)e2plen), 0, 0, 0, 0);
  if (hl1 >= 0 || hl2 >= 0) dp_wait();
  size_t e1socc = 0, g1socc = 0, f1slen = e1slen;
  if (sed > 0) { f1slen = (size_t)dp_res(hl1)[1]; e1socc = (size_t)dp_res(hl1)[2]; g1socc = (size_t)dp_res(hl1)[3]; }
  size_t e2pocc = 0, g2pocc = 0, f2plen = e2plen;
  if (ped > 0) { f2plen = (size_t)dp_res(hl2)[1]; e2pocc = (size_t)dp_res(hl2)[2]; g2pocc = (size_t)dp_res(hl2)[3]; }
  if (f1slen == e1slen && e2pocc > 0) {
    size_t nf = f1slen + 1;
    while (nf - f1slen < e2pocc && e[e1sstart + e1socc + f1slen] == g[g2pstart + nf - f1slen]) ++nf;
    if (nf - 1 > f1slen) f1slen = nf - 1;
  }
  const size_t elen = (e1slen - e1socc) + (e2pocc + f2plen) - (2 * MIN_PERFECT_BORDER);
  const size_t estart = e1sstart + e1socc + MIN_PERFECT_BORDER;
  const size_t allgstart = g1sstart + g1socc + MIN_PERFECT_BORDER;
  const size_t allglen = g2pstart + g2pocc + f2plen - MIN_PERFECT_BORDER - allgstart;
  const size_t MINI = (size_t)MAX2(4, c->min_intron_length);
  if (f1slen < MIN_PERFECT_BORDER || f2plen < MIN_PERFECT_BORDER || allglen < 2 * MINI + LB_SMALL_EXON || elen < LB_SMALL_EXON) return false;
  size_t cut[7];
  ef_small_exon_scan(g, glen_all, e, estart, elen, allgstart, allglen, f1slen, f2plen, MINI, cut);
  const size_t max_sexon = cut[0], ecut1 = cut[1], ecut2 = cut[2], gcut1_1 = cut[3], gcut1_2 = cut[4], gcut2_1 = cut[5], gcut2_2 = cut[6];
  if (max_sexon < LB_SMALL_EXON) return false;
  ef_factor nw = {(int)ecut1, (int)ecut2 - 1, (int)gcut1_2, (int)gcut2_1 - 1};
  p2->es = (int)ecut2; p2->gs = (int)gcut2_2;
  p1->ee = (int)ecut1 - 1; p1->ge = (int)gcut1_1 - 1;
  fz_insert(T, z, i + 1, nw);
  return true;
}

static void new_small_exons_of(ef_task *T, int k, void *user) {
  const par_fz *P = user;
  const ef_seq *est = P->est;
  ef_fz *z = P->L->v[k];
  if (z->n == 0) return;
  int first = 0;
  if (z->f[0].es > LB_SMALL_EXON) {
    const int before = z->n;
    small_exon_at_prefix(T, est, z);
    first = z->n - before;                /* the old first exon moved to index 1 when a new one was put in front */
  }
  for (int i = first; i + 1 < z->n;) {
    if (small_exon_between(T, est, z, i)) i += 2; else ++i;     /* the inserted exon is not examined */
  }
}
static void search_new_small_exons(ef_task *T, const ef_seq *est, ef_fzlist *L) {
  par_fz P = {est, L};
  dp_parallel_for(T, L->n, new_small_exons_of, &P);
}

static void clean_one(ef_task *T, int k, void *user) {
  const par_fz *P = user;
  ef_fz *z = P->L->v[k];
  clean_noisy_exons(T, z, T->gen->seq, P->est->orig);
  clean_external_exons(T, z, T->gen->seq, P->est->orig);
}

/* clean_factorizations :909-946 — the noisy / external exon cleaning again, this time against the ORIGINAL EST bytes */
static void clean_all(ef_task *T, const ef_seq *est, ef_fzlist *L) {
  ef_fzlist out = {0};
  { par_fz P = {est, L}; dp_parallel_for(T, L->n, clean_one, &P); }      /* the cleaning is per factorization ... */
  for (int k = 0; k < L->n; ++k) {                                       /* ... the containment test goes in list order */
    ef_fz *z = L->v[k];
    if (z->n == 0) continue;
    add_if_not_exists(T, z, &out);
  }
  *L = out;
}

void refine_factorizations(ef_task *T, const ef_seq *est, ef_fzlist *L) {
  remove_invalid(L);
  remove_duplicates(L);
  recover_lost_affixes(T, est, L);
  remove_false_small_exons(T, est, L);
  remove_duplicates(L);
  ef_phase(EF_PH_SMALLEX);
  search_new_small_exons(T, est, L);
  ef_phase(EF_PH_REFINE);
  clean_all(T, est, L);
  remove_very_small_exons(L);
  if (L->n) remove_duplicates(L);
}
