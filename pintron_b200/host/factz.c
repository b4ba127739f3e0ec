/* factz.c — from the MEG to the filtered factorizations of one EST.
 *
 * Follows reference src/est-factorizations.c: get_EST_factorizations :126-594, get_subtree_embeddings :597-762,
 * update_embedding :765-917, maximality_relation :1362-1460, get_factorizations_from_embeddings :1292-1356, and the
 * per-candidate checks handle_endpoints :2127, clean_external_exons :1706, clean_low_complexity_exons_2 :1667,
 * clean_noisy_exons :1842, check_est_coverage :2303, add_if_not_exists :2041 (with list.c:320-487 relaxed
 * containment), check_gap_errors :1462, plus detect-polya.c:42-166.
 * The candidates of one EST are independent until add_if_not_exists, so each check runs as one phase over all of
 * them and its DP jobs leave as one batch.
 */
#include "ef.h"
#include <math.h>

/* ---- small containers --------------------------------------------------------------------------------------- */
ef_fz *fz_new(ef_task *T, int cap) {
  ef_fz *z = ar_alloc(&T->ar, sizeof *z);
  z->cap = cap < 4 ? 4 : cap;
  z->f = ar_alloc(&T->ar, sizeof(ef_factor) * (size_t)z->cap);
  return z;
}

static void fz_grow(ef_task *T, ef_fz *z) {
  ef_factor *nf = ar_alloc(&T->ar, sizeof(ef_factor) * (size_t)z->cap * 2);
  memcpy(nf, z->f, sizeof(ef_factor) * (size_t)z->n);
  z->f = nf; z->cap *= 2;
}

void fz_push(ef_task *T, ef_fz *z, ef_factor f) { if (z->n == z->cap) fz_grow(T, z); z->f[z->n++] = f; }

void fz_insert(ef_task *T, ef_fz *z, int at, ef_factor f) {
  if (z->n == z->cap) fz_grow(T, z);
  memmove(z->f + at + 1, z->f + at, sizeof(ef_factor) * (size_t)(z->n - at));
  z->f[at] = f; ++z->n;
}

void fz_remove(ef_fz *z, int at) {
  memmove(z->f + at, z->f + at + 1, sizeof(ef_factor) * (size_t)(z->n - at - 1));
  --z->n;
}

void fzl_push(ef_task *T, ef_fzlist *L, ef_fz *z) {
  if (L->n == L->cap) {
    int nc = L->cap ? L->cap * 2 : 8;
    ef_fz **nv = ar_alloc(&T->ar, sizeof(ef_fz *) * (size_t)nc);
    if (L->n) memcpy(nv, L->v, sizeof(ef_fz *) * (size_t)L->n);
    L->v = nv; L->cap = nc;
  }
  L->v[L->n++] = z;
}

void fzl_remove(ef_fzlist *L, int at) {
  memmove(L->v + at, L->v + at + 1, sizeof(ef_fz *) * (size_t)(L->n - at - 1));
  --L->n;
}

bool ef_timeout_expired(ef_task *T) {
  return ef_ticks_to_s(ef_task_ticks(T) - T->t_start) > (double)T->cfg->max_single_factorization_time;
}

/* ---- embeddings ------------------------------------------------------------------------------------------------ */
typedef struct ptl { int p, t, l; } ptl;
typedef struct emb { ptl *v; int n; } emb;                  /* v[0] is the head (leftmost pairing) */
typedef struct emblist { emb **v; int n, cap; } emblist;

static void el_push(ef_task *T, emblist *L, emb *e) {
  if (L->n == L->cap) {
    int nc = L->cap ? L->cap * 2 : 4;
    emb **nv = ar_alloc_raw(&T->ar, sizeof(emb *) * (size_t)nc);
    if (L->n) memcpy(nv, L->v, sizeof(emb *) * (size_t)L->n);
    L->v = nv; L->cap = nc;
  }
  L->v[L->n++] = e;
}

/* one allocation per embedding: header and pairings side by side (not zeroed: every field is written here) */
static emb *emb_prepend(ef_task *T, const emb *tail, ptl head) {
  const int n = (tail ? tail->n : 0) + 1;
  emb *e = ar_alloc_raw(&T->ar, sizeof *e + sizeof(ptl) * (size_t)n);
  e->n = n;
  e->v = (ptl *)(e + 1);
  e->v[0] = head;
  if (tail) memcpy(e->v + 1, tail->v, sizeof(ptl) * (size_t)tail->n);
  return e;
}

static bool covers(const ptl *in, const ptl *out) {   /* `in` lies inside `out` on P and on T */
  return !(in->p < out->p || in->p + in->l > out->p + out->l || in->t < out->t || in->t + in->l > out->t + out->l);
}

/* Every one of the first m elements of `in` lies inside its counterpart of `out` (maximality_relation's loops,
 * est-factorizations.c:1362).  A conjunction, so the order of evaluation is free: embeddings that leave the same node
 * mostly share their first elements and part company further down, so after the two leading elements the scan runs
 * from the far end backwards. */
static bool covers_all(const ptl *in, const ptl *out, int m) {
  if (m > 0 && !covers(&in[0], &out[0])) return false;
  if (m > 1 && !covers(&in[1], &out[1])) return false;
  for (int i = m - 1; i >= 2; --i) if (!covers(&in[i], &out[i])) return false;
  return true;
}

/* 2: add dominates cmp, 0: cmp dominates add, 1: neither (maximality_relation) */
static int maximality(const emb *add, const emb *cmp) {
  const int m = MIN2(add->n, cmp->n);
  if (add->n > cmp->n) return covers_all(cmp->v, add->v, m) ? 2 : 1;
  if (add->n < cmp->n) return covers_all(add->v, cmp->v, m) ? 0 : 1;
  if (covers_all(add->v, cmp->v, m)) return 0;
  return covers_all(cmp->v, add->v, m) ? 2 : 1;
}

/* update_embedding: try to put `node` in front of `e`; NULL = not compatible */
static emb *link_embedding(ef_task *T, const emb *e, const ef_pairing *node) {
  const ef_config *c = T->cfg;
  const char *G = T->gen->seq;
  const ptl head = e->v[0];
  const ptl nd = {node->p, node->t, node->l};
  if (head.p == SINK_START) return node->p >= 0 ? emb_prepend(T, NULL, nd) : NULL;
  if (node->p < 0) { emb *cp = ar_alloc_raw(&T->ar, sizeof *cp); *cp = *e; return cp; }     /* the source adds nothing: same pairings */
  const int small_delta = head.p + head.l - nd.p, big_delta = head.t + head.l - nd.t;
  const int min_fl = (int)c->min_factor_len, fl = 2 * min_fl;
  if (!(small_delta >= fl && big_delta >= fl)) return NULL;
  if (!(small_delta - (nd.l + head.l) <= fl)) return NULL;
  if (!(small_delta - big_delta <= fl)) return NULL;
  int hl, hp, ht, nl;
  if (small_delta >= nd.l + head.l && big_delta >= nd.l + head.l) { hp = head.p; ht = head.t; hl = head.l; nl = nd.l; }
  else {
    const int ref = MIN2(small_delta, big_delta);
    int ln = ref / 2, lh = ref - ln;
    if (ln > nd.l) { ln = nd.l; lh = ref - ln; }
    else if (lh > head.l) { lh = head.l; ln = ref - lh; }
    hl = lh; hp = head.p + head.l - hl; ht = head.t + head.l - hl; nl = ln;
  }
  const bool overlap_p = small_delta < nd.l + head.l;
  const int gap_p = hp - nd.p - nl - 1, gap_t = ht - nd.t - nl - 1;
  const int intron_len = gap_t - MAX2(0, gap_p);
  const bool intron_t = intron_len >= 0 && (c->min_intron_length == 0 || intron_len >= c->min_intron_length);
  if (overlap_p && intron_t) {              /* slide the cut inside the overlap to the best Burset dinucleotides (last best wins) */
    int best = -1, best_cut = 0;
    const int lo = MAX2(nd.p + min_fl, head.p), hi = MIN2(head.p + head.l - min_fl, nd.p + nd.l);
    for (int cut = lo; cut <= hi; ++cut) {
      const int f = burset_adaptor(G, (size_t)(cut - nd.p + nd.t), (size_t)(cut - head.p + head.t));
      if (f >= best) { best = f; best_cut = cut; }
    }
    const int dh = best_cut - head.p;
    hl = head.l - dh; hp = head.p + dh; ht = head.t + dh;
    nl = nd.l - (nd.p + nd.l - best_cut);
  }
  if (!(gap_t <= fl || intron_t)) return NULL;
  emb *out = emb_prepend(T, e, (ptl){nd.p, nd.t, nl});
  out->v[1] = (ptl){hp, ht, hl};
  return out;
}

static emblist *subtree_embeddings(ef_task *T, ef_pairing *root, unsigned *tick) {
  if (root->memo) return root->memo;
  if (ef_timeout_expired(T)) return NULL;
  emblist *L = ar_alloc(&T->ar, sizeof *L);
  root->visited = true;
  if (root->adjs.n == 0) el_push(T, L, emb_prepend(T, NULL, (ptl){root->p, root->t, root->l}));
  for (int a = 0; a < root->adjs.n; ++a) {
    emblist *sub = subtree_embeddings(T, root->adjs.v[a], tick);
    if (!sub) return NULL;
    for (int s = 0; s < sub->n; ++s) {
      emb *add = link_embedding(T, sub->v[s], root);
      if (!add) continue;
      if (!*tick && ef_timeout_expired(T)) return NULL;
      *tick = (*tick + 1) & 1023u;
      int rel = 2;
      for (int k = 0; k < L->n && rel >= 1;) {
        rel = maximality(add, L->v[k]);
        if (rel == 2) { memmove(L->v + k, L->v + k + 1, sizeof(emb *) * (size_t)(L->n - k - 1)); --L->n; }
        else ++k;
      }
      if (rel >= 1) el_push(T, L, add);
      else ar_undo(&T->ar, add);          /* dominated: its bytes are the latest allocation, reuse them */
    }
  }
  root->memo = L;
  return L;
}

static ef_fz *factorization_of(ef_task *T, const emb *e) {
  const int fl = 2 * (int)T->cfg->min_factor_len;
  ef_fz *z = fz_new(T, e->n);
  for (int i = 0; i < e->n; ++i) {
    const ptl q = e->v[i];
    if (z->n == 0 || q.t - z->f[z->n - 1].ge - 1 > fl) fz_push(T, z, (ef_factor){q.p, q.p + q.l - 1, q.t, q.t + q.l - 1});
    else { z->f[z->n - 1].ee = q.p + q.l - 1; z->f[z->n - 1].ge = q.t + q.l - 1; }
  }
  return z;
}

/* ---- per-candidate checks --------------------------------------------------------------------------------------- */
static bool exon_bounds_ok(const ef_fz *z) {      /* check_exon_start_end */
  int pe = -1, pg = -1;
  for (int i = 0; i < z->n; ++i) {
    const ef_factor *f = &z->f[i];
    if (f->es > f->ee || f->gs > f->ge || f->es < pe || f->gs < pg) return false;
    pe = f->ee; pg = f->ge;
  }
  return true;
}

/* handle_endpoints, head side: advance to the first run of more than 5 matching columns */
static void trim_head(ef_fz *z, const ef_aln *A) {
  ef_factor *h = &z->f[0];
  int j = 0, matches = 0, cf = h->es, ce = h->gs;
  bool stop = false;
  while (j < A->dim && !stop) {
    if (matches > 5) stop = true;
    else {
      if (A->est[j] == A->gen[j]) { ++cf; ++ce; ++matches; }
      else { if (A->est[j] != '-') ++cf; if (A->gen[j] != '-') ++ce; matches = 0; }
      ++j;
    }
  }
  if (!stop) fz_remove(z, 0);
  else { h->es = cf - matches; h->gs = ce - matches; }
}

/* handle_endpoints, tail side: retreat to a run of more than 10 matches, then slide the cleavage over gap columns */
static void trim_tail(ef_fz *z, ef_aln *A) {
  ef_factor *t = &z->f[z->n - 1];
  int j = A->dim - 1, matches = 0, cf = t->ee, ce = t->ge;
  bool stop = false;
  while (j >= 0 && !stop) {
    if (matches > 10) stop = true;
    else {
      if (A->est[j] == A->gen[j]) { --cf; --ce; ++matches; }
      else { if (A->est[j] != '-') --cf; if (A->gen[j] != '-') --ce; matches = 0; }
      --j;
    }
  }
  int est_cl = cf + matches, gen_cl = ce + matches, cur = j + matches + 1;
  stop = false;
  while ((A->est[cur] == '-' || A->gen[cur] == '-') && cur < A->dim - 1 && !stop) {
    char *gap_row = A->est[cur] == '-' ? A->est : A->gen, *other = A->est[cur] == '-' ? A->gen : A->est;
    int tr = cur + 1;
    while (gap_row[tr] == '-') ++tr;
    if (tr < A->dim && gap_row[tr] == other[cur]) { gap_row[cur] = gap_row[tr]; gap_row[tr] = '-'; ++est_cl; ++gen_cl; }
    else stop = true;
    ++cur;
  }
  if (gen_cl >= t->gs) { t->ee = est_cl; t->ge = gen_cl; }
  else --z->n;
}

static bool is_ch(char c, char up) { return c == up || c == (char)(up + 32); }

/* clean_external_exons: the verdict for one end; *need_edit set when only a zero edit distance can still save it */
static bool external_ok_pre(const ef_fz *z, bool head, const char *g, bool *need_edit) {
  const ef_factor *x = head ? &z->f[0] : &z->f[z->n - 1];
  const int len = x->ge - x->gs + 1;
  *need_edit = false;
  if (len < 10) return false;
  if (len >= 20) return true;
  if (z->n < 2) return false;   /* the removed exon has no neighbour left in the list */
  if (head) {
    const ef_factor *nx = &z->f[1];
    if (!is_ch(g[x->ge + 1], 'G')) return false;
    if (!is_ch(g[x->ge + 2], 'T') && !is_ch(g[x->ge + 2], 'C')) return false;
    if (!is_ch(g[nx->gs - 2], 'A') || !is_ch(g[nx->gs - 1], 'G')) return false;
  } else {
    const ef_factor *pv = &z->f[z->n - 2];
    if (!is_ch(g[x->gs - 2], 'A') || !is_ch(g[x->gs - 1], 'G')) return false;
    if (!is_ch(g[pv->ge + 1], 'G')) return false;
    if (!is_ch(g[pv->ge + 2], 'T') && !is_ch(g[pv->ge + 2], 'C')) return false;
  }
  *need_edit = true;
  return true;
}

static int edit_job(const ef_factor *x, const char *g, const char *e) {
  return dp_push(PC_OP_EDIT, S_(e + x->es, x->ee - x->es + 1), S_(g + x->gs, x->ge - x->gs + 1), 0, 0, 0, 0);   /* symmetric */
}

void clean_external_exons(ef_task *T, ef_fz *z, const char *g, const char *e) {
  (void)T;
  if (z->n == 0) return;
  bool need;
  bool ok = external_ok_pre(z, true, g, &need);
  if (ok && need) { int h = edit_job(&z->f[0], g, e); dp_wait(); ok = dp_res(h)[1] == 0; }
  if (!ok) fz_remove(z, 0);
  if (z->n == 0) return;
  /* tail: judged with the head already settled; with a single exon left it has no neighbour */
  ok = external_ok_pre(z, false, g, &need);
  if (ok && need) { int h = edit_job(&z->f[z->n - 1], g, e); dp_wait(); ok = dp_res(h)[1] == 0; }
  if (!ok) --z->n;
}

static unsigned max_exon_errors(int len) {
  const double rate = len > 100 ? 0.030 : (len > 50 ? 0.035 : 0.040);
  return (unsigned)fmax(1.0, ceil((double)len * rate));
}

/* update_with_subfact_with_best_coverage: exons flagged in `bad` split the factorization; keep the run of good exons
 * with the widest EST span (first one wins ties) */
static void keep_best_run(ef_fz *z, const bool *bad) {
  int any = 0;
  for (int i = 0; i < z->n; ++i) any |= bad[i];
  if (!any) return;
  int best_l = -1, best_r = -1, best_cover = -1, start = 0;
  for (int i = 0; i <= z->n; ++i)
    if (i == z->n || bad[i]) {
      if (start < i) {
        const int cover = z->f[i - 1].ee - z->f[start].es + 1;
        if (cover > best_cover) { best_cover = cover; best_l = start; best_r = i - 1; }
      }
      start = i + 1;
    }
  if (best_l < 0) { z->n = 0; return; }
  memmove(z->f, z->f + best_l, sizeof(ef_factor) * (size_t)(best_r - best_l + 1));
  z->n = best_r - best_l + 1;
}

static void kband_jobs(ef_fz *z, const char *g, const char *e, int *handles) {
  for (int i = 0; i < z->n; ++i) {
    const ef_factor *x = &z->f[i];
    handles[i] = -1;
    if (x->gs <= x->ge)
      handles[i] = dp_push(PC_OP_KBAND, S_(e + x->es, x->ee - x->es + 1), S_(g + x->gs, x->ge - x->gs + 1),
                           (int)max_exon_errors(x->ge - x->gs + 1), 0, 0, 0);       /* symmetric: the genome side goes by reference */
  }
}

void clean_noisy_exons(ef_task *T, ef_fz *z, const char *g, const char *e) {
  if (z->n == 0) return;
  int *h = ar_alloc(&T->ar, sizeof(int) * (size_t)z->n);
  bool *bad = ar_alloc(&T->ar, (size_t)z->n);
  kband_jobs(z, g, e, h);
  dp_wait();
  for (int i = 0; i < z->n; ++i) bad[i] = h[i] < 0 || dp_res(h[i])[1] == 0;
  keep_best_run(z, bad);
}

/* ---- relaxed containment of factorizations (est-factorizations.c:1149-1256, list.c:320-487) --------------------- */
static int relaxed_cmp(const ef_factor *p1, const ef_factor *p2, int type, int diff, const ef_fz *l1) {
  if (p1->gs < p2->gs && p1->ge < p2->gs) return 1;
  if (p2->gs < p1->gs && p2->ge < p1->gs) return 1;
  const int max_unconf = 20;
  if (type == 0 && abs(p1->ge - p2->ge) <= diff && abs(p1->gs - p2->gs) <= diff) return 0;
  if (abs(type) == 2 && abs(p1->ge - p2->ge) <= diff) {
    if (type == 2) {
      if (p1->gs - p2->gs > max_unconf) return 1;
      if (p1->gs - p2->gs > 0) {
        int tot = 0;
        for (int i = 0; i < l1->n && l1->f[i].gs != p1->gs; ++i) tot += l1->f[i].ge - l1->f[i].gs + 1;
        if (abs(p1->gs - p2->gs - tot) < 10) return 1;
      }
    }
    return 0;
  }
  if (abs(type) == 1 && abs(p1->gs - p2->gs) <= diff) {
    if (type == 1) {
      if (p2->ge - p1->ge > max_unconf) return 1;
      if (p2->ge - p1->ge > 0) {
        int tot = 0;
        for (int i = l1->n - 1; i >= 0 && l1->f[i].gs != p1->gs; --i) tot += l1->f[i].ge - l1->f[i].gs + 1;
        if (abs(p2->ge - p1->ge - tot) < 20) return 1;
      }
    }
    return 0;
  }
  return 1;
}

static int relaxed_equal(const ef_fz *l1, const ef_fz *l2, int diff) {      /* relaxed_list_compare: -2 equal, 0 not */
  if (l1->n != l2->n || l1->n == 1) return 0;
  for (int i = 0; i < l1->n; ++i) {
    const int type = i == 0 ? -2 : (i + 1 == l1->n ? -1 : 0);
    if (relaxed_cmp(&l1->f[i], &l2->f[i], type, diff, l1) != 0) return 0;
  }
  return -2;
}

static int relaxed_contained(const ef_fz *l1, const ef_fz *l2, int diff) {  /* -1: l1 in l2, 1: l2 in l1, -2 equal, 0 */
  if (l1->n == l2->n) return relaxed_equal(l1, l2, diff);
  if (l1->n == 1 || l2->n == 1) return 0;
  const ef_fz *lg = l1->n > l2->n ? l1 : l2, *sh = l1->n > l2->n ? l2 : l1;
  int type = -2, il = 0, is = 0;
  unsigned count_long = 1;
  bool found = false;
  while (il < lg->n && !found) {
    if (relaxed_cmp(&lg->f[il], &sh->f[is], type, diff, lg) == 0) { found = true; ++is; } else ++count_long;
    ++il;
    if (type == -2) type = 2;
  }
  if (!found) return 0;
  unsigned count_factors = 1;
  bool stop = false;
  while (il < lg->n && is < sh->n && !stop) {
    type = count_factors + 1 == (unsigned)sh->n ? (count_long + 1 == (unsigned)lg->n ? -1 : 1) : 0;
    if (relaxed_cmp(&lg->f[il], &sh->f[is], type, diff, lg) == 0) { ++il; ++is; } else stop = true;
    ++count_factors; ++count_long;
  }
  if (stop) return 0;
  if (count_factors != (unsigned)sh->n) return 0;
  return l1->n >= l2->n ? 1 : -1;
}

bool add_if_not_exists(ef_task *T, ef_fz *z, ef_fzlist *L) {
  bool found = false;
  for (int k = 0; k < L->n && !found;) {
    ef_fz *c = L->v[k];
    int r = 0;
    if (c->n == z->n && c->n == 1) {
      const ef_factor *h1 = &z->f[0], *h2 = &c->f[0];
      if (h1->gs == h2->gs && h1->ge == h2->ge) r = -2;
      else if (h1->gs >= h2->gs && h1->ge <= h2->ge) r = -1;
      else if (h1->gs <= h2->gs && h1->ge >= h2->ge) r = 1;
    } else r = relaxed_contained(z, c, (int)T->cfg->max_site_difference);
    if (r < 0) {
      if (r == -2) {
        if (z->f[0].es < c->f[0].es) { c->f[0].es = z->f[0].es; c->f[0].gs = z->f[0].gs; }
        ef_factor *t1 = &z->f[z->n - 1], *t2 = &c->f[c->n - 1];
        if (t1->ee > t2->ee) { t2->ee = t1->ee; t2->ge = t1->ge; }
      }
      found = true;
    } else if (r == 1) { fzl_remove(L, k); continue; }
    ++k;
  }
  if (!found) fzl_push(T, L, z);
  return !found;
}

/* check_gap_errors: every gap left on P must fit inside its gap on T; total cost <= 20; then merge exons <= 3 nt apart */
static bool check_gap_errors(ef_task *T, ef_fz *z, const char *e, const char *g) {
  int *h = ar_alloc(&T->ar, sizeof(int) * (size_t)(z->n + 1));
  for (int i = 0; i + 1 < z->n; ++i) {
    const ef_factor *d = &z->f[i], *a = &z->f[i + 1];
    const size_t gp = (size_t)(a->es - d->ee - 1);
    h[i] = -1;
    if (gp > 0) {
      const size_t gt = (size_t)(a->gs - d->ge - 1);
      if (gp > gt) fprintf(stderr, "* FATAL ...the gap on P cannot be greater than the gap on T!\n");
      /* the reference hands refine_borders NUL-terminated copies: the byte after t reads as 0 there */
      h[i] = dp_push(PC_OP_BORDERS, S_(e + d->ee + 1, (int)gp), SZ_(g + d->ge + 1, (int)gt), (int)gp, 0, (int)gp, 0);
    }
  }
  dp_wait();
  unsigned tot = 0;
  for (int i = 0; i + 1 < z->n; ++i) {
    if (h[i] < 0) continue;
    const int32_t *r = dp_res(h[i]);
    if (!r[1]) return false;
    ef_factor *d = &z->f[i], *a = &z->f[i + 1];
    const int gt = a->gs - d->ge - 1;
    tot += (unsigned)r[5];
    d->ee += r[2]; a->es = d->ee + 1;
    d->ge += r[3]; a->gs -= gt - r[4];
  }
  if (tot > 20) return false;
  for (int i = 1; i < z->n;) {
    if (z->f[i].gs - z->f[i - 1].ge - 1 <= 3) { z->f[i - 1].ee = z->f[i].ee; z->f[i - 1].ge = z->f[i].ge; fz_remove(z, i); }
    else ++i;
  }
  return true;
}

/* detect-polya.c */
static void correct_tail(ef_fz *z, const char *g, int glen, const char *eo, int elen) {
  ef_factor *t = &z->f[z->n - 1];
  size_t i = (size_t)(t->ee + 1), j = (size_t)(t->ge + 1);
  while (i < (size_t)elen && j < (size_t)glen && g[j] == eo[i]) { ++i; ++j; }
  t->ee = (int)i - 1; t->ge = (int)j - 1;
}

static bool detect_polyA(const ef_fz *z, const char *g, const char *eo, int elen, bool *polyad) {
  const ef_factor *t = &z->f[z->n - 1];
  const char *cl = eo + t->ee + 1;
  const int cl_len = elen - t->ee - 1 > 0 ? (int)strlen(cl) : 0;
  int i = 0, matches = 0;
  bool stop = false;
  while (i < cl_len && !stop) {
    if (is_ch(cl[i], 'A')) { if (matches >= 8) stop = true; else { ++matches; ++i; } }
    else { if (matches >= 8) stop = true; else i = cl_len; }
  }
  *polyad = false;
  if (!stop) return false;
  for (i = MAX2(0, t->ge - 39); i <= t->ge && !*polyad; ++i)
    if (is_ch(g[i], 'A')) {
      char pas[7];
      strncpy(pas, g + i, 6); pas[6] = 0;
      *polyad = !strcmp(pas, "aataaa") || !strcmp(pas, "AATAAA") || !strcmp(pas, "attaaa") || !strcmp(pas, "ATTAAA");
    }
  i = MAX2(0, t->ge - 9); matches = 0;
  while (i <= t->ge + 10 && stop && g[i] != 0) {
    if (matches >= 6) stop = false;
    else { if (is_ch(g[i], 'A')) ++matches; else matches = 0; ++i; }
  }
  if (stop) {
    int count = 0;
    i = t->ge + 1;
    while (i <= t->ge + 10 && stop && g[i] != 0) {
      if (count >= 7) stop = false;
      else { if (is_ch(g[i], 'A')) ++count; ++i; }
    }
  }
  return stop;
}

/* ---- get_EST_factorizations ------------------------------------------------------------------------------------- */
static void candidate_phases(ef_task *T, const ef_seq *est, ef_fzlist *cand, ef_fzlist *out) {
  const char *g = T->gen->seq, *e = est->seq;
  const int n = cand->n, elen = est->len;
  bool *ok = ar_alloc(&T->ar, (size_t)n + 1);
  int *hh = ar_alloc(&T->ar, sizeof(int) * (size_t)(2 * n + 2));
  /* source/sink-only candidates and inverted exons */
  for (int c = 0; c < n; ++c) {
    ef_fz *z = cand->v[c];
    ok[c] = !(z->n <= 1 && (z->f[0].es < 0 || z->f[0].es >= elen)) && exon_bounds_ok(z);
  }
  /* handle_endpoints: head alignments of everybody, plus the tail alignment when it cannot depend on the head's */
  for (int c = 0; c < n; ++c) {
    hh[2 * c] = hh[2 * c + 1] = -1;
    if (!ok[c]) continue;
    ef_fz *z = cand->v[c];
    const ef_factor *h = &z->f[0], *t = &z->f[z->n - 1];
    hh[2 * c] = dp_push(PC_OP_ALIGN, S_(e + h->es, h->ee - h->es + 1), S_(g + h->gs, h->ge - h->gs + 1), 0, 0, 0, 0);
    if (z->n > 1) hh[2 * c + 1] = dp_push(PC_OP_ALIGN, S_(e + t->es, t->ee - t->es + 1), S_(g + t->gs, t->ge - t->gs + 1), 0, 0, 0, 0);
  }
  dp_wait();
  ef_aln *tails = ar_alloc(&T->ar, sizeof(ef_aln) * (size_t)(n + 1));
  for (int c = 0; c < n; ++c) {
    if (!ok[c]) continue;
    ef_fz *z = cand->v[c];
    const ef_factor h = z->f[0], t = z->f[z->n - 1];
    ef_aln A = aln_from_ops(T, dp_var(hh[2 * c]), dp_res(hh[2 * c])[2], e + h.es, g + h.gs);
    if (z->n > 1) tails[c] = aln_from_ops(T, dp_var(hh[2 * c + 1]), dp_res(hh[2 * c + 1])[2], e + t.es, g + t.gs);
    trim_head(z, &A);
  }
  for (int c = 0; c < n; ++c) {       /* single-exon candidates: the tail alignment sees the trimmed head */
    hh[2 * c] = -1;
    if (!ok[c]) continue;
    ef_fz *z = cand->v[c];
    if (z->n == 0) { ok[c] = false; continue; }
    if (hh[2 * c + 1] < 0) {
      const ef_factor *t = &z->f[z->n - 1];
      hh[2 * c] = dp_push(PC_OP_ALIGN, S_(e + t->es, t->ee - t->es + 1), S_(g + t->gs, t->ge - t->gs + 1), 0, 0, 0, 0);
    }
  }
  dp_wait();
  for (int c = 0; c < n; ++c) {
    if (!ok[c]) continue;
    ef_fz *z = cand->v[c];
    if (hh[2 * c] >= 0) {
      const ef_factor t = z->f[z->n - 1];
      tails[c] = aln_from_ops(T, dp_var(hh[2 * c]), dp_res(hh[2 * c])[2], e + t.es, g + t.gs);
    }
    trim_tail(z, &tails[c]);
    if (z->n == 0) ok[c] = false;
  }
  /* clean_external_exons: head verdicts, then tail verdicts (the tail is judged on the list without a removed head) */
  for (int side = 0; side < 2; ++side) {
    bool *need = ar_alloc(&T->ar, (size_t)n + 1);
    for (int c = 0; c < n; ++c) {
      hh[c] = -1;
      if (!ok[c]) continue;
      ef_fz *z = cand->v[c];
      bool nd;
      const bool pre = external_ok_pre(z, side == 0, g, &nd);
      need[c] = pre;
      if (pre && nd) hh[c] = edit_job(side == 0 ? &z->f[0] : &z->f[z->n - 1], g, e);
    }
    dp_wait();
    for (int c = 0; c < n; ++c) {
      if (!ok[c]) continue;
      ef_fz *z = cand->v[c];
      bool keep = need[c];
      if (keep && hh[c] >= 0) keep = dp_res(hh[c])[1] == 0;
      if (!keep) { if (side == 0) fz_remove(z, 0); else --z->n; }
      if (z->n == 0) ok[c] = false;
    }
  }
  /* clean_low_complexity_exons_2 (DUST on the genome and on the masked EST exon) */
  for (int c = 0; c < n; ++c) {
    if (!ok[c]) continue;
    ef_fz *z = cand->v[c];
    bool *bad = ar_alloc(&T->ar, (size_t)z->n + 1);
    for (int i = 0; i < z->n; ++i) {
      const ef_factor *x = &z->f[i];
      double gd = 0.0, ed = 0.0;
      if (x->gs <= x->ge) { gd = dust_score(g + x->gs, x->ge - x->gs + 1); ed = dust_score(e + x->es, x->ee - x->es + 1); }
      bad[i] = gd > T->cfg->complexity_threshold || ed > T->cfg->complexity_threshold;
    }
    keep_best_run(z, bad);
    if (z->n == 0) ok[c] = false;
  }
  /* clean_noisy_exons: one K-band job per exon of every surviving candidate */
  int **kh = ar_alloc(&T->ar, sizeof(int *) * (size_t)(n + 1));
  for (int c = 0; c < n; ++c) {
    if (!ok[c]) continue;
    kh[c] = ar_alloc(&T->ar, sizeof(int) * (size_t)cand->v[c]->n);
    kband_jobs(cand->v[c], g, e, kh[c]);
  }
  dp_wait();
  for (int c = 0; c < n; ++c) {
    if (!ok[c]) continue;
    ef_fz *z = cand->v[c];
    bool *bad = ar_alloc(&T->ar, (size_t)z->n + 1);
    for (int i = 0; i < z->n; ++i) bad[i] = kh[c][i] < 0 || dp_res(kh[c][i])[1] == 0;
    keep_best_run(z, bad);
    if (z->n == 0) ok[c] = false;
  }
  /* check_est_coverage (>= 35 %, float constant as in the reference) and the relaxed de-duplication, in order */
  for (int c = 0; c < n; ++c) {
    if (!ok[c]) continue;
    ef_fz *z = cand->v[c];
    const double cov = (double)(z->f[z->n - 1].ee - z->f[0].es + 1) / (double)(size_t)elen;
    if (!(cov >= 0.35f)) continue;
    add_if_not_exists(T, z, out);
  }
}

typedef struct par_fz { const ef_seq *est; ef_fzlist *L; } par_fz;
static void refine_introns_of(ef_task *T, int k, void *user) {
  const par_fz *P = user;
  ef_fz *z = P->L->v[k];
  if (z->n == 0) return;
  for (int i = 0; i + 1 < z->n; ++i) refine_intron(T, P->est, &z->f[i], &z->f[i + 1], i == 0);
  if (z->n > 1 && z->f[0].es == z->f[1].es) fz_remove(z, 0);
}

ef_fzlist *est_factorizations(ef_task *T, const ef_seq *est, ef_meg *M, bool *timed_out) {
  const ef_config *cfg = T->cfg;
  const char *g = T->gen->seq;
  const int elen = M->n - 2;
  *timed_out = false;
  ef_fzlist *L = ar_alloc(&T->ar, sizeof *L);
  unsigned tick = 0;
  for (int vi = 0; vi < M->nflat; ++vi) {
    ef_pairing *root = M->flat[vi];
    if (root->visited) continue;
    emblist *E = subtree_embeddings(T, root, &tick);
    if (!E) { *timed_out = true; return NULL; }
    ef_fzlist cand = {0};
    for (int x = 0; x < E->n; ++x) fzl_push(T, &cand, factorization_of(T, E->v[x]));
    ef_phase(EF_PH_CAND);
    candidate_phases(T, est, &cand, L);
    ef_phase(EF_PH_EMBED);
  }
  ef_phase(EF_PH_FILTER);
  /* FILTER 1: coverage on P relative to the best one */
  double *cov = ar_alloc(&T->ar, sizeof(double) * (size_t)(L->n + 1)), maxc = 0.0;
  for (int k = 0; k < L->n; ++k) {
    const ef_fz *z = L->v[k];
    if (z->n == 1 && (z->f[0].es < 0 || z->f[0].es >= elen)) { cov[k] = -1.0; continue; }
    const int cover = elen - (z->f[0].es + (elen - z->f[z->n - 1].ee - 1));
    cov[k] = (double)cover / (double)(unsigned)elen;
    if (maxc < cov[k]) maxc = cov[k];
  }
  {
    int w = 0;
    for (int k = 0; k < L->n; ++k) {
      const bool drop = cov[k] == -1.0 || maxc - cov[k] > cfg->max_coverage_diff || (maxc - cov[k]) * (double)est->len > 100;
      if (!drop) L->v[w++] = L->v[k];
    }
    L->n = w;
  }
  /* FILTER 3: total gap length on P */
  {
    int *gl = ar_alloc(&T->ar, sizeof(int) * (size_t)(L->n + 1)), mn = -1, w = 0;
    for (int k = 0; k < L->n; ++k) {
      const ef_fz *z = L->v[k];
      gl[k] = 0;
      for (int i = 1; i < z->n; ++i) gl[k] += z->f[i].es - z->f[i - 1].ee - 1;
      if (mn == -1 || mn > gl[k]) mn = gl[k];
    }
    for (int k = 0; k < L->n; ++k)
      if (!(cfg->max_gapLength_diff != -1 && gl[k] - mn > cfg->max_gapLength_diff)) L->v[w++] = L->v[k];
    L->n = w;
  }
  /* FILTER 4: errors inside the gaps (one borders job per gap; factorizations are independent) */
  {
    int w = 0;
    for (int k = 0; k < L->n; ++k) if (check_gap_errors(T, L->v[k], est->seq, g)) L->v[w++] = L->v[k];
    L->n = w;
  }
  if (cfg->max_number_of_factorizations != 0 && L->n > cfg->max_number_of_factorizations) L->n = 0;
  ef_phase(EF_PH_INTRON);
  /* splice-site refinement, intron by intron (the donor of intron k+1 is the acceptor refined by intron k) */
  { par_fz P = {est, L}; dp_parallel_for(T, L->n, refine_introns_of, &P); }      /* factorizations are independent of each other */
  for (int k = 0; k < L->n; ++k) {
    ef_fz *z = L->v[k];
    correct_tail(z, g, T->gen->len, est->orig, est->len);
    z->polya = detect_polyA(z, g, est->orig, est->len, &z->polyad);
  }
  return L;
}
