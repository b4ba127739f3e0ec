/* meg.c — the Maximal Embedding Graph of one EST: vertices from the device seeding kernel, then edges,
 * simplification, transitive reduction, short-edge compaction and the "too complex" retry loop.
 *
 * Behaviour follows reference src/compute-est-fact.c:90-152 (build_meg), src/max-emb-graph.c:382-672 (edges) and
 * src/meg-simplification.c:52-632; the order rules that the output depends on are the ones validated in
 * SURVEY.md Appendix D.  Lists are pointer arrays here; where the reference walks a linked list while appending to
 * it, its iterator has already cached the successor of the current node, so an element appended while the LAST
 * element is being processed is not visited in that sweep — walk_next() below reproduces that.
 */
#include "ef.h"

void pl_push(ef_task *T, ef_plist *l, ef_pairing *x) {
  if (l->n == l->cap) {
    int nc = l->cap ? l->cap * 2 : 4;
    ef_pairing **nv = ar_alloc_raw(&T->ar, sizeof(ef_pairing *) * (size_t)nc);
    if (l->n) memcpy(nv, l->v, sizeof(ef_pairing *) * (size_t)l->n);
    l->v = nv; l->cap = nc;
  }
  l->v[l->n++] = x;
}

static void pl_remove_at(ef_plist *l, int k) {
  memmove(l->v + k, l->v + k + 1, sizeof(ef_pairing *) * (size_t)(l->n - k - 1));
  --l->n;
}

bool pl_remove_first(ef_plist *l, ef_pairing *x) {
  for (int k = 0; k < l->n; ++k) if (l->v[k] == x) { pl_remove_at(l, k); return true; }
  return false;
}

static ef_pairing *new_pairing(ef_task *T, int p, int t, int l) {
  ef_pairing *q = ar_alloc(&T->ar, sizeof *q);
  q->p = p; q->t = t; q->l = l;
  return q;
}

void meg_stats(const ef_meg *M, size_t *pairings, size_t *edges) {
  size_t np = 0, ne = 0;
  for (int i = 0; i < M->n; ++i)
    for (int k = 0; k < M->V[i].n; ++k) { ++np; ne += (size_t)M->V[i].v[k]->adjs.n; }
  *pairings = np; *edges = ne;
}

/* ---- edges (max-emb-graph.c:393-672) -------------------------------------------------------------------- */
static bool edge_ok(const ef_pairing *I, const ef_pairing *J, int l, int fl, const ef_config *c) {
  if (J->p <= I->p || J->t <= I->t) return false;
  const bool simple_t = I->t + I->l <= J->t && (c->max_intron_length == 0 || J->t <= I->t + I->l + c->max_intron_length);
  const bool over_t = I->t + 2 * l <= J->t + J->l && J->t < I->t + I->l && J->p + I->t - I->p - J->t <= fl;
  if (I->p + I->l <= J->p && J->p <= I->p + I->l + fl) {          /* simple sequence on P */
    if (simple_t) return true;
    if (over_t) return !(I->l >= 5 * l && (double)(I->t + I->l - J->t) > 0.4 * (double)I->l);
    return false;                                                   /* does not fall through to the overlap-on-P case */
  }
  if (I->p + 2 * l <= J->p + J->l && J->p < I->p + I->l) return simple_t || over_t;   /* overlap on P */
  return false;
}

static bool disjoint(const ef_pairing *a, const ef_pairing *b) {
  return (a->p + a->l <= b->p || b->p + b->l <= a->p) && (a->t + a->l <= b->t || b->t + b->l <= a->t);
}

static void build_edges(ef_task *T, ef_meg *M, int l) {
  const ef_config *c = T->cfg;
  const int n = M->n, fl = 2 * l + 1, plen = n - 2;
  /* The reference scans V[0 .. ub) for every pairing I (max-emb-graph.c:540-550); most of those lists are empty, and
   * edge_ok needs J.p > I.p, i.e. j > i.  Same pairs in the same order from a flat copy of the non-empty lists. */
  int total = 0;
  for (int j = 0; j < n; ++j) total += M->V[j].n;
  ef_pairing **flat = ar_alloc_raw(&T->ar, sizeof(ef_pairing *) * (size_t)(total + 1));
  int *first_at = ar_alloc_raw(&T->ar, sizeof(int) * (size_t)(n + 1));      /* flat index of the first pairing in V[j ..] */
  {
    int k = 0;
    for (int j = 0; j < n; ++j) { first_at[j] = k; for (int b = 0; b < M->V[j].n; ++b) flat[k++] = M->V[j].v[b]; }
    first_at[n] = k;
  }
  /* two sweeps: count, size every adjacency list once (room for the source / sink edge added below), fill */
  for (int sweep = 0; sweep < 2; ++sweep) {
    for (int i = 1; i < n - 1; ++i)
      for (int a = 0; a < M->V[i].n; ++a) {
        ef_pairing *I = M->V[i].v[a];
        const int ub = MIN2(I->p + I->l + fl + 1, n - l);
        if (ub <= i + 1) continue;
        const int k1 = first_at[ub];
        for (int k = first_at[i + 1]; k < k1; ++k) {
          ef_pairing *J = flat[k];
          if (!edge_ok(I, J, l, fl, c)) continue;
          if (sweep == 0) { ++I->adjs.cap; ++J->incs.cap; }
          else { I->adjs.v[I->adjs.n++] = J; J->incs.v[J->incs.n++] = I; }
        }
      }
    if (sweep == 0)
      for (int k = first_at[1]; k < first_at[n - 1]; ++k) {
        ef_pairing *q = flat[k];
        q->adjs.cap += 1; q->incs.cap += 1;
        q->adjs.v = ar_alloc_raw(&T->ar, sizeof(ef_pairing *) * (size_t)q->adjs.cap);
        q->incs.v = ar_alloc_raw(&T->ar, sizeof(ef_pairing *) * (size_t)q->incs.cap);
      }
  }
  ef_pairing *source = M->V[0].v[0], *sink = M->V[n - 1].v[0];
  const int max_p = (int)((double)plen * c->max_prefix_discarded_rate);
  for (int i = 1; i <= max_p; ++i)
    for (int a = 0; a < M->V[i].n; ++a) {
      ef_pairing *I = M->V[i].v[a];
      bool ok = true;
      for (int k = 0; ok && k < I->incs.n; ++k) {
        const ef_pairing *q = I->incs.v[k];
        ok = !disjoint(q, I);
        ok = ok && (q->p + l > I->p || q->t + l > I->t);
      }
      if (ok) { pl_push(T, &source->adjs, I); pl_push(T, &I->incs, source); }
    }
  const int min_p = (int)((double)plen * (1.0 - c->max_suffix_discarded_rate));
  for (int i = 1; i <= plen; ++i)
    for (int a = 0; a < M->V[i].n; ++a) {
      ef_pairing *I = M->V[i].v[a];
      if (I->p + I->l < min_p) continue;
      bool ok = true;
      for (int k = 0; ok && k < I->adjs.n; ++k) {
        const ef_pairing *q = I->adjs.v[k];
        ok = !disjoint(q, I);
        ok = ok && (I->p + I->l + l > q->p + q->l || I->t + I->l + l > q->t + q->l);
      }
      if (ok) { pl_push(T, &sink->incs, I); pl_push(T, &I->adjs, sink); }
    }
}

/* ---- simplification (meg-simplification.c:142-258) ----------------------------------------------------------- */
static void prune_dead_ends(ef_meg *M) {
  bool removed;
  do {
    removed = false;
    for (int i = 1; i < M->n - 1; ++i)
      for (int a = 0; a < M->V[i].n;) {
        ef_pairing *I = M->V[i].v[a];
        if (I->adjs.n == 0 || I->incs.n == 0) {
          removed = true;
          for (int k = 0; k < I->adjs.n; ++k) pl_remove_first(&I->adjs.v[k]->incs, I);
          for (int k = 0; k < I->incs.n; ++k) pl_remove_first(&I->incs.v[k]->adjs, I);
          I->dead = true;
          pl_remove_at(&M->V[i], a);
        } else ++a;
      }
  } while (removed);
}

static void remove_useless_edges(ef_meg *M, int l, const ef_config *c) {
  const int gl = 2 * l + 3;
  for (int i = 1; i < M->n; ++i)
    for (int a = 0; a < M->V[i].n; ++a) {
      ef_pairing *p = M->V[i].v[a];
      for (int k = 0; k < p->adjs.n;) {
        ef_pairing *q = p->adjs.v[k];
        if (q->t != SINK_START) {
          const int gap = MAX2(q->t - q->p - p->t + p->p, 0);
          if (gap > gl && gap < c->min_intron_length) { pl_remove_at(&p->adjs, k); pl_remove_first(&q->incs, p); continue; }
        }
        ++k;
      }
    }
}

/* ---- transitive reduction (meg-simplification.c:333-632) ----------------------------------------------------- */
static int cmp_by_id(const void *a, const void *b) { return (*(ef_pairing *const *)a)->id - (*(ef_pairing *const *)b)->id; }

static void transitive_reduction(ef_task *T, ef_meg *M) {
  size_t nv, ne;
  meg_stats(M, &nv, &ne);
  ef_pairing **G = ar_alloc(&T->ar, sizeof(*G) * (nv + 1));
  int n = 0;
  for (int i = 0; i < M->n; ++i) for (int k = 0; k < M->V[i].n; ++k) { G[n] = M->V[i].v[k]; G[n]->id = n; ++n; }
  /* iterative DFS: the stack is the reference's int list used LIFO from its tail */
  int *color = ar_alloc(&T->ar, sizeof(int) * (size_t)(n + 1)), *ids = ar_alloc(&T->ar, sizeof(int) * (size_t)(n + 1));
  size_t scap = ne + (size_t)n + 16, sp = 0;
  int *S = ar_alloc(&T->ar, sizeof(int) * scap);
  bool acyclic = true;
  for (int i = 0; i < n; ++i) if (G[i]->incs.n == 0) S[sp++] = i;
  if (sp == 0) acyclic = false;
  int next_id = n;
  do {
    while (sp) {
      const int v = S[--sp];
      if (color[v] == 0) {
        color[v] = 1;
        S[sp++] = v;
        for (int k = 0; k < G[v]->adjs.n; ++k) {
          const int a = G[v]->adjs.v[k]->id;
          if (color[a] == 0) {
            if (sp == scap) { int *nS = ar_alloc(&T->ar, sizeof(int) * scap * 2); memcpy(nS, S, sizeof(int) * sp); S = nS; scap *= 2; }
            S[sp++] = a;
          } else if (color[a] == 1) acyclic = false;
        }
      } else if (color[v] == 1) { color[v] = 2; ids[v] = --next_id; }
    }
    for (int i = 0; i < n && sp == 0; ++i) if (color[i] == 0) { acyclic = false; S[sp++] = i; }
  } while (sp);
  if (!acyclic) { fprintf(stderr, "* FATAL The graph is cyclic. Transitive reduction not possible! Terminating.\n"); exit(1); }
  ef_pairing **H = ar_alloc(&T->ar, sizeof(*H) * (size_t)(n + 1));
  for (int i = 0; i < n; ++i) { G[i]->id = ids[i]; H[ids[i]] = G[i]; }
  for (int i = 0; i < n; ++i) {
    if (H[i]->adjs.n > 1) qsort(H[i]->adjs.v, (size_t)H[i]->adjs.n, sizeof(ef_pairing *), cmp_by_id);
    if (H[i]->incs.n > 1) qsort(H[i]->incs.v, (size_t)H[i]->incs.n, sizeof(ef_pairing *), cmp_by_id);
  }
  ef_plist *star = ar_alloc(&T->ar, sizeof(ef_plist) * (size_t)(n + 1));
  ef_plist *red = ar_alloc(&T->ar, sizeof(ef_plist) * (size_t)(n + 1));
  ef_plist *redinc = ar_alloc(&T->ar, sizeof(ef_plist) * (size_t)(n + 1));
  unsigned char *reach = ar_alloc(&T->ar, (size_t)n + 1);
  for (int i = n - 1; i >= 0; --i) {
    ef_pairing *v = H[i];
    memset(reach, 0, (size_t)n);
    reach[i] = 1;
    pl_push(T, &star[i], v);
    for (int k = 0; k < v->adjs.n; ++k) {
      ef_pairing *w = v->adjs.v[k];
      const bool early_end = w->p + w->l < v->p + v->l || w->t + w->l < v->t + v->l;
      if (!reach[w->id] || w->p < v->p || w->t < v->t || early_end) {
        pl_push(T, &red[i], w);
        pl_push(T, &redinc[w->id], v);
        if (!early_end)
          for (int q = 0; q < star[w->id].n; ++q) {
            ef_pairing *wa = star[w->id].v[q];
            if (!reach[wa->id] && v->t <= wa->t && v->p <= wa->p && v->t + v->l <= wa->t + wa->l && v->p + v->l <= wa->p + wa->l) {
              reach[wa->id] = 1;
              pl_push(T, &star[i], wa);
            }
          }
      }
    }
  }
  for (int i = 0; i < n; ++i) { H[i]->adjs = red[i]; H[i]->incs = redinc[i]; }
}

/* ---- short-edge compaction (meg-simplification.c:258-312) ---------------------------------------------------- */
static void compact_short_edges(ef_task *T, ef_meg *M) {
  bool changed;
  do {
    changed = false;
    for (int i = 1; i < M->n; ++i) {
      ef_plist *Vi = &M->V[i];
      for (int a = 0; a < Vi->n; ++a) {
        const bool was_last = a == Vi->n - 1;          /* successor cached when the walk reached this element */
        ef_pairing *p = Vi->v[a];
        for (int k = 0; k < p->adjs.n;) {
          ef_pairing *q = p->adjs.v[k];
          if (q->t != SINK_START && q->t + q->l - p->t == q->p + q->l - p->p && q->t >= p->t + p->l && q->t - p->t - p->l <= 3) {
            changed = true;
            pl_remove_at(&p->adjs, k);
            pl_remove_first(&q->incs, p);
            ef_pairing *nv = new_pairing(T, p->p, p->t, q->p + q->l - p->p);
            for (int x = 0; x < q->adjs.n; ++x) { pl_push(T, &nv->adjs, q->adjs.v[x]); pl_push(T, &q->adjs.v[x]->incs, nv); }
            for (int y = 0; y < p->incs.n; ++y) { pl_push(T, &nv->incs, p->incs.v[y]); pl_push(T, &p->incs.v[y]->adjs, nv); }
            pl_push(T, Vi, nv);
            continue;
          }
          ++k;
        }
        if (was_last) break;
      }
    }
    prune_dead_ends(M);
  } while (changed);
}

static bool too_complex(const ef_meg *M, int l, const ef_config *c) {     /* is_too_complex, :89-140 */
  int min_len = 0;
  size_t freq = 0, np = 0, ne = 0;
  const size_t est_len = (size_t)M->n - 2;
  for (int i = 0; i < M->n; ++i)
    for (int k = 0; k < M->V[i].n; ++k) {
      const ef_pairing *p = M->V[i].v[k];
      ++np;
      if (min_len == 0 || p->l < min_len) { min_len = p->l; freq = 1; } else if (p->l == min_len) ++freq;
      ne += (size_t)p->adjs.n;
    }
  if (np < 5 || ne < 4) return false;
  if (c->max_pairings_in_MEG != 0 && np > c->max_pairings_in_MEG && (double)freq > c->max_freq_shortest_pairing * (double)np) return true;
  return ne > 5 * np || np > (2 * est_len) / (size_t)l || (np > est_len / (size_t)l && np >= 50);
}

/* ---- vertex set from the device ------------------------------------------------------------------------------ */
static ef_meg *vertex_set(ef_task *T, const ef_seq *est, int mfl) {
  const int ph_ = ef_phase(EF_PH_SEED);
  int cap = 256, h;      /* triples; the kernel reports the needed count when this is too small */
  const int32_t *r;
  for (;;) {
    h = dp_push(PC_OP_SEED, S_(est->seq, est->len), S_(NULL, 0), mfl, 0, 0, cap);
    dp_wait();
    r = dp_res(h);
    if (r[0] != PC_E_OUTCAP) break;
    cap = r[1] + 16;
  }
  const int32_t *tri = (const int32_t *)dp_var(h);
  ef_meg *M = ar_alloc(&T->ar, sizeof *M);
  M->n = est->len + 2;
  M->V = ar_alloc(&T->ar, sizeof(ef_plist) * (size_t)M->n);
  pl_push(T, &M->V[0], new_pairing(T, SRC_START, SRC_START, SENTINEL_LEN));
  for (int k = 0; k < r[1]; ++k) pl_push(T, &M->V[tri[3 * k] + 1], new_pairing(T, tri[3 * k], tri[3 * k + 1], tri[3 * k + 2]));
  pl_push(T, &M->V[M->n - 1], new_pairing(T, SINK_START, SINK_START, SENTINEL_LEN));
  ef_phase(ph_);
  return M;
}

ef_meg *meg_build(ef_task *T, const ef_seq *est, unsigned *inc) {
  const ef_config *c = T->cfg;
  for (;;) {
    const int l = (int)(c->min_factor_len + *inc);
    ef_meg *M = vertex_set(T, est, l);
    build_edges(T, M, l);
    remove_useless_edges(M, l, c);
    prune_dead_ends(M);
    if (c->trans_red) transitive_reduction(T, M);
    size_t np, ne;
    meg_stats(M, &np, &ne);
    bool complex = ne > 1000 || np > 2000;
    if (!complex && c->short_edge_comp) compact_short_edges(T, M);
    complex = complex || too_complex(M, l, c);
    if (complex && (size_t)c->min_factor_len + *inc + 1 + 2 < (size_t)M->n) { ++*inc; continue; }
    return M;
  }
}

/* ---- megs.txt / meg-edges.txt (io-meg.c:147-190, max-emb-graph.c:677-707) ------------------------------------ */
void meg_write(ef_buf *b, ef_meg *M) {
  int id = 0;
  for (int i = 0; i < M->n; ++i)
    for (int k = 0; k < M->V[i].n; ++k) {
      ef_pairing *p = M->V[i].v[k];
      { const int v[3] = {p->p, p->t, p->l}; buf_ints(b, "(", v, 3, ',', ")\n"); }
      p->id = id++;
    }
  buf_printf(b, "#adj#\n");
  for (int i = 0; i < M->n; ++i)
    for (int k = 0; k < M->V[i].n; ++k) {
      const ef_pairing *p = M->V[i].v[k];
      for (int a = 0; a < p->adjs.n; ++a) { const int v[2] = {p->id, p->adjs.v[a]->id}; buf_ints(b, "", v, 2, '-', "\n"); }
    }
}

void meg_write_edges(ef_buf *b, const ef_meg *M) {
  for (int i = 0; i < M->n; ++i)
    for (int k = 0; k < M->V[i].n; ++k) {
      const ef_pairing *p = M->V[i].v[k];
      if (p->p == SRC_START || p->p == SINK_START) continue;
      for (int a = 0; a < p->adjs.n; ++a) {
        const ef_pairing *q = p->adjs.v[a];
        if (q->p == SINK_START) continue;
        const int v[9] = {p->t + p->l, q->t, p->p + p->l, q->p, q->t - p->t - p->l, q->p - p->p - p->l,
                          (q->t - p->t) - (q->p - p->p), p->l, q->l};
        buf_ints(b, "", v, 9, ' ', v[6] >= 50 ? " intronic\n" : "\n");
      }
    }
}
