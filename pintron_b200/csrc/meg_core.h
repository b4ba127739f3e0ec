/* meg_core.h — the Maximal Embedding Graph of one EST, from its vertex set to the finished graph: edges, simplification,
 * transitive reduction, short-edge compaction, the "too complex" test (PC_OP_SEED with p1 = PC_SEED_BUILD_MEG).
 *
 * Behaviour = reference src/compute-est-fact.c:90-152 (build_meg), src/max-emb-graph.c:382-672 (build_edge_set),
 * src/meg-simplification.c:52-632; the order rules the output bytes depend on are those of SURVEY.md Appendix D.
 * The reference keeps intrusive linked lists of heap nodes per EST position; here the graph is index-based and lives in
 * ONE flat int region (the warp's scratch slot on the device): a vertex table, the vertex ORDER (the concatenation of the
 * reference's per-position lists V[0], V[1], ... — sorted by p, list order inside one p), and a bump pool for the ordered
 * adjacency lists.  Where the reference walks a list while appending to it, its iterator has already cached the successor
 * of the current node, so an element appended while the LAST element is being processed is not visited in that sweep.
 *
 * The same source is compiled for the device (k_seed.cu: one lane walks the sequential, order-dependent part) and for the
 * host-side test backend (tests/cpu_backend): plain C subset, no recursion, no allocation, no library calls. */
#ifndef MEG_CORE_H
#define MEG_CORE_H
#include <stdint.h>

#ifdef __CUDACC__
#define MG_FN __device__ static
#else
#define MG_FN static
#endif

#define MG_SRC_START INT32_MIN
#define MG_SINK_START (INT32_MAX - 200)
#define MG_SENTINEL_LEN 200
enum { MG_OK = 0, MG_E_SCRATCH = 1, MG_E_CYCLIC = 2 };

typedef struct mg_list { int off, n, cap; } mg_list;             /* ordered list of vertex numbers inside the int pool */
typedef struct mg_vtx { int p, t, l, id; mg_list adjs, incs; } mg_vtx;
#define MG_VTX_INTS 10
typedef struct mg_graph {
  mg_vtx *vx; int nvx, vx_cap;
  int *order; int norder;            /* order[0] = source, order[norder-1] = sink */
  int *ip; int ip_used, ip_cap;
  int err;
} mg_graph;

MG_FN int mg_min(int a, int b) { return a < b ? a : b; }
MG_FN int mg_max(int a, int b) { return a > b ? a : b; }

MG_FN int mg_alloc(mg_graph *g, int n) {
  if (g->err) return 0;
  if (n < 0 || n > g->ip_cap - g->ip_used) { g->err = MG_E_SCRATCH; return 0; }
  const int o = g->ip_used;
  g->ip_used += n;
  return o;
}
MG_FN int mg_alloc_zero(mg_graph *g, int n) {
  const int o = mg_alloc(g, n);
  if (!g->err) for (int i = 0; i < n; ++i) g->ip[o + i] = 0;
  return o;
}

MG_FN void mg_push(mg_graph *g, mg_list *l, int x) {
  if (g->err) return;
  if (l->n == l->cap) {
    const int nc = l->cap ? l->cap * 2 : 4;
    const int o = mg_alloc(g, nc);
    if (g->err) return;
    for (int i = 0; i < l->n; ++i) g->ip[o + i] = g->ip[l->off + i];
    l->off = o; l->cap = nc;
  }
  g->ip[l->off + l->n++] = x;
}
MG_FN void mg_remove_at(mg_graph *g, mg_list *l, int k) {
  int *v = g->ip + l->off;
  for (int i = k; i + 1 < l->n; ++i) v[i] = v[i + 1];
  --l->n;
}
MG_FN void mg_remove_first(mg_graph *g, mg_list *l, int x) {
  const int *v = g->ip + l->off;
  for (int k = 0; k < l->n; ++k) if (v[k] == x) { mg_remove_at(g, l, k); return; }
}

MG_FN int mg_new_vtx(mg_graph *g, int p, int t, int l) {
  if (g->err) return 0;
  if (g->nvx >= g->vx_cap) { g->err = MG_E_SCRATCH; return 0; }
  mg_vtx *v = &g->vx[g->nvx];
  v->p = p; v->t = t; v->l = l; v->id = 0;
  v->adjs.off = v->adjs.n = v->adjs.cap = 0;
  v->incs.off = v->incs.n = v->incs.cap = 0;
  return g->nvx++;
}

/* Carves the graph out of `mem` (nints ints) and loads the vertex set: source, the ntri (p, t, l) triples (ascending p,
 * as build_vertex_set emits them), sink. */
MG_FN void mg_init(mg_graph *g, int *mem, long long nints, const int *tri, int ntri) {
  g->err = MG_OK; g->nvx = 0; g->norder = 0; g->ip_used = 0;
  long long vcap = (long long)ntri + 2 + 32 + nints / 64;
  if (vcap > 0x3fffffff) vcap = 0x3fffffff;
  const long long vints = vcap * (MG_VTX_INTS + 1);
  g->vx_cap = 0; g->ip_cap = 0; g->vx = (mg_vtx *)mem; g->order = mem; g->ip = mem;
  if (vints + 64 > nints) { g->err = MG_E_SCRATCH; return; }
  g->vx_cap = (int)vcap;
  g->order = mem + vcap * MG_VTX_INTS;
  g->ip = mem + vints;
  const long long rest = nints - vints;
  g->ip_cap = rest > 0x7fffffff ? 0x7fffffff : (int)rest;
  g->order[g->norder++] = mg_new_vtx(g, MG_SRC_START, MG_SRC_START, MG_SENTINEL_LEN);
  for (int k = 0; k < ntri; ++k) g->order[g->norder++] = mg_new_vtx(g, tri[3 * k], tri[3 * k + 1], tri[3 * k + 2]);
  g->order[g->norder++] = mg_new_vtx(g, MG_SINK_START, MG_SINK_START, MG_SENTINEL_LEN);
}

/* ---- edges (max-emb-graph.c:393-672) ------------------------------------------------------------------------------ */
MG_FN int mg_edge_ok(const mg_vtx *I, const mg_vtx *J, int l, int fl, const pc_meg_cfg *c) {
  if (J->p <= I->p || J->t <= I->t) return 0;
  const int simple_t = I->t + I->l <= J->t && (c->max_intron_length == 0 || J->t <= I->t + I->l + c->max_intron_length);
  const int over_t = I->t + 2 * l <= J->t + J->l && J->t < I->t + I->l && J->p + I->t - I->p - J->t <= fl;
  if (I->p + I->l <= J->p && J->p <= I->p + I->l + fl) {          /* simple sequence on P */
    if (simple_t) return 1;
    if (over_t) return !(I->l >= 5 * l && (double)(I->t + I->l - J->t) > 0.4 * (double)I->l);
    return 0;                                                       /* does not fall through to the overlap-on-P case */
  }
  if (I->p + 2 * l <= J->p + J->l && J->p < I->p + I->l) return simple_t || over_t;   /* overlap on P */
  return 0;
}
MG_FN int mg_disjoint(const mg_vtx *a, const mg_vtx *b) {
  return (a->p + a->l <= b->p || b->p + b->l <= a->p) && (a->t + a->l <= b->t || b->t + b->l <= a->t);
}

/* n = |P| + 2 (the reference's number of list slots), l = pairing length in force */
MG_FN void mg_build_edges(mg_graph *g, int n, int l, const pc_meg_cfg *c) {
  const int fl = 2 * l + 1, plen = n - 2, N = g->norder;
  /* The reference scans the slots (I.p + 1, ub) for every pairing I (max-emb-graph.c:540-550): the vertices after I's own
   * position group whose slot p + 1 is below ub, in order.  Two sweeps: count, size every list once (room for the source /
   * sink edge added below), fill. */
  for (int sweep = 0; sweep < 2 && !g->err; ++sweep) {
    for (int x = 1; x < N - 1; ++x) {
      const int ii = g->order[x];
      mg_vtx *I = &g->vx[ii];
      const int ub = mg_min(I->p + I->l + fl + 1, n - l);
      int y = x + 1;
      while (y < N - 1 && g->vx[g->order[y]].p == I->p) ++y;
      for (; y < N - 1; ++y) {
        const int jj = g->order[y];
        mg_vtx *J = &g->vx[jj];
        if (J->p + 1 >= ub) break;
        if (!mg_edge_ok(I, J, l, fl, c)) continue;
        if (sweep == 0) { ++I->adjs.cap; ++J->incs.cap; }
        else { g->ip[I->adjs.off + I->adjs.n++] = jj; g->ip[J->incs.off + J->incs.n++] = ii; }
      }
    }
    if (sweep == 0)
      for (int x = 1; x < N - 1; ++x) {
        mg_vtx *q = &g->vx[g->order[x]];
        q->adjs.cap += 1; q->incs.cap += 1;
        q->adjs.off = mg_alloc(g, q->adjs.cap);
        q->incs.off = mg_alloc(g, q->incs.cap);
      }
  }
  if (g->err) return;
  const int source = g->order[0], sink = g->order[N - 1];
  const int max_p = (int)((double)plen * c->max_prefix_discarded_rate);
  for (int x = 1; x < N - 1; ++x) {
    const int ii = g->order[x];
    mg_vtx *I = &g->vx[ii];
    if (I->p + 1 > max_p) break;                                    /* slots 1 .. max_p */
    int ok = 1;
    for (int k = 0; ok && k < I->incs.n; ++k) {
      const mg_vtx *q = &g->vx[g->ip[I->incs.off + k]];
      ok = !mg_disjoint(q, I);
      ok = ok && (q->p + l > I->p || q->t + l > I->t);
    }
    if (ok) { mg_push(g, &g->vx[source].adjs, ii); mg_push(g, &I->incs, source); }
  }
  const int min_p = (int)((double)plen * (1.0 - c->max_suffix_discarded_rate));
  for (int x = 1; x < N - 1; ++x) {
    const int ii = g->order[x];
    mg_vtx *I = &g->vx[ii];
    if (I->p + I->l < min_p) continue;
    int ok = 1;
    for (int k = 0; ok && k < I->adjs.n; ++k) {
      const mg_vtx *q = &g->vx[g->ip[I->adjs.off + k]];
      ok = !mg_disjoint(q, I);
      ok = ok && (I->p + I->l + l > q->p + q->l || I->t + I->l + l > q->t + q->l);
    }
    if (ok) { mg_push(g, &g->vx[sink].incs, ii); mg_push(g, &I->adjs, sink); }
  }
}

/* ---- simplification (meg-simplification.c:142-258) ---------------------------------------------------------------- */
MG_FN void mg_prune_dead_ends(mg_graph *g) {
  int removed;
  do {
    removed = 0;
    for (int x = 1; x < g->norder - 1;) {
      const int ii = g->order[x];
      mg_vtx *I = &g->vx[ii];
      if (I->adjs.n == 0 || I->incs.n == 0) {
        removed = 1;
        for (int k = 0; k < I->adjs.n; ++k) mg_remove_first(g, &g->vx[g->ip[I->adjs.off + k]].incs, ii);
        for (int k = 0; k < I->incs.n; ++k) mg_remove_first(g, &g->vx[g->ip[I->incs.off + k]].adjs, ii);
        for (int y = x; y + 1 < g->norder; ++y) g->order[y] = g->order[y + 1];
        --g->norder;
      } else ++x;
    }
  } while (removed);
}

MG_FN void mg_remove_useless_edges(mg_graph *g, int l, const pc_meg_cfg *c) {
  const int gl = 2 * l + 3;
  for (int x = 1; x < g->norder; ++x) {
    const int pi = g->order[x];
    mg_vtx *p = &g->vx[pi];
    for (int k = 0; k < p->adjs.n;) {
      const int qi = g->ip[p->adjs.off + k];
      mg_vtx *q = &g->vx[qi];
      if (q->t != MG_SINK_START) {
        const int gap = mg_max(q->t - q->p - p->t + p->p, 0);
        if (gap > gl && gap < c->min_intron_length) { mg_remove_at(g, &p->adjs, k); mg_remove_first(g, &q->incs, pi); continue; }
      }
      ++k;
    }
  }
}

MG_FN void mg_stats(const mg_graph *g, long long *np, long long *ne) {
  long long e = 0;
  for (int x = 0; x < g->norder; ++x) e += g->vx[g->order[x]].adjs.n;
  *np = g->norder; *ne = e;
}

/* ---- transitive reduction (meg-simplification.c:333-632) ---------------------------------------------------------- */
MG_FN void mg_sort_by_id(mg_graph *g, mg_list *l) {                 /* ids are distinct: any sort gives the reference's order */
  int *v = g->ip + l->off;
  for (int i = 1; i < l->n; ++i) {
    const int x = v[i], key = g->vx[x].id;
    int j = i - 1;
    while (j >= 0 && g->vx[v[j]].id > key) { v[j + 1] = v[j]; --j; }
    v[j + 1] = x;
  }
}

MG_FN void mg_transitive_reduction(mg_graph *g) {
  long long nv, ne;
  mg_stats(g, &nv, &ne);
  const int n = g->norder;
  if (ne + n + 16 > 0x3fffffff) { g->err = MG_E_SCRATCH; return; }
  const int scap = (int)ne + n + 16;
  const int oG = mg_alloc(g, n + 1), oC = mg_alloc_zero(g, n + 1), oI = mg_alloc_zero(g, n + 1), oS = mg_alloc(g, scap), oH = mg_alloc(g, n + 1);
  const int oStar = mg_alloc_zero(g, 3 * (n + 1)), oRed = mg_alloc_zero(g, 3 * (n + 1)), oRinc = mg_alloc_zero(g, 3 * (n + 1)), oReach = mg_alloc(g, n + 1);
  if (g->err) return;
  int *G = g->ip + oG, *color = g->ip + oC, *ids = g->ip + oI, *S = g->ip + oS, *H = g->ip + oH, *reach = g->ip + oReach;
  for (int i = 0; i < n; ++i) { G[i] = g->order[i]; g->vx[G[i]].id = i; }
  /* iterative DFS: the stack is the reference's int list used LIFO from its tail */
  int sp = 0, acyclic = 1;
  for (int i = 0; i < n; ++i) if (g->vx[G[i]].incs.n == 0) S[sp++] = i;
  if (sp == 0) acyclic = 0;
  int next_id = n;
  do {
    while (sp) {
      const int v = S[--sp];
      if (color[v] == 0) {
        color[v] = 1;
        S[sp++] = v;
        const mg_vtx *V = &g->vx[G[v]];
        for (int k = 0; k < V->adjs.n; ++k) {
          const int a = g->vx[g->ip[V->adjs.off + k]].id;
          if (color[a] == 0) {
            if (sp == scap) { g->err = MG_E_SCRATCH; return; }
            S[sp++] = a;
          } else if (color[a] == 1) acyclic = 0;
        }
      } else if (color[v] == 1) { color[v] = 2; ids[v] = --next_id; }
    }
    for (int i = 0; i < n && sp == 0; ++i) if (color[i] == 0) { acyclic = 0; S[sp++] = i; }
  } while (sp);
  if (!acyclic) { g->err = MG_E_CYCLIC; return; }
  for (int i = 0; i < n; ++i) { g->vx[G[i]].id = ids[i]; H[ids[i]] = G[i]; }
  for (int i = 0; i < n; ++i) {
    mg_vtx *h = &g->vx[H[i]];
    if (h->adjs.n > 1) mg_sort_by_id(g, &h->adjs);
    if (h->incs.n > 1) mg_sort_by_id(g, &h->incs);
  }
  /* the list headers live in the pool as well: take their addresses again after every push (the pool itself never moves) */
  for (int i = n - 1; i >= 0 && !g->err; --i) {
    const int vi = H[i];
    const mg_vtx *v = &g->vx[vi];
    for (int q = 0; q < n; ++q) reach[q] = 0;
    reach[i] = 1;
    mg_push(g, (mg_list *)(g->ip + oStar) + i, vi);
    for (int k = 0; k < v->adjs.n && !g->err; ++k) {
      const int wi = g->ip[v->adjs.off + k];
      const mg_vtx *w = &g->vx[wi];
      const int early_end = w->p + w->l < v->p + v->l || w->t + w->l < v->t + v->l;
      if (!reach[w->id] || w->p < v->p || w->t < v->t || early_end) {
        mg_push(g, (mg_list *)(g->ip + oRed) + i, wi);
        mg_push(g, (mg_list *)(g->ip + oRinc) + w->id, vi);
        if (!early_end) {
          const mg_list *sw = (const mg_list *)(g->ip + oStar) + w->id;
          for (int q = 0; q < sw->n && !g->err; ++q) {
            const int wai = g->ip[sw->off + q];
            const mg_vtx *wa = &g->vx[wai];
            if (!reach[wa->id] && v->t <= wa->t && v->p <= wa->p && v->t + v->l <= wa->t + wa->l && v->p + v->l <= wa->p + wa->l) {
              reach[wa->id] = 1;
              mg_push(g, (mg_list *)(g->ip + oStar) + i, wai);
            }
          }
        }
      }
    }
  }
  if (g->err) return;
  for (int i = 0; i < n; ++i) {
    mg_vtx *h = &g->vx[H[i]];
    h->adjs = ((const mg_list *)(g->ip + oRed))[i];
    h->incs = ((const mg_list *)(g->ip + oRinc))[i];
  }
}

/* ---- short-edge compaction (meg-simplification.c:258-312) --------------------------------------------------------- */
MG_FN void mg_compact_short_edges(mg_graph *g) {
  int changed;
  do {
    changed = 0;
    for (int gs = 1; gs < g->norder && !g->err;) {
      /* one list slot = the run of vertices with this p; vertices made here join the end of their slot */
      int ge = gs + 1;
      while (ge < g->norder && g->vx[g->order[ge]].p == g->vx[g->order[gs]].p) ++ge;
      for (int a = gs; a < ge && !g->err; ++a) {
        const int was_last = a == ge - 1;            /* successor cached when the walk reached this element */
        const int pi = g->order[a];
        for (int k = 0; k < g->vx[pi].adjs.n && !g->err;) {
          mg_vtx *p = &g->vx[pi];
          const int qi = g->ip[p->adjs.off + k];
          mg_vtx *q = &g->vx[qi];
          if (q->t != MG_SINK_START && q->t + q->l - p->t == q->p + q->l - p->p && q->t >= p->t + p->l && q->t - p->t - p->l <= 3) {
            changed = 1;
            mg_remove_at(g, &p->adjs, k);
            mg_remove_first(g, &q->incs, pi);
            const int ni = mg_new_vtx(g, p->p, p->t, q->p + q->l - p->p);
            if (g->err) break;
            mg_vtx *nv = &g->vx[ni];
            for (int x = 0; x < q->adjs.n; ++x) { const int ti = g->ip[q->adjs.off + x]; mg_push(g, &nv->adjs, ti); mg_push(g, &g->vx[ti].incs, ni); }
            for (int y = 0; y < p->incs.n; ++y) { const int si = g->ip[p->incs.off + y]; mg_push(g, &nv->incs, si); mg_push(g, &g->vx[si].adjs, ni); }
            for (int y = g->norder; y > ge; --y) g->order[y] = g->order[y - 1];      /* order has room for vx_cap entries */
            g->order[ge++] = ni; ++g->norder;
            continue;
          }
          ++k;
        }
        if (was_last) break;
      }
      gs = ge;
    }
    if (g->err) return;
    mg_prune_dead_ends(g);
  } while (changed);
}

MG_FN int mg_too_complex(const mg_graph *g, int n, int l, const pc_meg_cfg *c) {     /* is_too_complex, :89-140 */
  int min_len = 0;
  unsigned long long freq = 0, np = 0, ne = 0;
  const unsigned long long est_len = (unsigned long long)n - 2ull;
  for (int x = 0; x < g->norder; ++x) {
    const mg_vtx *p = &g->vx[g->order[x]];
    ++np;
    if (min_len == 0 || p->l < min_len) { min_len = p->l; freq = 1; } else if (p->l == min_len) ++freq;
    ne += (unsigned long long)p->adjs.n;
  }
  if (np < 5 || ne < 4) return 0;
  if (c->max_pairings_in_MEG != 0 && np > c->max_pairings_in_MEG && (double)freq > c->max_freq_shortest_pairing * (double)np) return 1;
  return ne > 5 * np || np > (2 * est_len) / (unsigned long long)l || (np > est_len / (unsigned long long)l && np >= 50);
}

/* One pass of build_meg's loop (compute-est-fact.c:103-150) for pairing length l.  Returns 1 when the graph is too complex
 * and a longer pairing length is still possible (the caller seeds again with l + 1), else 0. */
MG_FN int mg_build(mg_graph *g, int est_len, int l, const pc_meg_cfg *c) {
  const int n = est_len + 2;
  mg_build_edges(g, n, l, c);
  if (g->err) return 0;
  mg_remove_useless_edges(g, l, c);
  mg_prune_dead_ends(g);
  if (c->flags & PC_MEG_TRANS_RED) mg_transitive_reduction(g);
  if (g->err) return 0;
  long long np, ne;
  mg_stats(g, &np, &ne);
  int cx = ne > 1000 || np > 2000;
  if (!cx && (c->flags & PC_MEG_SHORT_EDGE_COMP)) mg_compact_short_edges(g);
  if (g->err) return 0;
  cx = cx || mg_too_complex(g, n, l, c);
  return cx && (long long)l + 1 + 2 < (long long)n;
}

/* int32 words of the MEG record (include/pintron_cuda.h) */
MG_FN long long mg_record_words(const mg_graph *g) {
  long long np, ne;
  mg_stats(g, &np, &ne);
  return 4 + 4 * np + ne;
}
MG_FN void mg_write_record(mg_graph *g, int retry, int32_t *out) {
  long long np, ne;
  mg_stats(g, &np, &ne);
  const int nv = g->norder;
  for (int x = 0; x < nv; ++x) g->vx[g->order[x]].id = x;
  out[0] = nv; out[1] = (int32_t)ne; out[2] = retry; out[3] = 0;
  int32_t *ptl = out + 4, *cnt = ptl + 3 * nv, *adj = cnt + nv;
  for (int x = 0; x < nv; ++x) {
    const mg_vtx *v = &g->vx[g->order[x]];
    ptl[3 * x] = v->p; ptl[3 * x + 1] = v->t; ptl[3 * x + 2] = v->l;
    cnt[x] = v->adjs.n;
    for (int k = 0; k < v->adjs.n; ++k) *adj++ = g->vx[g->ip[v->adjs.off + k]].id;
  }
}
#endif
