/* meg.c — the Maximal Embedding Graph of one EST.  The graph is BUILT ON THE DEVICE: one PC_OP_SEED job with
 * PC_SEED_BUILD_MEG runs build_vertex_set, build_edge_set, simplify_meg, transitive_reduction, compact_short_edges and
 * is_too_complex (reference src/compute-est-fact.c:90-152, src/max-emb-graph.c:217-672, src/meg-simplification.c:52-632;
 * ours: csrc/k_seed.cu + csrc/meg_core.h).  Vertex sets above EF_MEG_DEVICE_MAX pairings (mRNAs of several kbp) come back
 * as they are and the same meg_core.h walks them here: the walk is sequential and quadratic, one GPU lane needs tens of
 * milliseconds for what a host core does in well under one, and the whole batch would wait for it.  What is left here: asking for it (and again with a longer pairing length while
 * the device says "too complex"), turning the returned record into the pointer graph the embedding enumeration walks, and
 * the text of megs.txt / processed-megs.txt / meg-edges.txt (src/io-meg.c:147-190, src/max-emb-graph.c:677-707).
 */
#include "ef.h"
#include <pthread.h>
#include "../csrc/meg_core.h"      /* the MEG core of the device, plain C: the host runs it on the vertex sets the device hands back */

void meg_stats(const ef_meg *M, size_t *pairings, size_t *edges) { *pairings = M->np; *edges = M->ne; }

/* The options build_meg reads, in the form the device wants them (include/pintron_cuda.h); one copy per process: the
 * request holds a pointer to it until its batch is gathered. */
static pc_meg_cfg g_meg_cfg;
static const ef_config *g_meg_cfg_of;
static const pc_meg_cfg *meg_cfg(const ef_config *c) {
  if (__atomic_load_n(&g_meg_cfg_of, __ATOMIC_ACQUIRE) != c) {
    static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_mutex_lock(&mu);
    if (g_meg_cfg_of != c) {
      g_meg_cfg.min_intron_length = c->min_intron_length; g_meg_cfg.max_intron_length = c->max_intron_length;
      g_meg_cfg.max_pairings_in_MEG = c->max_pairings_in_MEG;
      g_meg_cfg.flags = (c->trans_red ? PC_MEG_TRANS_RED : 0u) | (c->short_edge_comp ? PC_MEG_SHORT_EDGE_COMP : 0u);
      g_meg_cfg.max_prefix_discarded_rate = c->max_prefix_discarded_rate; g_meg_cfg.max_suffix_discarded_rate = c->max_suffix_discarded_rate;
      g_meg_cfg.max_freq_shortest_pairing = c->max_freq_shortest_pairing;
      __atomic_store_n(&g_meg_cfg_of, c, __ATOMIC_RELEASE);
    }
    pthread_mutex_unlock(&mu);
  }
  return &g_meg_cfg;
}

/* One device job per attempt: vertex set, edges, simplification, transitive reduction, compaction and the complexity test
 * all run on the GPU (PC_OP_SEED with PC_SEED_BUILD_MEG; csrc/k_seed.cu + csrc/meg_core.h); what comes back is the
 * finished graph in list order.  The "too complex" retry (compute-est-fact.c:131-146) asks again with a longer pairing. */
#define EF_MEG_DEVICE_MAX_DEFAULT 32      /* measured: C3 does not care (16 .. 64 .. all on the device within noise), C4 prefers the low end (24: 4.3 k reads/s, 64: 3.6 k, all on the device: 0.28 k) */
static int meg_device_max(void) {
  static int v = -1;
  if (v < 0) { const char *e = getenv("EF_MEG_DEVICE_MAX"); v = e && atoi(e) >= 0 ? atoi(e) : EF_MEG_DEVICE_MAX_DEFAULT; }      /* 0 = always on the device */
  return v;
}

/* the device's MEG record (include/pintron_cuda.h) from a vertex set, on this core; caller frees */
static int32_t *meg_record_on_host(const int32_t *tri, int ntri, int est_len, int l, const pc_meg_cfg *mc) {
  for (long long nints = (1 << 15) + 256ll * ntri;; nints *= 2) {
    int *mem = malloc(sizeof(int) * (size_t)nints);
    if (!mem) { fprintf(stderr, "* FATAL est-fact: out of memory\n"); exit(1); }
    mg_graph g;
    mg_init(&g, mem, nints, tri, ntri);
    const int retry = g.err ? 0 : mg_build(&g, est_len, l, mc);
    if (g.err == MG_E_SCRATCH) { free(mem); continue; }
    if (g.err) { fprintf(stderr, "* FATAL est-fact: cyclic embedding graph\n"); exit(1); }
    int32_t *rec = malloc(sizeof(int32_t) * (size_t)(mg_record_words(&g) + 4));
    if (!rec) { fprintf(stderr, "* FATAL est-fact: out of memory\n"); exit(1); }
    mg_write_record(&g, retry, rec);
    free(mem);
    return rec;
  }
}

ef_meg *meg_build(ef_task *T, const ef_seq *est, unsigned *inc) {
  const ef_config *c = T->cfg;
  const pc_meg_cfg *mc = meg_cfg(c);
  const int ph_ = ef_phase(EF_PH_SEED);
  int cap = 256;       /* 12-byte units; the kernel reports the needed count when this is too small */
  for (;;) {
    const int l = (int)(c->min_factor_len + *inc);
    const int h = dp_push(PC_OP_SEED, S_(est->seq, est->len), S_((const char *)mc, (int)sizeof *mc), l, PC_SEED_BUILD_MEG, meg_device_max(), cap);
    dp_wait();
    const int32_t *r = dp_res(h);
    if (r[0] == PC_E_OUTCAP) { cap = r[1] + 16; continue; }
    const int32_t *rec = (const int32_t *)dp_var(h);
    int32_t *own = NULL;
    if (r[3] == PC_SEED_VERTEX_SET_ONLY) {
      ef_phase(EF_PH_MEG);
      rec = own = meg_record_on_host(rec, r[1], est->len, l, mc);
      ef_phase(EF_PH_SEED);
    }
    if (rec[2]) { ++*inc; free(own); continue; }
    ef_phase(EF_PH_MEG);
    const int nv = rec[0], ne = rec[1];
    const int32_t *ptl = rec + 4, *cnt = ptl + 3 * nv, *adj = cnt + nv;
    ef_meg *M = ar_alloc(&T->ar, sizeof *M);
    M->n = est->len + 2;
    M->nflat = nv; M->np = (size_t)nv; M->ne = (size_t)ne;
    ef_pairing *P = ar_alloc(&T->ar, sizeof(ef_pairing) * (size_t)(nv + 1));
    M->flat = ar_alloc_raw(&T->ar, sizeof(ef_pairing *) * (size_t)(nv + 1));
    ef_pairing **E = ar_alloc_raw(&T->ar, sizeof(ef_pairing *) * (size_t)(ne + 1));
    for (int x = 0; x < nv; ++x) {
      ef_pairing *q = &P[x];
      q->p = ptl[3 * x]; q->t = ptl[3 * x + 1]; q->l = ptl[3 * x + 2]; q->id = x;
      q->adjs.v = E; q->adjs.n = q->adjs.cap = cnt[x];
      for (int k = 0; k < cnt[x]; ++k) *E++ = &P[*adj++];
      M->flat[x] = q;
    }
    free(own);
    ef_phase(ph_);
    return M;
  }
}

/* ---- megs.txt / meg-edges.txt (io-meg.c:147-190, max-emb-graph.c:677-707) ------------------------------------ */
void meg_write(ef_buf *b, ef_meg *M) {
  for (int x = 0; x < M->nflat; ++x) {
    const ef_pairing *p = M->flat[x];
    const int v[3] = {p->p, p->t, p->l};
    buf_ints(b, "(", v, 3, ',', ")\n");
  }
  buf_printf(b, "#adj#\n");
  for (int x = 0; x < M->nflat; ++x) {
    const ef_pairing *p = M->flat[x];      /* ids = position in list order (set when the record was read) */
    for (int a = 0; a < p->adjs.n; ++a) { const int v[2] = {p->id, p->adjs.v[a]->id}; buf_ints(b, "", v, 2, '-', "\n"); }
  }
}

void meg_write_edges(ef_buf *b, const ef_meg *M) {
  for (int x = 0; x < M->nflat; ++x) {
    const ef_pairing *p = M->flat[x];
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
