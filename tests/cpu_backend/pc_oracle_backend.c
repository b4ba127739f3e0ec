/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * An implementation of include/pintron_cuda.h on top of the CPU oracle (oracle/port/dp_port.c), used ONLY by
 * tests/test_host_parity.py to exercise the C host program (pintron_b200/host/) in this GPU-less container:
 * tests/Makefile links the host sources with THIS file into tests/_build/est-fact-oracle-backend.  The shipped
 * est-fact links libpintron_cuda.so and nothing else; it has no CPU path (pc_ctx_create fails without a device).
 * Jobs run synchronously inside pc_submit; pc_stream_sync is a no-op.  The engine half (pc_engine_*, the lane protocol of
 * include/pintron_engine.h) is implemented too — memfd segments, one thread that serves posted lanes through the
 * pc_submit below — so that the in-process and the est-factd forms of the host can both be exercised without a GPU.
 */
#define _GNU_SOURCE
#include "pintron_cuda.h"
#include "pintron_engine.h"
#include <pthread.h>
#include <sys/mman.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../pintron_b200/csrc/meg_core.h"      /* the MEG core the device runs (PC_SEED_BUILD_MEG), compiled for the host */

int po_align(const char *est, int n, const char *gen, int m, uint8_t *ops, int *ops_len);
unsigned po_edit(const char *s1, int l1, const char *s2, int l2);
int po_kband(const char *a, int la, const char *b, int lb, unsigned k, unsigned *edit);
int po_borders(const char *p, int len_p, int min_cut, int max_cut, const char *t, int len_t, unsigned max_errs, int out[4]);
int po_gap(const char *est, int n, const char *gen, int m, uint8_t *ops, int pos[5]);
int po_affix(const char *est, int estl, const char *gen, int genl, int *ecut_out, int *gcut_out);
unsigned po_suffix_cut(const char *s1, int l1, const char *s2, int l2, int *c1, int *c2);
unsigned po_prefix_cut(const char *s1, int l1, const char *s2, int l2, int *c1, int *c2);
void po_lcs(const char *s1, long l1, const char *s2, long l2, long *occ1, long *occ2, long *len);
long po_seed(const char *T, long G, const char *P, long n, int mfl, double rate, int *out, long cap);

struct pc_ctx { char *genome; size_t len; int word; double rate; uint64_t ghash; };
struct pc_stream { pc_ctx *ctx; };
static unsigned long long g_jobs;

const char *pc_last_error(void) { return "oracle backend"; }
int pc_device_count(void) { return 1; }
pc_ctx *pc_ctx_create(int device) { (void)device; return calloc(1, sizeof(pc_ctx)); }
void pc_ctx_destroy(pc_ctx *c) { if (c) { free(c->genome); free(c); } }
int pc_genome_upload(pc_ctx *c, const char *genome, size_t len, int word_len, double depth_rate) {
  free(c->genome);
  c->genome = malloc(len + 16);
  memset(c->genome, 0, len + 16);
  memcpy(c->genome, genome, len);
  c->len = len; c->word = word_len; c->rate = depth_rate;
  uint64_t h = 1469598103934665603ull ^ (uint64_t)word_len ^ (uint64_t)(depth_rate * 1e6);
  for (size_t i = 0; i < len; ++i) h = (h ^ (uint8_t)genome[i]) * 1099511628211ull;
  c->ghash = h;
  return 0;
}
pc_stream *pc_stream_create(pc_ctx *c) { pc_stream *s = calloc(1, sizeof *s); s->ctx = c; return s; }
void pc_stream_destroy(pc_stream *s) { free(s); }
void *pc_host_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void pc_host_free(void *p) { free(p); }
int pc_stream_sync(pc_stream *s) { (void)s; return 0; }
uint64_t pc_launch_count(void) { return g_jobs; }
void pc_debug_dump(void) {}
void pc_set_blocking_sync(int on) { (void)on; }

/* ---- optional memo of job results (developer tool: PC_ORACLE_MEMO=<file>) ---------------------------------------------
 * The oracle is orders of magnitude slower than the device, so a profile of the HOST code over this backend is all
 * waiting.  With a memo file a second run answers every job from the table (key = hash of the job's parameters and
 * input bytes) and the host code runs at its own speed; the output bytes are checked as always. */
typedef struct memo_ent { uint64_t k0, k1; int32_t res[PC_RES_INTS]; uint32_t var_len; uint8_t *var; } memo_ent;
static struct { memo_ent *tab; size_t cap, n; pthread_mutex_t mu; const char *path; int loaded; size_t hits, misses; } g_memo = {.mu = PTHREAD_MUTEX_INITIALIZER};
static void memo_insert_locked(const memo_ent *e) {
  if ((g_memo.n + 1) * 2 > g_memo.cap) {
    const size_t ncap = g_memo.cap ? g_memo.cap * 2 : (size_t)1 << 16;
    memo_ent *nt = calloc(ncap, sizeof *nt);
    for (size_t i = 0; i < g_memo.cap; ++i) if (g_memo.tab[i].k0 | g_memo.tab[i].k1) { size_t h = g_memo.tab[i].k0 & (ncap - 1); while (nt[h].k0 | nt[h].k1) h = (h + 1) & (ncap - 1); nt[h] = g_memo.tab[i]; }
    free(g_memo.tab); g_memo.tab = nt; g_memo.cap = ncap;
  }
  size_t h = e->k0 & (g_memo.cap - 1);
  while (g_memo.tab[h].k0 | g_memo.tab[h].k1) { if (g_memo.tab[h].k0 == e->k0 && g_memo.tab[h].k1 == e->k1) return; h = (h + 1) & (g_memo.cap - 1); }
  g_memo.tab[h] = *e; ++g_memo.n;
}
static void memo_save(void) {
  if (!g_memo.path || !g_memo.misses) return;
  FILE *f = fopen(g_memo.path, "wb");
  if (!f) return;
  for (size_t i = 0; i < g_memo.cap; ++i) {
    const memo_ent *e = &g_memo.tab[i];
    if (!(e->k0 | e->k1)) continue;
    fwrite(&e->k0, 8, 1, f); fwrite(&e->k1, 8, 1, f); fwrite(e->res, sizeof e->res, 1, f); fwrite(&e->var_len, 4, 1, f);
    if (e->var_len) fwrite(e->var, 1, e->var_len, f);
  }
  fclose(f);
  fprintf(stderr, "oracle backend memo: %zu entries saved (%zu hits, %zu misses)\n", g_memo.n, g_memo.hits, g_memo.misses);
}
static void memo_load_locked(void) {
  g_memo.loaded = 1;
  g_memo.path = getenv("PC_ORACLE_MEMO");
  if (!g_memo.path) return;
  atexit(memo_save);
  FILE *f = fopen(g_memo.path, "rb");
  if (!f) return;
  memo_ent e;
  while (fread(&e.k0, 8, 1, f) == 1 && fread(&e.k1, 8, 1, f) == 1 && fread(e.res, sizeof e.res, 1, f) == 1 && fread(&e.var_len, 4, 1, f) == 1) {
    e.var = e.var_len ? malloc(e.var_len) : NULL;
    if (e.var_len && fread(e.var, 1, e.var_len, f) != e.var_len) break;
    memo_insert_locked(&e);
  }
  fclose(f);
}
static uint64_t mix(uint64_t h, const void *p, size_t n, uint64_t mul) {
  const uint8_t *b = p;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, b + i, 8); h = (h ^ w) * mul; h ^= h >> 29; }
  for (; i < n; ++i) h = (h ^ b[i]) * mul;
  return h ^ (h >> 31);
}
static void memo_key(const pc_ctx *c, const pc_job *j, const char *a, const char *b, uint64_t *k0, uint64_t *k1) {
  const uint32_t hdr[8] = {j->op, j->a_len, j->b_len, (uint32_t)j->p0, (uint32_t)j->p1, (uint32_t)j->p2, j->out_cap, j->flags & ~(uint32_t)PC_B_IN_GENOME};
  uint64_t x = mix(0x9E3779B97F4A7C15ull, hdr, sizeof hdr, 0xff51afd7ed558ccdull), y = mix(0xc4ceb9fe1a85ec53ull, hdr, sizeof hdr, 0x9fb21c651e98df25ull);
  x = mix(x, a, j->a_len, 0xff51afd7ed558ccdull); y = mix(y, a, j->a_len, 0x9fb21c651e98df25ull);
  if (j->op != PC_OP_SEED || j->p1 == PC_SEED_BUILD_MEG) { x = mix(x, b, j->b_len, 0xff51afd7ed558ccdull); y = mix(y, b, j->b_len, 0x9fb21c651e98df25ull); }
  if (j->op == PC_OP_SEED) { x ^= c->ghash; y += c->ghash * 0x9E3779B97F4A7C15ull; }        /* the genome, word length and depth rate of the session */
  *k0 = x | 1; *k1 = y;
}
static int memo_get(uint64_t k0, uint64_t k1, int32_t *res, uint8_t *var) {
  int hit = 0;
  pthread_mutex_lock(&g_memo.mu);
  if (!g_memo.loaded) memo_load_locked();
  if (g_memo.cap) {
    size_t h = k0 & (g_memo.cap - 1);
    while (g_memo.tab[h].k0 | g_memo.tab[h].k1) {
      if (g_memo.tab[h].k0 == k0 && g_memo.tab[h].k1 == k1) { memcpy(res, g_memo.tab[h].res, sizeof g_memo.tab[h].res); if (g_memo.tab[h].var_len) memcpy(var, g_memo.tab[h].var, g_memo.tab[h].var_len); hit = 1; break; }
      h = (h + 1) & (g_memo.cap - 1);
    }
  }
  if (hit) ++g_memo.hits; else ++g_memo.misses;
  pthread_mutex_unlock(&g_memo.mu);
  return hit;
}
static void memo_put(uint64_t k0, uint64_t k1, const int32_t *res, const uint8_t *var, uint32_t var_len) {
  memo_ent e; e.k0 = k0; e.k1 = k1; memcpy(e.res, res, sizeof e.res); e.var_len = var_len; e.var = var_len ? malloc(var_len) : NULL;
  if (var_len) memcpy(e.var, var, var_len);
  pthread_mutex_lock(&g_memo.mu);
  memo_insert_locked(&e);
  pthread_mutex_unlock(&g_memo.mu);
}

int pc_submit(pc_stream *st, const uint8_t *arena, size_t arena_bytes, const pc_job *jobs, int njobs, int32_t *res,
              uint8_t *var_out, size_t var_out_bytes) {
  (void)arena_bytes; (void)var_out_bytes;
  const pc_ctx *c = st->ctx;
  for (int i = 0; i < njobs; ++i) {
    const pc_job *j = &jobs[i];
    int32_t *r = res + (size_t)i * PC_RES_INTS;
    memset(r, 0, sizeof(int32_t) * PC_RES_INTS);
    const char *a = (const char *)arena + j->a_off;
    const char *b = (j->flags & PC_B_IN_GENOME) ? c->genome + j->b_off : (const char *)arena + j->b_off;
    const int la = (int)j->a_len, lb = (int)j->b_len;
    ++g_jobs;
    static int memo_on = -1;
    if (memo_on < 0) memo_on = getenv("PC_ORACLE_MEMO") != NULL;
    uint64_t k0 = 0, k1 = 0;
    if (memo_on) {
      memo_key(c, j, a, b, &k0, &k1);
      if (memo_get(k0, k1, r, var_out + j->out_off)) continue;
    }
    switch (j->op) {
      case PC_OP_ALIGN: {
        if ((uint32_t)(la + lb) > j->out_cap) { r[0] = PC_E_OUTCAP; break; }
        int n = 0; r[1] = po_align(a, la, b, lb, var_out + j->out_off, &n); r[2] = n; break;
      }
      case PC_OP_KBAND: { unsigned e = 0; r[1] = po_kband(a, la, b, lb, (unsigned)j->p0, &e); r[2] = (int32_t)e; break; }
      case PC_OP_EDIT: r[1] = (int32_t)po_edit(a, la, b, lb); break;
      case PC_OP_BORDERS: {
        if (j->p1 < 0 || j->p1 > j->p2 || j->p2 > la) { r[0] = PC_E_ARG; break; }
        char *tz = NULL;
        if (j->flags & PC_B_NUL_AFTER) { tz = calloc((size_t)lb + 2, 1); memcpy(tz, b, (size_t)lb); b = tz; }
        int out[4]; r[1] = po_borders(a, la, j->p1, j->p2, b, lb, (unsigned)j->p0, out);
        free(tz);
        r[2] = out[0]; r[3] = out[1]; r[4] = out[2]; r[5] = out[3]; break;
      }
      case PC_OP_GAP: {
        if ((uint32_t)(la + lb) > j->out_cap) { r[0] = PC_E_OUTCAP; break; }
        int pos[5]; r[1] = po_gap(a, la, b, lb, var_out + j->out_off, pos);
        for (int q = 0; q < 5; ++q) r[2 + q] = pos[q];
        break;
      }
      case PC_OP_AFFIX: { int e = 0, g = 0; r[1] = po_affix(a, la, b, lb, &e, &g); r[2] = r[1] ? e : 0; r[3] = r[1] ? g : 0; break; }
      case PC_OP_SUFCUT: { int c1, c2; r[1] = (int32_t)po_suffix_cut(a, la, b, lb, &c1, &c2); r[2] = c1; r[3] = c2; break; }
      case PC_OP_PRECUT: { int c1, c2; r[1] = (int32_t)po_prefix_cut(a, la, b, lb, &c1, &c2); r[2] = c1; r[3] = c2; break; }
      case PC_OP_LCS: { long o1, o2, ln; po_lcs(b, lb, a, la, &o1, &o2, &ln); r[1] = (int32_t)ln; r[2] = (int32_t)o1; r[3] = (int32_t)o2; break; }
      case PC_OP_SEED: {
        int *out = (int *)(var_out + j->out_off);
        if (j->p1 == PC_SEED_BUILD_MEG) {
          if (lb != (int)sizeof(pc_meg_cfg)) { r[0] = PC_E_ARG; break; }
          pc_meg_cfg cfg; memcpy(&cfg, b, sizeof cfg);
          long cap = 1024, n;
          int *tri = malloc(sizeof(int) * 3 * (size_t)cap);
          while ((n = po_seed(c->genome, (long)c->len, a, la, j->p0, c->rate, tri, cap)) < 0) { cap = -n + 16; tri = realloc(tri, sizeof(int) * 3 * (size_t)cap); }
          if (j->p2 > 0 && n > (long)j->p2) {                  /* the device's rule: a large vertex set goes back as it is */
            if (n > (long)j->out_cap) { r[0] = PC_E_OUTCAP; r[1] = (int32_t)n; }
            else { memcpy(out, tri, sizeof(int) * 3 * (size_t)n); r[1] = (int32_t)n; r[3] = PC_SEED_VERTEX_SET_ONLY; }
            free(tri);
            break;
          }
          long long nints = 1 << 16;
          for (;;) {
            int *mem = malloc(sizeof(int) * (size_t)nints);
            mg_graph g;
            mg_init(&g, mem, nints, tri, (int)n);
            const int retry = g.err ? 0 : mg_build(&g, la, j->p0, &cfg);
            if (g.err == MG_E_SCRATCH) { free(mem); nints *= 2; continue; }
            if (g.err) r[0] = PC_E_RANGE;
            else {
              const long long units = (mg_record_words(&g) + 2) / 3;
              if (units > (long long)j->out_cap) { r[0] = PC_E_OUTCAP; r[1] = (int32_t)units; }
              else { mg_write_record(&g, retry, out); r[1] = (int32_t)units; }
            }
            free(mem);
            break;
          }
          free(tri);
          break;
        }
        long n = po_seed(c->genome, (long)c->len, a, la, j->p0, c->rate, out, (long)j->out_cap);
        if (n < 0) { r[0] = PC_E_OUTCAP; r[1] = (int32_t)-n; } else r[1] = (int32_t)n;
        break;
      }
      default: r[0] = PC_E_ARG;
    }
    if (memo_on) {
      uint32_t vl = 0;
      if (r[0] == 0 && (j->op == PC_OP_ALIGN || j->op == PC_OP_GAP)) vl = (uint32_t)(la + lb);
      if (r[0] == 0 && j->op == PC_OP_SEED) vl = (uint32_t)r[1] * 12u;
      memo_put(k0, k1, r, var_out + j->out_off, vl);
    }
  }
  return 0;
}


/* ---- engine: lanes in memfd segments, one serving thread ------------------------------------------------------------ */
#define SEG_BYTES ((size_t)64 << 20)
typedef struct slab { uint32_t seg; size_t off, bytes; } slab;
typedef struct be_session { uint32_t id; pc_ctx *ctx; int nlanes; uint32_t lane[PCE_MAX_SESSION_LANES]; pc_session_stats st; int open; } be_session;
struct pc_engine {
  int nsegs, fd[PCE_MAX_SEGMENTS]; uint8_t *base[PCE_MAX_SEGMENTS]; size_t bytes[PCE_MAX_SEGMENTS], used[PCE_MAX_SEGMENTS];
  pce_hdr *hdr;
  be_session ses[64]; uint32_t next_id; int live;
  pthread_mutex_t mu; pthread_t th[8]; int nth; volatile int stop;
};

static size_t up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static int add_seg(pc_engine *e, size_t bytes) {
  if (e->nsegs >= PCE_MAX_SEGMENTS) return -1;
  const int k = e->nsegs;
  e->fd[k] = memfd_create("pintron-lanes-test", MFD_CLOEXEC);
  if (e->fd[k] < 0 || ftruncate(e->fd[k], (off_t)bytes)) return -1;
  e->base[k] = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, e->fd[k], 0);
  if (e->base[k] == MAP_FAILED) return -1;
  e->bytes[k] = bytes; e->used[k] = k == 0 ? PCE_HDR_BYTES : 0;
  if (k == 0) { e->hdr = (pce_hdr *)e->base[0]; memset(e->hdr, 0, sizeof(pce_hdr)); e->hdr->magic = PCE_MAGIC; e->hdr->version = PCE_VERSION; }
  return e->nsegs++;
}
static int lane_slab(pc_engine *e, pce_lane *l, uint64_t arena_cap, uint32_t jobs_cap, uint64_t var_cap) {
  const size_t a = up(arena_cap + 64, 256), j = up(sizeof(pc_job) * (size_t)jobs_cap, 256), r = up(4u * PC_RES_INTS * (size_t)jobs_cap, 256), v = up(var_cap + 64, 256);
  const size_t total = up(a + j + r + v, 4096);
  int k = -1;
  for (int q = 0; q < e->nsegs; ++q) if (e->used[q] + total <= e->bytes[q]) { k = q; break; }
  if (k < 0) k = add_seg(e, total > SEG_BYTES ? up(total, 1 << 21) : SEG_BYTES);
  if (k < 0) return PC_E_NOMEM;
  const size_t off = e->used[k];
  e->used[k] += total;
  l->seg = (uint32_t)k; l->jobs_cap = jobs_cap; l->arena_off = off; l->arena_cap = arena_cap; l->jobs_off = off + a; l->res_off = off + a + j;
  l->var_off = off + a + j + r; l->var_cap = var_cap;
  return 0;
}

static void *engine_main(void *arg) {
  pc_engine *e = arg;
  pc_stream st;
  while (!e->stop) {
    const uint32_t bell = __atomic_load_n(&e->hdr->doorbell, __ATOMIC_SEQ_CST);
    int served = 0;
    for (uint32_t i = 0; i < PCE_MAX_LANES; ++i) {
      pce_lane *l = &e->hdr->lanes[i];
      uint32_t expect = PCE_POSTED;
      if (!__atomic_compare_exchange_n(&l->state, &expect, (uint32_t)PCE_RUNNING, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) continue;
      be_session *S = NULL;
      pthread_mutex_lock(&e->mu);
      for (int q = 0; q < 64; ++q) if (e->ses[q].open && e->ses[q].id == l->session) S = &e->ses[q];
      pthread_mutex_unlock(&e->mu);
      int rc = PC_E_ARG;
      if (S && l->njobs <= l->jobs_cap && l->arena_len <= l->arena_cap && l->var_len <= l->var_cap) {
        uint8_t *b = e->base[l->seg];
        st.ctx = S->ctx;
        rc = pc_submit(&st, b + l->arena_off, (size_t)l->arena_len, (const pc_job *)(b + l->jobs_off), (int)l->njobs, (int32_t *)(b + l->res_off),
                       b + l->var_off, (size_t)l->var_len);
        S->st.batches++; S->st.lanes_merged++; S->st.jobs += l->njobs;
      }
      l->rc = rc;
      __atomic_store_n(&l->state, (uint32_t)PCE_DONE, __ATOMIC_RELEASE);
      pce_futex(&l->state, FUTEX_WAKE, 64, NULL);
      ++served;
    }
    if (!served) {
      __atomic_fetch_add(&e->hdr->sleepers, 1u, __ATOMIC_SEQ_CST);
      struct timespec to = {0, 20 * 1000 * 1000};
      if (__atomic_load_n(&e->hdr->doorbell, __ATOMIC_SEQ_CST) == bell) pce_futex(&e->hdr->doorbell, FUTEX_WAIT, bell, &to);
      __atomic_fetch_sub(&e->hdr->sleepers, 1u, __ATOMIC_SEQ_CST);
    }
  }
  return NULL;
}

pc_engine *pc_engine_create(const int *devices, int ndev, size_t segment_bytes) {
  (void)devices; (void)ndev;
  pc_engine *e = calloc(1, sizeof *e);
  pthread_mutex_init(&e->mu, NULL);
  e->next_id = 1;
  if (add_seg(e, segment_bytes ? up(segment_bytes, 1 << 21) : SEG_BYTES) < 0) { free(e); return NULL; }
  e->nth = 8;                                    /* the oracle is slow: serve lanes in parallel */
  for (int k = 0; k < e->nth; ++k) pthread_create(&e->th[k], NULL, engine_main, e);
  return e;
}
void pc_engine_destroy(pc_engine *e) { if (!e) return; e->stop = 1; for (int k = 0; k < e->nth; ++k) pthread_join(e->th[k], NULL); free(e); }
int pc_engine_gpu_count(const pc_engine *e) { (void)e; return 1; }
const char *pc_engine_backend(void) { return "oracle-test"; }
void pc_engine_enable_timers(pc_engine *e, int on) { (void)e; (void)on; }
int pc_engine_open(pc_engine *e, const pc_session_req *req, pc_session_info *out) {
  pthread_mutex_lock(&e->mu);
  be_session *S = NULL;
  for (int q = 0; q < 64 && !S; ++q) if (!e->ses[q].open) S = &e->ses[q];
  if (!S || req->nlanes > PCE_MAX_SESSION_LANES) { pthread_mutex_unlock(&e->mu); return PC_E_NOMEM; }
  memset(S, 0, sizeof *S);
  S->id = e->next_id++; S->ctx = pc_ctx_create(0); S->nlanes = req->nlanes;
  pc_genome_upload(S->ctx, req->genome, req->genome_len, req->word_len, req->depth_rate);
  memset(out, 0, sizeof *out);
  for (int k = 0; k < req->nlanes; ++k) {
    uint32_t li = 0;
    while (li < PCE_MAX_LANES && (e->hdr->lanes[li].session || e->hdr->lanes[li].state != PCE_FREE)) ++li;
    if (li == PCE_MAX_LANES || lane_slab(e, &e->hdr->lanes[li], req->arena_cap, req->jobs_cap, req->var_cap)) { pthread_mutex_unlock(&e->mu); return PC_E_NOMEM; }
    e->hdr->lanes[li].session = S->id; e->hdr->lanes[li].state = PCE_IDLE;
    S->lane[k] = li; out->lane[k] = li;
  }
  out->session = S->id; out->gpu = 0; out->nlanes = req->nlanes;
  S->open = 1; ++e->live;
  pthread_mutex_unlock(&e->mu);
  return 0;
}
int pc_engine_resize_lane(pc_engine *e, uint32_t session, uint32_t lane, uint64_t arena_cap, uint32_t jobs_cap, uint64_t var_cap, uint64_t keep_arena,
                          uint32_t keep_jobs) {
  (void)session;
  pthread_mutex_lock(&e->mu);
  pce_lane *l = &e->hdr->lanes[lane], old = *l;
  int rc = lane_slab(e, l, arena_cap, jobs_cap, var_cap);          /* the old slab is simply left behind: test runs are short */
  if (!rc) {
    memcpy(e->base[l->seg] + l->arena_off, e->base[old.seg] + old.arena_off, (size_t)keep_arena);
    memcpy(e->base[l->seg] + l->jobs_off, e->base[old.seg] + old.jobs_off, sizeof(pc_job) * (size_t)keep_jobs);
  }
  pthread_mutex_unlock(&e->mu);
  return rc;
}
int pc_engine_close(pc_engine *e, uint32_t session, pc_session_stats *stats) {
  pthread_mutex_lock(&e->mu);
  for (int q = 0; q < 64; ++q) {
    be_session *S = &e->ses[q];
    if (!S->open || S->id != session) continue;
    for (int k = 0; k < S->nlanes; ++k) {
      pce_lane *l = &e->hdr->lanes[S->lane[k]];
      while (__atomic_load_n(&l->state, __ATOMIC_ACQUIRE) == PCE_RUNNING) usleep(100);
      memset(l, 0, sizeof *l);
    }
    if (stats) *stats = S->st;
    pc_ctx_destroy(S->ctx);
    S->open = 0;
    if (--e->live == 0) for (int k = 0; k < e->nsegs; ++k) e->used[k] = k == 0 ? PCE_HDR_BYTES : 0;      /* nobody left: all slabs are free again */
  }
  pthread_mutex_unlock(&e->mu);
  return 0;
}
int pc_engine_segment_count(pc_engine *e, int gpu) { (void)gpu; return e->nsegs; }
int pc_engine_segment_fd(pc_engine *e, int gpu, int seg, size_t *bytes) { (void)gpu; if (seg < 0 || seg >= e->nsegs) return -1; if (bytes) *bytes = e->bytes[seg]; return e->fd[seg]; }
void *pc_engine_segment_base(pc_engine *e, int gpu, int seg) { (void)gpu; return (seg < 0 || seg >= e->nsegs) ? NULL : e->base[seg]; }
