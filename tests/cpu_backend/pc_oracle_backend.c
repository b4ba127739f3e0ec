/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * An implementation of include/pintron_cuda.h on top of the CPU oracle (oracle/port/dp_port.c), used ONLY by
 * tests/test_host_parity.py to exercise the C host program (pintron_b200/host/) in this GPU-less container:
 * tests/Makefile links the host sources with THIS file into tests/_build/est-fact-oracle-backend.  The shipped
 * est-fact links libpintron_cuda.so and nothing else; it has no CPU path (pc_ctx_create fails without a device).
 * Jobs run synchronously inside pc_submit; pc_stream_sync is a no-op.
 */
#include "pintron_cuda.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

struct pc_ctx { char *genome; size_t len; int word; double rate; };
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
  return 0;
}
pc_stream *pc_stream_create(pc_ctx *c) { pc_stream *s = calloc(1, sizeof *s); s->ctx = c; return s; }
void pc_stream_destroy(pc_stream *s) { free(s); }
void *pc_host_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void pc_host_free(void *p) { free(p); }
int pc_stream_sync(pc_stream *s) { (void)s; return 0; }
uint64_t pc_launch_count(void) { return g_jobs; }
void pc_debug_dump(void) {}

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
        long n = po_seed(c->genome, (long)c->len, a, la, j->p0, c->rate, out, (long)j->out_cap);
        if (n < 0) { r[0] = PC_E_OUTCAP; r[1] = (int32_t)-n; } else r[1] = (int32_t)n;
        break;
      }
      default: r[0] = PC_E_ARG;
    }
  }
  return 0;
}
