/* ef.h — shared declarations of the est-fact host program (C11).
 *
 * The host owns control flow only: FASTA prep, the maximal-embedding graph, embedding enumeration, the filter /
 * repair passes and output formatting.  Every DP (and the maximal-pairing discovery) is a job for
 * libpintron_cuda (include/pintron_cuda.h); the per-EST code pushes jobs with dp_push(), calls dp_wait() and is
 * suspended (it runs on a fiber) until the scheduler has run the jobs of many ESTs as one GPU batch.
 */
#ifndef EF_H
#define EF_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <limits.h>
#include "pintron_cuda.h"

#define MIN2(a, b) ((a) < (b) ? (a) : (b))
#define MAX2(a, b) ((a) > (b) ? (a) : (b))

/* ---- configuration (reference src/options.ggo, include/configuration.h:35-137) ------------------------ */
typedef struct ef_config {
  unsigned min_factor_len;
  int min_intron_length, max_intron_length;
  double min_string_depth_rate, max_prefix_discarded_rate, max_suffix_discarded_rate;
  int max_prefix_discarded, max_suffix_discarded;
  unsigned max_site_difference;
  int max_number_of_factorizations;
  double max_coverage_diff;
  int max_exonNUM_diff, max_gapLength_diff;
  char retain_externals;
  unsigned max_pairings_in_MEG;
  double max_freq_shortest_pairing;
  int suffpref_length_on_est, suffpref_length_for_intron, suffpref_length_on_gen;
  bool trans_red, short_edge_comp;
  unsigned max_single_factorization_time;
  double complexity_threshold;
  /* ours (not in the reference): execution knobs, none changes results */
  int threads, fibers, n_devices, devices[16];
  bool quiet, aux_outputs;
  char engine[16];               /* auto | daemon | inproc: where the batch engine runs (engine_client.h) */
} ef_config;

int ef_config_parse(ef_config *c, int argc, char **argv);   /* also writes ./config-dump.ini */

/* ---- per-EST bump arena --------------------------------------------------------------------------- */
typedef struct ef_chunk { struct ef_chunk *next; size_t cap, used; } ef_chunk;
typedef struct ef_arena { ef_chunk *head, *cur; } ef_arena;
void *ar_alloc(ef_arena *a, size_t bytes);         /* zeroed, 16-aligned */
void *ar_alloc_raw(ef_arena *a, size_t bytes);     /* the same, not zeroed */
void ar_undo(ef_arena *a, void *p);                /* rewind to p when it is in the current chunk (p = latest allocation) */
void ar_reset(ef_arena *a);                        /* keeps the first chunk */
void ar_free_all(ef_arena *a);

/* growable byte buffer (output text) */
typedef struct ef_buf { char *p; size_t len, cap; bool ext; } ef_buf;      /* ext: p is a slice of storage somebody else owns (never realloc / free it) */
void buf_printf(ef_buf *b, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void buf_write(ef_buf *b, const void *src, size_t n);
void buf_ints(ef_buf *b, const char *open, const int *v, int n, char sep, const char *close);
void buf_free(ef_buf *b);

/* ---- sequences ------------------------------------------------------------------------------------ */
typedef struct ef_seq {        /* mirrors struct _EST_info (include/types.h:140-174) where it matters */
  char *id;                    /* FASTA header without '>' */
  char *seq;                   /* EST_seq: strand-fixed, polyA/T masked ('*' / '#'); genome: N tails stripped */
  char *orig;                  /* original_EST_seq */
  char *gb;
  int len;
  int strand;
  bool fixed_strand;
  int pref_polyA, suff_polyA, pref_polyT, suff_polyT;
  int pref_N, suff_N;
  int abs_start, abs_end;
  char strand_as_read[12];
} ef_seq;

int ef_read_fasta(const char *path, ef_seq **out, size_t *n_out);
typedef struct ef_fasta ef_fasta;                     /* streaming reader: one record per call */
ef_fasta *ef_fasta_open(const char *path);
int ef_fasta_next(ef_fasta *r, ef_seq *out);
void ef_fasta_close(ef_fasta *r);
void ef_parse_genomic_header(ef_seq *g);
void ef_ntails_removal(ef_seq *g);
void ef_set_gb(ef_seq *e);
void ef_set_strand_and_rc(ef_seq *e);
void ef_reverse_complement(ef_seq *e);
void ef_polyAT_substitution(ef_seq *e);
void ef_make_rc_copy(const ef_seq *src, ef_seq *dst);

/* ---- MEG ------------------------------------------------------------------------------------------ */
typedef struct ef_pairing ef_pairing;
typedef struct ef_plist { ef_pairing **v; int n, cap; } ef_plist;      /* ordered list of pairing pointers */
struct ef_pairing {
  int p, t, l;
  int id;                       /* position in list order = the vertex number of megs.txt */
  ef_plist adjs;
  bool visited;
  void *memo;                   /* embeddings already enumerated from this vertex */
};
#define SRC_START INT_MIN
#define SINK_START (INT_MAX - 200)
#define SENTINEL_LEN 200

typedef struct ef_meg {
  int n;                        /* |P| + 2: the reference's list slots (V[0] = source, V[i+1] = pairings at p = i, V[n-1] = sink) */
  ef_pairing **flat; int nflat; /* the vertices in list order: V[0], V[1], ... concatenated (as the device returns them) */
  size_t np, ne;                /* pairings (source and sink included) and edges */
} ef_meg;

typedef struct ef_factor { int es, ee, gs, ge; } ef_factor;        /* EST_start, EST_end, GEN_start, GEN_end */
typedef struct ef_fz { ef_factor *f; int n, cap; bool polya, polyad; } ef_fz;   /* one factorization */
typedef struct ef_fzlist { ef_fz **v; int n, cap; } ef_fzlist;

/* ---- the per-EST task context ------------------------------------------------------------------------ */
typedef struct ef_task {
  const ef_config *cfg;
  const ef_seq *gen;
  ef_arena ar;
  /* outputs of this EST (concatenated in input order by the writer) */
  ef_buf out_raw, out_pest, out_megs, out_pmegs, out_edges;
  /* --max-single-factorization-time: charged with the time this EST's own code has RUN (fibers are suspended while
   * thousands of others run and while batches are with the engine; the reference times one EST running alone) */
  uint64_t run_ticks, run_mark, t_start;
} ef_task;

ef_fz *fz_new(ef_task *T, int cap);
void fz_push(ef_task *T, ef_fz *z, ef_factor f);
void fz_insert(ef_task *T, ef_fz *z, int at, ef_factor f);
void fz_remove(ef_fz *z, int at);
void fzl_push(ef_task *T, ef_fzlist *L, ef_fz *z);
void fzl_remove(ef_fzlist *L, int at);

/* ---- DP requests (sched.c) ----------------------------------------------------------------------------
 * dp_push queues one job for the calling fiber and returns its handle; dp_wait suspends the fiber until every
 * queued job has a result; dp_res / dp_var then read them (valid until the fiber's next dp_push). */
typedef struct ef_str { const char *p; int len; bool in_genome; int gen_off; bool nul_after; } ef_str;   /* b-side strings that lie inside
                                           the genome are sent as references (dp_push detects them); nul_after: BORDERS only */
static inline ef_str S_(const char *p, int len) { ef_str s = {p, len < 0 ? 0 : len, false, 0, false}; return s; }
static inline ef_str SZ_(const char *p, int len) { ef_str s = {p, len < 0 ? 0 : len, false, 0, true}; return s; }   /* t[len] reads as NUL */
int dp_push(int op, ef_str a, ef_str b, int p0, int p1, int p2, int out_cap);
void dp_wait(void);
/* body(T, k, user) for k = 0 .. n-1, as child fibers whose DP round trips overlap.  The iterations must touch disjoint data
 * (they share T and its arena cooperatively: same thread); the caller continues when all of them have returned. */
typedef void (*ef_par_fn)(ef_task *T, int k, void *user);
void dp_parallel_for(ef_task *T, int n, ef_par_fn body, void *user);
const int32_t *dp_res(int handle);
const uint8_t *dp_var(int handle);

/* blocking conveniences built on the three calls above */
unsigned dp_edit(const char *a, int la, const char *b, int lb);
bool dp_borders(const char *p, int len_p, int min_cut, int max_cut, const char *t, int len_t, unsigned max_errs,
                int *off_p, int *off_t1, int *off_t2, unsigned *ed);
void dp_lcs(const char *s1, long l1, const char *s2, long l2, long *occ1, long *occ2, long *len);

/* alignment rows rebuilt from column ops (1 spare NUL-filled margin of 16 bytes on both sides) */
typedef struct ef_aln { char *est, *gen; int dim; int score; } ef_aln;
ef_aln aln_from_ops(ef_task *T, const uint8_t *ops, int n_ops, const char *est, const char *gen);

/* ---- stages ------------------------------------------------------------------------------------------ */
ef_meg *meg_build(ef_task *T, const ef_seq *est, unsigned *inc_pairing_len);       /* incl. "too complex" retries */
void meg_stats(const ef_meg *M, size_t *pairings, size_t *edges);
void meg_write(ef_buf *b, ef_meg *M);
void meg_write_edges(ef_buf *b, const ef_meg *M);

ef_fzlist *est_factorizations(ef_task *T, const ef_seq *est, ef_meg *M, bool *timed_out);   /* get_EST_factorizations */
void refine_factorizations(ef_task *T, const ef_seq *est, ef_fzlist *L);                    /* refine_EST_factorizations + 2 removals */
bool refine_intron(ef_task *T, const ef_seq *est, ef_factor *donor, ef_factor *acceptor, bool first_intron);
int burset_freq(const char *donor, const char *acceptor);           /* getBursetFrequency */
int burset_adaptor(const char *t, size_t cut1, size_t cut2);        /* getBursetFrequency_adaptor */
char classify_intron(const char *gen, int glen, int start, int end);
void ef_small_exon_index_build(const char *g, size_t len);          /* 6-mer positions of the genome, once per run */
void ef_small_exon_scan(const char *g, int glen_all, const char *e, size_t estart, size_t elen, size_t allgstart, size_t allglen,
                        size_t f1slen, size_t f2plen, size_t MINI, size_t out[7]);      /* refine_fact.c */   /* 0 = U12, 1 = U2, 2 = not determined */
double dust_score(const char *s, int len);

/* shared by est_factorizations and refine_factorizations */
void clean_noisy_exons(ef_task *T, ef_fz *z, const char *gen, const char *est);
void clean_external_exons(ef_task *T, ef_fz *z, const char *gen, const char *est);
bool add_if_not_exists(ef_task *T, ef_fz *z, ef_fzlist *L);

bool ef_timeout_expired(ef_task *T);
double ef_now(void);
uint64_t ef_ticks(void);                    /* cheap monotonic counter (TSC) */
double ef_ticks_to_s(uint64_t ticks);
uint64_t ef_task_ticks(const ef_task *T);   /* ticks this task's code has run so far */

/* CPU-time accounting of the per-EST code (diagnostics; printed with the timers) */
enum { EF_PH_OTHER = 0, EF_PH_SEED, EF_PH_MEG, EF_PH_EMBED, EF_PH_CAND, EF_PH_FILTER, EF_PH_INTRON, EF_PH_REFINE, EF_PH_SMALLEX, EF_PH_OUTPUT, EF_PH_COUNT };
int ef_phase(int ph);                       /* returns the previous phase */
const double *sched_phase_seconds(void);
const uint64_t *sched_phase_yields(uint64_t *max_per_est);   /* engine round trips (dp_wait) per phase; the longest chain of one EST */

/* ---- scheduler entry ------------------------------------------------------------------------------------ */
typedef struct ef_job_result {          /* per input EST, filled by the workers */
  ef_buf raw, pest, megs, pmegs, edges;
  bool aligned;
  volatile int done;
} ef_job_result;

typedef void (*ef_task_fn)(ef_task *T, size_t handle, void *user);
/* the source of work: 1 = *handle is the next item, 0 = none available yet (the reader is behind), -1 = no more items, ever */
typedef int (*ef_next_fn)(void *user, size_t *handle);
void sched_prepare(const ef_config *cfg, const ef_seq *gen);   /* optional: start device set-up early, in the background */
/* n_hint: the number of items when the whole input is already known (short inputs get fewer threads / fibers), else SIZE_MAX */
int sched_run(const ef_config *cfg, const ef_seq *gen, size_t n_hint, ef_next_fn next, ef_task_fn fn, void *user);
void sched_stats(double *gpu_wait_s, uint64_t *batches, uint64_t *jobs);
struct pc_session_stats;
void sched_engine_stats(struct pc_session_stats *sum, const char **mode);   /* what the engine did for this run (all GPUs) */
void sched_bytes(uint64_t *h2d, uint64_t *d2h);                    /* bytes staged to / from the devices */
void sched_breakdown(double *fibers_s, double *gather_s, double *submit_s);   /* summed over worker threads */

#endif
