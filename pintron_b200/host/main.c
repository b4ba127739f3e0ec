/* main.c — the est-fact program: same command line, same files in the working directory, same output bytes as
 * reference src/main-est-fact.c:90-339, with the per-EST work (src/compute-est-fact.c:192-293) spread over fibers,
 * threads and GPUs by sched.c.  Inputs ./genomic.txt and ./ests.txt; outputs raw-multifasta-out.txt,
 * processed-ests.txt (the two the pipeline consumes), megs.txt, processed-megs.txt, meg-edges.txt,
 * processed-megs-info.txt, info-pid-<pid>.log, config-dump.ini.
 *
 * One scheduler item = one input EST: its forward copy and, when the strand is not fixed and the forward copy gave
 * no factorization, its reverse-complement copy (main-est-fact.c:249-291) — so the order of the records written
 * for one EST never depends on other ESTs, and shards of ests.txt concatenate to the single-run output.
 */
#define _GNU_SOURCE
#include "ef.h"
#include "pintron_engine.h"
#include <unistd.h>
#include <pthread.h>
#include <stdatomic.h>

typedef struct est_item {
  ef_seq fwd, rc;
  bool has_rc;
  ef_buf out[6];                  /* raw, processed-ests, megs, processed-megs, meg-edges, processed-megs-info */
  char *out_block;                /* the six buffers start as slices of this one allocation (est_task_body) */
  _Atomic uint32_t done;          /* set by the worker when every buffer above is final (futex word) */
  _Atomic uint32_t waited;        /* a writer sleeps on `done` */
  _Atomic int written;            /* writers that are through with this item (the last one frees its strings) */
} est_item;
enum { O_RAW = 0, O_PEST, O_MEGS, O_PMEGS, O_EDGES, O_INFO, O_COUNT };

/* Streaming (the reference loads every EST first, io-multifasta.c:93-167): a reader thread parses
 * ests.txt into WINDOWS of records; the scheduler hands the items of the windows out as they become ready (inside a
 * window longest first, so ESTs in flight together have similar sizes); one writer per output file emits the windows in
 * input order and releases them.  The reader stays at most WIN_AHEAD windows in front of the slowest writer, which bounds
 * the memory of a run whatever the size of ests.txt. */
#define WIN_BITS 20
#define WIN_RECORDS 8192u
#define WIN_BYTES ((size_t)48 << 20)
#define WIN_AHEAD 10
#define MAX_WINDOWS ((size_t)1 << 22)
/* Sleeping instead of polling (six writers polling every 100 us cost a good part of a core per process — eight processes
 * of them on a 32-core box): a futex wait on a 32-bit word that changes whenever the condition may have become true;
 * the 50 ms time-out is a safety net only. */
#include <linux/futex.h>
#include <sys/syscall.h>
static void word_wait(_Atomic uint32_t *w, uint32_t seen) {
  const struct timespec to = {0, 50000000};
  syscall(SYS_futex, (uint32_t *)w, FUTEX_WAIT_PRIVATE, seen, &to, NULL, 0);
}
static void word_wake(_Atomic uint32_t *w) { syscall(SYS_futex, (uint32_t *)w, FUTEX_WAKE_PRIVATE, INT32_MAX, NULL, NULL, 0); }
static void seq_bump(_Atomic uint32_t *w) { atomic_fetch_add(w, 1); word_wake(w); }

typedef struct window {
  uint32_t n;
  est_item *items;
  uint32_t *order;                /* dispatch order inside the window */
  _Atomic uint32_t next;          /* items handed out so far */
  _Atomic int writers_done;
} window;

typedef struct run_ctx {
  const ef_config *cfg;
  const ef_seq *gen;
  const char *gen_orig;          /* the genome before N-tail removal (output bytes come from here) */
  window **win;                  /* MAX_WINDOWS slots; slot w is valid once n_ready > w */
  _Atomic size_t n_ready, cur;
  _Atomic size_t n_written[O_COUNT];
  _Atomic int eof;
  _Atomic size_t n_records;
  _Atomic uint32_t ready_seq;    /* bumped when a window becomes ready and at end of input */
  _Atomic uint32_t written_seq;  /* bumped when a writer finishes a window */
  FILE *f[O_COUNT];
  double writer_busy[O_COUNT];
  int rc;
} run_ctx;
#define it_raw out[O_RAW]
#define it_pest out[O_PEST]
#define it_megs out[O_MEGS]
#define it_pmegs out[O_PMEGS]
#define it_edges out[O_EDGES]
#define it_info out[O_INFO]

static void write_est_record(ef_buf *b, const ef_seq *e) { buf_printf(b, ">%s\n", e->id); buf_write(b, e->orig, strlen(e->orig)); buf_write(b, "\n", 1); }

/* write_multifasta_output (src/io-multifasta.c:187-241) */
static void write_factorizations(ef_buf *b, const run_ctx *R, const ef_seq *e, const ef_fzlist *L, const bool *polya, const bool *polyad) {
  const bool keep_ext = R->cfg->retain_externals;
  for (int k = 0; k < L->n; ++k) {
    const ef_fz *z = L->v[k];
    if (!(keep_ext || z->n > 2 || (z->n == 2 && e->suff_polyA != -1))) continue;
    buf_printf(b, ">%s\n", e->id);
    buf_printf(b, "#polya=%d\n#polyad=%d\n", keep_ext ? (int)polya[k] : 0, keep_ext ? (int)polyad[k] : 0);
    const unsigned l_index = keep_ext ? 0u : 1u;
    const unsigned r_index = keep_ext ? (unsigned)z->n + 1u : (e->suff_polyA == -1 ? (unsigned)z->n : (unsigned)z->n + 1u);
    for (unsigned counter = 1; counter <= (unsigned)z->n; ++counter) {
      const ef_factor *f = &z->f[counter - 1];
      if (!(counter > l_index && counter < r_index)) continue;
      { const int v[4] = {f->es + 1, f->ee + 1, R->gen->pref_N + f->gs + 1, R->gen->pref_N + f->ge + 1}; buf_ints(b, "", v, 4, ' ', " "); }
      const int el = f->ee + 1 - f->es, gl = f->ge + 1 - f->gs;
      if (el > 0) buf_write(b, e->orig + f->es, strnlen(e->orig + f->es, (size_t)el));
      buf_write(b, " ", 1);
      if (gl > 0) buf_write(b, R->gen_orig + R->gen->pref_N + f->gs, strnlen(R->gen_orig + R->gen->pref_N + f->gs, (size_t)gl));
      buf_write(b, "\n", 1);
    }
  }
}

/* compute_est_fact (src/compute-est-fact.c:192-293) for one strand copy; true when it produced factorizations */
static bool compute_est_fact(ef_task *T, const run_ctx *R, est_item *it, const ef_seq *e) {
  unsigned inc = 0;
  size_t prev_p = 0, prev_e = 0;
  for (;;) {
    size_t tp, te;
    ef_meg *M;
    const double t_meg0 = ef_now();
    for (;;) {
      ef_phase(EF_PH_MEG);
      M = meg_build(T, e, &inc);
      meg_stats(M, &tp, &te);
      const bool same = prev_p > 2 && prev_e > 0 && (prev_p <= tp || prev_e <= te);
      if (!same) break;
      ++inc;
    }
    prev_p = tp; prev_e = te;
    const double t_meg1 = ef_now();
    T->t_start = ef_task_ticks(T);
    bool timed_out = false;
    ef_phase(EF_PH_EMBED);
    ef_fzlist *L = est_factorizations(T, e, M, &timed_out);
    bool *polya = NULL, *polyad = NULL;
    if (L) {
      polya = ar_alloc(&T->ar, (size_t)L->n + 1); polyad = ar_alloc(&T->ar, (size_t)L->n + 1);
      for (int k = 0; k < L->n; ++k) { polya[k] = L->v[k]->polya; polyad[k] = L->v[k]->polyad; }
      ef_phase(EF_PH_REFINE);
      refine_factorizations(T, e, L);
    }
    ef_phase(EF_PH_OUTPUT);
    timed_out = timed_out || ef_timeout_expired(T);
    const double t_comp1 = ef_now();
    const bool got = L && L->n > 0;
    size_t meg_at = 0, meg_len = 0;
    if ((!timed_out || got) && R->cfg->aux_outputs) {
      buf_printf(&it->it_megs, "\n\n***********\n\n");
      meg_at = it->it_megs.len;
      write_est_record(&it->it_megs, e);
      meg_write(&it->it_megs, M);
      meg_len = it->it_megs.len - meg_at;
    }
    if (got) {
      if (R->cfg->aux_outputs) {
        buf_printf(&it->it_edges, ">%s\n", e->id);
        meg_write_edges(&it->it_edges, M);
        buf_write(&it->it_pmegs, it->it_megs.p + meg_at, meg_len);      /* processed-megs.txt: the same record and graph, formatted once */
      }
      buf_printf(&it->it_info, "%llu %llu %zu\n", (unsigned long long)((t_meg1 - t_meg0) * 1e6), (unsigned long long)((t_comp1 - t_meg1) * 1e6), (size_t)L->n);
      write_factorizations(&it->it_raw, R, e, L, polya, polyad);
      write_est_record(&it->it_pest, e);
      return true;
    }
    if (!timed_out) return false;
    ++inc;                       /* timed out without a result: retry with longer pairings */
  }
}

static void est_task_body(ef_task *T, run_ctx *R, est_item *it);
static void est_task(ef_task *T, size_t handle, void *user) {
  run_ctx *R = user;
  est_item *it = &R->win[handle >> WIN_BITS]->items[handle & (((size_t)1 << WIN_BITS) - 1)];
  est_task_body(T, R, it);
  atomic_store(&it->done, 1);
  if (atomic_load(&it->waited)) word_wake(&it->done);
}

/* One EST's records are 0.6-1.2 KB per file.  Six buffers of their own were six mallocs by the worker and six frees by six
 * different writer threads (each one into the worker's malloc arena: its lock, its cache lines); now they are slices of ONE
 * block that the last writer frees, and only a record that outgrows its slice gets storage of its own (buf_reserve). */
static const uint16_t OUT_SLICE[O_COUNT] = {2048, 1024, 1536, 1536, 1024, 64};
#define OUT_BLOCK (2048 + 1024 + 1536 + 1536 + 1024 + 64)

static void est_task_body(ef_task *T, run_ctx *R, est_item *it) {
  it->out_block = malloc(OUT_BLOCK);
  if (!it->out_block) { fprintf(stderr, "* FATAL est-fact: out of memory\n"); exit(1); }
  for (size_t k = 0, at = 0; k < O_COUNT; at += OUT_SLICE[k], ++k) {
    it->out[k].p = it->out_block + at; it->out[k].len = 0; it->out[k].cap = OUT_SLICE[k]; it->out[k].ext = true;
    it->out[k].p[0] = 0;
  }
  /* EST preparation (main-est-fact.c:190-213) */
  ef_set_gb(&it->fwd);
  ef_set_strand_and_rc(&it->fwd);
  it->fwd.len = (int)strlen(it->fwd.seq);
  ef_polyAT_substitution(&it->fwd);
  if (!it->fwd.fixed_strand) { ef_make_rc_copy(&it->fwd, &it->rc); it->rc.len = it->fwd.len; it->has_rc = true; }
  if (compute_est_fact(T, R, it, &it->fwd)) return;
  if (it->has_rc) {
    ar_reset(&T->ar);
    compute_est_fact(T, R, it, &it->rc);
  }
}

/* ---- the source of work: windows filled by the reader ---------------------------------------------------------- */
static int next_item(void *user, size_t *handle) {
  run_ctx *R = user;
  for (;;) {
    const size_t w = atomic_load(&R->cur);
    const int eof = atomic_load(&R->eof);                           /* read before n_ready: the reader sets it last */
    if (w >= atomic_load_explicit(&R->n_ready, memory_order_acquire)) return eof ? -1 : 0;
    window *W = R->win[w];
    const uint32_t k = atomic_fetch_add(&W->next, 1);
    if (k < W->n) { *handle = (w << WIN_BITS) | W->order[k]; return 1; }
    size_t expect = w;
    atomic_compare_exchange_strong(&R->cur, &expect, w + 1);      /* this window is handed out: on to the next */
  }
}

static void close_window(run_ctx *R, window *W) {
  /* longest first (counting sort on min(len / 64, 63), descending) */
  uint32_t cnt[65];
  memset(cnt, 0, sizeof cnt);
  W->order = malloc(sizeof(uint32_t) * (W->n ? W->n : 1));
  for (uint32_t i = 0; i < W->n; ++i) { size_t b = (size_t)W->items[i].fwd.len >> 6; if (b > 63) b = 63; ++cnt[63 - b + 1]; }
  for (int b = 0; b < 64; ++b) cnt[b + 1] += cnt[b];
  for (uint32_t i = 0; i < W->n; ++i) { size_t b = (size_t)W->items[i].fwd.len >> 6; if (b > 63) b = 63; W->order[cnt[63 - b]++] = i; }
  const size_t w = atomic_load(&R->n_ready);
  R->win[w] = W;
  atomic_fetch_add(&R->n_records, W->n);
  atomic_store_explicit(&R->n_ready, w + 1, memory_order_release);
  seq_bump(&R->ready_seq);
}

static void *reader_main(void *arg) {
  run_ctx *R = arg;
  ef_fasta *fa = ef_fasta_open("ests.txt");
  if (!fa) { fprintf(stderr, "* FATAL File ests.txt not found! Terminating\n"); R->rc = 1; atomic_store(&R->eof, 1); seq_bump(&R->ready_seq); return NULL; }
  window *W = NULL;
  size_t bytes = 0, cap = 0;
  ef_seq e;
  uint32_t win_records = WIN_RECORDS;
  size_t win_ahead = WIN_AHEAD;
  { const char *ov = getenv("EF_WINDOW"); if (ov && atol(ov) > 0 && atol(ov) < (1 << WIN_BITS)) win_records = (uint32_t)atol(ov); }     /* tests: tiny windows */
  { const char *ov = getenv("EF_WINDOW_AHEAD"); if (ov && atol(ov) > 0) win_ahead = (size_t)atol(ov); }
  /* the first windows are short (1/8, 1/4, 1/2 of a window), so that the workers start a few ms after the process does */
  uint32_t win_now = win_records >= 8 * 256 ? win_records / 8 : win_records;
  while (ef_fasta_next(fa, &e)) {
    if (!W) { W = calloc(1, sizeof *W); cap = 0; bytes = 0; }
    if (W->n == cap) { cap = cap ? cap * 2 : 256; W->items = realloc(W->items, cap * sizeof(est_item)); }
    memset(&W->items[W->n], 0, sizeof(est_item));
    W->items[W->n++].fwd = e;       /* strand / reverse-complement / polyA masking happen in est_task, on the worker threads */
    bytes += (size_t)e.len;
    if (W->n >= win_now || bytes >= WIN_BYTES) {
      if (win_now < win_records) win_now *= 2;
      if (atomic_load(&R->n_ready) + 1 >= MAX_WINDOWS) { fprintf(stderr, "* FATAL est-fact: too many input windows\n"); exit(1); }
      close_window(R, W);
      W = NULL;
      /* stay at most WIN_AHEAD windows in front of the slowest writer */
      for (;;) {
        const uint32_t seen = atomic_load(&R->written_seq);
        size_t slowest = (size_t)-1;
        for (int k = 0; k < O_COUNT; ++k) { const size_t v = atomic_load(&R->n_written[k]); if (v < slowest) slowest = v; }
        if (atomic_load(&R->n_ready) < slowest + win_ahead) break;
        word_wait(&R->written_seq, seen);
      }
    }
  }
  if (W) close_window(R, W);
  ef_fasta_close(fa);
  atomic_store(&R->eof, 1);
  seq_bump(&R->ready_seq);
  return NULL;
}

/* ---- writers: one per output file, each streams the records of its file in INPUT order while the workers run ----- */
static void item_release(est_item *it) {
  free(it->out_block);
  free(it->fwd.id); free(it->fwd.gb); free(it->fwd.seq); free(it->fwd.orig);
  if (it->has_rc) { free(it->rc.id); free(it->rc.gb); free(it->rc.seq); free(it->rc.orig); }
}

typedef struct writer_arg { run_ctx *R; int k; } writer_arg;
static void *writer_main(void *arg) {
  run_ctx *R = ((writer_arg *)arg)->R;
  const int k = ((writer_arg *)arg)->k;
  for (size_t w = 0;; ++w) {
    for (;;) {
      const uint32_t seen = atomic_load(&R->ready_seq);
      if (w < atomic_load_explicit(&R->n_ready, memory_order_acquire)) break;
      if (atomic_load(&R->eof) && w >= atomic_load(&R->n_ready)) return NULL;
      word_wait(&R->ready_seq, seen);
    }
    window *W = R->win[w];
    for (uint32_t i = 0; i < W->n; ++i) {
      est_item *it = &W->items[i];
      while (!atomic_load_explicit(&it->done, memory_order_acquire)) {
        atomic_store(&it->waited, 1);
        if (atomic_load(&it->done)) break;
        word_wait(&it->done, 0);
      }
      const double t0 = ef_now();
      ef_buf *b = &it->out[k];
      if (b->len) fwrite_unlocked(b->p, 1, b->len, R->f[k]);
      buf_free(b);
      if (atomic_fetch_add(&it->written, 1) == O_COUNT - 1) item_release(it);
      R->writer_busy[k] += ef_now() - t0;
    }
    if (atomic_fetch_add(&W->writers_done, 1) == O_COUNT - 1) { free(W->items); W->items = NULL; }   /* the struct itself stays: late next_item calls read n */
    atomic_store(&R->n_written[k], w + 1);
    seq_bump(&R->written_seq);
  }
}

static FILE *open_out(const char *name) {
  FILE *f = fopen(name, "w");
  if (!f) { fprintf(stderr, "* FATAL Cannot create file %s! Terminating\n", name); exit(1); }
  setvbuf(f, NULL, _IOFBF, (size_t)4 << 20);       /* one write() per 4 MB: the writer thread emits hundreds of MB record by record */
  return f;
}

int main(int argc, char **argv) {
  const double t0 = ef_now();
  /* load every kernel when the context is created (one thread) instead of lazily at first launch, when two dozen
   * worker threads would queue up behind the module loader */
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  ef_config cfg;
  if (ef_config_parse(&cfg, argc, argv)) return 1;
  /* --devices on a multi-GPU box: show the CUDA runtime only the GPUs this process will use (initialising the driver
   * for eight GPUs to use one costs most of a second), and renumber them 0..n-1 */
  if (cfg.n_devices > 0 && !getenv("CUDA_VISIBLE_DEVICES")) {       /* only an in-process engine ever initialises CUDA here */
    char list[16 * 12 + 1]; size_t at = 0;
    for (int i = 0; i < cfg.n_devices; ++i) at += (size_t)snprintf(list + at, sizeof list - at, i ? ",%d" : "%d", cfg.devices[i]);
    setenv("CUDA_VISIBLE_DEVICES", list, 1);
    setenv("EF_DEVICES_NARROWED", "1", 1);                          /* engine_client.c: in-process ordinals are 0..n-1 now */
  }
  if (!cfg.quiet) fprintf(stderr, "* INFO  EST-FACTORIZATION v2 (B200 build)\n");
  char name[64];
  snprintf(name, sizeof name, "info-pid-%u.log", (unsigned)getpid());
  FILE *finfo = open_out(name);
  fprintf(finfo, "start\t%ld\n", (long)time(NULL));

  const double t_io0 = ef_now();
  ef_seq *gens = NULL; size_t ngen = 0;
  if (ef_read_fasta("genomic.txt", &gens, &ngen)) { fprintf(stderr, "* FATAL File genomic.txt not found! Terminating\n"); return 1; }
  if (ngen != 1) { fprintf(stderr, "* FATAL genomic.txt must hold exactly one sequence (found %zu)\n", ngen); return 1; }
  ef_seq *gen = &gens[0];
  ef_parse_genomic_header(gen);
  ef_ntails_removal(gen);
  const double tl_genome = ef_now();
  sched_prepare(&cfg, gen);
  ef_small_exon_index_build(gen->seq, (size_t)gen->len);      /* host-side 6-mer index, while the engine session opens */
  const double tl_index = ef_now();
  static run_ctx R;
  R.cfg = &cfg; R.gen = gen; R.gen_orig = gen->orig;
  R.win = calloc(MAX_WINDOWS, sizeof(window *));              /* 32 MB of address space, touched as windows appear */
  if (access("ests.txt", R_OK) != 0) { fprintf(stderr, "* FATAL File ests.txt not found! Terminating\n"); return 1; }
  R.f[O_RAW] = open_out("raw-multifasta-out.txt"); R.f[O_MEGS] = open_out("megs.txt"); R.f[O_PMEGS] = open_out("processed-megs.txt");
  R.f[O_INFO] = open_out("processed-megs-info.txt"); R.f[O_PEST] = open_out("processed-ests.txt"); R.f[O_EDGES] = open_out("meg-edges.txt");
  double t_io = ef_now() - t_io0;
  pthread_t rth, wth[O_COUNT];
  writer_arg wa[O_COUNT];
  if (pthread_create(&rth, NULL, reader_main, &R)) { perror("pthread_create"); return 1; }
  for (int k = 0; k < O_COUNT; ++k) { wa[k].R = &R; wa[k].k = k; if (pthread_create(&wth[k], NULL, writer_main, &wa[k])) { perror("pthread_create"); return 1; } }
  /* short inputs are known in full before the first window closes: they get fewer threads and fibers (sched_run) */
  for (;;) {
    const uint32_t seen = atomic_load(&R.ready_seq);
    if (atomic_load(&R.eof) || atomic_load(&R.n_ready) != 0) break;
    word_wait(&R.ready_seq, seen);
  }
  const size_t n_hint = atomic_load(&R.eof) ? atomic_load(&R.n_records) : SIZE_MAX;
  const double tl_ests = ef_now();
  const double t_alg0 = ef_now();
  if (sched_run(&cfg, gen, n_hint, next_item, est_task, &R)) return 1;
  const double t_alg = ef_now() - t_alg0;
  pthread_join(rth, NULL);
  if (R.rc) return 1;
  const size_t nest = atomic_load(&R.n_records);
  if (!cfg.quiet) fprintf(stderr, "* INFO  Read %zu sequences.\n", nest);
  if (!cfg.quiet)
    fprintf(stderr, "* INFO  timeline (s since start): genome read %.3f, small-exon index %.3f, first window / end of ests.txt %.3f, workers done %.3f\n",
            tl_genome - t0, tl_index - t0, tl_ests - t0, t_alg0 + t_alg - t0);

  const double t_io1 = ef_now();
  double wbusy = 0;
  double tl_join[O_COUNT], tl_close[O_COUNT];
  for (int k = 0; k < O_COUNT; ++k) { pthread_join(wth[k], NULL); tl_join[k] = ef_now(); fclose(R.f[k]); tl_close[k] = ef_now(); if (R.writer_busy[k] > wbusy) wbusy = R.writer_busy[k]; }
  t_io += ef_now() - t_io1 + wbusy;
  if (!cfg.quiet) {
    fprintf(stderr, "* INFO  tail (s after workers done), writer joined / file closed:");
    for (int k = 0; k < O_COUNT; ++k) fprintf(stderr, " %.3f/%.3f", tl_join[k] - t_io1, tl_close[k] - t_io1);
    fprintf(stderr, "; writer busy (s):");
    for (int k = 0; k < O_COUNT; ++k) fprintf(stderr, " %.3f", R.writer_busy[k]);
    fprintf(stderr, "\n");
  }
  fprintf(finfo, "end\t%ld\n", (long)time(NULL));
  fclose(finfo);

  double gpu_wait; uint64_t batches, jobs;
  sched_stats(&gpu_wait, &batches, &jobs);
  const double t_tot = ef_now() - t0;
  /* the five timers of the reference (main-est-fact.c:321-325); index build and per-EST work both live in "Algorithm" */
  fprintf(stderr, "@Timer Suffix Tree. Time elapsed: %llu microsec\n", 0ull);
  fprintf(stderr, "@Timer Algorithm. Time elapsed: %llu microsec\n", (unsigned long long)(t_alg * 1e6));
  fprintf(stderr, "@Timer Compositions. Time elapsed: %llu microsec\n", 0ull);
  fprintf(stderr, "@Timer IO. Time elapsed: %llu microsec\n", (unsigned long long)(t_io * 1e6));
  fprintf(stderr, "@Timer Total. Time elapsed: %llu microsec\n", (unsigned long long)(t_tot * 1e6));
  double tfib, tgat, tsub;
  sched_breakdown(&tfib, &tgat, &tsub);
  if (!cfg.quiet)
    fprintf(stderr, "* INFO  host thread-seconds: per-EST code %.3f, batch gather %.3f, submit %.3f, wait on device %.3f\n", tfib, tgat, tsub, gpu_wait);
  if (!cfg.quiet) {
    static const char *nm[EF_PH_COUNT] = {"other", "vertex-set", "meg", "embeddings", "candidates", "filters", "intron-refine", "fact-refine", "small-exons", "output"};
    const double *ph = sched_phase_seconds();
    fprintf(stderr, "* INFO  per-EST code by phase (s):");
    for (int i = 0; i < EF_PH_COUNT; ++i) fprintf(stderr, " %s %.3f", nm[i], ph[i]);
    fprintf(stderr, "\n");
    uint64_t ymax = 0;
    const uint64_t *py = sched_phase_yields(&ymax);
    fprintf(stderr, "* INFO  engine round trips by phase:");
    for (int i = 0; i < EF_PH_COUNT; ++i) if (py[i]) fprintf(stderr, " %s %llu", nm[i], (unsigned long long)py[i]);
    fprintf(stderr, "; longest chain of one EST: %llu\n", (unsigned long long)ymax);
  }
  uint64_t h2d, d2h;
  sched_bytes(&h2d, &d2h);
  if (!cfg.quiet) fprintf(stderr, "* INFO  bytes host->device: %llu, device->host: %llu\n", (unsigned long long)h2d, (unsigned long long)d2h);
  pc_session_stats es; const char *emode;
  sched_engine_stats(&es, &emode);
  if (!cfg.quiet)
    fprintf(stderr, "* INFO  lane batches: %llu, device jobs: %llu, kernel launches: %llu, summed wait on device: %.3f s, ESTs/s: %.1f\n",
            (unsigned long long)batches, (unsigned long long)jobs, (unsigned long long)es.launches, gpu_wait,
            t_alg > 0 ? (double)nest / t_alg : 0.0);
  if (!cfg.quiet)
    fprintf(stderr, "* INFO  engine (%s): %llu merged device batches from %llu lane batches (%.1f lanes per batch), %llu retry rounds, busy %.3f s\n",
            emode, (unsigned long long)es.batches, (unsigned long long)es.lanes_merged, es.batches ? (double)es.lanes_merged / (double)es.batches : 0.0,
            (unsigned long long)es.retries, es.busy_s);
  if (!cfg.quiet && getenv("PC_PROFILE")) {
    static const char *on[PC_OP_COUNT] = {"ALIGN", "KBAND", "EDIT", "BORDERS", "GAP", "AFFIX", "SUFCUT", "PRECUT", "LCS", "SEED"};
    fprintf(stderr, "* INFO  engine device ms per op:");
    for (int o = 0; o < PC_OP_COUNT; ++o) if (es.op_ms[o] > 0) fprintf(stderr, " %s %.1f", on[o], es.op_ms[o]);
    fprintf(stderr, "\n");
  }
  if (!strcmp(emode, "in-process")) pc_debug_dump();
  fflush(NULL);
#ifdef EF_GPROF
  exit(0);         /* profiling build: let gmon.out be written */
#endif
  _exit(0);        /* every output file is closed; skip the CUDA runtime's exit-time tear-down */
}
