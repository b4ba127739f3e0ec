/* sched.c — the EST batcher: worker threads, fibers, and the GPU batch round-trip.
 *
 * Replaces the sequential hot loop of the reference (src/main-est-fact.c:249-291): every EST runs the same
 * sequential per-EST code (compute_est_fact, src/compute-est-fact.c:192) on its own fiber; whenever that code needs
 * a DP it queues jobs and yields.  A worker thread owns two groups of fibers (EF_GROUPS: up to four) and one engine LANE per group
 * (include/pintron_engine.h): while the jobs of group A are with the engine it runs the fibers of group B, then swaps.
 * The engine — inside this process or in the resident server est-factd — merges the lanes of all threads that are posted
 * at the same moment into one device batch; no worker thread calls CUDA.  ESTs are dealt to threads from a shared
 * counter; threads are spread over the configured GPUs (genome + index replicated per GPU, no collective: SURVEY.md §8(e)).
 */
#define _GNU_SOURCE
#include "ef.h"
#include "engine_client.h"
#include <errno.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdatomic.h>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/time.h>
#include <ucontext.h>

/* Fiber switch.  glibc's swapcontext saves and restores the signal mask with a system call on every switch; an EST
 * yields about a hundred times, so on x86-64 the switch is done by hand: callee-saved registers + stack pointer. */
#if defined(__x86_64__)
#define EF_FAST_SWITCH 1
__attribute__((naked, noinline)) static void ctx_switch(void **save_sp, void *load_sp) {
  __asm__ volatile(
      "pushq %rbp\n\tpushq %rbx\n\tpushq %r12\n\tpushq %r13\n\tpushq %r14\n\tpushq %r15\n\t"
      "movq %rsp, (%rdi)\n\t"
      "movq %rsi, %rsp\n\t"
      "popq %r15\n\tpopq %r14\n\tpopq %r13\n\tpopq %r12\n\tpopq %rbx\n\tpopq %rbp\n\t"
      "ret\n");
}
#else
#define EF_FAST_SWITCH 0
#endif
#include <unistd.h>
#include <malloc.h>

#define FIBER_STACK (512u << 10)
#define EF_MAX_GROUPS 4

double ef_now(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

/* Phase accounting and the per-EST timeout read the clock twice per yield (a hundred yields per EST): the TSC costs a
 * third of a vDSO gettimeofday.  Its rate is measured against the wall clock over the life of the process. */
#if defined(__x86_64__)
#include <x86intrin.h>
uint64_t ef_ticks(void) { return __rdtsc(); }
#else
uint64_t ef_ticks(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec; }
#endif
static uint64_t g_tick0;
static double g_wall0;
__attribute__((constructor)) static void ticks_init(void) { g_wall0 = ef_now(); g_tick0 = ef_ticks(); }
double ef_ticks_to_s(uint64_t ticks) {
  double dt = ef_now() - g_wall0;
  if (dt < 0.002) { struct timespec ts = {0, 2000000}; nanosleep(&ts, NULL); dt = ef_now() - g_wall0; }      /* too early for a rate: wait 2 ms */
  const double rate = (double)(ef_ticks() - g_tick0) / dt;
  return (double)ticks / rate;
}
uint64_t ef_task_ticks(const ef_task *T) { return T->run_ticks + (ef_ticks() - T->run_mark); }

/* ---- arena / buffers ------------------------------------------------------------------------------- */
/* Chunks are kept across ESTs (ar_reset only rewinds): giving big chunks back to the system means munmap, and
 * every munmap interrupts all threads of the process to flush their TLBs — with a dozen workers handling long mRNAs
 * that alone was most of the run time.  head = the chunk being filled; chunks after it on the list are spare. */
#define AR_HDR (((sizeof(ef_chunk)) + 15u) & ~(size_t)15u)
void *ar_alloc_raw(ef_arena *a, size_t bytes) {
  bytes = (bytes + 15u) & ~(size_t)15u;
  ef_chunk *c = a->cur;
  while (!c || c->used + bytes > c->cap) {
    ef_chunk *nx = c ? c->next : a->head;
    if (nx && AR_HDR + bytes <= nx->cap) { nx->used = AR_HDR; a->cur = c = nx; continue; }
    size_t cap = 1u << 16;
    if (c && c->cap * 2 > cap) cap = MIN2(c->cap * 2, (size_t)8 << 20);
    if (cap < bytes + AR_HDR + 16) cap = bytes + AR_HDR + 16;
    ef_chunk *n = malloc(cap);
    if (!n) { fprintf(stderr, "* FATAL est-fact: out of memory\n"); exit(1); }
    n->cap = cap; n->used = AR_HDR;
    if (c) { n->next = c->next; c->next = n; } else { n->next = a->head; a->head = n; }
    a->cur = c = n;
  }
  void *p = (char *)c + c->used;
  c->used += bytes;
  return p;
}

void *ar_alloc(ef_arena *a, size_t bytes) { return memset(ar_alloc_raw(a, bytes), 0, bytes); }

/* give back the most recent allocation(s): everything from p on, if p lies in the chunk being filled */
void ar_undo(ef_arena *a, void *p) {
  ef_chunk *c = a->cur;
  if (c && (char *)p >= (char *)c + AR_HDR && (char *)p < (char *)c + c->used) c->used = (size_t)((char *)p - (char *)c);
}

void ar_reset(ef_arena *a) {
  /* rewind; keep at most 64 MB of chunks for the next EST */
  size_t kept = 0;
  ef_chunk **pp = &a->head;
  while (*pp) {
    ef_chunk *c = *pp;
    if (kept + c->cap > ((size_t)64 << 20) && c != a->head) { *pp = c->next; free(c); continue; }
    kept += c->cap; c->used = AR_HDR;
    pp = &c->next;
  }
  a->cur = a->head;
}

void ar_free_all(ef_arena *a) {
  while (a->head) { ef_chunk *n = a->head->next; free(a->head); a->head = n; }
  a->cur = NULL;
}

static void buf_reserve(ef_buf *b, size_t extra) {
  if (b->len + extra + 1 <= b->cap) return;
  size_t cap = b->cap ? b->cap * 2 : 2048;      /* one EST's records are a few KB per file: start there, not at 256 bytes */
  while (cap < b->len + extra + 1) cap *= 2;
  if (b->ext) {                                 /* outgrew its slice of the EST's output block: move to storage of its own */
    char *q = malloc(cap);
    if (q && b->len) memcpy(q, b->p, b->len + 1);
    b->p = q; b->ext = false;
  } else b->p = realloc(b->p, cap);
  if (!b->p) { fprintf(stderr, "* FATAL est-fact: out of memory\n"); exit(1); }
  b->cap = cap;
}

void buf_write(ef_buf *b, const void *src, size_t n) {
  buf_reserve(b, n);
  memcpy(b->p + b->len, src, n);
  b->len += n;
  b->p[b->len] = 0;
}

void buf_printf(ef_buf *b, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  char tmp[256];
  int n = vsnprintf(tmp, sizeof tmp, fmt, ap);
  va_end(ap);
  if (n < (int)sizeof tmp) { buf_write(b, tmp, (size_t)n); return; }
  buf_reserve(b, (size_t)n);
  va_start(ap, fmt);
  vsnprintf(b->p + b->len, (size_t)n + 1, fmt, ap);
  va_end(ap);
  b->len += (size_t)n;
}

/* decimal integers separated / terminated by single characters, without going through printf */
void buf_ints(ef_buf *b, const char *open, const int *v, int n, char sep, const char *close) {
  buf_reserve(b, (size_t)n * 12 + 16);
  char *p = b->p + b->len;
  while (*open) *p++ = *open++;
  for (int i = 0; i < n; ++i) {
    if (i) *p++ = sep;
    long long x = v[i];
    if (x < 0) { *p++ = '-'; x = -x; }
    char tmp[12]; int k = 0;
    do { tmp[k++] = (char)('0' + x % 10); x /= 10; } while (x);
    while (k) *p++ = tmp[--k];
  }
  while (*close) *p++ = *close++;
  *p = 0;
  b->len = (size_t)(p - b->p);
}

void buf_free(ef_buf *b) { if (!b->ext) free(b->p); b->p = NULL; b->len = b->cap = 0; b->ext = false; }

/* ---- fibers ---------------------------------------------------------------------------------------- */
enum { F_FREE = 0, F_RUNNABLE, F_WAITING, F_DONE, F_JOINING };      /* JOINING: waits for the child fibers of dp_parallel_for */

typedef struct ef_req { int op; ef_str a, b; int p0, p1, p2, out_cap; } ef_req;

typedef struct fiber {
  ucontext_t ctx;
  void *sp;                 /* saved stack pointer (fast switch) */
  void *stack;
  int state;
  size_t index;
  ef_task task;
  ef_req *reqs;
  int nreq, capreq;
  int base;                 /* index of this fiber's first job in the group's batch */
  int phase;
  bool has_results;
  size_t need_a, need_v;    /* staging bytes of the pending requests (valid while need_valid) */
  bool need_valid;
  bool submitted;           /* its requests are part of the batch in flight (false: deferred to the next one) */
  uint64_t yields;          /* round trips to the engine of the EST on this fiber */
  /* child fibers (dp_parallel_for): a child runs one iteration of its parent's loop on the parent's task */
  struct fiber *parent;
  int pending_children;
  ef_par_fn par_fn; void *par_user; int par_k; ef_task *par_T;
  struct group *grp;
} fiber;

typedef struct group {
  ef_conn *conn;              /* the engine session of this thread's GPU */
  int lane_k;                 /* which of the session's lanes is ours */
  fiber *fibers;
  int nfibers;                /* slots that take ESTs; slots [nfibers, nslots) are child fibers, started by dp_parallel_for only */
  int nslots;
  bool rerun;                 /* a fiber behind the cursor of the current pass became runnable */
  bool pending;
  /* batch buffers = the lane's slab in the engine's pinned (shared) memory; refreshed by lane_refresh */
  uint8_t *arena; size_t arena_cap, arena_len;
  pc_job *jobs; int jobs_cap, njobs;
  int32_t *res;
  uint8_t *var; size_t var_cap, var_len;
  struct worker *w;
  uint64_t deferred, grows;   /* fibers put off to a later batch; lane re-allocations */
  bool want_grow;             /* the last batch had to leave many fibers behind: take a larger slab before the next one */
  void *stacks;               /* nslots x FIBER_STACK, one mapping (fiber_prepare_stack) */
} group;

typedef struct worker {
  pthread_t th;
  int id, device;
  group g[EF_MAX_GROUPS];     /* g_ngroups of them in use */
  ucontext_t main_ctx;
  void *main_sp;
  const ef_config *cfg;
  const ef_seq *gen;
  ef_task_fn fn;
  void *user;
  int inflight;             /* ESTs started and not finished, both groups */
  uint64_t batches, jobs, h2d, d2h;
  double gpu_wait, t_fibers, t_gather, t_submit;
} worker;

static ef_next_fn g_next;            /* where items come from (main.c: the windows the FASTA reader has filled) */
static void *g_next_user;
static _Atomic int g_no_more;         /* the source has said -1 */
static int g_nthreads = 1;
/* Fiber groups (= engine lanes) per worker thread.  While the requests of one group are with the engine the thread runs the
 * others; with G groups a batch has the run time of G - 1 groups to come back before the thread would wait for it, at the
 * same number of suspended ESTs per thread (the cache footprint that decides the cost per EST).  EF_GROUPS=2..4. */
#define EF_GROUPS_DEFAULT 2
static int g_ngroups = EF_GROUPS_DEFAULT;
static __thread fiber *tl_fiber;
static __thread worker *tl_worker;
static uint64_t g_batches, g_jobs, g_h2d, g_d2h, g_deferred, g_grows;
static double g_gpu_wait, g_t_fibers, g_t_gather, g_t_submit, g_t_init, g_t_fini, g_t_end_sum, g_t_end_min;
static pthread_mutex_t g_stat_mu = PTHREAD_MUTEX_INITIALIZER;

/* ---- where the host time goes: CPU seconds per phase of the per-EST code (fibers only run between yields) ------ */
static __thread int tl_phase;
static __thread uint64_t tl_mark;
static __thread uint64_t tl_phase_s[EF_PH_COUNT];      /* ticks */
static double g_phase_s[EF_PH_COUNT];
static __thread uint64_t tl_phase_yields[EF_PH_COUNT];      /* dp_wait round trips per phase */
static uint64_t g_phase_yields[EF_PH_COUNT], g_max_yields;
static __thread uint64_t tl_max_yields;
static inline void phase_account(void) {
  const uint64_t t = ef_ticks();
  tl_phase_s[tl_phase] += t - tl_mark;
  tl_mark = t;
}
int ef_phase(int ph) { phase_account(); const int old = tl_phase; tl_phase = ph; return old; }
const double *sched_phase_seconds(void) { return g_phase_s; }
const uint64_t *sched_phase_yields(uint64_t *max_per_est) { if (max_per_est) *max_per_est = g_max_yields; return g_phase_yields; }

/* the group's view of its lane: pointers into the (shared, pinned) segment that holds the slab */
static void lane_refresh(group *g) {
  pce_lane *l = efc_lane(g->conn, g->lane_k);
  uint8_t *base = efc_seg(g->conn, l->seg);
  if (!base) { fprintf(stderr, "* FATAL est-fact: cannot map lane segment %u of the engine\n", l->seg); exit(1); }
  g->arena = base + l->arena_off; g->arena_cap = (size_t)l->arena_cap;
  g->jobs = (pc_job *)(base + l->jobs_off); g->jobs_cap = (int)l->jobs_cap;
  g->res = (int32_t *)(base + l->res_off);
  g->var = base + l->var_off; g->var_cap = (size_t)l->var_cap;
}
/* a fiber's requests outgrew the lane: the engine moves it to a larger slab (what is already gathered travels along) */
static void lane_grow(group *g, size_t arena_cap, int jobs_cap, size_t var_cap) {
  if (efc_resize(g->conn, g->lane_k, arena_cap, (uint32_t)jobs_cap, var_cap, g->arena_len, (uint32_t)g->njobs)) {
    fprintf(stderr, "* FATAL est-fact: the engine could not grow a lane to %zu + %zu bytes, %d jobs: %s\n", arena_cap, var_cap, jobs_cap, pc_last_error());
    exit(1);
  }
  ++g->grows;
  lane_refresh(g);
}

#define EF_STACK_CANARY 0x5ca1ab1e0ddba11ull
static inline void fiber_check_stack(const fiber *f) {
  if (f->stack && *(const uint64_t *)f->stack != EF_STACK_CANARY) {
    fprintf(stderr, "* FATAL est-fact: a fiber overran its %u KB stack\n", FIBER_STACK >> 10);
    abort();
  }
}
#define EF_PF_FAR 8
#define EF_PF_NEAR 3
static inline void fiber_prefetch(const group *g, const fiber *n) {
  if (n->state != F_RUNNABLE || !n->sp) return;
  if (n->has_results && n->submitted && n->nreq) {                         /* its results: DMA-written lines of the lane, in no cache */
    const char *r = (const char *)&g->res[(size_t)n->base * PC_RES_INTS];
    const size_t bytes = MIN2((size_t)n->nreq * PC_RES_INTS * sizeof(int32_t), (size_t)512);
    for (size_t o = 0; o < bytes; o += 64) __builtin_prefetch(r + o);
  }
  const char *sp = (const char *)n->sp;
  for (int i = 0; i < 16; ++i) __builtin_prefetch(sp + 64 * i, 1);       /* the frames it resumes into (dp_wait and its callers) */
  __builtin_prefetch(&n->task, 1);
  __builtin_prefetch((const char *)&n->task + 64, 1);
  if (n->reqs) __builtin_prefetch(n->reqs);
  if (n->task.ar.cur) __builtin_prefetch(n->task.ar.cur, 1);
}

static void fiber_entry(void) {
  fiber *f = tl_fiber;
  worker *w = tl_worker;
  tl_phase = EF_PH_OTHER; tl_mark = ef_ticks();
  f->task.run_ticks = 0; f->task.run_mark = tl_mark;
  w->fn(&f->task, f->index, w->user);
  phase_account();
  f->state = F_DONE;
#if EF_FAST_SWITCH
  ctx_switch(&f->sp, w->main_sp);
#else
  swapcontext(&f->ctx, &w->main_ctx);
#endif
  __builtin_unreachable();
}

static void fiber_prepare_stack(group *g, fiber *f, void (*entry)(void));
static void fiber_start(worker *w, group *g, fiber *f, size_t index) {
  fiber_prepare_stack(g, f, fiber_entry);
  f->index = index;
  f->yields = 0;
  f->state = F_RUNNABLE;
  f->nreq = 0;
  f->has_results = false;
  f->grp = g;
  f->parent = NULL; f->pending_children = 0;
  ar_reset(&f->task.ar);
  f->task.cfg = w->cfg;
  f->task.gen = w->gen;
}

/* ---- parallel loops over child fibers ----------------------------------------------------------------------------
 * A read with many candidate factorizations refines them one after the other in the reference, and every refinement is a
 * chain of dependent DP calls: on the device that is thousands of sequential round trips for ONE read (11 879 for the
 * worst mRNA of the C4 sample) while the rest of the machine waits.  Iterations that touch disjoint data run here as
 * child fibers of the same group, so their round trips overlap; the iteration bodies themselves are unchanged and the
 * result does not depend on the interleaving.  Without a free child slot the iteration simply runs inline. */
static void child_entry(void) {
  fiber *f = tl_fiber;
  worker *w = tl_worker;
  tl_mark = ef_ticks();
  f->par_fn(f->par_T, f->par_k, f->par_user);
  phase_account();
  fiber *p = f->parent;
  if (--p->pending_children == 0 && p->state == F_JOINING) { p->state = F_RUNNABLE; p->grp->rerun = true; }
  f->state = F_DONE;
#if EF_FAST_SWITCH
  ctx_switch(&f->sp, w->main_sp);
#else
  swapcontext(&f->ctx, &w->main_ctx);
#endif
  __builtin_unreachable();
}

static void fiber_prepare_stack(group *g, fiber *f, void (*entry)(void)) {
  if (!f->stack) {
    /* The stacks of a group are slices of ONE mapping (pages appear as they are touched).  One mmap + one mprotect per
     * fiber meant ~10^5 VMAs per process: every one of those calls takes the process-wide mmap lock while sixteen
     * workers ramp up, and the kernel walks them all again at exit.  Overflow is caught by a canary word at the low end
     * of each slice, checked when the fiber's EST is done (fiber_check_stack). */
    if (!g->stacks) {
      g->stacks = mmap(NULL, (size_t)g->nslots * FIBER_STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (g->stacks == MAP_FAILED) { perror("mmap fiber stacks"); exit(1); }
      madvise(g->stacks, (size_t)g->nslots * FIBER_STACK, MADV_NOHUGEPAGE);      /* a 2 MB page per touched stack would be zeroed in full */
    }
    f->stack = (char *)g->stacks + (size_t)(f - g->fibers) * FIBER_STACK;
    *(uint64_t *)f->stack = EF_STACK_CANARY;
  }
#if EF_FAST_SWITCH
  {
    uintptr_t *sp = (uintptr_t *)(((uintptr_t)f->stack + FIBER_STACK) & ~(uintptr_t)15);
    *--sp = 0;                          /* the return address the entry never uses (keeps the ABI stack alignment) */
    *--sp = (uintptr_t)entry;           /* where the first switch "returns" to */
    for (int r = 0; r < 6; ++r) *--sp = 0;   /* rbp rbx r12 r13 r14 r15 */
    f->sp = sp;
  }
#else
  getcontext(&f->ctx);
  f->ctx.uc_stack.ss_sp = f->stack;
  f->ctx.uc_stack.ss_size = FIBER_STACK;
  f->ctx.uc_link = NULL;
  makecontext(&f->ctx, entry, 0);
#endif
}

void dp_parallel_for(ef_task *T, int n, ef_par_fn body, void *user) {
  fiber *f = tl_fiber;
  group *g = f->grp;
  if (f->nreq && !f->has_results) dp_wait();                 /* nothing of the caller's may be in flight while it joins */
  int next_slot = g->nfibers;
  for (int k = 0; k < n; ++k) {
    fiber *c = NULL;
    if (n > 1)
      for (; next_slot < g->nslots; ++next_slot)
        if (g->fibers[next_slot].state == F_FREE) { c = &g->fibers[next_slot++]; break; }
    if (!c) { body(T, k, user); continue; }                  /* no child slot left (or a single iteration): inline */
    fiber_prepare_stack(g, c, child_entry);
    c->index = f->index; c->yields = 0; c->nreq = 0; c->has_results = false; c->grp = g;
    c->task.cfg = T->cfg; c->task.gen = T->gen;              /* dp_push looks at these; everything else goes through the parent's T */
    c->parent = f; c->pending_children = 0; c->par_fn = body; c->par_user = user; c->par_k = k; c->par_T = T;
    c->phase = tl_phase;
    c->state = F_RUNNABLE;
    ++f->pending_children;
  }
  if (f->pending_children == 0) return;
  f->state = F_JOINING;
  phase_account();
  f->task.run_ticks += tl_mark - f->task.run_mark;
  f->phase = tl_phase;
#if EF_FAST_SWITCH
  ctx_switch(&f->sp, tl_worker->main_sp);
#else
  swapcontext(&f->ctx, &tl_worker->main_ctx);
#endif
  tl_phase = f->phase; tl_mark = ef_ticks();
  f->task.run_mark = tl_mark;
}

int dp_push(int op, ef_str a, ef_str b, int p0, int p1, int p2, int out_cap) {
  fiber *f = tl_fiber;
  if (f->has_results) { f->nreq = 0; f->has_results = false; }
  if (f->nreq == f->capreq) {
    f->capreq = f->capreq ? f->capreq * 2 : 16;
    f->reqs = realloc(f->reqs, sizeof(ef_req) * (size_t)f->capreq);
  }
  ef_req *r = &f->reqs[f->nreq];
  /* a genome-side string that lies inside the genome travels as (offset, length) into the device-resident copy */
  const ef_seq *gen = f->task.gen;
  if (!b.in_genome && b.p && b.p >= gen->seq && b.p + b.len <= gen->seq + gen->len) { b.in_genome = true; b.gen_off = (int)(b.p - gen->seq); }
  r->op = op; r->a = a; r->b = b; r->p0 = p0; r->p1 = p1; r->p2 = p2; r->out_cap = out_cap;
  return f->nreq++;
}

void dp_wait(void) {
  fiber *f = tl_fiber;
  if (f->nreq == 0 || f->has_results) return;
  f->state = F_WAITING;
  f->need_valid = false;
  phase_account();
  ++tl_phase_yields[tl_phase];
  if (++f->yields > tl_max_yields) tl_max_yields = f->yields;
  f->task.run_ticks += tl_mark - f->task.run_mark;
  f->phase = tl_phase;
#if EF_FAST_SWITCH
  ctx_switch(&f->sp, tl_worker->main_sp);
#else
  swapcontext(&f->ctx, &tl_worker->main_ctx);
#endif
  /* resumed: results are in the group's batch buffers */
  tl_phase = f->phase; tl_mark = ef_ticks();
  f->task.run_mark = tl_mark;
  for (int i = 0; i < f->nreq; ++i) {
    const int32_t st = f->grp->res[(size_t)(f->base + i) * PC_RES_INTS];
    if (st < 0 && st != PC_E_OUTCAP) {
      fprintf(stderr, "* FATAL est-fact: device job (op %d) failed with status %d\n", f->reqs[i].op, st);
      exit(1);
    }
  }
}

const int32_t *dp_res(int h) { fiber *f = tl_fiber; return f->grp->res + (size_t)(f->base + h) * PC_RES_INTS; }
const uint8_t *dp_var(int h) { fiber *f = tl_fiber; return f->grp->var + f->grp->jobs[f->base + h].out_off; }

unsigned dp_edit(const char *a, int la, const char *b, int lb) {
  int h = dp_push(PC_OP_EDIT, S_(a, la), S_(b, lb), 0, 0, 0, 0);
  dp_wait();
  return (unsigned)dp_res(h)[1];
}

bool dp_borders(const char *p, int len_p, int min_cut, int max_cut, const char *t, int len_t, unsigned max_errs,
                int *off_p, int *off_t1, int *off_t2, unsigned *ed) {
  int h = dp_push(PC_OP_BORDERS, S_(p, len_p), S_(t, len_t), (int)max_errs, min_cut, max_cut, 0);
  dp_wait();
  const int32_t *r = dp_res(h);
  *off_p = r[2]; *off_t1 = r[3]; *off_t2 = r[4]; *ed = (unsigned)r[5];
  return r[1] != 0;
}

void dp_lcs(const char *s1, long l1, const char *s2, long l2, long *occ1, long *occ2, long *len) {
  int h = dp_push(PC_OP_LCS, S_(s2, (int)l2), S_(s1, (int)l1), 0, 0, 0, 0);
  dp_wait();
  const int32_t *r = dp_res(h);
  *len = r[1]; *occ1 = r[2]; *occ2 = r[3];
}

ef_aln aln_from_ops(ef_task *T, const uint8_t *ops, int n, const char *est, const char *gen) {
  ef_aln A;
  char *m = ar_alloc(&T->ar, 2 * ((size_t)n + 48));
  A.est = m + 16; A.gen = m + n + 48 + 16; A.dim = n; A.score = 0;
  int i = 0, j = 0;
  for (int k = 0; k < n; ++k) {
    if (ops[k] == 0) { A.est[k] = est[i++]; A.gen[k] = gen[j++]; }
    else if (ops[k] == 1) { A.est[k] = est[i++]; A.gen[k] = '-'; }
    else { A.est[k] = '-'; A.gen[k] = gen[j++]; }
  }
  return A;
}

/* ---- batching -------------------------------------------------------------------------------------------- */
#define LANE_MAX_BYTES ((size_t)32 << 20)
/* The strings of a DP request are mostly 10-100 bytes; two libc memcpy calls per job, 350 per EST, cost their call and size
 * dispatch more than the copy.  Sizes up to 64 go through overlapping fixed-size moves (never a byte outside [src, src+n)). */
static inline void copy_small(uint8_t *dst, const char *src, size_t n) {
  if (n > 64) { memcpy(dst, src, n); return; }
  if (n >= 16) {
    typedef struct { uint64_t a, b; } __attribute__((packed, may_alias)) v16;
    if (n > 32) { *(v16 *)(dst + 16) = *(const v16 *)(src + 16); *(v16 *)(dst + n - 32) = *(const v16 *)(src + n - 32); }
    *(v16 *)dst = *(const v16 *)src;
    *(v16 *)(dst + n - 16) = *(const v16 *)(src + n - 16);
    return;
  }
  if (n >= 8) { uint64_t a, b; memcpy(&a, src, 8); memcpy(&b, src + n - 8, 8); memcpy(dst, &a, 8); memcpy(dst + n - 8, &b, 8); return; }
  if (n >= 4) { uint32_t a, b; memcpy(&a, src, 4); memcpy(&b, src + n - 4, 4); memcpy(dst, &a, 4); memcpy(dst + n - 4, &b, 4); return; }
  for (size_t i = 0; i < n; ++i) dst[i] = (uint8_t)src[i];
}

static void gather(group *g) {
  g->arena_len = 0; g->njobs = 0; g->var_len = 0;
  /* Reads with thousands of candidate alignments (mRNAs) fill a lane with a handful of fibers; batches of a handful of
   * fibers keep neither the engine nor the workers busy.  When a batch had to defer more than a quarter of the waiting
   * fibers, the lane doubles (up to 32 MB per buffer) before the next batch is gathered. */
  if (g->want_grow) {
    g->want_grow = false;
    if (g->arena_cap < LANE_MAX_BYTES && !getenv("EF_STAGING_KB"))
      lane_grow(g, MIN2(g->arena_cap * 2, LANE_MAX_BYTES), MIN2(g->jobs_cap * 2, 1 << 20), MIN2(g->var_cap * 2, LANE_MAX_BYTES));
  }
  int waiting_now = 0, deferred_now = 0;
  for (int k = 0; k < g->nslots; ++k) {
    fiber *f = &g->fibers[k];
    if (f->state != F_WAITING) continue;
    ++waiting_now;
    /* Back-pressure instead of growth: the lane keeps the size it was given at start-up; a fiber whose requests do not
     * fit any more waits for the next batch.  Only a single fiber that is larger than an EMPTY lane makes the engine
     * move the lane to a larger slab. */
    if (!f->need_valid) {
      size_t need_a = 0, need_v = 0;
      for (int i = 0; i < f->nreq; ++i) {
        const ef_req *r = &f->reqs[i];
        need_a += (size_t)r->a.len + (size_t)(r->b.in_genome ? 0 : r->b.len) + 8;
        if (r->op == PC_OP_ALIGN || r->op == PC_OP_GAP) need_v += (size_t)r->a.len + (size_t)r->b.len;
        else if (r->op == PC_OP_SEED) need_v += 12u * (size_t)r->out_cap + 4;
      }
      f->need_a = need_a; f->need_v = need_v; f->need_valid = true;
    }
    if (g->njobs > 0 && (g->arena_len + f->need_a > g->arena_cap || g->var_len + f->need_v > g->var_cap || (size_t)g->njobs + (size_t)f->nreq > (size_t)g->jobs_cap)) {
      f->submitted = false;
      ++g->deferred; ++deferred_now;
      continue;
    }
    f->submitted = true;
    f->base = g->njobs;
    if (g->arena_len + f->need_a > g->arena_cap || g->var_len + f->need_v > g->var_cap || (size_t)g->njobs + (size_t)f->nreq > (size_t)g->jobs_cap)
      lane_grow(g, MAX2(g->arena_cap * 2, g->arena_len + f->need_a + 1024), MAX2(g->jobs_cap * 2, g->njobs + f->nreq),
                MAX2(g->var_cap * 2, g->var_len + f->need_v + 1024));
    for (int i = 0; i < f->nreq; ++i) {
      const ef_req *r = &f->reqs[i];
      pc_job *j = &g->jobs[g->njobs++];
      memset(j, 0, sizeof *j);
      j->op = (uint32_t)r->op;
      j->a_off = (uint32_t)g->arena_len; j->a_len = (uint32_t)r->a.len;
      if (r->a.len) copy_small(g->arena + g->arena_len, r->a.p, (size_t)r->a.len);
      g->arena_len += (size_t)r->a.len;
      if (r->b.nul_after) j->flags |= PC_B_NUL_AFTER;
      if (r->op == PC_OP_KBAND) j->flags |= PC_KBAND_OK_ONLY;      /* clean_noisy_exons reads the boolean only, like the reference's call sites */
      if (r->b.in_genome) { j->flags |= PC_B_IN_GENOME; j->b_off = (uint32_t)r->b.gen_off; j->b_len = (uint32_t)r->b.len; }
      else {
        j->b_off = (uint32_t)g->arena_len; j->b_len = (uint32_t)r->b.len;
        /* BORDERS reads the byte that follows t (refine.c:362-374); callers keep it addressable unless nul_after */
        const size_t nb = (size_t)r->b.len + ((r->op == PC_OP_BORDERS && !r->b.nul_after) ? 1u : 0u);
        if (nb) copy_small(g->arena + g->arena_len, r->b.p, nb);
        g->arena_len += nb;
      }
      j->p0 = r->p0; j->p1 = r->p1; j->p2 = r->p2;
      if (r->op == PC_OP_ALIGN || r->op == PC_OP_GAP) {
        j->out_cap = (uint32_t)(r->a.len + r->b.len);
        j->out_off = (uint32_t)g->var_len; g->var_len += j->out_cap;
      } else if (r->op == PC_OP_SEED) {
        g->var_len = (g->var_len + 3u) & ~(size_t)3u;
        j->out_cap = (uint32_t)r->out_cap;
        j->out_off = (uint32_t)g->var_len; g->var_len += 12u * (size_t)r->out_cap;
      }
    }
  }
  if (deferred_now * 4 > waiting_now) g->want_grow = true;
  if (g->njobs == 0) return;
  if (g->arena_len >= 0xfff00000u || g->var_len >= 0xfff00000u) {
    fprintf(stderr, "* FATAL est-fact: one batch exceeds 4 GiB; lower --fibers\n");
    exit(1);
  }
}

static bool run_group(worker *w, group *g) {
  /* returns false when the group has nothing left to do and no new item could be started */
  if (g->pending) {
    const double t0 = ef_now();
    const int rc = pce_wait(efc_lane(g->conn, g->lane_k), efc_alive, g->conn);
    if (rc) {
      if (rc > 0) fprintf(stderr, "* FATAL est-fact: the engine (est-factd) went away while a batch was in flight\n");
      else fprintf(stderr, "* FATAL est-fact: the engine failed a device batch (status %d; see the engine's log / stderr)\n", rc);
      exit(1);
    }
    w->gpu_wait += ef_now() - t0;
    g->pending = false;
    for (int k = 0; k < g->nslots; ++k)
      if (g->fibers[k].state == F_WAITING && g->fibers[k].submitted) { g->fibers[k].state = F_RUNNABLE; g->fibers[k].has_results = true; }
  }
  bool any = false;
  int started_now = 0;
  const double tf0 = ef_now();
again:
  g->rerun = false;
  any = false;
  for (int k = 0; k < g->nslots; ++k) {
    fiber *f = &g->fibers[k];
    /* A thousand suspended ESTs per group do not fit the caches: by the time a fiber runs again its stack top, its task
     * and its request list have been evicted.  Ask for them a few fibers ahead (the slots are visited in order). */
    if (k + EF_PF_FAR < g->nslots) __builtin_prefetch(&g->fibers[k + EF_PF_FAR].sp);
    if (k + EF_PF_NEAR < g->nslots) fiber_prefetch(g, &g->fibers[k + EF_PF_NEAR]);
    if (k >= g->nfibers) {                      /* child slots: run when runnable, never take a new EST */
      if (f->state == F_RUNNABLE) {
        tl_fiber = f;
        tl_phase = f->phase;
#if EF_FAST_SWITCH
        ctx_switch(&w->main_sp, f->sp);
#else
        swapcontext(&w->main_ctx, &f->ctx);
#endif
        tl_fiber = NULL;
      }
      if (f->state == F_DONE) { fiber_check_stack(f); f->state = F_FREE; }
      if (f->state == F_WAITING) any = true;
      continue;
    }
    for (;;) {
      if (f->state == F_FREE || f->state == F_DONE) {
        if (f->state == F_DONE) { --w->inflight; f->state = F_FREE; }
        /* Ramp: a group takes at most 64 new ESTs per round.  ESTs are dealt longest first; a thread that filled all
         * its fibers in one go would own the thousand longest ones (on mixed EST / mRNA inputs that is most of the
         * work of the run, fixed at t = 0).  Taking them in small bites while the other threads do the same spreads
         * the heavy ESTs evenly. */
        if (started_now >= 64 || atomic_load_explicit(&g_no_more, memory_order_relaxed)) break;
        size_t handle;
        const int got = g_next(g_next_user, &handle);
        if (got <= 0) { if (got < 0) atomic_store(&g_no_more, 1); break; }
        fiber_start(w, g, f, handle);
        ++w->inflight; ++started_now;
      }
      if (f->state != F_RUNNABLE) break;
      tl_fiber = f;
#if EF_FAST_SWITCH
      ctx_switch(&w->main_sp, f->sp);
#else
      swapcontext(&w->main_ctx, &f->ctx);
#endif
      tl_fiber = NULL;
      if (f->state == F_DONE) fiber_check_stack(f);       /* once per EST: the canary's line is in no cache */
      if (f->state == F_WAITING) break;       /* F_DONE: loop to pick the next item */
    }
    if (f->state == F_WAITING) any = true;
  }
  if (g->rerun) goto again;                     /* a parent whose last child just finished sits behind the cursor */
  const double tf1 = ef_now();
  w->t_fibers += tf1 - tf0;
  if (!any) return false;
  gather(g);
  const double tf2 = ef_now();
  w->t_gather += tf2 - tf1;
  if (g->njobs) {
    pce_lane *l = efc_lane(g->conn, g->lane_k);
    l->njobs = (uint32_t)g->njobs; l->arena_len = g->arena_len; l->var_len = g->var_len;
    pce_post(g->conn->hdr, l);
    g->pending = true;
    w->batches++; w->jobs += (uint64_t)g->njobs;
    w->h2d += g->arena_len + sizeof(pc_job) * (uint64_t)g->njobs;
    w->d2h += g->var_len + sizeof(int32_t) * PC_RES_INTS * (uint64_t)g->njobs;
  }
  w->t_submit += ef_now() - tf2;
  return true;
}

static void *worker_main(void *arg) {
  worker *w = arg;
  tl_worker = w;
  const double tw0 = ef_now();
  const int ng = g_ngroups;
  for (int i = 0; i < ng; ++i) { w->g[i].w = w; lane_refresh(&w->g[i]); }
  const double tw1 = ef_now();
  bool alive[EF_MAX_GROUPS];
  for (int i = 0; i < ng; ++i) alive[i] = true;
  for (;;) {
    bool any_alive = false;
    for (int i = 0; i < ng; ++i) {
      if (alive[i] || w->g[i].pending) alive[i] = run_group(w, &w->g[i]);
      any_alive |= alive[i];
    }
    if (any_alive) continue;
    if (atomic_load(&g_no_more)) break;
    struct timespec ts = {0, 100 * 1000};                          /* nothing in flight, but the reader may still deliver */
    nanosleep(&ts, NULL);
    alive[0] = true;
  }
  const double tw2 = ef_now();
  /* No tear-down: est-fact exits right after the last EST, and freeing pinned / device memory (two dozen streams,
   * each call synchronising the device) only delays the threads that are still working. */
  pthread_mutex_lock(&g_stat_mu);
  g_batches += w->batches; g_jobs += w->jobs; g_gpu_wait += w->gpu_wait; g_h2d += w->h2d; g_d2h += w->d2h;
  for (int i = 0; i < ng; ++i) { g_deferred += w->g[i].deferred; g_grows += w->g[i].grows; }
  g_t_fibers += w->t_fibers; g_t_gather += w->t_gather; g_t_submit += w->t_submit;
  g_t_init += tw1 - tw0; g_t_fini += ef_now() - tw2;
  g_t_end_sum += tw2; if (g_t_end_min == 0 || tw2 < g_t_end_min) g_t_end_min = tw2;
  for (int i = 0; i < EF_PH_COUNT; ++i) { g_phase_s[i] += ef_ticks_to_s(tl_phase_s[i]); g_phase_yields[i] += tl_phase_yields[i]; }
  if (tl_max_yields > g_max_yields) g_max_yields = tl_max_yields;
  pthread_mutex_unlock(&g_stat_mu);
  return NULL;
}

void sched_bytes(uint64_t *h2d, uint64_t *d2h) { *h2d = g_h2d; *d2h = g_d2h; }

void sched_breakdown(double *fibers_s, double *gather_s, double *submit_s) {
  *fibers_s = g_t_fibers; *gather_s = g_t_gather; *submit_s = g_t_submit;
}

void sched_stats(double *gpu_wait_s, uint64_t *batches, uint64_t *jobs) {
  if (gpu_wait_s) *gpu_wait_s = g_gpu_wait;
  if (batches) *batches = g_batches;
  if (jobs) *jobs = g_jobs;
}

/* Opening the engine sessions (in-process: CUDA context + pinned lanes + genome index, about a second; server: a
 * handshake and the index build, milliseconds) runs on its own thread as soon as the genome is in memory, so it overlaps
 * reading and preparing the ESTs. */
static struct {
  pthread_t th; bool started; int rc, nuse, use[16], nthreads, per_group; ef_conn *conns[16];
  const ef_config *cfg; const ef_seq *gen; double secs;
} g_prep;
static pc_session_stats g_engine_stats;
static const char *g_engine_mode = "";

/* VA budget: pintron.py runs est-fact under `ulimit -v` (3000 MiB by default, dist-scripts/pintron.py:207-213).  Fiber
 * stacks are address space, not memory, but they count: keep them (and malloc's per-thread arenas) inside a third of it. */
static size_t va_limit(void) {
  struct rlimit rl;
  if (getrlimit(RLIMIT_AS, &rl) != 0 || rl.rlim_cur == RLIM_INFINITY) return 0;
  return (size_t)rl.rlim_cur;
}

static void *prepare_main(void *arg) {
  (void)arg;
  const double t0 = ef_now();
  const ef_config *cfg = g_prep.cfg;
  const ef_seq *gen = g_prep.gen;
  if (cfg->n_devices > 0) for (int i = 0; i < cfg->n_devices; ++i) g_prep.use[g_prep.nuse++] = cfg->devices[i];
  else g_prep.use[g_prep.nuse++] = -1;                 /* none named: the server picks its least loaded GPU, an in-process engine takes GPU 0 */
  const int nuse = g_prep.nuse, nthreads = g_prep.nthreads;
  const size_t per_fiber = 4096;
  ef_conn_req req;
  memset(&req, 0, sizeof req);
  req.genome = gen->seq; req.genome_len = (size_t)gen->len; req.word_len = (int)cfg->min_factor_len; req.depth_rate = cfg->min_string_depth_rate;
  req.arena_cap = MAX2((size_t)1 << 20, (size_t)g_prep.per_group * per_fiber);
  req.jobs_cap = (uint32_t)MAX2(4096, g_prep.per_group * 32);
  {
    const char *kb = getenv("EF_STAGING_KB");        /* tests: tiny lanes, so that back-pressure and growth both happen */
    if (kb && atol(kb) > 0) { req.arena_cap = (size_t)atol(kb) << 10; req.jobs_cap = (uint32_t)MAX2(64, atol(kb)); }
  }
  req.var_cap = req.arena_cap;
  req.timers = getenv("PC_PROFILE") != NULL;
  for (int d = 0; d < nuse; ++d) {
    int nth = 0;
    for (int i = 0; i < nthreads; ++i) if (i % nuse == d) ++nth;
    if (nth == 0) continue;
    req.nlanes = g_ngroups * nth;
    req.device = d;
    char err[768];
    ef_conn *c = efc_open(cfg->engine, g_prep.use, g_prep.use[0] >= 0 ? nuse : 0, &req, err, sizeof err);
    if (!c) {
      fprintf(stderr, "* FATAL est-fact: no CUDA device available (no GPU engine: %s).  This build has no CPU path", err);
      if (va_limit()) fprintf(stderr, " (note: this process runs under a %zu MiB address-space limit, which a CUDA context does not fit: "
                                      "start est-factd outside the limit and est-fact will use it)", va_limit() >> 20);
      fprintf(stderr, "\n");
      g_prep.rc = 1;
      return NULL;
    }
    if (d > 0 && c->daemon != g_prep.conns[0]->daemon) { fprintf(stderr, "* FATAL est-fact: engine sessions ended up in different places\n"); g_prep.rc = 1; return NULL; }
    g_prep.conns[d] = c;
  }
  g_engine_mode = g_prep.conns[0]->daemon ? "est-factd" : "in-process";
  g_prep.secs = ef_now() - t0;
  return NULL;
}

void sched_prepare(const ef_config *cfg, const ef_seq *gen) {
  /* no mmap / munmap / trim behind malloc: see ar_alloc */
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TRIM_THRESHOLD, 1 << 30);
  memset(&g_prep, 0, sizeof g_prep);
  g_prep.cfg = cfg; g_prep.gen = gen;
  /* defaults from measurements on a 16-core B200 host (tools/e2e_probe.py, profiles/r2_summary.md): one worker per core (the
   * engine's submission loop and the writers float over the same cores: 16 workers beat 14), 768 ESTs in flight per group */
  const int ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN);
  int nthreads = cfg->threads > 0 ? cfg->threads : ncpu;
  if (nthreads < 1) nthreads = 1;
  { const char *e = getenv("EF_GROUPS"); g_ngroups = e && atoi(e) >= 2 && atoi(e) <= EF_MAX_GROUPS ? atoi(e) : EF_GROUPS_DEFAULT; }
  if (nthreads > PCE_MAX_SESSION_LANES / g_ngroups) nthreads = PCE_MAX_SESSION_LANES / g_ngroups;
  int per_group = cfg->fibers > 0 ? cfg->fibers : 768;       /* measured (C3, 16 and 4 threads): fewer leave the threads waiting for the engine, more outgrow the caches */
  const size_t lim = va_limit();
  if (lim) {
    mallopt(M_ARENA_MAX, 2);
    const size_t budget = lim / 3 / FIBER_STACK * 2 / 3;       /* EST fibers in total (child fibers add half as many again) */
    if ((size_t)per_group * (size_t)g_ngroups * (size_t)nthreads > budget) per_group = (int)MAX2((size_t)8, budget / ((size_t)g_ngroups * (size_t)nthreads));
  }
  g_prep.nthreads = nthreads; g_prep.per_group = per_group;
  if (pthread_create(&g_prep.th, NULL, prepare_main, NULL)) { perror("pthread_create"); exit(1); }
  g_prep.started = true;
}

void sched_engine_stats(pc_session_stats *sum, const char **mode) { if (sum) *sum = g_engine_stats; if (mode) *mode = g_engine_mode; }

int sched_run(const ef_config *cfg, const ef_seq *gen, size_t n_items, ef_next_fn next, ef_task_fn fn, void *user) {
  const double ts0 = ef_now();
  if (!g_prep.started) sched_prepare(cfg, gen);
  pthread_join(g_prep.th, NULL);
  g_prep.started = false;
  if (g_prep.rc) return 1;
  const int nuse = g_prep.nuse;
  int *use = g_prep.use;
  int nthreads = g_prep.nthreads;            /* lanes exist for this many; a short input uses fewer */
  if ((size_t)nthreads > n_items) {
    /* keep the thread -> lane mapping: only the first n threads run */
    nthreads = n_items ? (int)n_items : 1;
  }
  int per_group = g_prep.per_group;
  if ((size_t)per_group * (size_t)g_ngroups * (size_t)nthreads > n_items) per_group = (int)(n_items / ((size_t)g_ngroups * (size_t)nthreads)) + 1;
  g_next = next; g_next_user = user;
  atomic_store(&g_no_more, 0);
  g_nthreads = nthreads;
  g_batches = g_jobs = g_h2d = g_d2h = g_deferred = g_grows = 0; g_gpu_wait = 0; g_t_fibers = g_t_gather = g_t_submit = g_t_init = g_t_fini = g_t_end_sum = g_t_end_min = 0;
  const double ts1 = ef_now();
  worker *ws = calloc((size_t)nthreads, sizeof(worker));
  for (int i = 0; i < nthreads; ++i) {
    worker *w = &ws[i];
    w->id = i; w->device = use[i % nuse];
    w->cfg = cfg; w->gen = gen; w->fn = fn; w->user = user;
    for (int k = 0; k < g_ngroups; ++k) {
      w->g[k].conn = g_prep.conns[i % nuse];
      w->g[k].lane_k = g_ngroups * (i / nuse) + k;
      w->g[k].nfibers = per_group;
      w->g[k].nslots = per_group + MAX2(32, per_group / 2);       /* child fibers of dp_parallel_for */
      w->g[k].fibers = calloc((size_t)w->g[k].nslots, sizeof(fiber));
    }
    if (pthread_create(&w->th, NULL, worker_main, w)) { perror("pthread_create"); return 1; }
  }
  for (int i = 0; i < nthreads; ++i) pthread_join(ws[i].th, NULL);
  const double ts2 = ef_now();
  free(ws);
  memset(&g_engine_stats, 0, sizeof g_engine_stats);
  for (int d = 0; d < nuse; ++d) {
    if (!g_prep.conns[d]) continue;
    pc_session_stats st;
    efc_close(g_prep.conns[d], &st);
    g_engine_stats.batches += st.batches; g_engine_stats.lanes_merged += st.lanes_merged; g_engine_stats.jobs += st.jobs;
    g_engine_stats.launches += st.launches; g_engine_stats.h2d_bytes += st.h2d_bytes; g_engine_stats.d2h_bytes += st.d2h_bytes;
    g_engine_stats.retries += st.retries; g_engine_stats.busy_s += st.busy_s;
    for (int o = 0; o < PC_OP_COUNT; ++o) g_engine_stats.op_ms[o] += st.op_ms[o];
    g_prep.conns[d] = NULL;
  }
  if (!cfg->quiet)
    fprintf(stderr, "* INFO  scheduler: %d thread(s) x %d x %d fibers on %d GPU(s), engine: %s; engine session(s) open after %.3f s (%.3f s of it still to wait for), workers %.3f s "
            "(set-up %.3f s per thread on average), session close %.3f s; "
            "%llu fiber deferrals, %llu lane re-allocations; threads idle at the end %.3f s on average (first done %.3f s before the last)\n",
            nthreads, g_ngroups, per_group, nuse, g_engine_mode, g_prep.secs, ts1 - ts0, ts2 - ts1, g_t_init / nthreads, ef_now() - ts2,
            (unsigned long long)g_deferred, (unsigned long long)g_grows, ts2 - g_t_end_sum / nthreads, ts2 - g_t_end_min);
  return 0;
}
