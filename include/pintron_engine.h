/*
 * pintron_engine.h — the batch ENGINE of libpintron_cuda.so and the shared-memory LANE protocol its clients speak.
 *
 * Why it exists.  The reference runs est-fact as one process per gene (dist-scripts/pintron.py:878-884) whose hot loop
 * (src/main-est-fact.c:249-291) calls the DP routines one by one.  Our host runs thousands of ESTs as fibers on several
 * worker threads; each thread-group gathers the DP requests of its fibers into a LANE (arena + jobs + results).  The
 * engine owns the GPU: one submission loop per GPU claims every lane that is posted at that moment and runs them as ONE
 * device batch (one ordering pass, one launch per job class, copies per lane) — instead of every worker thread driving
 * its own CUDA stream with its own forty launches.  The engine runs either
 *   - inside the est-fact process (`--engine inproc`), or
 *   - inside the resident server `est-factd`, with est-fact as a CUDA-free client: lanes live in memfd segments the
 *     server has pinned once; the client maps them (fd passed over a UNIX socket) and rings a futex doorbell.  No CUDA
 *     context is created per job, several est-fact processes share the GPUs of a box, and the client stays inside
 *     pintron.py's default `ulimit -v` (dist-scripts/pintron.py:207-213), which a CUDA context does not.
 * Both forms use the same lane protocol below; results never depend on which one runs, nor on how lanes get merged.
 *
 * Plain C (C11 / C++17), no CUDA types.  The lane protocol is header-only; the pc_engine_* functions are exported by
 * libpintron_cuda.so.
 */
#ifndef PINTRON_ENGINE_H
#define PINTRON_ENGINE_H

#include <stddef.h>
#include <stdint.h>
#include "pintron_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- lane protocol (shared memory) ------------------------------------------------------------------------------ */
#define PCE_MAGIC 0x50434532u        /* "PCE2" */
#define PCE_VERSION 2u
#define PCE_MAX_LANES 512            /* per GPU */
#define PCE_MAX_SEGMENTS 64          /* per GPU */
#define PCE_MAX_SESSION_LANES 256

enum pce_lane_state { PCE_FREE = 0, PCE_IDLE = 1, PCE_POSTED = 2, PCE_RUNNING = 3, PCE_DONE = 4 };

/* One lane = the staging of one thread-group of one client.  The slab [arena | jobs | res | var] lies in segment `seg`
 * at the byte offsets below.  The client fills arena / jobs, sets njobs / arena_len / var_len and posts; the engine
 * writes res / var, sets rc and marks the lane DONE.  Job offsets are relative to the lane's own arena / var. */
typedef struct pce_lane {
  uint32_t state;                    /* enum pce_lane_state; atomic; futex word the client sleeps on */
  int32_t rc;                        /* 0 or a negative pc_status for the whole batch */
  uint32_t session;                  /* owner (0 = none) */
  uint32_t njobs;
  uint64_t arena_len, var_len;
  uint32_t seg, jobs_cap;
  uint64_t arena_off, arena_cap, jobs_off, res_off, var_off, var_cap;
  uint64_t batches, jobs_total;      /* statistics (engine side) */
  uint8_t pad[128 - 104];
} pce_lane;

typedef struct pce_hdr {             /* at offset 0 of segment 0 of a GPU */
  uint32_t magic, version;
  uint32_t doorbell;                 /* atomic; bumped by a client after posting; the engine sleeps on it */
  uint32_t sleepers;                 /* engine threads asleep on the doorbell (clients skip the wake call when 0) */
  uint8_t pad[128 - 16];
  pce_lane lanes[PCE_MAX_LANES];
} pce_hdr;

#if defined(__cplusplus)
static_assert(sizeof(pce_lane) == 128, "pce_lane layout");
#else
_Static_assert(sizeof(pce_lane) == 128, "pce_lane layout");
#endif
#define PCE_HDR_BYTES ((sizeof(pce_hdr) + 4095u) & ~(size_t)4095u)

/* ---- engine (server side; exported by libpintron_cuda.so) ------------------------------------------------------- */
typedef struct pc_engine pc_engine;

/* devices: CUDA device ordinals to serve (ndev >= 1).  segment_bytes: size of the first pinned shared-memory segment
 * per GPU (0 = default); more segments are added when lanes do not fit.  NULL on failure (pc_last_error()). */
pc_engine *pc_engine_create(const int *devices, int ndev, size_t segment_bytes);
void pc_engine_destroy(pc_engine *e);
int pc_engine_gpu_count(const pc_engine *e);
/* which implementation serves: "cuda-sm100a" for libpintron_cuda.so.  A client only accepts a server of its own kind (the
 * test suite has a CPU stand-in for GPU-less containers; the shipped est-fact must never end up talking to it). */
const char *pc_engine_backend(void);

typedef struct pc_session_req {
  int gpu;                           /* index into the engine's device list, or -1: least loaded */
  const char *genome;                /* N-tail-stripped genome (replaces the suffix tree: pc_genome_upload) */
  size_t genome_len;
  int word_len;                      /* min-factor-length */
  double depth_rate;                 /* min-string-depth-rate */
  int nlanes;
  uint64_t arena_cap, var_cap;       /* per lane */
  uint32_t jobs_cap;                 /* per lane */
} pc_session_req;

typedef struct pc_session_info {
  uint32_t session;                  /* > 0 */
  int gpu, nlanes;
  uint32_t lane[PCE_MAX_SESSION_LANES];   /* indices into pce_hdr.lanes */
} pc_session_info;

typedef struct pc_session_stats {
  uint64_t batches, lanes_merged, jobs, launches, h2d_bytes, d2h_bytes, retries;
  double busy_s;                     /* wall time the engine spent on this session's batches */
  double op_ms[PC_OP_COUNT];         /* device time per op (CUDA events; 0 unless timers were on) */
} pc_session_stats;

int pc_engine_open(pc_engine *e, const pc_session_req *req, pc_session_info *out);
/* Move a lane to a larger slab (a fiber's requests outgrew it).  keep_arena bytes of the arena and keep_jobs jobs are
 * carried over.  The lane must be IDLE or DONE. */
int pc_engine_resize_lane(pc_engine *e, uint32_t session, uint32_t lane, uint64_t arena_cap, uint32_t jobs_cap, uint64_t var_cap,
                          uint64_t keep_arena, uint32_t keep_jobs);
int pc_engine_close(pc_engine *e, uint32_t session, pc_session_stats *stats);    /* waits for the session's lanes in flight */
void pc_engine_enable_timers(pc_engine *e, int on);

/* segments of one GPU: file descriptor (memfd, for SCM_RIGHTS), size, and the engine's own mapping */
int pc_engine_segment_count(pc_engine *e, int gpu);
int pc_engine_segment_fd(pc_engine *e, int gpu, int seg, size_t *bytes);
void *pc_engine_segment_base(pc_engine *e, int gpu, int seg);

/* ---- multi-part submission (what the engine runs; also usable directly) ------------------------------------------ */
typedef struct pc_part {
  const uint8_t *arena; size_t arena_bytes;
  const pc_job *jobs; int njobs;
  int32_t *res;                      /* njobs * PC_RES_INTS */
  uint8_t *var_out; size_t var_out_bytes;
} pc_part;
/* Runs the jobs of all parts as ONE device batch on `st` against the genome of `genome_ctx` (NULL: the stream's own
 * context).  Offsets inside a part's jobs are relative to that part's arena / var_out.  Asynchronous like pc_submit:
 * follow with pc_stream_sync(st); the parts array itself may be freed after the call returns. */
int pc_submit_parts(pc_stream *st, pc_ctx *genome_ctx, const pc_part *parts, int nparts);

/* ---- client-side helpers (header-only; used by the est-fact host and by tests) ---------------------------------- */
#if defined(__linux__)
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <time.h>
static inline long pce_futex(uint32_t *addr, int op, uint32_t val, const struct timespec *to) {
  return syscall(SYS_futex, addr, op, val, to, NULL, 0);          /* shared (non-private) futex: works across processes */
}
static inline void pce_post(pce_hdr *h, pce_lane *l) {
  __atomic_store_n(&l->state, (uint32_t)PCE_POSTED, __ATOMIC_RELEASE);
  __atomic_fetch_add(&h->doorbell, 1u, __ATOMIC_SEQ_CST);
  if (__atomic_load_n(&h->sleepers, __ATOMIC_SEQ_CST)) pce_futex(&h->doorbell, FUTEX_WAKE, 64, NULL);
}
/* Waits until the lane is DONE (returns its rc) — or until `alive` (optional) says the engine is gone (returns 1). */
static inline int pce_wait(pce_lane *l, int (*alive)(void *), void *alive_arg) {
  for (int spin = 0;; ++spin) {
    const uint32_t s = __atomic_load_n(&l->state, __ATOMIC_ACQUIRE);
    if (s == PCE_DONE) return l->rc;
    if (spin < 200) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      continue;
    }
    struct timespec to = {0, 100 * 1000 * 1000};
    pce_futex(&l->state, FUTEX_WAIT, s, &to);
    if (alive && (spin & 15) == 0 && !alive(alive_arg)) {
      if (__atomic_load_n(&l->state, __ATOMIC_ACQUIRE) == PCE_DONE) return l->rc;
      return 1;
    }
  }
}
#endif

#ifdef __cplusplus
}
#endif
#endif
