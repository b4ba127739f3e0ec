/*
 * pintron_cuda.h — the C ABI of libpintron_cuda.so: the B200 (sm_100a) implementation of the integer DP
 * and seeding routines of PIntron's est-fact stage.
 *
 * The reference has no FFI: est-fact is one single-threaded C process whose per-EST code calls the
 * routines below directly.  This header is the seam our C host (pintron_b200/host/)
 * uses instead, and what a maintainer of the reference would bind to replace those calls (see
 * INTEGRATION.md).  Plain C: pointers, sizes, ints; no C++/torch types; errors are negative return values
 * plus pc_last_error().  There is NO CPU fallback: without a CUDA device pc_ctx_create fails.
 *
 * Work is submitted in BATCHES of jobs.  A job names two byte strings inside one caller-provided arena
 * (or, for the genome side, inside the genome uploaded with pc_genome_upload) and one operation:
 *
 *   op              replaces (reference file:line)                          res[1..] (res[0] = status)
 *   PC_OP_ALIGN     compute_alignment           src/compute-alignments.c:39   score, ops_len      (+ops bytes)
 *   PC_OP_KBAND     K_band_edit_distance        src/compute-alignments.c:319  ok, edit
 *   PC_OP_EDIT      edit_distance (last cell)   src/refine.c:51
 *                   compute_edit_distance       src/compute-alignments.c:235  distance
 *   PC_OP_BORDERS   general_refine_borders      src/refine.c:106              ok, off_p, off_t1, off_t2, ed
 *   PC_OP_GAP       compute_gap_alignment       src/refine-intron.c:560       dim, factor_cut, intron_start,
 *                                                                             intron_end, intron_start_on_align,
 *                                                                             intron_end_on_align (+ops bytes)
 *   PC_OP_AFFIX     find_longest_affix          src/factorization-refinement.c:1134  valid, est_cut, gen_cut
 *   PC_OP_SUFCUT    compute_best_suffix_cut     src/compute-alignments.c:246  ed, cut1, cut2
 *   PC_OP_PRECUT    compute_best_prefix_cut     src/compute-alignments.c:290  ed, cut1, cut2
 *   PC_OP_LCS       find_longest_common_factor_dp  src/factorization-refinement.c:255  len, occ1, occ2
 *   PC_OP_SEED      build_vertex_set            src/max-emb-graph.c:217       count        (+(p,t,l) int32 triples)
 *     with p1 = PC_SEED_BUILD_MEG the job goes on, on the device, through the rest of build_meg
 *     (src/compute-est-fact.c:90-152): build_edge_set (src/max-emb-graph.c:649), simplify_meg,
 *     transitive_reduction, compact_short_edges and is_too_complex (src/meg-simplification.c:314,518,258,89);
 *     the result is the finished graph (struct pc_meg_cfg / "MEG record" below), count = its size in 12-byte units
 *
 * Alignment ops: one byte per alignment column, left to right: 0 = EST char over genome char,
 * 1 = EST char over '-', 2 = '-' over genome char (the reference's EST_alignment / GEN_alignment rows,
 * include/types.h:221-253, are rebuilt from them by the host).
 */
#ifndef PINTRON_CUDA_H
#define PINTRON_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pc_ctx pc_ctx;        /* one per GPU: device-resident genome + k-mer index */
typedef struct pc_stream pc_stream;  /* one per host worker thread: CUDA stream, staging, scratch pool */

enum pc_op {
  PC_OP_ALIGN = 0, PC_OP_KBAND = 1, PC_OP_EDIT = 2, PC_OP_BORDERS = 3, PC_OP_GAP = 4, PC_OP_AFFIX = 5,
  PC_OP_SUFCUT = 6, PC_OP_PRECUT = 7, PC_OP_LCS = 8, PC_OP_SEED = 9, PC_OP_COUNT = 10
};

enum pc_flags {
  PC_B_IN_GENOME = 1u,  /* b_off/b_len index the genome uploaded with pc_genome_upload, not the arena */
  PC_B_NUL_AFTER = 2u,  /* BORDERS: the byte after t reads as NUL (the reference passes a NUL-terminated copy of t,
                           src/est-factorizations.c:1490), whatever follows it in the arena or the genome */
  PC_KBAND_OK_ONLY = 4u /* KBAND: the caller reads only `ok`, as every call site of the reference does
                           (src/est-factorizations.c:1876-1888, src/factorization-refinement.c:931).  When the distance is
                           above k the reference's `edit` is the value of its band-restricted matrix, which only a banded
                           sweep reproduces; with this flag res[2] then holds the true edit distance instead and the job
                           never leaves the bit-parallel kernel */
};

enum pc_status {
  PC_OK = 0,
  PC_E_POOL = -1,       /* internal: device scratch pool exhausted (pc_stream_sync grows it and retries) */
  PC_E_OUTCAP = -2,     /* out_cap too small; res[1] = needed count */
  PC_E_RANGE = -3,      /* a length outside what the kernel supports */
  PC_E_CUDA = -10, PC_E_ARG = -11, PC_E_NOMEM = -12
};

#define PC_RES_INTS 8

typedef struct pc_job {
  uint32_t op;            /* enum pc_op */
  uint32_t flags;         /* enum pc_flags */
  uint32_t a_off, a_len;  /* EST-side string (ALIGN/GAP/AFFIX/SEED: the EST; BORDERS: p; LCS: s2; else s1) */
  uint32_t b_off, b_len;  /* genome-side string (BORDERS: t; LCS: s1 = the long one; unused for SEED) */
  int32_t  p0, p1, p2;    /* KBAND: p0 = upper bound k.  BORDERS: p0 = max_errs, p1 = min_p_cut, p2 = max_p_cut.
                             SEED: p0 = min factor length in force (config + inc_pairing_len); p1 = 0 (vertex set only) or
                             PC_SEED_BUILD_MEG (b = one struct pc_meg_cfg in the arena; p2 = largest vertex set built on the
                             device, 0 = any). */
  uint32_t out_off;       /* byte offset of this job's variable output inside var_out (4-aligned for SEED) */
  uint32_t out_cap;       /* ALIGN/GAP: bytes (>= a_len + b_len); SEED: capacity in (p,t,l) triples (= 12-byte units) */
} pc_job;

/* ---- PC_OP_SEED with p1 = PC_SEED_BUILD_MEG: the whole Maximal Embedding Graph on the device -------------------
 * b (in the arena, any alignment) = the options build_meg reads (src/options.ggo; include/configuration.h).
 * MEG record written at out_off (int32 words; PC_E_OUTCAP + needed 12-byte units in res[1] when it does not fit):
 *   [0] nv   vertices, in the order the reference's lists hold them (V[0] = source, V[i+1] = pairings at p = i in list
 *            order, V[|P|+1] = sink) = the numbering of megs.txt (src/io-meg.c:147-190)
 *   [1] ne   edges            [2] 1 = "too complex": build again with p0 + 1 (src/compute-est-fact.c:131-146)      [3] 0
 *   then nv x (p, t, l)  (source: INT32_MIN, INT32_MIN, 200; sink: INT32_MAX - 200, INT32_MAX - 200, 200),
 *   then nv adjacency counts, then the ne adjacency targets (vertex numbers), list by list in list order. */
#define PC_SEED_BUILD_MEG 1
/* p2 > 0 on such a job: build the graph on the device only when the vertex set has at most p2 pairings; for a larger one
 * the job answers res[3] = PC_SEED_VERTEX_SET_ONLY, res[1] = the number of (p, t, l) triples, written at out_off
 * (PC_E_OUTCAP + that number when they do not fit), and the caller builds the graph from them (one sequential walk per
 * graph: quadratic in the vertices, milliseconds for one GPU lane on an mRNA, microseconds on a host core). */
#define PC_SEED_VERTEX_SET_ONLY 1
typedef struct pc_meg_cfg {
  int32_t min_intron_length, max_intron_length;          /* --min-intron-length, --max-intron-length (0 = unlimited) */
  uint32_t max_pairings_in_MEG;                          /* --max-pairings-in-CMEG */
  uint32_t flags;                                        /* 1 = transitive reduction, 2 = short-edge compaction */
  double max_prefix_discarded_rate, max_suffix_discarded_rate, max_freq_shortest_pairing;
} pc_meg_cfg;
#define PC_MEG_TRANS_RED 1u
#define PC_MEG_SHORT_EDGE_COMP 2u

/* ---- context / errors ---------------------------------------------------------------------------- */
const char *pc_last_error(void);                 /* thread-local message of the last failing call */
int pc_device_count(void);                       /* < 0 on error (no driver / no device) */
pc_ctx *pc_ctx_create(int device);               /* NULL on failure */
void pc_ctx_destroy(pc_ctx *ctx);

/* Replaces lst_stree_new + preprocess_text + stree_preprocess (src/main-est-fact.c:223-240,
 * stree_src/lst_stree.c:816, src/aug_suffix_tree.c:69,248): copies the (N-tail-stripped) genome to HBM and
 * builds the device k-mer index over it.  word_len = configured min-factor-length (options.ggo -l, default 15);
 * depth_rate = min-string-depth-rate (-d, default 0.2). */
int pc_genome_upload(pc_ctx *ctx, const char *genome, size_t len, int word_len, double depth_rate);

/* ---- streams ------------------------------------------------------------------------------------- */
pc_stream *pc_stream_create(pc_ctx *ctx);
void pc_stream_destroy(pc_stream *st);
void *pc_host_alloc(size_t bytes);               /* pinned host memory for arenas / results (optional) */
void pc_host_free(void *p);

/* Enqueue one mixed batch: H2D of arena + jobs, one kernel per op present, D2H of res and var_out.
 * Asynchronous; res / var_out / jobs / arena must stay valid until pc_stream_sync returns.
 * res: njobs * PC_RES_INTS int32.  Returns 0 or a negative pc_status. */
int pc_submit(pc_stream *st, const uint8_t *arena, size_t arena_bytes, const pc_job *jobs, int njobs,
              int32_t *res, uint8_t *var_out, size_t var_out_bytes);

/* Same, for inputs that are ALREADY in device memory (bench "value" leg, chained device pipelines):
 * d_arena / d_jobs / d_res / d_var_out are device pointers; nothing is copied.  d_arena must be readable (and zero)
 * for 16 bytes past arena_bytes, d_var_out writable for 16 bytes past var_out_bytes (pc_submit pads its own copies).
 * h_jobs (a host copy of the jobs) may be NULL: the batch is then validated, keyed and ordered on the device whatever
 * its size (what the engine does with merged lanes, include/pintron_engine.h). */
int pc_submit_device(pc_stream *st, const uint8_t *d_arena, size_t arena_bytes, const pc_job *d_jobs,
                     const pc_job *h_jobs, int njobs, int32_t *d_res, uint8_t *d_var_out, size_t var_out_bytes);

int pc_stream_sync(pc_stream *st);               /* waits; transparently retries jobs that hit PC_E_POOL */

/* Per-routine entry points (thin wrappers over pc_submit + pc_stream_sync that insist every job has the
 * matching op) — the names a binding for the reference call sites would use. */
int pc_compute_alignment_batch(pc_stream *, const uint8_t *arena, size_t, const pc_job *, int, int32_t *res, uint8_t *ops, size_t);
int pc_kband_edit_distance_batch(pc_stream *, const uint8_t *arena, size_t, const pc_job *, int, int32_t *res);
int pc_edit_distance_batch(pc_stream *, const uint8_t *arena, size_t, const pc_job *, int, int32_t *res);
int pc_refine_borders_batch(pc_stream *, const uint8_t *arena, size_t, const pc_job *, int, int32_t *res);
int pc_gap_alignment_batch(pc_stream *, const uint8_t *arena, size_t, const pc_job *, int, int32_t *res, uint8_t *ops, size_t);
int pc_longest_affix_batch(pc_stream *, const uint8_t *arena, size_t, const pc_job *, int, int32_t *res);
int pc_best_cut_batch(pc_stream *, const uint8_t *arena, size_t, const pc_job *, int, int32_t *res);
int pc_longest_common_factor_batch(pc_stream *, const uint8_t *arena, size_t, const pc_job *, int, int32_t *res);
int pc_build_vertex_set_batch(pc_stream *, const uint8_t *arena, size_t, const pc_job *, int, int32_t *res, uint8_t *triples, size_t);

/* ---- instrumentation ----------------------------------------------------------------------------- */
/* Environment (read once when the library is loaded): PC_PROFILE=1 makes pc_debug_dump() print host-side phase times,
 * device time / jobs per op, hand-over and re-run counts; PC_CAPTURE=<file> appends every batch given to pc_submit to
 * <file> (u32 njobs, u64 arena_bytes, jobs, arena; pc_submit_parts writes u32 0xffffffff, u64 nparts before the parts it ran
 * as one device batch) — bench.py replays such a recording as its device workload and tools/check_capture.py checks every
 * recorded job against the oracle. */
uint64_t pc_launch_count(void);                  /* kernels launched by this library in this process */
/* Device time (ms, CUDA events on the stream) and launches of the kernels of `op` since the last reset. */
int pc_stream_op_time(pc_stream *st, int op, double *ms, uint64_t *launches);
void pc_stream_reset_timers(pc_stream *st);
void pc_stream_enable_timers(pc_stream *st, int on);
void *pc_stream_cuda_stream(pc_stream *st);    /* the cudaStream_t, for callers that interoperate */
/* INT32 ALU micro-benchmark: lane-operations per second of independent add/min chains (the roofline
 * denominator for the DP kernels; SURVEY.md §8(d) asks for a measured figure). */
double pc_measure_int_peak(pc_ctx *ctx);
void pc_debug_dump(void);                     /* prints library-side timings to stderr when PC_PROFILE (or PC_PROFILE_HOST: phase clock only) is set */
/* How a submitting thread waits for its stream: 0 = the driver's spinning cudaStreamSynchronize (lowest latency, one core per
 * submission loop), 1 = sleep on a cudaEventBlockingSync event (measured: 0.3-0.5 ms later per wake-up; nothing turns it on by default).
 * PC_SYNC=spin|block in the environment wins over this call.  PC_SIDE_STREAMS=1..8: side streams a small batch forks over. */
void pc_set_blocking_sync(int on);

#ifdef __cplusplus
}
#endif
#endif
