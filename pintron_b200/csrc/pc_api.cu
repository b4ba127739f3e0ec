// pc_api.cu — host side of the C ABI (include/pintron_cuda.h): contexts, streams, batch plumbing.
#include "pc_device.cuh"
#include "pintron_engine.h"
#include <algorithm>
#include <map>
#include <set>
#include <tuple>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

unsigned long long g_pc_launches = 0;
thread_local unsigned long long tl_pc_launches = 0;
#include <chrono>
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const bool g_prof = getenv("PC_PROFILE") != nullptr;
static const bool g_prof_host = g_prof || getenv("PC_PROFILE_HOST") != nullptr;      /* only the host-side phase clock: side streams stay on */
/* how many side streams a small batch forks over (launch_selected): every (op, class) kernel of such a batch is bound by the
 * latency of its longest job, so the batch takes the longest chain on one stream, not the sum.  PC_SIDE_STREAMS=1..8. */
/* compute_alignment through the bit-parallel kernel (k_myers.cu: k_align_bp); PC_ALIGN_BP=0 keeps every job on the wavefront kernel */
static const bool g_align_bp = [] { const char *v = getenv("PC_ALIGN_BP"); return !v || atoi(v) != 0; }();
static const int g_side_streams = [] { const char *v = getenv("PC_SIDE_STREAMS"); const int n = v ? atoi(v) : 0; return n >= 1 && n <= 8 ? n : 8; }();
#define PC_MULTI_STREAM_MAX ((size_t)1 << 18)        /* batches below this many jobs run their segments on side streams */
static const bool g_serial = getenv("PC_SERIAL_SEGMENTS") != nullptr;      /* experiments: keep every batch on one stream */
#define PC_POOL_MB_DEFAULT 320         /* scratch pool per stream (direction words, wavefront matrices); grows on demand */
#define PC_PARTS_HOST_ORDER_MAX ((size_t)1 << 14)   /* merged batches below this are keyed and ordered by the submitting thread: no mid-batch read-back */
#define PC_DEVICE_ORDER_MIN ((size_t)1 << 16)      /* batches from this size up are ordered on the device (k_order.cu) */
/* PC_CAPTURE=<file>: every batch handed to pc_submit is appended to <file> (bench.py replays the job stream of a real
 * est-fact run as its device-resident workload).  Record = u32 njobs, u64 arena_bytes, jobs, arena. */
#include <mutex>
static FILE *g_capture = getenv("PC_CAPTURE") ? fopen(getenv("PC_CAPTURE"), "wb") : nullptr;
static std::mutex g_capture_mu;
static double g_t[8];      /* check, reserve, h2d, sort, launch, d2h, sync (racy sums: diagnostics only) */
static unsigned long long g_dev_grows, g_host_allocs, g_host_alloc_bytes; static double g_host_alloc_s;
static unsigned long long g_op_jobs[PC_OP_COUNT], g_op_suma[PC_OP_COUNT], g_op_sumb[PC_OP_COUNT], g_op_maxa[PC_OP_COUNT], g_op_maxb[PC_OP_COUNT], g_op_cells[PC_OP_COUNT];
static unsigned long long g_op_slow[PC_OP_COUNT];
static double g_op_ms[PC_OP_COUNT]; static unsigned long long g_op_launches[PC_OP_COUNT], g_retry_rounds, g_retry_jobs, g_pool_grows;
extern "C" void pc_debug_dump(void) {
  if (g_capture) fflush(g_capture);
  if (g_prof_host && !g_prof) fprintf(stderr, "[pc profile] host: check %.3f reserve %.3f h2d %.3f sort %.3f launch %.3f d2h %.3f sync %.3f s\n", g_t[0], g_t[1], g_t[2], g_t[3], g_t[4], g_t[5], g_t[6]);
  if (!g_prof) return;
  static const char *nm[PC_OP_COUNT] = {"ALIGN", "KBAND", "EDIT", "BORDERS", "GAP", "AFFIX", "SUFCUT", "PRECUT", "LCS", "SEED"};
  fprintf(stderr, "[pc profile] host: check %.3f reserve %.3f h2d %.3f sort %.3f launch %.3f d2h %.3f sync %.3f s\n", g_t[0], g_t[1], g_t[2], g_t[3], g_t[4], g_t[5], g_t[6]);
  fprintf(stderr, "[pc profile] device ms per op:");
  for (int i = 0; i < PC_OP_COUNT; ++i) if (g_op_launches[i]) fprintf(stderr, " %s %.1f (%llu launches)", nm[i], g_op_ms[i], g_op_launches[i]);
  fprintf(stderr, "\n[pc profile] jobs per op (count, mean a_len, mean b_len, max a_len, max b_len, a*b cells):");
  for (int i = 0; i < PC_OP_COUNT; ++i) if (g_op_jobs[i]) fprintf(stderr, " %s %llu %.1f %.1f %llu %llu %.3g;", nm[i], g_op_jobs[i], (double)g_op_suma[i] / g_op_jobs[i], (double)g_op_sumb[i] / g_op_jobs[i], g_op_maxa[i], g_op_maxb[i], (double)g_op_cells[i]);
  fprintf(stderr, "\n[pc profile] jobs the bit-parallel kernel handed to the wavefront kernel: EDIT %llu, KBAND %llu", g_op_slow[PC_OP_EDIT], g_op_slow[PC_OP_KBAND]);
  fprintf(stderr, "\n[pc profile] pool retries: %llu rounds, %llu jobs, %llu pool growths\n", g_retry_rounds, g_retry_jobs, g_pool_grows);
  fprintf(stderr, "[pc profile] device buffer growths (cudaMalloc during the run): %llu; pinned host allocations: %llu, %.1f MB, %.3f s\n",
          g_dev_grows, g_host_allocs, g_host_alloc_bytes / 1048576.0, g_host_alloc_s);
}
struct ProfDump { ~ProfDump() { if (g_prof) fprintf(stderr, "[pc profile] check %.3f reserve %.3f h2d %.3f sort %.3f launch %.3f d2h %.3f sync %.3f s\n", g_t[0], g_t[1], g_t[2], g_t[3], g_t[4], g_t[5], g_t[6]); } } g_prof_dump;
#define PROF(slot, t0) do { if (g_prof_host) { double t1_ = now_s(); g_t[slot] += t1_ - (t0); (t0) = t1_; } } while (0)
static thread_local char g_err[512] = "";

/* How the submitting thread waits for its stream.  The driver's default spins: lowest latency, one core per submission loop.
 * PC_SYNC=block / pc_set_blocking_sync(1) make it sleep on a blocking event instead — measured slower everywhere we tried
 * (the wake-up costs 0.3-0.5 ms per batch), kept as a knob for boxes with fewer cores than GPUs. */
static const int g_sync_env = [] { const char *v = getenv("PC_SYNC"); return !v ? -1 : (strcmp(v, "block") == 0 ? 1 : (strcmp(v, "spin") == 0 ? 0 : -1)); }();
static int g_block_sync = g_sync_env > 0;
extern "C" void pc_set_blocking_sync(int on) { if (g_sync_env < 0) g_block_sync = on != 0; }

static int fail(int code, const char *fmt, const char *detail = "") {
  snprintf(g_err, sizeof g_err, fmt, detail);
  return code;
}
int pc_set_error(int code, const char *msg) { return fail(code, "%s", msg); }      /* for the other translation units */
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(PC_E_CUDA, #call ": %s", cudaGetErrorString(e_)); } while (0)

// ---- per-device launch facts, asked once ----------------------------------------------------------------------------
static std::mutex g_occ_mu;
static std::map<std::tuple<int, const void *, size_t>, int> g_occ;
int pc_cached_occupancy(const void *func, int tpb, size_t smem) {
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t kb = (smem + 1023) >> 10;                     /* asked for the size rounded up to 1 KB: never optimistic */
  const auto key = std::make_tuple(dev, func, kb * 4096 + (size_t)tpb);
  {
    std::lock_guard<std::mutex> lk(g_occ_mu);
    auto it = g_occ.find(key);
    if (it != g_occ.end()) return it->second;
  }
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, tpb, kb << 10) != cudaSuccess || per_sm < 1) { per_sm = 1; cudaGetLastError(); }
  std::lock_guard<std::mutex> lk(g_occ_mu);
  g_occ[key] = per_sm;
  return per_sm;
}
static std::set<std::pair<int, const void *>> g_optin;
int pc_smem_optin(const void *func, int bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_occ_mu);
  if (g_optin.count({dev, func})) return 0;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { fail(PC_E_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize): %s", cudaGetErrorString(e)); return PC_E_CUDA; }
  g_optin.insert({dev, func});
  return 0;
}

struct pc_ctx {
  int device = 0, sm_count = 148;
  uint8_t *d_genome = nullptr;
  uint32_t genome_len = 0;
  unsigned long long *ix_keys = nullptr;
  uint32_t *ix_pos = nullptr, *ix_bstart = nullptr;
  int ix_shift = 63;
  uint32_t ix_n = 0;
  int ix_word = 0;
  double depth_rate = 0.2;
  const uint32_t *gplanes = nullptr; uint32_t gplane_words = 0;
  PcGrowBuf genome_buf;         /* the buffers behind d_genome / ix_*: kept (and grown) across pc_genome_upload calls */
  PcIndexBufs ix_bufs;
  cudaStream_t up = nullptr;    /* uploads and index builds run here, not on the legacy stream */
};

/* PC_GUARD=1 (debugging aid; compute-sanitizer is not available on every pool): every device buffer of a stream is
 * allocated at exactly the size asked for, between two 4 KB guard bands filled with 0xA5 — the buffers themselves start
 * out as 0xA5 too — and pc_stream_sync checks the bands after every batch: an out-of-bounds WRITE by any kernel fails the
 * batch loudly, and a result that depends on bytes nobody wrote changes with the fill pattern. */
static const bool g_guard = getenv("PC_GUARD") != nullptr;
constexpr size_t GUARD = 4096;
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool in_slab = false;        /* carved from the stream's single start-up allocation: never freed on its own */
  void *g_base = nullptr; size_t g_bytes = 0;      /* guard mode: the real allocation and the exact size in use */
  void carve(uint8_t *&cursor, size_t bytes) { p = cursor; cap = bytes; in_slab = true; cursor += (bytes + 255u) & ~(size_t)255u; }
  int reserve_guarded(size_t bytes) {
    const size_t b16 = (bytes + 15u) & ~(size_t)15u;
    if (g_base && b16 == g_bytes) return 0;
    if (g_base) cudaFree(g_base);
    g_base = nullptr; p = nullptr; cap = 0;
    if (cudaMalloc(&g_base, b16 + 2 * GUARD) != cudaSuccess) { g_base = nullptr; return fail(PC_E_NOMEM, "%s", "cudaMalloc (guard mode)"); }
    cudaMemset(g_base, 0xA5, b16 + 2 * GUARD);
    cudaDeviceSynchronize();                          /* the fill runs on the legacy stream, which our non-blocking streams do not wait for */
    p = (uint8_t *)g_base + GUARD; cap = b16; g_bytes = b16;
    return 0;
  }
  int check_guards(const char *name) const {
    if (!g_base) return 0;
    static thread_local std::vector<uint8_t> h(2 * GUARD);
    if (cudaMemcpy(h.data(), g_base, GUARD, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(h.data() + GUARD, (uint8_t *)g_base + GUARD + g_bytes, GUARD, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(PC_E_CUDA, "%s", "guard read-back");
    for (size_t i = 0; i < 2 * GUARD; ++i)
      if (h[i] != 0xA5) {
        static thread_local char msg[160];
        snprintf(msg, sizeof msg, "PC_GUARD: a kernel wrote %s the '%s' buffer (%zu bytes): guard byte %zu", i < GUARD ? "BEFORE" : "PAST THE END of", name, g_bytes, i % GUARD);
        return fail(PC_E_CUDA, "%s", msg);
      }
    return 0;
  }
  int reserve(size_t bytes) {
    if (g_guard) return reserve_guarded(bytes);
    if (bytes <= cap) return 0;
    if (p && !in_slab) cudaFree(p);
    p = nullptr; in_slab = false;
    size_t want = std::max(bytes, cap * 2);
    __atomic_fetch_add(&g_dev_grows, 1ull, __ATOMIC_RELAXED);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cap = 0; return fail(PC_E_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cap = want;
    return 0;
  }
  void release() { if (g_base) cudaFree(g_base); else if (p && !in_slab) cudaFree(p); p = nullptr; g_base = nullptr; cap = 0; g_bytes = 0; in_slab = false; }
};

// Pinned host staging that only grows (job order and LCS block prefix travel to the device from here: a pageable
// source would make cudaMemcpyAsync wait for the stream).
struct PinBuf {
  uint32_t *p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    size_t want = std::max<size_t>(std::max(n, cap * 2), 1u << 16);
    if (cudaHostAlloc((void **)&p, want * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) { cap = 0; return fail(PC_E_NOMEM, "%s", "cudaHostAlloc failed"); }
    cap = want;
    return 0;
  }
};

struct Pending {          // what pc_stream_sync needs to re-run jobs that ran out of pool
  bool active = false, device_mode = false;
  const pc_job *jobs = nullptr;   // host copy
  int njobs = 0;
  int32_t *res = nullptr;         // host (or device in device_mode)
  uint8_t *var_out = nullptr;
  size_t var_out_bytes = 0;
  const uint8_t *d_arena = nullptr;
  const pc_job *d_jobs = nullptr;
  int32_t *d_res = nullptr;
  uint8_t *d_var = nullptr;
  /* multi-part batches (pc_submit_parts): where every part went inside the merged device buffers */
  struct Part { pc_part p; size_t j_base, a_base, v_base; };
  std::vector<Part> parts;
  std::vector<pc_job> merged;     // host copy of all jobs, built only when a retry round needs it
  pc_ctx *ctx = nullptr;          // genome the batch ran against
};

struct pc_stream {
  pc_ctx *ctx = nullptr;
  cudaStream_t s = nullptr;
  cudaEvent_t ev_block = nullptr;           /* cudaEventBlockingSync: stream_wait() sleeps on it when g_block_sync is on */
  DevBuf arena, jobs, idx, res, var, pool, lcs_best;
  void *slab = nullptr;
  unsigned long long *d_pool_need = nullptr, *h_pool_need = nullptr;   /* device counter + pinned mirror */
  int max_warps = 0;
  std::vector<uint32_t> h_bins;
  PinBuf pin_idx, pin_lcs, pin_seg, pin_parts;
  DevBuf d_parts;
  uint64_t n_retry_rounds = 0;
  uint32_t *last_slow_count = nullptr;      /* device: per-segment counts of jobs handed from k_myers to the wavefront kernel (profiling) */
  std::vector<uint16_t> h_key;
  std::vector<int32_t> h_status;
  Pending pend;
  bool timers = false;
  double op_ms[PC_OP_COUNT] = {0};
  uint64_t op_launches[PC_OP_COUNT] = {0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  /* side streams: the (op, class) segments of a SMALL batch are independent and each kernel is latency-bound (a grid
   * of a few dozen CTAs waiting for its longest job), so they run side by side instead of one after the other */
  static constexpr int NSIDE = 8;
  cudaStream_t side[NSIDE] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[NSIDE] = {};
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_pending;
  std::vector<cudaEvent_t> ev_free;
};

extern "C" const char *pc_last_error(void) { return g_err; }

extern "C" int pc_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(PC_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

extern "C" pc_ctx *pc_ctx_create(int device) {
  int n = pc_device_count();
  if (n <= 0) { if (n == 0) fail(PC_E_CUDA, "%s", "no CUDA device: libpintron_cuda has no CPU fallback"); return nullptr; }
  if (device < 0 || device >= n) { fail(PC_E_ARG, "%s", "device index out of range"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { fail(PC_E_CUDA, "%s", "cudaSetDevice failed"); return nullptr; }
  pc_ctx *c = new pc_ctx();
  c->device = device;
  cudaDeviceProp prop;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) c->sm_count = sms;
  else if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->up, cudaStreamNonBlocking) != cudaSuccess) { fail(PC_E_CUDA, "%s", "pc_ctx_create: stream"); delete c; return nullptr; }
  return c;
}

extern "C" void pc_ctx_destroy(pc_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  c->genome_buf.release();
  for (PcGrowBuf *b : {&c->ix_bufs.keys_in, &c->ix_bufs.keys_out, &c->ix_bufs.pos_in, &c->ix_bufs.pos_out, &c->ix_bufs.tmp, &c->ix_bufs.bstart, &c->ix_bufs.planes}) b->release();
  if (c->up) cudaStreamDestroy(c->up);
  delete c;
}

extern "C" int pc_genome_upload(pc_ctx *c, const char *genome, size_t len, int word_len, double depth_rate) {
  if (!c || !genome || word_len <= 0 || len >= 0xfffffff0ull) return fail(PC_E_ARG, "%s", "pc_genome_upload: bad argument");
  CU(cudaSetDevice(c->device));
  c->d_genome = nullptr; c->ix_keys = nullptr; c->ix_pos = nullptr; c->ix_bstart = nullptr;
  if (c->genome_buf.reserve(len + 16)) return fail(PC_E_NOMEM, "%s", "pc_genome_upload: device allocation failed");
  c->d_genome = (uint8_t *)c->genome_buf.p;
  CU(cudaMemsetAsync(c->d_genome + len, 0, 16, c->up));
  CU(cudaMemcpyAsync(c->d_genome, genome, len, cudaMemcpyHostToDevice, c->up));
  c->genome_len = (uint32_t)len;
  c->ix_word = word_len;
  c->depth_rate = depth_rate;
  int rc = pc_build_index(c->d_genome, c->genome_len, word_len, c->ix_bufs, &c->ix_keys, &c->ix_pos, &c->ix_n, &c->ix_bstart, &c->ix_shift, c->up);
  if (rc) return fail(rc, "%s", "pc_genome_upload: index build failed");
  c->gplanes = nullptr;
  if (pc_build_planes(c->d_genome, c->genome_len, c->ix_bufs.planes, &c->gplane_words, c->up)) return fail(PC_E_NOMEM, "%s", "pc_genome_upload: bit planes");
  c->gplanes = (const uint32_t *)c->ix_bufs.planes.p;
  CU(cudaStreamSynchronize(c->up));
  return 0;
}

extern "C" pc_stream *pc_stream_create(pc_ctx *c) {
  if (!c) { fail(PC_E_ARG, "%s", "pc_stream_create: null context"); return nullptr; }
  cudaSetDevice(c->device);
  pc_stream *st = new pc_stream();
  st->ctx = c;
  if (cudaStreamCreateWithFlags(&st->s, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&st->d_pool_need, 8) != cudaSuccess ||
      cudaHostAlloc((void **)&st->h_pool_need, 8, cudaHostAllocDefault) != cudaSuccess) {
    fail(PC_E_CUDA, "%s", "pc_stream_create: stream / counter allocation failed");
    if (st->s) cudaStreamDestroy(st->s);
    cudaFree(st->d_pool_need); cudaFreeHost(st->h_pool_need);
    delete st;
    return nullptr;
  }
  /* scratch pool and staging as ONE allocation made up front: cudaMalloc / cudaFree while other streams run
   * serialise the whole device.  Buffers that outgrow their share later get their own allocation. */
  static const size_t pool_mb = getenv("PC_POOL_MB") && atol(getenv("PC_POOL_MB")) >= 16 ? (size_t)atol(getenv("PC_POOL_MB")) : PC_POOL_MB_DEFAULT;
  const size_t sizes[7] = {pool_mb << 20, 8u << 20, 8u << 20, sizeof(pc_job) << 16, (sizeof(int32_t) * PC_RES_INTS) << 16, 4u << 16, 8u << 14};
  DevBuf *bufs[7] = {&st->pool, &st->arena, &st->var, &st->jobs, &st->res, &st->idx, &st->lcs_best};
  if (g_guard) {               /* no slab: every buffer on its own, exactly sized, between guard bands (only the pool is sized up front) */
    if (st->pool.reserve(sizes[0])) { pc_stream_destroy(st); return nullptr; }
  } else {
    size_t total = 0;
    for (size_t z : sizes) total += (z + 255u) & ~(size_t)255u;
    if (cudaMalloc(&st->slab, total) != cudaSuccess) { st->slab = nullptr; fail(PC_E_NOMEM, "%s", "pc_stream_create: device allocation failed"); pc_stream_destroy(st); return nullptr; }
    uint8_t *cur = (uint8_t *)st->slab;
    for (int i = 0; i < 7; ++i) bufs[i]->carve(cur, sizes[i]);
  }
  if (st->pin_idx.reserve(1u << 16) || st->pin_lcs.reserve(1u << 14)) { pc_stream_destroy(st); return nullptr; }
  bool ok = cudaEventCreateWithFlags(&st->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&st->ev_block, cudaEventBlockingSync | cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < pc_stream::NSIDE && ok; ++k)
    ok = cudaStreamCreateWithFlags(&st->side[k], cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&st->ev_join[k], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { fail(PC_E_CUDA, "%s", "pc_stream_create: side streams"); pc_stream_destroy(st); return nullptr; }
  return st;
}

extern "C" void pc_stream_destroy(pc_stream *st) {
  if (!st) return;
  cudaSetDevice(st->ctx->device);
  cudaStreamSynchronize(st->s);
  for (DevBuf *b : {&st->arena, &st->jobs, &st->idx, &st->res, &st->var, &st->pool, &st->lcs_best, &st->d_parts}) b->release();
  cudaFree(st->slab);
  cudaFree(st->d_pool_need); cudaFreeHost(st->h_pool_need);
  if (st->pin_idx.p) cudaFreeHost(st->pin_idx.p);
  if (st->pin_lcs.p) cudaFreeHost(st->pin_lcs.p);
  if (st->pin_seg.p) cudaFreeHost(st->pin_seg.p);
  if (st->pin_parts.p) cudaFreeHost(st->pin_parts.p);
  for (auto &e : st->ev_pending) { cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second); }
  for (auto e : st->ev_free) cudaEventDestroy(e);
  if (st->ev_fork) cudaEventDestroy(st->ev_fork);
  if (st->ev_block) cudaEventDestroy(st->ev_block);
  for (int k = 0; k < pc_stream::NSIDE; ++k) { if (st->ev_join[k]) cudaEventDestroy(st->ev_join[k]); if (st->side[k]) cudaStreamDestroy(st->side[k]); }
  cudaStreamDestroy(st->s);
  delete st;
}

extern "C" void *pc_host_alloc(size_t bytes) {
  void *p = nullptr;
  const double t0 = g_prof ? now_s() : 0;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { fail(PC_E_NOMEM, "%s", "cudaHostAlloc failed"); return nullptr; }
  if (g_prof) { __atomic_fetch_add(&g_host_allocs, 1ull, __ATOMIC_RELAXED); __atomic_fetch_add(&g_host_alloc_bytes, (unsigned long long)bytes, __ATOMIC_RELAXED); g_host_alloc_s += now_s() - t0; }
  return p;
}
double pc_int_peak_run(cudaStream_t s, int sm_count, float *ms_out);
/* Measured INT32 ALU-pipe throughput of this GPU in lane-operations per second (add + min chains). */
extern "C" double pc_measure_int_peak(pc_ctx *c) {
  if (!c) return 0.0;
  cudaSetDevice(c->device);
  float ms = 0;
  double best = 0;
  for (int r = 0; r < 3; ++r) { double ops = pc_int_peak_run(0, c->sm_count, &ms); if (ms > 0) best = std::max(best, ops / (ms * 1e-3)); }
  return best;
}
extern "C" void pc_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" uint64_t pc_launch_count(void) { return g_pc_launches; }
extern "C" void *pc_stream_cuda_stream(pc_stream *st) { return st ? (void *)st->s : nullptr; }
extern "C" void pc_stream_enable_timers(pc_stream *st, int on) { if (st) st->timers = on != 0; }

static void drain_events(pc_stream *st) {
  for (auto &e : st->ev_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e.second.first, e.second.second) == cudaSuccess) { st->op_ms[e.first] += ms; if (g_prof) g_op_ms[e.first] += ms; }
    st->ev_free.push_back(e.second.first);
    st->ev_free.push_back(e.second.second);
  }
  st->ev_pending.clear();
}

extern "C" void pc_stream_reset_timers(pc_stream *st) {
  if (!st) return;
  drain_events(st);
  for (int i = 0; i < PC_OP_COUNT; ++i) { st->op_ms[i] = 0; st->op_launches[i] = 0; }
}

extern "C" int pc_stream_op_time(pc_stream *st, int op, double *ms, uint64_t *launches) {
  if (!st || op < 0 || op >= PC_OP_COUNT) return fail(PC_E_ARG, "%s", "pc_stream_op_time: bad argument");
  drain_events(st);
  if (ms) *ms = st->op_ms[op];
  if (launches) *launches = st->op_launches[op];
  return 0;
}

static cudaError_t stream_wait(pc_stream *st) {
  if (g_block_sync && st->ev_block) {
    const cudaError_t e = cudaEventRecord(st->ev_block, st->s);
    return e != cudaSuccess ? e : cudaEventSynchronize(st->ev_block);
  }
  return cudaStreamSynchronize(st->s);
}

static cudaEvent_t get_event(pc_stream *st) {
  if (!st->ev_free.empty()) { cudaEvent_t e = st->ev_free.back(); st->ev_free.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}

// Launch the kernels for the jobs whose indices are in `sel` (nullptr = all njobs of the batch, which is already
// resident on the device).  One sequential pass over the host copy of the jobs gives every job its (op, class, cost)
// key and every (op, class) segment its size and longest strings; a counting sort then lays the job indices out
// segment by segment, heaviest first, for the persistent kernels.
static int launch_selected(pc_stream *st, pc_ctx *c, const pc_job *h_jobs, const uint32_t *sel, size_t nsel, const uint8_t *d_arena,
                           const pc_job *d_jobs, int32_t *d_res, uint8_t *d_var, size_t arena_bytes, size_t var_bytes,
                           bool force_device_order = false) {
  double tp = g_prof_host ? now_s() : 0;
  constexpr int NSEG = PC_ORDER_SEGS, NB = PC_ORDER_BINS;
  struct Seg { uint32_t n, max_a, max_b, max_t; unsigned long long lcs_blocks; } seg[NSEG];
  memset(seg, 0, sizeof seg);
  /* device buffer: [job order | list of the jobs the bit-parallel kernel leaves to the wavefront kernel | one counter per
   * segment | (device ordering only) keys, histogram work area, segment statistics] */
  const size_t idx_words = (nsel + 63) & ~(size_t)63;
  const bool on_device = sel == nullptr && (force_device_order || nsel >= PC_DEVICE_ORDER_MIN);
  const size_t extra_words = on_device ? idx_words / 2 + 3 * NB + 64 + NSEG * (sizeof(PcSegStat) / 4) : 0;
  if (st->idx.reserve((2 * idx_words + 64 + extra_words) * 4 + 256)) return PC_E_NOMEM;
  uint32_t *d_order = (uint32_t *)st->idx.p, *d_slow = d_order + idx_words, *d_slow_count = d_order + 2 * idx_words;
  st->last_slow_count = d_slow_count;
  CU(cudaMemsetAsync(d_slow_count, 0, 64 * sizeof(uint32_t), st->s));      /* every segment's hand-over counter, once per batch */
  PinBuf &order = st->pin_idx;
  if (on_device) {
    // millions of jobs already in HBM: key, histogram, scan and scatter run there (k_order.cu); the host only reads
    // the 40 segment records back.  The job check of pc_submit is part of the key kernel.
    uint16_t *d_keys = (uint16_t *)(d_slow_count + 64);
    uint32_t *d_work = (uint32_t *)(d_keys) + idx_words / 2;
    PcSegStat *d_seg = (PcSegStat *)(((uintptr_t)(d_work + 3 * NB + 2) + 15u) & ~(uintptr_t)15u);
    if (st->pin_seg.reserve(NSEG * (sizeof(PcSegStat) / 4) + 4)) return PC_E_NOMEM;
    pc_order_jobs(d_jobs, (int)nsel, arena_bytes, c->genome_len, var_bytes, PC_LCS_TPB, PC_LCS_MAX_S2, d_keys, d_work, d_seg, d_order,
                  st->s, c->sm_count);
    PcSegStat *h_seg = (PcSegStat *)st->pin_seg.p;
    uint32_t *h_invalid = st->pin_seg.p + NSEG * (sizeof(PcSegStat) / 4);
    CU(cudaMemcpyAsync(h_seg, d_seg, sizeof(PcSegStat) * NSEG, cudaMemcpyDeviceToHost, st->s));
    CU(cudaMemcpyAsync(h_invalid, d_work + 3 * NB + 1, 4, cudaMemcpyDeviceToHost, st->s));
    CU(stream_wait(st));
    if (*h_invalid) return fail(PC_E_ARG, "%s", "pc_submit: a job references bytes outside its buffers (or has an unknown op)");
    for (int sg = 0; sg < NSEG; ++sg) { seg[sg].n = h_seg[sg].n; seg[sg].max_a = h_seg[sg].max_a; seg[sg].max_b = h_seg[sg].max_b; seg[sg].max_t = h_seg[sg].max_t; seg[sg].lcs_blocks = h_seg[sg].lcs_blocks; }
    PROF(3, tp);
  } else {
    if (order.reserve(nsel + 1)) return PC_E_NOMEM;
    std::vector<uint32_t> &bins = st->h_bins;
    bins.assign(NB + 1, 0);
    std::vector<uint16_t> &key = st->h_key;
    key.resize(nsel);
    for (size_t q = 0; q < nsel; ++q) {
      const pc_job &j = h_jobs[sel ? sel[q] : q];
      if (j.op >= PC_OP_COUNT) return fail(PC_E_ARG, "%s", "pc_submit: unknown op");
      const int cls = pc_job_class(j);
      const int lg = pc_job_cost(j, cls);        // packed kernels run max(columns) steps per warp: ordered by that, finely
      const int sg = (int)j.op * 4 + cls;
      Seg &S = seg[sg];
      ++S.n; S.max_a = std::max(S.max_a, j.a_len); S.max_b = std::max(S.max_b, j.b_len);
      if (j.op == PC_OP_BORDERS) S.max_t = std::max(S.max_t, pc_borders_window(j));
      key[q] = (uint16_t)(sg * 64 + (63 - lg));
      ++bins[key[q] + 1];
    }
    for (int b = 0; b < NB; ++b) bins[b + 1] += bins[b];
    uint32_t *out = order.p;
    if (sel) for (size_t q = 0; q < nsel; ++q) out[bins[key[q]]++] = sel[q];
    else for (size_t q = 0; q < nsel; ++q) out[bins[key[q]]++] = (uint32_t)q;
    PROF(3, tp);
    CU(cudaMemcpyAsync(d_order, order.p, nsel * 4, cudaMemcpyHostToDevice, st->s));
  }
  if (g_prof)
    for (size_t q = 0; q < nsel && !on_device; ++q) {
      const pc_job &j = h_jobs[sel ? sel[q] : q];
      ++g_op_jobs[j.op]; g_op_suma[j.op] += j.a_len; g_op_sumb[j.op] += j.b_len; g_op_cells[j.op] += (unsigned long long)j.a_len * j.b_len;
      if (j.a_len > g_op_maxa[j.op]) g_op_maxa[j.op] = j.a_len;
      if (j.b_len > g_op_maxb[j.op]) g_op_maxb[j.op] = j.b_len;
    }
  CU(cudaMemsetAsync(st->d_pool_need, 0, 8, st->s));
  PcDevBatch B;
  B.arena = d_arena; B.genome = c->d_genome; B.genome_len = c->genome_len;
  B.jobs = d_jobs; B.res = d_res; B.var_out = d_var;
  B.pool = (uint8_t *)st->pool.p; B.pool_cap = st->pool.cap; B.pool_need = st->d_pool_need;
  B.slots = 1; B.max_warps = st->max_warps; B.n_dev = nullptr;
  B.gplanes = c->gplanes; B.gplane_words = c->gplane_words;
  B.ix_keys = c->ix_keys; B.ix_pos = c->ix_pos; B.ix_bstart = c->ix_bstart; B.ix_shift = c->ix_shift; B.ix_n = c->ix_n; B.ix_word = c->ix_word; B.depth_rate = c->depth_rate;
  int nseg_live = 0;
  for (int sg = 0; sg < NSEG; ++sg) nseg_live += seg[sg].n != 0;
  // small batch, several segments, no per-op timing wanted, not a retry round: fork over the side streams, each with
  // its own quarter of the scratch pool
  const bool multi = nseg_live > 1 && nsel < PC_MULTI_STREAM_MAX && !(st->timers || g_prof) && st->max_warps == 0 && !g_serial;
  const int NS = g_side_streams;
  if (multi) CU(cudaEventRecord(st->ev_fork, st->s));
  unsigned used_side = 0;
  // Launch order and stream of every live segment.  The kernels of a small batch are bound by the latency of their longest
  // job (measured per batch of ~17 k jobs, one after the other: SEED 250 us, AFFIX 170, BORDERS 260 over its four classes,
  // KBAND 200 and EDIT 190 over three, GAP 110, ALIGN 60, LCS 50), so the batch lasts as long as the longest chain on one
  // stream: the expensive segments are launched first (SEED used to go last) and each segment goes to the side stream with
  // the least expected work so far (longest-processing-time-first).
  size_t seg_off[NSEG + 1];
  seg_off[0] = 0;
  for (int sg = 0; sg < NSEG; ++sg) seg_off[sg + 1] = seg_off[sg] + seg[sg].n;
  auto expected_us = [](int sg) {
    const int op = sg / 4, cls = sg & 3;
    switch (op) {
      case PC_OP_SEED: return 250;
      case PC_OP_AFFIX: case PC_OP_SUFCUT: case PC_OP_PRECUT: return 170;
      case PC_OP_BORDERS: return cls == 3 ? 120 : 50;
      case PC_OP_KBAND: return 70;
      case PC_OP_EDIT: return 60;
      case PC_OP_GAP: return cls == 3 ? 100 : 55;
      case PC_OP_ALIGN: return 60;
      default: return 50;
    }
  };
  int launch_order[NSEG], n_live = 0;
  for (int sg = 0; sg < NSEG; ++sg) if (seg[sg].n) launch_order[n_live++] = sg;
  std::stable_sort(launch_order, launch_order + n_live, [&](int x, int y) { return expected_us(x) > expected_us(y); });
  int side_load[pc_stream::NSIDE] = {};
  for (int li = 0; li < n_live; ++li) {
    const int sg = launch_order[li];
    const size_t i = seg_off[sg];
    cudaStream_t ss = st->s;
    if (multi) {
      int k = 0;
      for (int q = 1; q < NS; ++q) if (side_load[q] < side_load[k]) k = q;
      side_load[k] += expected_us(sg);
      ss = st->side[k];
      if (!(used_side >> k & 1u)) { CU(cudaStreamWaitEvent(ss, st->ev_fork, 0)); used_side |= 1u << k; }
      const unsigned long long share = (st->pool.cap / NS) & ~255ull;
      B.pool = (uint8_t *)st->pool.p + share * (unsigned long long)k; B.pool_cap = share;
    }
    const uint32_t op = (uint32_t)(sg / 4);
    const int cls = sg & 3;
    const size_t j = i + seg[sg].n;
    const long long max_l1 = seg[sg].max_b;
    const int max_l2 = (int)seg[sg].max_a;
    B.idx = d_order + i;
    B.n = (int)(j - i);
    if (op == PC_OP_SEED && !c->d_genome) return fail(PC_E_ARG, "%s", "PC_OP_SEED before pc_genome_upload");
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (st->timers || g_prof) { e0 = get_event(st); e1 = get_event(st); cudaEventRecord(e0, ss); }
    const unsigned long long before = tl_pc_launches;
    if (op == PC_OP_SEED) {
      pc_launch_seed(B, max_l2, ss, c->sm_count);
    } else if (op == PC_OP_LCS) {
      // [best slots | block prefix] in one device buffer; the prefix comes from the host copy of the jobs
      const size_t best_b = (8ull * B.n + 255u) & ~(size_t)255u, prefix_b = (4ull * (B.n + 1) + 255u) & ~(size_t)255u;
      const uint64_t tiles_bound = on_device ? seg[sg].lcs_blocks : [&] { uint64_t z = 0; for (size_t q = i; q < j; ++q) { const pc_job &jb = h_jobs[order.p[q]]; z += (uint64_t)pc_lcs_blocks(jb.b_len, (int)jb.a_len); } return z; }();
      if (tiles_bound >= 0x7fffffffull) return fail(PC_E_RANGE, "%s", "LCS batch too large for one launch");
      if (st->lcs_best.reserve(best_b + prefix_b + 4ull * (2 * tiles_bound + 8))) { if (e0) { st->ev_free.push_back(e0); st->ev_free.push_back(e1); } return PC_E_NOMEM; }
      uint32_t *d_prefix = (uint32_t *)((uint8_t *)st->lcs_best.p + best_b);
      uint64_t tot = 0;
      if (on_device) {
        tot = seg[sg].lcs_blocks;
        pc_lcs_prefix(d_jobs, d_order + i, B.n, PC_LCS_TPB, PC_LCS_MAX_S2, d_prefix, ss);
      } else {
        PinBuf &pl = st->pin_lcs;
        if (pl.reserve((size_t)B.n + 1)) return PC_E_NOMEM;
        for (size_t q = i; q < j; ++q) {
          pl.p[q - i] = (uint32_t)tot;
          const pc_job &jb = h_jobs[order.p[q]];
          tot += (uint64_t)pc_lcs_blocks(jb.b_len, (int)jb.a_len);
        }
        pl.p[B.n] = (uint32_t)tot;
        CU(cudaMemcpyAsync(d_prefix, pl.p, 4ull * (B.n + 1), cudaMemcpyHostToDevice, ss));
      }
      if (tot >= 0x7fffffffull) return fail(PC_E_RANGE, "%s", "LCS batch too large for one launch");
      pc_launch_lcs(B, (unsigned long long *)st->lcs_best.p, d_prefix, (uint32_t)tot, max_l2, (uint32_t *)((uint8_t *)st->lcs_best.p + best_b + prefix_b), ss);
    } else if (op == PC_OP_GAP && cls < 3) {
      pc_launch_gap_pairs(cls, B, (int)max_l1, d_slow_count + sg, ss, c->sm_count);
    } else if (op == PC_OP_BORDERS && cls < 3) {
      pc_launch_borders_packed(cls, B, (int)std::min<uint32_t>(seg[sg].max_t, PC_BORDERS_FAST_MAX_T), ss, c->sm_count);
    } else if (op == PC_OP_BORDERS) {
      // taller than the packed classes: row-chunked packed sweep; what does not fit 16-bit scores goes to the wavefront kernel
      pc_launch_borders_chunked(B, d_slow + i, d_slow_count + sg, ss, c->sm_count);
      PcDevBatch S = B;
      S.idx = d_slow + i; S.n_dev = d_slow_count + sg;
      pc_launch_dp((int)op, S, ss, c->sm_count);
    } else if (op == PC_OP_ALIGN && g_align_bp) {
      // one job per thread, bit-parallel with a traceback from stored delta vectors (align_core.h); long ESTs, bytes outside
      // the alphabet and jobs whose columns do not fit the thread's share of the pool go on to the wavefront kernel
      pc_launch_align_bp(B, d_slow + i, d_slow_count + sg, ss, c->sm_count);
      PcDevBatch S = B;
      S.idx = d_slow + i; S.n_dev = d_slow_count + sg;
      pc_launch_dp((int)op, S, ss, c->sm_count);
    } else if ((op == PC_OP_EDIT || op == PC_OP_KBAND) && cls < 3) {
      // one job per thread, bit-parallel; what it cannot answer bit-exactly is listed for the wavefront kernel
      pc_launch_myers((int)op, cls, B, d_slow + i, d_slow_count + sg, ss, c->sm_count);
      PcDevBatch S = B;
      S.idx = d_slow + i; S.n_dev = d_slow_count + sg;
      pc_launch_dp((int)op, S, ss, c->sm_count);
    } else {
      pc_launch_dp((int)op, B, ss, c->sm_count);
    }
    st->op_launches[op] += tl_pc_launches - before;
    if (g_guard) {                                       /* debugging mode: name the segment whose kernel faulted */
      const cudaError_t ge = cudaStreamSynchronize(ss);
      if (ge != cudaSuccess) {
        static thread_local char msg[200];
        snprintf(msg, sizeof msg, "PC_GUARD: kernel of op %u class %d (%d jobs, longest a %d, longest b %lld) faulted: %s", op, cls, B.n, max_l2, max_l1, cudaGetErrorString(ge));
        return fail(PC_E_CUDA, "%s", msg);
      }
    }
    if (g_prof) g_op_launches[op] += 1;
    if (st->timers || g_prof) { cudaEventRecord(e1, ss); st->ev_pending.push_back({(int)op, {e0, e1}}); }
  }
  if (multi)
    for (int k = 0; k < NS; ++k)
      if (used_side >> k & 1u) { CU(cudaEventRecord(st->ev_join[k], st->side[k])); CU(cudaStreamWaitEvent(st->s, st->ev_join[k], 0)); }
  CU(cudaMemcpyAsync(st->h_pool_need, st->d_pool_need, 8, cudaMemcpyDeviceToHost, st->s));
  CU(cudaGetLastError());
  PROF(4, tp);
  return 0;
}

static int check_jobs(const pc_ctx *c, const pc_job *jobs, int njobs, size_t arena_bytes, size_t var_out_bytes) {
  const size_t glen = c->genome_len;
  for (int i = 0; i < njobs; ++i) {
    const pc_job &j = jobs[i];
    if ((size_t)j.a_off + j.a_len > arena_bytes) return fail(PC_E_ARG, "%s", "pc_submit: job string a outside the arena");
    if (j.op != PC_OP_SEED) {
      const size_t lim = (j.flags & PC_B_IN_GENOME) ? glen : arena_bytes;
      if ((size_t)j.b_off + j.b_len > lim) return fail(PC_E_ARG, "%s", "pc_submit: job string b out of range");
    } else if (j.p1 == PC_SEED_BUILD_MEG) {
      if ((j.flags & PC_B_IN_GENOME) || j.b_len != sizeof(pc_meg_cfg) || (size_t)j.b_off + j.b_len > arena_bytes || j.p0 < 1)
        return fail(PC_E_ARG, "%s", "pc_submit: PC_SEED_BUILD_MEG wants one struct pc_meg_cfg as string b (in the arena) and p0 >= 1");
    }
    if (j.op == PC_OP_ALIGN || j.op == PC_OP_GAP) {
      if ((size_t)j.out_off + j.out_cap > var_out_bytes) return fail(PC_E_ARG, "%s", "pc_submit: ops region outside var_out");
    } else if (j.op == PC_OP_SEED) {
      if ((j.out_off & 3u) || (size_t)j.out_off + 12ull * j.out_cap > var_out_bytes)
        return fail(PC_E_ARG, "%s", "pc_submit: triples region misaligned or outside var_out");
    }
  }
  return 0;
}

extern "C" int pc_submit(pc_stream *st, const uint8_t *arena, size_t arena_bytes, const pc_job *jobs, int njobs,
                         int32_t *res, uint8_t *var_out, size_t var_out_bytes) {
  if (!st || njobs < 0 || (njobs && (!jobs || !res))) return fail(PC_E_ARG, "%s", "pc_submit: bad argument");
  if (st->pend.active) return fail(PC_E_ARG, "%s", "pc_submit: previous batch not synced");
  if (njobs == 0) return 0;
  double tp = g_prof_host ? now_s() : 0;
  CU(cudaSetDevice(st->ctx->device));
  int rc = (size_t)njobs >= PC_DEVICE_ORDER_MIN ? 0 : check_jobs(st->ctx, jobs, njobs, arena_bytes, var_out_bytes);   /* large batches: checked by the key kernel */
  if (rc) return rc;
  if (g_capture) {
    std::lock_guard<std::mutex> lk(g_capture_mu);
    const uint32_t n32 = (uint32_t)njobs; const uint64_t ab = arena_bytes;
    fwrite(&n32, 4, 1, g_capture); fwrite(&ab, 8, 1, g_capture);
    fwrite(jobs, sizeof(pc_job), (size_t)njobs, g_capture); fwrite(arena, 1, arena_bytes, g_capture);
  }
  PROF(0, tp);
  // one spare readable byte after the arena: general_refine_borders reads t[len_t] (refine.c:362-374 adaptor)
  if (st->arena.reserve(arena_bytes + 16) || st->jobs.reserve(sizeof(pc_job) * (size_t)njobs) ||
      st->res.reserve(sizeof(int32_t) * PC_RES_INTS * (size_t)njobs) || st->var.reserve(var_out_bytes + 16))
    return PC_E_NOMEM;
  PROF(1, tp);
  if (arena_bytes) CU(cudaMemcpyAsync(st->arena.p, arena, arena_bytes, cudaMemcpyHostToDevice, st->s));
  CU(cudaMemsetAsync((uint8_t *)st->arena.p + arena_bytes, 0, 16, st->s));
  CU(cudaMemcpyAsync(st->jobs.p, jobs, sizeof(pc_job) * (size_t)njobs, cudaMemcpyHostToDevice, st->s));
  PROF(2, tp);
  rc = launch_selected(st, st->ctx, jobs, nullptr, (size_t)njobs, (const uint8_t *)st->arena.p, (const pc_job *)st->jobs.p, (int32_t *)st->res.p,
                       (uint8_t *)st->var.p, arena_bytes, var_out_bytes);
  if (rc) return rc;
  Pending &P = st->pend;
  P.active = true; P.device_mode = false; P.jobs = jobs; P.njobs = njobs; P.res = res; P.var_out = var_out;
  P.var_out_bytes = var_out_bytes; P.d_arena = (const uint8_t *)st->arena.p; P.d_jobs = (const pc_job *)st->jobs.p;
  P.d_res = (int32_t *)st->res.p; P.d_var = (uint8_t *)st->var.p; P.parts.clear(); P.ctx = st->ctx;
  if (g_prof_host) tp = now_s();
  CU(cudaMemcpyAsync(res, st->res.p, sizeof(int32_t) * PC_RES_INTS * (size_t)njobs, cudaMemcpyDeviceToHost, st->s));
  if (var_out_bytes) CU(cudaMemcpyAsync(var_out, st->var.p, var_out_bytes, cudaMemcpyDeviceToHost, st->s));
  PROF(5, tp);
  return 0;
}

extern "C" int pc_submit_device(pc_stream *st, const uint8_t *d_arena, size_t arena_bytes, const pc_job *d_jobs,
                                const pc_job *h_jobs, int njobs, int32_t *d_res, uint8_t *d_var_out, size_t var_out_bytes) {
  if (!st || njobs < 0 || (njobs && (!d_jobs || !d_res))) return fail(PC_E_ARG, "%s", "pc_submit_device: bad argument");
  if (st->pend.active) return fail(PC_E_ARG, "%s", "pc_submit_device: previous batch not synced");
  if (njobs == 0) return 0;
  CU(cudaSetDevice(st->ctx->device));
  /* no host copy of the jobs: validated, keyed and ordered on the device whatever the size (the engine's path) */
  int rc = ((size_t)njobs >= PC_DEVICE_ORDER_MIN || !h_jobs) ? 0 : check_jobs(st->ctx, h_jobs, njobs, arena_bytes, var_out_bytes);
  if (rc) return rc;
  st->pend.merged.clear();
  rc = launch_selected(st, st->ctx, h_jobs, nullptr, (size_t)njobs, d_arena, d_jobs, d_res, d_var_out, arena_bytes, var_out_bytes, h_jobs == nullptr);
  if (rc) return rc;
  Pending &P = st->pend;
  P.active = true; P.device_mode = true; P.jobs = h_jobs; P.njobs = njobs; P.res = nullptr; P.var_out = nullptr;
  P.var_out_bytes = var_out_bytes; P.d_arena = d_arena; P.d_jobs = d_jobs; P.d_res = d_res; P.d_var = d_var_out; P.parts.clear(); P.ctx = st->ctx;
  return 0;
}

// D2H of every part's results out of the merged device buffers
static int copy_parts_back(pc_stream *st) {
  Pending &P = st->pend;
  for (const Pending::Part &q : P.parts) {
    CU(cudaMemcpyAsync(q.p.res, P.d_res + q.j_base * PC_RES_INTS, sizeof(int32_t) * PC_RES_INTS * (size_t)q.p.njobs, cudaMemcpyDeviceToHost, st->s));
    if (q.p.var_out_bytes) CU(cudaMemcpyAsync(q.p.var_out, P.d_var + q.v_base, q.p.var_out_bytes, cudaMemcpyDeviceToHost, st->s));
  }
  return 0;
}

// Several (arena, jobs, res, var) quadruples as ONE device batch: the parts are laid one after the other in the stream's
// device buffers (16-byte aligned), a kernel rebases the job offsets, and from there on the batch is an ordinary
// device-resident one (validated, keyed and ordered on the device whatever its size).  What the engine runs for the
// lanes it merged; results go back part by part.
extern "C" int pc_submit_parts(pc_stream *st, pc_ctx *genome_ctx, const pc_part *parts, int nparts) {
  if (!st || nparts < 0 || (nparts && !parts)) return fail(PC_E_ARG, "%s", "pc_submit_parts: bad argument");
  if (st->pend.active) return fail(PC_E_ARG, "%s", "pc_submit_parts: previous batch not synced");
  pc_ctx *c = genome_ctx ? genome_ctx : st->ctx;
  if (c->device != st->ctx->device) return fail(PC_E_ARG, "%s", "pc_submit_parts: genome context and stream live on different devices");
  Pending &P = st->pend;
  P.parts.clear(); P.merged.clear();
  size_t nj = 0, na = 0, nv = 0;
  for (int q = 0; q < nparts; ++q) {
    const pc_part &p = parts[q];
    if (p.njobs < 0 || (p.njobs && (!p.jobs || !p.res)) || (p.arena_bytes && !p.arena) || (p.var_out_bytes && !p.var_out))
      return fail(PC_E_ARG, "%s", "pc_submit_parts: bad part");
    if (p.njobs == 0) continue;
    P.parts.push_back({p, nj, na, nv});
    nj += (size_t)p.njobs;
    na = (na + p.arena_bytes + 16 + 15u) & ~(size_t)15u;       /* 16 spare bytes per part: BORDERS reads t[len_t], word loads run past the end */
    nv = (nv + p.var_out_bytes + 15u) & ~(size_t)15u;
  }
  if (nj == 0) return 0;
  if (g_capture) {                                            /* every part as the batch its lane was */
    std::lock_guard<std::mutex> lk(g_capture_mu);
    const uint32_t mark = 0xffffffffu; const uint64_t npart = P.parts.size();      /* "the next npart records ran as one device batch" */
    fwrite(&mark, 4, 1, g_capture); fwrite(&npart, 8, 1, g_capture);
    for (const Pending::Part &q : P.parts) {
      const uint32_t n32 = (uint32_t)q.p.njobs; const uint64_t ab = q.p.arena_bytes;
      fwrite(&n32, 4, 1, g_capture); fwrite(&ab, 8, 1, g_capture);
      fwrite(q.p.jobs, sizeof(pc_job), (size_t)q.p.njobs, g_capture); fwrite(q.p.arena, 1, q.p.arena_bytes, g_capture);
    }
  }
  if (nj >= 0x7fffffffull || na >= 0xfff00000ull || nv >= 0xfff00000ull) return fail(PC_E_RANGE, "%s", "pc_submit_parts: merged batch exceeds 32-bit offsets");
  CU(cudaSetDevice(c->device));
  const size_t np = P.parts.size();
  if (st->arena.reserve(na + 16) || st->jobs.reserve(sizeof(pc_job) * nj) || st->res.reserve(sizeof(int32_t) * PC_RES_INTS * nj) ||
      st->var.reserve(nv + 16) || st->d_parts.reserve(4 * (3 * np + 4)) || st->pin_parts.reserve(3 * np + 4))
    return PC_E_NOMEM;
  uint32_t *tab = st->pin_parts.p;
  for (size_t q = 0; q < np; ++q) {
    const Pending::Part &pp = P.parts[q];
    tab[3 * q] = (uint32_t)pp.j_base; tab[3 * q + 1] = (uint32_t)pp.a_base; tab[3 * q + 2] = (uint32_t)pp.v_base;
    if (pp.p.arena_bytes) CU(cudaMemcpyAsync((uint8_t *)st->arena.p + pp.a_base, pp.p.arena, pp.p.arena_bytes, cudaMemcpyHostToDevice, st->s));
    CU(cudaMemcpyAsync((pc_job *)st->jobs.p + pp.j_base, pp.p.jobs, sizeof(pc_job) * (size_t)pp.p.njobs, cudaMemcpyHostToDevice, st->s));
  }
  tab[3 * np] = (uint32_t)nj;
  CU(cudaMemcpyAsync(st->d_parts.p, tab, 4 * (3 * np + 1), cudaMemcpyHostToDevice, st->s));
  pc_rebase_jobs((pc_job *)st->jobs.p, (int)nj, (const uint32_t *)st->d_parts.p, (int)np, st->s, c->sm_count);
  // Small merged batches (a few lanes of a few hundred fibers: most of what the host sends, and all of it when a handful of
  // long reads are the only ones left) are validated, keyed and ordered right here from the lanes' own job arrays: the
  // device ordering path costs three launches, a read-back of the segment table and a synchronisation in mid-batch.
  const pc_job *h_jobs = nullptr;
  if (nj < PC_PARTS_HOST_ORDER_MAX) {
    P.merged.resize(nj);
    for (const Pending::Part &q : P.parts) {
      memcpy(P.merged.data() + q.j_base, q.p.jobs, sizeof(pc_job) * (size_t)q.p.njobs);
      const int bad = check_jobs(c, q.p.jobs, q.p.njobs, q.p.arena_bytes, q.p.var_out_bytes);       // offsets are still lane-relative here
      if (bad) { P.parts.clear(); P.merged.clear(); cudaStreamSynchronize(st->s); return bad; }
    }
    h_jobs = P.merged.data();
  }
  int rc = launch_selected(st, c, h_jobs, nullptr, nj, (const uint8_t *)st->arena.p, (const pc_job *)st->jobs.p, (int32_t *)st->res.p,
                           (uint8_t *)st->var.p, na, nv, h_jobs == nullptr);
  if (rc) { P.parts.clear(); return rc; }
  P.active = true; P.device_mode = true; P.jobs = nullptr; P.njobs = (int)nj; P.res = nullptr; P.var_out = nullptr;
  P.var_out_bytes = nv; P.d_arena = (const uint8_t *)st->arena.p; P.d_jobs = (const pc_job *)st->jobs.p;
  P.d_res = (int32_t *)st->res.p; P.d_var = (uint8_t *)st->var.p; P.ctx = c;
  return copy_parts_back(st);
}

extern "C" int pc_stream_sync(pc_stream *st) {
  if (!st) return fail(PC_E_ARG, "%s", "pc_stream_sync: null stream");
  CU(cudaSetDevice(st->ctx->device));
  double tp = g_prof_host ? now_s() : 0;
  CU(stream_wait(st));
  PROF(6, tp);
  if (g_prof) drain_events(st);
  if (g_prof && st->last_slow_count) {
    uint32_t h[64];
    if (cudaMemcpy(h, st->last_slow_count, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess)
      for (int sg = 0; sg < PC_ORDER_SEGS; ++sg) g_op_slow[sg / 4] += h[sg];
    st->last_slow_count = nullptr;
  }
  Pending &P = st->pend;
  if (!P.active) return 0;
  // every job that ran out of scratch also raised the pool_need counter (pc_pool_alloc, k_gap): zero = nothing to re-run,
  // and the statuses need not be read at all
  if (g_guard) {
    const char *nm[8] = {"arena", "jobs", "idx", "res", "var", "pool", "lcs_best", "parts"};
    DevBuf *bb[8] = {&st->arena, &st->jobs, &st->idx, &st->res, &st->var, &st->pool, &st->lcs_best, &st->d_parts};
    for (int i = 0; i < 8; ++i) if (bb[i]->check_guards(nm[i])) { P.active = false; return PC_E_CUDA; }
  }
  if (*st->h_pool_need == 0) { P.active = false; return 0; }
  // Jobs whose scratch did not fit their warp's pool slot are re-run with fewer warps (= larger slots: the launchers
  // give exactly max_warps warps a slot); the pool itself grows only when a single job needs more than all of it.
  const pc_job *h_jobs = P.jobs;
  if (!P.parts.empty() && P.merged.size() == (size_t)P.njobs) h_jobs = P.merged.data();      // small merged batch: copy already there
  else if (!P.parts.empty()) {      // merged batch: the host copy of the jobs is only put together now that it is needed
    P.merged.resize((size_t)P.njobs);     // (lengths, op and parameters only: the offsets stay un-rebased and are not read)
    for (const Pending::Part &q : P.parts) memcpy(P.merged.data() + q.j_base, q.p.jobs, sizeof(pc_job) * (size_t)q.p.njobs);
    h_jobs = P.merged.data();
  } else if (!h_jobs) {             // device-resident batch submitted without a host copy: fetch the jobs for the retry's ordering pass
    P.merged.resize((size_t)P.njobs);
    CU(cudaMemcpy(P.merged.data(), P.d_jobs, sizeof(pc_job) * (size_t)P.njobs, cudaMemcpyDeviceToHost));
    h_jobs = P.merged.data();
  }
  size_t left = 0;
  for (int round = 0; round < 64; ++round) {
    std::vector<uint32_t> redo;
    if (P.device_mode) {
      // device-resident results: the failed jobs are listed by a kernel, only the list comes back
      if (st->lcs_best.reserve(4ull * ((size_t)P.njobs + 4))) { P.active = false; return PC_E_NOMEM; }
      uint32_t *d_list = (uint32_t *)st->lcs_best.p + 4, *d_count = (uint32_t *)st->lcs_best.p;
      pc_collect_status(P.d_res, P.njobs, PC_E_POOL, d_list, d_count, st->s, st->ctx->sm_count);
      uint32_t cnt = 0;
      CU(cudaMemcpyAsync(&cnt, d_count, 4, cudaMemcpyDeviceToHost, st->s));
      CU(cudaStreamSynchronize(st->s));
      redo.resize(cnt);
      if (cnt) CU(cudaMemcpy(redo.data(), d_list, 4ull * cnt, cudaMemcpyDeviceToHost));
      std::sort(redo.begin(), redo.end());
    } else {
      const int32_t *status = P.res;
      for (int i = 0; i < P.njobs; ++i) if (status[(size_t)i * PC_RES_INTS] == PC_E_POOL) redo.push_back((uint32_t)i);
    }
    left = redo.size();
    if (redo.empty()) break;
    ++st->n_retry_rounds;
    if (g_prof) { ++g_retry_rounds; g_retry_jobs += redo.size(); }
    const unsigned long long need = std::max<unsigned long long>(*st->h_pool_need, 4096) + 4096;
    if (need > st->pool.cap) {
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      if (need > st->pool.cap + free_b - (free_b >> 3)) {
        P.active = false; st->max_warps = 0;
        return fail(PC_E_NOMEM, "%s", "a single job needs more scratch than the device has free");
      }
      size_t want = std::min<size_t>(need * std::min<size_t>(redo.size(), 32), st->pool.cap + free_b / 2);
      if (g_prof) ++g_pool_grows;
      int rc = st->pool.reserve(std::max<size_t>(want, need));
      if (rc) { P.active = false; st->max_warps = 0; return rc; }
    }
    st->max_warps = (int)std::max<unsigned long long>(1, std::min<unsigned long long>(st->pool.cap / need, 1u << 20));
    int rc = launch_selected(st, P.ctx, h_jobs, redo.data(), redo.size(), P.d_arena, P.d_jobs, P.d_res, P.d_var, 0, P.var_out_bytes);
    st->max_warps = 0;
    if (rc) { P.active = false; return rc; }
    if (!P.parts.empty()) { rc = copy_parts_back(st); if (rc) { P.active = false; return rc; } }
    else if (!P.device_mode) {
      CU(cudaMemcpyAsync(P.res, P.d_res, sizeof(int32_t) * PC_RES_INTS * (size_t)P.njobs, cudaMemcpyDeviceToHost, st->s));
      if (P.var_out_bytes) CU(cudaMemcpyAsync(P.var_out, P.d_var, P.var_out_bytes, cudaMemcpyDeviceToHost, st->s));
    }
    CU(cudaStreamSynchronize(st->s));
    if (*st->h_pool_need == 0) { left = 0; break; }
  }
  P.active = false;
  if (left) return fail(PC_E_NOMEM, "%s", "jobs still short of scratch after 64 retry rounds");
  return 0;
}

// ---- per-routine entry points ------------------------------------------------------------------------------
static int run_typed(pc_stream *st, uint32_t op_a, uint32_t op_b, const uint8_t *arena, size_t ab, const pc_job *jobs, int n,
                     int32_t *res, uint8_t *var, size_t vb) {
  for (int i = 0; i < n; ++i)
    if (jobs[i].op != op_a && jobs[i].op != op_b) return fail(PC_E_ARG, "%s", "typed batch: job with a different op");
  int rc = pc_submit(st, arena, ab, jobs, n, res, var, vb);
  if (rc) return rc;
  return pc_stream_sync(st);
}

extern "C" int pc_compute_alignment_batch(pc_stream *st, const uint8_t *a, size_t ab, const pc_job *j, int n, int32_t *r, uint8_t *ops, size_t ob) {
  return run_typed(st, PC_OP_ALIGN, PC_OP_ALIGN, a, ab, j, n, r, ops, ob);
}
extern "C" int pc_kband_edit_distance_batch(pc_stream *st, const uint8_t *a, size_t ab, const pc_job *j, int n, int32_t *r) {
  return run_typed(st, PC_OP_KBAND, PC_OP_KBAND, a, ab, j, n, r, nullptr, 0);
}
extern "C" int pc_edit_distance_batch(pc_stream *st, const uint8_t *a, size_t ab, const pc_job *j, int n, int32_t *r) {
  return run_typed(st, PC_OP_EDIT, PC_OP_EDIT, a, ab, j, n, r, nullptr, 0);
}
extern "C" int pc_refine_borders_batch(pc_stream *st, const uint8_t *a, size_t ab, const pc_job *j, int n, int32_t *r) {
  return run_typed(st, PC_OP_BORDERS, PC_OP_BORDERS, a, ab, j, n, r, nullptr, 0);
}
extern "C" int pc_gap_alignment_batch(pc_stream *st, const uint8_t *a, size_t ab, const pc_job *j, int n, int32_t *r, uint8_t *ops, size_t ob) {
  return run_typed(st, PC_OP_GAP, PC_OP_GAP, a, ab, j, n, r, ops, ob);
}
extern "C" int pc_longest_affix_batch(pc_stream *st, const uint8_t *a, size_t ab, const pc_job *j, int n, int32_t *r) {
  return run_typed(st, PC_OP_AFFIX, PC_OP_AFFIX, a, ab, j, n, r, nullptr, 0);
}
extern "C" int pc_best_cut_batch(pc_stream *st, const uint8_t *a, size_t ab, const pc_job *j, int n, int32_t *r) {
  return run_typed(st, PC_OP_SUFCUT, PC_OP_PRECUT, a, ab, j, n, r, nullptr, 0);
}
extern "C" int pc_longest_common_factor_batch(pc_stream *st, const uint8_t *a, size_t ab, const pc_job *j, int n, int32_t *r) {
  return run_typed(st, PC_OP_LCS, PC_OP_LCS, a, ab, j, n, r, nullptr, 0);
}
extern "C" int pc_build_vertex_set_batch(pc_stream *st, const uint8_t *a, size_t ab, const pc_job *j, int n, int32_t *r, uint8_t *tr, size_t tb) {
  return run_typed(st, PC_OP_SEED, PC_OP_SEED, a, ab, j, n, r, tr, tb);
}
