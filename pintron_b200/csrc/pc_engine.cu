// pc_engine.cu — the batch engine (include/pintron_engine.h): per GPU, pinned shared-memory segments that hold the
// clients' lanes, and submission threads that run every lane posted at the same moment as ONE device batch.
//
// Replaces, together with the host's fibers, the sequential hot loop of the reference (src/main-est-fact.c:249-291):
// there one thread walks the ESTs and calls each DP routine in turn; here the DP requests of thousands of ESTs in
// flight arrive lane by lane and leave as merged batches.  Round 1 let every worker thread drive its own CUDA stream
// (40 launches + 4 copies per thread-batch, all contending on the driver lock: the submit calls cost 2.7x the per-EST
// code).  Now only the engine's threads talk to CUDA, whatever the number of workers or client processes.
#include "pintron_engine.h"
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

extern thread_local unsigned long long tl_pc_launches;
int pc_set_error(int code, const char *msg);      // pc_api.cu: sets the thread-local pc_last_error() text

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
constexpr size_t ALIGN = 4096;
size_t up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Segment {
  int fd = -1;
  uint8_t *base = nullptr;
  size_t bytes = 0;
  bool registered = false;
  std::map<size_t, size_t> free_ranges;        // offset -> length

  size_t alloc(size_t n) {                     // first fit; (size_t)-1 when nothing is large enough
    for (auto it = free_ranges.begin(); it != free_ranges.end(); ++it) {
      if (it->second < n) continue;
      const size_t off = it->first, len = it->second;
      free_ranges.erase(it);
      if (len > n) free_ranges[off + n] = len - n;
      return off;
    }
    return (size_t)-1;
  }
  void release(size_t off, size_t n) {
    auto it = free_ranges.emplace(off, n).first;
    auto nx = std::next(it);
    if (nx != free_ranges.end() && it->first + it->second == nx->first) { it->second += nx->second; free_ranges.erase(nx); }
    if (it != free_ranges.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) { pv->second += it->second; free_ranges.erase(it); }
    }
  }
};

struct Session {
  uint32_t id = 0;
  int gpu = 0;
  pc_ctx *ctx = nullptr;
  std::vector<uint32_t> lanes;
  std::vector<std::pair<uint32_t, std::pair<size_t, size_t>>> slabs;   // per lane: segment, (offset, bytes)
  bool closing = false;
  pc_session_stats stats{};
};

struct Gpu {
  int device = 0;
  pc_ctx *base_ctx = nullptr;
  std::vector<Segment> segs;
  pce_hdr *hdr = nullptr;
  std::mutex mu;                               // sessions, segments, lane ownership
  std::map<uint32_t, Session *> sessions;
  std::vector<std::thread> threads;
  std::vector<pc_ctx *> idle_ctx;              // contexts of closed sessions: their device buffers serve the next genome
  size_t default_seg = 0;
};

}  // namespace

struct pc_engine {
  std::vector<Gpu *> gpus;
  std::atomic<bool> stop{false};
  std::atomic<uint32_t> next_session{1};
  std::atomic<int> timers{0};
};

namespace {

int add_segment(Gpu *g, size_t bytes) {        // g->mu held (or engine not yet running)
  if (g->segs.size() >= PCE_MAX_SEGMENTS) return pc_set_error(PC_E_NOMEM, "engine: too many shared-memory segments");
  bytes = up(bytes, 1u << 21);
  Segment s;
  s.fd = memfd_create("pintron-lanes", MFD_CLOEXEC);
  if (s.fd < 0 || ftruncate(s.fd, (off_t)bytes) != 0) { if (s.fd >= 0) close(s.fd); return pc_set_error(PC_E_NOMEM, "engine: memfd_create / ftruncate failed"); }
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, s.fd, 0);
  if (p == MAP_FAILED) { close(s.fd); return pc_set_error(PC_E_NOMEM, "engine: mmap of a lane segment failed"); }
  s.base = (uint8_t *)p; s.bytes = bytes;
  cudaSetDevice(g->device);
  // pinned once, here: every later copy to / from a lane is a plain asynchronous DMA
  if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) == cudaSuccess) s.registered = true;
  else { cudaGetLastError(); fprintf(stderr, "* WARN  pc_engine: cudaHostRegister of a %zu MB lane segment failed; copies will be staged\n", bytes >> 20); }
  size_t first = 0;
  if (g->segs.empty()) {                       // segment 0 starts with the lane table
    first = PCE_HDR_BYTES;
    g->hdr = (pce_hdr *)p;
    memset(g->hdr, 0, sizeof(pce_hdr));
    g->hdr->magic = PCE_MAGIC; g->hdr->version = PCE_VERSION;
  }
  s.free_ranges[first] = bytes - first;
  g->segs.push_back(std::move(s));
  return (int)g->segs.size() - 1;
}

// slab for one lane: [arena | jobs | res | var]; returns the segment index or a negative status
int alloc_slab(Gpu *g, uint64_t arena_cap, uint32_t jobs_cap, uint64_t var_cap, size_t *off_out, size_t *bytes_out) {
  const size_t a = up(arena_cap + 64, 256), j = up(sizeof(pc_job) * (size_t)jobs_cap, 256), r = up(sizeof(int32_t) * PC_RES_INTS * (size_t)jobs_cap, 256),
               v = up(var_cap + 64, 256);
  const size_t total = up(a + j + r + v, ALIGN);
  for (size_t k = 0; k < g->segs.size(); ++k) {
    const size_t off = g->segs[k].alloc(total);
    if (off != (size_t)-1) { *off_out = off; *bytes_out = total; return (int)k; }
  }
  const int k = add_segment(g, std::max(g->default_seg, total + PCE_HDR_BYTES));
  if (k < 0) return k;
  const size_t off = g->segs[(size_t)k].alloc(total);
  if (off == (size_t)-1) return pc_set_error(PC_E_NOMEM, "engine: lane slab does not fit a fresh segment");
  *off_out = off; *bytes_out = total;
  return k;
}

void lay_out(pce_lane *l, int seg, size_t off, uint64_t arena_cap, uint32_t jobs_cap, uint64_t var_cap) {
  const size_t a = up(arena_cap + 64, 256), j = up(sizeof(pc_job) * (size_t)jobs_cap, 256), r = up(sizeof(int32_t) * PC_RES_INTS * (size_t)jobs_cap, 256);
  l->seg = (uint32_t)seg; l->jobs_cap = jobs_cap;
  l->arena_off = off; l->arena_cap = arena_cap;
  l->jobs_off = off + a; l->res_off = off + a + j; l->var_off = off + a + j + r; l->var_cap = var_cap;
}

void wake_lane(pce_lane *l, int rc) {
  l->rc = rc;
  __atomic_store_n(&l->state, (uint32_t)PCE_DONE, __ATOMIC_RELEASE);
  pce_futex(&l->state, FUTEX_WAKE, 64, nullptr);
}

// One submission loop.  Claims every POSTED lane of one session (a batch runs against one genome), runs them as one
// device batch, marks them DONE.  One loop per GPU by default (PC_ENGINE_THREADS overrides).
void engine_loop(pc_engine *e, Gpu *g, int which) {
  cudaSetDevice(g->device);
  pc_stream *st = pc_stream_create(g->base_ctx);
  if (!st) { fprintf(stderr, "* FATAL pc_engine: %s\n", pc_last_error()); return; }
  pce_hdr *h = g->hdr;
  std::vector<uint32_t> mine;
  std::vector<pc_part> parts;
  uint32_t rot = (uint32_t)which * 7u;
  bool timers_on = false;
  while (!e->stop.load(std::memory_order_acquire)) {
    const uint32_t bell = __atomic_load_n(&h->doorbell, __ATOMIC_SEQ_CST);
    mine.clear();
    uint32_t sid = 0;
    for (uint32_t q = 0; q < PCE_MAX_LANES; ++q) {
      const uint32_t i = (q + rot) % PCE_MAX_LANES;
      pce_lane *l = &h->lanes[i];
      if (__atomic_load_n(&l->state, __ATOMIC_ACQUIRE) != PCE_POSTED) continue;
      const uint32_t s = l->session;
      if (s == 0 || (sid && s != sid)) continue;
      uint32_t expect = PCE_POSTED;
      if (__atomic_compare_exchange_n(&l->state, &expect, (uint32_t)PCE_RUNNING, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) { sid = s; mine.push_back(i); }
    }
    if (mine.empty()) {
      __atomic_fetch_add(&h->sleepers, 1u, __ATOMIC_SEQ_CST);
      if (__atomic_load_n(&h->doorbell, __ATOMIC_SEQ_CST) == bell) {
        struct timespec to = {0, 20 * 1000 * 1000};
        pce_futex(&h->doorbell, FUTEX_WAIT, bell, &to);
      }
      __atomic_fetch_sub(&h->sleepers, 1u, __ATOMIC_SEQ_CST);
      continue;
    }
    rot = mine.back() + 1;
    const double t0 = now_s();
    Session *S = nullptr;
    {
      std::lock_guard<std::mutex> lk(g->mu);
      auto it = g->sessions.find(sid);
      if (it != g->sessions.end()) S = it->second;
    }
    int rc = 0;
    if (!S) rc = pc_set_error(PC_E_ARG, "engine: lane posted for an unknown session");
    parts.clear();
    uint64_t njobs = 0, h2d = 0, d2h = 0;
    if (!rc) {
      for (uint32_t i : mine) {
        pce_lane *l = &h->lanes[i];
        if (l->seg >= g->segs.size() || l->njobs > l->jobs_cap || l->arena_len > l->arena_cap || l->var_len > l->var_cap) {
          rc = pc_set_error(PC_E_ARG, "engine: lane posted with sizes beyond its slab");
          break;
        }
        uint8_t *base = g->segs[l->seg].base;
        pc_part p;
        p.arena = base + l->arena_off; p.arena_bytes = (size_t)l->arena_len;
        p.jobs = (const pc_job *)(base + l->jobs_off); p.njobs = (int)l->njobs;
        p.res = (int32_t *)(base + l->res_off);
        p.var_out = base + l->var_off; p.var_out_bytes = (size_t)l->var_len;
        parts.push_back(p);
        njobs += l->njobs;
        h2d += l->arena_len + sizeof(pc_job) * (uint64_t)l->njobs;
        d2h += l->var_len + sizeof(int32_t) * PC_RES_INTS * (uint64_t)l->njobs;
      }
    }
    const bool want_timers = e->timers.load() != 0;
    if (want_timers != timers_on) { pc_stream_enable_timers(st, want_timers); pc_stream_reset_timers(st); timers_on = want_timers; }
    const unsigned long long launches0 = tl_pc_launches;
    if (!rc) rc = pc_submit_parts(st, S->ctx, parts.data(), (int)parts.size());
    if (!rc) rc = pc_stream_sync(st);
    else { pc_stream_sync(st); cudaDeviceSynchronize(); }      // drain whatever was enqueued before the failure (side streams included)
    if (rc && S) fprintf(stderr, "* ERROR pc_engine: batch of %zu lane(s) failed: %s\n", mine.size(), pc_last_error());
    if (S) {
      std::lock_guard<std::mutex> lk(g->mu);
      pc_session_stats &T = S->stats;
      T.batches += 1; T.lanes_merged += mine.size(); T.jobs += njobs; T.h2d_bytes += h2d; T.d2h_bytes += d2h;
      T.launches += tl_pc_launches - launches0;
      T.busy_s += now_s() - t0;
      if (timers_on) {
        for (int op = 0; op < PC_OP_COUNT; ++op) { double ms = 0; uint64_t k = 0; pc_stream_op_time(st, op, &ms, &k); T.op_ms[op] += ms; }
        pc_stream_reset_timers(st);
      }
    }
    for (uint32_t i : mine) {
      pce_lane *l = &h->lanes[i];
      l->batches += 1; l->jobs_total += l->njobs;
      wake_lane(l, rc);
    }
  }
  pc_stream_destroy(st);
}

}  // namespace

extern "C" pc_engine *pc_engine_create(const int *devices, int ndev, size_t segment_bytes) {
  if (!devices || ndev < 1) { pc_set_error(PC_E_ARG, "pc_engine_create: no device given"); return nullptr; }
  pc_engine *e = new pc_engine();
  const char *env = getenv("PC_ENGINE_SEGMENT_MB");
  if (segment_bytes == 0) segment_bytes = (env && atol(env) > 0 ? (size_t)atol(env) : 384) << 20;
  // The submission loops wait with the driver's spinning synchronize.  Sleeping on a blocking event instead (PC_SYNC=block,
  // pc_set_blocking_sync) was measured on the 16-core box and under a 4-core budget (taskset, the share of one GPU of an 8-GPU
  // box): 1.1-1.4 ms per batch instead of 0.5-0.8, whole program 20-30 % slower in both settings (profiles/r2_summary.md).
  // Submission loops per GPU.  A batch is latency-bound (copies in, ~20 small kernels, copies out: 0.5-0.8 ms whatever its
  // size), so a second loop that forms the next batch while the first one waits for its stream cuts the queueing time of a
  // lane: measured on the 16-core box 111-114 k -> 119-120 k ESTs/s; three or four loops only split the batches further
  // (110-118 k).  Each loop spins on a core while it waits, so with fewer than 8 cores per GPU one loop is better (4-core
  // budget: 48 k vs 43-46 k ESTs/s).  PC_ENGINE_THREADS overrides.
  const long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
  int nthreads = ncpu > 0 && ncpu / ndev >= 8 ? 2 : 1;
  if (const char *t = getenv("PC_ENGINE_THREADS")) if (atoi(t) >= 1 && atoi(t) <= 8) nthreads = atoi(t);
  for (int i = 0; i < ndev; ++i) {
    Gpu *g = new Gpu();
    g->device = devices[i];
    g->default_seg = segment_bytes;
    g->segs.reserve(PCE_MAX_SEGMENTS);          // engine threads index segs without the lock: never reallocate
    g->base_ctx = pc_ctx_create(devices[i]);
    if (!g->base_ctx || add_segment(g, segment_bytes) < 0) { delete g; pc_engine_destroy(e); return nullptr; }
    e->gpus.push_back(g);
  }
  for (Gpu *g : e->gpus)
    for (int k = 0; k < nthreads; ++k) g->threads.emplace_back(engine_loop, e, g, k);
  return e;
}

extern "C" void pc_engine_destroy(pc_engine *e) {
  if (!e) return;
  e->stop.store(true, std::memory_order_release);
  for (Gpu *g : e->gpus) {
    if (g->hdr) { __atomic_fetch_add(&g->hdr->doorbell, 1u, __ATOMIC_SEQ_CST); pce_futex(&g->hdr->doorbell, FUTEX_WAKE, 64, nullptr); }
    for (auto &t : g->threads) t.join();
    for (auto &kv : g->sessions) { pc_ctx_destroy(kv.second->ctx); delete kv.second; }
    for (pc_ctx *c : g->idle_ctx) pc_ctx_destroy(c);
    cudaSetDevice(g->device);
    for (Segment &s : g->segs) {
      if (s.registered) cudaHostUnregister(s.base);
      munmap(s.base, s.bytes);
      close(s.fd);
    }
    pc_ctx_destroy(g->base_ctx);
    delete g;
  }
  delete e;
}

extern "C" const char *pc_engine_backend(void) { return "cuda-sm100a"; }
extern "C" int pc_engine_gpu_count(const pc_engine *e) { return e ? (int)e->gpus.size() : 0; }
extern "C" void pc_engine_enable_timers(pc_engine *e, int on) { if (e) e->timers.store(on != 0); }

extern "C" int pc_engine_open(pc_engine *e, const pc_session_req *req, pc_session_info *out) {
  if (!e || !req || !out || !req->genome || req->nlanes < 1 || req->nlanes > PCE_MAX_SESSION_LANES || req->jobs_cap < 1)
    return pc_set_error(PC_E_ARG, "pc_engine_open: bad argument");
  int gi = req->gpu;
  if (gi < 0) {                                // least loaded: fewest open sessions
    size_t best = (size_t)-1;
    for (size_t k = 0; k < e->gpus.size(); ++k) {
      std::lock_guard<std::mutex> lk(e->gpus[k]->mu);
      if (e->gpus[k]->sessions.size() < best) { best = e->gpus[k]->sessions.size(); gi = (int)k; }
    }
  }
  if (gi < 0 || gi >= (int)e->gpus.size()) return pc_set_error(PC_E_ARG, "pc_engine_open: no such GPU in this engine");
  Gpu *g = e->gpus[(size_t)gi];
  Session *S = new Session();
  S->id = e->next_session.fetch_add(1);
  S->gpu = gi;
  // genome + k-mer index of this session (replaces the per-process suffix tree, src/main-est-fact.c:223-240)
  {
    std::lock_guard<std::mutex> lk(g->mu);
    if (!g->idle_ctx.empty()) { S->ctx = g->idle_ctx.back(); g->idle_ctx.pop_back(); }
  }
  if (!S->ctx) S->ctx = pc_ctx_create(g->device);
  if (!S->ctx || pc_genome_upload(S->ctx, req->genome, req->genome_len, req->word_len, req->depth_rate)) {
    pc_ctx_destroy(S->ctx); delete S;
    return PC_E_CUDA;
  }
  std::lock_guard<std::mutex> lk(g->mu);
  memset(out, 0, sizeof *out);
  for (int k = 0; k < req->nlanes; ++k) {
    uint32_t li = PCE_MAX_LANES;
    for (uint32_t i = 0; i < PCE_MAX_LANES; ++i) if (g->hdr->lanes[i].session == 0 && g->hdr->lanes[i].state == PCE_FREE) { li = i; break; }
    size_t off = 0, bytes = 0;
    const int seg = li < PCE_MAX_LANES ? alloc_slab(g, req->arena_cap, req->jobs_cap, req->var_cap, &off, &bytes) : -1;
    if (seg < 0) {
      if (li >= PCE_MAX_LANES) pc_set_error(PC_E_NOMEM, "pc_engine_open: no free lane on this GPU");
      for (size_t q = 0; q < S->lanes.size(); ++q) {
        g->segs[S->slabs[q].first].release(S->slabs[q].second.first, S->slabs[q].second.second);
        memset(&g->hdr->lanes[S->lanes[q]], 0, sizeof(pce_lane));
      }
      pc_ctx_destroy(S->ctx); delete S;
      return PC_E_NOMEM;
    }
    pce_lane *l = &g->hdr->lanes[li];
    memset(l, 0, sizeof *l);
    lay_out(l, seg, off, req->arena_cap, req->jobs_cap, req->var_cap);
    l->session = S->id;
    __atomic_store_n(&l->state, (uint32_t)PCE_IDLE, __ATOMIC_RELEASE);
    S->lanes.push_back(li);
    S->slabs.push_back({(uint32_t)seg, {off, bytes}});
    out->lane[k] = li;
  }
  out->session = S->id; out->gpu = gi; out->nlanes = req->nlanes;
  g->sessions[S->id] = S;
  return 0;
}

static Session *find_session(pc_engine *e, uint32_t session, Gpu **g_out) {
  for (Gpu *g : e->gpus) {
    std::lock_guard<std::mutex> lk(g->mu);
    auto it = g->sessions.find(session);
    if (it != g->sessions.end()) { *g_out = g; return it->second; }
  }
  return nullptr;
}

extern "C" int pc_engine_resize_lane(pc_engine *e, uint32_t session, uint32_t lane, uint64_t arena_cap, uint32_t jobs_cap, uint64_t var_cap,
                                     uint64_t keep_arena, uint32_t keep_jobs) {
  Gpu *g = nullptr;
  Session *S = e ? find_session(e, session, &g) : nullptr;
  if (!S || lane >= PCE_MAX_LANES) return pc_set_error(PC_E_ARG, "pc_engine_resize_lane: unknown session or lane");
  std::lock_guard<std::mutex> lk(g->mu);
  size_t q = 0;
  while (q < S->lanes.size() && S->lanes[q] != lane) ++q;
  pce_lane *l = &g->hdr->lanes[lane];
  const uint32_t st = __atomic_load_n(&l->state, __ATOMIC_ACQUIRE);
  if (q == S->lanes.size() || (st != PCE_IDLE && st != PCE_DONE)) return pc_set_error(PC_E_ARG, "pc_engine_resize_lane: lane not owned or busy");
  if (keep_arena > l->arena_cap || keep_arena > arena_cap || keep_jobs > l->jobs_cap || keep_jobs > jobs_cap)
    return pc_set_error(PC_E_ARG, "pc_engine_resize_lane: more to keep than fits");
  size_t off = 0, bytes = 0;
  const int seg = alloc_slab(g, arena_cap, jobs_cap, var_cap, &off, &bytes);
  if (seg < 0) return seg;
  pce_lane old = *l;
  lay_out(l, seg, off, arena_cap, jobs_cap, var_cap);
  memcpy(g->segs[(size_t)seg].base + l->arena_off, g->segs[old.seg].base + old.arena_off, (size_t)keep_arena);
  memcpy(g->segs[(size_t)seg].base + l->jobs_off, g->segs[old.seg].base + old.jobs_off, sizeof(pc_job) * (size_t)keep_jobs);
  g->segs[S->slabs[q].first].release(S->slabs[q].second.first, S->slabs[q].second.second);
  S->slabs[q] = {(uint32_t)seg, {off, bytes}};
  return 0;
}

extern "C" int pc_engine_close(pc_engine *e, uint32_t session, pc_session_stats *stats) {
  Gpu *g = nullptr;
  Session *S = e ? find_session(e, session, &g) : nullptr;
  if (!S) return pc_set_error(PC_E_ARG, "pc_engine_close: unknown session");
  { std::lock_guard<std::mutex> lk(g->mu); if (S->closing) return 0; S->closing = true; }
  // lanes still posted are withdrawn, lanes in flight are waited for (the engine threads look the session up by id)
  for (uint32_t li : S->lanes) {
    pce_lane *l = &g->hdr->lanes[li];
    for (;;) {
      uint32_t st = __atomic_load_n(&l->state, __ATOMIC_ACQUIRE);
      if (st == PCE_POSTED) { if (__atomic_compare_exchange_n(&l->state, &st, (uint32_t)PCE_IDLE, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) break; continue; }
      if (st != PCE_RUNNING) break;
      usleep(200);
    }
  }
  {
    std::lock_guard<std::mutex> lk(g->mu);
    for (size_t q = 0; q < S->lanes.size(); ++q) {
      pce_lane *l = &g->hdr->lanes[S->lanes[q]];
      l->session = 0;
      __atomic_store_n(&l->state, (uint32_t)PCE_FREE, __ATOMIC_RELEASE);
      g->segs[S->slabs[q].first].release(S->slabs[q].second.first, S->slabs[q].second.second);
    }
    g->sessions.erase(S->id);
    if (stats) *stats = S->stats;
    if (getenv("PC_PROFILE") || getenv("PC_PROFILE_HOST")) pc_debug_dump();      /* cumulative phase clock of the submission loops */
    if (g->idle_ctx.size() < 8) { g->idle_ctx.push_back(S->ctx); S->ctx = nullptr; }
  }
  pc_ctx_destroy(S->ctx);
  delete S;
  return 0;
}

extern "C" int pc_engine_segment_count(pc_engine *e, int gpu) {
  if (!e || gpu < 0 || gpu >= (int)e->gpus.size()) return 0;
  std::lock_guard<std::mutex> lk(e->gpus[(size_t)gpu]->mu);
  return (int)e->gpus[(size_t)gpu]->segs.size();
}
extern "C" int pc_engine_segment_fd(pc_engine *e, int gpu, int seg, size_t *bytes) {
  if (!e || gpu < 0 || gpu >= (int)e->gpus.size()) return -1;
  Gpu *g = e->gpus[(size_t)gpu];
  std::lock_guard<std::mutex> lk(g->mu);
  if (seg < 0 || seg >= (int)g->segs.size()) return -1;
  if (bytes) *bytes = g->segs[(size_t)seg].bytes;
  return g->segs[(size_t)seg].fd;
}
extern "C" void *pc_engine_segment_base(pc_engine *e, int gpu, int seg) {
  if (!e || gpu < 0 || gpu >= (int)e->gpus.size()) return nullptr;
  Gpu *g = e->gpus[(size_t)gpu];
  std::lock_guard<std::mutex> lk(g->mu);
  if (seg < 0 || seg >= (int)g->segs.size()) return nullptr;
  return g->segs[(size_t)seg].base;
}
