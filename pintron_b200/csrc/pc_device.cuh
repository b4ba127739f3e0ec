// pc_device.cuh — shared device-side declarations for libpintron_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pintron_cuda.h"

#define PC_WARPS_PER_CTA 4
#define PC_SMEM_INTS_PER_WARP 3072          /* 12 KB per warp of anti-diagonal state */
#define PC_INF 0x3fffffffu

struct PcDevBatch {
  const uint8_t *arena;
  const uint8_t *genome;
  uint32_t genome_len;
  const pc_job *jobs;
  const uint32_t *idx;        /* job indices of this op, heaviest first */
  int n;
  const uint32_t *n_dev;      /* when set: the job count is read from device memory (list built by an earlier kernel) */
  int32_t *res;
  uint8_t *var_out;
  uint8_t *pool;              /* per-stream scratch pool: one equal slot per warp of the launch, reused job after job */
  unsigned long long pool_cap;
  unsigned long long *pool_need;   /* atomicMax of the bytes a job wanted when its slot was too small */
  int slots;                  /* warps in this launch (set by the launcher) */
  int max_warps;              /* host hint: cap on the warps of a launch (retry rounds use fewer, larger slots) */
  /* k-mer index (SEED) */
  const unsigned long long *ix_keys;
  const uint32_t *ix_pos;
  const uint32_t *ix_bstart;  /* bucket starts: bucket = hash >> ix_shift */
  int ix_shift;
  uint32_t ix_n;
  int ix_word;
  double depth_rate;
  /* nine bit planes of the genome (LCS scan), or nullptr */
  const uint32_t *gplanes;
  uint32_t gplane_words;
};

/* device-side job ordering (k_order.cu) */
#define PC_ORDER_SEGS (PC_OP_COUNT * 4)
#define PC_ORDER_BINS (PC_ORDER_SEGS * 64)
#define PC_LCS_TPB 1024        /* diagonals per tile of the LCS scan (tile counts are computed with this on host and device) */
#define PC_LCS_MAX_S2 4096
struct PcSegStat { uint32_t n, max_a, max_b, max_t; unsigned long long lcs_blocks; };      /* max_t: widest BORDERS window */
__host__ __device__ inline uint32_t pc_borders_window(const pc_job &j) {
  const unsigned long long tw = (unsigned long long)j.a_len + (uint32_t)j.p0;
  return (uint32_t)(tw < j.b_len ? tw : j.b_len);
}

#define PC_BORDERS_FAST_MAX_T 1024
#define PC_GAP_FAST_MAX_M 3000      /* packed GAP kernel: 64 * (m + 1) B of column codes per CTA must fit the 200 KB opt-in */
/* Kernel class of a job inside its op (shared by the host and the device ordering): GAP and BORDERS jobs that fit the
 * packed register kernels are classed by their row count (0 / 1 / 2 = 8 / 16 / 32 lanes per job), EDIT / KBAND jobs by
 * the words per column of the bit-parallel kernel; 3 = generic wavefront kernel. */
__host__ __device__ inline int pc_job_class(const pc_job &j) {
  if (j.op == PC_OP_GAP) {
    if (j.a_len < 1 || j.b_len < 1 || j.b_len > PC_GAP_FAST_MAX_M || j.a_len > 256) return 3;
    return j.a_len <= 64 ? 0 : (j.a_len <= 128 ? 1 : 2);
  }
  if (j.op == PC_OP_BORDERS) {
    const unsigned long long tw = (unsigned long long)j.a_len + (uint32_t)j.p0;
    const unsigned long long t_win = tw < j.b_len ? tw : j.b_len;
    if (j.a_len < 1 || j.a_len > 256 || t_win < 1 || t_win > PC_BORDERS_FAST_MAX_T) return 3;
    return j.a_len <= 64 ? 0 : (j.a_len <= 128 ? 1 : 2);
  }
  if (j.op == PC_OP_EDIT || j.op == PC_OP_KBAND) {        /* bit-parallel kernel: words per column from the shorter string */
    const uint32_t m = j.a_len < j.b_len ? j.a_len : j.b_len;
    return m <= 64 ? 0 : (m <= 128 ? 1 : (m <= 320 ? 2 : 3));
  }
  if (j.op == PC_OP_SEED) return j.a_len <= 1024 ? 0 : (j.a_len <= 3072 ? 1 : 2);      /* scratch per warp grows with the read: long reads get larger slots */
  return 0;
}
/* cost class inside a segment, 0..63, larger = heavier (packed kernels: by the number of column steps, finely) */
__host__ __device__ inline int pc_job_cost(const pc_job &j, int cls) {
  if ((j.op == PC_OP_GAP || j.op == PC_OP_BORDERS) && cls < 3) {
    unsigned long long m = j.b_len;
    if (j.op == PC_OP_BORDERS) { const unsigned long long tw = (unsigned long long)j.a_len + (uint32_t)j.p0; if (tw < m) m = tw; }
    return m < 512 ? (int)(m >> 4) : 32 + (int)((m - 512) >> 7 < 31 ? (m - 512) >> 7 : 31);
  }
  unsigned long long c;
  if (j.op == PC_OP_LCS || j.op == PC_OP_SEED) c = (unsigned long long)j.a_len + j.b_len + 1ull;
  else c = ((unsigned long long)j.a_len + 1ull) * ((unsigned long long)j.b_len + 1ull) + 1ull;
#ifdef __CUDA_ARCH__
  return 63 - __clzll((long long)c);
#else
  return 63 - __builtin_clzll(c);
#endif
}

struct WarpPool { uint8_t *base; unsigned long long size, used; };

__device__ __forceinline__ WarpPool pc_warp_pool(const PcDevBatch &B, int slot) {
  WarpPool wp;
  wp.size = (B.pool_cap / (unsigned long long)B.slots) & ~255ull;
  wp.base = B.pool + wp.size * (unsigned long long)slot;
  wp.used = 0;
  return wp;
}

// All lanes of the warp call this with the same arguments and get the same pointer (nullptr = slot too small).
__device__ __forceinline__ uint8_t *pc_pool_alloc(const PcDevBatch &B, WarpPool &wp, unsigned long long bytes, int lane) {
  const unsigned long long sz = (bytes + 255ull) & ~255ull;
  if (wp.used + sz > wp.size) {
    if (lane == 0) atomicMax(B.pool_need, wp.used + sz);
    return nullptr;
  }
  uint8_t *p = wp.base + wp.used;
  wp.used += sz;
  return p;
}

// Per-halfword unsigned min / max that also say which operand won (pred = "a is the result": a <= b for min, a >= b
// for max); ptxas turns each into one VIMNMX.U16x2 with two predicate outputs.  Same PTX as CUDA's __vibmin_u16x2 /
// __vibmax_u16x2, but with early-clobber outputs: the header's version lets the result share a register with `a`
// (seen with a loop-carried running minimum), after which its own compare reads the overwritten value and the
// predicates come out constant.
__device__ __forceinline__ uint32_t pc_vibmin_u16x2(uint32_t a, uint32_t b, bool &pred_hi, bool &pred_lo) {
  uint32_t val, h, l;
  asm("{.reg .pred pu, pv; \n\t"
      ".reg .u16 rs0, rs1, rs2, rs3; \n\t"
      "min.u16x2 %0, %3, %4; \n\t"
      "mov.b32 {rs0, rs1}, %0; \n\t"
      "mov.b32 {rs2, rs3}, %3; \n\t"
      "setp.eq.u16 pv, rs0, rs2; \n\t"
      "setp.eq.u16 pu, rs1, rs3; \n\t"
      "selp.b32 %1, 1, 0, pu; \n\t"
      "selp.b32 %2, 1, 0, pv;} \n\t"
      : "=&r"(val), "=&r"(h), "=&r"(l) : "r"(a), "r"(b));
  pred_hi = h != 0; pred_lo = l != 0;
  return val;
}
__device__ __forceinline__ uint32_t pc_vibmax_u16x2(uint32_t a, uint32_t b, bool &pred_hi, bool &pred_lo) {
  uint32_t val, h, l;
  asm("{.reg .pred pu, pv; \n\t"
      ".reg .u16 rs0, rs1, rs2, rs3; \n\t"
      "max.u16x2 %0, %3, %4; \n\t"
      "mov.b32 {rs0, rs1}, %0; \n\t"
      "mov.b32 {rs2, rs3}, %3; \n\t"
      "setp.eq.u16 pv, rs0, rs2; \n\t"
      "setp.eq.u16 pu, rs1, rs3; \n\t"
      "selp.b32 %1, 1, 0, pu; \n\t"
      "selp.b32 %2, 1, 0, pv;} \n\t"
      : "=&r"(val), "=&r"(h), "=&r"(l) : "r"(a), "r"(b));
  pred_hi = h != 0; pred_lo = l != 0;
  return val;
}

__device__ __forceinline__ bool pc_is_n(uint8_t c) { return c == 'n' || c == 'N'; }

void pc_launch_dp(int op, const PcDevBatch &B, cudaStream_t s, int sm_count);
void pc_launch_myers(int op, int cls, const PcDevBatch &B, uint32_t *slow_list, uint32_t *slow_count, cudaStream_t s, int sm_count);
void pc_launch_align_bp(const PcDevBatch &B, uint32_t *slow_list, uint32_t *slow_count, cudaStream_t s, int sm_count);
void pc_order_jobs(const pc_job *d_jobs, int n, size_t arena_bytes, size_t genome_len, size_t var_bytes, int lcs_tpb, int lcs_max_s2,
                   uint16_t *d_keys, uint32_t *d_work, PcSegStat *d_seg, uint32_t *d_order, cudaStream_t s, int sm_count);
void pc_collect_status(const int32_t *d_res, int n, int code, uint32_t *d_list, uint32_t *d_count, cudaStream_t s, int sm_count);
void pc_lcs_prefix(const pc_job *d_jobs, const uint32_t *d_order, int n, int lcs_tpb, int lcs_max_s2, uint32_t *d_prefix, cudaStream_t s);
void pc_launch_borders_chunked(const PcDevBatch &B, uint32_t *slow_list, uint32_t *slow_count, cudaStream_t s, int sm_count);
void pc_launch_borders_packed(int cls, const PcDevBatch &B, int tcap, cudaStream_t s, int sm_count);
void pc_launch_gap_pairs(int cls, const PcDevBatch &B, int max_m, uint32_t *work, cudaStream_t s, int sm_count);
void pc_launch_seed(const PcDevBatch &B, int max_len, cudaStream_t s, int sm_count);
int pc_lcs_blocks(long long l1, int l2);
int pc_launch_lcs(const PcDevBatch &B, unsigned long long *best, const uint32_t *d_blk_prefix, uint32_t total_blocks, int max_l2, uint32_t *work,
                  cudaStream_t s);
/* device buffers of one genome index; they only grow, so a context that is reused for the next genome (the engine keeps
 * idle contexts) builds its index without a single cudaMalloc / cudaFree */
struct PcGrowBuf {
  void *p = nullptr; size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return PC_E_NOMEM; }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PcIndexBufs { PcGrowBuf keys_in, keys_out, pos_in, pos_out, tmp, bstart, planes; };
int pc_build_planes(const uint8_t *d_genome, uint32_t len, PcGrowBuf &planes, uint32_t *nwords_out, cudaStream_t s);
int pc_build_index(const uint8_t *d_genome, uint32_t len, int word, PcIndexBufs &bufs, unsigned long long **keys, uint32_t **pos,
                   uint32_t *n_out, uint32_t **bstart, int *shift, cudaStream_t s);
extern unsigned long long g_pc_launches;
extern thread_local unsigned long long tl_pc_launches;      /* launches issued by the calling thread (per-stream accounting) */
#define PC_COUNT_LAUNCH(n) do { __atomic_fetch_add(&g_pc_launches, (unsigned long long)(n), __ATOMIC_RELAXED); tl_pc_launches += (n); } while (0)
/* cudaOccupancyMaxActiveBlocksPerMultiprocessor, asked once per (device, kernel, shared-memory KB) instead of per launch */
int pc_cached_occupancy(const void *func, int tpb, size_t smem);
/* cudaFuncSetAttribute(MaxDynamicSharedMemorySize), once per (device, kernel): the attribute is per device */
int pc_smem_optin(const void *func, int bytes);
void pc_rebase_jobs(pc_job *d_jobs, int n, const uint32_t *d_parts, int nparts, cudaStream_t s, int sm_count);
/* warps of a launch that own a scratch slot: retry rounds may allow fewer warps than one CTA holds */
__device__ __forceinline__ int pc_active_warps(const PcDevBatch &B, int launched) { return B.slots < launched ? B.slots : launched; }
