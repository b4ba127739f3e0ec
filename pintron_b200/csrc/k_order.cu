// k_order.cu — device-side job ordering for large batches (sm_100a).
//
// pc_submit lays the job indices of a batch out segment by segment ((op, class) pairs), heaviest first inside a
// segment, for the persistent kernels.  For the batches the est-fact host sends (a few thousand jobs) that is a
// 10-microsecond counting sort on the submitting thread; for a device-resident batch of millions of jobs the host
// pass would cost more than the kernels it feeds, so the same counting sort runs here:
//   k_job_keys    one thread per job: validates the job, computes its 16-bit key (segment * 64 + cost class, the
//                 formulas of pc_api.cu), a histogram of the keys (shared-memory bins per block) and per-segment
//                 statistics (job count, longest a / b string, LCS blocks) the host needs to shape the launches
//   k_scan_bins   exclusive scan of the 2560 bins (one block)
//   k_scatter     order[cursor[key]++] = job  (warp-aggregated atomics; the order inside a bin is irrelevant)
//   k_lcs_prefix  exclusive scan of the per-job block counts of the LCS segment (k_lcs finds its job by binary search)
#include "pc_device.cuh"

namespace {

__device__ __forceinline__ bool job_valid(const pc_job &j, size_t arena_bytes, size_t genome_len, size_t var_bytes) {
  if (j.op >= PC_OP_COUNT) return false;
  if ((size_t)j.a_off + j.a_len > arena_bytes) return false;
  if (j.op != PC_OP_SEED) {
    const size_t lim = (j.flags & PC_B_IN_GENOME) ? genome_len : arena_bytes;
    if ((size_t)j.b_off + j.b_len > lim) return false;
  } else if (j.p1 == PC_SEED_BUILD_MEG) {            /* b = the MEG options, in the arena */
    if ((j.flags & PC_B_IN_GENOME) || j.b_len != sizeof(pc_meg_cfg) || (size_t)j.b_off + j.b_len > arena_bytes || j.p0 < 1) return false;
  }
  if (j.op == PC_OP_ALIGN || j.op == PC_OP_GAP) { if ((size_t)j.out_off + j.out_cap > var_bytes) return false; }
  else if (j.op == PC_OP_SEED) { if ((j.out_off & 3u) || (size_t)j.out_off + 12ull * j.out_cap > var_bytes) return false; }
  return true;
}

__global__ void __launch_bounds__(256) k_job_keys(const pc_job *jobs, int n, size_t arena_bytes, size_t genome_len, size_t var_bytes,
                                                  int lcs_tpb, int lcs_max_s2, uint16_t *keys, uint32_t *bins, PcSegStat *seg,
                                                  uint32_t *invalid) {
  __shared__ uint32_t sh_bins[PC_ORDER_BINS];
  __shared__ PcSegStat sh_seg[PC_ORDER_SEGS];
  for (int x = threadIdx.x; x < PC_ORDER_BINS; x += blockDim.x) sh_bins[x] = 0;
  for (int x = threadIdx.x; x < PC_ORDER_SEGS; x += blockDim.x) { sh_seg[x].n = 0; sh_seg[x].max_a = 0; sh_seg[x].max_b = 0; sh_seg[x].max_t = 0; sh_seg[x].lcs_blocks = 0; }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const pc_job j = jobs[i];
    if (!job_valid(j, arena_bytes, genome_len, var_bytes)) { keys[i] = 0xffffu; atomicAdd(invalid, 1u); continue; }
    const int cls = pc_job_class(j);
    const int lg = pc_job_cost(j, cls);
    const int sg = (int)j.op * 4 + cls;
    const uint16_t key = (uint16_t)(sg * 64 + (63 - lg));
    keys[i] = key;
    atomicAdd(&sh_bins[key], 1u);
    atomicAdd(&sh_seg[sg].n, 1u);
    atomicMax(&sh_seg[sg].max_a, j.a_len);
    atomicMax(&sh_seg[sg].max_b, j.b_len);
    if (j.op == PC_OP_BORDERS) atomicMax(&sh_seg[sg].max_t, pc_borders_window(j));
    if (j.op == PC_OP_LCS) {
      const long long l1 = j.b_len; const int l2 = (int)j.a_len;
      const unsigned long long blocks = (l2 > lcs_max_s2 || l2 <= 0 || l1 <= 0) ? 0ull : (unsigned long long)((l1 + l2 - 1 + lcs_tpb - 1) / lcs_tpb);
      atomicAdd(&sh_seg[sg].lcs_blocks, blocks);
    }
  }
  __syncthreads();
  for (int x = threadIdx.x; x < PC_ORDER_BINS; x += blockDim.x) if (sh_bins[x]) atomicAdd(&bins[x], sh_bins[x]);
  for (int x = threadIdx.x; x < PC_ORDER_SEGS; x += blockDim.x)
    if (sh_seg[x].n) {
      atomicAdd(&seg[x].n, sh_seg[x].n); atomicMax(&seg[x].max_a, sh_seg[x].max_a); atomicMax(&seg[x].max_b, sh_seg[x].max_b); atomicMax(&seg[x].max_t, sh_seg[x].max_t);
      if (sh_seg[x].lcs_blocks) atomicAdd(&seg[x].lcs_blocks, sh_seg[x].lcs_blocks);
    }
}

// bins[0 .. NB) counts -> start[] (exclusive prefix, NB + 1 entries) and cursor[] (a working copy for the scatter)
__global__ void __launch_bounds__(1024) k_scan_bins(const uint32_t *bins, uint32_t *start, uint32_t *cursor) {
  __shared__ uint32_t part[1024];
  constexpr int PER = (PC_ORDER_BINS + 1023) / 1024;
  uint32_t local[PER], sum = 0;
  for (int q = 0; q < PER; ++q) { const int x = threadIdx.x * PER + q; local[q] = x < PC_ORDER_BINS ? bins[x] : 0u; sum += local[q]; }
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - sum;
  for (int q = 0; q < PER; ++q) {
    const int x = threadIdx.x * PER + q;
    if (x < PC_ORDER_BINS) { start[x] = run; cursor[x] = run; }
    run += local[q];
  }
  if (threadIdx.x == 1023) start[PC_ORDER_BINS] = part[1023];
}

__global__ void __launch_bounds__(256) k_scatter(const uint16_t *keys, int n, uint32_t *cursor, uint32_t *order) {
  const int lane = threadIdx.x & 31;
  for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < n; i0 += gridDim.x * blockDim.x) {
    const int i = i0 + lane;
    const uint32_t key = i < n ? keys[i] : 0xffffu;
    const bool live = key != 0xffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (live && lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live) order[base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)i;
  }
}

// prefix[q] = blocks of the LCS jobs order[0 .. q) (one block of 1024 threads walks the segment in tiles)
__global__ void __launch_bounds__(1024) k_lcs_prefix(const pc_job *jobs, const uint32_t *order, int n, int lcs_tpb, int lcs_max_s2, uint32_t *prefix) {
  __shared__ uint32_t part[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int q = t0 + threadIdx.x;
    uint32_t v = 0;
    if (q < n) {
      const pc_job &j = jobs[order[q]];
      const long long l1 = j.b_len; const int l2 = (int)j.a_len;
      v = (l2 > lcs_max_s2 || l2 <= 0 || l1 <= 0) ? 0u : (uint32_t)((l1 + l2 - 1 + lcs_tpb - 1) / lcs_tpb);
    }
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const uint32_t add = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
      __syncthreads();
      part[threadIdx.x] += add;
      __syncthreads();
    }
    if (q < n) prefix[q] = carry + part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) prefix[n] = carry;
}

// Multi-part batches (pc_submit_parts): the jobs of part q were written with offsets relative to that part's own arena /
// var_out; parts[3q .. 3q+2] = (first job, arena base, var base) of part q inside the merged device buffers, parts[3*nparts]
// = total jobs.  One thread per job finds its part by binary search and rebases the three offsets.
__global__ void __launch_bounds__(256) k_rebase(pc_job *jobs, int n, const uint32_t *parts, int nparts) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int lo = 0, hi = nparts - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (parts[3 * mid] <= (uint32_t)i) lo = mid; else hi = mid - 1; }
    const uint32_t ab = parts[3 * lo + 1], vb = parts[3 * lo + 2];
    pc_job j = jobs[i];
    j.a_off += ab;
    if (!(j.flags & PC_B_IN_GENOME)) j.b_off += ab;
    j.out_off += vb;
    jobs[i] = j;
  }
}

// jobs whose status is `code` (PC_E_POOL after a pass): their indices, for the re-run with larger scratch slots
__global__ void __launch_bounds__(256) k_collect_status(const int32_t *res, int n, int code, uint32_t *list, uint32_t *count) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (res[(size_t)i * PC_RES_INTS] == code) list[atomicAdd(count, 1u)] = (uint32_t)i;
}

}  // namespace

void pc_collect_status(const int32_t *d_res, int n, int code, uint32_t *d_list, uint32_t *d_count, cudaStream_t s, int sm_count) {
  cudaMemsetAsync(d_count, 0, sizeof(uint32_t), s);
  k_collect_status<<<min((n + 255) / 256, sm_count * 8), 256, 0, s>>>(d_res, n, code, d_list, d_count);
  PC_COUNT_LAUNCH(1);
}

// Enqueue keys + histogram + scan + scatter on `s`.  work = [bins NB | start NB+1 | cursor NB | invalid 1] uint32 (zeroed here),
// seg = PC_ORDER_SEGS PcSegStat (zeroed here).  The caller copies seg / invalid back and synchronises before it
// shapes the per-segment launches; start[sg * 64] is the offset of segment sg inside order.
void pc_order_jobs(const pc_job *d_jobs, int n, size_t arena_bytes, size_t genome_len, size_t var_bytes, int lcs_tpb, int lcs_max_s2,
                   uint16_t *d_keys, uint32_t *d_work, PcSegStat *d_seg, uint32_t *d_order, cudaStream_t s, int sm_count) {
  uint32_t *bins = d_work, *start = d_work + PC_ORDER_BINS, *cursor = start + PC_ORDER_BINS + 1, *invalid = cursor + PC_ORDER_BINS;
  cudaMemsetAsync(d_work, 0, sizeof(uint32_t) * (3 * PC_ORDER_BINS + 2), s);
  cudaMemsetAsync(d_seg, 0, sizeof(PcSegStat) * PC_ORDER_SEGS, s);
  const int grid = min((n + 255) / 256, sm_count * 4);
  k_job_keys<<<grid, 256, 0, s>>>(d_jobs, n, arena_bytes, genome_len, var_bytes, lcs_tpb, lcs_max_s2, d_keys, bins, d_seg, invalid);
  k_scan_bins<<<1, 1024, 0, s>>>(bins, start, cursor);
  k_scatter<<<min((n + 255) / 256, sm_count * 8), 256, 0, s>>>(d_keys, n, cursor, d_order);
  PC_COUNT_LAUNCH(3);
}

void pc_rebase_jobs(pc_job *d_jobs, int n, const uint32_t *d_parts, int nparts, cudaStream_t s, int sm_count) {
  if (n <= 0 || nparts <= 0) return;
  k_rebase<<<min((n + 255) / 256, sm_count * 8), 256, 0, s>>>(d_jobs, n, d_parts, nparts);
  PC_COUNT_LAUNCH(1);
}

void pc_lcs_prefix(const pc_job *d_jobs, const uint32_t *d_order, int n, int lcs_tpb, int lcs_max_s2, uint32_t *d_prefix, cudaStream_t s) {
  k_lcs_prefix<<<1, 1024, 0, s>>>(d_jobs, d_order, n, lcs_tpb, lcs_max_s2, d_prefix);
  PC_COUNT_LAUNCH(1);
}
