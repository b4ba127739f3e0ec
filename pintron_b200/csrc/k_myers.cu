// k_myers.cu — edit_distance (reference src/refine.c:51, src/compute-alignments.c:235) and K_band_edit_distance
// (src/compute-alignments.c:319-453) as ONE JOB PER THREAD, bit-parallel (myers_core.h), for sm_100a.
//
// est-fact issues these two by the million on short strings (splice-shift checks of 10-40 nt, one K-band test per exon
// and candidate factorization): a warp-wide wavefront spends its time in __syncwarp and shared-memory round trips on
// matrices of a few hundred cells.  Here 64 rows of a column are one 64-bit word, a column costs ~17 integer
// instructions per block of 64 rows, and a warp works on 32 jobs at a time.  The match vectors (10 symbols x MAXW
// words per thread) live in shared memory, interleaved by thread (conflict-free).
//
// Jobs this form cannot answer go to a device-side list that the generic wavefront kernel (k_dp.cu) then runs, so
// every result stays bit-exact: strings with bytes outside ACGT/acgt/N/n, shorter strings above 64*MAXW letters, and
// K-band jobs whose true distance is above k while the reference would report the value of its band-restricted
// matrix (2k+1 < n): that value is not the edit distance and only the banded sweep reproduces it.
#include "pc_device.cuh"

// symbol table in shared memory (filled per block) and word-wise string loads for myers_core.h
#define MY_HD __device__ __forceinline__
#define MY_SYM(c) ((int)my_symtab[(c)])
#define MY_LOAD4(p) my_load4_dev(p)
__device__ __forceinline__ uint32_t my_load4_dev(const uint8_t *p) {
  const uintptr_t a = (uintptr_t)p;
  const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);       // both buffers are readable 16 bytes past their end
  return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3) * 8u);
}
__shared__ int8_t my_symtab[256];
#include "myers_core.h"
#include "align_core.h"

namespace {

constexpr int MY_TPB = 64;

// The wide variant serves 3, 4 and 5 words per column: the instance with exactly the job's word count runs, so a
// 150-letter string does not pay for two predicated-off blocks per column (the Peq array is sized for 5).
template <int MAXW>
__device__ __forceinline__ uint32_t myers_dispatch(const uint8_t *pat, int m, const uint8_t *txt, int n, unsigned long long *peq, int stride) {
  if (MAXW <= 2) return my_edit_distance<MAXW>(pat, m, txt, n, peq, stride);
  const int W = (m + 63) >> 6;
  if (W <= 3) return my_edit_distance<3>(pat, m, txt, n, peq, stride);
  if (W == 4) return my_edit_distance<4>(pat, m, txt, n, peq, stride);
  return my_edit_distance<5>(pat, m, txt, n, peq, stride);
}

template <int OP, int MAXW>
__global__ void __launch_bounds__(MY_TPB) k_myers(PcDevBatch B, uint32_t *slow_list, uint32_t *slow_count) {
  __shared__ unsigned long long peq[MY_NSYM * MAXW * MY_TPB];
  for (int c = threadIdx.x; c < 256; c += MY_TPB) my_symtab[c] = (int8_t)my_sym_switch((uint8_t)c);
  __syncthreads();
  for (int w = blockIdx.x * MY_TPB + threadIdx.x; w < B.n; w += gridDim.x * MY_TPB) {
    const uint32_t ji = B.idx[w];
    const pc_job *job = B.jobs + ji;
    int32_t *res = B.res + (size_t)ji * PC_RES_INTS;
    const uint8_t *a = B.arena + job->a_off;
    const uint8_t *b = ((job->flags & PC_B_IN_GENOME) ? B.genome : B.arena) + job->b_off;
    const int la = (int)job->a_len, lb = (int)job->b_len;
    const uint8_t *pat = a, *txt = b;
    int m = la, n = lb;
    if (m > n) { pat = b; txt = a; m = lb; n = la; }                 // the distance is symmetric; the shorter string goes along the bits
    bool slow = m > 64 * MAXW;
    if (OP == PC_OP_EDIT) {
      if (!slow) {
        const uint32_t d = myers_dispatch<MAXW>(pat, m, txt, n, peq + threadIdx.x, MY_TPB);
        if (d == MY_UNSUPPORTED) slow = true;
        else { res[0] = PC_OK; res[1] = (int32_t)d; }
      }
    } else {
      // K_band_edit_distance(seq1, seq2, k, &edit): the order of the reference's tests
      const uint32_t k = (uint32_t)job->p0;
      if (la != lb && k == 0) { res[0] = PC_OK; res[1] = 0; res[2] = 1; }                                     // not equal, no error allowed
      else if (la != lb && (uint32_t)(n - m) > k) { res[0] = PC_OK; res[1] = 0; res[2] = n - m; }               // length gap alone is too much
      else if (!slow) {
        const uint32_t d = myers_dispatch<MAXW>(pat, m, txt, n, peq + threadIdx.x, MY_TPB);
        if (d == MY_UNSUPPORTED) slow = true;
        else if (la == lb && d == 0) { res[0] = PC_OK; res[1] = 1; res[2] = 0; }                               // equal strings
        else if (k == 0) { res[0] = PC_OK; res[1] = 0; res[2] = 1; }
        else if (2ull * k + 1ull >= (unsigned long long)n || d <= k) { res[0] = PC_OK; res[1] = d <= k; res[2] = (int32_t)d; }   // full matrix, or inside the band: exact
        else if (job->flags & PC_KBAND_OK_ONLY) { res[0] = PC_OK; res[1] = 0; res[2] = (int32_t)d; }             // not ok; the caller does not read `edit`
        else slow = true;                                                                                      // band-restricted value wanted
      }
    }
    if (slow) slow_list[atomicAdd(slow_count, 1u)] = ji;
  }
}

template <int OP, int MAXW>
void launch_myers(const PcDevBatch &B, uint32_t *slow_list, uint32_t *slow_count, cudaStream_t s, int sm_count) {
  const int per_sm = pc_cached_occupancy((const void *)k_myers<OP, MAXW>, MY_TPB, 0);
  const int needed = (B.n + MY_TPB - 1) / MY_TPB;
  const int grid = needed < sm_count * per_sm ? needed : sm_count * per_sm;
  k_myers<OP, MAXW><<<grid < 1 ? 1 : grid, MY_TPB, 0, s>>>(B, slow_list, slow_count);
  PC_COUNT_LAUNCH(1);
}

// ---- compute_alignment, one job per thread (align_core.h) -----------------------------------------------------------
// The traceback words (two per genome column and 64-row block) go to the stream's scratch pool, interleaved by thread:
// word r of thread t lives at pool[t + r * T] (T = threads of the grid), so a warp's 32 jobs store 256 contiguous bytes
// per word.  A job whose columns do not fit its share of the pool, whose EST is longer than 320 nt or that holds a byte
// outside ACGTacgtNn is listed for the wavefront kernel (op_align, k_dp.cu), like the jobs k_myers cannot answer.
template <int MAXW>
__device__ __forceinline__ uint32_t align_dispatch(const uint8_t *pat, int n, const uint8_t *txt, int m, unsigned long long *peq, int stride,
                                                   unsigned long long *tb, long long tbs, long long cap, uint8_t *ops, int *k) {
  const int W = (n + 63) >> 6;
  if (W <= 1) return my_align<1>(pat, n, txt, m, peq, stride, tb, tbs, cap, ops, k);
  if (W == 2) return my_align<2>(pat, n, txt, m, peq, stride, tb, tbs, cap, ops, k);
  if (W == 3) return my_align<3>(pat, n, txt, m, peq, stride, tb, tbs, cap, ops, k);
  if (W == 4) return my_align<4>(pat, n, txt, m, peq, stride, tb, tbs, cap, ops, k);
  return my_align<MY_MAXW>(pat, n, txt, m, peq, stride, tb, tbs, cap, ops, k);
}

__global__ void __launch_bounds__(MY_TPB) k_align_bp(PcDevBatch B, uint32_t *slow_list, uint32_t *slow_count) {
  __shared__ unsigned long long peq[MY_NSYM * MY_MAXW * MY_TPB];
  for (int c = threadIdx.x; c < 256; c += MY_TPB) my_symtab[c] = (int8_t)my_sym_switch((uint8_t)c);
  __syncthreads();
  const long long T = (long long)gridDim.x * MY_TPB;
  const long long tid = (long long)blockIdx.x * MY_TPB + threadIdx.x;
  const long long cap = (long long)(B.pool_cap / (16ull * (unsigned long long)T));      // (column, block) entries per thread
  unsigned long long *tb = (unsigned long long *)B.pool + tid;
  for (long long w = tid; w < B.n; w += T) {
    const uint32_t ji = B.idx[w];
    const pc_job *job = B.jobs + ji;
    int32_t *res = B.res + (size_t)ji * PC_RES_INTS;
    const uint8_t *a = B.arena + job->a_off;
    const uint8_t *b = ((job->flags & PC_B_IN_GENOME) ? B.genome : B.arena) + job->b_off;
    const int n = (int)job->a_len, m = (int)job->b_len;
    if ((uint32_t)(n + m) > job->out_cap) { res[0] = PC_E_OUTCAP; continue; }
    bool slow = n > 64 * MY_MAXW;
    if (!slow) {
      int k = 0;
      const uint32_t d = align_dispatch<MY_MAXW>(a, n, b, m, peq + threadIdx.x, MY_TPB, tb, T, cap, B.var_out + job->out_off, &k);
      if (d == MY_UNSUPPORTED || d == MY_NOSPACE) slow = true;
      else { res[0] = PC_OK; res[1] = (int32_t)d; res[2] = k; }
    }
    if (slow) slow_list[atomicAdd(slow_count, 1u)] = ji;
  }
}

}  // namespace

// compute_alignment jobs of one segment: bit-parallel kernel first, the rest through slow_list to the wavefront kernel
void pc_launch_align_bp(const PcDevBatch &B, uint32_t *slow_list, uint32_t *slow_count, cudaStream_t s, int sm_count) {
  const int needed = (B.n + MY_TPB - 1) / MY_TPB;
  const int grid = needed < 2 * sm_count ? needed : 2 * sm_count;
  k_align_bp<<<grid < 1 ? 1 : grid, MY_TPB, 0, s>>>(B, slow_list, slow_count);
  PC_COUNT_LAUNCH(1);
}

// op = PC_OP_EDIT or PC_OP_KBAND.  cls = length class of the segment (pc_job_class: shorter string <= 64 / 128 / 320
// letters -> 1 / 2 / 5 words per column).  slow_list (B.n entries) / slow_count (zeroed by the caller) receive the job indices
// left for the generic kernel.
void pc_launch_myers(int op, int cls, const PcDevBatch &B, uint32_t *slow_list, uint32_t *slow_count, cudaStream_t s, int sm_count) {
  if (op == PC_OP_EDIT) {
    if (cls == 0) launch_myers<PC_OP_EDIT, 1>(B, slow_list, slow_count, s, sm_count);
    else if (cls == 1) launch_myers<PC_OP_EDIT, 2>(B, slow_list, slow_count, s, sm_count);
    else launch_myers<PC_OP_EDIT, MY_MAXW>(B, slow_list, slow_count, s, sm_count);
  } else {
    if (cls == 0) launch_myers<PC_OP_KBAND, 1>(B, slow_list, slow_count, s, sm_count);
    else if (cls == 1) launch_myers<PC_OP_KBAND, 2>(B, slow_list, slow_count, s, sm_count);
    else launch_myers<PC_OP_KBAND, MY_MAXW>(B, slow_list, slow_count, s, sm_count);
  }
}
