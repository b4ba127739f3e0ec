// k_gap.cu — compute_gap_alignment (reference src/refine-intron.c:560-890) as a register-resident, 16x2-packed
// wavefront for sm_100a.
//
// Shape of the work: n = EST window (about 60 nt + the gap on P), m = donor-suffix30 |
// intron-prefix70 | intron-suffix70 | acceptor-prefix30 (<= 200 with the default windows); three score planes
// L / G / R, match +1, mismatch -1, indel -1, a free jump L->G->R along a row, no end-gap penalty on the last row
// of R.  One direction byte per cell feeds the traceback.
//
// Mapping.  TWO jobs share every instruction: their scores travel as the low and the high 16-bit half of one
// register and are combined with the Blackwell DPX forms (VIMNMX.U16x2 with its two predicate outputs,
// VIMNMX3.U16x2); the predicates ARE the direction bits, so no extra compare is spent on the traceback matrix.
// A group of LANES lanes (8, 16 or 32, picked from n) owns one pair; lane k owns rows 8k+1 .. 8k+8 and sweeps the
// columns, one column per step, one step behind lane k-1 (skewed wavefront): the cell state lives in registers,
// the only traffic between lanes is two boundary values and the genome code, by shuffle.  Scores are kept with a
// +0x4000 bias per half so that every add is a plain 32-bit add (either pipe) and every max is unsigned.
//   stored per row:  YL = L-1, VL = L, VG = G, YR = R-1 (R on the job's last row: that folds the free end gap in)
// Direction bytes go to the warp's scratch slot as [lane][column][8 rows] = one 8-byte store per job per step and
// stay L2/L1-resident for the traceback, which lanes 0 and 1 of the group walk for the two jobs.
#include "pc_device.cuh"
#include <cstdlib>
#define PC_GAP_MINB_DEFAULT 4

namespace {

constexpr uint32_t BIAS2 = 0x40004000u, ONE2 = 0x00010001u, TWO2 = 0x00020002u;
constexpr int ROWS = 8;

__device__ __forceinline__ uint32_t sym_code(uint8_t c) { return (c == 'N' || c == 'n') ? 0u : ((uint32_t)c << 1); }

struct GapJob {
  const uint8_t *est, *gen;
  int n, m;
  int32_t *res;
  uint8_t *ops;
  bool ok;
};

__device__ __forceinline__ GapJob gap_job(const PcDevBatch &B, int slot) {
  GapJob J;
  const uint32_t ji = B.idx[slot];
  const pc_job *job = B.jobs + ji;
  J.res = B.res + (size_t)ji * PC_RES_INTS;
  J.est = B.arena + job->a_off;
  J.n = (int)job->a_len;
  J.gen = ((job->flags & PC_B_IN_GENOME) ? B.genome : B.arena) + job->b_off;
  J.m = (int)job->b_len;
  J.ops = B.var_out + job->out_off;
  J.ok = (uint32_t)(J.n + J.m) <= job->out_cap;
  return J;
}

// Walk the direction bytes back from (n, m).  Same bookkeeping as the reference's recursive
// TracebackGapAlignment (refine-intron.c:828-890): ops are produced last column first, the caller reverses them.
//
// The whole group walks together.  A single lane chasing one direction byte per step through global memory spends
// the latency of a dependent L2 / DRAM load on every alignment column (~260 of them) while the other lanes of its warp
// wait: that was half of the kernel's time (ncu: long-scoreboard stalls 2.6 per issue).  Here lane t fetches the
// 8-row word of column j - t, so one round of loads covers LANES columns of the current row block; the walk inside
// that window reads the words by shuffle.  Every lane follows the same (i, j, plane) state; lane 0 writes.
template <int LANES>
__device__ int gap_traceback(const GapJob &J, const uint8_t *dir, int mstride, int Le, int Ge, int Re, int k, unsigned gmask) {
  int state;
  if (Re >= Ge) state = (Re >= Le) ? 2 : 0; else state = (Ge >= Le) ? 1 : 0;
  int pos0 = 0, pos1 = 0, pos2 = 0, k_end = -1, k_start = -1;
  int i = J.n, j = J.m, n_ops = 0;
  uint8_t *ops = J.ops;
  const uint2 *dir64 = reinterpret_cast<const uint2 *>(dir);
  while (i > 0 && j > 0) {
    const int lb = (i - 1) >> 3, j0 = j;
    const int col = j0 - k;
    uint2 w = make_uint2(0u, 0u);
    if (col >= 1) w = dir64[(size_t)lb * mstride + col];
    while (i > 0 && j > 0 && ((i - 1) >> 3) == lb && j > j0 - LANES) {
      const int src = j0 - j;
      const uint32_t lo = __shfl_sync(gmask, w.x, src, LANES), hi = __shfl_sync(gmask, w.y, src, LANES);
      const int r = (i - 1) & 7;
      const uint32_t c = ((r < 4 ? lo : hi) >> (8 * (r & 3))) & 0xffu;
      int dd;
      if (state == 2) dd = (c & 32) ? 1 : ((c & 16) ? 3 : ((c & 8) ? 2 : 0));
      else if (state == 1) dd = (c & 4) ? 3 : 2;
      else dd = (c & 2) ? 2 : ((c & 1) ? 1 : 0);
      uint8_t op;
      if (dd == 0) { op = 0; --i; --j; }
      else if (dd == 1) { op = 1; --i; }
      else {
        if (dd == 3) {
          if (state == 2) { pos2 = j - 1; pos0 = i; k_end = n_ops; } else { pos1 = j - 1; k_start = n_ops; }
          --state;
        }
        op = 2; --j;
      }
      if (k == 0) ops[n_ops] = op;
      ++n_ops;
    }
  }
  for (int a = k; a < i; a += LANES) ops[n_ops + a] = 1;            // what is left runs along a border
  n_ops += i;
  for (int a = k; a < j; a += LANES) ops[n_ops + a] = 2;
  n_ops += j;
  if (k == 0) {
    J.res[0] = PC_OK; J.res[1] = n_ops;
    J.res[2] = pos0; J.res[3] = pos1; J.res[4] = pos2;
    J.res[5] = k_start >= 0 ? n_ops - 1 - k_start : 0;
    J.res[6] = k_end >= 0 ? n_ops - 1 - k_end : 0;
  }
  return n_ops;
}

// MINB = resident CTAs per SM the register allocation is held to (occupancy against register pressure: the sweep is a
// chain of dependent DPX operations, so the issue rate follows the number of resident warps)
template <int LANES, int MINB>
__global__ void __launch_bounds__(128, MINB) k_gap_pairs(PcDevBatch B, int mcap) {
  extern __shared__ uint32_t sh_codes[];                   // [4 warps][G groups][mcap + 1] packed column codes
  constexpr int G = 32 / LANES;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int k = lane % LANES, grp = lane / LANES;
  uint32_t *codes = sh_codes + (size_t)(wib * G + grp) * (mcap + 1);
  const int npairs = (B.n + 1) >> 1;
  const int warp_id = blockIdx.x * 4 + wib, nwarps = pc_active_warps(B, gridDim.x * 4), ngroups = nwarps * G;
  if (warp_id >= nwarps) return;                          // retry rounds: fewer scratch slots than launched warps (no block-wide sync below)
  WarpPool wp = pc_warp_pool(B, warp_id);
  const unsigned long long gshare = (wp.size / G) & ~255ull;
  uint8_t *gbase = wp.base + gshare * grp;

  for (int q0 = warp_id * G; q0 < npairs; q0 += ngroups) {       // warp-uniform trip count
    const int q = q0 + grp;
    const bool live = q < npairs;
    GapJob A, Bj;
    bool hasB = false;
    if (live) {
      A = gap_job(B, 2 * q);
      hasB = 2 * q + 1 < B.n;
      Bj = hasB ? gap_job(B, 2 * q + 1) : A;
    } else {
      A.est = A.gen = nullptr; A.n = A.m = 0; A.res = nullptr; A.ops = nullptr; A.ok = false; Bj = A;
    }
    const int mmax = max(A.m, Bj.m);
    const int mstride = mmax + 1;
    const unsigned long long need = 2ull * LANES * mstride * 8ull;
    bool fits = need <= gshare;
    if (live && !fits && k == 0) atomicMax(B.pool_need, need * G + 1024ull);
    uint8_t *dirA = gbase, *dirB = gbase + (size_t)LANES * mstride * 8;

    // packed column codes into shared memory, this lane's packed row codes and last-row markers into registers
    if (live && fits)
      for (int j = k + 1; j <= mmax; j += LANES)
        codes[j] = (j <= A.m ? sym_code(A.gen[j - 1]) : 0u) | ((j <= Bj.m ? sym_code(Bj.gen[j - 1]) : 0u) << 16);
    uint32_t e[ROWS], e2[ROWS], subv[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int i = k * ROWS + r + 1;
      const uint32_t ea = (live && i <= A.n) ? sym_code(A.est[i - 1]) : 0u;
      const uint32_t eb = (live && i <= Bj.n) ? sym_code(Bj.est[i - 1]) : 0u;
      e[r] = ea | (eb << 16);
      e2[r] = __vminu2(e[r], TWO2);                        // 0 for N, else 2
      subv[r] = (i == A.n ? 0u : 1u) | ((i == Bj.n ? 0u : 1u) << 16);
    }
    const int laneA = (A.n - 1) >> 3, rowA = (A.n - 1) & 7, laneB = (Bj.n - 1) >> 3, rowB = (Bj.n - 1) & 7;
    int steps = (live && fits) ? mmax + LANES - 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
    __syncwarp();

    uint32_t YL[ROWS], VL[ROWS], VG[ROWS], YR[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { YL[r] = BIAS2 - ONE2; VL[r] = BIAS2; VG[r] = BIAS2; YR[r] = BIAS2 - subv[r]; }   // column 0: all planes 0
    uint32_t inYL = BIAS2 - ONE2, inYR = BIAS2 - ONE2, gcur = 0;
    uint32_t fL = BIAS2, fG = BIAS2, fR = BIAS2;          // packed final cells (n, m) of the two jobs
    for (int s = 1; s <= steps; ++s) {
      const int j = s - k;
      const uint32_t dL0 = inYL, dR0 = inYR;              // (row 8k, column j-1)
      inYL = __shfl_up_sync(0xffffffffu, YL[ROWS - 1], 1, LANES);
      inYR = __shfl_up_sync(0xffffffffu, YR[ROWS - 1], 1, LANES);
      gcur = __shfl_up_sync(0xffffffffu, gcur, 1, LANES);
      if (k == 0) { inYL = BIAS2 - ONE2; inYR = BIAS2 - ONE2; gcur = codes[min(s, mmax)]; }
      if (j >= 1 && j <= mmax && live && fits) {
        const uint32_t g2 = __vminu2(gcur, TWO2);
        uint32_t dL = dL0, dR = dR0, upL = inYL, upR = inYR;
        uint32_t wa0 = 0, wa1 = 0, wb0 = 0, wb1 = 0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          bool ph, pl;
          // 2 on a match (equal bytes, or either side N), 0 on a mismatch, per half
          const uint32_t c2m = TWO2 - __vimin3_u16x2(e[r] ^ gcur, e2[r], g2);
          uint32_t bits_a = 0, bits_b = 0;
          // L plane: diagonal, then up if strictly better, then left if strictly better
          uint32_t v = pc_vibmax_u16x2(dL + c2m, upL, ph, pl);
          if (!pl) bits_a |= 1u; if (!ph) bits_b |= 1u;
          v = pc_vibmax_u16x2(v, YL[r], ph, pl);
          if (!pl) bits_a |= 2u; if (!ph) bits_b |= 2u;
          const uint32_t llf = VL[r];
          dL = YL[r]; VL[r] = v; YL[r] = v - ONE2; upL = YL[r];
          // G plane: stay in the gap, or enter it from L
          const uint32_t glf = VG[r];
          VG[r] = pc_vibmax_u16x2(glf, llf, ph, pl);
          if (!pl) bits_a |= 4u; if (!ph) bits_b |= 4u;
          // R plane: diagonal, left (free on the job's last row), jump from G, up
          v = pc_vibmax_u16x2(dR + c2m, YR[r], ph, pl);
          if (!pl) bits_a |= 8u; if (!ph) bits_b |= 8u;
          v = pc_vibmax_u16x2(v, glf, ph, pl);
          if (!pl) bits_a |= 16u; if (!ph) bits_b |= 16u;
          v = pc_vibmax_u16x2(v, upR, ph, pl);
          if (!pl) bits_a |= 32u; if (!ph) bits_b |= 32u;
          dR = YR[r]; YR[r] = v - subv[r]; upR = v - ONE2;
          if (r < 4) { wa0 |= bits_a << (8 * r); wb0 |= bits_b << (8 * r); }
          else { wa1 |= bits_a << (8 * (r - 4)); wb1 |= bits_b << (8 * (r - 4)); }
        }
        *reinterpret_cast<uint2 *>(dirA + ((size_t)k * mstride + j) * 8) = make_uint2(wa0, wa1);
        *reinterpret_cast<uint2 *>(dirB + ((size_t)k * mstride + j) * 8) = make_uint2(wb0, wb1);
        const bool capA = j == A.m && k == laneA, capB = j == Bj.m && k == laneB;
        if (capA | capB) {                                  // once per job: pick the row holding (n, m)
#pragma unroll
          for (int r = 0; r < ROWS; ++r) {
            const uint32_t vr = YR[r] + subv[r];
            if (capA && r == rowA) { fL = (fL & 0xffff0000u) | (VL[r] & 0xffffu); fG = (fG & 0xffff0000u) | (VG[r] & 0xffffu); fR = (fR & 0xffff0000u) | (vr & 0xffffu); }
            if (capB && r == rowB) { fL = (fL & 0xffffu) | (VL[r] & 0xffff0000u); fG = (fG & 0xffffu) | (VG[r] & 0xffff0000u); fR = (fR & 0xffffu) | (vr & 0xffff0000u); }
          }
        }
      }
    }
    __syncwarp();
    // final cells to the tracing lanes (group lane 0 walks job A, group lane 1 job B)
    const int base = grp * LANES;
    const uint32_t aL = __shfl_sync(0xffffffffu, fL, base + (live ? laneA : 0)), aG = __shfl_sync(0xffffffffu, fG, base + (live ? laneA : 0)),
                   aR = __shfl_sync(0xffffffffu, fR, base + (live ? laneA : 0));
    const uint32_t bL = __shfl_sync(0xffffffffu, fL, base + (live ? laneB : 0)), bG = __shfl_sync(0xffffffffu, fG, base + (live ? laneB : 0)),
                   bR = __shfl_sync(0xffffffffu, fR, base + (live ? laneB : 0));
    int lenA = 0, lenB = 0;
    if (live) {                                            // live / fits / ok are the same for all lanes of a group
      const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << base);
      if (!fits) { if (k == 0) A.res[0] = PC_E_POOL; if (k == 1 && hasB) Bj.res[0] = PC_E_POOL; }
      else {
        if (!A.ok) { if (k == 0) A.res[0] = PC_E_OUTCAP; }
        else lenA = gap_traceback<LANES>(A, dirA, mstride, (int)(aL & 0xffffu) - 0x4000, (int)(aG & 0xffffu) - 0x4000, (int)(aR & 0xffffu) - 0x4000, k, gmask);
        if (hasB) {
          if (!Bj.ok) { if (k == 0) Bj.res[0] = PC_E_OUTCAP; }
          else lenB = gap_traceback<LANES>(Bj, dirB, mstride, (int)(bL >> 16) - 0x4000, (int)(bG >> 16) - 0x4000, (int)(bR >> 16) - 0x4000, k, gmask);
        }
      }
    }
    __syncwarp();
    if (live) {
      for (int a = k; a < lenA / 2; a += LANES) { uint8_t t = A.ops[a]; A.ops[a] = A.ops[lenA - 1 - a]; A.ops[lenA - 1 - a] = t; }
      for (int a = k; a < lenB / 2; a += LANES) { uint8_t t = Bj.ops[a]; Bj.ops[a] = Bj.ops[lenB - 1 - a]; Bj.ops[lenB - 1 - a] = t; }
    }
    __syncwarp();
  }
}

template <int LANES, int MINB>
void launch_pairs(const PcDevBatch &B, int mcap, cudaStream_t s, int sm_count) {
  constexpr int G = 32 / LANES;
  const int npairs = (B.n + 1) / 2;
  const int ctas_needed = (npairs + 4 * G - 1) / (4 * G);
  const size_t sh = (size_t)4 * G * (mcap + 1) * sizeof(uint32_t);
  pc_smem_optin((const void *)k_gap_pairs<LANES, MINB>, 200 * 1024);
  // persistent CTAs: exactly as many as are resident at once (registers limit this kernel), else the rest runs as a tail wave
  const int per_sm = pc_cached_occupancy((const void *)k_gap_pairs<LANES, MINB>, 128, sh);
  int grid = ctas_needed < sm_count * per_sm ? ctas_needed : sm_count * per_sm;
  if (B.max_warps > 0 && grid > (B.max_warps + 3) / 4) grid = (B.max_warps + 3) / 4;
  if (grid < 1) grid = 1;
  PcDevBatch C = B;
  C.slots = (B.max_warps > 0 && B.max_warps < grid * 4) ? B.max_warps : grid * 4;
  k_gap_pairs<LANES, MINB><<<grid, 128, sh, s>>>(C, mcap);
  PC_COUNT_LAUNCH(1);
}

}  // namespace

// cls 0/1/2: n <= 64 / 128 / 256 with 1 <= m <= PC_GAP_FAST_MAX_M (the generic wavefront kernel takes the rest).
template <int MINB>
static void launch_cls(int cls, const PcDevBatch &B, int max_m, cudaStream_t s, int sm_count) {
  if (cls == 0) launch_pairs<8, MINB>(B, max_m, s, sm_count);
  else if (cls == 1) launch_pairs<16, MINB>(B, max_m, s, sm_count);
  else launch_pairs<32, MINB>(B, max_m, s, sm_count);
}
void pc_launch_gap_pairs(int cls, const PcDevBatch &B, int max_m, cudaStream_t s, int sm_count) {
  static const int variant = getenv("PC_GAP_MINB") ? atoi(getenv("PC_GAP_MINB")) : PC_GAP_MINB_DEFAULT;      /* experiments */
  if (variant == 3) launch_cls<3>(cls, B, max_m, s, sm_count);
  else if (variant == 5) launch_cls<5>(cls, B, max_m, s, sm_count);
  else if (variant == 6) launch_cls<6>(cls, B, max_m, s, sm_count);
  else launch_cls<4>(cls, B, max_m, s, sm_count);
}
