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
// Direction bytes go to the warp's scratch slot as [step][lane] x 16 bytes (8 rows of job A, 8 rows of job B): one coalesced
// 512-byte store per warp-step; the traceback is walked by the whole group, both jobs together, a run at a time.
#include "pc_device.cuh"
#include <cstdlib>
#include <type_traits>
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

// ---- traceback ---------------------------------------------------------------------------------------------------
// Same bookkeeping as the reference's recursive TracebackGapAlignment (refine-intron.c:828-890): ops are produced last
// column first and handed over reversed.
//
// The whole group walks BOTH jobs of its pair together.  A single lane chasing one direction byte per step through
// global memory spends the latency of a dependent L2 / DRAM load on every alignment column while the other lanes of its
// warp wait (ncu, round 2: 37 % of the kernel's warp time, long-scoreboard stalls).  Here one round of loads covers a
// window of LANES columns x two row blocks (16 rows) of each job — lane t fetches the 8-row words of column j0 - t — and
// the walk inside the window reads them by shuffle; every lane follows the same (i, j, plane) state.  The ops go to a
// shared-memory strip and leave for global memory already in their final order (no read-modify-write reversal pass);
// alignments longer than the strip fall back to writing reversed and swapping in place.
constexpr int OPS_STRIP = 384;      // bytes of shared memory per job for the ops of one alignment

struct Walk {
  int i, j, state, n_ops, pos0, pos1, pos2, k_end, k_start;
  bool live;
};

__device__ __forceinline__ Walk walk_start(const GapJob &J, int Le, int Ge, int Re, bool live) {
  Walk w;
  if (Re >= Ge) w.state = (Re >= Le) ? 2 : 0; else w.state = (Ge >= Le) ? 1 : 0;
  w.i = J.n; w.j = J.m; w.n_ops = 0; w.pos0 = w.pos1 = w.pos2 = 0; w.k_end = w.k_start = -1;
  w.live = live && J.n > 0 && J.m > 0;
  return w;
}

// The direction the walk takes from row `row` (1-based) of this lane's column, in plane `state`; W0 / W1 = this lane's
// 8-row words of row blocks lb / lb - 1.
__device__ __forceinline__ int dir_at(uint2 W0, uint2 W1, int lb, int row, int state) {
  const uint2 ws = ((row - 1) >> 3) == lb ? W0 : W1;
  const int r = (row - 1) & 7;
  const uint32_t c = ((r < 4 ? ws.x : ws.y) >> (8 * (r & 3))) & 0xffu;
  if (state == 2) return (c & 32) ? 1 : ((c & 16) ? 3 : ((c & 8) ? 2 : 0));
  if (state == 1) return (c & 4) ? 3 : 2;
  return (c & 2) ? 2 : ((c & 1) ? 1 : 0);
}

// One window of one job: lane k holds column j0 - k (W0: row block lb, W1: row block lb - 1).  An alignment is made of RUNS —
// diagonal moves through matching stretches, left moves through the intron — so the walk advances a run at a time: the
// lane of the current column broadcasts its direction; if that is "diagonal" or "left", every later lane looks at the
// cell the walk would reach in ITS column if the run went on and votes; the run is as long as the votes agree.  No
// dependent chain per alignment column: one shuffle and one ballot per run.
template <int LANES>
__device__ __forceinline__ void walk_window(Walk &w, uint2 W0, uint2 W1, int lb, int j0, uint8_t *strip, uint8_t *gops, bool in_smem, int k, int gl0,
                                            unsigned gmask) {
  const int col = j0 - k;
  while (w.i > 0 && w.j > 0 && ((w.i - 1) >> 3) >= lb - 1 && w.j > j0 - LANES) {
    const int d = j0 - w.j;                                  // the lane that holds the current column
    const int dd0 = __shfl_sync(gmask, dir_at(W0, W1, lb, w.i, w.state), d, LANES);
    if (dd0 == 0 || dd0 == 2) {
      const int t = k - d;                                   // moves from the current cell to this lane's column
      const int row = dd0 == 0 ? w.i - t : w.i;
      bool ok = t >= 0 && row >= 1 && col >= 1 && ((row - 1) >> 3) >= lb - 1;
      if (ok) ok = dir_at(W0, W1, lb, row, w.state) == dd0;
      unsigned vote = __ballot_sync(gmask, ok);
      if constexpr (LANES < 32) vote = (vote >> gl0) & ((1u << (LANES & 31)) - 1u);
      const unsigned inv = ~(vote >> d);
      const int run = inv ? __ffs(inv) - 1 : 32 - d;         // consecutive agreeing lanes from lane d on (at least lane d itself)
      if (t >= 0 && t < run) { const uint8_t op = dd0 == 0 ? 0 : 2; if (in_smem) strip[w.n_ops + t] = op; else gops[w.n_ops + t] = op; }
      w.n_ops += run; w.j -= run;
      if (dd0 == 0) w.i -= run;
    } else {
      uint8_t op;
      if (dd0 == 1) { op = 1; --w.i; }
      else {                                                 // the jump between planes: bookkeeping of TracebackGapAlignment
        if (w.state == 2) { w.pos2 = w.j - 1; w.pos0 = w.i; w.k_end = w.n_ops; } else { w.pos1 = w.j - 1; w.k_start = w.n_ops; }
        --w.state;
        op = 2; --w.j;
      }
      if (k == 0) { if (in_smem) strip[w.n_ops] = op; else gops[w.n_ops] = op; }
      ++w.n_ops;
    }
  }
  if (w.i == 0 || w.j == 0) w.live = false;
}

// what is left runs along a border; then the results, and the ops in their final (left to right) order
template <int LANES>
__device__ __forceinline__ void walk_finish(const GapJob &J, Walk &w, uint8_t *strip, bool in_smem, int k, unsigned gmask) {
  const int tail_i = w.i, tail_j = w.j, body = w.n_ops, total = body + tail_i + tail_j;
  if (k == 0) {
    J.res[0] = PC_OK; J.res[1] = total;
    J.res[2] = w.pos0; J.res[3] = w.pos1; J.res[4] = w.pos2;
    J.res[5] = w.k_start >= 0 ? total - 1 - w.k_start : 0;
    J.res[6] = w.k_end >= 0 ? total - 1 - w.k_end : 0;
  }
  uint8_t *ops = J.ops;
  // reversed order = [body (as walked) | tail_i times 1 | tail_j times 2]; final order is that read backwards
  const int sh = tail_i + tail_j;
  if (in_smem) {
    __syncwarp(gmask);                                              // lane 0's strip writes, visible to the group
    for (int a = k; a < tail_j; a += LANES) ops[a] = 2;
    for (int a = k; a < tail_i; a += LANES) ops[tail_j + a] = 1;
    for (int a = k; a < body; a += LANES) ops[sh + a] = strip[body - 1 - a];
  } else {
    // (rare: alignments longer than the strip)  the walked part sits reversed at ops[0 .. body): swap it in place, shift it
    // behind the tails — descending, a whole round read before it is written — then put the tails in front
    __syncwarp(gmask);
    for (int a = k; a < body / 2; a += LANES) { const uint8_t t = ops[a]; ops[a] = ops[body - 1 - a]; ops[body - 1 - a] = t; }
    __syncwarp(gmask);
    if (sh > 0) {
      for (int base = body - 1; base >= 0; base -= LANES) {
        const int a = base - k;
        const uint8_t t = a >= 0 ? ops[a] : 0;
        __syncwarp(gmask);
        if (a >= 0) ops[a + sh] = t;
        __syncwarp(gmask);
      }
      for (int a = k; a < tail_j; a += LANES) ops[a] = 2;
      for (int a = k; a < tail_i; a += LANES) ops[tail_j + a] = 1;
    }
  }
}

template <int LANES>
__device__ void gap_traceback_pair(const GapJob &A, const GapJob &Bj, bool doA, bool doB, const uint4 *dirW, int gl0,
                                   const int fa[3], const int fb[3], uint8_t *stripA, uint8_t *stripB, int k, unsigned gmask) {
  Walk wa = walk_start(A, fa[0], fa[1], fa[2], doA), wb = walk_start(Bj, fb[0], fb[1], fb[2], doB);
  const bool smA = A.n + A.m <= OPS_STRIP, smB = Bj.n + Bj.m <= OPS_STRIP;
  // the 8-row word of (row block lb, column c) was stored by lane gl0 + lb at step c + lb: dirW[step * 32 + lane], .xy = job A, .zw = job B
  const uint2 *d2 = reinterpret_cast<const uint2 *>(dirW);
  while (wa.live || wb.live) {
    // the loads of both windows go out together
    const int lba = (wa.i - 1) >> 3, ja = wa.j, lbb = (wb.i - 1) >> 3, jb = wb.j;
    uint2 a0 = make_uint2(0u, 0u), a1 = a0, b0 = a0, b1 = a0;
    if (wa.live && ja - k >= 1) {
      a0 = d2[((size_t)(ja - k + lba) * 32 + gl0 + lba) * 2];
      if (lba > 0) a1 = d2[((size_t)(ja - k + lba - 1) * 32 + gl0 + lba - 1) * 2];
    }
    if (wb.live && jb - k >= 1) {
      b0 = d2[((size_t)(jb - k + lbb) * 32 + gl0 + lbb) * 2 + 1];
      if (lbb > 0) b1 = d2[((size_t)(jb - k + lbb - 1) * 32 + gl0 + lbb - 1) * 2 + 1];
    }
    if (wa.live) walk_window<LANES>(wa, a0, a1, lba, ja, stripA, A.ops, smA, k, gl0, gmask);
    if (wb.live) walk_window<LANES>(wb, b0, b1, lbb, jb, stripB, Bj.ops, smB, k, gl0, gmask);
  }
  __syncwarp(gmask);
  if (doA) walk_finish<LANES>(A, wa, stripA, smA, k, gmask);
  if (doB) walk_finish<LANES>(Bj, wb, stripB, smB, k, gmask);
}

// MINB = resident CTAs per SM the register allocation is held to (occupancy against register pressure: the sweep is a
// chain of dependent DPX operations, so the issue rate follows the number of resident warps)
template <int LANES, int MINB>
__global__ void __launch_bounds__(128, MINB) k_gap_pairs(PcDevBatch B, int mcap, uint32_t *work, int dbg) {
  extern __shared__ uint32_t sh_codes[];                   // [4 warps][G groups]: (mcap + 1) packed column codes, then two ops strips
  constexpr int G = 32 / LANES;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int k = lane % LANES, grp = lane / LANES;
  const int per_group = (mcap + 1) + 2 * OPS_STRIP / 4;
  uint32_t *codes = sh_codes + (size_t)(wib * G + grp) * per_group;
  uint8_t *stripA = reinterpret_cast<uint8_t *>(codes + mcap + 1), *stripB = stripA + OPS_STRIP;
  const int npairs = (B.n + 1) >> 1;
  const int warp_id = blockIdx.x * 4 + wib, nwarps = pc_active_warps(B, gridDim.x * 4);
  if (warp_id >= nwarps) return;                          // retry rounds: fewer scratch slots than launched warps (no block-wide sync below)
  WarpPool wp = pc_warp_pool(B, warp_id);
  // direction words of the warp: [step][lane] x 16 bytes (8 rows of job A, 8 rows of job B): one coalesced 512-byte store
  // per step.  (Indexed by column and lane, every lane wrote into a sector of its own: 64 store wavefronts per step, and
  // the stores alone were 2 ms of a 4.5 ms kernel.)
  uint4 *dirW = reinterpret_cast<uint4 *>(wp.base);

  // pairs are handed out G at a time from a counter (zeroed by the host with the batch): jobs come heaviest first, and a
  // static deal leaves the warps that got one round more than the others running alone at the end
  for (;;) {
    int q0 = 0;
    if (lane == 0) q0 = (int)atomicAdd(work, (uint32_t)G);
    q0 = __shfl_sync(0xffffffffu, q0, 0);
    if (q0 >= npairs) break;
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
    int mwarp = live ? mmax : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mwarp = max(mwarp, __shfl_xor_sync(0xffffffffu, mwarp, o));
    const unsigned long long need = (unsigned long long)(mwarp + LANES + 1) * 512ull;
    const bool fits = need <= wp.size;                       // the same for the whole warp
    if (live && !fits && k == 0) atomicMax(B.pool_need, need + 1024ull);

    // packed column codes into shared memory, this lane's packed row codes and last-row markers into registers
    if (live && fits)
      for (int j0 = k + 1; j0 <= mmax; j0 += 4 * LANES) {     // four columns per round: eight byte loads in flight, then the stores
        uint8_t ga[4], gb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u * LANES;
          ga[u] = j <= A.m ? A.gen[j - 1] : (uint8_t)'N';
          gb[u] = j <= Bj.m ? Bj.gen[j - 1] : (uint8_t)'N';
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u * LANES;
          if (j <= mmax) codes[j] = sym_code(ga[u]) | (sym_code(gb[u]) << 16);
        }
      }
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
    int steps = (live && fits && !(dbg & 4)) ? mmax + LANES - 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
    __syncwarp();

    uint32_t YL[ROWS], VL[ROWS], VG[ROWS], YR[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { YL[r] = BIAS2 - ONE2; VL[r] = BIAS2; VG[r] = BIAS2; YR[r] = BIAS2 - subv[r]; }   // column 0: all planes 0
    uint32_t inYL = BIAS2 - ONE2, inYR = BIAS2 - ONE2, gcur = 0;
    uint32_t fL = BIAS2, fG = BIAS2, fR = BIAS2;          // packed final cells (n, m) of the two jobs
    // One step of the wavefront.  `guarded` steps test whether the lane is inside its matrix; the steady state needs no
    // test, and runs two steps a round with the L and G values of a row ping-ponging between two register sets (written in
    // place they cost 16 register moves a step: the old value of a cell is still read after the new one exists).
    uint32_t VL2[ROWS], VG2[ROWS];
    const int capSA = k == laneA ? A.m + k : -1, capSB = k == laneB ? Bj.m + k : -1;     // the step in which this lane computes cell (n, m)
    auto step = [&](auto guarded, const int s, const uint32_t (&VLi)[ROWS], const uint32_t (&VGi)[ROWS], uint32_t (&VLo)[ROWS], uint32_t (&VGo)[ROWS]) {
      const int j = s - k;
      const uint32_t dL0 = inYL, dR0 = inYR;              // (row 8k, column j-1)
      inYL = __shfl_up_sync(0xffffffffu, YL[ROWS - 1], 1, LANES);
      inYR = __shfl_up_sync(0xffffffffu, YR[ROWS - 1], 1, LANES);
      gcur = __shfl_up_sync(0xffffffffu, gcur, 1, LANES);
      if (k == 0) { inYL = BIAS2 - ONE2; inYR = BIAS2 - ONE2; gcur = codes[min(s, mmax)]; }
      if (!decltype(guarded)::value || (j >= 1 && j <= mmax && live && fits)) {
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
          const uint32_t llf = VLi[r];
          dL = YL[r]; VLo[r] = v; YL[r] = v - ONE2; upL = YL[r];
          // G plane: stay in the gap, or enter it from L
          const uint32_t glf = VGi[r];
          VGo[r] = pc_vibmax_u16x2(glf, llf, ph, pl);
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
        if (!(dbg & 2)) dirW[(size_t)s * 32 + lane] = make_uint4(wa0, wa1, wb0, wb1);
        else if ((wa0 ^ wb0 ^ wa1 ^ wb1) == 0x12345678u) fL = 0;
        const bool capA = s == capSA, capB = s == capSB;
        if (capA | capB) {                                  // once per job: pick the row holding (n, m)
#pragma unroll
          for (int r = 0; r < ROWS; ++r) {
            const uint32_t vr = YR[r] + subv[r];
            if (capA && r == rowA) { fL = (fL & 0xffff0000u) | (VLo[r] & 0xffffu); fG = (fG & 0xffff0000u) | (VGo[r] & 0xffffu); fR = (fR & 0xffff0000u) | (vr & 0xffffu); }
            if (capB && r == rowB) { fL = (fL & 0xffffu) | (VLo[r] & 0xffff0000u); fG = (fG & 0xffffu) | (VGo[r] & 0xffff0000u); fR = (fR & 0xffffu) | (vr & 0xffff0000u); }
          }
        }
      }
    };
    // lead-in (lanes joining the wavefront), steady state (every lane of the warp inside its matrix: no guard, two steps a
    // round), lead-out
    int s_hi = live ? mmax : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s_hi = min(s_hi, __shfl_xor_sync(0xffffffffu, s_hi, o));
    if (steps == 0 || (dbg & 8)) s_hi = 0;
    int s = 1;
    for (; s <= steps && s < LANES; ++s) step(std::true_type{}, s, VL, VG, VL, VG);
    for (; s + 1 <= s_hi; s += 2) {
      step(std::false_type{}, s, VL, VG, VL2, VG2);
      step(std::false_type{}, s + 1, VL2, VG2, VL, VG);
    }
    for (; s <= steps; ++s) step(std::true_type{}, s, VL, VG, VL, VG);
    __syncwarp();
    // final cells to the tracing lanes (group lane 0 walks job A, group lane 1 job B)
    const int base = grp * LANES;
    const uint32_t aL = __shfl_sync(0xffffffffu, fL, base + (live ? laneA : 0)), aG = __shfl_sync(0xffffffffu, fG, base + (live ? laneA : 0)),
                   aR = __shfl_sync(0xffffffffu, fR, base + (live ? laneA : 0));
    const uint32_t bL = __shfl_sync(0xffffffffu, fL, base + (live ? laneB : 0)), bG = __shfl_sync(0xffffffffu, fG, base + (live ? laneB : 0)),
                   bR = __shfl_sync(0xffffffffu, fR, base + (live ? laneB : 0));
    if (live) {                                            // live / fits / ok are the same for all lanes of a group
      const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << base);
      if (!fits) { if (k == 0) A.res[0] = PC_E_POOL; if (k == 1 && hasB) Bj.res[0] = PC_E_POOL; }
      else {
        if (!A.ok && k == 0) A.res[0] = PC_E_OUTCAP;
        if (hasB && !Bj.ok && k == 0) Bj.res[0] = PC_E_OUTCAP;
        const int fa[3] = {(int)(aL & 0xffffu) - 0x4000, (int)(aG & 0xffffu) - 0x4000, (int)(aR & 0xffffu) - 0x4000};
        const int fb[3] = {(int)(bL >> 16) - 0x4000, (int)(bG >> 16) - 0x4000, (int)(bR >> 16) - 0x4000};
        if (!(dbg & 1)) gap_traceback_pair<LANES>(A, Bj, A.ok, hasB && Bj.ok, dirW, base, fa, fb, stripA, stripB, k, gmask);
      }
    }
    __syncwarp();
  }
}

template <int LANES, int MINB>
void launch_pairs(const PcDevBatch &B, int mcap, uint32_t *work, cudaStream_t s, int sm_count) {
  constexpr int G = 32 / LANES;
  const int npairs = (B.n + 1) / 2;
  const int ctas_needed = (npairs + 4 * G - 1) / (4 * G);
  const size_t sh = (size_t)4 * G * ((mcap + 1) + 2 * OPS_STRIP / 4) * sizeof(uint32_t);
  pc_smem_optin((const void *)k_gap_pairs<LANES, MINB>, 200 * 1024);
  // persistent CTAs: exactly as many as are resident at once (registers limit this kernel), else the rest runs as a tail wave
  const int per_sm = pc_cached_occupancy((const void *)k_gap_pairs<LANES, MINB>, 128, sh);
  int grid = ctas_needed < sm_count * per_sm ? ctas_needed : sm_count * per_sm;
  if (B.max_warps > 0 && grid > (B.max_warps + 3) / 4) grid = (B.max_warps + 3) / 4;
  if (grid < 1) grid = 1;
  PcDevBatch C = B;
  C.slots = (B.max_warps > 0 && B.max_warps < grid * 4) ? B.max_warps : grid * 4;
  static const int dbg = getenv("PC_GAP_DBG") ? atoi(getenv("PC_GAP_DBG")) : 0;      /* timing experiments only: 1 = no traceback, 2 = no direction stores, 4 = no sweep, 8 = guarded steps only */
  k_gap_pairs<LANES, MINB><<<grid, 128, sh, s>>>(C, mcap, work, dbg);
  PC_COUNT_LAUNCH(1);
}

}  // namespace

// cls 0/1/2: n <= 64 / 128 / 256 with 1 <= m <= PC_GAP_FAST_MAX_M (the generic wavefront kernel takes the rest).
template <int MINB>
static void launch_cls(int cls, const PcDevBatch &B, int max_m, uint32_t *work, cudaStream_t s, int sm_count) {
  if (cls == 0) launch_pairs<8, MINB>(B, max_m, work, s, sm_count);
  else if (cls == 1) launch_pairs<16, MINB>(B, max_m, work, s, sm_count);
  else launch_pairs<32, MINB>(B, max_m, work, s, sm_count);
}
// work: a device counter, zero at launch (the segment's otherwise unused hand-over counter)
void pc_launch_gap_pairs(int cls, const PcDevBatch &B, int max_m, uint32_t *work, cudaStream_t s, int sm_count) {
  static const int variant = getenv("PC_GAP_MINB") ? atoi(getenv("PC_GAP_MINB")) : PC_GAP_MINB_DEFAULT;      /* experiments */
  if (variant == 3) launch_cls<3>(cls, B, max_m, work, s, sm_count);      /* 3 CTAs/SM: what the kernel ran at before the store layout changed */
  else launch_cls<4>(cls, B, max_m, work, s, sm_count);
}
