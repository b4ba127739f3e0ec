// k_dp.cu — warp-synchronous integer DP kernels for est-fact (sm_100a).
//
// One warp owns one job.  Every DP here is evaluated as an anti-diagonal wavefront: the cells with
// i + j = s are independent, so the lanes of the warp take them 32 at a time.  The state of the sweep is one
// value per DIAGONAL k = j - i (the most recent cell on that diagonal), kept in shared memory: at step s the
// cell on diagonal k needs the value of its own diagonal (written at step s-2) and of diagonals k-1 and k+1
// (written at step s-1), so one array and one __syncwarp per step are enough.
//
// Banding: the unit-cost alignments only sweep the diagonals [lo, hi] that can hold an optimal path
// (everything outside reads as +inf).  A banded result s with s <= half-width d is exact, and so is every
// direction on the traceback path (cells of value <= d lie within d diagonals of the origin),
// which is what lets compute_alignment run in O((n+m)·d) here instead of the reference's O(n·m) bytes of Mdir.
#include "pc_device.cuh"

namespace {

struct Str {            // a byte string walked forwards (step +1) or backwards (step -1)
  const uint8_t *p;
  int step;
  __device__ __forceinline__ uint8_t at(int i) const { return p[i * step]; }
};

__device__ __forceinline__ int ceil_half(int a) { return a >= 0 ? (a + 1) >> 1 : -((-a) >> 1); }
__device__ __forceinline__ int floor_half(int a) { return a >= 0 ? a >> 1 : -((-a + 1) >> 1); }

enum { ST_NONE = 0, ST_DIR = 1, ST_MAT = 2 };

// Unit-cost edit DP over rows (length nr) x cols (length nc) restricted to diagonals lo..hi (lo <= 0,
// lo <= nc-nr <= hi).  Tie-break of the reference (compute-alignments.c:114-136): diagonal, then up
// (row char against '-') only if strictly cheaper, then left only if strictly cheaper.
// STORE = ST_DIR: dir[i*W + (k-lo)] gets 0/1/2.  STORE = ST_MAT: mat[i*(nc+1)+j] gets the cell value.
template <bool WILD, int STORE>
__device__ uint32_t banded_dp(Str rows, int nr, Str cols, int nc, int lo, int hi, uint32_t *H, uint8_t *dir,
                              uint32_t *mat, int lane) {
  const int W = hi - lo + 1;
  for (int x = lane; x < W; x += 32) H[x] = PC_INF;
  __syncwarp();
  const int last = nr + nc;
  for (int s = 0; s <= last; ++s) {
    int i_lo = max(max(0, s - nc), ceil_half(s - hi));
    int i_hi = min(min(nr, s), floor_half(s - lo));
    for (int i = i_lo + lane; i <= i_hi; i += 32) {
      const int j = s - i;
      const int x = j - i - lo;
      uint32_t v;
      uint8_t d = 0;
      if (i == 0) { v = (uint32_t)j; d = 2; }
      else if (j == 0) { v = (uint32_t)i; d = 1; }
      else {
        const uint8_t rc = rows.at(i - 1), cc = cols.at(j - 1);
        bool eq = rc == cc;
        if (WILD) eq = eq || pc_is_n(rc) || pc_is_n(cc);
        v = H[x] + (eq ? 0u : 1u);
        const uint32_t up = (x + 1 < W ? H[x + 1] : PC_INF) + 1u;     // (i-1, j)
        const uint32_t lf = (x > 0 ? H[x - 1] : PC_INF) + 1u;         // (i, j-1)
        if (v > up) { v = up; d = 1; }
        if (v > lf) { v = lf; d = 2; }
      }
      H[x] = v;
      if (STORE == ST_DIR) dir[(size_t)i * W + x] = d;
      if (STORE == ST_MAT) mat[(size_t)i * (nc + 1) + j] = v;
    }
    __syncwarp();
  }
  return H[nc - nr - lo];
}

__device__ __forceinline__ void warp_reverse(uint8_t *p, int len, int lane) {
  for (int a = lane; a < len / 2; a += 32) {
    uint8_t t = p[a]; p[a] = p[len - 1 - a]; p[len - 1 - a] = t;
  }
}

struct JobView {
  const pc_job *job;
  int32_t *res;
  const uint8_t *a, *b;
  int la, lb;
};

__device__ __forceinline__ JobView view(const PcDevBatch &B, int w) {
  JobView v;
  const uint32_t ji = B.idx[w];
  v.job = B.jobs + ji;
  v.res = B.res + (size_t)ji * PC_RES_INTS;
  v.a = B.arena + v.job->a_off;
  v.la = (int)v.job->a_len;
  v.b = ((v.job->flags & PC_B_IN_GENOME) ? B.genome : B.arena) + v.job->b_off;
  v.lb = (int)v.job->b_len;
  return v;
}

__device__ __forceinline__ uint32_t *state_mem(const PcDevBatch &B, WarpPool &wp, uint32_t *smem, size_t ints, int lane, bool &ok) {
  ok = true;
  if (ints <= PC_SMEM_INTS_PER_WARP) return smem;
  uint32_t *p = (uint32_t *)pc_pool_alloc(B, wp, ints * 4ull, lane);
  ok = p != nullptr;
  return p;
}

#define PC_FAIL(code) do { if (lane == 0) J.res[0] = (code); return; } while (0)

// ---- DP A: compute_alignment (src/compute-alignments.c:39-207) --------------------------------------------
__device__ void op_align(const PcDevBatch &B, WarpPool &wp, int w, uint32_t *smem, int lane) {
  JobView J = view(B, w);
  const int n = J.la, m = J.lb;
  uint8_t *ops = B.var_out + J.job->out_off;
  if ((uint32_t)(n + m) > J.job->out_cap) PC_FAIL(PC_E_OUTCAP);
  Str rows{J.a, 1}, cols{J.b, 1};
  const int off = m - n;
  // pass 1: score, doubling the half-width until the banded score fits inside it
  int d = max(4, (62 - abs(off)) / 2);
  uint32_t score;
  for (;;) {
    int lo = max(min(0, off) - d, -n), hi = min(max(0, off) + d, m);
    bool ok;
    uint32_t *H = state_mem(B, wp, smem, (size_t)(hi - lo + 1), lane, ok);
    if (!ok) PC_FAIL(PC_E_POOL);
    score = banded_dp<true, ST_NONE>(rows, n, cols, m, lo, hi, H, nullptr, nullptr, lane);
    if (score <= (uint32_t)d || (lo == -n && hi == m)) break;
    d = (score < 2u * (uint32_t)d) ? 2 * d : (int)min(score, (uint32_t)(n + m));
  }
  // pass 2: directions inside the exact band of half-width = score
  const int lo = max(min(0, off) - (int)score, -n), hi = min(max(0, off) + (int)score, m);
  const int W = hi - lo + 1;
  bool ok;
  uint32_t *H = state_mem(B, wp, smem, (size_t)W, lane, ok);
  if (!ok) PC_FAIL(PC_E_POOL);
  uint8_t *dir = pc_pool_alloc(B, wp, (unsigned long long)(n + 1) * W, lane);
  if (!dir) PC_FAIL(PC_E_POOL);
  banded_dp<true, ST_DIR>(rows, n, cols, m, lo, hi, H, dir, nullptr, lane);
  __syncwarp();
  int k = 0;
  if (lane == 0) {
    int i = n, j = m;
    while (i > 0 || j > 0) {            // border cells carry the pure-gap direction
      const uint8_t dd = dir[(size_t)i * W + (j - i - lo)];
      ops[k++] = dd;
      if (dd == 0) { --i; --j; } else if (dd == 1) --i; else --j;
    }
    J.res[0] = PC_OK; J.res[1] = (int32_t)score; J.res[2] = k;
  }
  k = __shfl_sync(0xffffffffu, k, 0);
  __syncwarp();
  warp_reverse(ops, k, lane);
}

// ---- DP B: K_band_edit_distance (src/compute-alignments.c:319-453) ----------------------------------------
__device__ bool warp_equal(const uint8_t *a, const uint8_t *b, int len, int lane) {
  bool diff = false;
  for (int i = lane; i < len; i += 32) diff |= a[i] != b[i];
  return !__any_sync(0xffffffffu, diff);
}

__device__ void op_kband(const PcDevBatch &B, WarpPool &wp, int w, uint32_t *smem, int lane) {
  JobView J = view(B, w);
  const uint32_t k = (uint32_t)J.job->p0;
  int ok_flag; uint32_t edit;
  const uint8_t *s1 = J.a, *s2 = J.b;
  int n = J.la, m = J.lb;
  if (n == m && warp_equal(s1, s2, n, lane)) { ok_flag = 1; edit = 0; }
  else if (k == 0) { ok_flag = 0; edit = 1; }
  else {
    if (n < m) { const uint8_t *t = s1; s1 = s2; s2 = t; int q = n; n = m; m = q; }
    if ((uint32_t)(n - m) > k) { ok_flag = 0; edit = (uint32_t)(n - m); }
    else {
      // rows over the shorter string s2, columns over s1; full matrix when 2k+1 >= n
      int lo, hi;
      if (2ull * k + 1ull >= (unsigned long long)n) { lo = -m; hi = n; } else { lo = -(int)k; hi = (int)k; }
      lo = max(lo, -m); hi = min(hi, n);
      bool ok;
      uint32_t *H = state_mem(B, wp, smem, (size_t)(hi - lo + 1), lane, ok);
      if (!ok) PC_FAIL(PC_E_POOL);
      edit = banded_dp<false, ST_NONE>(Str{s2, 1}, m, Str{s1, 1}, n, lo, hi, H, nullptr, nullptr, lane);
      ok_flag = edit <= k;
    }
  }
  if (lane == 0) { J.res[0] = PC_OK; J.res[1] = ok_flag; J.res[2] = (int32_t)edit; }
}

// ---- plain edit distance, last cell (src/refine.c:51, src/compute-alignments.c:235) -----------------------
__device__ void op_edit(const PcDevBatch &B, WarpPool &wp, int w, uint32_t *smem, int lane) {
  JobView J = view(B, w);
  bool ok;
  uint32_t *H = state_mem(B, wp, smem, (size_t)J.la + J.lb + 1, lane, ok);
  if (!ok) PC_FAIL(PC_E_POOL);
  uint32_t d = banded_dp<false, ST_NONE>(Str{J.a, 1}, J.la, Str{J.b, 1}, J.lb, -J.la, J.lb, H, nullptr, nullptr, lane);
  if (lane == 0) { J.res[0] = PC_OK; J.res[1] = (int32_t)d; }
}

// ---- Burset frequencies (src/refine-intron.c:362-556) -----------------------------------------------------
__constant__ uint8_t c_burset[58][5] = {
  {'A','A','A','G',1},{'A','A','A','T',1},{'A','A','G','T',1},{'A','C','C','C',1},{'A','G','A','C',1},
  {'A','G','A','G',5},{'A','G','C','T',2},{'A','G','G','C',1},{'A','G','T','G',2},{'A','T','A','A',1},
  {'A','T','A','C',8},{'A','T','A','G',7},{'A','T','A','T',2},{'A','T','G','C',1},{'A','T','G','T',1},
  {'C','A','A','G',1},{'C','A','T','T',1},{'C','C','A','G',2},{'C','G','A','G',1},{'C','G','C','A',1},
  {'C','T','A','C',2},{'C','T','C','A',1},{'G','A','A','G',8},{'G','A','G','T',1},{'G','A','T','C',1},
  {'G','A','T','G',1},{'G','C','A','G',126},{'G','C','G','G',1},{'G','C','T','A',1},{'G','G','A','C',1},
  {'G','G','A','G',11},{'G','G','C','A',1},{'G','G','G','A',2},{'G','G','T','C',2},{'G','T','A','G',200},
  {'G','T','A','C',4},{'G','T','A','T',2},{'G','T','C','A',9},{'G','T','C','G',4},{'G','T','C','T',3},
  {'G','T','G','C',1},{'G','T','G','G',10},{'G','T','G','T',1},{'G','T','T','A',7},{'G','T','T','C',2},
  {'G','T','T','G',8},{'G','T','T','T',2},{'T','A','A','G',6},{'T','A','C','G',1},{'T','A','T','C',1},
  {'T','C','A','G',1},{'T','C','G','G',1},{'T','G','A','C',1},{'T','G','A','G',7},{'T','G','G','G',2},
  {'T','T','A','G',5},{'T','T','A','T',1},{'T','T','G','G',1}};

__device__ __forceinline__ uint8_t up_c(uint8_t c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

// len_t >= 0: bytes at or after len_t read as NUL (PC_B_NUL_AFTER); len_t < 0: read whatever follows t
__device__ int burset_freq(const uint8_t *t, int cut1, int cut2, int len_t) {
  if (cut2 < 2) return 0;
  const uint8_t d0 = (len_t >= 0 && cut1 >= len_t) ? 0 : up_c(t[cut1]);
  const uint8_t d1 = (!d0 || (len_t >= 0 && cut1 + 1 >= len_t)) ? 0 : up_c(t[cut1 + 1]);
  const uint8_t a0 = up_c(t[cut2 - 2]), a1 = up_c(t[cut2 - 1]);
  for (int e = 0; e < 58; ++e)
    if (c_burset[e][0] == d0 && c_burset[e][1] == d1 && c_burset[e][2] == a0 && c_burset[e][3] == a1)
      return c_burset[e][4];
  return 0;
}

// ---- DP C: general_refine_borders (src/refine.c:106-190) --------------------------------------------------
// a = p, b = t (the byte b[lb] must be readable unless PC_B_NUL_AFTER: it is what the reference reads after t).
// The reference fills two (len_p+1) x (t_win+1) matrices and then takes, per row, the minimum and its FIRST column.
// Along the wavefront the cells of one row arrive in increasing column order, so the same (minimum, first argmin)
// pair is kept per row while the sweep runs and no matrix is stored: O(len_p) scratch instead of O(len_p * t_win).
__device__ void rowmin_dp(Str rows, int nr, Str cols, int nc, uint32_t *H, uint32_t *mn, uint32_t *pos, int lane) {
  const int lo = -nr, W = nr + nc + 1;
  for (int x = lane; x < W; x += 32) H[x] = PC_INF;
  for (int i = lane; i <= nr; i += 32) { mn[i] = PC_INF; pos[i] = 0; }
  __syncwarp();
  for (int s = 0; s <= nr + nc; ++s) {
    const int i_lo = max(0, s - nc), i_hi = min(nr, s);
    for (int i = i_lo + lane; i <= i_hi; i += 32) {
      const int j = s - i, x = j - i - lo;
      uint32_t v;
      if (i == 0) v = (uint32_t)j;
      else if (j == 0) v = (uint32_t)i;
      else {
        v = H[x] + (rows.at(i - 1) == cols.at(j - 1) ? 0u : 1u);
        v = min(v, H[x + 1] + 1u);
        v = min(v, H[x - 1] + 1u);
      }
      H[x] = v;
      if (mn[i] > v) { mn[i] = v; pos[i] = (uint32_t)j; }      // strict: the first minimum of the row stays
    }
    __syncwarp();
  }
}

__device__ void op_borders(const PcDevBatch &B, WarpPool &wp, int w, uint32_t *smem, int lane) {
  JobView J = view(B, w);
  const int len_p = J.la, len_t = J.lb;
  const uint32_t max_errs = (uint32_t)J.job->p0;
  const int min_cut = J.job->p1, max_cut = J.job->p2;
  if (min_cut < 0 || min_cut > max_cut || max_cut > len_p) PC_FAIL(PC_E_ARG);
  const int t_win = (int)min((unsigned long long)len_p + max_errs, (unsigned long long)len_t);
  bool ok;
  uint32_t *H = state_mem(B, wp, smem, (size_t)len_p + t_win + 1 + 4ull * (len_p + 1), lane, ok);
  if (!ok) PC_FAIL(PC_E_POOL);
  uint32_t *mn = H + (len_p + t_win + 1), *pos = mn + 2 * (len_p + 1);
  // rows over p, columns over the first t_win chars of t (prefix side) / of reversed t (suffix side)
  rowmin_dp(Str{J.a, 1}, len_p, Str{J.b, 1}, t_win, H, mn, pos, lane);
  rowmin_dp(Str{J.a + len_p - 1, -1}, len_p, Str{J.b + len_t - 1, -1}, t_win, H, mn + len_p + 1, pos + len_p + 1, lane);
  __syncwarp();
  if (lane == 0) {
    mn[0] = 0; pos[0] = 0; mn[len_p + 1] = 0; pos[len_p + 1] = 0;      // row 0: min_pp[0] = min_sp[0] = 0 at column 0
    const uint32_t *mn_p = mn, *mn_s = mn + len_p + 1, *pos_p = pos, *pos_s = pos + len_p + 1;
    int off_p = min_cut;
    uint32_t off_t1 = pos_p[min_cut], off_t2 = pos_s[len_p - min_cut];
    uint32_t best = mn_p[min_cut] + mn_s[len_p - min_cut];
    const int nul_at = (J.job->flags & PC_B_NUL_AFTER) ? len_t : -1;
    int best_freq = burset_freq(J.b, (int)off_t1, len_t - (int)off_t2, nul_at);
    for (int i = min_cut + 1; i <= max_cut; ++i) {
      const uint32_t c = mn_p[i] + mn_s[len_p - i];
      if (best < c) continue;                                  // a worse split never wins: skip its Burset lookup
      const int freq = burset_freq(J.b, (int)pos_p[i], len_t - (int)pos_s[len_p - i], nul_at);
      if (best > c || (best == c && freq > best_freq)) {
        best = c; off_p = i; off_t1 = pos_p[i]; off_t2 = pos_s[len_p - i]; best_freq = freq;
      }
    }
    J.res[0] = PC_OK; J.res[1] = best <= max_errs; J.res[2] = off_p; J.res[3] = (int32_t)off_t1;
    J.res[4] = len_t - (int32_t)off_t2; J.res[5] = (int32_t)best;
  }
}

// ---- general_refine_borders, register-resident and 16x2-packed --------------------------------------------------
// The two matrices of one job (p against the head of t, reversed p against the reversed tail of t) have the same
// shape, so they travel as the low and the high 16-bit half of one register and share every instruction (the layout
// of k_gap.cu: a group of LANES lanes owns a job, lane k holds rows 8k+1 .. 8k+8 in registers and sweeps the columns one
// step behind lane k-1, boundary values by shuffle).  The per-row (minimum, first argmin) pair is followed in registers
// with the DPX min whose predicates say which operand won.  The split is then chosen by the whole group: candidate i
// goes to lane i mod LANES, and the reference's scan order (cost, then Burset frequency, then the smaller i;
// refine.c:161-178) is the lexicographic minimum of (cost, -freq, i).
template <int LANES>
__global__ void __launch_bounds__(128) k_borders_packed(PcDevBatch B, int tcap) {
  extern __shared__ uint32_t sh_b[];
  constexpr int G = 32 / LANES, RMAX = 8 * LANES;
  constexpr uint32_t ONE2 = 0x00010001u;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int k = lane % LANES, grp = lane / LANES;
  const int per_group = (tcap + 1) + 2 * (RMAX + 1);
  uint32_t *codes = sh_b + (size_t)(wib * G + grp) * per_group, *mnS = codes + tcap + 1, *posS = mnS + RMAX + 1;
  const int warp_id = blockIdx.x * 4 + wib, ngroups = gridDim.x * 4 * G;
  for (int q0 = warp_id * G; q0 < B.n; q0 += ngroups) {            // warp-uniform trip count
    const int q = q0 + grp;
    bool live = q < B.n;
    JobView J;
    int len_p = 0, len_t = 0, t_win = 0, min_cut = 0, max_cut = 0;
    uint32_t max_errs = 0;
    if (live) {
      J = view(B, q);
      len_p = J.la; len_t = J.lb;
      max_errs = (uint32_t)J.job->p0; min_cut = J.job->p1; max_cut = J.job->p2;
      t_win = (int)min((unsigned long long)len_p + max_errs, (unsigned long long)len_t);
      if (min_cut < 0 || min_cut > max_cut || max_cut > len_p) { if (k == 0) J.res[0] = PC_E_ARG; live = false; }
    }
    if (live)
      for (int j = k + 1; j <= t_win; j += LANES) codes[j] = (uint32_t)J.b[j - 1] | ((uint32_t)J.b[len_t - j] << 16);
    uint32_t e[8], V[8], mn[8], pos[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = k * 8 + r + 1;
      e[r] = (live && i <= len_p) ? ((uint32_t)J.a[i - 1] | ((uint32_t)J.a[len_p - i] << 16)) : 0u;
      V[r] = (uint32_t)i * ONE2;                                 // column 0: D[i][0] = i, which is also the row minimum so far
      mn[r] = V[r]; pos[r] = 0u;
    }
    int steps = live ? t_win + LANES - 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
    __syncwarp();
    uint32_t inV = (uint32_t)(8 * k) * ONE2, gcur = 0;           // (row 8k, column 0)
    for (int s = 1; s <= steps; ++s) {
      const int j = s - k;
      uint32_t dprev = inV;                                      // (row 8k, column j-1)
      inV = __shfl_up_sync(0xffffffffu, V[7], 1, LANES);         // (row 8k, column j): lane k-1 finished it one step ago
      gcur = __shfl_up_sync(0xffffffffu, gcur, 1, LANES);
      if (k == 0) { dprev = (uint32_t)(j - 1) * ONE2; inV = (uint32_t)j * ONE2; gcur = codes[min(max(s, 1), max(t_win, 1))]; }   // row 0: D[0][j] = j
      if (live && j >= 1 && j <= t_win) {
        const uint32_t j2 = (uint32_t)j * ONE2;
        uint32_t diag = dprev, up = inV;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t c = __vminu2(e[r] ^ gcur, ONE2);        // 1 where the bytes differ, per half
          uint32_t v = __vminu2(diag + c, up + ONE2);
          v = __vminu2(v, V[r] + ONE2);
          diag = V[r]; V[r] = v; up = v;
          bool ph, pl;                                           // "the old minimum is still <= v": a later equal value never replaces it
          const uint32_t m2 = pc_vibmin_u16x2(mn[r], v, ph, pl);
          const uint32_t lo = pl ? pos[r] : j2, hi = ph ? pos[r] : j2;
          pos[r] = __byte_perm(lo, hi, 0x7610);
          mn[r] = m2;
        }
      }
    }
    __syncwarp();
    if (live) {
      if (k == 0) { mnS[0] = 0u; posS[0] = 0u; }                 // row 0: minimum 0 at column 0 on both sides
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = k * 8 + r + 1;
        if (i <= len_p) { mnS[i] = mn[r]; posS[i] = pos[r]; }
      }
    }
    __syncwarp();
    unsigned long long best_key = ~0ull;
    if (live) {
      const int nul_at = (J.job->flags & PC_B_NUL_AFTER) ? len_t : -1;
      uint32_t bc = 0xffffffffu;
      for (int i = min_cut + k; i <= max_cut; i += LANES) {
        const uint32_t c = (mnS[i] & 0xffffu) + (mnS[len_p - i] >> 16);
        if (bc < c) continue;                                    // a worse split never wins: skip its Burset lookup
        const int freq = burset_freq(J.b, (int)(posS[i] & 0xffffu), len_t - (int)(posS[len_p - i] >> 16), nul_at);
        const unsigned long long key = ((unsigned long long)c << 32) | ((unsigned long long)(0xffff - freq) << 16) | (unsigned long long)i;
        if (key < best_key) { best_key = key; bc = c; }
      }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best_key, o, LANES);
      best_key = other < best_key ? other : best_key;
    }
    if (live && k == 0) {
      const int i = (int)(best_key & 0xffffu);
      const uint32_t best = (uint32_t)(best_key >> 32);
      J.res[0] = PC_OK; J.res[1] = best <= max_errs; J.res[2] = i; J.res[3] = (int32_t)(posS[i] & 0xffffu);
      J.res[4] = len_t - (int32_t)(posS[len_p - i] >> 16); J.res[5] = (int32_t)best;
    }
    __syncwarp();
  }
}

// The same sweep for jobs of any height: one warp per job, rows in chunks of 256 (32 lanes x 8 rows).  The bottom row
// of a chunk is the top boundary of the next one: lane 31 stores it column by column into a scratch row that lane 0
// reads one column ahead (in place: lane 31 trails lane 0 by 31 columns).  Column letters come straight from t (two
// bytes per step, fetched one step ahead), per-row minima go to scratch as each chunk ends.  Values must fit 16 bits
// (len_p + t_win <= 65000); anything larger is listed for the wavefront kernel.
__global__ void __launch_bounds__(128) k_borders_chunked(PcDevBatch B, uint32_t *slow_list, uint32_t *slow_count) {
  constexpr uint32_t ONE2 = 0x00010001u;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp_id = blockIdx.x * 4 + wib, nwarps = pc_active_warps(B, gridDim.x * 4);
  if (warp_id >= nwarps) return;
  WarpPool wp = pc_warp_pool(B, warp_id);
  for (int q = warp_id; q < B.n; q += nwarps) {
    wp.used = 0;
    JobView J = view(B, q);
    const int len_p = J.la, len_t = J.lb;
    const uint32_t max_errs = (uint32_t)J.job->p0;
    const int min_cut = J.job->p1, max_cut = J.job->p2;
    if (min_cut < 0 || min_cut > max_cut || max_cut > len_p) { if (lane == 0) J.res[0] = PC_E_ARG; continue; }
    const unsigned long long tw64 = min((unsigned long long)len_p + max_errs, (unsigned long long)len_t);
    if ((unsigned long long)len_p + tw64 > 65000ull) { if (lane == 0) slow_list[atomicAdd(slow_count, 1u)] = B.idx[q]; continue; }
    const int t_win = (int)tw64;
    uint32_t *rowbuf = (uint32_t *)pc_pool_alloc(B, wp, 4ull * ((unsigned long long)t_win + 2 + 2ull * (len_p + 1)), lane);
    if (!rowbuf) { if (lane == 0) J.res[0] = PC_E_POOL; continue; }
    uint32_t *mnS = rowbuf + t_win + 2, *posS = mnS + len_p + 1;
    for (int j = lane; j <= t_win; j += 32) rowbuf[j] = (uint32_t)j * ONE2;             // row 0: D[0][j] = j
    if (lane == 0) { mnS[0] = 0u; posS[0] = 0u; }
    __syncwarp();
    const int steps = t_win + 31;
    for (int base = 0; base < len_p; base += 256) {
      uint32_t e[8], V[8], mn[8], pos[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = base + lane * 8 + r + 1;
        e[r] = i <= len_p ? ((uint32_t)J.a[i - 1] | ((uint32_t)J.a[len_p - i] << 16)) : 0u;
        V[r] = (uint32_t)i * ONE2; mn[r] = V[r]; pos[r] = 0u;
      }
      uint32_t inV = (uint32_t)(base + 8 * lane) * ONE2, gcur = 0;
      // lane 0 runs one column ahead on its two inputs: the boundary row and the letters of column j
      uint32_t top_next = 0, code_next = 0;
      if (lane == 0 && t_win >= 1) { top_next = rowbuf[1]; code_next = (uint32_t)J.b[0] | ((uint32_t)J.b[len_t - 1] << 16); }
      for (int s = 1; s <= steps; ++s) {
        const int j = s - lane;
        uint32_t dprev = inV;
        inV = __shfl_up_sync(0xffffffffu, V[7], 1);
        gcur = __shfl_up_sync(0xffffffffu, gcur, 1);
        if (lane == 0) {
          inV = top_next; gcur = code_next;                        // (row base, column j), letters of column j; dprev = (row base, column j-1)
          if (s + 1 <= t_win) { top_next = rowbuf[s + 1]; code_next = (uint32_t)J.b[s] | ((uint32_t)J.b[len_t - s - 1] << 16); }
        }
        if (j >= 1 && j <= t_win) {
          const uint32_t j2 = (uint32_t)j * ONE2;
          uint32_t diag = dprev, up = inV;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const uint32_t c = __vminu2(e[r] ^ gcur, ONE2);
            uint32_t v = __vminu2(diag + c, up + ONE2);
            v = __vminu2(v, V[r] + ONE2);
            diag = V[r]; V[r] = v; up = v;
            bool ph, pl;
            const uint32_t m2 = pc_vibmin_u16x2(mn[r], v, ph, pl);
            const uint32_t lo = pl ? pos[r] : j2, hi = ph ? pos[r] : j2;
            pos[r] = __byte_perm(lo, hi, 0x7610);
            mn[r] = m2;
          }
          if (lane == 31) rowbuf[j] = V[7];                        // bottom row of this chunk = top boundary of the next
        }
      }
      if (lane == 31) rowbuf[0] = (uint32_t)(base + 256) * ONE2;   // column 0 of that boundary row
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = base + lane * 8 + r + 1;
        if (i <= len_p) { mnS[i] = mn[r]; posS[i] = pos[r]; }
      }
      __syncwarp();
    }
    __syncwarp();
    const int nul_at = (J.job->flags & PC_B_NUL_AFTER) ? len_t : -1;
    unsigned long long best_key = ~0ull;
    uint32_t bc = 0xffffffffu;
    for (int i = min_cut + lane; i <= max_cut; i += 32) {
      const uint32_t c = (mnS[i] & 0xffffu) + (mnS[len_p - i] >> 16);
      if (bc < c) continue;
      const int freq = burset_freq(J.b, (int)(posS[i] & 0xffffu), len_t - (int)(posS[len_p - i] >> 16), nul_at);
      const unsigned long long key = ((unsigned long long)c << 32) | ((unsigned long long)(0xffff - freq) << 16) | (unsigned long long)i;   // i <= len_p < 65536
      if (key < best_key) { best_key = key; bc = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best_key, o);
      best_key = other < best_key ? other : best_key;
    }
    if (lane == 0) {
      const int i = (int)(best_key & 0xffffu);
      const uint32_t best = (uint32_t)(best_key >> 32);
      J.res[0] = PC_OK; J.res[1] = best <= max_errs; J.res[2] = i; J.res[3] = (int32_t)(posS[i] & 0xffffu);
      J.res[4] = len_t - (int32_t)(posS[len_p - i] >> 16); J.res[5] = (int32_t)best;
    }
    __syncwarp();
  }
}

template <int LANES>
void launch_borders_packed(const PcDevBatch &B, int tcap, cudaStream_t s, int sm_count) {
  constexpr int G = 32 / LANES;
  const size_t sh = (size_t)4 * G * ((tcap + 1) + 2 * (8 * LANES + 1)) * sizeof(uint32_t);
  pc_smem_optin((const void *)k_borders_packed<LANES>, 100 * 1024);
  const int per_sm = pc_cached_occupancy((const void *)k_borders_packed<LANES>, 128, sh);
  const int needed = (B.n + 4 * G - 1) / (4 * G);
  int grid = needed < sm_count * per_sm ? needed : sm_count * per_sm;
  if (grid < 1) grid = 1;
  k_borders_packed<LANES><<<grid, 128, sh, s>>>(B, tcap);
  PC_COUNT_LAUNCH(1);
}

// ---- DP E: find_longest_affix (src/factorization-refinement.c:1134-1172) ----------------------------------
// Wanted: among the cells (e, g) with equal end characters and weight 2*D/(e+g) <= 0.17, the minimal weight, LAST in
// row-major order on ties.  A qualifying cell has D <= 0.085*(e+g) <= wd := floor(0.085*(el+gl)) + 1, and every cell
// whose true distance is <= wd lies within wd diagonals of the main one and is computed exactly by a DP restricted
// to that band (off-band cells can only be over-estimated, which never makes a cell qualify): the sweep is
// O((el+gl)*wd) instead of el*gl, and the reduction runs inside it, so nothing but the wavefront state is stored.
__device__ void op_affix(const PcDevBatch &B, WarpPool &wp, int w, uint32_t *smem, int lane) {
  JobView J = view(B, w);
  const int el = J.la, gl = J.lb;
  const int wd = (int)(0.085 * ((double)el + (double)gl)) + 1;
  const int lo = max(-wd, -el), hi = min(wd, gl), W = hi - lo + 1;
  bool ok;
  uint32_t *H = state_mem(B, wp, smem, (size_t)W, lane, ok);
  if (!ok) PC_FAIL(PC_E_POOL);
  for (int x = lane; x < W; x += 32) H[x] = PC_INF;
  __syncwarp();
  double best = 2.0; long long best_idx = -1;
  for (int s = 0; s <= el + gl; ++s) {
    const int i_lo = max(max(0, s - gl), ceil_half(s - hi)), i_hi = min(min(el, s), floor_half(s - lo));
    for (int i = i_lo + lane; i <= i_hi; i += 32) {
      const int j = s - i, x = j - i - lo;
      uint32_t v;
      if (i == 0) v = (uint32_t)j;
      else if (j == 0) v = (uint32_t)i;
      else {
        const bool eq = J.a[i - 1] == J.b[j - 1];
        v = H[x] + (eq ? 0u : 1u);
        v = min(v, (x + 1 < W ? H[x + 1] : PC_INF) + 1u);
        v = min(v, (x > 0 ? H[x - 1] : PC_INF) + 1u);
        if (eq) {
          const double wgt = 2.0 * ((double)v) / (double)((size_t)i + (size_t)j);
          const long long idx = (long long)(i - 1) * gl + (j - 1);
          if (wgt <= 0.17 && (wgt < best || (wgt == best && idx > best_idx) || best_idx < 0)) { best = wgt; best_idx = idx; }
        }
      }
      H[x] = v;
    }
    __syncwarp();
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
    if (oi >= 0 && (best_idx < 0 || ob < best || (ob == best && oi > best_idx))) { best = ob; best_idx = oi; }
  }
  if (lane == 0) {
    J.res[0] = PC_OK; J.res[1] = best_idx >= 0;
    J.res[2] = best_idx >= 0 ? (int)(best_idx / gl) + 1 : 0;
    J.res[3] = best_idx >= 0 ? (int)(best_idx % gl) + 1 : 0;
  }
}

// ---- compute_best_suffix_cut / compute_best_prefix_cut (src/compute-alignments.c:246-316) -----------------
__device__ void op_cut(const PcDevBatch &B, WarpPool &wp, int w, uint32_t *smem, int lane, bool prefix) {
  JobView J = view(B, w);
  const int l1 = J.la, l2 = J.lb;
  if (l1 == l2 && warp_equal(J.a, J.b, l1, lane)) {
    if (lane == 0) { J.res[0] = PC_OK; J.res[1] = 0; J.res[2] = prefix ? 0 : l1; J.res[3] = prefix ? 0 : l2; }
    return;
  }
  bool ok;
  uint32_t *H = state_mem(B, wp, smem, (size_t)l1 + l2 + 1, lane, ok);
  if (!ok) PC_FAIL(PC_E_POOL);
  const size_t Wd = (size_t)l2 + 1;
  uint32_t *M = (uint32_t *)pc_pool_alloc(B, wp, (size_t)(l1 + 1) * Wd * 4ull, lane);
  if (!M) PC_FAIL(PC_E_POOL);
  Str r = prefix ? Str{J.a + l1 - 1, -1} : Str{J.a, 1};
  Str c = prefix ? Str{J.b + l2 - 1, -1} : Str{J.b, 1};
  banded_dp<false, ST_MAT>(r, l1, c, l2, -l1, l2, H, nullptr, M, lane);
  __syncwarp();
  if (lane == 0) {
    const uint32_t corner = M[(size_t)l1 * Wd + l2];
    uint32_t mincol = corner, minrow = corner;
    int colpos = l1, rowpos = l2;
    for (int i = 0; i < l1; ++i) if (mincol >= M[i * Wd + l2]) { mincol = M[i * Wd + l2]; colpos = i; }
    for (int j = 0; j < l2; ++j) if (minrow >= M[(size_t)l1 * Wd + j]) { minrow = M[(size_t)l1 * Wd + j]; rowpos = j; }
    int c1, c2; uint32_t ed;
    if (minrow < mincol) { c1 = l1; c2 = rowpos; ed = minrow; } else { c1 = colpos; c2 = l2; ed = mincol; }
    if (prefix) { c1 = l1 - c1; c2 = l2 - c2; }
    J.res[0] = PC_OK; J.res[1] = (int32_t)ed; J.res[2] = c1; J.res[3] = c2;
  }
}

// ---- DP D: compute_gap_alignment (src/refine-intron.c:560-890) --------------------------------------------
// Three score planes L/G/R swept together on the same wavefront; one direction byte per cell:
// bits 0-1 = L dir, bit 2 = G jump, bits 3-4 = R dir (3 = jump to G).
__device__ void op_gap(const PcDevBatch &B, WarpPool &wp, int w, uint32_t *smem, int lane) {
  JobView J = view(B, w);
  const int n = J.la, m = J.lb;
  uint8_t *ops = B.var_out + J.job->out_off;
  if ((uint32_t)(n + m) > J.job->out_cap) PC_FAIL(PC_E_OUTCAP);
  const int W = n + m + 1, lo = -n;
  bool ok;
  int32_t *HL = (int32_t *)state_mem(B, wp, smem, 3 * (size_t)W, lane, ok);
  if (!ok) PC_FAIL(PC_E_POOL);
  int32_t *HG = HL + W, *HR = HG + W;
  const size_t Wd = (size_t)m + 1;
  uint8_t *dir = pc_pool_alloc(B, wp, (unsigned long long)(n + 1) * Wd, lane);
  if (!dir) PC_FAIL(PC_E_POOL);
  for (int x = lane; x < 3 * W; x += 32) HL[x] = 0;
  __syncwarp();
  for (int s = 0; s <= n + m; ++s) {
    const int i_lo = max(0, s - m), i_hi = min(n, s);
    for (int i = i_lo + lane; i <= i_hi; i += 32) {
      const int j = s - i, x = j - i - lo;
      int32_t vl = 0, vg = 0, vr = 0;
      uint8_t d = 0;
      if (i > 0 && j > 0) {
        const uint8_t ec = J.a[i - 1], gc = J.b[j - 1];
        const int sc = (ec == gc || pc_is_n(ec) || pc_is_n(gc)) ? 1 : -1;
        const int32_t Lup = HL[x + 1], Llf = HL[x - 1], Glf = HG[x - 1], Rup = HR[x + 1], Rlf = HR[x - 1];
        uint8_t dl = 0, dg = 0, dr = 0;
        vl = HL[x] + sc;
        if (vl < Lup - 1) { vl = Lup - 1; dl = 1; }
        if (vl < Llf - 1) { vl = Llf - 1; dl = 2; }
        vg = Glf;
        if (vg < Llf) { vg = Llf; dg = 1; }
        vr = HR[x] + sc;
        const int32_t hgap = (i != n) ? Rlf - 1 : Rlf;
        if (vr < hgap) { vr = hgap; dr = 2; }
        if (vr < Glf) { vr = Glf; dr = 3; }
        if (vr < Rup - 1) { vr = Rup - 1; dr = 1; }
        d = dl | (dg << 2) | (dr << 3);
      }
      HL[x] = vl; HG[x] = vg; HR[x] = vr;
      dir[(size_t)i * Wd + j] = d;
    }
    __syncwarp();
  }
  int k = 0;
  if (lane == 0) {
    const int xe = m - n - lo;
    const int32_t Le = HL[xe], Ge = HG[xe], Re = HR[xe];
    int state;
    if (Re >= Ge) state = (Re >= Le) ? 2 : 0; else state = (Ge >= Le) ? 1 : 0;
    int pos[5] = {0, 0, 0, 0, 0};
    int i = n, j = m, k_end = -1, k_start = -1;
    while (i > 0 || j > 0) {
      if (i > 0 && j > 0) {
        const uint8_t c = dir[(size_t)i * Wd + j];
        const int dd = state == 2 ? (c >> 3) & 3 : state == 1 ? (((c >> 2) & 1) ? 3 : 2) : (c & 3);
        if (dd == 0) { ops[k++] = 0; --i; --j; }
        else if (dd == 1) { ops[k++] = 1; --i; }
        else {
          if (dd == 3) {
            if (state == 2) { pos[2] = j - 1; pos[0] = i; k_end = k; } else { pos[1] = j - 1; k_start = k; }
            --state;
          }
          ops[k++] = 2; --j;
        }
      } else if (i > 0) { ops[k++] = 1; --i; }
      else { ops[k++] = 2; --j; }
    }
    if (k_end >= 0) pos[4] = k - 1 - k_end;
    if (k_start >= 0) pos[3] = k - 1 - k_start;
    J.res[0] = PC_OK; J.res[1] = k;
    for (int q = 0; q < 5; ++q) J.res[2 + q] = pos[q];
  }
  k = __shfl_sync(0xffffffffu, k, 0);
  __syncwarp();
  warp_reverse(ops, k, lane);
}

template <int OP>
__global__ void __launch_bounds__(PC_WARPS_PER_CTA * 32) k_warp_per_job(PcDevBatch B) {
  __shared__ uint32_t smem[PC_WARPS_PER_CTA][PC_SMEM_INTS_PER_WARP];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nwarps = pc_active_warps(B, gridDim.x * PC_WARPS_PER_CTA);
  if (blockIdx.x * PC_WARPS_PER_CTA + wib >= nwarps) return;      // no block-wide barrier in this kernel: a warp may leave
  // jobs are sorted heaviest first; deal them round-robin over the resident warps
  WarpPool wp = pc_warp_pool(B, blockIdx.x * PC_WARPS_PER_CTA + wib);
  const int njobs = B.n_dev ? (int)*B.n_dev : B.n;
  for (int w = blockIdx.x * PC_WARPS_PER_CTA + wib; w < njobs; w += nwarps) {
    wp.used = 0;
    if (OP == PC_OP_ALIGN) op_align(B, wp, w, smem[wib], lane);
    else if (OP == PC_OP_KBAND) op_kband(B, wp, w, smem[wib], lane);
    else if (OP == PC_OP_EDIT) op_edit(B, wp, w, smem[wib], lane);
    else if (OP == PC_OP_BORDERS) op_borders(B, wp, w, smem[wib], lane);
    else if (OP == PC_OP_GAP) op_gap(B, wp, w, smem[wib], lane);
    else if (OP == PC_OP_AFFIX) op_affix(B, wp, w, smem[wib], lane);
    else if (OP == PC_OP_SUFCUT) op_cut(B, wp, w, smem[wib], lane, false);
    else if (OP == PC_OP_PRECUT) op_cut(B, wp, w, smem[wib], lane, true);
    __syncwarp();
  }
}

template <int OP>
void launch_wpj(const PcDevBatch &B, cudaStream_t s, int sm_count) {
  const int ctas_needed = (B.n + PC_WARPS_PER_CTA - 1) / PC_WARPS_PER_CTA;
  const int resident = sm_count * 4;      // 48 KB static smem per CTA -> 4 CTAs per SM
  int grid = ctas_needed < resident ? ctas_needed : resident;
  if (B.max_warps > 0 && grid > (B.max_warps + PC_WARPS_PER_CTA - 1) / PC_WARPS_PER_CTA)
    grid = (B.max_warps + PC_WARPS_PER_CTA - 1) / PC_WARPS_PER_CTA;
  if (grid < 1) grid = 1;
  PcDevBatch C = B;
  C.slots = (B.max_warps > 0 && B.max_warps < grid * PC_WARPS_PER_CTA) ? B.max_warps : grid * PC_WARPS_PER_CTA;
  k_warp_per_job<OP><<<grid, PC_WARPS_PER_CTA * 32, 0, s>>>(C);
  PC_COUNT_LAUNCH(1);
}

}  // namespace

// cls 0/1/2: len_p <= 64 / 128 / 256 with 1 <= t_win <= PC_BORDERS_FAST_MAX_T (the generic wavefront kernel takes the rest)
void pc_launch_borders_packed(int cls, const PcDevBatch &B, int tcap, cudaStream_t s, int sm_count) {
  if (cls == 0) launch_borders_packed<8>(B, tcap, s, sm_count);
  else if (cls == 1) launch_borders_packed<16>(B, tcap, s, sm_count);
  else launch_borders_packed<32>(B, tcap, s, sm_count);
}

// BORDERS jobs outside the packed classes (taller than 256 rows, or a window above PC_BORDERS_FAST_MAX_T columns)
void pc_launch_borders_chunked(const PcDevBatch &B, uint32_t *slow_list, uint32_t *slow_count, cudaStream_t s, int sm_count) {
  const int per_sm = pc_cached_occupancy((const void *)k_borders_chunked, 128, 0);
  const int needed = (B.n + 3) / 4;
  int grid = needed < sm_count * per_sm ? needed : sm_count * per_sm;
  if (B.max_warps > 0 && grid > (B.max_warps + 3) / 4) grid = (B.max_warps + 3) / 4;
  if (grid < 1) grid = 1;
  PcDevBatch C = B;
  C.slots = (B.max_warps > 0 && B.max_warps < grid * 4) ? B.max_warps : grid * 4;
  k_borders_chunked<<<grid, 128, 0, s>>>(C, slow_list, slow_count);
  PC_COUNT_LAUNCH(1);
}

void pc_launch_dp(int op, const PcDevBatch &B, cudaStream_t s, int sm_count) {
  switch (op) {
    case PC_OP_ALIGN: launch_wpj<PC_OP_ALIGN>(B, s, sm_count); break;
    case PC_OP_KBAND: launch_wpj<PC_OP_KBAND>(B, s, sm_count); break;
    case PC_OP_EDIT: launch_wpj<PC_OP_EDIT>(B, s, sm_count); break;
    case PC_OP_BORDERS: launch_wpj<PC_OP_BORDERS>(B, s, sm_count); break;
    case PC_OP_GAP: launch_wpj<PC_OP_GAP>(B, s, sm_count); break;
    case PC_OP_AFFIX: launch_wpj<PC_OP_AFFIX>(B, s, sm_count); break;
    case PC_OP_SUFCUT: launch_wpj<PC_OP_SUFCUT>(B, s, sm_count); break;
    case PC_OP_PRECUT: launch_wpj<PC_OP_PRECUT>(B, s, sm_count); break;
    default: break;
  }
}

// ---- INT32 ALU peak micro-benchmark (roofline denominator for the DP kernels) ------------
// 8 independent add/min chains per thread; each add+min pair becomes one VIADDMNMX on the ALU pipe (checked
// with cuobjdump -sass).  Returns executed ALU lane-INSTRUCTIONS; the host divides by the CUDA-event time.
__global__ void __launch_bounds__(256) k_int_peak(int iters, int a, int b, int *sink) {
  int v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
#define STEP(v) asm volatile("add.s32 %0, %0, %1;\n\tmin.s32 %0, %0, %2;" : "+r"(v) : "r"(a), "r"(b));
    STEP(v0) STEP(v1) STEP(v2) STEP(v3) STEP(v4) STEP(v5) STEP(v6) STEP(v7)
#undef STEP
  }
  if ((v0 ^ v1 ^ v2 ^ v3 ^ v4 ^ v5 ^ v6 ^ v7) == 0x7fffffff) *sink = v0;
}

double pc_int_peak_run(cudaStream_t s, int sm_count, float *ms_out) {
  int *sink; cudaMalloc(&sink, 4);
  const int iters = 1 << 14, grid = sm_count * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_int_peak<<<grid, 256, 0, s>>>(iters, 3, 1 << 30, sink);          // warm-up
  cudaEventRecord(e0, s);
  k_int_peak<<<grid, 256, 0, s>>>(iters, 3, 1 << 30, sink);
  cudaEventRecord(e1, s);
  cudaEventSynchronize(e1);
  PC_COUNT_LAUNCH(2);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
  if (ms_out) *ms_out = ms;
  return (double)grid * 256.0 * iters * 8.0;    // ptxas fuses each add+min pair into ONE VIADDMNMX: 8 ALU instructions / iteration
}
