// k_seed.cu — genome k-mer index, maximal-pairing discovery (seed-and-extend) and the genome LCS scan.
//
// Index (replaces the Ukkonen suffix tree, reference stree_src/lst_stree.c:816 + src/aug_suffix_tree.c:151-264):
// a 64-bit hash of every `word`-byte window of the genome, radix-sorted with the window start as payload, so
// that all occurrences of a word are one contiguous, position-ascending run.  Words are hashed as BYTES
// (the reference matches literal bytes: 'N' equals 'N', case matters, the mask characters '*' and '#'
// never occur in a genome), and every hit is re-verified by the byte-wise extension, so hash collisions cost
// time, never correctness.
#include "pc_device.cuh"
#include "meg_core.h"
#include <cub/device/device_radix_sort.cuh>
#include <climits>

namespace {

__device__ __forceinline__ unsigned long long hash_word(const uint8_t *p, int word) {
  unsigned long long h = 0xcbf29ce484222325ull;
  for (int i = 0; i < word; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
  return h;
}

__global__ void k_hash_windows(const uint8_t *g, uint32_t n_win, int word, unsigned long long *keys, uint32_t *pos) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_win; t += gridDim.x * blockDim.x) {
    keys[t] = hash_word(g + t, word);
    pos[t] = t;
  }
}

// bucket b = top `bits` bits of the hash; bstart[b] = first sorted entry of bucket b (bstart[nb] = n): two loads
// replace a 17-step dependent binary search
__global__ void k_bucket_starts(const unsigned long long *keys, uint32_t n, int shift, uint32_t nb, uint32_t *bstart) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
    const uint32_t b_prev = i == 0 ? 0u : (uint32_t)(keys[i - 1] >> shift) + 1u;
    const uint32_t b_cur = i == n ? nb : (uint32_t)(keys[i] >> shift);
    for (uint32_t b = b_prev; b <= b_cur; ++b) bstart[b] = i;     // buckets (b_prev-1, b_cur] start at i
  }
}

__device__ __forceinline__ int lcp(const uint8_t *P, int n, int p, const uint8_t *T, uint32_t G, uint32_t t) {
  const int lim = min(n - p, (int)(G - t));
  const uint8_t *a = P + p, *b = T + t;
  int l = 0;
  // eight independent byte loads per round instead of a chain of dependent ones
  while (l + 4 <= lim) {
    const uint8_t a0 = a[l], a1 = a[l + 1], a2 = a[l + 2], a3 = a[l + 3];
    const uint8_t b0 = b[l], b1 = b[l + 1], b2 = b[l + 2], b3 = b[l + 3];
    if (a0 != b0) return l;
    if (a1 != b1) return l + 1;
    if (a2 != b2) return l + 2;
    if (a3 != b3) return l + 3;
    l += 4;
  }
  while (l < lim && a[l] == b[l]) ++l;
  return l;
}

// exclusive scan of v[0..n) in place by one warp; returns the total
__device__ int warp_exscan(int *v, int n, int lane) {
  int carry = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const int x = i < n ? v[i] : 0;
    int inc = x;
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (i < n) v[i] = carry + inc - x;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  return carry;
}

// ---- build_vertex_set (src/max-emb-graph.c:217-380), one warp per EST ------------------------------------
// Spec = SURVEY.md Appendix A (validated against the reference on 4 259 MEG builds):
//  per EST position p: the occurrences t of P[p..p+word) that are left-maximal (t==0 or p==0 or
//  T[t-1]!=P[p-1]) and extend to l = LCP >= mfl; D = max l; keep l >= max(floor(D*rate), mfl);
//  ascending t; filter A inside one p; filter B against the list of p-1; emit (p,t,l) by ascending p.
struct TL { int t, l; };

#define SEED_TILE 2048      /* EST bytes staged in shared memory per warp (longer ESTs are read through L1) */
// Returns the number of triples, written to the job's output region — or, for a PC_SEED_BUILD_MEG job, to the warp's scratch
// slot (*tri_out); -1 when the job is already answered (status written).
__device__ int seed_one(const PcDevBatch &B, WarpPool &wp, int w, int lane, uint8_t *tile, int32_t **tri_out) {
  const uint32_t ji = B.idx[w];
  const pc_job *job = B.jobs + ji;
  const bool meg = job->p1 == PC_SEED_BUILD_MEG;
  int32_t *res = B.res + (size_t)ji * PC_RES_INTS;
  const uint8_t *P = B.arena + job->a_off;
  const int n = (int)job->a_len, mfl = job->p0, word = B.ix_word;
  if (n <= SEED_TILE) {                      // the EST is read ~ word + LCP times per position: keep it on chip
    for (int i = lane; i < n; i += 32) tile[i] = P[i];
    __syncwarp();
    P = tile;
  }
  const uint8_t *T = B.genome;
  const uint32_t G = B.genome_len;
  int32_t *out = (int32_t *)(B.var_out + job->out_off);
  *tri_out = out;
  if (mfl < word) { if (lane == 0) res[0] = PC_E_ARG; return -1; }
  const int np = n - word + 1;               // positions that can start a word
  if (np <= 0 || G < (uint32_t)word) { if (!meg && lane == 0) { res[0] = PC_OK; res[1] = 0; } return meg ? 0 : -1; }
  // per-position arrays: bucket start, bucket end, D / offsets, counts
  int *arr = (int *)pc_pool_alloc(B, wp, 10ull * np * sizeof(int), lane);
  if (!arr) { if (lane == 0) res[0] = PC_E_POOL; return -1; }
  int *b_lo = arr, *b_hi = arr + np, *offs = arr + 2 * np, *cnt = arr + 3 * np, *thr_a = arr + 4 * np, *ncand = arr + 5 * np;
  TL *first2 = (TL *)(arr + 6 * np);           // the first two candidates of every position: S2 rarely has to extend again
  // S1: bucket, D(p), number of candidates >= mfl
  for (int p = lane; p < np; p += 32) {
    const unsigned long long h = hash_word(P + p, word);
    const uint32_t bk = (uint32_t)(h >> B.ix_shift);
    uint32_t k = B.ix_bstart[bk];
    const uint32_t kend = B.ix_bstart[bk + 1];
    while (k < kend && B.ix_keys[k] != h) ++k;               // entries of one word are contiguous inside the bucket
    const uint32_t k0 = k;
    int c = 0, D = 0;
    const uint8_t prev = p > 0 ? P[p - 1] : 0;
    for (; k < kend && B.ix_keys[k] == h; ++k) {
      const uint32_t t = B.ix_pos[k];
      if (p > 0 && t > 0 && T[t - 1] == prev) continue;
      const int l = lcp(P, n, p, T, G, t);
      if (l >= mfl) { if (c < 2) { first2[2 * p + c].t = (int)t; first2[2 * p + c].l = l; } ++c; D = max(D, l); }
    }
    b_lo[p] = (int)k0; b_hi[p] = (int)k;
    int thr = (int)(size_t)((double)D * B.depth_rate);
    thr_a[p] = max(thr, mfl);
    offs[p] = c; ncand[p] = c;
  }
  __syncwarp();
  const int total = warp_exscan(offs, np, lane);
  __syncwarp();
  TL *cand = nullptr; uint8_t *keep = nullptr;
  if (total > 0) {
    cand = (TL *)pc_pool_alloc(B, wp, (unsigned long long)total * (sizeof(TL) + 1), lane);
    if (!cand) { if (lane == 0) res[0] = PC_E_POOL; return -1; }
    keep = (uint8_t *)(cand + total);
  }
  // S2: emit candidates >= thr (ascending t), filter A in place
  for (int p = lane; p < np; p += 32) {
    TL *v = cand + offs[p];
    const int thr = thr_a[p];
    const uint8_t prev = p > 0 ? P[p - 1] : 0;
    int c = 0;
    if (ncand[p] == 0) { cnt[p] = 0; continue; }
    if (ncand[p] <= 2) {
      for (int k = 0; k < ncand[p]; ++k) if (first2[2 * p + k].l >= thr) v[c++] = first2[2 * p + k];
    } else
    for (int k = b_lo[p]; k < b_hi[p]; ++k) {
      const uint32_t t = B.ix_pos[k];
      if (p > 0 && t > 0 && T[t - 1] == prev) continue;
      const int l = lcp(P, n, p, T, G, t);
      if (l >= thr) { v[c].t = (int)t; v[c].l = l; ++c; }
    }
    // filter A, judged against the UNfiltered list.  The list is strictly ascending in t, so "an earlier entry ends
    // at or after this one" is a running maximum of t+l, and "t == earlier t + 1" can only be the direct predecessor.
    int q = 0, max_end = INT_MIN, pt = INT_MIN, pl = -1;
    for (int j = 0; j < c; ++j) {
      const int t = v[j].t, l = v[j].l;
      const bool drop = (j > 0 && t + l <= max_end) || (t == pt + 1 && l == pl);
      max_end = max(max_end, t + l); pt = t; pl = l;
      if (!drop) { v[q].t = t; v[q].l = l; ++q; }
    }
    cnt[p] = q;
  }
  __syncwarp();
  // S3: filter B — list(p) against the post-A list(p-1); only flags are written, lists stay intact
  int *outc = b_lo;                             // buckets are no longer needed
  for (int p = lane; p < np; p += 32) {
    const TL *v = cand + offs[p];
    uint8_t *kp = keep + offs[p];
    int q = 0, y = 0;
    const TL *u = p > 0 ? cand + offs[p - 1] : nullptr;
    const int nu = p > 0 ? cnt[p - 1] : 0;
    for (int x = 0; x < cnt[p]; ++x) {                       // both lists ascend in t: one merge pass
      while (y < nu && u[y].t < v[x].t) ++y;
      const bool drop = y < nu && u[y].t == v[x].t && u[y].l >= v[x].l;
      kp[x] = !drop; q += !drop;
    }
    outc[p] = q;
  }
  __syncwarp();
  const int n_out = warp_exscan(outc, np, lane);
  __syncwarp();
  if (meg) {
    out = (int32_t *)pc_pool_alloc(B, wp, 12ull * (unsigned long long)n_out + 4ull, lane);
    if (!out) { if (lane == 0) res[0] = PC_E_POOL; return -1; }
    *tri_out = out;
  } else if ((uint32_t)n_out > job->out_cap) { if (lane == 0) { res[0] = PC_E_OUTCAP; res[1] = n_out; } return -1; }
  for (int p = lane; p < np; p += 32) {
    const TL *v = cand + offs[p];
    const uint8_t *kp = keep + offs[p];
    int o = outc[p];
    for (int x = 0; x < cnt[p]; ++x)
      if (kp[x]) { out[3 * o] = p; out[3 * o + 1] = v[x].t; out[3 * o + 2] = v[x].l; ++o; }
  }
  if (!meg && lane == 0) { res[0] = PC_OK; res[1] = n_out; }
  return n_out;
}

// ---- the rest of build_meg on the vertex set just found (PC_SEED_BUILD_MEG; meg_core.h) -----------------------------
// The triples move to the bottom of the warp's scratch slot (everything else in it is dead by now) and the graph is carved
// out of what follows.  Edge rules, list orders and the simplification passes are sequential and order-dependent (the bytes
// of megs.txt follow the list orders): one lane walks them; graphs are a few dozen vertices, and the other warps of the SM
// run meanwhile.
__device__ void meg_one(const PcDevBatch &B, WarpPool &wp, int w, int lane, int32_t *tri, int ntri) {
  const uint32_t ji = B.idx[w];
  const pc_job *job = B.jobs + ji;
  int32_t *res = B.res + (size_t)ji * PC_RES_INTS;
  int32_t *base = (int32_t *)wp.base;
  const int nw = 3 * ntri;
  __syncwarp();
  if (tri != base)
    for (int r0 = 0; r0 < nw; r0 += 32) {              // downwards, a whole round read before it is written
      const int i = r0 + lane;
      const int32_t v = i < nw ? tri[i] : 0;
      __syncwarp();
      if (i < nw) base[i] = v;
      __syncwarp();
    }
  __syncwarp();
  // Large vertex sets (mRNAs of several kbp: hundreds of pairings, build_edge_set alone is quadratic) would keep this one
  // lane — and the batch that waits for it — busy for tens of milliseconds; a host core does the same walk in well under
  // one.  Above the caller's limit (p2 > 0) the job answers with the vertex set only (res[3] = 1): the whole warp copies
  // the triples out and the caller runs the very same meg_core.h on them.
  if (job->p2 > 0 && ntri > job->p2) {
    if ((uint32_t)ntri > job->out_cap) { if (lane == 0) { res[0] = PC_E_OUTCAP; res[1] = ntri; } }
    else {
      int32_t *out = (int32_t *)(B.var_out + job->out_off);
      for (int i = lane; i < nw; i += 32) out[i] = base[i];
      if (lane == 0) { res[0] = PC_OK; res[1] = ntri; res[3] = PC_SEED_VERTEX_SET_ONLY; }
    }
    __syncwarp();
    return;
  }
  if (lane == 0) {
    res[3] = 0;
    const unsigned long long tri_bytes = (12ull * (unsigned long long)ntri + 255ull) & ~255ull;
    pc_meg_cfg cfg;
    const uint8_t *cb = B.arena + job->b_off;
    for (int i = 0; i < (int)sizeof(pc_meg_cfg); ++i) ((uint8_t *)&cfg)[i] = cb[i];
    mg_graph g;
    g.err = MG_E_SCRATCH;
    int retry = 0;
    if (tri_bytes + 1024ull <= wp.size) {
      mg_init(&g, (int *)(wp.base + tri_bytes), (long long)((wp.size - tri_bytes) / 4ull), base, ntri);
      if (!g.err) retry = mg_build(&g, (int)job->a_len, job->p0, &cfg);
    }
    if (g.err == MG_E_SCRATCH) { atomicMax(B.pool_need, 2ull * wp.size + 4096ull); res[0] = PC_E_POOL; }
    else if (g.err) res[0] = PC_E_RANGE;                  // a cyclic graph: cannot happen (edges ascend in p); the host reports it
    else {
      const long long units = (mg_record_words(&g) + 2) / 3;
      if (units > (long long)job->out_cap) { res[0] = PC_E_OUTCAP; res[1] = (int32_t)units; }
      else { mg_write_record(&g, retry, (int32_t *)(B.var_out + job->out_off)); res[0] = PC_OK; res[1] = (int32_t)units; }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(128) k_seed(PcDevBatch B) {
  __shared__ uint8_t tiles[4][SEED_TILE];
  const int lane = threadIdx.x & 31;
  const int nwarps = pc_active_warps(B, gridDim.x * (blockDim.x >> 5));
  if (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) >= nwarps) return;
  WarpPool wp = pc_warp_pool(B, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < B.n; w += nwarps) {
    wp.used = 0;
    int32_t *tri = nullptr;
    const int ntri = seed_one(B, wp, w, lane, tiles[threadIdx.x >> 5], &tri);
    if (ntri >= 0 && B.jobs[B.idx[w]].p1 == PC_SEED_BUILD_MEG) meg_one(B, wp, w, lane, tri, ntri);
    __syncwarp();
  }
}

// ---- find_longest_common_factor_dp (src/factorization-refinement.c:255-315) -------------------------------
// s1 = b (long, usually a genome prefix), s2 = a (short EST piece).  A common run ending at (i1,i2) lives on
// diagonal i1-i2; one thread walks one diagonal, a block stages its slice of s1 in shared memory.  The
// reference keeps the FIRST strictly longer run in (i1 outer, i2 inner) order = max len, then min i1, then
// min i2: packed so that one 64-bit atomicMax per block picks it.
#define LCS_TPB 256                 /* threads (= diagonals) of one block of the locate pass */
#define LCS_TILE PC_LCS_TPB          /* diagonals of one tile of the length pass (one warp, 32 per lane) */
#define LCS_MAX_S2 PC_LCS_MAX_S2

__device__ __forceinline__ unsigned long long lcs_key(int len, uint32_t i1, int i2) {
  return ((unsigned long long)len << 48) | ((unsigned long long)(0xffffffffu - i1) << 16) |
         (unsigned long long)(0xffffu - (uint32_t)i2);
}

// Bit-parallel form for s2 of at most 64 bytes (the est-fact callers pass <= 46): the cells of one diagonal become one
// 64-bit mask = (bytes equal) | (s1 byte is N) | (s2 byte is N), built four cells per step with packed byte compares;
// the longest run of ones and its first position come from x &= x << 1.
__device__ __forceinline__ uint32_t eq_bytes(uint32_t a, uint32_t b) {         // 0x80 in every byte where a == b
  const uint32_t x = a ^ b;
  return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}

__device__ __forceinline__ void lcs_block_bits(const uint8_t *s1, long long l1, const uint8_t *s2, int l2, long long d0,
                                               uint8_t *sh, unsigned long long *best_w) {
  // shared: t1 words (slice of s1, zero outside [0,l1)), N-flag bits of the slice, s2 words
  uint32_t *t1w = reinterpret_cast<uint32_t *>(sh);                 // (LCS_TPB + 64 + 8) bytes
  uint32_t *n1 = t1w + (LCS_TPB + 64 + 8) / 4;                       // (LCS_TPB + 64) / 32 + 1 words
  uint32_t *s2w = n1 + (LCS_TPB + 64) / 32 + 2;                      // 16 words of s2 + 2 words of its N mask
  uint8_t *t1 = reinterpret_cast<uint8_t *>(t1w);
  const int tid = threadIdx.x, lane = tid & 31;
  for (int base = 0; base < 2 * LCS_TPB; base += LCS_TPB) {           // every thread runs both rounds: the ballots need whole warps
    const int i = base + tid;
    const long long g = d0 + i;
    const uint8_t c = (g >= 0 && g < l1 && i < LCS_TPB + 64) ? s1[g] : 0;
    if (i < LCS_TPB + 64 + 8) t1[i] = c;
    const uint32_t nb = __ballot_sync(0xffffffffu, pc_is_n(c));
    if (lane == 0 && i < LCS_TPB + 64 + 64) n1[i >> 5] = nb;         // words 10, 11 come out zero
  }
  if (tid < 16) {
    uint32_t w = 0;
    for (int b = 0; b < 4; ++b) { const int i = tid * 4 + b; w |= (uint32_t)(i < l2 ? s2[i] : 0xffu) << (8 * b); }   // 0xff never equals the 0 padding
    s2w[tid] = w;
  }
  if (tid < 32) {                                                    // N positions of s2 as a 64-bit mask (two ballots of warp 0)
    const uint32_t f0 = __ballot_sync(0xffffffffu, tid < l2 && pc_is_n(s2[tid]));
    const uint32_t f1 = __ballot_sync(0xffffffffu, tid + 32 < l2 && pc_is_n(s2[tid + 32]));
    if (tid == 0) { s2w[16] = f0; s2w[17] = f1; }
  }
  __syncthreads();
  const long long d = d0 + tid;
  unsigned long long key = 0;
  if (d <= l1 - 1) {
    const unsigned long long n2 = (unsigned long long)s2w[16] | ((unsigned long long)s2w[17] << 32);
    const int r8 = (tid & 3) * 8, wbase = tid >> 2;
    unsigned long long eq = 0;
    uint32_t lo = t1w[wbase];
    const int nw = (l2 + 3) >> 2;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < nw) {
        const uint32_t hi = t1w[wbase + k + 1];
        const uint32_t w1 = __funnelshift_r(lo, hi, r8);
        lo = hi;
        const uint32_t f = eq_bytes(w1, s2w[k]);
        eq |= (unsigned long long)((((f >> 7) * 0x00204081u) >> 21) & 15u) << (4 * k);
      }
    }
    // N flags of s1 along the diagonal: 64 bits of the slice's flag array starting at bit tid
    const int wq = tid >> 5, bq = tid & 31;
    const uint32_t a0 = n1[wq], a1 = n1[wq + 1], a2 = n1[wq + 2];
    const unsigned long long nfl = (unsigned long long)__funnelshift_r(a0, a1, bq) | ((unsigned long long)__funnelshift_r(a1, a2, bq) << 32);
    // valid cells: 0 <= d + i2 < l1 and i2 < l2
    const int v_lo = d < 0 ? (int)-d : 0;
    const long long room = l1 - d;
    const int v_hi = (int)(room < l2 ? room : l2);                 // exclusive
    unsigned long long valid = v_hi >= 64 ? ~0ull : ((1ull << v_hi) - 1ull);
    valid = v_lo >= 64 ? 0ull : (valid & (~0ull << v_lo));
    unsigned long long x = (eq | nfl | n2) & valid, prev = 0;
    int len = 0;
    while (x) { prev = x; x &= x << 1; ++len; }
    if (len > 0) {
      const int e = __ffsll((long long)prev) - 1;                 // END of the first longest run on this diagonal
      key = lcs_key(len, (uint32_t)(d + e), e);
    }
  }
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o); key = max(key, k2); }
  if (lane == 0 && key) atomicMax(best_w, key);
}

__device__ __forceinline__ void lcs_block_generic(const uint8_t *s1, long long l1, const uint8_t *s2, int l2, long long d0,
                                                  uint8_t *sh, unsigned long long *best_w) {
  uint8_t *t1 = sh;                 // s1[d0 .. d0 + LCS_TPB + l2 - 1)
  uint8_t *t2 = sh + LCS_TPB + l2;  // s2
  for (int i = threadIdx.x; i < LCS_TPB + l2 - 1; i += LCS_TPB) {
    const long long g = d0 + i;
    t1[i] = (g >= 0 && g < l1) ? s1[g] : 0;
  }
  for (int i = threadIdx.x; i < l2; i += LCS_TPB) t2[i] = s2[i];
  __syncthreads();
  const long long d = d0 + threadIdx.x;
  unsigned long long key = 0;
  if (d <= l1 - 1) {
    int run = 0, bl = 0, bi2 = 0;
    const int i2_lo = d < 0 ? (int)-d : 0;
    for (int i2 = i2_lo; i2 < l2; ++i2) {
      const long long i1 = d + i2;
      if (i1 >= l1) break;
      const uint8_t c1 = t1[threadIdx.x + i2], c2 = t2[i2];
      run = (c1 == c2 || pc_is_n(c1) || pc_is_n(c2)) ? run + 1 : 0;
      if (run > bl) { bl = run; bi2 = i2; }
    }
    if (bl > 0) key = lcs_key(bl, (uint32_t)(d + bi2), bi2);
  }
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o); key = max(key, k2); }
  if ((threadIdx.x & 31) == 0 && key) atomicMax(best_w, key);
}

// ---- the scan in two passes ----------------------------------------------------------------------------------------
// Pass 1 (k_lcs_len) only asks HOW LONG the longest common run of every tile is, with the DP transposed: a lane owns 32
// consecutive diagonals as the 32 bits of a word, so one AND serves 32 cells.  The bytes of s1 are read as nine bit planes
// (the eight bits of the byte + "is N"): for the genome they are built once per upload (pc_build_planes) and a lane just
// loads its three words per plane; for strings in the arena the warp builds them with ballots.  For column i2 the cells
// of the lane's diagonals are the plane bits shifted by i2, compared with the (uniform) byte s2[i2]; the 32-bit equality
// words of all columns go to shared memory, and "runs of length >= k+1 ending here" = R_k[i2] & R_k[i2-1] is applied until
// nothing is left: the number of rounds is the longest run.  Pass 2 (k_lcs_pick) lists the tiles that reach their job's
// maximum; pass 3 (k_lcs_locate) runs the position-exact per-diagonal form above on those few tiles only.  The reference's
// answer — first strictly longer run in (i1 outer, i2 inner) order — is max length, then min i1, then min i2, whatever
// the order of evaluation, so the split changes nothing.  (One thread per diagonal spent ~360 instructions on ~20 cells.)
constexpr int LCS_WPB = 4;                      // warps per block of the length pass
constexpr int LCS_ROWS = LCS_TILE / 32 + 3;     // plane words a tile needs: 1024 diagonals + up to 63 more columns
struct LcsWarpMem {
  uint32_t planes[9][LCS_ROWS + 1];
  uint32_t R[64][32];
  uint32_t s2b[64];
};

__device__ __forceinline__ uint32_t range_mask(long long lo, long long hi, long long base) {   // bits b with lo <= base + b < hi
  const long long a = lo - base, e = hi - base;
  if (e <= 0 || a >= 32) return 0u;
  const uint32_t m_hi = e >= 32 ? 0xffffffffu : ((1u << (int)e) - 1u);
  const uint32_t m_lo = a <= 0 ? 0xffffffffu : (0xffffffffu << (int)a);
  return m_hi & m_lo;
}

__global__ void __launch_bounds__(LCS_WPB * 32) k_lcs_len(PcDevBatch B, unsigned long long *best, const uint32_t *blk_prefix, uint32_t total_tiles,
                                                          uint32_t *tile_len) {
  __shared__ LcsWarpMem smem[LCS_WPB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  LcsWarpMem &M = smem[wib];
  const uint32_t gw = blockIdx.x * LCS_WPB + wib, nw = gridDim.x * LCS_WPB;
  const uint32_t t0 = (uint32_t)(((unsigned long long)total_tiles * gw) / nw), t1 = (uint32_t)(((unsigned long long)total_tiles * (gw + 1)) / nw);
  if (t0 >= t1) return;
  int w;
  {
    int lo = 0, hi = B.n - 1;                             // last job whose first tile is <= t0 (every lane: same loads, broadcast)
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (blk_prefix[mid] <= t0) lo = mid; else hi = mid - 1; }
    w = lo;
  }
  int cur = -1, l2 = 0;
  const uint8_t *s1 = nullptr, *s2 = nullptr;
  long long l1 = 0, goff = -1;                            // goff >= 0: s1 lies in the genome at this offset (planes precomputed)
  uint32_t first = 0;
  for (uint32_t t = t0; t < t1; ++t) {
    while (w + 1 < B.n && blk_prefix[w + 1] <= t) ++w;
    if (w != cur) {
      cur = w;
      const pc_job *job = B.jobs + B.idx[w];
      s2 = B.arena + job->a_off; l2 = (int)job->a_len;
      const bool ing = (job->flags & PC_B_IN_GENOME) != 0;
      s1 = (ing ? B.genome : B.arena) + job->b_off; l1 = job->b_len;
      goff = (ing && B.gplanes) ? (long long)job->b_off : -1;
      first = blk_prefix[w];
      __syncwarp();
      if (l2 <= 64) for (int i = lane; i < l2; i += 32) M.s2b[i] = s2[i];
      __syncwarp();
    }
    if (l2 > 64 || l2 <= 0) { if (lane == 0) tile_len[t] = 0xffffffffu; continue; }      // the locate pass handles these itself
    const long long tile_d0 = (long long)(t - first) * LCS_TILE - (l2 - 1);                // first diagonal of the tile
    if (tile_d0 > l1 - 1) { if (lane == 0) tile_len[t] = 0; continue; }
    const long long p0 = tile_d0 + 32 * lane;                                              // s1 position of this lane's bit 0 at column 0
    uint32_t X[9][3];
    if (goff >= 0) {
      const long long g0 = goff + p0;                                                      // absolute genome position (may be negative by < 64)
      const long long wi = g0 >> 5;                                                        // arithmetic shift: floor
      const int bit = (int)(g0 & 31);
#pragma unroll
      for (int p = 0; p < 9; ++p) {
        const uint32_t *pl = B.gplanes + (size_t)p * B.gplane_words;
        uint32_t W[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const long long x = wi + q; W[q] = (x >= 0 && x < (long long)B.gplane_words) ? pl[x] : 0u; }
#pragma unroll
        for (int q = 0; q < 3; ++q) X[p][q] = __funnelshift_r(W[q], W[q + 1], bit);
      }
    } else {
      __syncwarp();
      for (int r = 0; r < LCS_ROWS; ++r) {                                                 // 32 positions per round, one ballot per plane
        const long long pos = tile_d0 + 32 * r + lane;
        const uint8_t c = (pos >= 0 && pos < l1) ? s1[pos] : 0;
        uint32_t bal[9];
#pragma unroll
        for (int p = 0; p < 8; ++p) bal[p] = __ballot_sync(0xffffffffu, (c >> p) & 1u);
        bal[8] = __ballot_sync(0xffffffffu, pc_is_n(c));
        if (lane < 9) {
          uint32_t v = bal[0];
#pragma unroll
          for (int p = 1; p < 9; ++p) v = lane == p ? bal[p] : v;
          M.planes[lane][r] = v;
        }
      }
      __syncwarp();
#pragma unroll
      for (int p = 0; p < 9; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) X[p][q] = M.planes[p][lane + q];
    }
    uint32_t V[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) V[q] = range_mask(0, l1, p0 + 32 * q);                     // positions that exist in s1
    // equality words of all columns
    uint32_t any = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {                 // columns 0..31 read words 0 and 1 of the window, columns 32..63 words 1 and 2
      const int i2_end = min(l2, 32 * (half + 1));
      for (int i2 = 32 * half; i2 < i2_end; ++i2) {
        const uint32_t sym = M.s2b[i2];
        const int sh = i2 & 31;
        uint32_t diff = 0;
#pragma unroll
        for (int p = 0; p < 8; ++p) diff |= __funnelshift_r(X[p][half], X[p][half + 1], sh) ^ (0u - ((sym >> p) & 1u));
        const uint32_t an = __funnelshift_r(X[8][half], X[8][half + 1], sh), av = __funnelshift_r(V[half], V[half + 1], sh);
        const uint32_t e = (~diff | an | (pc_is_n((uint8_t)sym) ? 0xffffffffu : 0u)) & av;
        M.R[i2][lane] = e;
        any |= e;
      }
    }
    // longest run: R_{k+1}[i2] = R_k[i2] & R_k[i2-1]
    int len = any ? 1 : 0;
    unsigned alive = __ballot_sync(0xffffffffu, any != 0);
    for (int k = 1; alive && k < l2; ++k) {
      uint32_t nz = 0, up = M.R[l2 - 1][lane];
      for (int i2 = l2 - 1; i2 >= k; --i2) {
        const uint32_t dn = M.R[i2 - 1][lane];
        const uint32_t v = up & dn;
        M.R[i2][lane] = v;
        nz |= v;
        up = dn;
      }
      if (nz) len = k + 1;
      alive = __ballot_sync(0xffffffffu, nz != 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) {
      tile_len[t] = (uint32_t)len;
      if (len) atomicMax(best + w, (unsigned long long)len << 48);                           // positions come from the locate pass
    }
  }
}

// the tiles that reach their job's longest run (and the tiles the length pass left alone)
__global__ void __launch_bounds__(256) k_lcs_pick(const unsigned long long *best, const uint32_t *blk_prefix, int n, uint32_t total_tiles,
                                                  const uint32_t *tile_len, uint32_t *list, uint32_t *count) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total_tiles; t += gridDim.x * blockDim.x) {
    const uint32_t tl = tile_len[t];
    if (tl == 0) continue;
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (blk_prefix[mid] <= t) lo = mid; else hi = mid - 1; }
    while (lo + 1 < n && blk_prefix[lo + 1] <= t) ++lo;
    if (tl == 0xffffffffu || tl == (uint32_t)(best[lo] >> 48)) list[atomicAdd(count, 1u)] = t;
  }
}

// position-exact pass over the listed tiles: LCS_TILE / LCS_TPB blocks of the per-diagonal form each
__global__ void __launch_bounds__(LCS_TPB) k_lcs_locate(PcDevBatch B, unsigned long long *best, const uint32_t *blk_prefix, const uint32_t *list,
                                                        const uint32_t *count) {
  extern __shared__ __align__(16) uint8_t sh[];
  constexpr int SUBS = LCS_TILE / LCS_TPB;
  const uint32_t n_items = *count * SUBS;
  for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    const uint32_t t = list[it / SUBS], sub = it % SUBS;
    int lo = 0, hi = B.n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (blk_prefix[mid] <= t) lo = mid; else hi = mid - 1; }
    while (lo + 1 < B.n && blk_prefix[lo + 1] <= t) ++lo;
    const int w = lo;
    const pc_job *job = B.jobs + B.idx[w];
    const uint8_t *s2 = B.arena + job->a_off;
    const int l2 = (int)job->a_len;
    const uint8_t *s1 = ((job->flags & PC_B_IN_GENOME) ? B.genome : B.arena) + job->b_off;
    const long long l1 = job->b_len;
    const long long d0 = (long long)(t - blk_prefix[w]) * LCS_TILE + (long long)sub * LCS_TPB - (l2 - 1);
    if (l2 <= LCS_MAX_S2 && l2 > 0 && d0 <= l1 - 1) {
      if (l2 <= 64) lcs_block_bits(s1, l1, s2, l2, d0, sh, best + w);
      else lcs_block_generic(s1, l1, s2, l2, d0, sh, best + w);
    }
    __syncthreads();
  }
}

// nine bit planes of the genome (bit 0..7 of every byte, "is N"), built once per upload: plane p = words [p * nwords, (p + 1) * nwords)
__global__ void __launch_bounds__(256) k_genome_planes(const uint8_t *g, uint32_t len, uint32_t nwords, uint32_t *planes) {
  const int lane = threadIdx.x & 31;
  for (uint32_t wd = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wd < nwords; wd += (gridDim.x * blockDim.x) >> 5) {
    const uint32_t pos = wd * 32 + lane;
    const uint8_t c = pos < len ? g[pos] : 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) { const uint32_t b = __ballot_sync(0xffffffffu, (c >> p) & 1u); if (lane == p) planes[(size_t)p * nwords + wd] = b; }
    const uint32_t bn = __ballot_sync(0xffffffffu, pc_is_n(c));
    if (lane == 8) planes[(size_t)8 * nwords + wd] = bn;
  }
}

__global__ void k_lcs_finish(PcDevBatch B, const unsigned long long *best) {
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < B.n; w += gridDim.x * blockDim.x) {
    const uint32_t ji = B.idx[w];
    int32_t *res = B.res + (size_t)ji * PC_RES_INTS;
    if (B.jobs[ji].a_len > LCS_MAX_S2) { res[0] = PC_E_RANGE; continue; }
    const unsigned long long key = best[w];
    const int len = (int)(key >> 48);
    const uint32_t i1 = 0xffffffffu - (uint32_t)((key >> 16) & 0xffffffffu);
    const int i2 = (int)(0xffffu - (uint32_t)(key & 0xffffu));
    res[0] = PC_OK;
    res[1] = len;
    res[2] = len ? (int32_t)(i1 + 1 - (uint32_t)len) : 0;
    res[3] = len ? i2 + 1 - len : 0;
  }
}

}  // namespace

void pc_launch_seed(const PcDevBatch &B, int max_len, cudaStream_t s, int sm_count) {
  const int ctas = (B.n + 3) / 4;
  int grid = ctas < sm_count * 8 ? ctas : sm_count * 8;
  // a warp needs about 40 B of scratch per read position plus its candidate lists: as many warps as get a slot of that size
  // (a segment of 6 kbp mRNAs launched at full occupancy had every job come back for a re-run)
  const unsigned long long need = 56ull * (unsigned long long)(max_len > 0 ? max_len : 1) + 16384ull;
  const unsigned long long fit = B.pool_cap / need;
  if (fit < (unsigned long long)grid * 4ull) grid = (int)(fit / 4ull > 0 ? fit / 4ull : 1);
  if (B.max_warps > 0 && grid > (B.max_warps + 3) / 4) grid = (B.max_warps + 3) / 4;
  if (grid < 1) grid = 1;
  PcDevBatch C = B;
  C.slots = (B.max_warps > 0 && B.max_warps < grid * 4) ? B.max_warps : grid * 4;
  k_seed<<<grid, 128, 0, s>>>(C);
  PC_COUNT_LAUNCH(1);
}

// best: device array of B.n 64-bit slots (zeroed here); max_l1/max_l2 over the jobs of the batch
int pc_lcs_blocks(long long l1, int l2) {                       // blocks one job needs (0 for an oversized s2: reported by the finish kernel)
  if (l2 > LCS_MAX_S2 || l2 <= 0 || l1 <= 0) return 0;
  return (int)((l1 + l2 - 1 + LCS_TILE - 1) / LCS_TILE);
}

// best: B.n 64-bit slots; work: 2 * total_tiles + 4 uint32 (tile lengths, tile list, list counter); all zeroed / filled here
int pc_launch_lcs(const PcDevBatch &B, unsigned long long *best, const uint32_t *d_blk_prefix, uint32_t total_tiles, int max_l2, uint32_t *work,
                  cudaStream_t s) {
  if (max_l2 > LCS_MAX_S2) max_l2 = LCS_MAX_S2;
  cudaMemsetAsync(best, 0, sizeof(unsigned long long) * B.n, s);
  if (total_tiles > 0) {
    uint32_t *tile_len = work, *list = work + total_tiles, *count = work + 2 * (size_t)total_tiles;
    cudaMemsetAsync(count, 0, 16, s);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = pc_cached_occupancy((const void *)k_lcs_len, LCS_WPB * 32, 0);
    const uint32_t need1 = (total_tiles + LCS_WPB - 1) / LCS_WPB;
    const uint32_t grid1 = need1 < (uint32_t)(sms * per_sm) ? need1 : (uint32_t)(sms * per_sm);
    k_lcs_len<<<grid1, LCS_WPB * 32, 0, s>>>(B, best, d_blk_prefix, total_tiles, tile_len);
    const uint32_t need2 = (total_tiles + 255) / 256;
    k_lcs_pick<<<need2 < (uint32_t)(sms * 8) ? need2 : (uint32_t)(sms * 8), 256, 0, s>>>(best, d_blk_prefix, B.n, total_tiles, tile_len, list, count);
    size_t sh = LCS_TPB + 2 * (size_t)max_l2 + 8;
    if (sh < LCS_TPB + 64 + 8 + 4 * ((LCS_TPB + 64) / 32 + 2) + 72 + 16) sh = LCS_TPB + 64 + 8 + 4 * ((LCS_TPB + 64) / 32 + 2) + 72 + 16;
    const uint32_t need3 = total_tiles * (LCS_TILE / LCS_TPB);
    const uint32_t cap3 = (uint32_t)(sms * pc_cached_occupancy((const void *)k_lcs_locate, LCS_TPB, sh));
    k_lcs_locate<<<need3 < cap3 ? need3 : cap3, LCS_TPB, sh, s>>>(B, best, d_blk_prefix, list, count);
    PC_COUNT_LAUNCH(3);
  }
  k_lcs_finish<<<(B.n + 127) / 128, 128, 0, s>>>(B, best);
  PC_COUNT_LAUNCH(1);
  return 0;
}

// bit planes of the genome for the LCS scan; planes: 9 * nwords uint32 (grown here), *nwords_out = words per plane
int pc_build_planes(const uint8_t *d_genome, uint32_t len, PcGrowBuf &planes, uint32_t *nwords_out, cudaStream_t s) {
  const uint32_t nwords = (len + 31) / 32 + 8;
  if (planes.reserve(9ull * nwords * sizeof(uint32_t))) return PC_E_NOMEM;
  const uint32_t blocks = (nwords * 32 + 255) / 256;
  k_genome_planes<<<blocks < 4096 ? blocks : 4096, 256, 0, s>>>(d_genome, len, nwords, (uint32_t *)planes.p);
  PC_COUNT_LAUNCH(1);
  *nwords_out = nwords;
  return 0;
}

int pc_build_index(const uint8_t *d_genome, uint32_t len, int word, PcIndexBufs &bufs, unsigned long long **keys_out, uint32_t **pos_out,
                   uint32_t *n_out, uint32_t **bstart_out, int *shift_out, cudaStream_t s) {
  *keys_out = nullptr; *pos_out = nullptr; *n_out = 0; *bstart_out = nullptr; *shift_out = 63;
  if (len < (uint32_t)word) {                        // empty index: one empty bucket pair
    if (bufs.bstart.reserve(3 * sizeof(uint32_t)) || cudaMemsetAsync(bufs.bstart.p, 0, 3 * sizeof(uint32_t), s)) return PC_E_NOMEM;
    if (cudaStreamSynchronize(s) != cudaSuccess) return PC_E_CUDA;
    *bstart_out = (uint32_t *)bufs.bstart.p;
    return 0;
  }
  const uint32_t n = len - word + 1;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 64, s);
  int bits = 12;
  while (bits < 24 && (1u << bits) < 2u * n) ++bits;          // about half an entry per bucket
  const uint32_t nb = 1u << bits;
  if (bufs.keys_in.reserve(8ull * n) || bufs.keys_out.reserve(8ull * n) || bufs.pos_in.reserve(4ull * n) || bufs.pos_out.reserve(4ull * n) ||
      bufs.tmp.reserve(tmp_bytes) || bufs.bstart.reserve((nb + 2ull) * sizeof(uint32_t)))
    return PC_E_NOMEM;
  unsigned long long *k_in = (unsigned long long *)bufs.keys_in.p, *k_out = (unsigned long long *)bufs.keys_out.p;
  uint32_t *p_in = (uint32_t *)bufs.pos_in.p, *p_out = (uint32_t *)bufs.pos_out.p, *bstart = (uint32_t *)bufs.bstart.p;
  k_hash_windows<<<(n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048, 256, 0, s>>>(d_genome, n, word, k_in, p_in);
  PC_COUNT_LAUNCH(1);
  cub::DeviceRadixSort::SortPairs(bufs.tmp.p, tmp_bytes, k_in, k_out, p_in, p_out, (int)n, 0, 64, s);   // stable: equal keys keep ascending t
  k_bucket_starts<<<(n + 256) / 256 < 2048 ? (n + 256) / 256 : 2048, 256, 0, s>>>(k_out, n, 64 - bits, nb, bstart);
  PC_COUNT_LAUNCH(1);
  if (cudaStreamSynchronize(s) != cudaSuccess) return PC_E_CUDA;
  *keys_out = k_out; *pos_out = p_out; *n_out = n; *bstart_out = bstart; *shift_out = 64 - bits;
  return 0;
}
