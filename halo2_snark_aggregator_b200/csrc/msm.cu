// K1/K4: Pippenger windowed-bucket MSM over BN254 G1 for sm_100a.
//
// Replaces halo2_proofs `best_multiexp(coeffs, bases)` (external crate; restated in SURVEY.md
// App. B1) as reached through `ParamsKZG::commit_lagrange` / `commit` inside the reference's
// `create_proof` call, halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994, and
// `keygen_vk` at :760-761.  The result Sum_i s_i * P_i is a unique group element, so any
// correct schedule is bit-exact after normalisation to affine.
//
// Two modes share every kernel:
//   plain  bases given per call: W = ceil(255/c) windows x 2^(c-1) signed-digit buckets each,
//          Horner over windows at the end.
//   table  bases registered as an SRS (ParamsKZG::g / g_lagrange are immutable, so the 180 GB of
//          HBM buy arithmetic): 2^(c w) P_i is precomputed for every window w, all W digits of a
//          scalar fall into ONE set of 2^(c-1) buckets, c grows to 20 (13 windows instead of 16),
//          the per-window reductions and the 240-doubling Horner chain disappear.
//
// Pipeline (one stream, no host synchronisation inside):
//   K4 msm_digits<count>   Montgomery -> canonical scalar, signed c-bit digits, bucket histogram
//      msm_scan1/2/3       exclusive scan -> bucket offsets
//   K4 msm_digits<scatter> counting sort of (base index | sign) by bucket
//   K1 msm_accumulate      the sorted list is cut into chunks of exactly T entries, one thread per
//                          chunk, whatever bucket boundaries fall inside: every lane of a warp does
//                          the same number of XYZZ mixed additions (8M+2S) and a hot bucket
//                          (witness columns are full of 0/1/17-bit values) is split across threads
//                          by construction.  Bases are gathered as 4 x 16-byte read-only loads.
//      msm_fold / msm_fold_hot  stitch the pieces of buckets that straddle chunks (CTA tree for hot ones)
//      msm_wsum            sum_d d * B_d by a 4-ary (S, C) reduction tree: the big levels one launch each, every level
//      msm_wsum_tail       below 2048 items + the Horner over windows (plain mode) in ONE single-CTA launch (XYZZ out)
//      g1_normalize_batch  affine + Jacobian(z=1) for all results of a round with one inversion
#include "bn254_g1.cuh"
#include "ctx.hpp"
#include <algorithm>
#include <cstring>

namespace h2agg {

static constexpr int MSM_THREADS = 128;
static constexpr uint32_t HOT_PIECES = 8;  // buckets cut into more pieces than this get a whole CTA
static constexpr uint32_t WSUM_L = 4;      // arity of the window-sum tree
static constexpr uint32_t WSUM_QUAD_MAX = 16384;  // levels with at most this many outputs use the four-lane kernel

struct MsmGeom {
  uint32_t n;          // scalars in this MSM
  uint32_t c;          // window bits
  uint32_t nwin;       // ceil(255 / c) digit positions
  uint32_t bpw;        // buckets per window = 2^(c-1)
  uint32_t nsets;      // bucket sets: nwin (plain) or 1 (table)
  uint32_t nb;         // nsets * bpw
  uint32_t chunk;      // T (upper bound; see msm_chunk_len)
  uint32_t max_chunks; // launch size of msm_accumulate = capacity of the head / tail partial arrays
  uint32_t win_begin, win_end;
  uint32_t table;      // 1: entries index the precomputed table [w][srs_n]
  uint32_t srs_n;      // row length of the table
};

int msm_window_config(size_t n, int forced_c, int* c_out, int* nwin_out) {
  int c;
  if (forced_c > 0) {
    c = forced_c;
  } else {
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    c = lg - 5;
    if (c > 16) c = 16;
    if (c < 4) c = 4;
  }
  if (c < 2) c = 2;
  if (c > 20) c = 20;
  *c_out = c;
  *nwin_out = (255 + c - 1) / c;
  return 0;
}

int msm_table_config(size_t srs_n, int* c_out, int* nwin_out) {
  int lg = 0;
  while (((size_t)1 << (lg + 1)) <= srs_n) lg++;
  int c = lg - 2;
  if (c > 20) c = 20;
  if (c < 4) c = 4;
  *c_out = c;
  *nwin_out = (255 + c - 1) / c;
  return 0;
}

// signed digit of window w (canonical scalar in v[8]); carry feeds the next window
__device__ __forceinline__ int32_t take_digit(const uint32_t* v, uint32_t w, uint32_t c, uint32_t& carry) {
  uint32_t bit = w * c;
  uint32_t limb = bit >> 5, sh = bit & 31;
  uint64_t two = v[limb];
  if (limb + 1 < 8) two |= (uint64_t)v[limb + 1] << 32;
  uint32_t raw = (uint32_t)(two >> sh) & ((1u << c) - 1);
  raw += carry;
  if (raw > (1u << (c - 1))) {
    carry = 1;
    return (int32_t)raw - (int32_t)(1u << c);
  }
  carry = 0;
  return (int32_t)raw;
}

// Digits are produced in groups of DG windows: all DG atomics of a group are issued before any of
// their results is consumed, so a thread keeps DG L2 round trips in flight instead of one.
static constexpr int DG = 8;

template <bool SCATTER>
__global__ void __launch_bounds__(256) msm_digits(const uint4* __restrict__ scalars, MsmGeom g,
                                                   uint32_t* __restrict__ counts_or_cursor,
                                                   uint32_t* __restrict__ entries) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
    Fr s = Fr::from_halves(__ldg(scalars + 2 * (size_t)i), __ldg(scalars + 2 * (size_t)i + 1));
    if (s.is_zero()) continue;
    s = fp_from_mont(s);
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = s.v[k];
    uint32_t carry = 0;
    for (uint32_t w0 = 0; w0 < g.nwin; w0 += DG) {
      uint32_t key[DG], val[DG], pos[DG];
#pragma unroll
      for (int j = 0; j < DG; j++) {
        uint32_t w = w0 + j;
        key[j] = 0xffffffffu;
        if (w < g.nwin) {
          int32_t d = take_digit(v, w, g.c, carry);
          if (d != 0 && w >= g.win_begin && w < g.win_end) {
            uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
            key[j] = (g.table ? 0u : w * g.bpw) + (mag - 1);
            val[j] = (g.table ? (w * g.srs_n + i) : i) | (d < 0 ? 0x80000000u : 0u);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < DG; j++)
        if (key[j] != 0xffffffffu) pos[j] = atomicAdd(counts_or_cursor + key[j], 1u);
      if (SCATTER) {
#pragma unroll
        for (int j = 0; j < DG; j++)
          if (key[j] != 0xffffffffu) entries[pos[j]] = val[j];
      }
    }
  }
}

// ---- exclusive scan of the bucket histogram, 2048 per block -------------------------------------------
static constexpr uint32_t SCAN_ITEMS = 8, SCAN_THREADS = 256, SCAN_BLOCK = SCAN_ITEMS * SCAN_THREADS;

__device__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* sh /*[SCAN_THREADS/32]*/) {
  uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += x;
  }
  if (lane == 31) sh[wid] = inc;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
  for (uint32_t k = 0; k < SCAN_THREADS / 32; k++) {
    if (k < wid) woff += sh[k];
    tot += sh[k];
  }
  __syncthreads();
  *total = tot;
  return woff + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) msm_scan1(const uint32_t* __restrict__ counts, uint32_t nb,
                                                           uint32_t* block_sums) {
  __shared__ uint32_t sh[SCAN_THREADS / 32];
  uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  uint32_t acc = 0;
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) acc += (base + k < nb) ? counts[base + k] : 0;
  uint32_t tot;
  block_exclusive_scan(acc, &tot, sh);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) msm_scan2(uint32_t* block_sums, uint32_t nblocks) {
  __shared__ uint32_t sh[SCAN_THREADS / 32];
  uint32_t base = threadIdx.x * SCAN_ITEMS;
  uint32_t loc[SCAN_ITEMS];
  uint32_t acc = 0;
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    loc[k] = (base + k < nblocks) ? block_sums[base + k] : 0;
    acc += loc[k];
  }
  uint32_t tot;
  uint32_t off = block_exclusive_scan(acc, &tot, sh);
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < nblocks) block_sums[base + k] = off;
    off += loc[k];
  }
}

// offsets[nb+1] and cursor[nb] (= offsets, consumed by the scatter)
__global__ void __launch_bounds__(SCAN_THREADS) msm_scan3(const uint32_t* __restrict__ counts, uint32_t nb,
                                                           const uint32_t* __restrict__ block_sums, uint32_t* offsets,
                                                           uint32_t* cursor) {
  __shared__ uint32_t sh[SCAN_THREADS / 32];
  uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  uint32_t cnt[SCAN_ITEMS];
  uint32_t acc = 0;
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    cnt[k] = (base + k < nb) ? counts[base + k] : 0;
    acc += cnt[k];
  }
  uint32_t tot;
  uint32_t off = block_exclusive_scan(acc, &tot, sh) + block_sums[blockIdx.x];
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    if (base + k <= nb) {  // also writes the closing element [nb]
      offsets[base + k] = off;
      if (base + k < nb) cursor[base + k] = off;
    }
    off += cnt[k];
  }
}

// Chunk length actually used: g.chunk when the sorted list is long, shorter (down to MSM_MIN_CHUNK, a power of two) when
// it is short -- a column of small cells has n entries instead of n * W, and with 128-entry chunks only a few hundred
// threads per SM would share the work (latency-bound).  The number of chunks never exceeds the launch size
// max_chunks = max_entries / g.chunk + 1.  Every kernel evaluates this on the same `total`.
static constexpr uint32_t MSM_MIN_CHUNK = 16;
__host__ __device__ __forceinline__ uint32_t msm_chunk_len(uint32_t total, uint32_t chunk, uint32_t max_chunks) {
  uint32_t c = chunk;
  while (c > MSM_MIN_CHUNK && (unsigned long long)(total + (c >> 1) - 1) / (c >> 1) <= (unsigned long long)max_chunks) c >>= 1;
  return c;
}

// ---- K1: bucket accumulation over fixed-size chunks of the sorted entry list ---------------------------
// Piece bookkeeping for a bucket [beg, end) cut by chunk boundaries (chunk t = [tT, (t+1)T)):
//   entirely inside one chunk           -> written straight to bucket_sums[b]
//   first piece (bucket starts in t)    -> tail_part[t]
//   every later piece (chunks t+1..)    -> head_part[t']
__global__ void __launch_bounds__(MSM_THREADS) msm_accumulate(const uint8_t* __restrict__ bases,
                                                               const uint32_t* __restrict__ entries,
                                                               const uint32_t* __restrict__ offsets, MsmGeom g,
                                                               uint8_t* __restrict__ head_part,
                                                               uint8_t* __restrict__ tail_part,
                                                               uint8_t* __restrict__ bucket_sums) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = __ldg(offsets + g.nb);
  const uint32_t chunk = msm_chunk_len(total, g.chunk, g.max_chunks);
  const unsigned long long start64 = (unsigned long long)t * chunk;
  if (start64 >= total) return;
  const uint32_t start = (uint32_t)start64;
  const uint32_t end = (uint32_t)min((unsigned long long)total, start64 + chunk);
  // largest b with offsets[b] <= start: the (non-empty) bucket that owns slot `start`
  uint32_t lo = 0, hi = g.nb;  // offsets[lo] <= start < offsets[hi]
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(offsets + mid) <= start) lo = mid; else hi = mid;
  }
  uint32_t b = lo;
  uint32_t bbeg = __ldg(offsets + b), bend = __ldg(offsets + b + 1);
  G1Xyzz acc = G1Xyzz::identity();
  // ONE flat loop over the chunk: the bucket switch is a short predicated side branch, so the
  // lanes of a warp stay converged on the expensive mixed addition (a nested per-bucket loop
  // would make every lane wait for the longest segment of its neighbours).
  for (uint32_t k = start; k < end; k++) {
    if (k == bend) {
      // bucket b is finished inside this chunk (it cannot continue past `end` here)
      if (bbeg < start) acc.store(head_part + (size_t)t * 128);
      else acc.store(bucket_sums + (size_t)b * 128);
      acc = G1Xyzz::identity();
      // next non-empty bucket = the one that owns slot k.  Usually the neighbour; but a column of small values with
      // a few full-width ones (every real halo2 column: 17-bit cells + random blinding rows) leaves gaps of 10^5 empty
      // buckets, so after a few linear steps fall back to the binary search (offsets[b+1] <= k < offsets[nb]).
      b++;
      for (uint32_t tries = 0; __ldg(offsets + b + 1) <= k; ) {
        b++;
        if (++tries == 4) {
          uint32_t lo2 = b, hi2 = g.nb;  // offsets[lo2] <= k < offsets[hi2]
          while (hi2 - lo2 > 1) {
            uint32_t mid = (lo2 + hi2) >> 1;
            if (__ldg(offsets + mid) <= k) lo2 = mid; else hi2 = mid;
          }
          b = lo2;
          break;
        }
      }
      bbeg = __ldg(offsets + b);
      bend = __ldg(offsets + b + 1);
    }
    uint32_t e = __ldg(entries + k);
    G1Affine q = G1Affine::load_nc(bases + (size_t)(e & 0x7fffffffu) * 64);
    if (e >> 31) q.y = fp_neg(q.y);
    xyzz_madd(acc, q);
  }
  if (bbeg < start) acc.store(head_part + (size_t)t * 128);
  else if (bend > end) acc.store(tail_part + (size_t)t * 128);
  else acc.store(bucket_sums + (size_t)b * 128);
}

// a whole XYZZ point from another lane of the warp
__device__ __forceinline__ G1Xyzz xyzz_shfl(const G1Xyzz& p, int src_lane) {
  G1Xyzz r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.x.v[i] = __shfl_sync(0xffffffffu, p.x.v[i], src_lane);
    r.y.v[i] = __shfl_sync(0xffffffffu, p.y.v[i], src_lane);
    r.zz.v[i] = __shfl_sync(0xffffffffu, p.zz.v[i], src_lane);
    r.zzz.v[i] = __shfl_sync(0xffffffffu, p.zzz.v[i], src_lane);
  }
  return r;
}

// stitch buckets that straddle chunk boundaries; empty buckets -> identity; hot ones are queued
__global__ void __launch_bounds__(MSM_THREADS) msm_fold(const uint32_t* __restrict__ offsets, MsmGeom g,
                                                         const uint8_t* __restrict__ head_part,
                                                         const uint8_t* __restrict__ tail_part,
                                                         uint8_t* __restrict__ bucket_sums, uint32_t* hot_count,
                                                         uint32_t* hot_list) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= g.nb) return;
  uint32_t beg = offsets[b], end = offsets[b + 1];
  if (beg == end) {
    G1Xyzz::identity().store(bucket_sums + (size_t)b * 128);
    return;
  }
  const uint32_t chunk = msm_chunk_len(offsets[g.nb], g.chunk, g.max_chunks);
  uint32_t t0 = beg / chunk, t1 = (end - 1) / chunk;
  if (t0 == t1) return;  // complete, already written by its chunk
  if (t1 - t0 + 1 > HOT_PIECES) {
    hot_list[atomicAdd(hot_count, 1u)] = b;
    return;
  }
  G1Xyzz acc = G1Xyzz::load(tail_part + (size_t)t0 * 128);
  for (uint32_t t = t0 + 1; t <= t1; t++) xyzz_add(acc, G1Xyzz::load(head_part + (size_t)t * 128));
  acc.store(bucket_sums + (size_t)b * 128);
}

// A hot bucket's pieces are cut into HOT_SPLIT slices, one CTA each (a column of small cells has a handful of very hot
// buckets -- the value 1 -- and a single CTA walking 10^4 pieces would be the longest kernel of the MSM): stage 1 reduces
// every slice into the head_part slot at the slice's start (in place: slices are disjoint), stage 2 adds the slice sums
// and the bucket's first piece.
static constexpr uint32_t HOT_SPLIT = 16;

__device__ __forceinline__ void hot_slice(uint32_t t0, uint32_t t1, uint32_t j, uint32_t& lo, uint32_t& hi) {
  const uint32_t pieces = t1 - t0;  // head pieces t0+1 .. t1
  const uint32_t sz = (pieces + HOT_SPLIT - 1) / HOT_SPLIT;
  lo = t0 + 1 + j * sz;
  hi = min(lo + sz, t1 + 1);
}

__global__ void __launch_bounds__(256) msm_fold_hot(const uint32_t* __restrict__ offsets, MsmGeom g, uint8_t* head_part,
                                                     const uint32_t* __restrict__ hot_count,
                                                     const uint32_t* __restrict__ hot_list) {
  __shared__ uint4 sh[256 * 8];  // one XYZZ point (128 B) per thread
  const uint32_t nhot = *hot_count;
  const uint32_t chunk = msm_chunk_len(offsets[g.nb], g.chunk, g.max_chunks);
  for (uint32_t item = blockIdx.x; item < nhot * HOT_SPLIT; item += gridDim.x) {
    const uint32_t b = hot_list[item / HOT_SPLIT];
    const uint32_t beg = offsets[b], end = offsets[b + 1];
    uint32_t lo, hi;
    hot_slice(beg / chunk, (end - 1) / chunk, item % HOT_SPLIT, lo, hi);
    if (lo >= hi) continue;  // CTA-uniform
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t t = lo + threadIdx.x; t < hi; t += blockDim.x) xyzz_add(acc, G1Xyzz::load(head_part + (size_t)t * 128));
    acc.store(sh + threadIdx.x * 8);
    __syncthreads();
    for (uint32_t o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o && threadIdx.x + o < hi - lo) {  // partners beyond the slice hold the identity
        G1Xyzz a = G1Xyzz::load(sh + threadIdx.x * 8);
        xyzz_add(a, G1Xyzz::load(sh + (threadIdx.x + o) * 8));
        a.store(sh + threadIdx.x * 8);
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) G1Xyzz::load(sh).store(head_part + (size_t)lo * 128);
    __syncthreads();
  }
}

// stage 2: one warp per hot bucket; lane j < HOT_SPLIT holds slice j's sum, lane HOT_SPLIT the bucket's first piece
__global__ void __launch_bounds__(128) msm_fold_hot2(const uint32_t* __restrict__ offsets, MsmGeom g,
                                                      const uint8_t* __restrict__ head_part,
                                                      const uint8_t* __restrict__ tail_part, uint8_t* __restrict__ bucket_sums,
                                                      const uint32_t* __restrict__ hot_count,
                                                      const uint32_t* __restrict__ hot_list) {
  static_assert(HOT_SPLIT < 32, "slice sums + the first piece must fit one warp");
  const uint32_t nhot = *hot_count;
  const uint32_t chunk = msm_chunk_len(offsets[g.nb], g.chunk, g.max_chunks);
  const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t h = warp; h < nhot; h += nwarps) {  // warp-uniform
    const uint32_t b = hot_list[h];
    const uint32_t beg = offsets[b], end = offsets[b + 1];
    const uint32_t t0 = beg / chunk, t1 = (end - 1) / chunk;
    G1Xyzz acc = G1Xyzz::identity();
    if (lane < HOT_SPLIT) {
      uint32_t lo, hi;
      hot_slice(t0, t1, lane, lo, hi);
      if (lo < hi) acc = G1Xyzz::load(head_part + (size_t)lo * 128);
    } else if (lane == HOT_SPLIT) {
      acc = G1Xyzz::load(tail_part + (size_t)t0 * 128);
    }
    for (int o = 16; o > 0; o >>= 1) {
      G1Xyzz other = xyzz_shfl(acc, (int)((lane + o) & 31));  // wrapped reads are ignored below
      if (lane + o < 32) xyzz_add(acc, other);
    }
    if (lane == 0) acc.store(bucket_sums + (size_t)b * 128);
  }
}

// ---- window sums: sum_idx (idx * S_idx + C_idx) by an L-ary tree -------------------------------------
// level input: per bucket set m items (S, C); output ceil(m/L) items.  At level 0, C aliases S
// (bucket idx holds digit idx+1).
__global__ void __launch_bounds__(MSM_THREADS) msm_wsum(const uint8_t* __restrict__ s_in, const uint8_t* __restrict__ c_in,
                                                         uint32_t nsets, uint32_t m, uint8_t* __restrict__ s_out,
                                                         uint8_t* __restrict__ c_out) {
  uint32_t mo = (m + WSUM_L - 1) / WSUM_L;
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nsets * mo) return;
  uint32_t w = gid / mo, t = gid % mo;
  uint32_t first = t * WSUM_L;
  uint32_t cnt = min(WSUM_L, m - first);
  size_t base = ((size_t)w * m + first) * 128;
  G1Xyzz running = G1Xyzz::identity(), acc = G1Xyzz::identity(), csum = G1Xyzz::identity();
  for (int i = (int)cnt - 1; i >= 0; i--) {
    xyzz_add(running, G1Xyzz::load(s_in + base + (size_t)i * 128));
    if (i > 0) xyzz_add(acc, running);
    xyzz_add(csum, G1Xyzz::load(c_in + base + (size_t)i * 128));
  }
  xyzz_add(acc, csum);
  for (uint32_t l = 1; l < WSUM_L; l <<= 1) running = xyzz_dbl(running);
  size_t ob = ((size_t)w * mo + t) * 128;
  running.store(s_out + ob);
  acc.store(c_out + ob);
}

// The same node, spread over four adjacent lanes so that the serial chain of a level is 5 point operations instead of
// 14 (the levels with few outputs are pure latency: ten of them cost ~1.3 ms per MSM, which is what bounds the small
// configurations).  Lane roles within a quad:  R running sums (S3, +S2, +S1, +S0, then two doublings),
// A  acc = r3 + r2 + r1 (fed by R through shuffles) and finally + csum,  C  csum = C3 + C2 + C1 + C0,  the 4th lane idles.
// All lanes execute the same xyzz_add / xyzz_dbl calls in lockstep on role-selected operands.
__global__ void __launch_bounds__(MSM_THREADS) msm_wsum_quad(const uint8_t* __restrict__ s_in, const uint8_t* __restrict__ c_in,
                                                              uint32_t nsets, uint32_t m, uint8_t* __restrict__ s_out,
                                                              uint8_t* __restrict__ c_out) {
  static_assert(WSUM_L == 4, "the quad kernel is written for arity 4");
  const uint32_t mo = (m + WSUM_L - 1) / WSUM_L;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t gid = tid >> 2, role = tid & 3;  // 0 = R, 1 = A, 2 = C, 3 = idle
  const int lane = threadIdx.x & 31, base_lane = lane & ~3;
  const bool live = gid < nsets * mo;             // whole quads are live or not; dead quads still take part in the shuffles
  const uint32_t w = live ? gid / mo : 0, t = live ? gid % mo : 0;
  const uint32_t first = t * WSUM_L;
  const uint32_t cnt = live ? min(WSUM_L, m - first) : 0;
  const size_t base = ((size_t)w * m + first) * 128;
  auto item = [&](const uint8_t* arr, uint32_t i) { return (live && i < cnt) ? G1Xyzz::load(arr + base + (size_t)i * 128) : G1Xyzz::identity(); };
  // step 0: R holds S3, C holds C3; operands of the first addition
  G1Xyzz acc = G1Xyzz::identity();
  if (role == 0) acc = item(s_in, 3);
  if (role == 2) acc = item(c_in, 3);
  G1Xyzz r_prev = xyzz_shfl(acc, base_lane);      // r3, for A
  // step 1: R += S2, C += C2
  {
    G1Xyzz y = G1Xyzz::identity();
    if (role == 0) y = item(s_in, 2);
    if (role == 2) y = item(c_in, 2);
    xyzz_add(acc, y);
  }
  G1Xyzz r_cur = xyzz_shfl(acc, base_lane);       // r2
  if (role == 1) acc = r_prev;                     // A starts at r3 ...
  // step 2: R += S1, A += r2, C += C1
  {
    G1Xyzz y = G1Xyzz::identity();
    if (role == 0) y = item(s_in, 1);
    if (role == 1) y = r_cur;
    if (role == 2) y = item(c_in, 1);
    xyzz_add(acc, y);
  }
  r_cur = xyzz_shfl(acc, base_lane);              // r1
  // step 3: R += S0, A += r1, C += C0
  {
    G1Xyzz y = G1Xyzz::identity();
    if (role == 0) y = item(s_in, 0);
    if (role == 1) y = r_cur;
    if (role == 2) y = item(c_in, 0);
    xyzz_add(acc, y);
  }
  G1Xyzz csum = xyzz_shfl(acc, base_lane + 2);
  // step 4: A += csum; R doubles twice (4 * running)
  if (role == 1) xyzz_add(acc, csum);
  if (role == 0) {
    acc = xyzz_dbl(acc);
    acc = xyzz_dbl(acc);
  }
  if (!live) return;
  const size_t ob = ((size_t)w * mo + t) * 128;
  if (role == 0) acc.store(s_out + ob);
  if (role == 1) acc.store(c_out + ob);
}

// One quad node of the tree (see msm_wsum_quad): lanes (R, A, C, idle) of a quad cooperate on a node; all lanes of the
// warp execute the same calls.  `live` quads write their node; dead quads only take part in the shuffles.
__device__ __forceinline__ void wsum_quad_node(const uint8_t* __restrict__ s_in, const uint8_t* __restrict__ c_in, uint32_t m,
                                               uint32_t mo, uint32_t gid, bool live, uint32_t role, int base_lane,
                                               uint8_t* __restrict__ s_out, uint8_t* __restrict__ c_out) {
  const uint32_t w = live ? gid / mo : 0, t = live ? gid % mo : 0;
  const uint32_t first = t * WSUM_L;
  const uint32_t cnt = live ? min(WSUM_L, m - first) : 0;
  const size_t base = ((size_t)w * m + first) * 128;
  auto item = [&](const uint8_t* arr, uint32_t i) { return (live && i < cnt) ? G1Xyzz::load(arr + base + (size_t)i * 128) : G1Xyzz::identity(); };
  G1Xyzz acc = G1Xyzz::identity();
  if (role == 0) acc = item(s_in, 3);
  if (role == 2) acc = item(c_in, 3);
  G1Xyzz r_prev = xyzz_shfl(acc, base_lane);
  {
    G1Xyzz y = G1Xyzz::identity();
    if (role == 0) y = item(s_in, 2);
    if (role == 2) y = item(c_in, 2);
    xyzz_add(acc, y);
  }
  G1Xyzz r_cur = xyzz_shfl(acc, base_lane);
  if (role == 1) acc = r_prev;
  {
    G1Xyzz y = G1Xyzz::identity();
    if (role == 0) y = item(s_in, 1);
    if (role == 1) y = r_cur;
    if (role == 2) y = item(c_in, 1);
    xyzz_add(acc, y);
  }
  r_cur = xyzz_shfl(acc, base_lane);
  {
    G1Xyzz y = G1Xyzz::identity();
    if (role == 0) y = item(s_in, 0);
    if (role == 1) y = r_cur;
    if (role == 2) y = item(c_in, 0);
    xyzz_add(acc, y);
  }
  G1Xyzz csum = xyzz_shfl(acc, base_lane + 2);
  if (role == 1) xyzz_add(acc, csum);
  if (role == 0) {
    acc = xyzz_dbl(acc);
    acc = xyzz_dbl(acc);
  }
  if (!live) return;
  const size_t ob = ((size_t)w * mo + t) * 128;
  if (role == 0) acc.store(s_out + ob);
  if (role == 1) acc.store(c_out + ob);
}

// The tail of an MSM in ONE launch of one CTA: the last levels of the window-sum tree (from <= WSUM_TAIL_MAX items over all
// sets down to one per set) and the Horner over the windows -- four launches of msm_wsum_quad plus msm_final before.  The
// result stays in XYZZ form (no inversion here: the caller normalises all results of a round with ONE inversion,
// g1_normalize_batch, instead of a 170 us Fermat ladder on the critical path of every MSM).
static constexpr uint32_t WSUM_TAIL_THREADS = 256;   // 255 registers per thread: the quad node keeps several XYZZ points live
static constexpr uint32_t WSUM_TAIL_MAX = 256;    // = one iteration per level for the CTA (measured: a larger tail serialises ~17 us point operations and loses to one launch per level)

__global__ void __launch_bounds__(WSUM_TAIL_THREADS) msm_wsum_tail(const uint8_t* s_in, const uint8_t* c_in, MsmGeom g, uint32_t m,
                                                                    uint8_t* s0, uint8_t* c0, uint8_t* s1, uint8_t* c1,
                                                                    uint8_t* out_xyzz) {
  const uint32_t role = threadIdx.x & 3;
  const int base_lane = (threadIdx.x & 31) & ~3;
  uint8_t* lvl_s[2] = {s0, s1};
  uint8_t* lvl_c[2] = {c0, c1};
  int flip = 0;
  bool first = true;
  while (m > 1 || first) {  // at least one level so that C holds sum (idx + 1) * B_idx
    first = false;
    const uint32_t mo = (m + WSUM_L - 1) / WSUM_L;
    const uint32_t total = g.nsets * mo;
    for (uint32_t q0 = 0; q0 < total; q0 += WSUM_TAIL_THREADS / 4) {   // CTA-uniform trip count
      const uint32_t gid = q0 + (threadIdx.x >> 2);
      wsum_quad_node(s_in, c_in, m, mo, gid, gid < total, role, base_lane, lvl_s[flip], lvl_c[flip]);
    }
    __syncthreads();   // global writes of this CTA are visible to it after the barrier
    s_in = lvl_s[flip];
    c_in = lvl_c[flip];
    flip ^= 1;
    m = mo;
  }
  if (threadIdx.x != 0) return;
  G1Xyzz acc = G1Xyzz::identity();
  if (g.table) {
    acc = G1Xyzz::load(c_in);
  } else {  // plain mode: Horner over windows [wb, we), times 2^(c*wb)
    for (int w = (int)g.win_end - 1; w >= (int)g.win_begin; w--) {
      if (!acc.is_identity())
        for (uint32_t k = 0; k < g.c; k++) acc = xyzz_dbl(acc);
      xyzz_add(acc, G1Xyzz::load(c_in + (size_t)w * 128));
    }
    if (!acc.is_identity())
      for (uint32_t k = 0; k < g.c * g.win_begin; k++) acc = xyzz_dbl(acc);
  }
  acc.store(out_xyzz);
}

// XYZZ results (128 B at the start of each 160-byte slot) -> affine (64 B) + normalised Jacobian (96 B) in place, with ONE
// inversion for the whole batch (Montgomery's trick over the ZZZ's; 1/ZZ = ZZZ^-2 ZZ^2).  One thread: n is a commit
// round's column count.
__global__ void g1_normalize_batch(uint8_t* out160s, uint32_t n) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Fq run = Fq::one();
  for (uint32_t i = 0; i < n; i++) {   // prefix products parked in the slot's spare 32 bytes
    G1Xyzz p = G1Xyzz::load(out160s + (size_t)i * 160);
    run.store(out160s + (size_t)i * 160 + 128);
    if (!p.is_identity()) run = run * p.zzz;
  }
  Fq inv = fp_inv(run);
  for (uint32_t i = n; i-- > 0;) {
    uint8_t* o = out160s + (size_t)i * 160;
    G1Xyzz p = G1Xyzz::load(o);
    const bool id = p.is_identity();
    Fq x = Fq::zero(), y = Fq::zero();
    if (!id) {
      Fq izzz = inv * Fq::load(o + 128);
      inv = inv * p.zzz;
      Fq izz = fp_sqr(izzz) * fp_sqr(p.zz);
      x = p.x * izz;
      y = p.y * izzz;
    }
    x.store(o);
    y.store(o + 32);
    x.store(o + 64);
    (id ? Fq::one() : y).store(o + 96);
    (id ? Fq::zero() : Fq::one()).store(o + 128);
  }
}

int g1_normalize(h2agg_ctx* ctx, cudaStream_t st, void* d_out160s, size_t n) {
  if (n == 0) return 0;
  g1_normalize_batch<<<1, 32, 0, st>>>((uint8_t*)d_out160s, (uint32_t)n);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

__device__ __forceinline__ void write_out160(const G1Xyzz& acc, uint8_t* out160) {
  G1Affine a = xyzz_to_affine(acc);
  bool id = acc.is_identity();
  a.x.store(out160);
  a.y.store(out160 + 32);
  a.x.store(out160 + 64);
  (id ? Fq::one() : a.y).store(out160 + 96);
  (id ? Fq::zero() : Fq::one()).store(out160 + 128);
}

// one thread per output: out[j] = sum_i point(i, j); point(i, j) at pts + i * stride + j * 160 (Jacobian, 96 B)
__global__ void g1_sum_kernel(const uint8_t* __restrict__ pts, uint32_t m, size_t stride, uint32_t n_out, uint8_t* out160) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const uint8_t* pts96 = pts + (size_t)j * 160;
  out160 += (size_t)j * 160;
  G1Xyzz acc = G1Xyzz::identity();
  for (uint32_t i = 0; i < m; i++) {
    Fq x = Fq::load(pts96 + (size_t)i * stride), y = Fq::load(pts96 + (size_t)i * stride + 32),
       z = Fq::load(pts96 + (size_t)i * stride + 64);
    if (z.is_zero()) continue;
    G1Xyzz p;  // general Jacobian -> XYZZ: (X, Y, Z^2, Z^3)
    p.x = x; p.y = y; p.zz = fp_sqr(z); p.zzz = p.zz * z;
    xyzz_add(acc, p);
  }
  write_out160(acc, out160);
}

int g1_sum_jacobian(h2agg_ctx* ctx, const void* d_points96, size_t m, void* d_out160, size_t stride, size_t n_out) {
  g1_sum_kernel<<<(unsigned)((n_out + 31) / 32), 32, 0, ctx->stream>>>((const uint8_t*)d_points96, (uint32_t)m, stride,
                                                                     (uint32_t)n_out, (uint8_t*)d_out160);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ---- SRS table: rows w = 0..W-1 of 2^(c w) P_i in affine form ---------------------------------------------
// One thread per point: Jacobian doubling chain, Z's batch-inverted per thread (Montgomery's trick).
static constexpr int MAX_TABLE_ROWS = 64;

__global__ void __launch_bounds__(128) msm_build_table(const uint8_t* __restrict__ bases, uint32_t n, uint32_t c,
                                                        uint32_t nwin, uint8_t* __restrict__ table) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = G1Affine::load_nc(bases + (size_t)i * 64);
  const size_t row = (size_t)n * 64;
  p.x.store(table + (size_t)i * 64);
  p.y.store(table + (size_t)i * 64 + 32);
  if (p.is_identity()) {
    for (uint32_t w = 1; w < nwin; w++) {
      p.x.store(table + w * row + (size_t)i * 64);
      p.y.store(table + w * row + (size_t)i * 64 + 32);
    }
    return;
  }
  // Jacobian chain; X, Y parked in the table slot, Z prefix products in local memory
  Fq X = p.x, Y = p.y, Z = Fq::one();
  Fq prefix[MAX_TABLE_ROWS];
  Fq zs[MAX_TABLE_ROWS];
  Fq run = Fq::one();
  for (uint32_t w = 1; w < nwin; w++) {
    for (uint32_t k = 0; k < c; k++) {  // dbl-2009-l (a = 0); BN254 G1 has no 2-torsion so Y != 0
      Fq A = fp_sqr(X), B = fp_sqr(Y), C = fp_sqr(B);
      Fq t = X + B;
      Fq D = fp_dbl(fp_sqr(t) - A - C);
      Fq E = fp_dbl(A) + A, F = fp_sqr(E);
      Fq Z3 = fp_dbl(Y * Z);
      X = F - fp_dbl(D);
      Y = E * (D - X) - fp_dbl(fp_dbl(fp_dbl(C)));
      Z = Z3;
    }
    X.store(table + w * row + (size_t)i * 64);
    Y.store(table + w * row + (size_t)i * 64 + 32);
    zs[w] = Z;
    prefix[w] = run;  // product of Z_1 .. Z_{w-1}
    run = run * Z;
  }
  Fq inv = fp_inv(run);
  for (uint32_t w = nwin - 1; w >= 1; w--) {
    Fq zi = inv * prefix[w];  // 1 / Z_w
    inv = inv * zs[w];
    Fq zi2 = fp_sqr(zi);
    Fq x = Fq::load(table + w * row + (size_t)i * 64) * zi2;
    Fq y = Fq::load(table + w * row + (size_t)i * 64 + 32) * zi2 * zi;
    x.store(table + w * row + (size_t)i * 64);
    y.store(table + w * row + (size_t)i * 64 + 32);
  }
}

int msm_build_srs_table(h2agg_ctx* ctx, Srs& s) {
  int c, nwin;
  msm_table_config(s.n, &c, &nwin);
  if (nwin > MAX_TABLE_ROWS || (size_t)nwin * s.n >= (1ull << 31)) return 0;  // stay in plain mode
  void* t = nullptr;
  if (cudaMalloc(&t, (size_t)nwin * s.n * 64) != cudaSuccess) {
    cudaGetLastError();
    return 0;  // not enough memory for the table: plain mode still works
  }
  msm_build_table<<<(unsigned)((s.n + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)s.d_bases, (uint32_t)s.n,
                                                                          (uint32_t)c, (uint32_t)nwin, (uint8_t*)t);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  s.d_table = t;
  s.table_c = c;
  s.table_nwin = nwin;
  return 0;
}

int lanes_init(h2agg_ctx* ctx) {
  if (ctx->fork_ev) return 0;
  H2AGG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  for (int i = 0; i < N_LANES; i++) {
    H2AGG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->lanes[i].st, cudaStreamNonBlocking));
    H2AGG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->lanes[i].done, cudaEventDisableTiming));
    H2AGG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->lanes[i].up, cudaEventDisableTiming));
  }
  return 0;
}

int msm_run_batch(h2agg_ctx* ctx, const MsmBases& bases, const void* const* cols, size_t n_cols, size_t n,
                  uint8_t* d_out160s, bool host_cols, int win_begin, int win_end, const int* win_begins, const int* win_ends) {
  if (n_cols == 0) return 0;
  int rc;
  if (n_cols == 1 && !host_cols)
    return msm_run(ctx, ctx->stream, ctx->msm_ws, bases, cols[0], n, d_out160s, win_begins ? win_begins[0] : win_begin,
                   win_ends ? win_ends[0] : win_end, true);
  if ((rc = lanes_init(ctx))) return rc;
  LaneFork lf(ctx);
  if ((rc = lf.fork())) return rc;
  for (size_t i = 0; i < n_cols; i++) {
    Lane& ln = ctx->lanes[i % N_LANES];
    const void* d_col = cols[i];
    if (host_cols) {
      if ((rc = ensure(ctx, ln.io, n * 32 + 64))) return rc;
      if (n) H2AGG_CUDA(ctx, cudaMemcpyAsync(ln.io.p, cols[i], n * 32, cudaMemcpyHostToDevice, ln.st));
      d_col = ln.io.p;
    }
    if ((rc = msm_run(ctx, ln.st, ln.ws, bases, d_col, n, d_out160s + i * 160, win_begins ? win_begins[i] : win_begin,
                      win_ends ? win_ends[i] : win_end, false)))
      return rc;
  }
  if ((rc = lf.join())) return rc;
  return g1_normalize(ctx, ctx->stream, d_out160s, n_cols);   // ONE inversion for the whole round
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int msm_run(h2agg_ctx* ctx, cudaStream_t st, DevBuf& wsbuf, const MsmBases& bases, const void* d_scalars, size_t n,
            void* d_out160, int win_begin, int win_end, bool normalize) {
  if (n >= (1ull << 31)) {
    ctx->last_error = "msm: n must be < 2^31";
    return 1;
  }
  MsmGeom g;
  int c, nwin;
  const bool table = bases.d_table != nullptr && ctx->msm_window_bits == 0;
  if (table) {
    c = bases.table_c;
    nwin = bases.table_nwin;
  } else {
    msm_window_config(n ? n : 1, ctx->msm_window_bits, &c, &nwin);
  }
  g.n = (uint32_t)n;
  g.c = (uint32_t)c;
  g.nwin = (uint32_t)nwin;
  g.bpw = 1u << (c - 1);
  g.nsets = table ? 1u : g.nwin;
  g.nb = g.nsets * g.bpw;
  g.table = table ? 1u : 0u;
  g.srs_n = (uint32_t)bases.srs_n;
  g.win_begin = win_begin < 0 ? 0 : (uint32_t)win_begin;
  g.win_end = (win_end < 0 || win_end > nwin) ? (uint32_t)nwin : (uint32_t)win_end;
  if (g.win_begin > g.win_end) g.win_begin = g.win_end;
  g.chunk = 128;
  const uint8_t* d_points = (const uint8_t*)(table ? bases.d_table : bases.d_bases);
  const size_t max_entries = (size_t)n * (g.win_end - g.win_begin);
  if (max_entries >= (1ull << 32) - 1) {
    ctx->last_error = "msm: n * windows exceeds the 32-bit entry index";
    return 1;
  }
  const size_t max_chunks = max_entries / g.chunk + 1;
  g.max_chunks = (uint32_t)max_chunks;
  const uint32_t scan_blocks = (g.nb + 1 + SCAN_BLOCK - 1) / SCAN_BLOCK;
  if (scan_blocks > SCAN_BLOCK) {
    ctx->last_error = "msm: too many buckets for the scan";
    return 1;
  }

  // workspace carve-up
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  size_t o_counts = carve((size_t)(g.nb + 1) * 4);
  size_t o_offsets = carve((size_t)(g.nb + 1) * 4);
  size_t o_cursor = carve((size_t)(g.nb + 1) * 4);
  size_t o_bsums = carve((size_t)scan_blocks * 4);
  size_t o_hot = carve((size_t)(g.nb + 1) * 4 + 256);
  size_t o_entries = carve((max_entries + 1) * 4);
  size_t o_head = carve(max_chunks * 128);
  size_t o_tail = carve(max_chunks * 128);
  size_t o_buckets = carve((size_t)g.nb * 128);
  size_t lvl_items = (size_t)g.nsets * ((g.bpw + WSUM_L - 1) / WSUM_L);
  size_t o_lvl_s0 = carve(lvl_items * 128), o_lvl_c0 = carve(lvl_items * 128);
  size_t o_lvl_s1 = carve(lvl_items * 128), o_lvl_c1 = carve(lvl_items * 128);
  int rc = ensure(ctx, wsbuf, off);
  if (rc) return rc;
  uint8_t* ws = (uint8_t*)wsbuf.p;
  uint32_t* counts = (uint32_t*)(ws + o_counts);
  uint32_t* offsets = (uint32_t*)(ws + o_offsets);
  uint32_t* cursor = (uint32_t*)(ws + o_cursor);
  uint32_t* bsums = (uint32_t*)(ws + o_bsums);
  uint32_t* hot_count = (uint32_t*)(ws + o_hot);
  uint32_t* hot_list = hot_count + 64;
  uint32_t* entries = (uint32_t*)(ws + o_entries);
  uint8_t* head_part = ws + o_head;
  uint8_t* tail_part = ws + o_tail;
  uint8_t* buckets = ws + o_buckets;

  ScopedKernelTimer t_total(ctx, KC_MSM_TOTAL, st);
  H2AGG_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)(g.nb + 1) * 4, st));
  H2AGG_CUDA(ctx, cudaMemsetAsync(hot_count, 0, 256, st));
  const uint32_t dgrid = (uint32_t)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
  if (n) {
    ScopedKernelTimer tk(ctx, KC_MSM_DIGITS, st);
    msm_digits<false><<<dgrid, 256, 0, st>>>((const uint4*)d_scalars, g, counts, nullptr);
    ctx->launches++;
  }
  msm_scan1<<<scan_blocks, SCAN_THREADS, 0, st>>>(counts, g.nb, bsums);
  msm_scan2<<<1, SCAN_THREADS, 0, st>>>(bsums, scan_blocks);
  msm_scan3<<<scan_blocks, SCAN_THREADS, 0, st>>>(counts, g.nb, bsums, offsets, cursor);
  ctx->launches += 3;
  if (n) {
    {
      ScopedKernelTimer tk(ctx, KC_MSM_DIGITS, st);
      msm_digits<true><<<dgrid, 256, 0, st>>>((const uint4*)d_scalars, g, cursor, entries);
      ctx->launches++;
    }
    ScopedKernelTimer tk(ctx, KC_MSM_ACCUMULATE, st);
    msm_accumulate<<<(uint32_t)((max_chunks + MSM_THREADS - 1) / MSM_THREADS), MSM_THREADS, 0, st>>>(
        d_points, entries, offsets, g, head_part, tail_part, buckets);
    ctx->launches++;
  }
  ScopedKernelTimer t_red(ctx, KC_MSM_REDUCE, st);
  msm_fold<<<(g.nb + MSM_THREADS - 1) / MSM_THREADS, MSM_THREADS, 0, st>>>(offsets, g, head_part, tail_part, buckets,
                                                                            hot_count, hot_list);
  msm_fold_hot<<<ctx->sm_count * 2, 256, 0, st>>>(offsets, g, head_part, hot_count, hot_list);
  msm_fold_hot2<<<16, 128, 0, st>>>(offsets, g, head_part, tail_part, buckets, hot_count, hot_list);
  ctx->launches += 3;
  H2AGG_CUDA(ctx, cudaGetLastError());

  // window sums
  const uint8_t *s_in = buckets, *c_in = buckets;
  uint8_t* lvl_s[2] = {ws + o_lvl_s0, ws + o_lvl_s1};
  uint8_t* lvl_c[2] = {ws + o_lvl_c0, ws + o_lvl_c1};
  uint32_t m = g.bpw;
  int flip = 0;
  while ((size_t)g.nsets * m > WSUM_TAIL_MAX) {   // the big levels: one launch each
    uint32_t mo = (m + WSUM_L - 1) / WSUM_L;
    uint32_t total = g.nsets * mo;
    if (total <= WSUM_QUAD_MAX)  // few outputs: latency-bound, spread every node over four lanes
      msm_wsum_quad<<<(4 * total + MSM_THREADS - 1) / MSM_THREADS, MSM_THREADS, 0, st>>>(s_in, c_in, g.nsets, m, lvl_s[flip],
                                                                                      lvl_c[flip]);
    else
      msm_wsum<<<(total + MSM_THREADS - 1) / MSM_THREADS, MSM_THREADS, 0, st>>>(s_in, c_in, g.nsets, m, lvl_s[flip],
                                                                                lvl_c[flip]);
    ctx->launches++;
    s_in = lvl_s[flip];
    c_in = lvl_c[flip];
    flip ^= 1;
    m = mo;
  }
  // every remaining level + the Horner over the windows: one CTA, one launch; XYZZ result into the 160-byte slot
  msm_wsum_tail<<<1, WSUM_TAIL_THREADS, 0, st>>>(s_in, c_in, g, m, lvl_s[flip], lvl_c[flip], lvl_s[flip ^ 1], lvl_c[flip ^ 1],
                                                 (uint8_t*)d_out160);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  if (normalize) return g1_normalize(ctx, st, d_out160, 1);
  return 0;
}

}  // namespace h2agg
