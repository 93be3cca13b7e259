// K1/K4: Pippenger windowed-bucket MSM over BN254 G1 for sm_100a.
//
// Replaces halo2_proofs `best_multiexp(coeffs, bases)` (external crate; restated in SURVEY.md
// App. B1) as reached through `ParamsKZG::commit_lagrange` / `commit` inside the reference's
// `create_proof` call, halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994, and
// `keygen_vk` at :760-761.  The result Sum_i s_i * P_i is a unique group element, so any
// correct schedule is bit-exact after normalisation to affine.
//
// Pipeline (all on one stream, no host synchronisation inside):
//   K4 msm_count    Montgomery -> canonical scalar, signed c-bit digits, bucket histogram
//      scan         bucket offsets + per-bucket task counts (buckets longer than T points are
//                   split into T-point tasks so one hot bucket -- witness columns are full of
//                   0/1/17-bit values -- cannot serialise the kernel)
//   K4 msm_scatter  counting sort of (point index | sign) by bucket
//   K1 msm_accumulate  one thread per task: gather 64-byte affine bases as 4 x uint4,
//                   XYZZ mixed additions (8M + 2S), exact handling of P+P / P-P / identity
//      msm_fold / msm_fold_hot   combine the tasks of split buckets (CTA tree for hot ones)
//      msm_wsum     sum_d d * B_d per window by a 4-ary (S, C) reduction tree
//      msm_final    Horner over windows, one inversion, affine + Jacobian(z=1) out
#include "bn254_g1.cuh"
#include "ctx.hpp"
#include <cstring>

namespace h2agg {

static constexpr int MSM_THREADS = 128;
static constexpr uint32_t HOT_TASKS = 8;   // buckets with more tasks than this get a whole CTA
static constexpr uint32_t WSUM_L = 4;      // arity of the window-sum tree

struct MsmGeom {
  uint32_t n;
  uint32_t c;         // window bits
  uint32_t nwin;      // ceil(255 / c)
  uint32_t bpw;       // buckets per window = 2^(c-1)
  uint32_t nb;        // nwin * bpw
  uint32_t task_len;  // T
  uint32_t win_begin, win_end;
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

// signed digit of window w (canonical scalar in v[8]); returns carry for the next window
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
    for (uint32_t w = 0; w < g.nwin; w++) {
      int32_t d = take_digit(v, w, g.c, carry);
      if (d == 0 || w < g.win_begin || w >= g.win_end) continue;
      uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      uint32_t key = w * g.bpw + (mag - 1);
      if (SCATTER) {
        uint32_t pos = atomicAdd(counts_or_cursor + key, 1u);
        entries[pos] = i | (d < 0 ? 0x80000000u : 0u);
      } else {
        atomicAdd(counts_or_cursor + key, 1u);
      }
    }
  }
}

// ---- exclusive scan of (count, ceil(count/T)) pairs over nb buckets, 2048 per block -------------
static constexpr uint32_t SCAN_ITEMS = 8, SCAN_THREADS = 256, SCAN_BLOCK = SCAN_ITEMS * SCAN_THREADS;

__device__ __forceinline__ uint2 add2(uint2 a, uint2 b) { return make_uint2(a.x + b.x, a.y + b.y); }

__device__ uint2 block_exclusive_scan(uint2 v, uint2* total, uint2* sh /*[SCAN_THREADS/32]*/) {
  uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint2 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t x = __shfl_up_sync(0xffffffffu, inc.x, o), y = __shfl_up_sync(0xffffffffu, inc.y, o);
    if (lane >= (uint32_t)o) { inc.x += x; inc.y += y; }
  }
  if (lane == 31) sh[wid] = inc;
  __syncthreads();
  uint2 woff = make_uint2(0, 0), tot = make_uint2(0, 0);
  for (uint32_t k = 0; k < SCAN_THREADS / 32; k++) {
    if (k < wid) woff = add2(woff, sh[k]);
    tot = add2(tot, sh[k]);
  }
  __syncthreads();
  *total = tot;
  return make_uint2(woff.x + inc.x - v.x, woff.y + inc.y - v.y);
}

__global__ void __launch_bounds__(SCAN_THREADS) msm_scan1(const uint32_t* __restrict__ counts, uint32_t nb,
                                                           uint32_t task_len, uint2* block_sums) {
  __shared__ uint2 sh[SCAN_THREADS / 32];
  uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  uint2 acc = make_uint2(0, 0);
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    uint32_t cnt = (base + k < nb) ? counts[base + k] : 0;
    acc.x += cnt;
    acc.y += (cnt + task_len - 1) / task_len;
  }
  uint2 tot;
  block_exclusive_scan(acc, &tot, sh);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) msm_scan2(uint2* block_sums, uint32_t nblocks) {
  __shared__ uint2 sh[SCAN_THREADS / 32];
  uint32_t base = threadIdx.x * SCAN_ITEMS;
  uint2 loc[SCAN_ITEMS];
  uint2 acc = make_uint2(0, 0);
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    loc[k] = (base + k < nblocks) ? block_sums[base + k] : make_uint2(0, 0);
    acc = add2(acc, loc[k]);
  }
  uint2 tot;
  uint2 off = block_exclusive_scan(acc, &tot, sh);
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < nblocks) block_sums[base + k] = off;
    off = add2(off, loc[k]);
  }
}

// offsets[nb+1], cursor[nb] (= offsets, consumed by the scatter), task_off[nb+1]
__global__ void __launch_bounds__(SCAN_THREADS) msm_scan3(const uint32_t* __restrict__ counts, uint32_t nb,
                                                           uint32_t task_len, const uint2* __restrict__ block_sums,
                                                           uint32_t* offsets, uint32_t* cursor, uint32_t* task_off) {
  __shared__ uint2 sh[SCAN_THREADS / 32];
  uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  uint32_t cnt[SCAN_ITEMS];
  uint2 acc = make_uint2(0, 0);
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    cnt[k] = (base + k < nb) ? counts[base + k] : 0;
    acc.x += cnt[k];
    acc.y += (cnt[k] + task_len - 1) / task_len;
  }
  uint2 tot;
  uint2 off = add2(block_exclusive_scan(acc, &tot, sh), block_sums[blockIdx.x]);
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    if (base + k <= nb) {  // also writes the closing element [nb]
      offsets[base + k] = off.x;
      task_off[base + k] = off.y;
      if (base + k < nb) cursor[base + k] = off.x;
    }
    off.x += cnt[k];
    off.y += (cnt[k] + task_len - 1) / task_len;
  }
}

// ---- K1: bucket accumulation --------------------------------------------------------------------
__global__ void __launch_bounds__(MSM_THREADS) msm_accumulate(const uint8_t* __restrict__ bases,
                                                               const uint32_t* __restrict__ entries,
                                                               const uint32_t* __restrict__ offsets,
                                                               const uint32_t* __restrict__ task_off, MsmGeom g,
                                                               uint8_t* __restrict__ partials,
                                                               uint8_t* __restrict__ bucket_sums) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t ntasks = task_off[g.nb];
  if (t >= ntasks) return;
  // largest b with task_off[b] <= t  (buckets without tasks have task_off[b] == task_off[b+1])
  uint32_t lo = 0, hi = g.nb;  // invariant: task_off[lo] <= t < task_off[hi]
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(task_off + mid) <= t) lo = mid; else hi = mid;
  }
  uint32_t b = lo;
  uint32_t t0 = __ldg(task_off + b), t1 = __ldg(task_off + b + 1);
  uint32_t beg = __ldg(offsets + b), end = __ldg(offsets + b + 1);
  uint32_t start = beg + (t - t0) * g.task_len;
  uint32_t stop = min(start + g.task_len, end);

  G1Xyzz acc = G1Xyzz::identity();
  for (uint32_t k = start; k < stop; k++) {
    uint32_t e = __ldg(entries + k);
    G1Affine q = G1Affine::load_nc(bases + (size_t)(e & 0x7fffffffu) * 64);
    if (e >> 31) q.y = fp_neg(q.y);
    xyzz_madd(acc, q);
  }
  if (t1 - t0 == 1) acc.store(bucket_sums + (size_t)b * 128);
  else acc.store(partials + (size_t)t * 128);
}

// buckets with 0 or 2..HOT tasks; hot ones are queued
__global__ void __launch_bounds__(MSM_THREADS) msm_fold(const uint32_t* __restrict__ task_off, MsmGeom g,
                                                         const uint8_t* __restrict__ partials,
                                                         uint8_t* __restrict__ bucket_sums, uint32_t* hot_count,
                                                         uint32_t* hot_list) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= g.nb) return;
  uint32_t t0 = task_off[b], t1 = task_off[b + 1];
  uint32_t nt = t1 - t0;
  if (nt == 1) return;
  if (nt == 0) {
    G1Xyzz::identity().store(bucket_sums + (size_t)b * 128);
    return;
  }
  if (nt > HOT_TASKS) {
    hot_list[atomicAdd(hot_count, 1u)] = b;
    return;
  }
  G1Xyzz acc = G1Xyzz::load(partials + (size_t)t0 * 128);
  for (uint32_t t = t0 + 1; t < t1; t++) xyzz_add(acc, G1Xyzz::load(partials + (size_t)t * 128));
  acc.store(bucket_sums + (size_t)b * 128);
}

__global__ void __launch_bounds__(256) msm_fold_hot(const uint32_t* __restrict__ task_off,
                                                     const uint8_t* __restrict__ partials,
                                                     uint8_t* __restrict__ bucket_sums,
                                                     const uint32_t* __restrict__ hot_count,
                                                     const uint32_t* __restrict__ hot_list) {
  __shared__ uint4 sh[256 * 8];  // one XYZZ point (128 B) per thread
  uint32_t nhot = *hot_count;
  for (uint32_t h = blockIdx.x; h < nhot; h += gridDim.x) {
    uint32_t b = hot_list[h];
    uint32_t t0 = task_off[b], t1 = task_off[b + 1];
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) xyzz_add(acc, G1Xyzz::load(partials + (size_t)t * 128));
    acc.store(sh + threadIdx.x * 8);
    __syncthreads();
    for (uint32_t o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        G1Xyzz a = G1Xyzz::load(sh + threadIdx.x * 8);
        xyzz_add(a, G1Xyzz::load(sh + (threadIdx.x + o) * 8));
        a.store(sh + threadIdx.x * 8);
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) G1Xyzz::load(sh).store(bucket_sums + (size_t)b * 128);
    __syncthreads();
  }
}

// ---- window sums: sum_idx (idx * S_idx + C_idx) by an L-ary tree -------------------------------------
// level input: per window m items (S, C); output ceil(m/L) items.  At level 0, C aliases S
// (bucket idx holds digit idx+1).
__global__ void __launch_bounds__(MSM_THREADS) msm_wsum(const uint8_t* __restrict__ s_in, const uint8_t* __restrict__ c_in,
                                                         uint32_t nwin, uint32_t m, uint8_t* __restrict__ s_out,
                                                         uint8_t* __restrict__ c_out) {
  uint32_t mo = (m + WSUM_L - 1) / WSUM_L;
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nwin * mo) return;
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

// Horner over windows [wb, we), times 2^(c*wb); affine + jacobian out
__global__ void msm_final(const uint8_t* __restrict__ wsum_c, MsmGeom g, uint8_t* out160) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  G1Xyzz acc = G1Xyzz::identity();
  for (int w = (int)g.win_end - 1; w >= (int)g.win_begin; w--) {
    if (!acc.is_identity())
      for (uint32_t k = 0; k < g.c; k++) acc = xyzz_dbl(acc);
    xyzz_add(acc, G1Xyzz::load(wsum_c + (size_t)w * 128));
  }
  if (!acc.is_identity())
    for (uint32_t k = 0; k < g.c * g.win_begin; k++) acc = xyzz_dbl(acc);
  G1Affine a = xyzz_to_affine(acc);
  a.x.store(out160);
  a.y.store(out160 + 32);
  bool id = acc.is_identity();
  a.x.store(out160 + 64);
  (id ? Fq::one() : a.y).store(out160 + 96);
  (id ? Fq::zero() : Fq::one()).store(out160 + 128);
}

__global__ void g1_sum_kernel(const uint8_t* __restrict__ pts96, uint32_t m, uint8_t* out160) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  G1Xyzz acc = G1Xyzz::identity();
  for (uint32_t i = 0; i < m; i++) {
    Fq x = Fq::load(pts96 + (size_t)i * 96), y = Fq::load(pts96 + (size_t)i * 96 + 32),
       z = Fq::load(pts96 + (size_t)i * 96 + 64);
    if (z.is_zero()) continue;
    G1Xyzz p;  // general Jacobian -> XYZZ: (X, Y, Z^2, Z^3)
    p.x = x; p.y = y; p.zz = fp_sqr(z); p.zzz = p.zz * z;
    xyzz_add(acc, p);
  }
  G1Affine a = xyzz_to_affine(acc);
  bool id = acc.is_identity();
  a.x.store(out160);
  a.y.store(out160 + 32);
  a.x.store(out160 + 64);
  (id ? Fq::one() : a.y).store(out160 + 96);
  (id ? Fq::zero() : Fq::one()).store(out160 + 128);
}

int g1_sum_jacobian(h2agg_ctx* ctx, const void* d_points96, size_t m, void* d_out160) {
  g1_sum_kernel<<<1, 32, 0, ctx->stream>>>((const uint8_t*)d_points96, (uint32_t)m, (uint8_t*)d_out160);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int lanes_init(h2agg_ctx* ctx) {
  if (ctx->fork_ev) return 0;
  H2AGG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  for (int i = 0; i < N_LANES; i++) {
    H2AGG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->lanes[i].st, cudaStreamNonBlocking));
    H2AGG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->lanes[i].done, cudaEventDisableTiming));
  }
  return 0;
}

int msm_run_batch(h2agg_ctx* ctx, const void* d_bases, const void* const* cols, size_t n_cols, size_t n,
                  uint8_t* d_out160s, bool host_cols) {
  if (n_cols == 0) return 0;
  int rc;
  if (n_cols == 1 && !host_cols) return msm_run(ctx, ctx->stream, ctx->msm_ws, d_bases, cols[0], n, d_out160s, 0, -1);
  if ((rc = lanes_init(ctx))) return rc;
  H2AGG_CUDA(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));
  for (int l = 0; l < N_LANES; l++) H2AGG_CUDA(ctx, cudaStreamWaitEvent(ctx->lanes[l].st, ctx->fork_ev, 0));
  for (size_t i = 0; i < n_cols; i++) {
    Lane& ln = ctx->lanes[i % N_LANES];
    const void* d_col = cols[i];
    if (host_cols) {
      if ((rc = ensure(ctx, ln.io, n * 32 + 64))) return rc;
      if (n) H2AGG_CUDA(ctx, cudaMemcpyAsync(ln.io.p, cols[i], n * 32, cudaMemcpyHostToDevice, ln.st));
      d_col = ln.io.p;
    }
    if ((rc = msm_run(ctx, ln.st, ln.ws, d_bases, d_col, n, d_out160s + i * 160, 0, -1))) return rc;
  }
  for (int l = 0; l < N_LANES; l++) {
    H2AGG_CUDA(ctx, cudaEventRecord(ctx->lanes[l].done, ctx->lanes[l].st));
    H2AGG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lanes[l].done, 0));
  }
  return 0;
}

int msm_run(h2agg_ctx* ctx, cudaStream_t st, DevBuf& wsbuf, const void* d_bases, const void* d_scalars, size_t n,
            void* d_out160, int win_begin, int win_end) {
  if (n >= (1ull << 31)) {
    ctx->last_error = "msm: n must be < 2^31";
    return 1;
  }
  MsmGeom g;
  int c, nwin;
  msm_window_config(n ? n : 1, ctx->msm_window_bits, &c, &nwin);
  g.n = (uint32_t)n;
  g.c = (uint32_t)c;
  g.nwin = (uint32_t)nwin;
  g.bpw = 1u << (c - 1);
  g.nb = g.nwin * g.bpw;
  g.win_begin = win_begin < 0 ? 0 : (uint32_t)win_begin;
  g.win_end = (win_end < 0 || win_end > nwin) ? (uint32_t)nwin : (uint32_t)win_end;
  if (g.win_begin > g.win_end) g.win_begin = g.win_end;
  {
    size_t avg = n / g.bpw;
    uint32_t T = 32;
    while (T < 2 * avg && T < 512) T <<= 1;
    g.task_len = T;
  }
  const size_t max_entries = (size_t)n * (g.win_end - g.win_begin);
  const size_t max_tasks = max_entries / g.task_len + g.nb + 1;
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
  size_t o_taskoff = carve((size_t)(g.nb + 1) * 4);
  size_t o_bsums = carve((size_t)scan_blocks * 8);
  size_t o_hot = carve((size_t)(g.nb + 1) * 4 + 256);
  size_t o_entries = carve((max_entries + 1) * 4);
  size_t o_partials = carve(max_tasks * 128);
  size_t o_buckets = carve((size_t)g.nb * 128);
  size_t lvl_items = (size_t)g.nwin * ((g.bpw + WSUM_L - 1) / WSUM_L);
  size_t o_lvl_s0 = carve(lvl_items * 128), o_lvl_c0 = carve(lvl_items * 128);
  size_t o_lvl_s1 = carve(lvl_items * 128), o_lvl_c1 = carve(lvl_items * 128);
  int rc = ensure(ctx, wsbuf, off);
  if (rc) return rc;
  uint8_t* ws = (uint8_t*)wsbuf.p;
  uint32_t* counts = (uint32_t*)(ws + o_counts);
  uint32_t* offsets = (uint32_t*)(ws + o_offsets);
  uint32_t* cursor = (uint32_t*)(ws + o_cursor);
  uint32_t* task_off = (uint32_t*)(ws + o_taskoff);
  uint2* bsums = (uint2*)(ws + o_bsums);
  uint32_t* hot_count = (uint32_t*)(ws + o_hot);
  uint32_t* hot_list = hot_count + 64;
  uint32_t* entries = (uint32_t*)(ws + o_entries);
  uint8_t* partials = ws + o_partials;
  uint8_t* buckets = ws + o_buckets;

  ScopedKernelTimer t_total(ctx, KC_MSM_TOTAL, st);
  H2AGG_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)(g.nb + 1) * 4, st));
  H2AGG_CUDA(ctx, cudaMemsetAsync(hot_count, 0, 256, st));
  if (n) {
    uint32_t grid = (uint32_t)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
    ScopedKernelTimer tk(ctx, KC_MSM_DIGITS, st);
    msm_digits<false><<<grid, 256, 0, st>>>((const uint4*)d_scalars, g, counts, nullptr);
    ctx->launches++;
  }
  msm_scan1<<<scan_blocks, SCAN_THREADS, 0, st>>>(counts, g.nb, g.task_len, bsums);
  msm_scan2<<<1, SCAN_THREADS, 0, st>>>(bsums, scan_blocks);
  msm_scan3<<<scan_blocks, SCAN_THREADS, 0, st>>>(counts, g.nb, g.task_len, bsums, offsets, cursor, task_off);
  ctx->launches += 3;
  if (n) {
    uint32_t grid = (uint32_t)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
    {
      ScopedKernelTimer tk(ctx, KC_MSM_DIGITS, st);
      msm_digits<true><<<grid, 256, 0, st>>>((const uint4*)d_scalars, g, cursor, entries);
      ctx->launches++;
    }
    ScopedKernelTimer tk(ctx, KC_MSM_ACCUMULATE, st);
    msm_accumulate<<<(uint32_t)((max_tasks + MSM_THREADS - 1) / MSM_THREADS), MSM_THREADS, 0, st>>>(
        (const uint8_t*)d_bases, entries, offsets, task_off, g, partials, buckets);
    ctx->launches++;
  }
  ScopedKernelTimer t_red(ctx, KC_MSM_REDUCE, st);
  msm_fold<<<(g.nb + MSM_THREADS - 1) / MSM_THREADS, MSM_THREADS, 0, st>>>(task_off, g, partials, buckets, hot_count,
                                                                            hot_list);
  msm_fold_hot<<<ctx->sm_count * 2, 256, 0, st>>>(task_off, partials, buckets, hot_count, hot_list);
  ctx->launches += 2;
  H2AGG_CUDA(ctx, cudaGetLastError());

  // window sums
  const uint8_t *s_in = buckets, *c_in = buckets;
  uint8_t* lvl_s[2] = {ws + o_lvl_s0, ws + o_lvl_s1};
  uint8_t* lvl_c[2] = {ws + o_lvl_c0, ws + o_lvl_c1};
  uint32_t m = g.bpw;
  int flip = 0;
  // at least one level so that the final C holds sum (idx+1) * B_idx
  do {
    uint32_t mo = (m + WSUM_L - 1) / WSUM_L;
    uint32_t total = g.nwin * mo;
    msm_wsum<<<(total + MSM_THREADS - 1) / MSM_THREADS, MSM_THREADS, 0, st>>>(s_in, c_in, g.nwin, m, lvl_s[flip],
                                                                              lvl_c[flip]);
    ctx->launches++;
    s_in = lvl_s[flip];
    c_in = lvl_c[flip];
    flip ^= 1;
    m = mo;
  } while (m > 1);
  msm_final<<<1, 32, 0, st>>>(c_in, g, (uint8_t*)d_out160);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace h2agg
