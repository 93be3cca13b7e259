// N3 ("next" row), sorting half: the lookup argument's `permute_expression_pair`.
//
// halo2_proofs plonk/lookup/prover.rs (external crate; reached from create_proof,
// halo2-snark-aggregator-circuit/src/verify_circuit.rs:986; the verifier's view of the result is
// halo2-snark-aggregator-api/src/systems/halo2/lookup.rs:58-119) does, per lookup, over the usable rows:
//     a' = sort(input)                                   Fr::cmp = numeric order of the CANONICAL value
//     s'[row] = a'[row] where a' changes value, and one instance of that value leaves the table multiset
//     the remaining table values, ascending, fill the rows of repeated inputs taken from the BACK
// and fails if an input value is not in the table.  The CPU code is a std sort, a BTreeMap and a serial loop over
// 2^k rows, 7 times per proof.  Here every step is a data-parallel pass:
//     canonicalise -> LSD radix sort (8-bit digits; digits on which all keys agree are skipped: witness columns
//     are 17-bit values, so 3 of 32 passes run) of input and table -> "repeated row" flags + scan ->
//     leftover flags (first table occurrence of a value that occurs in a' is consumed; membership by binary
//     search) + scan -> scatter leftover q to repeated row m-1-q -> back to Montgomery form.
// All keys are moved as two 16-byte vectors; a radix pass reads the keys twice and writes them once.
#include "../../include/h2agg.h"
#include "bn254_field.cuh"
#include "ctx.hpp"
#include <algorithm>
#include <cstring>
#include <string>

namespace h2agg {

static constexpr uint32_t SORT_THREADS = 256, SORT_ROUNDS = 8, SORT_TILE = SORT_THREADS * SORT_ROUNDS;

struct Key {
  uint32_t v[8];
};
__device__ __forceinline__ Key key_load(const uint4* p, size_t i) {
  uint4 a = p[2 * i], b = p[2 * i + 1];
  Key k;
  k.v[0] = a.x; k.v[1] = a.y; k.v[2] = a.z; k.v[3] = a.w;
  k.v[4] = b.x; k.v[5] = b.y; k.v[6] = b.z; k.v[7] = b.w;
  return k;
}
__device__ __forceinline__ void key_store(uint4* p, size_t i, const Key& k) {
  p[2 * i] = make_uint4(k.v[0], k.v[1], k.v[2], k.v[3]);
  p[2 * i + 1] = make_uint4(k.v[4], k.v[5], k.v[6], k.v[7]);
}
__device__ __forceinline__ uint32_t key_byte(const Key& k, uint32_t byte) { return (k.v[byte >> 2] >> ((byte & 3) * 8)) & 0xffu; }
// -1 / 0 / +1 as 256-bit little-endian integers
__device__ __forceinline__ int key_cmp(const Key& a, const Key& b) {
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    if (a.v[i] != b.v[i]) return a.v[i] < b.v[i] ? -1 : 1;
  }
  return 0;
}

// Montgomery -> canonical, plus the bitwise OR / AND of all keys (a byte position where they agree needs no pass)
__global__ void __launch_bounds__(256) sort_canon(const Fr* __restrict__ in, uint4* __restrict__ out, size_t n,
                                                  uint32_t* __restrict__ or_and /* [8] OR, [8] AND */) {
  uint32_t o[8], a[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { o[j] = 0; a[j] = 0xffffffffu; }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr c = fp_from_mont(Fr::load_nc(in + i));
    Key k;
#pragma unroll
    for (int j = 0; j < 8; j++) { k.v[j] = c.v[j]; o[j] |= c.v[j]; a[j] &= c.v[j]; }
    key_store(out, i, k);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    o[j] = __reduce_or_sync(0xffffffffu, o[j]);
    a[j] = __reduce_and_sync(0xffffffffu, a[j]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int j = 0; j < 8; j++) { atomicOr(or_and + j, o[j]); atomicAnd(or_and + 8 + j, a[j]); }
  }
}

__global__ void __launch_bounds__(256) sort_to_mont(const uint4* __restrict__ in, Fr* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Key k = key_load(in, i);
  Fr c;
#pragma unroll
  for (int j = 0; j < 8; j++) c.v[j] = k.v[j];
  fp_to_mont(c).store(out + i);
}

// per-tile histogram of one byte position, bin-major: hist[bin * ntiles + tile]
__global__ void __launch_bounds__(SORT_THREADS) sort_hist(const uint4* __restrict__ keys, uint32_t n, uint32_t byte,
                                                           uint32_t* __restrict__ hist, uint32_t ntiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * SORT_TILE;
  const uint32_t limb = byte >> 2, sh = (byte & 3) * 8;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(keys);
#pragma unroll
  for (uint32_t r = 0; r < SORT_ROUNDS; r++) {
    uint32_t p = base + r * SORT_THREADS + threadIdx.x;
    if (p < n) atomicAdd(&h[(__ldg(words + (size_t)p * 8 + limb) >> sh) & 0xffu], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// stable scatter of one tile: local order = (round, warp, lane) = key position
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter(const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t n,
                                                              uint32_t byte, const uint32_t* __restrict__ offsets,
                                                              uint32_t ntiles) {
  __shared__ uint32_t warp_cnt[SORT_THREADS / 32][256];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t run = offsets[tid * ntiles + blockIdx.x];  // thread t owns digit t's running output position
  const uint32_t base = blockIdx.x * SORT_TILE;
  for (uint32_t r = 0; r < SORT_ROUNDS; r++) {
#pragma unroll
    for (uint32_t w = 0; w < SORT_THREADS / 32; w++) warp_cnt[w][tid] = 0;
    __syncthreads();
    const uint32_t p = base + r * SORT_THREADS + tid;
    const bool valid = p < n;
    Key k;
    uint32_t digit = 256;  // out-of-range lanes match among themselves and write nothing
    if (valid) {
      k = key_load(in, p);
      digit = key_byte(k, byte);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1));
    if (valid && rank == 0) warp_cnt[warp][digit] = __popc(peers);
    __syncthreads();
#pragma unroll
    for (uint32_t w = 0; w < SORT_THREADS / 32; w++) {
      uint32_t c = warp_cnt[w][tid];
      warp_cnt[w][tid] = run;
      run += c;
    }
    __syncthreads();
    if (valid) key_store(out, warp_cnt[warp][digit] + rank, k);
    __syncthreads();
  }
}

// ---- exclusive scan of n uint32 (n <= 2048 * 2048): out[i] = sum_{j<i} in[j], out[n] = total ------------------
static constexpr uint32_t XS_ITEMS = 8, XS_THREADS = 256, XS_BLOCK = XS_ITEMS * XS_THREADS;

__device__ __forceinline__ uint32_t xs_block_scan(uint32_t v, uint32_t* total, uint32_t* sh) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += x;
  }
  if (lane == 31) sh[wid] = inc;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
  for (uint32_t w = 0; w < XS_THREADS / 32; w++) {
    if (w < wid) woff += sh[w];
    tot += sh[w];
  }
  __syncthreads();
  *total = tot;
  return woff + inc - v;
}
__global__ void __launch_bounds__(XS_THREADS) xs_sums(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t sh[XS_THREADS / 32];
  const uint32_t base = blockIdx.x * XS_BLOCK + threadIdx.x * XS_ITEMS;
  uint32_t acc = 0;
#pragma unroll
  for (uint32_t j = 0; j < XS_ITEMS; j++) acc += (base + j < n) ? in[base + j] : 0;
  uint32_t tot;
  xs_block_scan(acc, &tot, sh);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(XS_THREADS) xs_top(uint32_t* block_sums, uint32_t nblocks) {
  __shared__ uint32_t sh[XS_THREADS / 32];
  const uint32_t base = threadIdx.x * XS_ITEMS;
  uint32_t loc[XS_ITEMS], acc = 0;
#pragma unroll
  for (uint32_t j = 0; j < XS_ITEMS; j++) {
    loc[j] = (base + j < nblocks) ? block_sums[base + j] : 0;
    acc += loc[j];
  }
  uint32_t tot;
  uint32_t off = xs_block_scan(acc, &tot, sh);
#pragma unroll
  for (uint32_t j = 0; j < XS_ITEMS; j++) {
    if (base + j < nblocks) block_sums[base + j] = off;
    off += loc[j];
  }
}
// `in` and `out` may be the same array (every thread reads its own items before it writes them): no __restrict__
__global__ void __launch_bounds__(XS_THREADS) xs_apply(const uint32_t* in, uint32_t n, const uint32_t* __restrict__ block_sums,
                                                        uint32_t* out) {
  __shared__ uint32_t sh[XS_THREADS / 32];
  const uint32_t base = blockIdx.x * XS_BLOCK + threadIdx.x * XS_ITEMS;
  uint32_t loc[XS_ITEMS], acc = 0;
#pragma unroll
  for (uint32_t j = 0; j < XS_ITEMS; j++) {
    loc[j] = (base + j < n) ? in[base + j] : 0;
    acc += loc[j];
  }
  uint32_t tot;
  uint32_t off = xs_block_scan(acc, &tot, sh) + block_sums[blockIdx.x];
#pragma unroll
  for (uint32_t j = 0; j < XS_ITEMS; j++) {
    if (base + j <= n) out[base + j] = off;  // also the closing element out[n] = total
    off += loc[j];
  }
}
// in and out may alias
static int exclusive_scan_u32(h2agg_ctx* ctx, cudaStream_t st, const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* block_sums) {
  const uint32_t nblocks = (n + 1 + XS_BLOCK - 1) / XS_BLOCK;
  if (nblocks > XS_BLOCK) {
    ctx->last_error = "sort: scan longer than 2048 * 2048 elements";
    return 1;
  }
  xs_sums<<<nblocks, XS_THREADS, 0, st>>>(in, n, block_sums);
  xs_top<<<1, XS_THREADS, 0, st>>>(block_sums, nblocks);
  xs_apply<<<nblocks, XS_THREADS, 0, st>>>(in, n, block_sums, out);
  ctx->launches += 3;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ---- the permutation itself (all keys canonical, both arrays sorted ascending) -----------------------------------
// does `key` occur in sorted[0..n)?
__device__ __forceinline__ bool sorted_contains(const uint4* __restrict__ sorted, uint32_t n, const Key& key) {
  uint32_t lo = 0, hi = n;  // first index with sorted[idx] >= key
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (key_cmp(key_load(sorted, mid), key) < 0) lo = mid + 1; else hi = mid;
  }
  return lo < n && key_cmp(key_load(sorted, lo), key) == 0;
}

// rows of a': repeated[r] = 1 when a'[r] == a'[r-1]; first occurrences copy their value into s' and must be in the table
__global__ void __launch_bounds__(256) lookup_mark_rows(const uint4* __restrict__ a_sorted, const uint4* __restrict__ s_sorted,
                                                        uint32_t u, uint32_t* __restrict__ repeated, uint4* __restrict__ s_out,
                                                        uint32_t* __restrict__ missing) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= u) return;
  Key k = key_load(a_sorted, r);
  bool rep = r > 0 && key_cmp(key_load(a_sorted, r - 1), k) == 0;
  repeated[r] = rep ? 1u : 0u;
  if (!rep) {
    key_store(s_out, r, k);
    if (!sorted_contains(s_sorted, u, k)) atomicAdd(missing, 1u);
  }
}
// rep_rows[i] = the i-th repeated row (ascending)
__global__ void __launch_bounds__(256) lookup_collect_rows(const uint4* __restrict__ a_sorted, uint32_t u,
                                                           const uint32_t* __restrict__ rep_index, uint32_t* __restrict__ rep_rows) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= u) return;
  if (rep_index[r + 1] != rep_index[r]) rep_rows[rep_index[r]] = r;
}
// table entry j is consumed by a first occurrence iff it is the first of its value in the sorted table and the value is in a'
__global__ void __launch_bounds__(256) lookup_mark_leftover(const uint4* __restrict__ s_sorted, const uint4* __restrict__ a_sorted,
                                                            uint32_t u, uint32_t* __restrict__ leftover) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= u) return;
  Key k = key_load(s_sorted, j);
  bool first = j == 0 || key_cmp(key_load(s_sorted, j - 1), k) != 0;
  leftover[j] = (first && sorted_contains(a_sorted, u, k)) ? 0u : 1u;
}
// leftover q (ascending) -> repeated row m - 1 - q   (Vec::pop from the back)
__global__ void __launch_bounds__(256) lookup_place_leftover(const uint4* __restrict__ s_sorted, uint32_t u,
                                                             const uint32_t* __restrict__ left_index,
                                                             const uint32_t* __restrict__ rep_index,
                                                             const uint32_t* __restrict__ rep_rows, uint4* __restrict__ s_out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= u) return;
  if (left_index[j + 1] == left_index[j]) return;
  const uint32_t q = left_index[j], m = rep_index[u];
  if (q >= m) return;  // only when an input value is missing from the table (reported through `missing`)
  key_store(s_out, rep_rows[m - 1 - q], key_load(s_sorted, j));
}

// sort n canonical keys held in buf[0] (ping-pong with buf[1]); returns which buffer holds the result
static int radix_sort_canonical(h2agg_ctx* ctx, cudaStream_t st, uint4* buf[2], uint32_t n, const uint32_t or_and[16],
                                uint32_t* hist, uint32_t* block_sums, int* result_in) {
  const uint32_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
  int cur = 0;
  for (uint32_t byte = 0; byte < 32; byte++) {
    const uint32_t sh = (byte & 3) * 8;
    if ((((or_and[byte >> 2] ^ or_and[8 + (byte >> 2)]) >> sh) & 0xffu) == 0) continue;  // every key has the same digit
    sort_hist<<<ntiles, SORT_THREADS, 0, st>>>(buf[cur], n, byte, hist, ntiles);
    ctx->launches++;
    int rc = exclusive_scan_u32(ctx, st, hist, 256 * ntiles, hist, block_sums);
    if (rc) return rc;
    sort_scatter<<<ntiles, SORT_THREADS, 0, st>>>(buf[cur], buf[cur ^ 1], n, byte, hist, ntiles);
    ctx->launches++;
    cur ^= 1;
  }
  H2AGG_CUDA(ctx, cudaGetLastError());
  *result_in = cur;
  return 0;
}

struct SortWs {
  uint4* keys[2][2];   // [array][ping-pong]
  uint32_t *hist, *block_sums, *flags_a, *flags_s, *rep_rows, *small;  // small: [0..15] OR/AND array 0, [16..31] array 1, [32] missing
  uint4* s_out;
};

static int carve_ws(h2agg_ctx* ctx, size_t n, int n_arrays, bool lookup, SortWs* w) {
  const size_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
  size_t o_keys[2][2];
  for (int a = 0; a < n_arrays; a++)
    for (int b = 0; b < 2; b++) o_keys[a][b] = carve(n * 32);
  size_t o_hist = carve((256 * ntiles + 1) * 4);
  size_t o_bs = carve((XS_BLOCK + 1) * 4);
  size_t o_small = carve(256);
  size_t o_fa = 0, o_fs = 0, o_rr = 0, o_so = 0;
  if (lookup) {
    o_fa = carve((n + 1) * 4);
    o_fs = carve((n + 1) * 4);
    o_rr = carve((n + 1) * 4);
    o_so = carve(n * 32);
  }
  int rc = ensure(ctx, ctx->sort_ws, off);
  if (rc) return rc;
  uint8_t* p = (uint8_t*)ctx->sort_ws.p;
  for (int a = 0; a < n_arrays; a++)
    for (int b = 0; b < 2; b++) w->keys[a][b] = (uint4*)(p + o_keys[a][b]);
  w->hist = (uint32_t*)(p + o_hist);
  w->block_sums = (uint32_t*)(p + o_bs);
  w->small = (uint32_t*)(p + o_small);
  w->flags_a = (uint32_t*)(p + o_fa);
  w->flags_s = (uint32_t*)(p + o_fs);
  w->rep_rows = (uint32_t*)(p + o_rr);
  w->s_out = (uint4*)(p + o_so);
  return 0;
}

// canonicalise `n_arrays` device arrays and fetch their OR/AND masks (one small D2H + sync: the host picks the passes)
static int canon_and_masks(h2agg_ctx* ctx, cudaStream_t st, const void* const* d_in, int n_arrays, size_t n, SortWs& w,
                           uint32_t masks[2][16]) {
  static const uint32_t init[32] = {0, 0, 0, 0, 0, 0, 0, 0, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u,
                                    0, 0, 0, 0, 0, 0, 0, 0, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u};
  memcpy(ctx->pinned, init, sizeof(init));
  memset((uint8_t*)ctx->pinned + 128, 0, 4);
  H2AGG_CUDA(ctx, cudaMemcpyAsync(w.small, ctx->pinned, 132, cudaMemcpyHostToDevice, st));
  const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
  for (int a = 0; a < n_arrays; a++) {
    sort_canon<<<grid, 256, 0, st>>>((const Fr*)d_in[a], w.keys[a][0], n, w.small + 16 * a);
    ctx->launches++;
  }
  H2AGG_CUDA(ctx, cudaGetLastError());
  H2AGG_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ctx->pinned + 256, w.small, 128, cudaMemcpyDeviceToHost, st));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(masks, (uint8_t*)ctx->pinned + 256, 128);
  return 0;
}

static int sort_fr_dev(h2agg_ctx* ctx, void* d_a, size_t n) {
  if (n >= (1ull << 31)) { ctx->last_error = "sort_fr: n must be < 2^31"; return 1; }
  if (n <= 1) return 0;
  cudaStream_t st = ctx->stream;
  SortWs w;
  int rc = carve_ws(ctx, n, 1, false, &w);
  if (rc) return rc;
  uint32_t masks[2][16];
  const void* in[1] = {d_a};
  if ((rc = canon_and_masks(ctx, st, in, 1, n, w, masks))) return rc;
  int res;
  if ((rc = radix_sort_canonical(ctx, st, w.keys[0], (uint32_t)n, masks[0], w.hist, w.block_sums, &res))) return rc;
  sort_to_mont<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.keys[0][res], (Fr*)d_a, n);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

// returns 0, or 4 when an input value does not occur in the table (outputs are then unspecified)
static int permute_expression_pair_dev(h2agg_ctx* ctx, const void* d_input, const void* d_table, size_t u, void* d_pin,
                                       void* d_ptab) {
  if (u >= (1ull << 31)) { ctx->last_error = "permute_expression_pair: too many rows"; return 1; }
  if (u == 0) return 0;
  cudaStream_t st = ctx->stream;
  SortWs w;
  int rc = carve_ws(ctx, u, 2, true, &w);
  if (rc) return rc;
  uint32_t masks[2][16];
  const void* in[2] = {d_input, d_table};
  if ((rc = canon_and_masks(ctx, st, in, 2, u, w, masks))) return rc;
  int ra, rs;
  if ((rc = radix_sort_canonical(ctx, st, w.keys[0], (uint32_t)u, masks[0], w.hist, w.block_sums, &ra))) return rc;
  if ((rc = radix_sort_canonical(ctx, st, w.keys[1], (uint32_t)u, masks[1], w.hist, w.block_sums, &rs))) return rc;
  const uint4* a_sorted = w.keys[0][ra];
  const uint4* s_sorted = w.keys[1][rs];
  const uint32_t n32 = (uint32_t)u;
  const unsigned grid = (unsigned)((u + 255) / 256);
  uint32_t* missing = w.small + 32;
  lookup_mark_rows<<<grid, 256, 0, st>>>(a_sorted, s_sorted, n32, w.flags_a, w.s_out, missing);
  lookup_mark_leftover<<<grid, 256, 0, st>>>(s_sorted, a_sorted, n32, w.flags_s);
  ctx->launches += 2;
  if ((rc = exclusive_scan_u32(ctx, st, w.flags_a, n32, w.flags_a, w.block_sums))) return rc;
  if ((rc = exclusive_scan_u32(ctx, st, w.flags_s, n32, w.flags_s, w.block_sums))) return rc;
  lookup_collect_rows<<<grid, 256, 0, st>>>(a_sorted, n32, w.flags_a, w.rep_rows);
  lookup_place_leftover<<<grid, 256, 0, st>>>(s_sorted, n32, w.flags_s, w.flags_a, w.rep_rows, w.s_out);
  sort_to_mont<<<grid, 256, 0, st>>>(a_sorted, (Fr*)d_pin, u);
  sort_to_mont<<<grid, 256, 0, st>>>(w.s_out, (Fr*)d_ptab, u);
  ctx->launches += 4;
  H2AGG_CUDA(ctx, cudaGetLastError());
  H2AGG_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ctx->pinned + 512, missing, 4, cudaMemcpyDeviceToHost, st));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(st));
  uint32_t miss;
  memcpy(&miss, (uint8_t*)ctx->pinned + 512, 4);
  if (miss) {
    ctx->last_error = "permute_expression_pair: " + std::to_string(miss) +
                      " distinct input value(s) do not occur in the table (halo2: Error::ConstraintSystemFailure)";
    return 4;
  }
  return 0;
}

}  // namespace h2agg

using namespace h2agg;

extern "C" {

int h2agg_sort_fr_dev(h2agg_ctx* ctx, void* d_a, size_t n) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_a && n) { ctx->last_error = "sort_fr: null argument"; return 1; }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  return sort_fr_dev(ctx, d_a, n);
}

int h2agg_sort_fr(h2agg_ctx* ctx, uint64_t* a, size_t n) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!a && n) { ctx->last_error = "sort_fr: null argument"; return 1; }
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure(ctx, ctx->io_a, n * 32);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = sort_fr_dev(ctx, ctx->io_a.p, n))) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(a, ctx->io_a.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int h2agg_permute_expression_pair_dev(h2agg_ctx* ctx, const void* d_input, const void* d_table, size_t usable_rows,
                                      void* d_permuted_input, void* d_permuted_table) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (usable_rows && (!d_input || !d_table || !d_permuted_input || !d_permuted_table)) {
    ctx->last_error = "permute_expression_pair: null argument";
    return 1;
  }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  return permute_expression_pair_dev(ctx, d_input, d_table, usable_rows, d_permuted_input, d_permuted_table);
}

int h2agg_permute_expression_pair(h2agg_ctx* ctx, const uint64_t* input, const uint64_t* table, size_t usable_rows,
                                  uint64_t* permuted_input, uint64_t* permuted_table) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (usable_rows && (!input || !table || !permuted_input || !permuted_table)) {
    ctx->last_error = "permute_expression_pair: null argument";
    return 1;
  }
  if (usable_rows == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t u = usable_rows;
  int rc = ensure(ctx, ctx->io_a, u * 64);
  if (rc) return rc;
  uint8_t* d = (uint8_t*)ctx->io_a.p;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(d, input, u * 32, cudaMemcpyHostToDevice, ctx->stream));
  H2AGG_CUDA(ctx, cudaMemcpyAsync(d + u * 32, table, u * 32, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = permute_expression_pair_dev(ctx, d, d + u * 32, u, d, d + u * 32))) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(permuted_input, d, u * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaMemcpyAsync(permuted_table, d + u * 32, u * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
