// N3 ("next" row): the two scan primitives behind halo2's grand products.
//
//   batch_invert(a)            a_i <- a_i^-1, zeros left at zero          (halo2_proofs `BatchInvert`)
//   grand_product(num, den)    z_0 = 1, z_{i+1} = z_i * num_i / den_i     (the running products of the
//                              permutation argument and of the 7 lookup arguments: per row the prover forms
//                              a numerator and a denominator from the column values and beta/gamma, batch-inverts
//                              the denominators and takes the running product -- plonk/permutation/prover.rs,
//                              plonk/lookup/prover.rs of the external crate; reference
//                              call site halo2-snark-aggregator-circuit/src/verify_circuit.rs:986)
//
// Both are cut into chunks of L = 64: a forward sweep leaves every chunk's running products, the vector of
// chunk totals is handled by the same routine one level up (4M -> 64K -> 1K -> 16 -> one thread), and a
// backward sweep finishes each chunk from its carry.  One Fermat inversion per call; ~3 products per element
// for the inversion, ~3 for the product scan.
#include "../../include/h2agg.h"
#include "bn254_field.cuh"
#include "ctx.hpp"
#include <cstring>

namespace h2agg {

static constexpr uint32_t SCAN_L = 64;

// forward: prefix[j] = product of the non-zero a[lo..j-1] of the chunk (prefix[lo] = 1); total[c] = product of the chunk
__global__ void __launch_bounds__(128) inv_forward(const Fr* __restrict__ a, size_t n, Fr* __restrict__ prefix,
                                                    Fr* __restrict__ total, size_t m) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  size_t lo = c * SCAN_L, hi = min(lo + (size_t)SCAN_L, n);
  Fr run = Fr::one();
  for (size_t i = lo; i < hi; i++) {
    run.store(prefix + i);
    Fr v = Fr::load_nc(a + i);
    if (!v.is_zero()) run = run * v;
  }
  run.store(total + c);
}

// backward: given inv_total[c] = 1 / total[c], out[j] = 1 / a[j] (0 stays 0)
__global__ void __launch_bounds__(128) inv_backward(const Fr* __restrict__ a, size_t n, const Fr* __restrict__ prefix,
                                                     const Fr* __restrict__ inv_total, Fr* __restrict__ out, size_t m) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  size_t lo = c * SCAN_L, hi = min(lo + (size_t)SCAN_L, n);
  Fr inv = Fr::load(inv_total + c);
  for (size_t i = hi; i-- > lo;) {
    Fr v = Fr::load_nc(a + i);
    Fr pre = Fr::load(prefix + i);
    if (v.is_zero()) {
      Fr::zero().store(out + i);
    } else {
      (inv * pre).store(out + i);
      inv = inv * v;
    }
  }
}

// top of the recursion: at most SCAN_L elements, one thread, one inversion
__global__ void inv_serial(const Fr* __restrict__ a, size_t n, Fr* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  Fr pre[SCAN_L];
  Fr run = Fr::one();
  for (size_t i = 0; i < n; i++) {
    pre[i] = run;
    Fr v = Fr::load(a + i);
    if (!v.is_zero()) run = run * v;
  }
  Fr inv = fp_inv(run);
  for (size_t i = n; i-- > 0;) {
    Fr v = Fr::load(a + i);
    if (v.is_zero()) {
      Fr::zero().store(out + i);
    } else {
      (inv * pre[i]).store(out + i);
      inv = inv * v;
    }
  }
}

// ratio and chunk totals of the product scan: r[i] = num[i] * dinv[i]; total[c] = prod r over the chunk
__global__ void __launch_bounds__(128) prod_forward(const Fr* __restrict__ num, const Fr* __restrict__ dinv, size_t n,
                                                     Fr* __restrict__ r, Fr* __restrict__ total, size_t m) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  size_t lo = c * SCAN_L, hi = min(lo + (size_t)SCAN_L, n);
  Fr run = Fr::one();
  for (size_t i = lo; i < hi; i++) {
    Fr v = Fr::load_nc(num + i);
    if (dinv) {
      v = v * Fr::load_nc(dinv + i);
      v.store(r + i);
    }
    run = run * v;
  }
  run.store(total + c);
}

// exclusive prefix products of a short vector (one thread): out[i] = prod_{j<i} a[j]
__global__ void prod_exclusive_serial(const Fr* __restrict__ a, size_t n, Fr* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  Fr run = Fr::one();
  for (size_t i = 0; i < n; i++) {
    Fr v = Fr::load(a + i);
    run.store(out + i);
    run = run * v;
  }
}

// out[i] = carry[c] * prod_{lo <= j < i} r[j]   (exclusive running product seeded by the chunk's carry)
__global__ void __launch_bounds__(128) prod_backfill(const Fr* __restrict__ r, size_t n, const Fr* __restrict__ carry,
                                                      Fr* __restrict__ out, size_t m) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  size_t lo = c * SCAN_L, hi = min(lo + (size_t)SCAN_L, n);
  Fr run = Fr::load(carry + c);
  for (size_t i = lo; i < hi; i++) {
    Fr v = Fr::load_nc(r + i);
    run.store(out + i);
    run = run * v;
  }
}

struct ScanWs {
  Fr* a;         // n: prefix products (inversion) / ratios (product scan)
  Fr* b;         // n: inverted denominators
  Fr* lvl[8];    // chunk totals per level
  Fr* aux[8];    // per level: prefix scratch (inversion) or exclusive carries (product scan)
  Fr* aux2[8];   // per level: inverted totals / carries
  Fr* spare;     // 64 elements
  size_t len[9];
  int levels;
};

static int scan_ws(h2agg_ctx* ctx, size_t n, ScanWs* w, DevBuf* buf_in = nullptr) {
  DevBuf& buf = buf_in ? *buf_in : ctx->scan_ws;
  size_t total = 2 * n + 64;
  size_t m = n;
  w->levels = 0;
  w->len[0] = n;
  while (m > SCAN_L && w->levels < 8) {
    m = (m + SCAN_L - 1) / SCAN_L;
    w->len[++w->levels] = m;
    total += 3 * m;
  }
  int rc = ensure(ctx, buf, total * 32 + 256);
  if (rc) return rc;
  Fr* p = (Fr*)buf.p;
  w->a = p; p += n;
  w->b = p; p += n;
  w->spare = p; p += 64;
  for (int i = 0; i < w->levels; i++) {
    size_t sz = w->len[i + 1];
    w->lvl[i] = p; p += sz;
    w->aux[i] = p; p += sz;
    w->aux2[i] = p; p += sz;
  }
  return 0;
}

// out[i] = 1 / a[i] (zeros stay zero); `prefix0` is n elements of scratch for level 0; out may alias a
static int batch_invert_levels(h2agg_ctx* ctx, const ScanWs& w, const Fr* a, size_t n, Fr* prefix0, Fr* out,
                               cudaStream_t st_in = nullptr) {
  cudaStream_t st = st_in ? st_in : ctx->stream;
  if (w.levels == 0) {
    inv_serial<<<1, 32, 0, st>>>(a, n, out);
    ctx->launches++;
    return 0;
  }
  const Fr* src = a;
  Fr* prefix = prefix0;
  for (int l = 0; l < w.levels; l++) {
    size_t m = w.len[l + 1];
    inv_forward<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(src, w.len[l], prefix, w.lvl[l], m);
    ctx->launches++;
    src = w.lvl[l];
    prefix = (l + 1 < w.levels) ? w.aux[l] : nullptr;  // prefixes of the totals of level l live in aux[l]
  }
  // top: invert the shortest vector of totals into aux2[levels-1]
  inv_serial<<<1, 32, 0, st>>>(w.lvl[w.levels - 1], w.len[w.levels], w.aux2[w.levels - 1]);
  ctx->launches++;
  for (int l = w.levels - 1; l >= 0; l--) {
    const Fr* lsrc = (l == 0) ? a : w.lvl[l - 1];
    const Fr* lpre = (l == 0) ? prefix0 : w.aux[l - 1];
    Fr* lout = (l == 0) ? out : w.aux2[l - 1];
    size_t m = w.len[l + 1];
    inv_backward<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(lsrc, w.len[l], lpre, w.aux2[l], lout, m);
    ctx->launches++;
  }
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

int batch_invert_dev(h2agg_ctx* ctx, void* d_a, size_t n) {
  if (n == 0) return 0;
  ScanWs w;
  int rc = scan_ws(ctx, n, &w);
  if (rc) return rc;
  return batch_invert_levels(ctx, w, (const Fr*)d_a, n, w.a, (Fr*)d_a);
}

// z[0] = 1, z[i+1] = z[i] * num[i] / den[i]
// `st` / `ws`: run on another stream with its own scratch (lane-parallel batches); default = the context's
int grand_product_dev(h2agg_ctx* ctx, const void* d_num, const void* d_den, size_t n, void* d_z, cudaStream_t st_in, DevBuf* ws) {
  if (n == 0) return 0;
  ScanWs w;
  int rc = scan_ws(ctx, n, &w, ws);
  if (rc) return rc;
  cudaStream_t st = st_in ? st_in : ctx->stream;
  // 1. inverted denominators -> w.b (prefix scratch in w.a)
  rc = batch_invert_levels(ctx, w, (const Fr*)d_den, n, w.a, w.b, st);
  if (rc) return rc;
  // 2. ratios -> w.a, chunk totals up the levels
  if (w.levels == 0) {  // n <= SCAN_L: one chunk forms the ratios, one thread scans them
    prod_forward<<<1, 128, 0, st>>>((const Fr*)d_num, w.b, n, w.a, w.spare, 1);
    prod_exclusive_serial<<<1, 32, 0, st>>>(w.a, n, (Fr*)d_z);
    ctx->launches += 2;
    H2AGG_CUDA(ctx, cudaGetLastError());
    return 0;
  }
  const Fr* src = (const Fr*)d_num;
  for (int l = 0; l < w.levels; l++) {
    size_t m = w.len[l + 1];
    prod_forward<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(src, l == 0 ? w.b : nullptr, w.len[l], l == 0 ? w.a : nullptr,
                                                             w.lvl[l], m);
    ctx->launches++;
    src = w.lvl[l];
  }
  // 3. exclusive carries of the top level, then back-fill downwards
  prod_exclusive_serial<<<1, 32, 0, st>>>(w.lvl[w.levels - 1], w.len[w.levels], w.aux2[w.levels - 1]);
  ctx->launches++;
  for (int l = w.levels - 1; l >= 0; l--) {
    const Fr* lsrc = (l == 0) ? w.a : w.lvl[l - 1];
    Fr* lout = (l == 0) ? (Fr*)d_z : w.aux2[l - 1];
    size_t m = w.len[l + 1];
    prod_backfill<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(lsrc, w.len[l], w.aux2[l], lout, m);
    ctx->launches++;
  }
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace h2agg

using namespace h2agg;

extern "C" {

int h2agg_batch_invert_dev(h2agg_ctx* ctx, void* d_a, size_t n) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_a) { ctx->last_error = "batch_invert: null argument"; return 1; }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  return batch_invert_dev(ctx, d_a, n);
}

int h2agg_batch_invert(h2agg_ctx* ctx, uint64_t* a, size_t n) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!a) { ctx->last_error = "batch_invert: null argument"; return 1; }
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure(ctx, ctx->io_a, n * 32);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  rc = batch_invert_dev(ctx, ctx->io_a.p, n);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(a, ctx->io_a.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int h2agg_grand_product_dev(h2agg_ctx* ctx, const void* d_num, const void* d_den, size_t n, void* d_z) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_num || !d_den || !d_z) { ctx->last_error = "grand_product: null argument"; return 1; }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  return grand_product_dev(ctx, d_num, d_den, n, d_z, nullptr, nullptr);
}

int h2agg_grand_product(h2agg_ctx* ctx, const uint64_t* num, const uint64_t* den, size_t n, uint64_t* z) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!num || !den || !z) { ctx->last_error = "grand_product: null argument"; return 1; }
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure(ctx, ctx->io_a, n * 64);
  if (rc) return rc;
  rc = ensure(ctx, ctx->io_b, n * 32);
  if (rc) return rc;
  uint8_t* d_num = (uint8_t*)ctx->io_a.p;
  uint8_t* d_den = d_num + n * 32;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(d_num, num, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  H2AGG_CUDA(ctx, cudaMemcpyAsync(d_den, den, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  rc = grand_product_dev(ctx, d_num, d_den, n, ctx->io_b.p, nullptr, nullptr);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(z, ctx->io_b.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
