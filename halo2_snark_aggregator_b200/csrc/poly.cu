// N2 (first "next" row): polynomial evaluation and Kate division on the device, so
// that coefficient vectors produced by the iNTTs never have to leave HBM for the evaluation round and
// the GWC multi-opening of create_proof (halo2_proofs arithmetic.rs `eval_polynomial`, `kate_division`;
// external crate, restated in SURVEY.md App. B4 steps 8-9; reference call site
// halo2-snark-aggregator-circuit/src/verify_circuit.rs:986).
//
//   eval_polynomial(a, x)  = sum_i a_i x^i
//   kate_division(a, b)    = q with q_i = sum_{j>i} a_j b^(j-i-1)   (quotient of a(X) by (X - b), remainder dropped)
//
// Both are linear recurrences; they are cut into chunks of L = 64 coefficients.  One kernel computes
// every chunk's Horner value S_c = sum_j a_{cL+j} y^j; the vector S is itself a polynomial in y^L, so
// the same routine recurses on it (4M -> 64K -> 1K -> 16 -> serial).  Evaluation returns the top of
// the recursion; division additionally runs each chunk's recurrence seeded with its carry.
// One multiplication per coefficient per sweep: like everything on this path these kernels are bound
// by the 256-bit multiplier (68 G mul/s = 2.2 TB/s of 32-byte elements), not by HBM.
#include "../../include/h2agg.h"
#include "bn254_field.cuh"
#include "ctx.hpp"
#include <algorithm>
#include <cstring>

namespace h2agg {

static constexpr uint32_t POLY_L = 64;
static constexpr uint32_t POLY_SERIAL = 64;  // at or below this length one thread finishes the job

// out[0] = y^(2^0) ... out[k] = y^(2^k)
__global__ void poly_pow2_table(Fr y, uint32_t k, Fr* out) {
  if (threadIdx.x || blockIdx.x) return;
  Fr v = y;
  for (uint32_t i = 0; i <= k; i++) {
    v.store(out + i);
    v = fp_sqr(v);
  }
}

// S_c = sum_{j < L} a[c L + j] y^j   (missing tail coefficients count as zero)
__global__ void __launch_bounds__(128) poly_chunk_horner(const Fr* __restrict__ a, size_t n, const Fr* __restrict__ y_ptr,
                                                          Fr* __restrict__ s, size_t m) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  const Fr y = Fr::load(y_ptr);
  size_t lo = c * POLY_L, hi = lo + POLY_L;
  if (hi > n) hi = n;
  Fr acc = Fr::zero();
  for (size_t i = hi; i-- > lo;) acc = acc * y + Fr::load_nc(a + i);
  acc.store(s + c);
}

// serial evaluation of a short vector
__global__ void poly_eval_serial(const Fr* __restrict__ a, size_t n, const Fr* __restrict__ y_ptr, Fr* out) {
  if (threadIdx.x || blockIdx.x) return;
  const Fr y = Fr::load(y_ptr);
  Fr acc = Fr::zero();
  for (size_t i = n; i-- > 0;) acc = acc * y + Fr::load(a + i);
  acc.store(out);
}

// serial Kate division of a short vector: q[i-1] = a[i] + y q[i], q has n entries with q[n-1] = 0
__global__ void poly_kate_serial(const Fr* __restrict__ a, size_t n, const Fr* __restrict__ y_ptr, Fr* q) {
  if (threadIdx.x || blockIdx.x) return;
  const Fr y = Fr::load(y_ptr);
  Fr run = Fr::zero();
  for (size_t i = n; i-- > 0;) {
    run.store(q + i);
    run = Fr::load(a + i) + y * run;
  }
}

// per chunk: q[i] for i in the chunk, seeded with the carry C_c = q[(c+1)L - 1]
__global__ void __launch_bounds__(128) poly_kate_chunks(const Fr* __restrict__ a, size_t n, const Fr* __restrict__ y_ptr,
                                                         const Fr* __restrict__ carry, Fr* __restrict__ q, size_t m) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  const Fr y = Fr::load(y_ptr);
  size_t lo = c * POLY_L, hi = lo + POLY_L;
  if (hi > n) hi = n;
  Fr run = Fr::load(carry + c);
  for (size_t i = hi; i-- > lo;) {
    run.store(q + i);
    run = Fr::load_nc(a + i) + y * run;
  }
}

// workspace layout: pow table (64 entries) then the recursion levels
struct PolyWs {
  Fr* pow2;
  Fr* lvl[8];
  Fr* carry[8];
};

static int poly_ws(h2agg_ctx* ctx, size_t n, PolyWs* w) {
  size_t total = 64;
  size_t m = n;
  int levels = 0;
  size_t sizes[8];
  while (m > POLY_SERIAL && levels < 8) {
    m = (m + POLY_L - 1) / POLY_L;
    sizes[levels++] = m;
    total += 2 * m;
  }
  int rc = ensure(ctx, ctx->poly_ws, total * 32 + 256);
  if (rc) return rc;
  Fr* p = (Fr*)ctx->poly_ws.p;
  w->pow2 = p;
  p += 64;
  for (int i = 0; i < levels; i++) {
    w->lvl[i] = p;
    p += sizes[i];
    w->carry[i] = p;
    p += sizes[i];
  }
  return 0;
}

// evaluation: d_out receives one Fr
int poly_eval_dev(h2agg_ctx* ctx, const void* d_a, size_t n, const uint64_t point[4], void* d_out) {
  PolyWs w;
  int rc = poly_ws(ctx, n, &w);
  if (rc) return rc;
  Fr x;
  memcpy(x.v, point, 32);
  cudaStream_t st = ctx->stream;
  poly_pow2_table<<<1, 32, 0, st>>>(x, 48, w.pow2);
  ctx->launches++;
  const Fr* cur = (const Fr*)d_a;
  size_t m = n;
  int lvl = 0;
  uint32_t lg = 0;  // cur is a polynomial in x^(2^lg)
  while (m > POLY_SERIAL) {
    size_t mo = (m + POLY_L - 1) / POLY_L;
    poly_chunk_horner<<<(unsigned)((mo + 127) / 128), 128, 0, st>>>(cur, m, w.pow2 + lg, w.lvl[lvl], mo);
    ctx->launches++;
    cur = w.lvl[lvl++];
    m = mo;
    lg += 6;
  }
  poly_eval_serial<<<1, 32, 0, st>>>(cur, m, w.pow2 + lg, (Fr*)d_out);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ---- many polynomials of the same length at ONE point (the evaluation round of create_proof opens ~50 polynomials at
// x and a handful at each rotated point): the same chunked recursion with the polynomial index as blockIdx.y, so the
// whole group costs five launches instead of five per polynomial.
static constexpr uint32_t EVAL_BATCH = 64;
struct EvalBatchArgs {
  const Fr* polys[EVAL_BATCH];  // level 0: one pointer per polynomial
  const Fr* base;               // later levels: vector p at base + p * stride
  size_t stride;
  size_t n, m;                  // input length, number of chunks
  const Fr* y_ptr;
  Fr* s;                        // out: [p][m]
};
__global__ void __launch_bounds__(128) poly_chunk_horner_batch(const __grid_constant__ EvalBatchArgs a) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.m) return;
  const uint32_t p = blockIdx.y;
  const Fr* __restrict__ v = a.base ? a.base + p * a.stride : a.polys[p];
  const Fr y = Fr::load(a.y_ptr);
  size_t lo = c * POLY_L, hi = lo + POLY_L;
  if (hi > a.n) hi = a.n;
  Fr acc = Fr::zero();
  for (size_t i = hi; i-- > lo;) acc = acc * y + Fr::load_nc(v + i);
  acc.store(a.s + p * a.m + c);
}
// one thread per polynomial finishes its short vector
__global__ void poly_eval_serial_batch(const EvalBatchArgs a, uint32_t n_polys, Fr* out) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_polys) return;
  const Fr* v = a.base ? a.base + p * a.stride : a.polys[p];
  const Fr y = Fr::load(a.y_ptr);
  Fr acc = Fr::zero();
  for (size_t i = a.n; i-- > 0;) acc = acc * y + Fr::load(v + i);
  acc.store(out + p);
}

int poly_eval_many_dev(h2agg_ctx* ctx, const void* const* d_polys, size_t n_polys, size_t n, const uint64_t point[4], void* d_out) {
  if (n_polys == 0) return 0;
  // workspace: pow table + two ping-pong level buffers of EVAL_BATCH x ceil(n / L) entries
  const size_t m1 = (n + POLY_L - 1) / POLY_L;
  int rc = ensure(ctx, ctx->poly_many_ws, (64 + 2 * EVAL_BATCH * m1) * 32 + 256);
  if (rc) return rc;
  Fr* pow2 = (Fr*)ctx->poly_many_ws.p;
  Fr* lvl[2] = {pow2 + 64, pow2 + 64 + EVAL_BATCH * m1};
  Fr x;
  memcpy(x.v, point, 32);
  cudaStream_t st = ctx->stream;
  poly_pow2_table<<<1, 32, 0, st>>>(x, 48, pow2);
  ctx->launches++;
  for (size_t done = 0; done < n_polys; done += EVAL_BATCH) {
    const uint32_t P = (uint32_t)std::min<size_t>(EVAL_BATCH, n_polys - done);
    EvalBatchArgs a;
    memset(&a, 0, sizeof(a));
    for (uint32_t p = 0; p < P; p++) a.polys[p] = (const Fr*)d_polys[done + p];
    a.n = n;
    uint32_t lg = 0;
    int flip = 0;
    while (a.n > POLY_SERIAL) {
      a.m = (a.n + POLY_L - 1) / POLY_L;
      a.y_ptr = pow2 + lg;
      a.s = lvl[flip];
      poly_chunk_horner_batch<<<dim3((unsigned)((a.m + 127) / 128), P), 128, 0, st>>>(a);
      ctx->launches++;
      a.base = lvl[flip];
      a.stride = a.m;
      a.n = a.m;
      flip ^= 1;
      lg += 6;
    }
    a.y_ptr = pow2 + lg;
    poly_eval_serial_batch<<<(P + 31) / 32, 32, 0, st>>>(a, P, (Fr*)d_out + done);
    ctx->launches++;
  }
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

// division: d_q receives n entries, q[n-1] = 0 (halo2 returns the first n-1)
int poly_kate_dev(h2agg_ctx* ctx, const void* d_a, size_t n, const uint64_t b[4], void* d_q) {
  if (n == 0) return 0;
  PolyWs w;
  int rc = poly_ws(ctx, n, &w);
  if (rc) return rc;
  Fr x;
  memcpy(x.v, b, 32);
  cudaStream_t st = ctx->stream;
  poly_pow2_table<<<1, 32, 0, st>>>(x, 48, w.pow2);
  ctx->launches++;
  // downward sweep: chunk values of every level
  const Fr* src[9];
  size_t len[9];
  src[0] = (const Fr*)d_a;
  len[0] = n;
  int levels = 0;
  uint32_t lg = 0;
  while (len[levels] > POLY_SERIAL) {
    size_t mo = (len[levels] + POLY_L - 1) / POLY_L;
    poly_chunk_horner<<<(unsigned)((mo + 127) / 128), 128, 0, st>>>(src[levels], len[levels], w.pow2 + lg, w.lvl[levels], mo);
    ctx->launches++;
    src[levels + 1] = w.lvl[levels];
    len[levels + 1] = mo;
    levels++;
    lg += 6;
  }
  // top: serial division of the shortest vector (into its carry array, or straight into q when there is one level)
  Fr* top_out = (levels == 0) ? (Fr*)d_q : w.carry[levels - 1];
  poly_kate_serial<<<1, 32, 0, st>>>(src[levels], len[levels], w.pow2 + lg, top_out);
  ctx->launches++;
  // upward sweep: carries of level l are the quotient entries of level l+1
  for (int l = levels - 1; l >= 0; l--) {
    lg -= 6;
    Fr* out = (l == 0) ? (Fr*)d_q : w.carry[l - 1];
    size_t mo = len[l + 1];
    poly_kate_chunks<<<(unsigned)((mo + 127) / 128), 128, 0, st>>>(src[l], len[l], w.pow2 + lg, w.carry[l], out, mo);
    ctx->launches++;
  }
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace h2agg

using namespace h2agg;

extern "C" {

int h2agg_eval_polynomial_dev(h2agg_ctx* ctx, const void* d_poly, size_t n, const uint64_t point[4], void* d_out32) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_poly || !point || !d_out32) { ctx->last_error = "eval_polynomial: null argument"; return 1; }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  return poly_eval_dev(ctx, d_poly, n, point, d_out32);
}

int h2agg_eval_polynomials_dev(h2agg_ctx* ctx, const void* const* d_polys, size_t n_polys, size_t n, const uint64_t point[4],
                               void* d_out) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!point || (n_polys && (!d_polys || !d_out))) { ctx->last_error = "eval_polynomials: null argument"; return 1; }
  for (size_t i = 0; i < n_polys; i++)
    if (!d_polys[i]) { ctx->last_error = "eval_polynomials: null polynomial"; return 1; }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  return poly_eval_many_dev(ctx, d_polys, n_polys, n, point, d_out);
}

int h2agg_eval_polynomial(h2agg_ctx* ctx, const uint64_t* poly, size_t n, const uint64_t point[4], uint64_t out[4]) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!poly || !point || !out) { ctx->last_error = "eval_polynomial: null argument"; return 1; }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure(ctx, ctx->io_a, n * 32 + 64);
  if (rc) return rc;
  if (n) H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, poly, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  rc = poly_eval_dev(ctx, ctx->io_a.p, n, point, ctx->small.p);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->small.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->pinned, 32);
  return 0;
}

int h2agg_kate_division_dev(h2agg_ctx* ctx, const void* d_a, size_t n, const uint64_t b[4], void* d_q) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_a || !b || !d_q) { ctx->last_error = "kate_division: null argument"; return 1; }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  return poly_kate_dev(ctx, d_a, n, b, d_q);
}

int h2agg_kate_division(h2agg_ctx* ctx, const uint64_t* a, size_t n, const uint64_t b[4], uint64_t* q) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!a || !b || !q || n == 0) { ctx->last_error = "kate_division: null argument or empty polynomial"; return 1; }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure(ctx, ctx->io_a, n * 32 + 64);
  if (rc) return rc;
  rc = ensure(ctx, ctx->io_b, n * 32 + 64);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  rc = poly_kate_dev(ctx, ctx->io_a.p, n, b, ctx->io_b.p);
  if (rc) return rc;
  if (n > 1) H2AGG_CUDA(ctx, cudaMemcpyAsync(q, ctx->io_b.p, (n - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
