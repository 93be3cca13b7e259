// N3 ("next" row), the remaining row-wise steps of the lookup and permutation arguments, so that
// create_proof's second and third commit rounds can be produced from resident columns without leaving HBM:
//
//   compress_expressions   lookup::Argument::commit_permuted: every input / table expression is evaluated on the
//                          Lagrange domain (rotations wrap mod n) and folded  acc = acc * theta + e_i
//   lookup_product         lookup::Permuted::commit_product:  z_0 = 1,
//                          z_{i+1} = z_i (A_i + beta)(S_i + gamma) / ((A'_i + beta)(S'_i + gamma))
//   permutation_product    permutation::Argument::commit, one column set:  z_0 = last z of the previous set,
//                          z_{i+1} = z_i prod_j (v_j(i) + beta delta^j omega^i + gamma) / (v_j(i) + beta sigma_j(i) + gamma)
//
// (halo2_proofs plonk/lookup/prover.rs, plonk/permutation/prover.rs -- external crate, restated in
// oracle/py/lookup_ref.py; the reference's verifier checks exactly these recurrences:
// halo2-snark-aggregator-api/src/systems/halo2/lookup.rs:58-119, permutation.rs:54-136.)
// Numerators and denominators are formed row-wise here; the running product itself is scan.cu's grand_product.
// The blinding rows at the end of every column are the caller's random values and are written by the caller.
#include "../../include/h2agg.h"
#include "bn254_field.cuh"
#include "ctx.hpp"
#include <cstring>
#include <string>
#include <vector>

namespace h2agg {

int grand_product_dev(h2agg_ctx* ctx, const void* d_num, const void* d_den, size_t n, void* d_z, cudaStream_t st,
                      DevBuf* ws);  // scan.cu

static constexpr uint32_t ANOCONST = 0xffffffffu;

struct CompressArgs {
  const uint32_t* plan;  // n_exprs, then per expression a sum of products (POLY layout of include/h2agg.h)
  const Fr* const* cols;
  const Fr* consts;
  Fr* out;
  uint32_t k;
  Fr theta;
};

__global__ void __launch_bounds__(256) compress_expressions_kernel(const __grid_constant__ CompressArgs a) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t mask = (1u << a.k) - 1;
  if (idx > mask) return;
  const uint32_t* __restrict__ plan = a.plan;
  uint32_t pc = 0;
  const uint32_t ne = plan[pc++];
  Fr acc = Fr::zero();
  for (uint32_t e = 0; e < ne; e++) {
    const uint32_t nt = plan[pc++];
    Fr sum = Fr::zero();
    for (uint32_t t = 0; t < nt; t++) {
      const uint32_t ci = plan[pc++], nf = plan[pc++];
      Fr prod = Fr::one();
      bool have = false;
      if (ci != ANOCONST) {
        prod = Fr::load_nc(a.consts + ci);
        have = true;
      }
      for (uint32_t f = 0; f < nf; f++) {
        const uint32_t w = plan[pc++];
        const uint32_t r = (idx + (uint32_t)(int)(int16_t)(w >> 16)) & mask;
        Fr v = Fr::load_nc(a.cols[w & 0xffffu] + r);
        prod = have ? prod * v : v;
        have = true;
      }
      sum = sum + prod;
    }
    acc = (e == 0) ? sum : acc * a.theta + sum;
  }
  acc.store(a.out + idx);
}

__global__ void __launch_bounds__(256) lookup_num_den_kernel(const Fr* __restrict__ A, const Fr* __restrict__ S,
                                                             const Fr* __restrict__ Ap, const Fr* __restrict__ Sp, size_t n,
                                                             Fr beta, Fr gamma, Fr* __restrict__ num, Fr* __restrict__ den) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ((Fr::load_nc(A + i) + beta) * (Fr::load_nc(S + i) + gamma)).store(num + i);
  ((Fr::load_nc(Ap + i) + beta) * (Fr::load_nc(Sp + i) + gamma)).store(den + i);
}

static constexpr uint32_t PERM_ROWS = 8;    // consecutive rows per thread: one omega^row power, then steps of omega
static constexpr uint32_t PERM_MAX_COLS = 16;

struct PermArgs {
  const Fr* values[PERM_MAX_COLS];
  const Fr* sigmas[PERM_MAX_COLS];
  uint32_t n_cols, k;
  Fr omega, beta, gamma, delta, beta_delta_start;  // beta * delta^(index of the set's first column)
  Fr *num, *den;
};

__global__ void __launch_bounds__(128) perm_num_den_kernel(const __grid_constant__ PermArgs a) {
  const size_t n = (size_t)1 << a.k;
  const size_t row0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * PERM_ROWS;
  if (row0 >= n) return;
  Fr w = fp_pow_u64(a.omega, row0);
  for (uint32_t r = 0; r < PERM_ROWS && row0 + r < n; r++) {
    const size_t i = row0 + r;
    Fr num = Fr::one(), den = Fr::one();
    Fr d = a.beta_delta_start;
    for (uint32_t j = 0; j < a.n_cols; j++) {
      Fr v = Fr::load_nc(a.values[j] + i);
      Fr vg = v + a.gamma;
      Fr t_den = vg + a.beta * Fr::load_nc(a.sigmas[j] + i);
      Fr t_num = vg + d * w;
      if (j == 0) { num = t_num; den = t_den; }
      else { num = num * t_num; den = den * t_den; }
      d = d * a.delta;
    }
    num.store(a.num + i);
    den.store(a.den + i);
    w = w * a.omega;
  }
}

__global__ void __launch_bounds__(256) scale_by_element_kernel(Fr* __restrict__ z, size_t n, const Fr* __restrict__ factor) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  (Fr::load(z + i) * Fr::load_nc(factor)).store(z + i);
}

static const char* check_expr_list(const uint32_t* p, size_t n, size_t n_cols, size_t n_consts) {
  size_t pc = 0;
  if (n == 0) return "empty";
  const uint32_t ne = p[pc++];
  if (ne == 0) return "no expressions";
  for (uint32_t e = 0; e < ne; e++) {
    if (pc >= n) return "truncated";
    const uint32_t nt = p[pc++];
    for (uint32_t t = 0; t < nt; t++) {
      if (pc + 2 > n) return "truncated";
      const uint32_t ci = p[pc++], nf = p[pc++];
      if (ci != ANOCONST && ci >= n_consts) return "constant index out of range";
      if (nf > 64 || pc + nf > n) return "truncated";
      for (uint32_t f = 0; f < nf; f++)
        if ((p[pc++] & 0xffffu) >= n_cols) return "column out of range";
    }
  }
  return pc == n ? nullptr : "trailing words";
}

// scratch: [num | den], n x 32 B each
static int num_den_ws(h2agg_ctx* ctx, size_t n, Fr** num, Fr** den) {
  int rc = ensure(ctx, ctx->args_ws, 2 * n * 32);
  if (rc) return rc;
  *num = (Fr*)ctx->args_ws.p;
  *den = *num + n;
  return 0;
}

}  // namespace h2agg

using namespace h2agg;

extern "C" {

int h2agg_compress_expressions_dev(h2agg_ctx* ctx, const uint32_t* exprs, size_t n_words, const void* const* d_columns,
                                   size_t n_columns, const uint64_t* consts, size_t n_consts, uint32_t k,
                                   const uint64_t theta[4], void* d_out) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!exprs || !d_columns || !theta || !d_out || (n_consts && !consts)) { ctx->last_error = "compress_expressions: null argument"; return 1; }
  if (k == 0 || k > 28 || n_columns == 0 || n_columns > 65535) { ctx->last_error = "compress_expressions: bad k or column count"; return 1; }
  if (const char* why = check_expr_list(exprs, n_words, n_columns, n_consts)) {
    ctx->last_error = std::string("compress_expressions: invalid expression list: ") + why;
    return 1;
  }
  for (size_t i = 0; i < n_columns; i++)
    if (!d_columns[i]) { ctx->last_error = "compress_expressions: null column pointer"; return 1; }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t off_cols = (n_words * 4 + 31) & ~(size_t)31;
  const size_t off_consts = (off_cols + n_columns * 8 + 31) & ~(size_t)31;
  const size_t total = ((off_consts + n_consts * 32 + 32) + 255) & ~(size_t)255;
  // a previous launch may still be reading its copy: rotate through four slots
  int rc = ensure(ctx, ctx->args_meta, 4 * total);
  if (rc) return rc;
  ctx->args_flip = (ctx->args_flip + 1) & 3;
  uint8_t* base = (uint8_t*)ctx->args_meta.p + (size_t)ctx->args_flip * total;
  std::vector<uint8_t> stage(total, 0);
  memcpy(stage.data(), exprs, n_words * 4);
  memcpy(stage.data() + off_cols, d_columns, n_columns * 8);
  if (n_consts) memcpy(stage.data() + off_consts, consts, n_consts * 32);
  H2AGG_CUDA(ctx, cudaMemcpyAsync(base, stage.data(), total, cudaMemcpyHostToDevice, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `stage` is pageable and dies with this frame
  CompressArgs a;
  a.plan = (const uint32_t*)base;
  a.cols = (const Fr* const*)(base + off_cols);
  a.consts = (const Fr*)(base + off_consts);
  a.out = (Fr*)d_out;
  a.k = k;
  memcpy(a.theta.v, theta, 32);
  const size_t n = (size_t)1 << k;
  compress_expressions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

int h2agg_lookup_product_dev(h2agg_ctx* ctx, const void* d_input, const void* d_table, const void* d_permuted_input,
                             const void* d_permuted_table, size_t n, const uint64_t beta[4], const uint64_t gamma[4],
                             void* d_z) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_input || !d_table || !d_permuted_input || !d_permuted_table || !beta || !gamma || !d_z) {
    ctx->last_error = "lookup_product: null argument";
    return 1;
  }
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  Fr *num, *den;
  int rc = num_den_ws(ctx, n, &num, &den);
  if (rc) return rc;
  Fr b, g;
  memcpy(b.v, beta, 32);
  memcpy(g.v, gamma, 32);
  lookup_num_den_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      (const Fr*)d_input, (const Fr*)d_table, (const Fr*)d_permuted_input, (const Fr*)d_permuted_table, n, b, g, num, den);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return grand_product_dev(ctx, num, den, n, d_z, nullptr, nullptr);
}

// All lookups of a constraint system at once: lookup i runs on lane i % N_LANES with that lane's scratch, so the
// latency chains of the grand products (an inversion and ~15 small launches each) overlap instead of queueing up.
int h2agg_lookup_products_dev(h2agg_ctx* ctx, size_t n_lookups, const void* const* d_inputs, const void* const* d_tables,
                              const void* const* d_permuted_inputs, const void* const* d_permuted_tables, size_t n,
                              const uint64_t beta[4], const uint64_t gamma[4], void* const* d_z) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!beta || !gamma || (n_lookups && (!d_inputs || !d_tables || !d_permuted_inputs || !d_permuted_tables || !d_z))) {
    ctx->last_error = "lookup_products: null argument";
    return 1;
  }
  for (size_t i = 0; i < n_lookups; i++)
    if (!d_inputs[i] || !d_tables[i] || !d_permuted_inputs[i] || !d_permuted_tables[i] || !d_z[i]) {
      ctx->last_error = "lookup_products: null column pointer";
      return 1;
    }
  if (n == 0 || n_lookups == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = lanes_init(ctx);
  if (rc) return rc;
  Fr b, g;
  memcpy(b.v, beta, 32);
  memcpy(g.v, gamma, 32);
  LaneFork lf(ctx);
  if ((rc = lf.fork())) return rc;
  for (size_t i = 0; i < n_lookups; i++) {
    Lane& ln = ctx->lanes[i % N_LANES];
    if ((rc = ensure(ctx, ln.args_ws, 2 * n * 32))) return rc;
    Fr* num = (Fr*)ln.args_ws.p;
    Fr* den = num + n;
    lookup_num_den_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ln.st>>>((const Fr*)d_inputs[i], (const Fr*)d_tables[i],
                                                                          (const Fr*)d_permuted_inputs[i],
                                                                          (const Fr*)d_permuted_tables[i], n, b, g, num, den);
    ctx->launches++;
    H2AGG_CUDA(ctx, cudaGetLastError());
    if ((rc = grand_product_dev(ctx, num, den, n, d_z[i], ln.st, &ln.scan_ws))) return rc;
  }
  return lf.join();
}

int h2agg_permutation_product_dev(h2agg_ctx* ctx, const void* const* d_values, const void* const* d_sigmas, size_t n_cols,
                                  uint32_t k, const uint64_t omega[4], const uint64_t beta_delta_start[4],
                                  const uint64_t delta[4], const uint64_t beta[4], const uint64_t gamma[4],
                                  const void* d_last_z, void* d_z) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_values || !d_sigmas || !omega || !beta_delta_start || !delta || !beta || !gamma || !d_z) {
    ctx->last_error = "permutation_product: null argument";
    return 1;
  }
  if (n_cols == 0 || n_cols > PERM_MAX_COLS || k == 0 || k > 28) {
    ctx->last_error = "permutation_product: 1..16 columns per set, 1 <= k <= 28";
    return 1;
  }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << k;
  PermArgs a;
  memset(&a, 0, sizeof(a));
  for (size_t j = 0; j < n_cols; j++) {
    if (!d_values[j] || !d_sigmas[j]) { ctx->last_error = "permutation_product: null column pointer"; return 1; }
    a.values[j] = (const Fr*)d_values[j];
    a.sigmas[j] = (const Fr*)d_sigmas[j];
  }
  a.n_cols = (uint32_t)n_cols;
  a.k = k;
  memcpy(a.omega.v, omega, 32);
  memcpy(a.beta.v, beta, 32);
  memcpy(a.gamma.v, gamma, 32);
  memcpy(a.delta.v, delta, 32);
  memcpy(a.beta_delta_start.v, beta_delta_start, 32);
  int rc = num_den_ws(ctx, n, &a.num, &a.den);
  if (rc) return rc;
  const size_t threads = (n + PERM_ROWS - 1) / PERM_ROWS;
  perm_num_den_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(a);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  if ((rc = grand_product_dev(ctx, a.num, a.den, n, d_z, nullptr, nullptr))) return rc;
  if (d_last_z) {
    scale_by_element_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((Fr*)d_z, n, (const Fr*)d_last_z);
    ctx->launches++;
    H2AGG_CUDA(ctx, cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
