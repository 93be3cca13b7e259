// N1: the quotient numerator of create_proof on the extended coset, fused with the
// division by the vanishing polynomial -- halo2_proofs plonk/evaluation.rs `Evaluator::evaluate_h` followed by
// `EvaluationDomain::divide_by_vanishing_poly` (external crate; reference call site
// halo2-snark-aggregator-circuit/src/verify_circuit.rs:986).  The reference holds the VERIFIER's copy of the same
// equations, which is what pins the order of the y-fold and every formula here:
//   gates, then permutation, then lookups          halo2-snark-aggregator-api/src/systems/halo2/params.rs:95-150
//   permutation terms (first, last, links, sets)   .../permutation.rs:54-136
//   lookup terms (five per lookup)                 .../lookup.rs:58-119
//   fold with y, divide by x^n - 1                 .../vanish.rs:28-29
//
// One thread per row of the 2^ext_k coset.  Every column (fixed, advice, instance, sigma, l_0, l_last,
// l_active_row, permutation z, lookup z / a' / s') is an extended-coset evaluation vector resident in HBM;
// consecutive threads read consecutive 32-byte elements of each column, rotations are (idx + rot * 2^(ext_k-k))
// mod 2^ext_k and hit the same or a neighbouring line.  The constraint system arrives as a small word program
// (the "plan"; layout in include/h2agg.h): every polynomial is a sum of products of column queries, which any
// halo2 Expression expands to, so the kernel needs no expression stack.
//
// Work per row for the aggregation circuit (1 gate, 6 permutation columns in 2 sets, 7 lookups): ~180 Fr products
// against ~75 x 32 B read: like everything on this path the kernel is bound by the 256-bit multiplier, not HBM.
#include "../../include/h2agg.h"
#include "bn254_field.cuh"
#include "ctx.hpp"
#include <cstring>
#include <vector>

namespace h2agg {

static constexpr uint32_t QPLAN_MAGIC = 0x31485148u;  // "HQH1"
static constexpr uint32_t QPLAN_HEADER = 9;
static constexpr uint32_t QNOCONST = 0xffffffffu;
static constexpr uint32_t QUOT_LO_BITS = 12;

struct QuotKernelArgs {
  const uint32_t* plan;
  const Fr* const* cols;
  const Fr* consts;
  const Fr* t_lo;   // omega_ext^i, i < 2^lo_bits
  const Fr* t_hi;   // omega_ext^(j << lo_bits)
  const Fr* t_evals;  // 1 / ((zeta omega_ext^i)^n - 1), i < t_mask + 1 (or null)
  const Fr* y_pow;    // y^e, e < n_terms
  uint32_t n_terms;   // number of folded terms: h = sum_j term_j * y^(n_terms - 1 - j)
  Fr* out;
  uint32_t lo_bits, ext_k, rot_scale, t_mask;
  // row window (multi-GPU row sharding): this launch produces rows [row_begin, row_begin + row_count); column c's buffer
  // starts at global row col_row0[c] (cyclically), so a buffer only needs the window plus its rotation halo
  uint32_t row_begin, row_count;
  const uint32_t* col_row0;  // null = every buffer is the whole column (row0 = 0)
  Fr y, beta, gamma, theta, zeta, delta;
};

__global__ void quot_gen_tables(Fr omega, uint32_t lo_bits, uint32_t hi_bits, Fr* t_lo, Fr* t_hi) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nlo = 1u << lo_bits, nhi = 1u << hi_bits;
  if (i < nlo) {
    fp_pow_u64(omega, i).store(t_lo + i);
  } else if (i < nlo + nhi) {
    uint32_t j = i - nlo;
    fp_pow_u64(omega, (uint64_t)j << lo_bits).store(t_hi + j);
  }
}

__global__ void quot_y_powers(Fr y, uint32_t n, Fr* out) {
  if (threadIdx.x || blockIdx.x) return;
  Fr acc = Fr::one();
  for (uint32_t e = 0; e < n; e++) {
    acc.store(out + e);
    acc = acc * y;
  }
}

__device__ __forceinline__ Fr q_load(const QuotKernelArgs& a, uint32_t col, int rot, uint32_t idx, uint32_t mask) {
  uint32_t r = idx + (uint32_t)(rot * (int)a.rot_scale);
  if (a.col_row0) r -= __ldg(a.col_row0 + col);
  return Fr::load_nc(a.cols[col] + (r & mask));
}

// sum of products: n_terms, then per term { const index or QNOCONST, n_factors, factor words (col | rot << 16) }
__device__ __forceinline__ Fr q_sop(const QuotKernelArgs& a, uint32_t& pc, uint32_t idx, uint32_t mask) {
  const uint32_t* __restrict__ plan = a.plan;
  uint32_t nt = plan[pc++];
  Fr acc = Fr::zero();
  for (uint32_t t = 0; t < nt; t++) {
    uint32_t ci = plan[pc++], nf = plan[pc++];
    Fr prod = Fr::one();
    bool have = false;
    if (ci != QNOCONST) {
      prod = Fr::load_nc(a.consts + ci);
      have = true;
    }
    for (uint32_t f = 0; f < nf; f++) {
      uint32_t w = plan[pc++];
      Fr v = q_load(a, w & 0xffffu, (int)(int16_t)(w >> 16), idx, mask);
      if (have) prod = prod * v;
      else prod = v;
      have = true;
    }
    acc = acc + prod;
  }
  return acc;
}

// theta-compression of a list of expressions: acc = acc * theta + e_i (halo2 Calculation::Horner(0, parts, Theta))
__device__ __forceinline__ Fr q_compress(const QuotKernelArgs& a, uint32_t& pc, uint32_t idx, uint32_t mask) {
  uint32_t ne = a.plan[pc++];
  Fr acc = Fr::zero();
  for (uint32_t e = 0; e < ne; e++) {
    Fr v = q_sop(a, pc, idx, mask);
    acc = (e == 0) ? v : acc * a.theta + v;
  }
  return acc;
}

__global__ void __launch_bounds__(256) quot_evaluate_h(const __grid_constant__ QuotKernelArgs a) {
  const uint32_t local = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t mask = (1u << a.ext_k) - 1;
  if (local >= a.row_count) return;
  const uint32_t idx = (a.row_begin + local) & mask;
  const uint32_t* __restrict__ plan = a.plan;
  const uint32_t n_gates = plan[1], n_pcols = plan[2], chunk_len = plan[3];
  const int last_rot = (int)plan[4];
  const uint32_t n_lookups = plan[5];
  const Fr one = Fr::one();
  uint32_t pc = QPLAN_HEADER;
  // halo2 folds h = h * y + term over the ordered term list, i.e. h = sum_j term_j * y^(N-1-j).  Most terms carry one
  // of three selectors (l_0, l_last, l_active_row); summing the y-weighted terms per selector and multiplying by the
  // selector ONCE replaces "times selector, times y" by "times y^e" for every such term (same field element, ~20 %
  // fewer products per row for the aggregation circuit).
  Fr acc_plain = Fr::zero(), acc_l0 = Fr::zero(), acc_last = Fr::zero(), acc_active = Fr::zero();
  uint32_t e = a.n_terms;  // exponent of the NEXT term is e - 1
  auto weighted = [&](const Fr& term) {
    e--;
    return e ? term * Fr::load_nc(a.y_pow + e) : term;
  };
  auto fold = [&](const Fr& term) { acc_plain = acc_plain + weighted(term); };
  auto fold_l0 = [&](const Fr& term) { acc_l0 = acc_l0 + weighted(term); };
  auto fold_last = [&](const Fr& term) { acc_last = acc_last + weighted(term); };
  auto fold_active = [&](const Fr& term) { acc_active = acc_active + weighted(term); };

  // custom gates, in gate order
  for (uint32_t g = 0; g < n_gates; g++) fold(q_sop(a, pc, idx, mask));

  // permutation argument
  if (n_pcols) {
    const uint32_t n_sets = (n_pcols + chunk_len - 1) / chunk_len;
    const uint32_t pcols = pc;            // n_pcols x {value column, sigma column}
    const uint32_t zcols = pc + 2 * n_pcols;  // n_sets x z column
    pc = zcols + n_sets;
    {
      Fr z0 = q_load(a, plan[zcols], 0, idx, mask);
      fold_l0(one - z0);
      Fr zl = q_load(a, plan[zcols + n_sets - 1], 0, idx, mask);
      fold_last(fp_sqr(zl) - zl);
    }
    for (uint32_t s = 1; s < n_sets; s++) {
      Fr zc = q_load(a, plan[zcols + s], 0, idx, mask);
      Fr zp = q_load(a, plan[zcols + s - 1], last_rot, idx, mask);
      fold_l0(zc - zp);
    }
    // beta * X with X = zeta * omega_ext^idx, then * delta per column
    Fr beta_term = Fr::load_nc(a.t_lo + (idx & ((1u << a.lo_bits) - 1)));
    if (idx >> a.lo_bits) beta_term = beta_term * Fr::load_nc(a.t_hi + (idx >> a.lo_bits));
    Fr current_delta = (a.beta * a.zeta) * beta_term;
    for (uint32_t s = 0; s < n_sets; s++) {
      Fr left = q_load(a, plan[zcols + s], 1, idx, mask);
      Fr right = q_load(a, plan[zcols + s], 0, idx, mask);
      uint32_t j0 = s * chunk_len, j1 = j0 + chunk_len;
      if (j1 > n_pcols) j1 = n_pcols;
      for (uint32_t j = j0; j < j1; j++) {
        Fr v = q_load(a, plan[pcols + 2 * j], 0, idx, mask);
        Fr sg = q_load(a, plan[pcols + 2 * j + 1], 0, idx, mask);
        left = left * (v + a.beta * sg + a.gamma);
        right = right * (v + current_delta + a.gamma);
        current_delta = current_delta * a.delta;
      }
      fold_active(left - right);
    }
  }

  // lookups
  for (uint32_t l = 0; l < n_lookups; l++) {
    Fr cin = q_compress(a, pc, idx, mask);
    Fr ctab = q_compress(a, pc, idx, mask);
    Fr table_value = (cin + a.beta) * (ctab + a.gamma);
    uint32_t zc = plan[pc++], ac = plan[pc++], sc = plan[pc++];
    Fr z = q_load(a, zc, 0, idx, mask);
    Fr z_next = q_load(a, zc, 1, idx, mask);
    Fr ain = q_load(a, ac, 0, idx, mask);
    Fr ain_prev = q_load(a, ac, -1, idx, mask);
    Fr stab = q_load(a, sc, 0, idx, mask);
    Fr a_minus_s = ain - stab;
    fold_l0(one - z);
    fold_last(fp_sqr(z) - z);
    fold_active(fp_mul_sub2(z_next * (ain + a.beta), stab + a.gamma, z, table_value));
    fold_l0(a_minus_s);
    fold_active(a_minus_s * (ain - ain_prev));
  }
  // the three selector columns are only needed now: loading them up front kept 24 registers live through every loop
  // (the kernel sits at the 128-register budget of two CTAs per SM and spilled inside the lookup loop)
  Fr value = acc_plain + fp_mul_add2(acc_l0, q_load(a, plan[6], 0, idx, mask), acc_last, q_load(a, plan[7], 0, idx, mask));
  value = value + acc_active * q_load(a, plan[8], 0, idx, mask);

  if (a.t_evals) value = value * Fr::load_nc(a.t_evals + (idx & a.t_mask));
  value.store(a.out + local);
}

// out[j] = sum_i polys[i][j] * v^(m-1-i)   (Horner over the list: acc = acc * v + poly_i), m <= 64 per launch
struct FoldArgs {
  const Fr* polys[64];
  uint32_t m;
  uint32_t accumulate;  // start from out[] (continuing a longer list) instead of zero
  Fr v;
  Fr* out;
  size_t n;
};

__global__ void __launch_bounds__(256) poly_fold_kernel(const __grid_constant__ FoldArgs a) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  Fr acc;
  uint32_t i = 0;
  if (a.accumulate) acc = Fr::load(a.out + j);
  else acc = Fr::load_nc(a.polys[i++] + j);
  for (; i < a.m; i++) acc = acc * a.v + Fr::load_nc(a.polys[i] + j);
  acc.store(a.out + j);
}

// ---- plan validation on the host: every index the kernel will dereference is checked here ----
static bool plan_sop(const uint32_t* p, size_t n, size_t& pc, size_t n_cols, size_t n_consts) {
  if (pc >= n) return false;
  uint32_t nt = p[pc++];
  for (uint32_t t = 0; t < nt; t++) {
    if (pc + 2 > n) return false;
    uint32_t ci = p[pc++], nf = p[pc++];
    if (ci != QNOCONST && ci >= n_consts) return false;
    if (nf > 64 || pc + nf > n) return false;
    for (uint32_t f = 0; f < nf; f++)
      if ((p[pc++] & 0xffffu) >= n_cols) return false;
  }
  return true;
}

static const char* plan_check(const uint32_t* p, size_t n, size_t n_cols, size_t n_consts) {
  if (n < QPLAN_HEADER || p[0] != QPLAN_MAGIC) return "bad magic / short header";
  uint32_t n_gates = p[1], n_pcols = p[2], chunk_len = p[3], n_lookups = p[5];
  if (p[6] >= n_cols || p[7] >= n_cols || p[8] >= n_cols) return "l_0 / l_last / l_active_row column out of range";
  size_t pc = QPLAN_HEADER;
  for (uint32_t g = 0; g < n_gates; g++)
    if (!plan_sop(p, n, pc, n_cols, n_consts)) return "malformed gate polynomial";
  if (n_pcols) {
    if (chunk_len == 0) return "permutation chunk_len is zero";
    size_t n_sets = (n_pcols + chunk_len - 1) / chunk_len;
    if (pc + 2 * (size_t)n_pcols + n_sets > n) return "permutation section truncated";
    for (size_t i = 0; i < 2 * (size_t)n_pcols + n_sets; i++)
      if (p[pc++] >= n_cols) return "permutation column out of range";
  }
  for (uint32_t l = 0; l < n_lookups; l++) {
    for (int side = 0; side < 2; side++) {
      if (pc >= n) return "lookup section truncated";
      uint32_t ne = p[pc++];
      if (ne == 0) return "lookup with no expressions";
      for (uint32_t e = 0; e < ne; e++)
        if (!plan_sop(p, n, pc, n_cols, n_consts)) return "malformed lookup expression";
    }
    if (pc + 3 > n) return "lookup section truncated";
    for (int i = 0; i < 3; i++)
      if (p[pc++] >= n_cols) return "lookup column out of range";
  }
  if (pc != n) return "trailing words after the last section";
  return nullptr;
}

static int quot_tables(h2agg_ctx* ctx, const uint64_t* omega_ext, uint32_t ext_k, uint32_t* lo_bits_out) {
  uint32_t lo_bits = ext_k < QUOT_LO_BITS ? ext_k : QUOT_LO_BITS, hi_bits = ext_k - lo_bits;
  *lo_bits_out = lo_bits;
  if (ctx->quot_tw.p && ctx->quot_ext_k == ext_k && memcmp(ctx->quot_omega, omega_ext, 32) == 0) return 0;
  size_t total = ((size_t)1 << lo_bits) + ((size_t)1 << hi_bits);
  int rc = ensure(ctx, ctx->quot_tw, total * 32);
  if (rc) return rc;
  Fr w;
  memcpy(w.v, omega_ext, 32);
  Fr* lo = (Fr*)ctx->quot_tw.p;
  quot_gen_tables<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(w, lo_bits, hi_bits, lo, lo + ((size_t)1 << lo_bits));
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  memcpy(ctx->quot_omega, omega_ext, 32);
  ctx->quot_ext_k = ext_k;
  return 0;
}

}  // namespace h2agg

using namespace h2agg;

extern "C" {

static int evaluate_h_impl(h2agg_ctx* ctx, const h2agg_quotient_args* q, uint64_t row_begin, uint64_t row_count,
                           const uint64_t* col_row0, void* d_out);

int h2agg_evaluate_h_dev(h2agg_ctx* ctx, const h2agg_quotient_args* q, void* d_out) {
  if (!ctx) return 1;
  if (!q) { ctx->last_error = "evaluate_h: null argument"; return 1; }
  return evaluate_h_impl(ctx, q, 0, q->ext_k <= 28 ? ((uint64_t)1 << q->ext_k) : 0, nullptr, d_out);
}

int h2agg_evaluate_h_rows_dev(h2agg_ctx* ctx, const h2agg_quotient_args* q, uint64_t row_begin, uint64_t row_count,
                              const uint64_t* col_row0, void* d_out) {
  if (!ctx) return 1;
  if (!q) { ctx->last_error = "evaluate_h: null argument"; return 1; }
  return evaluate_h_impl(ctx, q, row_begin, row_count, col_row0, d_out);
}

}  // extern "C"

static int evaluate_h_impl(h2agg_ctx* ctx, const h2agg_quotient_args* q, uint64_t row_begin, uint64_t row_count,
                           const uint64_t* col_row0, void* d_out) {
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!q || !d_out || !q->plan || !q->d_columns || !q->y || !q->beta || !q->gamma || !q->theta || !q->omega_ext ||
      !q->zeta || !q->delta || (q->n_consts && !q->consts)) {
    ctx->last_error = "evaluate_h: null argument";
    return 1;
  }
  if (q->ext_k < q->k || q->ext_k > 28 || q->k == 0) { ctx->last_error = "evaluate_h: need 0 < k <= ext_k <= 28"; return 1; }
  if (q->n_columns == 0 || q->n_columns > 65535) { ctx->last_error = "evaluate_h: 1..65535 columns"; return 1; }
  if (q->t_evaluations && (q->t_len == 0 || (q->t_len & (q->t_len - 1)) || q->t_len > ((size_t)1 << q->ext_k))) {
    ctx->last_error = "evaluate_h: t_len must be a power of two <= 2^ext_k";
    return 1;
  }
  if (const char* why = plan_check(q->plan, q->n_plan_words, q->n_columns, q->n_consts)) {
    ctx->last_error = std::string("evaluate_h: invalid plan: ") + why;
    return 1;
  }
  for (size_t i = 0; i < q->n_columns; i++)
    if (!q->d_columns[i]) { ctx->last_error = "evaluate_h: null column pointer"; return 1; }
  const uint64_t size = (uint64_t)1 << q->ext_k;
  if (row_begin >= size || row_count > size) { ctx->last_error = "evaluate_h: row window outside the extended domain"; return 1; }
  if (col_row0)
    for (size_t i = 0; i < q->n_columns; i++)
      if (col_row0[i] >= size) { ctx->last_error = "evaluate_h: column window start outside the extended domain"; return 1; }
  if (row_count == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  uint32_t lo_bits;
  int rc = quot_tables(ctx, q->omega_ext, q->ext_k, &lo_bits);
  if (rc) return rc;
  // device copy of { plan | column pointers | constants | t_evaluations }
  size_t off_cols = (q->n_plan_words * 4 + 31) & ~(size_t)31;
  size_t off_consts = (off_cols + q->n_columns * 8 + 31) & ~(size_t)31;
  size_t off_t = off_consts + q->n_consts * 32;
  uint32_t n_terms = q->plan[1] + 5 * q->plan[5];
  if (q->plan[2]) {
    uint32_t n_sets = (q->plan[2] + q->plan[3] - 1) / q->plan[3];
    n_terms += 2 + (n_sets - 1) + n_sets;
  }
  size_t off_y = off_t + q->t_len * 32;
  size_t off_row0 = off_y + (size_t)n_terms * 32 + 32;
  size_t total = off_row0 + (col_row0 ? q->n_columns * 4 : 0);
  // the previous call's kernel may still be reading the old copy: alternate two halves of the buffer
  rc = ensure(ctx, ctx->quot_ws, 2 * total + 64);
  if (rc) return rc;
  ctx->quot_flip ^= 1;
  uint8_t* base = (uint8_t*)ctx->quot_ws.p + (ctx->quot_flip ? ((total + 31) & ~(size_t)31) : 0);
  std::vector<uint8_t> stage(total, 0);
  memcpy(stage.data(), q->plan, q->n_plan_words * 4);
  memcpy(stage.data() + off_cols, q->d_columns, q->n_columns * 8);
  if (q->n_consts) memcpy(stage.data() + off_consts, q->consts, q->n_consts * 32);
  if (q->t_evaluations) memcpy(stage.data() + off_t, q->t_evaluations, q->t_len * 32);
  if (col_row0)
    for (size_t i = 0; i < q->n_columns; i++) {
      uint32_t r0 = (uint32_t)col_row0[i];
      memcpy(stage.data() + off_row0 + 4 * i, &r0, 4);
    }
  H2AGG_CUDA(ctx, cudaMemcpyAsync(base, stage.data(), total, cudaMemcpyHostToDevice, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `stage` is pageable and dies with this frame
  QuotKernelArgs a;
  a.plan = (const uint32_t*)base;
  a.cols = (const Fr* const*)(base + off_cols);
  a.consts = (const Fr*)(base + off_consts);
  a.t_lo = (const Fr*)ctx->quot_tw.p;
  a.t_hi = a.t_lo + ((size_t)1 << lo_bits);
  a.t_evals = q->t_evaluations ? (const Fr*)(base + off_t) : nullptr;
  a.t_mask = q->t_evaluations ? (uint32_t)(q->t_len - 1) : 0;
  a.y_pow = (const Fr*)(base + off_y);
  a.n_terms = n_terms;
  a.out = (Fr*)d_out;
  a.row_begin = (uint32_t)row_begin;
  a.row_count = (uint32_t)row_count;
  a.col_row0 = col_row0 ? (const uint32_t*)(base + off_row0) : nullptr;
  a.lo_bits = lo_bits;
  a.ext_k = q->ext_k;
  a.rot_scale = 1u << (q->ext_k - q->k);
  memcpy(a.y.v, q->y, 32);
  memcpy(a.beta.v, q->beta, 32);
  memcpy(a.gamma.v, q->gamma, 32);
  memcpy(a.theta.v, q->theta, 32);
  memcpy(a.zeta.v, q->zeta, 32);
  memcpy(a.delta.v, q->delta, 32);
  size_t rows = (size_t)row_count;
  if (n_terms) {
    quot_y_powers<<<1, 32, 0, ctx->stream>>>(a.y, n_terms, (Fr*)(base + off_y));
    ctx->launches++;
  }
  {
    ScopedKernelTimer tm(ctx, KC_QUOTIENT);
    quot_evaluate_h<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(a);
  }
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

// out[j] = sum_i w_i * polys[i][j]: the general linear combination (a rank's share of a GWC fold: its own polynomials
// with the powers of v they carry in the full list; weights 1 add the gathered partial folds)
struct LincombArgs {
  const Fr* polys[32];
  Fr w[32];
  uint32_t m;
  uint32_t accumulate;
  Fr* out;
  size_t n;
};

__global__ void __launch_bounds__(256) poly_lincomb_kernel(const __grid_constant__ LincombArgs a) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  Fr acc = a.accumulate ? Fr::load(a.out + j) : Fr::zero();
  for (uint32_t i = 0; i < a.m; i++) acc = acc + a.w[i] * Fr::load_nc(a.polys[i] + j);
  acc.store(a.out + j);
}

extern "C" {

int h2agg_poly_lincomb_dev(h2agg_ctx* ctx, const void* const* d_polys, const uint64_t* weights, size_t n_polys, size_t n,
                           void* d_out) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_out || (n_polys && (!d_polys || !weights))) { ctx->last_error = "poly_lincomb: null argument"; return 1; }
  for (size_t i = 0; i < n_polys; i++)
    if (!d_polys[i]) { ctx->last_error = "poly_lincomb: null polynomial"; return 1; }
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_polys == 0) {
    H2AGG_CUDA(ctx, cudaMemsetAsync(d_out, 0, n * 32, ctx->stream));
    return 0;
  }
  size_t done = 0;
  while (done < n_polys) {
    LincombArgs a;
    a.accumulate = done ? 1 : 0;
    size_t m = n_polys - done;
    if (m > 32) m = 32;
    for (size_t i = 0; i < m; i++) {
      a.polys[i] = (const Fr*)d_polys[done + i];
      memcpy(a.w[i].v, weights + 4 * (done + i), 32);
    }
    a.m = (uint32_t)m;
    a.out = (Fr*)d_out;
    a.n = n;
    poly_lincomb_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    done += m;
  }
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

int h2agg_poly_fold_dev(h2agg_ctx* ctx, const void* const* d_polys, size_t n_polys, size_t n, const uint64_t v[4],
                        void* d_out) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_polys || !v || !d_out || n_polys == 0) { ctx->last_error = "poly_fold: null argument or empty list"; return 1; }
  for (size_t i = 0; i < n_polys; i++)
    if (!d_polys[i]) { ctx->last_error = "poly_fold: null polynomial"; return 1; }
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t done = 0;
  while (done < n_polys) {
    FoldArgs a;
    a.accumulate = done ? 1 : 0;
    size_t m = n_polys - done;
    if (m > 64) m = 64;
    for (size_t i = 0; i < m; i++) a.polys[i] = (const Fr*)d_polys[done + i];
    a.m = (uint32_t)m;
    memcpy(a.v.v, v, 32);
    a.out = (Fr*)d_out;
    a.n = n;
    poly_fold_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    done += m;
  }
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // extern "C"
