// h2agg.hpp -- C++ host-side mirror of the reference-facing surface, over the C ABI (h2agg.h).
//
// The reference is Rust; this image has no Rust toolchain, so the host layer above the C ABI is
// written in C++ with the same names, argument meaning and error behaviour as the Rust functions
// the reference's create_proof reaches (halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994):
//   best_multiexp(coeffs, bases) -> G1            asserts coeffs.len() == bases.len()
//   best_fft(a, omega, log_n)                     asserts a.len() == 1 << log_n, in place
//   ParamsKZG::{commit_lagrange, commit}          blind ignored by KZG
//   EvaluationDomain::{lagrange_to_coeff, coeff_to_extended, extended_to_coeff}
// Types are the Rust memory layouts.  Failures panic (throw), like the Rust
// functions which have no error channel; there is no CPU fallback.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "h2agg.h"

namespace h2agg_host {

struct Fr { uint64_t l[4]; };                 // Montgomery, little-endian limbs
struct Fq { uint64_t l[4]; };
struct G1Affine { Fq x, y; };                 // identity = (0, 0)
struct G1 { Fq x, y, z; };                    // Jacobian, identity z = 0
static_assert(sizeof(Fr) == 32 && sizeof(G1Affine) == 64 && sizeof(G1) == 96, "layout must match halo2curves");

class Context {
 public:
  explicit Context(int device = 0) {
    if (int rc = h2agg_init(device, &ctx_)) throw std::runtime_error(std::string("h2agg_init: ") + h2agg_last_error(nullptr) + " (" + std::to_string(rc) + ")");
  }
  ~Context() { h2agg_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  h2agg_ctx* raw() const { return ctx_; }
  void check(int rc) const {
    if (rc) throw std::runtime_error(std::string("h2agg: ") + h2agg_last_error(ctx_));
  }

 private:
  h2agg_ctx* ctx_ = nullptr;
};

// halo2_proofs::arithmetic::best_multiexp
inline G1 best_multiexp(const Context& c, const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
  if (coeffs.size() != bases.size()) throw std::invalid_argument("best_multiexp: coeffs.len() != bases.len()");
  G1 out;
  c.check(h2agg_msm_g1(c.raw(), 0, reinterpret_cast<const uint64_t*>(bases.data()),
                       reinterpret_cast<const uint64_t*>(coeffs.data()), coeffs.size(), reinterpret_cast<uint64_t*>(&out)));
  return out;
}

// halo2_proofs::arithmetic::best_fft
inline void best_fft(const Context& c, std::vector<Fr>& a, const Fr& omega, uint32_t log_n) {
  if (a.size() != (size_t(1) << log_n)) throw std::invalid_argument("best_fft: a.len() != 1 << log_n");
  c.check(h2agg_ntt_fr(c.raw(), reinterpret_cast<uint64_t*>(a.data()), omega.l, log_n));
}

// halo2_proofs::poly::kzg::commitment::ParamsKZG (the two SRS vectors stay resident in HBM)
class ParamsKZG {
 public:
  ParamsKZG(const Context& c, uint32_t k, const std::vector<G1Affine>& g, const std::vector<G1Affine>& g_lagrange)
      : c_(c), k_(k), n_(size_t(1) << k) {
    if (g.size() != n_ || g_lagrange.size() != n_) throw std::invalid_argument("ParamsKZG: |g| != 2^k");
    c_.check(h2agg_srs_register(c_.raw(), reinterpret_cast<const uint64_t*>(g.data()), n_, &g_id_));
    c_.check(h2agg_srs_register(c_.raw(), reinterpret_cast<const uint64_t*>(g_lagrange.data()), n_, &gl_id_));
  }
  ~ParamsKZG() {
    h2agg_srs_release(c_.raw(), g_id_);
    h2agg_srs_release(c_.raw(), gl_id_);
  }
  G1 commit_lagrange(const std::vector<Fr>& values) const {
    if (values.size() != n_) throw std::invalid_argument("commit_lagrange: poly.len() != n");
    return msm(gl_id_, values);
  }
  G1 commit(const std::vector<Fr>& coeffs) const {
    if (coeffs.size() > n_) throw std::invalid_argument("commit: poly.len() > n");
    return msm(g_id_, coeffs);
  }
  // one commit round (e.g. the 5 advice columns): affine results
  std::vector<G1Affine> commit_lagrange_many(const std::vector<const Fr*>& columns) const {
    std::vector<G1Affine> out(columns.size());
    std::vector<const uint64_t*> ptrs;
    for (auto p : columns) ptrs.push_back(reinterpret_cast<const uint64_t*>(p));
    c_.check(h2agg_msm_g1_batch(c_.raw(), gl_id_, ptrs.data(), ptrs.size(), n_, reinterpret_cast<uint64_t*>(out.data())));
    return out;
  }

 private:
  G1 msm(uint64_t id, const std::vector<Fr>& s) const {
    G1 out;
    c_.check(h2agg_msm_g1(c_.raw(), id, nullptr, reinterpret_cast<const uint64_t*>(s.data()), s.size(), reinterpret_cast<uint64_t*>(&out)));
    return out;
  }
  const Context& c_;
  uint32_t k_;
  size_t n_;
  uint64_t g_id_ = 0, gl_id_ = 0;
};

// halo2_proofs::poly::EvaluationDomain -- the caller supplies the field constants it already holds
// (omega, omega_inv, extended_omega(_inv), ifft divisors, g_coset = Fr::ZETA) exactly as the Rust
// struct stores them, so no field arithmetic happens on the host.
struct EvaluationDomain {
  const Context& c;
  uint32_t k, extended_k;
  Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, ifft_divisor, extended_ifft_divisor;
  uint32_t quotient_poly_degree;

  void lagrange_to_coeff(std::vector<Fr>& a) const {
    if (a.size() != (size_t(1) << k)) throw std::invalid_argument("lagrange_to_coeff: wrong length");
    c.check(h2agg_intt_fr(c.raw(), reinterpret_cast<uint64_t*>(a.data()), omega_inv.l, ifft_divisor.l, k));
  }
  std::vector<Fr> coeff_to_extended(const std::vector<Fr>& a) const {
    if (a.size() != (size_t(1) << k)) throw std::invalid_argument("coeff_to_extended: wrong length");
    std::vector<Fr> out(size_t(1) << extended_k);
    c.check(h2agg_coeff_to_extended(c.raw(), reinterpret_cast<const uint64_t*>(a.data()), k, extended_k, g_coset.l,
                                    extended_omega.l, reinterpret_cast<uint64_t*>(out.data())));
    return out;
  }
  void extended_to_coeff(std::vector<Fr>& a) const {
    if (a.size() != (size_t(1) << extended_k)) throw std::invalid_argument("extended_to_coeff: wrong length");
    size_t out_len = (size_t(1) << k) * quotient_poly_degree;
    c.check(h2agg_extended_to_coeff(c.raw(), reinterpret_cast<uint64_t*>(a.data()), extended_k, extended_omega_inv.l,
                                    extended_ifft_divisor.l, g_coset.l, out_len));
    a.resize(out_len);  // halo2 truncates to n * (j - 1)
  }
};

// halo2_proofs::arithmetic::{eval_polynomial, kate_division} and plonk::lookup::prover::permute_expression_pair
// (usable rows only; the caller appends its blinding rows) -- host-vector forms of the "next" rows.
inline Fr eval_polynomial(const Context& c, const std::vector<Fr>& poly, const Fr& point) {
  Fr out;
  c.check(h2agg_eval_polynomial(c.raw(), reinterpret_cast<const uint64_t*>(poly.data()), poly.size(), point.l, out.l));
  return out;
}
inline std::vector<Fr> kate_division(const Context& c, const std::vector<Fr>& a, const Fr& b) {
  if (a.empty()) throw std::invalid_argument("kate_division: empty polynomial");
  std::vector<Fr> q(a.size() - 1);
  c.check(h2agg_kate_division(c.raw(), reinterpret_cast<const uint64_t*>(a.data()), a.size(), b.l,
                              reinterpret_cast<uint64_t*>(q.data())));
  return q;
}
// returns false where halo2 returns Error::ConstraintSystemFailure (an input value is not in the table)
inline bool permute_expression_pair(const Context& c, const std::vector<Fr>& input, const std::vector<Fr>& table,
                                    std::vector<Fr>& permuted_input, std::vector<Fr>& permuted_table) {
  if (input.size() != table.size()) throw std::invalid_argument("permute_expression_pair: lengths differ");
  permuted_input.resize(input.size());
  permuted_table.resize(input.size());
  int rc = h2agg_permute_expression_pair(c.raw(), reinterpret_cast<const uint64_t*>(input.data()),
                                         reinterpret_cast<const uint64_t*>(table.data()), input.size(),
                                         reinterpret_cast<uint64_t*>(permuted_input.data()),
                                         reinterpret_cast<uint64_t*>(permuted_table.data()));
  if (rc == 4) return false;
  c.check(rc);
  return true;
}

}  // namespace h2agg_host
