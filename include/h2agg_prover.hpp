// h2agg_prover.hpp -- C++ host-side driver of the device-resident create_proof pipeline, over the C ABI (h2agg.h).
//
// The reference drives proving from Rust (halo2_proofs plonk/prover.rs `create_proof`, reached from
// halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994); this image has no Rust toolchain, so the host
// layer is C++.  ResidentProver issues, in halo2's order, the calls a patched `create_proof` would issue between
// its transcript operations (the Python twin is halo2_snark_aggregator_b200/prover.py):
//     commit_columns        round 1: witness columns arrive from host memory, stay resident in three forms
//     lookup_round          round 2: compress_expressions, permute_expression_pair, blinding rows, commit
//     product_round         round 3: permutation products (chained sets), lookup products, blinding rows, commit
//     commit_coeff          the vanishing argument's random polynomial
//     quotient              evaluate_h / (X^n - 1), extended_to_coeff, commit the pieces of h
//     fold_h, evaluate      the evaluation round (one batched call per opening point)
//     open                  GWC: fold with v, kate_division, commit, per point
// No field arithmetic happens on the host: everything that depends on a challenge (x^n, x * omega^rot,
// beta * delta^i) is passed in by the caller, who owns the transcript and the field type, exactly as the constants of
// EvaluationDomain are (h2agg.hpp).  Challenges and blinding values are inputs.  Failures throw.
#pragma once
#include <cstring>
#include <map>

#include "h2agg.hpp"

namespace h2agg_host {

struct ProverShape {
  uint32_t k = 0, ext_k = 0, blinding_factors = 0, chunk_len = 0, quotient_pieces = 0;
  uint32_t n_columns = 0;                     // columns of the quotient plan, ids 0 .. n_columns-1 (plonk.py order)
  std::vector<uint32_t> plan;                 // h2agg_evaluate_h_dev word program
  std::vector<Fr> plan_consts;
  std::vector<Fr> t_evaluations;              // 1 / ((zeta omega_ext^i)^n - 1)
  Fr omega, omega_inv, n_inv, omega_ext, omega_ext_inv, ext_n_inv, zeta, delta;
  std::vector<uint32_t> expr_columns;         // plan ids of the Lagrange columns the lookup expressions index into
  struct Lookup {
    std::vector<uint32_t> input_exprs, table_exprs;  // { n_exprs, POLY... } over expr_columns
    std::vector<Fr> input_consts, table_consts;
    uint32_t z_col = 0, input_col = 0, table_col = 0;  // plan ids of z, a', s'
  };
  std::vector<Lookup> lookups;
  std::vector<uint32_t> perm_values, perm_sigmas, perm_z;  // plan ids, permutation column order / set order
  // ids beyond the plan's columns
  uint32_t random_id() const { return n_columns; }
  uint32_t h_id() const { return n_columns + 1; }
  uint32_t h_piece_id(uint32_t i) const { return n_columns + 2 + i; }
};

struct Query {
  uint32_t poly;   // column id
  int32_t rotation;
};

class ResidentProver {
 public:
  ResidentProver(const Context& c, const ProverShape& s, uint64_t srs_lagrange, uint64_t srs_g)
      : c_(c), s_(s), gl_(srs_lagrange), g_(srs_g), n_(size_t(1) << s.k), ext_n_(size_t(1) << s.ext_k) {
    if (s.k == 0 || s.ext_k < s.k || s.quotient_pieces == 0 || ((size_t)s.quotient_pieces << s.k) > ext_n_)
      throw std::invalid_argument("ResidentProver: bad shape");
  }
  ~ResidentProver() {
    h2agg_synchronize(c_.raw());
    for (void* p : owned_) h2agg_dev_free(c_.raw(), p);
  }
  ResidentProver(const ResidentProver&) = delete;
  ResidentProver& operator=(const ResidentProver&) = delete;

  // Deferred transforms (h2agg_set_defer_transforms): the NTT passes of every commit round run on the context's
  // background stream and fill the latency-bound stretches of the following rounds; they are joined where the quotient /
  // the evaluation round / the opening first read a coefficient or extended form.  On by default, like the Python twin.
  bool defer_transforms = true;
  void join_transforms() { c_.check(h2agg_transforms_join(c_.raw())); }

  size_t usable_rows() const { return n_ - (s_.blinding_factors + 1); }
  void* lagrange(uint32_t id) { return slot(lag_, id, n_ * 32); }
  void* coeff(uint32_t id) { return slot(coeff_, id, n_ * 32); }
  void* extended(uint32_t id) { return slot(ext_, id, ext_n_ * 32); }

  // round 1 (and keygen): HOST Lagrange columns -> commitments; the three forms stay resident under `ids`
  std::vector<G1Affine> commit_columns(const std::vector<uint32_t>& ids, const std::vector<const Fr*>& cols, bool with_extended = true,
                                       bool keep_lagrange = true) {
    if (ids.size() != cols.size()) throw std::invalid_argument("commit_columns: ids / columns mismatch");
    std::vector<const uint64_t*> src;
    std::vector<void*> lo, co, eo;
    for (size_t i = 0; i < ids.size(); i++) {
      src.push_back(reinterpret_cast<const uint64_t*>(cols[i]));
      lo.push_back((keep_lagrange || defer_transforms) ? lagrange(ids[i]) : nullptr);   // deferred passes read the column later
      co.push_back(coeff(ids[i]));
      eo.push_back(with_extended ? extended(ids[i]) : nullptr);
    }
    std::vector<G1Affine> out(ids.size());
    Deferred scope(*this);
    c_.check(h2agg_commit_round_resident(c_.raw(), gl_, src.data(), ids.size(), s_.k, s_.omega_inv.l, s_.n_inv.l,
                                         reinterpret_cast<uint64_t*>(out.data()), lo.data(), co.data(), with_extended ? s_.ext_k : 0,
                                         with_extended ? s_.zeta.l : nullptr, with_extended ? s_.omega_ext.l : nullptr,
                                         with_extended ? eo.data() : nullptr));
    return out;
  }

  // commit round for Lagrange columns already in HBM under lagrange(id)
  std::vector<G1Affine> commit_device_columns(const std::vector<uint32_t>& ids) {
    std::vector<const void*> src;
    std::vector<void*> co, eo;
    for (uint32_t id : ids) {
      src.push_back(lagrange(id));
      co.push_back(coeff(id));
      eo.push_back(extended(id));
    }
    std::vector<G1Affine> out(ids.size());
    Deferred scope(*this);
    c_.check(h2agg_commit_round_dev(c_.raw(), gl_, src.data(), ids.size(), s_.k, s_.omega_inv.l, s_.n_inv.l,
                                    reinterpret_cast<uint64_t*>(out.data()), co.data(), s_.ext_k, s_.zeta.l, s_.omega_ext.l, eo.data()));
    return out;
  }

  // round 2.  blinds: per lookup two vectors (permuted input, permuted table) of blinding_factors + 1 values
  std::vector<G1Affine> lookup_round(const Fr& theta, const std::vector<std::vector<Fr>>& blinds) {
    const size_t L = s_.lookups.size(), u = usable_rows(), tail = s_.blinding_factors + 1;
    if (blinds.size() != 2 * L) throw std::invalid_argument("lookup_round: two blinding vectors per lookup");
    std::vector<const void*> cols;
    for (uint32_t id : s_.expr_columns) cols.push_back(lagrange(id));
    std::vector<uint32_t> out_ids;
    for (size_t i = 0; i < L; i++) {
      const ProverShape::Lookup& lk = s_.lookups[i];
      void* a = scratch(key_compressed(i, 0), n_ * 32);
      void* t = scratch(key_compressed(i, 1), n_ * 32);
      c_.check(h2agg_compress_expressions_dev(c_.raw(), lk.input_exprs.data(), lk.input_exprs.size(), cols.data(), cols.size(),
                                              lk.input_consts.empty() ? nullptr : lk.input_consts[0].l, lk.input_consts.size(), s_.k,
                                              theta.l, a));
      c_.check(h2agg_compress_expressions_dev(c_.raw(), lk.table_exprs.data(), lk.table_exprs.size(), cols.data(), cols.size(),
                                              lk.table_consts.empty() ? nullptr : lk.table_consts[0].l, lk.table_consts.size(), s_.k,
                                              theta.l, t));
      void* pa = lagrange(lk.input_col);
      void* pt = lagrange(lk.table_col);
      c_.check(h2agg_permute_expression_pair_dev(c_.raw(), a, t, u, pa, pt));
      for (int side = 0; side < 2; side++) {
        const std::vector<Fr>& b = blinds[2 * i + side];
        if (b.size() != tail) throw std::invalid_argument("lookup_round: blinding vector length");
        c_.check(h2agg_memcpy_h2d(c_.raw(), (uint8_t*)(side ? pt : pa) + 32 * u, b.data(), tail * 32));
      }
      out_ids.push_back(lk.input_col);
      out_ids.push_back(lk.table_col);
    }
    return commit_device_columns(out_ids);
  }

  // round 3.  beta_delta_start[s] = beta * delta^(s * chunk_len); blinds: per z column (permutation sets first, then
  // lookups) blinding_factors values
  std::vector<G1Affine> product_round(const Fr& beta, const Fr& gamma, const std::vector<Fr>& beta_delta_start,
                                      const std::vector<std::vector<Fr>>& blinds) {
    const size_t sets = s_.perm_z.size(), L = s_.lookups.size(), bf = s_.blinding_factors, u = usable_rows();
    if (beta_delta_start.size() != sets || blinds.size() != sets + L) throw std::invalid_argument("product_round: argument sizes");
    std::vector<uint32_t> out_ids;
    const void* last = nullptr;
    for (size_t st = 0; st < sets; st++) {
      std::vector<const void*> vals, sigs;
      for (size_t j = st * s_.chunk_len; j < std::min<size_t>((st + 1) * s_.chunk_len, s_.perm_values.size()); j++) {
        vals.push_back(lagrange(s_.perm_values[j]));
        sigs.push_back(lagrange(s_.perm_sigmas[j]));
      }
      void* z = lagrange(s_.perm_z[st]);
      c_.check(h2agg_permutation_product_dev(c_.raw(), vals.data(), sigs.data(), vals.size(), s_.k, s_.omega.l,
                                             beta_delta_start[st].l, s_.delta.l, beta.l, gamma.l, last, z));
      put_tail(z, blinds[st], bf);
      last = (const uint8_t*)z + 32 * u;
      out_ids.push_back(s_.perm_z[st]);
    }
    std::vector<const void*> in_a, in_s, in_ap, in_sp;
    std::vector<void*> zs;
    for (size_t i = 0; i < L; i++) {
      const ProverShape::Lookup& lk = s_.lookups[i];
      in_a.push_back(scratch(key_compressed(i, 0), n_ * 32));
      in_s.push_back(scratch(key_compressed(i, 1), n_ * 32));
      in_ap.push_back(lagrange(lk.input_col));
      in_sp.push_back(lagrange(lk.table_col));
      zs.push_back(lagrange(lk.z_col));
    }
    // all lookup products in one call: their latency chains overlap on the lanes
    c_.check(h2agg_lookup_products_dev(c_.raw(), L, in_a.data(), in_s.data(), in_ap.data(), in_sp.data(), n_, beta.l, gamma.l, zs.data()));
    for (size_t i = 0; i < L; i++) {
      put_tail(zs[i], blinds[sets + i], bf);
      out_ids.push_back(s_.lookups[i].z_col);
    }
    return commit_device_columns(out_ids);
  }

  // a polynomial given in COEFFICIENT form (the random polynomial of the vanishing argument): ParamsKZG::commit
  G1Affine commit_coeff(uint32_t id, const Fr* coeffs) {
    void* d = coeff(id);
    c_.check(h2agg_memcpy_h2d(c_.raw(), d, coeffs, n_ * 32));
    return commit_dev({d})[0];
  }

  std::vector<G1Affine> quotient(const Fr& y, const Fr& beta, const Fr& gamma, const Fr& theta) {
    std::vector<const void*> cols;
    for (uint32_t id = 0; id < s_.n_columns; id++) {
      auto it = ext_.find(id);
      if (it == ext_.end()) throw std::logic_error("quotient: column " + std::to_string(id) + " has no extended form yet");
      cols.push_back(it->second);
    }
    join_transforms();
    void* h = scratch(KEY_H, ext_n_ * 32);
    h2agg_quotient_args a;
    memset(&a, 0, sizeof(a));
    a.k = s_.k;
    a.ext_k = s_.ext_k;
    a.plan = s_.plan.data();
    a.n_plan_words = s_.plan.size();
    a.d_columns = cols.data();
    a.n_columns = cols.size();
    a.consts = s_.plan_consts.empty() ? nullptr : s_.plan_consts[0].l;
    a.n_consts = s_.plan_consts.size();
    a.y = y.l; a.beta = beta.l; a.gamma = gamma.l; a.theta = theta.l;
    a.omega_ext = s_.omega_ext.l; a.zeta = s_.zeta.l; a.delta = s_.delta.l;
    a.t_evaluations = s_.t_evaluations[0].l;
    a.t_len = s_.t_evaluations.size();
    c_.check(h2agg_evaluate_h_dev(c_.raw(), &a, h));
    c_.check(h2agg_extended_to_coeff_dev(c_.raw(), h, s_.ext_k, s_.omega_ext_inv.l, s_.ext_n_inv.l, s_.zeta.l,
                                         (size_t)s_.quotient_pieces << s_.k));
    std::vector<const void*> pieces;
    for (uint32_t i = 0; i < s_.quotient_pieces; i++) {
      void* p = (uint8_t*)h + (size_t)i * n_ * 32;
      coeff_[s_.h_piece_id(i)] = p;
      pieces.push_back(p);
    }
    return commit_dev(pieces);
  }

  // h(X) = sum_i x^(n i) h_i(X), kept under h_id(); xn = x^n
  void fold_h(const Fr& xn) {
    std::vector<const void*> pieces;
    for (uint32_t i = s_.quotient_pieces; i-- > 0;) pieces.push_back(coeff_.at(s_.h_piece_id(i)));
    void* d = scratch(KEY_H_FOLDED, n_ * 32);
    c_.check(h2agg_poly_fold_dev(c_.raw(), pieces.data(), pieces.size(), n_, xn.l, d));
    coeff_[s_.h_id()] = d;
  }

  // points: rotation -> x * omega^rotation.  One batched call per opening point; results in query order
  std::vector<Fr> evaluate(const std::vector<Query>& queries, const std::map<int32_t, Fr>& points) {
    join_transforms();
    void* d_ev = scratch(KEY_EVALS, 32 * std::max<size_t>(queries.size(), 1));
    std::map<int32_t, std::vector<size_t>> groups;
    for (size_t i = 0; i < queries.size(); i++) groups[queries[i].rotation].push_back(i);
    std::vector<size_t> where(queries.size());
    size_t off = 0;
    for (auto& kv : groups) {
      std::vector<const void*> polys;
      for (size_t i : kv.second) polys.push_back(coeff_.at(queries[i].poly));
      c_.check(h2agg_eval_polynomials_dev(c_.raw(), polys.data(), polys.size(), n_, points.at(kv.first).l, (uint8_t*)d_ev + 32 * off));
      for (size_t j = 0; j < kv.second.size(); j++) where[kv.second[j]] = off + j;
      off += kv.second.size();
    }
    std::vector<Fr> flat(queries.size()), out(queries.size());
    c_.check(h2agg_memcpy_d2h(c_.raw(), flat.data(), d_ev, queries.size() * 32));
    for (size_t i = 0; i < queries.size(); i++) out[i] = flat[where[i]];
    return out;
  }

  // GWC: points in order of first appearance; W = commit(kate_division(fold_v(polys at the point), point)).
  // (The constant eval_batch halo2 subtracts only changes the remainder kate_division drops.)
  std::vector<G1Affine> open(const std::vector<Query>& queries, const std::map<int32_t, Fr>& points, const Fr& v,
                             std::vector<int32_t>* order_out = nullptr) {
    join_transforms();
    std::vector<int32_t> order;
    std::map<int32_t, std::vector<const void*>> groups;
    for (const Query& q : queries) {
      if (!groups.count(q.rotation)) order.push_back(q.rotation);
      groups[q.rotation].push_back(coeff_.at(q.poly));
    }
    void* fold = scratch(KEY_FOLD, n_ * 32);
    std::vector<const void*> ws;
    for (size_t j = 0; j < order.size(); j++) {
      void* w = scratch(KEY_W + j, n_ * 32);
      // GWC folds query i of a point with v^i (halo2_proofs gwc/prover.rs zips the queries with powers(v); the reference's
      // verifier rebuilds exactly that, halo2-snark-aggregator-api/src/systems/halo2/multiopen.rs:55-61):
      // poly_fold gives its first argument the highest power, so the group goes in reversed
      std::vector<const void*> g(groups[order[j]].rbegin(), groups[order[j]].rend());
      c_.check(h2agg_poly_fold_dev(c_.raw(), g.data(), g.size(), n_, v.l, fold));
      c_.check(h2agg_kate_division_dev(c_.raw(), fold, n_, points.at(order[j]).l, w));
      ws.push_back(w);
    }
    if (order_out) *order_out = order;
    return commit_dev(ws);
  }

 private:
  // scope of a commit round issued with deferred transforms (the setting is per context, so it is switched on only here)
  struct Deferred {
    ResidentProver& p;
    explicit Deferred(ResidentProver& pr) : p(pr) {
      if (p.defer_transforms) h2agg_set_defer_transforms(p.c_.raw(), 1);
    }
    ~Deferred() {
      if (p.defer_transforms) h2agg_set_defer_transforms(p.c_.raw(), 0);
    }
  };
  enum : uint64_t { KEY_H = 1, KEY_H_FOLDED, KEY_EVALS, KEY_FOLD, KEY_PTS, KEY_W = 100, KEY_COMPRESSED = 1000 };
  static uint64_t key_compressed(size_t lookup, int side) { return KEY_COMPRESSED + 2 * lookup + side; }

  void* alloc(size_t bytes) {
    void* p = nullptr;
    c_.check(h2agg_dev_alloc(c_.raw(), bytes, &p));
    owned_.push_back(p);
    return p;
  }
  void* slot(std::map<uint32_t, void*>& m, uint32_t id, size_t bytes) {
    auto it = m.find(id);
    if (it != m.end()) return it->second;
    return m[id] = alloc(bytes);
  }
  void* scratch(uint64_t key, size_t bytes) {
    auto it = scratch_.find(key);
    if (it != scratch_.end()) return it->second;
    return scratch_[key] = alloc(bytes);
  }
  void put_tail(void* z, const std::vector<Fr>& b, size_t bf) {
    if (b.size() != bf) throw std::invalid_argument("product_round: blinding vector length");
    if (bf) c_.check(h2agg_memcpy_h2d(c_.raw(), (uint8_t*)z + 32 * (n_ - bf), b.data(), bf * 32));
  }
  std::vector<G1Affine> commit_dev(const std::vector<const void*>& polys) {
    void* d_out = scratch(KEY_PTS, 160 * 64);
    if (polys.size() > 64) throw std::invalid_argument("commit_dev: at most 64 polynomials per call");
    c_.check(h2agg_msm_g1_batch_dev(c_.raw(), g_, nullptr, polys.data(), polys.size(), n_, d_out));
    std::vector<uint64_t> raw(20 * polys.size());
    c_.check(h2agg_memcpy_d2h(c_.raw(), raw.data(), d_out, raw.size() * 8));
    std::vector<G1Affine> out(polys.size());
    for (size_t i = 0; i < polys.size(); i++) memcpy(&out[i], raw.data() + 20 * i, 64);
    return out;
  }

  const Context& c_;
  ProverShape s_;
  uint64_t gl_, g_;
  size_t n_, ext_n_;
  std::map<uint32_t, void*> lag_, coeff_, ext_;
  std::map<uint64_t, void*> scratch_;
  std::vector<void*> owned_;
};

}  // namespace h2agg_host
