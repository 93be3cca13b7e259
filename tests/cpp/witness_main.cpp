// Compiled-host driver of the witness path: replays the op mix of an aggregation of `n_proofs` inner proofs -- Poseidon
// transcript (T = 9), ScalarChip expression mix, one scalar_mul_constant + add per instance, two multi_exps, the final
// pair -- call by call through the recording C ABI (include/h2agg.h), exactly the calls the reference's chips make
// through `verify_aggregation_proofs_in_chip` (halo2-snark-aggregator-api/src/systems/halo2/verify.rs:835-942) and the
// same stream halo2_snark_aggregator_b200/witness_workload.py drives from Python.  The reference's host is compiled
// code, so this is the representative cost of the host side: a chip call is a C function call, not a ctypes round trip.
// Values are arbitrary (the row layout does not depend on them); points come from a file of affine Montgomery points.
//
// usage: witness_main <points.bin> <n_proofs> [expand_k]     (expand_k: also run the expansion kernels into 5 device
//        columns of 2^expand_k rows; needs a B200)          -> one JSON line on stdout
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "h2agg.h"

typedef unsigned __int128 u128;

static const uint64_t R_MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};

struct Rng {  // splitmix64; canonical scalars by masking to 253 bits (always < r)
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  void scalar(uint64_t out[4]) {
    for (int i = 0; i < 4; i++) out[i] = next();
    out[3] &= 0x1fffffffffffffffULL;
    (void)R_MOD;
  }
};

struct Chips {
  h2agg_witness* w;
  Rng rng;
  std::vector<uint64_t> points;  // n x 8
  int64_t ok(int64_t h) {
    if (h < 0) throw std::runtime_error(std::string("recorder: ") + h2agg_wit_error(w));
    return h;
  }
  const uint64_t* point(size_t i) const { return points.data() + 8 * (i % (points.size() / 8)); }
  // ScalarChip
  int64_t assign_var() { uint64_t v[4]; rng.scalar(v); return ok(h2agg_wit_assign_scalar(w, v)); }
  int64_t assign_const() { uint64_t v[4]; rng.scalar(v); return ok(h2agg_wit_field_assign_const(w, v)); }
  int64_t add(int64_t a, int64_t b) { return ok(h2agg_wit_field_add(w, a, b)); }
  int64_t sub(int64_t a, int64_t b) { return ok(h2agg_wit_field_sub(w, a, b)); }
  int64_t mul(int64_t a, int64_t b) { return ok(h2agg_wit_field_mul(w, a, b)); }
  int64_t mul_add_constant(int64_t a, int64_t b) { uint64_t c[4]; rng.scalar(c); return ok(h2agg_wit_field_mul_add_constant(w, a, b, c)); }
  int64_t sum(const std::vector<int64_t>& e, bool random_coeffs, bool random_constant) {
    std::vector<uint64_t> co(4 * e.size() + 4, 0);
    for (size_t i = 0; i < e.size(); i++) {
      if (random_coeffs) rng.scalar(co.data() + 4 * i);
      else co[4 * i] = 1;
    }
    uint64_t c[4] = {0, 0, 0, 0};
    if (random_constant) rng.scalar(c);
    return ok(h2agg_wit_field_sum_with_coeff_and_constant(w, e.data(), co.data(), e.size(), c));
  }
  int64_t sum2(int64_t s0, int64_t x) {  // [(s0, c), (x, 1)]
    int64_t e[2] = {s0, x};
    uint64_t co[8] = {0, 0, 0, 0, 1, 0, 0, 0}, c[4] = {0, 0, 0, 0};
    rng.scalar(co);
    return ok(h2agg_wit_field_sum_with_coeff_and_constant(w, e, co, 2, c));
  }
  // EccChip / Encode
  int64_t assign_point(size_t i) { return ok(h2agg_wit_assign_point(w, point(i))); }
  int64_t ecc_add(int64_t a, int64_t b) { return ok(h2agg_wit_ecc_add(w, a, b)); }
  int64_t normalize(int64_t a) { return ok(h2agg_wit_ecc_reduce(w, a)); }
  int64_t constant_mul(int64_t s, size_t i) { return ok(h2agg_wit_ecc_constant_mul(w, point(i), s)); }
  void encode_point(int64_t p, std::vector<int64_t>& out) {
    int64_t n2[2];
    if (h2agg_wit_encode_point(w, p, n2) != 0) ok(-1);
    out.push_back(n2[0]);
    out.push_back(n2[1]);
  }
};

// PoseidonChip's op pattern (halo2-snark-aggregator-api/src/hash/poseidon.rs:150-231) with random constants
struct Poseidon {
  static const int T = 9, R_F = 8, R_P = 63;
  Chips& c;
  std::vector<int64_t> s, absorbing;
  explicit Poseidon(Chips& chips) : c(chips) {
    for (int i = 0; i < T; i++) s.push_back(c.assign_const());
  }
  void update(const std::vector<int64_t>& e) { absorbing.insert(absorbing.end(), e.begin(), e.end()); }
  int64_t x5(int64_t x) {
    int64_t x2 = c.mul(x, x), x4 = c.mul(x2, x2);
    return c.mul_add_constant(x, x4);
  }
  void full() {
    for (auto& x : s) x = x5(x);
    std::vector<int64_t> n;
    for (int i = 0; i < T; i++) n.push_back(c.sum(s, true, false));
    s = n;
  }
  void partial() {
    s[0] = x5(s[0]);
    std::vector<int64_t> n;
    n.push_back(c.sum(s, true, false));
    for (int i = 1; i < T; i++) n.push_back(c.sum2(s[0], s[i]));
    s = n;
  }
  void permutation(const std::vector<int64_t>& in) {
    s[0] = c.sum({s[0]}, false, true);
    for (size_t i = 0; i < in.size(); i++) s[i + 1] = c.sum({s[i + 1], in[i]}, false, true);
    for (size_t i = in.size() + 1; i < (size_t)T; i++) s[i] = c.sum({s[i]}, false, true);
    for (int r = 0; r < R_F / 2; r++) full();
    for (int r = 0; r < R_P; r++) partial();
    for (int r = 0; r < R_F / 2; r++) full();
  }
  int64_t squeeze() {
    std::vector<int64_t> in;
    in.swap(absorbing);
    size_t pad = 0;
    for (size_t i = 0; i < in.size(); i += T - 1) {
      std::vector<int64_t> chunk(in.begin() + i, in.begin() + std::min(in.size(), i + T - 1));
      pad = T - 1 - chunk.size();
      permutation(chunk);
    }
    if (pad == 0) permutation({});
    return s[1];
  }
};

static void record(Chips& c, int n_proofs, int points_per_proof = 23, int scalars_per_proof = 71, int field_ops_per_proof = 600,
                   int instances_per_proof = 2) {
  std::vector<int64_t> all_pts, all_scs;
  for (int p = 0; p < n_proofs; p++) {
    Poseidon tr(c);
    std::vector<int64_t> inst;
    for (int i = 0; i < instances_per_proof; i++) inst.push_back(c.assign_var());
    int64_t acc = -1;
    for (size_t i = 0; i < inst.size(); i++) {  // assign_instance_commitment: scalar_mul_constant + add
      int64_t ls = c.constant_mul(inst[i], 1000 * p + i);
      acc = acc < 0 ? ls : c.ecc_add(acc, ls);
    }
    std::vector<int64_t> pts = {c.normalize(acc)}, enc, scs;
    c.encode_point(pts[0], enc);
    tr.update(enc);
    for (int i = 0; i < points_per_proof - 1; i++) {  // read_point: assign_var (on-curve check) + encode
      int64_t pt = c.assign_point(1000 * p + 100 + i);
      enc.clear();
      c.encode_point(pt, enc);
      tr.update(enc);
      pts.push_back(pt);
      if (i % 5 == 4) scs.push_back(tr.squeeze());
    }
    for (int i = 0; i < scalars_per_proof; i++) {  // read_scalar
      int64_t s = c.assign_var();
      tr.update({s});
      scs.push_back(s);
    }
    scs.push_back(tr.squeeze());
    int64_t acc_s = scs[0];
    for (int i = 0; i < field_ops_per_proof; i++) {  // the mul / add / sub mix of the expression evaluation
      int64_t o = scs[(7 * i + 3) % scs.size()];
      acc_s = i % 3 == 0 ? c.mul(acc_s, o) : (i % 3 == 1 ? c.add(acc_s, o) : c.sub(acc_s, o));
    }
    all_pts.insert(all_pts.end(), pts.begin(), pts.end());
    for (size_t i = 0; i < pts.size(); i++) all_scs.push_back(c.mul(acc_s, scs[i % scs.size()]));
  }
  int64_t w_g = c.ok(h2agg_wit_ecc_shamir(c.w, all_pts.data(), all_scs.data(), all_pts.size()));
  int64_t w_x = c.ok(h2agg_wit_ecc_shamir(c.w, all_pts.data(), all_scs.data(), 4));
  int64_t cells[4];
  if (h2agg_wit_expose_final_pair(c.w, w_x, w_g, cells) != 0) c.ok(-1);
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s points.bin n_proofs [expand_k]\n", argv[0]);
    return 2;
  }
  try {
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) throw std::runtime_error("cannot open the points file");
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (raw.size() < 64 || raw.size() % 64) throw std::runtime_error("points file: need n x 64 bytes");
    const int n_proofs = atoi(argv[2]);
    const int expand_k = argc > 3 ? atoi(argv[3]) : 0;
    h2agg_ctx* ctx = nullptr;
    void* d_cols[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (expand_k) {
      if (h2agg_init(0, &ctx) != 0) throw std::runtime_error(std::string("h2agg_init: ") + h2agg_last_error(nullptr));
      for (int c = 0; c < 5; c++)
        if (h2agg_dev_alloc(ctx, (size_t)32 << expand_k, &d_cols[c]) != 0) throw std::runtime_error("dev_alloc");
    }
    double rec_s = 0, exp_s = 0;
    uint64_t rows = 0, ops = 0;
    const int reps = expand_k ? 4 : 2;  // the first pass is untimed: it page-locks / faults in the record chunks
    for (int rep = 0; rep < reps; rep++) {
      Chips c;
      c.w = h2agg_wit_new();
      c.rng.s = 0x1234 + 1;
      c.points.resize(raw.size() / 8);
      memcpy(c.points.data(), raw.data(), raw.size());
      const double t0 = now();
      record(c, n_proofs);
      const double t1 = now();
      rows = h2agg_wit_rows(c.w);
      ops = h2agg_wit_ops(c.w);
      if (expand_k) {
        if (rows > ((uint64_t)1 << expand_k)) throw std::runtime_error("expand_k too small for the recorded layout");
        if (h2agg_witness_expand_dev(ctx, c.w, d_cols, (size_t)1 << expand_k) != 0 || h2agg_synchronize(ctx) != 0)
          throw std::runtime_error(std::string("expand: ") + h2agg_last_error(ctx));
      }
      const double t2 = now();
      if (rep > 0) {
        rec_s += t1 - t0;
        exp_s += t2 - t1;
      }
      h2agg_wit_free(c.w);
    }
    rec_s /= reps - 1;
    exp_s /= reps - 1;
    printf("{\"n_proofs\": %d, \"rows\": %llu, \"op_records\": %llu, \"host_record_s\": %.6f, \"expand_incl_h2d_s\": %s, \"rows_per_s_total\": %.1f, "
           "\"host_threads\": %d}\n",
           n_proofs, (unsigned long long)rows, (unsigned long long)ops, rec_s, expand_k ? std::to_string(exp_s).c_str() : "null",
           rows / (rec_s + exp_s), h2agg_wit_set_threads(0));
    if (ctx) h2agg_destroy(ctx);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "witness_main: %s\n", e.what());
    return 1;
  }
}
