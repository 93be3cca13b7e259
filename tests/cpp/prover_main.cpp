// Test driver for include/h2agg_prover.hpp: runs the whole resident pipeline from a section file written by
// tests/test_gpu_prover.py and dumps every commitment / evaluation, which the test compares bit for bit with the
// Python twin (prover.py) and, through it, with the CPU oracles.  usage: prover_main <in.bin> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "h2agg_prover.hpp"

using namespace h2agg_host;
typedef std::vector<uint64_t> Sec;

static std::map<uint64_t, std::vector<Sec>> read_sections(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open input");
  std::map<uint64_t, std::vector<Sec>> out;
  uint64_t hdr[2];
  while (f.read(reinterpret_cast<char*>(hdr), 16)) {
    Sec s(hdr[1]);
    if (hdr[1]) f.read(reinterpret_cast<char*>(s.data()), hdr[1] * 8);
    out[hdr[0]].push_back(std::move(s));
  }
  return out;
}
static std::vector<Fr> frs(const Sec& s) {
  std::vector<Fr> v(s.size() / 4);
  if (!v.empty()) memcpy(v.data(), s.data(), v.size() * 32);
  return v;
}
static std::vector<uint32_t> u32s(const Sec& s) { return std::vector<uint32_t>(s.begin(), s.end()); }

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  try {
    auto sec = read_sections(argv[1]);
    auto one = [&](uint64_t tag) -> const Sec& { return sec.at(tag).at(0); };
    ProverShape sh;
    const Sec& h = one(1);
    sh.k = h[0]; sh.ext_k = h[1]; sh.blinding_factors = h[2]; sh.chunk_len = h[3]; sh.quotient_pieces = h[4]; sh.n_columns = h[5];
    sh.plan = u32s(one(2));
    sh.plan_consts = frs(one(3));
    auto fc = frs(one(4));
    sh.omega = fc[0]; sh.omega_inv = fc[1]; sh.n_inv = fc[2]; sh.omega_ext = fc[3]; sh.omega_ext_inv = fc[4]; sh.ext_n_inv = fc[5];
    sh.zeta = fc[6]; sh.delta = fc[7];
    sh.t_evaluations = frs(one(5));
    sh.expr_columns = u32s(one(6));
    const size_t L = sec.count(7) ? sec[7].size() : 0;
    for (size_t i = 0; i < L; i++) {
      ProverShape::Lookup lk;
      lk.input_exprs = u32s(sec[7][i]); lk.input_consts = frs(sec[8][i]);
      lk.table_exprs = u32s(sec[9][i]); lk.table_consts = frs(sec[10][i]);
      lk.z_col = sec[11][i][0]; lk.input_col = sec[11][i][1]; lk.table_col = sec[11][i][2];
      sh.lookups.push_back(lk);
    }
    sh.perm_values = u32s(one(12)); sh.perm_sigmas = u32s(one(13)); sh.perm_z = u32s(one(14));
    const size_t n = size_t(1) << sh.k;

    Context ctx(0);
    uint64_t gl = 0, g = 0;
    ctx.check(h2agg_srs_register(ctx.raw(), one(15).data(), n, &gl));
    ctx.check(h2agg_srs_register(ctx.raw(), one(16).data(), n, &g));
    std::vector<uint64_t> out;
    auto dump_pts = [&](const std::vector<G1Affine>& p) { for (auto& q : p) out.insert(out.end(), (const uint64_t*)&q, (const uint64_t*)&q + 8); };
    {
      ResidentProver pr(ctx, sh, gl, g);
      auto cols_of = [&](uint64_t tag) { std::vector<const Fr*> v; for (auto& s : sec[tag]) v.push_back(reinterpret_cast<const Fr*>(s.data())); return v; };
      dump_pts(pr.commit_columns(u32s(one(17)), cols_of(18)));      // keygen-side columns
      dump_pts(pr.commit_columns(u32s(one(19)), cols_of(20)));      // round 1
      auto ch = frs(one(21));                                        // theta, beta, gamma, y, v
      std::vector<std::vector<Fr>> b2, b3;
      for (auto& s : sec[23]) b2.push_back(frs(s));
      for (auto& s : sec[24]) b3.push_back(frs(s));
      dump_pts(pr.lookup_round(ch[0], b2));
      dump_pts(pr.product_round(ch[1], ch[2], frs(one(22)), b3));
      dump_pts({pr.commit_coeff(sh.random_id(), reinterpret_cast<const Fr*>(one(25).data()))});
      dump_pts(pr.quotient(ch[3], ch[1], ch[2], ch[0]));
      pr.fold_h(frs(one(26))[0]);
      std::vector<Query> qs;
      const Sec& q = one(27);
      for (size_t i = 0; i + 1 < q.size(); i += 2) qs.push_back(Query{(uint32_t)q[i], (int32_t)(int64_t)q[i + 1]});
      std::map<int32_t, Fr> pts;
      const Sec& p = one(28);
      for (size_t i = 0; i + 4 < p.size(); i += 5) { Fr f; memcpy(f.l, &p[i + 1], 32); pts[(int32_t)(int64_t)p[i]] = f; }
      for (auto& e : pr.evaluate(qs, pts)) out.insert(out.end(), e.l, e.l + 4);
      dump_pts(pr.open(qs, pts, ch[4]));
    }
    h2agg_srs_release(ctx.raw(), gl);
    h2agg_srs_release(ctx.raw(), g);
    std::ofstream o(argv[2], std::ios::binary);
    o.write(reinterpret_cast<const char*>(out.data()), out.size() * 8);
    printf("prover_main: %zu words written\n", out.size());
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "prover_main: %s\n", e.what());
    return 1;
  }
}
