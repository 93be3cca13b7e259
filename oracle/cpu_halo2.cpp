// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// CPU restatement (C++17, 4 x 64-bit Montgomery limbs, unsigned __int128) of the algorithms the
// reference's hot path runs.  The arithmetic of that path is NOT under /root/reference: it lives
// in the un-vendored git dependencies
//     halo2_proofs 0.2.0  scroll-tech/halo2 @ 3370852d (Cargo.lock:1549-1551; manifests: PSE tag v2022_09_10)
//     halo2curves  0.2.1  @ f75ed26c                   (Cargo.lock:1569-1571)
// so this file restates their published algorithms and anchors on the
// reference's call sites:
//     best_multiexp   <- ParamsKZG::commit_lagrange/commit <- create_proof
//                        halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994, keygen_vk :760-761
//     best_fft / EvaluationDomain::{ifft, coeff_to_extended, extended_to_coeff}
//                     <- create_proof :986, keygen_pk :974
// PARITY UNPINNED against reference-held vectors: the reference's tests never look at a commitment
// or an FFT output, and the reference cannot be built here (no cargo/rustc).  The
// oracle is instead pinned against an independent Python big-integer implementation
// (oracle/py/bn254_ref.py -> tests/golden/*.json) and the algebraic identities in tests/.
//
// Deliberately shares NO code with the CUDA path (different limb width, different Montgomery
// schedule, Jacobian instead of XYZZ coordinates, unsigned windows instead of signed digits).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;

namespace {

struct FrP {
  static constexpr u64 P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr u64 R[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
  static constexpr u64 R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
  static constexpr u64 INV = 0xc2e1f593efffffffULL;
};
struct FqP {
  static constexpr u64 P[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr u64 R[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
  static constexpr u64 R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
  static constexpr u64 INV = 0x87d20782e4866389ULL;
};

template <class M>
struct Fp {
  u64 v[4];
  static Fp zero() { return Fp{{0, 0, 0, 0}}; }
  static Fp one() { return Fp{{M::R[0], M::R[1], M::R[2], M::R[3]}}; }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  bool operator==(const Fp& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
  bool operator!=(const Fp& o) const { return !(*this == o); }
};

template <class M>
static inline bool geq_p(const u64* a) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > M::P[i]) return true;
    if (a[i] < M::P[i]) return false;
  }
  return true;
}
template <class M>
static inline void sub_p(u64* a) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - M::P[i] - (u64)b;
    a[i] = (u64)d;
    b = (d >> 64) & 1;
  }
}
template <class M>
static inline Fp<M> add(const Fp<M>& a, const Fp<M>& b) {
  Fp<M> r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.v[i] + b.v[i];
    r.v[i] = (u64)c;
    c >>= 64;
  }
  if (geq_p<M>(r.v)) sub_p<M>(r.v);
  return r;
}
template <class M>
static inline Fp<M> sub(const Fp<M>& a, const Fp<M>& b) {
  Fp<M> r;
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a.v[i] - b.v[i] - borrow;
    r.v[i] = (u64)d;
    borrow = (u64)(d >> 64) & 1;
  }
  if (borrow) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)r.v[i] + M::P[i];
      r.v[i] = (u64)c;
      c >>= 64;
    }
  }
  return r;
}
template <class M>
static inline Fp<M> neg(const Fp<M>& a) {
  return sub(Fp<M>::zero(), a);
}
template <class M>
static inline Fp<M> dbl(const Fp<M>& a) {
  return add(a, a);
}
// Montgomery product (coarsely integrated operand scanning)
template <class M>
static inline Fp<M> mul(const Fp<M>& a, const Fp<M>& b) {
  u64 t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.v[j] * b.v[i] + t[j];
      t[j] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (u64)c;
    t[5] = (u64)(c >> 64);
    u64 m = t[0] * M::INV;
    c = ((u128)m * M::P[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * M::P[j] + t[j];
      t[j - 1] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (u64)c;
    t[4] = t[5] + (u64)(c >> 64);
  }
  Fp<M> r{{t[0], t[1], t[2], t[3]}};
  if (t[4] || geq_p<M>(r.v)) sub_p<M>(r.v);
  return r;
}
template <class M>
static inline Fp<M> sqr(const Fp<M>& a) {
  return mul(a, a);
}
template <class M>
static inline Fp<M> to_mont(const u64* canon) {
  Fp<M> a{{canon[0], canon[1], canon[2], canon[3]}};
  Fp<M> r2{{M::R2[0], M::R2[1], M::R2[2], M::R2[3]}};
  return mul(a, r2);
}
template <class M>
static inline void from_mont(const Fp<M>& a, u64* canon) {
  Fp<M> o{{1, 0, 0, 0}};
  Fp<M> r = mul(a, o);
  memcpy(canon, r.v, 32);
}
template <class M>
static Fp<M> pow_words(const Fp<M>& a, const u64* e) {
  Fp<M> r = Fp<M>::one();
  for (int i = 255; i >= 0; i--) {
    r = sqr(r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = mul(r, a);
  }
  return r;
}
template <class M>
static Fp<M> inv(const Fp<M>& a) {
  u64 e[4] = {M::P[0] - 2, M::P[1], M::P[2], M::P[3]};
  return pow_words(a, e);
}

typedef Fp<FrP> Fr;
typedef Fp<FqP> Fq;

// ---- G1: affine {x,y} (identity = (0,0)), Jacobian {x,y,z} (identity z = 0) as in halo2curves ----
struct Aff {
  Fq x, y;
  bool is_identity() const { return x.is_zero() && y.is_zero(); }
};
struct Jac {
  Fq x, y, z;
  static Jac identity() { return Jac{Fq::zero(), Fq::one(), Fq::zero()}; }
  bool is_identity() const { return z.is_zero(); }
};

static Jac jac_double(const Jac& p) {  // dbl-2009-l
  if (p.is_identity()) return p;
  Fq a = sqr(p.x), b = sqr(p.y), c = sqr(b);
  Fq d = dbl(sub(sub(sqr(add(p.x, b)), a), c));
  Fq e = add(dbl(a), a), f = sqr(e);
  Jac r;
  r.z = dbl(mul(p.y, p.z));
  r.x = sub(f, dbl(d));
  Fq c8 = dbl(dbl(dbl(c)));
  r.y = sub(mul(e, sub(d, r.x)), c8);
  return r;
}
static Jac jac_add(const Jac& p, const Jac& q) {  // add-2007-bl
  if (p.is_identity()) return q;
  if (q.is_identity()) return p;
  Fq z1z1 = sqr(p.z), z2z2 = sqr(q.z);
  Fq u1 = mul(p.x, z2z2), u2 = mul(q.x, z1z1);
  Fq s1 = mul(mul(p.y, q.z), z2z2), s2 = mul(mul(q.y, p.z), z1z1);
  if (u1 == u2) {
    if (s1 == s2) return jac_double(p);
    return Jac::identity();
  }
  Fq h = sub(u2, u1), i = sqr(dbl(h)), j = mul(h, i), r = dbl(sub(s2, s1)), v = mul(u1, i);
  Jac o;
  o.x = sub(sub(sqr(r), j), dbl(v));
  o.y = sub(mul(r, sub(v, o.x)), dbl(mul(s1, j)));
  o.z = mul(sub(sub(sqr(add(p.z, q.z)), z1z1), z2z2), h);
  return o;
}
static Jac jac_add_mixed(const Jac& p, const Aff& q) {  // madd-2007-bl
  if (q.is_identity()) return p;
  if (p.is_identity()) return Jac{q.x, q.y, Fq::one()};
  Fq z1z1 = sqr(p.z), u2 = mul(q.x, z1z1), s2 = mul(mul(q.y, p.z), z1z1);
  if (p.x == u2) {
    if (p.y == s2) return jac_double(p);
    return Jac::identity();
  }
  Fq h = sub(u2, p.x), hh = sqr(h), i = dbl(dbl(hh)), j = mul(h, i), r = dbl(sub(s2, p.y)), v = mul(p.x, i);
  Jac o;
  o.x = sub(sub(sqr(r), j), dbl(v));
  o.y = sub(mul(r, sub(v, o.x)), dbl(mul(p.y, j)));
  o.z = sub(sub(sqr(add(p.z, h)), z1z1), hh);
  return o;
}
static Aff jac_to_affine(const Jac& p) {
  if (p.is_identity()) return Aff{Fq::zero(), Fq::zero()};
  Fq zi = inv(p.z), zi2 = sqr(zi);
  return Aff{mul(p.x, zi2), mul(mul(p.y, zi2), zi)};
}
static void write_normalised(const Jac& p, u64* out12) {
  Aff a = jac_to_affine(p);
  if (p.is_identity()) {
    Jac id = Jac::identity();
    memcpy(out12, &id, 96);
    return;
  }
  Fq one = Fq::one();
  memcpy(out12, a.x.v, 32);
  memcpy(out12 + 4, a.y.v, 32);
  memcpy(out12 + 8, one.v, 32);
}

// ---- best_multiexp ------------------------------------------------------------
struct Bucket {
  int kind = 0;  // 0 None, 1 Affine, 2 Projective
  Aff a;
  Jac j;
  void add_assign(const Aff& other) {
    if (kind == 0) { kind = 1; a = other; }
    else if (kind == 1) { j = jac_add_mixed(Jac{a.x, a.y, a.is_identity() ? Fq::zero() : Fq::one()}, other); kind = 2; }
    else j = jac_add_mixed(j, other);
  }
  Jac add(const Jac& other) const {
    if (kind == 0) return other;
    if (kind == 1) return jac_add_mixed(other, a);
    return jac_add(other, j);
  }
};

static inline u64 get_at(size_t segment, size_t c, const uint8_t* bytes) {
  size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
  if (skip_bytes >= 32) return 0;
  uint8_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < 8 && skip_bytes + i < 32; i++) v[i] = bytes[skip_bytes + i];
  u64 tmp;
  memcpy(&tmp, v, 8);
  tmp >>= (skip_bits - skip_bytes * 8);
  return tmp % ((u64)1 << c);
}

static void multiexp_serial(const Fr* coeffs, const Aff* bases, size_t len, Jac* acc) {
  std::vector<uint8_t> repr(len * 32);
  for (size_t i = 0; i < len; i++) from_mont(coeffs[i], (u64*)&repr[i * 32]);  // to_repr()
  size_t c;
  if (len < 4) c = 1;
  else if (len < 32) c = 3;
  else c = (size_t)std::ceil(std::log((double)len));
  size_t segments = 256 / c + 1;
  std::vector<Bucket> buckets(((size_t)1 << c) - 1);
  for (size_t seg = segments; seg-- > 0;) {
    for (size_t k = 0; k < c; k++) *acc = jac_double(*acc);
    for (auto& b : buckets) b.kind = 0;
    for (size_t i = 0; i < len; i++) {
      u64 d = get_at(seg, c, &repr[i * 32]);
      if (d != 0) buckets[d - 1].add_assign(bases[i]);
    }
    Jac running = Jac::identity();
    for (size_t b = buckets.size(); b-- > 0;) {
      running = buckets[b].add(running);
      *acc = jac_add(*acc, running);
    }
  }
}

static Jac best_multiexp(const Fr* coeffs, const Aff* bases, size_t len, unsigned threads) {
  if (threads < 1) threads = 1;
  if (len > threads) {
    size_t chunk = len / threads;
    size_t nchunks = (len + chunk - 1) / chunk;
    std::vector<Jac> results(nchunks, Jac::identity());
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nchunks; t++) {
      size_t lo = t * chunk, hi = std::min(len, lo + chunk);
      pool.emplace_back([&, t, lo, hi] { multiexp_serial(coeffs + lo, bases + lo, hi - lo, &results[t]); });
    }
    for (auto& th : pool) th.join();
    Jac acc = Jac::identity();
    for (auto& r : results) acc = jac_add(acc, r);
    return acc;
  }
  Jac acc = Jac::identity();
  multiexp_serial(coeffs, bases, len, &acc);
  return acc;
}

// ---- best_fft -------------------------------------------------------------------
static inline uint32_t bitreverse(uint32_t n, uint32_t l) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < l; i++) {
    r = (r << 1) | (n & 1);
    n >>= 1;
  }
  return r;
}

static void recursive_butterfly(Fr* a, size_t n, size_t twiddle_chunk, const Fr* tw, int spawn_depth) {
  if (n == 2) {
    Fr t = a[1];
    a[1] = a[0];
    a[0] = add(a[0], t);
    a[1] = sub(a[1], t);
    return;
  }
  Fr* left = a;
  Fr* right = a + n / 2;
  if (spawn_depth > 0) {
    std::thread th([&] { recursive_butterfly(left, n / 2, twiddle_chunk * 2, tw, spawn_depth - 1); });
    recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, spawn_depth - 1);
    th.join();
  } else {
    recursive_butterfly(left, n / 2, twiddle_chunk * 2, tw, 0);
    recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, 0);
  }
  {  // twiddle factor one
    Fr t = right[0];
    right[0] = left[0];
    left[0] = add(left[0], t);
    right[0] = sub(right[0], t);
  }
  auto body = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      Fr t = mul(right[i], tw[i * twiddle_chunk]);
      right[i] = left[i];
      left[i] = add(left[i], t);
      right[i] = sub(right[i], t);
    }
  };
  size_t half = n / 2;
  if (spawn_depth > 0 && half >= 4096) {  // the rayon pool also splits these combine loops
    unsigned parts = 1u << spawn_depth;
    std::vector<std::thread> pool;
    size_t per = (half - 1 + parts - 1) / parts;
    for (unsigned p = 0; p < parts; p++) {
      size_t lo = 1 + p * per, hi = std::min(half, lo + per);
      if (lo < hi) pool.emplace_back(body, lo, hi);
    }
    for (auto& th : pool) th.join();
  } else {
    body(1, half);
  }
}

static void best_fft(Fr* a, const Fr& omega, uint32_t log_n, unsigned threads) {
  if (threads < 1) threads = 1;
  uint32_t log_threads = 0;
  while ((2u << log_threads) <= threads) log_threads++;
  size_t n = (size_t)1 << log_n;
  for (size_t k = 0; k < n; k++) {
    size_t rk = bitreverse((uint32_t)k, log_n);
    if (k < rk) std::swap(a[rk], a[k]);
  }
  std::vector<Fr> tw(n / 2 ? n / 2 : 1);
  {
    Fr w = Fr::one();
    for (size_t i = 0; i < n / 2; i++) {
      tw[i] = w;
      w = mul(w, omega);
    }
  }
  if (log_n == 0) return;
  if (log_n <= log_threads) {
    size_t chunk = 2, twiddle_chunk = n / 2;
    for (uint32_t s = 0; s < log_n; s++) {
      for (size_t base = 0; base < n; base += chunk) {
        Fr* left = a + base;
        Fr* right = a + base + chunk / 2;
        {
          Fr t = right[0];
          right[0] = left[0];
          left[0] = add(left[0], t);
          right[0] = sub(right[0], t);
        }
        for (size_t i = 1; i < chunk / 2; i++) {
          Fr t = mul(right[i], tw[i * twiddle_chunk]);
          right[i] = left[i];
          left[i] = add(left[i], t);
          right[i] = sub(right[i], t);
        }
      }
      chunk *= 2;
      twiddle_chunk /= 2;
    }
  } else {
    recursive_butterfly(a, n, 1, tw.data(), (int)log_threads);
  }
}

template <class F>
static void parallelize(size_t n, unsigned threads, F f) {
  if (threads <= 1 || n < 1024) {
    f(0, n);
    return;
  }
  std::vector<std::thread> pool;
  size_t per = (n + threads - 1) / threads;
  for (unsigned t = 0; t < threads; t++) {
    size_t lo = t * per, hi = std::min(n, lo + per);
    if (lo < hi) pool.emplace_back(f, lo, hi);
  }
  for (auto& th : pool) th.join();
}

static void distribute_powers_zeta(Fr* a, size_t n, const Fr& zeta, bool into_coset, unsigned threads) {
  Fr z2 = sqr(zeta);
  Fr pw[2] = {into_coset ? zeta : z2, into_coset ? z2 : zeta};
  parallelize(n, threads, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      size_t m = i % 3;
      if (m) a[i] = mul(a[i], pw[m - 1]);
    }
  });
}

// ---- deterministic synthetic inputs, counter-based so any slice can be generated
static inline u64 splitmix64(u64& s) {
  u64 z = (s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static inline u64 stream_seed(u64 seed, u64 index) {
  u64 s = seed ^ (index * 0xd1342543de82ef95ULL + 0x2545f4914f6cdd1dULL);
  splitmix64(s);
  return s;
}
template <class M>
static void draw_below_modulus(u64& s, u64* out) {  // uniform in [0, p) by rejection on 254 bits
  for (;;) {
    for (int i = 0; i < 4; i++) out[i] = splitmix64(s);
    out[3] &= 0x3fffffffffffffffULL;
    if (!geq_p<M>(out)) return;
  }
}

static Fq fq_sqrt_candidate(const Fq& a) {  // p = 3 mod 4: a^((p+1)/4)
  static const u64 e[4] = {0x4f082305b61f3f52ULL, 0x65e05aa45a1c72a3ULL, 0x6e14116da0605617ULL, 0x0c19139cb84c680aULL};
  return pow_words(a, e);
}

}  // namespace

extern "C" {

// kind 0: uniform Fr; 1: witness-like small columns a0..a3 (70% 17-bit, 10% {0,1}, 20% zero);
// 2: witness-like wide column a4 (50% 68-bit, 20% full width, 30% zero); 3: uniform 17-bit (permuted lookup columns)
void oracle_gen_scalars(uint64_t seed, int kind, size_t first, size_t n, uint64_t* out, unsigned threads) {
  parallelize(n, threads, [&](size_t lo, size_t hi) {
    for (size_t k = lo; k < hi; k++) {
      u64 s = stream_seed(seed, first + k);
      u64 c[4] = {0, 0, 0, 0};
      u64 sel = splitmix64(s) % 100;
      if (kind == 0) {
        draw_below_modulus<FrP>(s, c);
      } else if (kind == 1) {
        if (sel < 70) c[0] = splitmix64(s) & 0x1ffff;
        else if (sel < 80) c[0] = splitmix64(s) & 1;
      } else if (kind == 2) {
        if (sel < 50) { c[0] = splitmix64(s); c[1] = splitmix64(s) & 0xf; }
        else if (sel < 70) draw_below_modulus<FrP>(s, c);
      } else {
        c[0] = splitmix64(s) & 0x1ffff;
      }
      Fr m = to_mont<FrP>(c);
      memcpy(out + 4 * k, m.v, 32);
    }
  });
}

// bases: try-and-increment on y^2 = x^3 + 3, even canonical y  (BN254 G1 has cofactor 1)
void oracle_gen_bases(uint64_t seed, size_t first, size_t n, uint64_t* out, unsigned threads) {
  u64 three_c[4] = {3, 0, 0, 0};
  Fq three = to_mont<FqP>(three_c);
  parallelize(n, threads, [&](size_t lo, size_t hi) {
    for (size_t k = lo; k < hi; k++) {
      u64 s = stream_seed(seed, first + k);
      u64 xc[4];
      draw_below_modulus<FqP>(s, xc);
      Fq x = to_mont<FqP>(xc);
      for (;;) {
        Fq rhs = add(mul(sqr(x), x), three);
        Fq y = fq_sqrt_candidate(rhs);
        if (sqr(y) == rhs && !y.is_zero()) {
          u64 yc[4];
          from_mont(y, yc);
          if (yc[0] & 1) y = neg(y);
          memcpy(out + 8 * k, x.v, 32);
          memcpy(out + 8 * k + 4, y.v, 32);
          break;
        }
        x = add(x, Fq::one());
      }
    }
  });
}

void oracle_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (field == 0) {
      Fr x, y = Fr::zero(), r;
      memcpy(x.v, a + 4 * i, 32);
      if (b) memcpy(y.v, b + 4 * i, 32);
      r = op == 0 ? add(x, y) : op == 1 ? sub(x, y) : op == 2 ? inv(x) : mul(x, y);
      memcpy(out + 4 * i, r.v, 32);
    } else {
      Fq x, y = Fq::zero(), r;
      memcpy(x.v, a + 4 * i, 32);
      if (b) memcpy(y.v, b + 4 * i, 32);
      r = op == 0 ? add(x, y) : op == 1 ? sub(x, y) : op == 2 ? inv(x) : mul(x, y);
      memcpy(out + 4 * i, r.v, 32);
    }
  }
}
void oracle_to_mont(int field, const uint64_t* canon, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (field == 0) { Fr r = to_mont<FrP>(canon + 4 * i); memcpy(out + 4 * i, r.v, 32); }
    else { Fq r = to_mont<FqP>(canon + 4 * i); memcpy(out + 4 * i, r.v, 32); }
  }
}
void oracle_from_mont(int field, const uint64_t* mont, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (field == 0) { Fr x; memcpy(x.v, mont + 4 * i, 32); from_mont(x, out + 4 * i); }
    else { Fq x; memcpy(x.v, mont + 4 * i, 32); from_mont(x, out + 4 * i); }
  }
}

// halo2 best_multiexp -> normalised Jacobian (x, y, 1) / identity (0, 1, 0)
void oracle_best_multiexp(const uint64_t* scalars, const uint64_t* bases, size_t n, unsigned threads, uint64_t* out12) {
  Jac r = best_multiexp((const Fr*)scalars, (const Aff*)bases, n, threads);
  write_normalised(r, out12);
}
// naive sum of double-and-add products (independent of the bucket method)
void oracle_msm_naive(const uint64_t* scalars, const uint64_t* bases, size_t n, uint64_t* out12) {
  Jac acc = Jac::identity();
  for (size_t i = 0; i < n; i++) {
    Fr s;
    memcpy(s.v, scalars + 4 * i, 32);
    u64 c[4];
    from_mont(s, c);
    Aff b;
    memcpy(&b, bases + 8 * i, 64);
    Jac t = Jac::identity();
    for (int bit = 255; bit >= 0; bit--) {
      t = jac_double(t);
      if ((c[bit >> 6] >> (bit & 63)) & 1) t = jac_add_mixed(t, b);
    }
    acc = jac_add(acc, t);
  }
  write_normalised(acc, out12);
}
void oracle_g1_sum(const uint64_t* pts12, size_t m, uint64_t* out12) {
  Jac acc = Jac::identity();
  for (size_t i = 0; i < m; i++) {
    Jac p;
    memcpy(&p, pts12 + 12 * i, 96);
    acc = jac_add(acc, p);
  }
  write_normalised(acc, out12);
}
int oracle_g1_on_curve(const uint64_t* aff8) {
  Aff a;
  memcpy(&a, aff8, 64);
  if (a.is_identity()) return 1;
  u64 three_c[4] = {3, 0, 0, 0};
  return sqr(a.y) == add(mul(sqr(a.x), a.x), to_mont<FqP>(three_c)) ? 1 : 0;
}

void oracle_best_fft(uint64_t* a, const uint64_t* omega, uint32_t log_n, unsigned threads) {
  Fr w;
  memcpy(w.v, omega, 32);
  best_fft((Fr*)a, w, log_n, threads);
}
void oracle_dft_naive(const uint64_t* a, const uint64_t* omega, uint32_t log_n, uint64_t* out) {
  size_t n = (size_t)1 << log_n;
  Fr w;
  memcpy(w.v, omega, 32);
  std::vector<Fr> pw(n);
  pw[0] = Fr::one();
  for (size_t i = 1; i < n; i++) pw[i] = mul(pw[i - 1], w);
  for (size_t k = 0; k < n; k++) {
    Fr acc = Fr::zero();
    for (size_t j = 0; j < n; j++) {
      Fr x;
      memcpy(x.v, a + 4 * j, 32);
      acc = add(acc, mul(x, pw[(j * k) & (n - 1)]));
    }
    memcpy(out + 4 * k, acc.v, 32);
  }
}
// EvaluationDomain::ifft: best_fft(a, omega_inv) then * divisor
void oracle_ifft(uint64_t* a, const uint64_t* omega_inv, const uint64_t* divisor, uint32_t log_n, unsigned threads) {
  Fr w, d;
  memcpy(w.v, omega_inv, 32);
  memcpy(d.v, divisor, 32);
  Fr* p = (Fr*)a;
  best_fft(p, w, log_n, threads);
  parallelize((size_t)1 << log_n, threads, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) p[i] = mul(p[i], d);
  });
}
// EvaluationDomain::coeff_to_extended
void oracle_coeff_to_extended(const uint64_t* coeffs, uint32_t k, uint32_t ext_k, const uint64_t* zeta,
                              const uint64_t* omega_ext, uint64_t* out, unsigned threads) {
  size_t n = (size_t)1 << k, en = (size_t)1 << ext_k;
  Fr z, w;
  memcpy(z.v, zeta, 32);
  memcpy(w.v, omega_ext, 32);
  memcpy(out, coeffs, n * 32);
  distribute_powers_zeta((Fr*)out, n, z, true, threads);
  memset(out + 4 * n, 0, (en - n) * 32);
  best_fft((Fr*)out, w, ext_k, threads);
}
// EvaluationDomain::extended_to_coeff (caller truncates to out_len)
void oracle_extended_to_coeff(uint64_t* a, uint32_t ext_k, const uint64_t* omega_ext_inv, const uint64_t* ext_n_inv,
                              const uint64_t* zeta, unsigned threads) {
  Fr z, w, d;
  memcpy(z.v, zeta, 32);
  memcpy(w.v, omega_ext_inv, 32);
  memcpy(d.v, ext_n_inv, 32);
  Fr* p = (Fr*)a;
  size_t en = (size_t)1 << ext_k;
  best_fft(p, w, ext_k, threads);
  parallelize(en, threads, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) p[i] = mul(p[i], d);
  });
  distribute_powers_zeta(p, en, z, false, threads);
}

// halo2_proofs::arithmetic::eval_polynomial (Horner; the parallel version splits into chunks and
// recombines with powers of the point -- same value)
void oracle_eval_polynomial(const uint64_t* poly, size_t n, const uint64_t* point, uint64_t* out) {
  Fr x, acc = Fr::zero();
  memcpy(x.v, point, 32);
  for (size_t i = n; i-- > 0;) {
    Fr c;
    memcpy(c.v, poly + 4 * i, 32);
    acc = add(mul(acc, x), c);
  }
  memcpy(out, acc.v, 32);
}
// halo2_proofs::arithmetic::kate_division: b = -b; tmp = 0; for (q, r) in q.rev().zip(a.rev()):
//   lead = r - tmp; q = lead; tmp = lead * b      -> q has a.len() - 1 entries
void oracle_kate_division(const uint64_t* a, size_t n, const uint64_t* b_in, uint64_t* q) {
  if (n < 2) return;
  Fr b;
  memcpy(b.v, b_in, 32);
  b = neg(b);
  Fr tmp = Fr::zero();
  for (size_t k = 0; k + 1 < n; k++) {
    size_t qi = n - 2 - k, ai = n - 1 - k;
    Fr lead;
    memcpy(lead.v, a + 4 * ai, 32);
    lead = sub(lead, tmp);
    memcpy(q + 4 * qi, lead.v, 32);
    tmp = mul(lead, b);
  }
}

// halo2_proofs BatchInvert: zeros are skipped (stay zero), everything else is inverted
void oracle_batch_invert(uint64_t* a, size_t n) {
  std::vector<Fr> pre(n);
  Fr run = Fr::one();
  Fr* p = (Fr*)a;
  for (size_t i = 0; i < n; i++) {
    pre[i] = run;
    if (!p[i].is_zero()) run = mul(run, p[i]);
  }
  Fr iv = inv(run);
  for (size_t i = n; i-- > 0;) {
    if (p[i].is_zero()) continue;
    Fr v = p[i];
    p[i] = mul(iv, pre[i]);
    iv = mul(iv, v);
  }
}
// running product column: z[0] = 1, z[i+1] = z[i] * num[i] / den[i]
void oracle_grand_product(const uint64_t* num, const uint64_t* den, size_t n, uint64_t* z) {
  std::vector<uint64_t> d(den, den + 4 * n);
  oracle_batch_invert(d.data(), n);
  Fr run = Fr::one();
  for (size_t i = 0; i < n; i++) {
    memcpy(z + 4 * i, run.v, 32);
    Fr a, b;
    memcpy(a.v, num + 4 * i, 32);
    memcpy(b.v, d.data() + 4 * i, 32);
    run = mul(run, mul(a, b));
  }
}

// halo2_proofs plonk/lookup/prover.rs `permute_expression_pair` (external crate) over the usable rows,
// WITHOUT the random blinding rows the caller appends.  Restated step by step: sort the input (Fr::cmp = numeric order
// of the canonical value), count the table values in an ordered map (BTreeMap), give every first occurrence of an input
// value its own table cell and take one instance out of the map, remember the rows of repeated inputs, then hand the
// leftover table values out in ascending order to the remembered rows popped from the BACK.
// Returns 0, or 1 when an input value does not occur in the table (halo2: Error::ConstraintSystemFailure).
struct CanonLess {
  bool operator()(const std::array<u64, 4>& a, const std::array<u64, 4>& b) const {
    for (int i = 3; i >= 0; i--)
      if (a[i] != b[i]) return a[i] < b[i];
    return false;
  }
};
int oracle_permute_expression_pair(const uint64_t* input, const uint64_t* table, size_t usable_rows, uint64_t* permuted_input,
                                   uint64_t* permuted_table) {
  const size_t u = usable_rows;
  auto canon = [](const uint64_t* p) {
    Fr a;
    memcpy(a.v, p, 32);
    std::array<u64, 4> r;
    from_mont(a, r.data());
    return r;
  };
  auto mont = [](const std::array<u64, 4>& c, uint64_t* out) {
    Fr m = to_mont<FrP>(c.data());
    memcpy(out, m.v, 32);
  };
  std::vector<std::array<u64, 4>> in(u);
  for (size_t i = 0; i < u; i++) in[i] = canon(input + 4 * i);
  std::sort(in.begin(), in.end(), CanonLess());
  std::map<std::array<u64, 4>, uint32_t, CanonLess> leftover;
  for (size_t i = 0; i < u; i++) leftover[canon(table + 4 * i)]++;
  std::vector<std::array<u64, 4>> tab(u, std::array<u64, 4>{0, 0, 0, 0});
  std::vector<size_t> repeated;
  for (size_t row = 0; row < u; row++) {
    if (row == 0 || in[row] != in[row - 1]) {
      tab[row] = in[row];
      auto it = leftover.find(in[row]);
      if (it == leftover.end() || it->second == 0) return 1;
      it->second--;
    } else {
      repeated.push_back(row);
    }
  }
  for (auto& kv : leftover)
    for (uint32_t c = 0; c < kv.second; c++) {
      if (repeated.empty()) return 2;  // cannot happen: |table| = |input|
      tab[repeated.back()] = kv.first;
      repeated.pop_back();
    }
  if (!repeated.empty()) return 2;
  for (size_t i = 0; i < u; i++) {
    mont(in[i], permuted_input + 4 * i);
    mont(tab[i], permuted_table + 4 * i);
  }
  return 0;
}

// halo2_proofs plonk/lookup/prover.rs compress_expressions: each expression of the list (word layout of
// include/h2agg.h: { n_exprs, POLY x n_exprs }) is evaluated on the Lagrange domain -- rotation r of row i reads row
// (i + r) mod n -- and folded acc = acc * theta + e.  Row-parallel like halo2's `parallelize`.
void oracle_compress_expressions(const uint32_t* exprs, const uint64_t* const* cols, const uint64_t* consts, uint32_t k,
                                 const uint64_t* theta_, uint64_t* out, unsigned threads) {
  const size_t n = (size_t)1 << k;
  Fr theta;
  memcpy(theta.v, theta_, 32);
  const Fr one = Fr::one();
  parallelize(n, threads, [&](size_t lo, size_t hi) {
    for (size_t idx = lo; idx < hi; idx++) {
      size_t pc = 0;
      const uint32_t ne = exprs[pc++];
      Fr acc = Fr::zero();
      for (uint32_t e = 0; e < ne; e++) {
        const uint32_t nt = exprs[pc++];
        Fr sum = Fr::zero();
        for (uint32_t t = 0; t < nt; t++) {
          const uint32_t ci = exprs[pc++], nf = exprs[pc++];
          Fr prod = one;
          if (ci != 0xffffffffu) memcpy(prod.v, consts + 4 * (size_t)ci, 32);
          for (uint32_t f = 0; f < nf; f++) {
            const uint32_t w = exprs[pc++];
            const size_t r = (size_t)(((int64_t)idx + (int16_t)(w >> 16)) & (int64_t)(n - 1));
            Fr v;
            memcpy(v.v, cols[w & 0xffffu] + 4 * r, 32);
            prod = mul(prod, v);
          }
          sum = add(sum, prod);
        }
        acc = add(mul(acc, theta), sum);
      }
      memcpy(out + 4 * idx, acc.v, 32);
    }
  });
}

// halo2_proofs plonk/lookup/prover.rs commit_product before blinding: denominators (a' + beta)(s' + gamma) batch-inverted,
// times (a + beta)(s + gamma), running product from 1; n values.
void oracle_lookup_product(const uint64_t* A, const uint64_t* S, const uint64_t* Ap, const uint64_t* Sp, size_t n,
                           const uint64_t* beta_, const uint64_t* gamma_, uint64_t* z) {
  Fr beta, gamma;
  memcpy(beta.v, beta_, 32);
  memcpy(gamma.v, gamma_, 32);
  std::vector<uint64_t> num(4 * n), den(4 * n);
  auto at = [](const uint64_t* p, size_t i) { Fr v; memcpy(v.v, p + 4 * i, 32); return v; };
  for (size_t i = 0; i < n; i++) {
    Fr d = mul(add(at(Ap, i), beta), add(at(Sp, i), gamma));
    Fr m = mul(add(at(A, i), beta), add(at(S, i), gamma));
    memcpy(den.data() + 4 * i, d.v, 32);
    memcpy(num.data() + 4 * i, m.v, 32);
  }
  oracle_grand_product(num.data(), den.data(), n, z);
}

// halo2_proofs plonk/permutation/prover.rs commit for ONE column set before blinding: modified_values = prod_j
// (v_j + beta sigma_j + gamma), batch-inverted, times prod_j (v_j + delta^j' omega^i beta + gamma) with deltaomega stepping
// by omega per row and by delta per column; z[0] = last_z.
void oracle_permutation_product(const uint64_t* const* values, const uint64_t* const* sigmas, size_t n_cols, uint32_t k,
                                const uint64_t* omega_, const uint64_t* beta_delta_start_, const uint64_t* delta_,
                                const uint64_t* beta_, const uint64_t* gamma_, const uint64_t* last_z, uint64_t* z) {
  const size_t n = (size_t)1 << k;
  Fr omega, bds, delta, beta, gamma;
  memcpy(omega.v, omega_, 32); memcpy(bds.v, beta_delta_start_, 32); memcpy(delta.v, delta_, 32);
  memcpy(beta.v, beta_, 32); memcpy(gamma.v, gamma_, 32);
  std::vector<Fr> num(n, Fr::one()), den(n, Fr::one());
  auto at = [](const uint64_t* p, size_t i) { Fr v; memcpy(v.v, p + 4 * i, 32); return v; };
  for (size_t j = 0; j < n_cols; j++)
    for (size_t i = 0; i < n; i++) den[i] = mul(den[i], add(add(mul(beta, at(sigmas[j], i)), gamma), at(values[j], i)));
  Fr col_start = bds;  // beta * delta^(first + j)
  for (size_t j = 0; j < n_cols; j++) {
    Fr deltaomega = col_start;
    for (size_t i = 0; i < n; i++) {
      num[i] = mul(num[i], add(add(deltaomega, gamma), at(values[j], i)));
      deltaomega = mul(deltaomega, omega);
    }
    col_start = mul(col_start, delta);
  }
  oracle_grand_product((const uint64_t*)num.data(), (const uint64_t*)den.data(), n, z);
  if (last_z) {
    Fr lz;
    memcpy(lz.v, last_z, 32);
    for (size_t i = 0; i < n; i++) {
      Fr v = mul(at(z, i), lz);
      memcpy(z + 4 * i, v.v, 32);
    }
  }
}

// halo2_proofs plonk/evaluation.rs Evaluator::evaluate_h (+ divide_by_vanishing_poly when t_evals != null), driven
// by the same word program as the device (layout: include/h2agg.h).  Follows halo2's CPU structure: the rows are
// split over threads (`parallelize`), each chunk starts beta_term = omega_ext^start and steps it per row.  Formulas
// and fold order are pinned by the reference's verifier (oracle/py/quotient_ref.py, tests/test_quotient_cpu.py).
void oracle_evaluate_h(const uint32_t* plan, size_t n_words, const uint64_t* const* cols, const uint64_t* consts,
                       uint32_t k, uint32_t ext_k, const uint64_t* y_, const uint64_t* beta_, const uint64_t* gamma_,
                       const uint64_t* theta_, const uint64_t* omega_ext_, const uint64_t* zeta_, const uint64_t* delta_,
                       const uint64_t* t_evals, size_t t_len, uint64_t* out, unsigned threads) {
  (void)n_words;
  const size_t size = (size_t)1 << ext_k;
  const int64_t rs = (int64_t)1 << (ext_k - k);
  Fr y, beta, gamma, theta, w_ext, zeta, delta;
  memcpy(y.v, y_, 32); memcpy(beta.v, beta_, 32); memcpy(gamma.v, gamma_, 32); memcpy(theta.v, theta_, 32);
  memcpy(w_ext.v, omega_ext_, 32); memcpy(zeta.v, zeta_, 32); memcpy(delta.v, delta_, 32);
  const uint32_t n_gates = plan[1], n_pcols = plan[2], chunk = plan[3], n_lookups = plan[5];
  const int32_t last_rot = (int32_t)plan[4];
  const Fr one = Fr::one();
  const Fr delta_start = mul(beta, zeta);
  parallelize(size, threads, [&](size_t lo, size_t hi) {
    u64 e[4] = {lo, 0, 0, 0};
    Fr beta_term = pow_words(w_ext, e);
    for (size_t idx = lo; idx < hi; idx++) {
      auto load = [&](uint32_t col, int64_t rot) {
        size_t r = (size_t)(((int64_t)idx + rot * rs) & (int64_t)(size - 1));
        Fr v;
        memcpy(v.v, cols[col] + 4 * r, 32);
        return v;
      };
      size_t pc = 9;
      auto poly = [&]() {
        uint32_t nt = plan[pc++];
        Fr acc = Fr::zero();
        for (uint32_t t = 0; t < nt; t++) {
          uint32_t ci = plan[pc++], nf = plan[pc++];
          Fr prod = one;
          if (ci != 0xffffffffu) memcpy(prod.v, consts + 4 * (size_t)ci, 32);
          for (uint32_t f = 0; f < nf; f++) {
            uint32_t w = plan[pc++];
            prod = mul(prod, load(w & 0xffffu, (int16_t)(w >> 16)));
          }
          acc = add(acc, prod);
        }
        return acc;
      };
      auto compress = [&]() {
        uint32_t ne = plan[pc++];
        Fr acc = Fr::zero();
        for (uint32_t i = 0; i < ne; i++) acc = add(mul(acc, theta), poly());
        return acc;
      };
      const Fr l0 = load(plan[6], 0), l_last = load(plan[7], 0), l_active = load(plan[8], 0);
      Fr value = Fr::zero();
      auto fold = [&](const Fr& term) { value = add(mul(value, y), term); };
      for (uint32_t g = 0; g < n_gates; g++) fold(poly());
      if (n_pcols) {
        uint32_t n_sets = (n_pcols + chunk - 1) / chunk;
        size_t pcols = pc, zcols = pc + 2 * (size_t)n_pcols;
        pc = zcols + n_sets;
        Fr z0 = load(plan[zcols], 0), zl = load(plan[zcols + n_sets - 1], 0);
        fold(mul(sub(one, z0), l0));
        fold(mul(sub(mul(zl, zl), zl), l_last));
        for (uint32_t s = 1; s < n_sets; s++)
          fold(mul(sub(load(plan[zcols + s], 0), load(plan[zcols + s - 1], last_rot)), l0));
        Fr current_delta = mul(delta_start, beta_term);
        for (uint32_t s = 0; s < n_sets; s++) {
          Fr left = load(plan[zcols + s], 1), right = load(plan[zcols + s], 0);
          uint32_t j1 = std::min((s + 1) * chunk, n_pcols);
          for (uint32_t j = s * chunk; j < j1; j++) {
            Fr v = load(plan[pcols + 2 * j], 0), sg = load(plan[pcols + 2 * j + 1], 0);
            left = mul(left, add(add(v, mul(beta, sg)), gamma));
            right = mul(right, add(add(v, current_delta), gamma));
            current_delta = mul(current_delta, delta);
          }
          fold(mul(sub(left, right), l_active));
        }
      }
      for (uint32_t l = 0; l < n_lookups; l++) {
        Fr cin = compress(), ctab = compress();
        Fr table_value = mul(add(cin, beta), add(ctab, gamma));
        uint32_t zc = plan[pc], ac = plan[pc + 1], sc = plan[pc + 2];
        pc += 3;
        Fr z = load(zc, 0), zn = load(zc, 1), a = load(ac, 0), ap = load(ac, -1), st = load(sc, 0);
        Fr ams = sub(a, st);
        fold(mul(sub(one, z), l0));
        fold(mul(sub(mul(z, z), z), l_last));
        fold(mul(sub(mul(mul(zn, add(a, beta)), add(st, gamma)), mul(z, table_value)), l_active));
        fold(mul(ams, l0));
        fold(mul(mul(ams, sub(a, ap)), l_active));
      }
      if (t_evals) {
        Fr t;
        memcpy(t.v, t_evals + 4 * (idx & (t_len - 1)), 32);
        value = mul(value, t);
      }
      memcpy(out + 4 * idx, value.v, 32);
      beta_term = mul(beta_term, w_ext);
    }
  });
}

unsigned oracle_hw_threads(void) {
  unsigned t = std::thread::hardware_concurrency();
  return t ? t : 1;
}

}  // extern "C"
