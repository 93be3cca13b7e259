// Host-side modular inverse of a 256-bit value modulo an odd 254-bit prime, by batches of 62 division steps
// (Bernstein-Yang "safegcd", variable time): the pair (f, g) = (M, x) is driven to (+-1, 0) while the 2x2 transition
// matrix of each batch, computed on the low 62 bits only, is applied to (f, g) and -- modulo M, with an exact division by
// 2^62 -- to the cofactors (d, e).  About ten batches for 254-bit inputs instead of ~380 full-width shift/subtract
// rounds of a binary extended Euclid: `div` in the recorded chip chain needs one native inverse per curve operation
// (witness_recorder.cu: int_div), and that inverse was 38 % of the host recording time.
#pragma once
#include <cstdint>

namespace h2agg {
namespace modinv {

typedef __int128 i128;
typedef unsigned __int128 u128;

struct S62 {  // signed, value = sum v[i] 2^(62 i); limbs 0..3 in [0, 2^62), limb 4 carries the sign
  int64_t v[5];
};
struct Modulus {
  S62 m;
  uint64_t m_inv62;  // M^-1 mod 2^62
};
struct T2x2 {
  int64_t u, v, q, r;
};

static const uint64_t M62 = UINT64_MAX >> 2;

static inline S62 to_s62(const uint64_t a[4]) {
  S62 r;
  r.v[0] = (int64_t)(a[0] & M62);
  r.v[1] = (int64_t)(((a[0] >> 62) | (a[1] << 2)) & M62);
  r.v[2] = (int64_t)(((a[1] >> 60) | (a[2] << 4)) & M62);
  r.v[3] = (int64_t)(((a[2] >> 58) | (a[3] << 6)) & M62);
  r.v[4] = (int64_t)(a[3] >> 56);
  return r;
}
static inline void from_s62(const S62& a, uint64_t out[4]) {  // a in [0, 2^256)
  const uint64_t a0 = (uint64_t)a.v[0], a1 = (uint64_t)a.v[1], a2 = (uint64_t)a.v[2], a3 = (uint64_t)a.v[3], a4 = (uint64_t)a.v[4];
  out[0] = a0 | (a1 << 62);
  out[1] = (a1 >> 2) | (a2 << 60);
  out[2] = (a2 >> 4) | (a3 << 58);
  out[3] = (a3 >> 6) | (a4 << 56);
}
static inline Modulus make_modulus(const uint64_t m[4]) {
  Modulus r;
  r.m = to_s62(m);
  uint64_t x = m[0];  // Newton: x <- x (2 - m x) doubles the correct low bits; m odd => m*m = 1 mod 8
  for (int i = 0; i < 6; i++) x *= 2 - m[0] * x;
  r.m_inv62 = x & M62;
  return r;
}

// 62 division steps on the low words; returns the new eta, fills the transition matrix t with
// t * [f, g] = 2^62 * [f', g'].
static inline int64_t divsteps_62_var(int64_t eta, uint64_t f0, uint64_t g0, T2x2* t) {
  uint64_t u = 1, v = 0, q = 0, r = 1;
  uint64_t f = f0, g = g0;
  int i = 62;
  for (;;) {
    const int zeros = __builtin_ctzll(g | (UINT64_MAX << i));
    g >>= zeros;
    u <<= zeros;
    v <<= zeros;
    eta -= zeros;
    i -= zeros;
    if (i == 0) break;
    uint64_t w, m;
    int limit;
    if (eta < 0) {  // swap: (f, g) <- (g, -f)
      uint64_t tmp;
      eta = -eta;
      tmp = f; f = g; g = (uint64_t)0 - tmp;
      tmp = u; u = q; q = (uint64_t)0 - tmp;
      tmp = v; v = r; r = (uint64_t)0 - tmp;
      // cancel up to 6 low bits of g at once (never more than the steps left, nor than eta + 1: eta flips sign
      // again after that many): -1/f mod 2^6 = f (f^2 - 2), one Newton step from f^-1 = f mod 8
      limit = ((int)eta + 1) > i ? i : ((int)eta + 1);
      m = (UINT64_MAX >> (64 - limit)) & 63U;
      w = (f * g * (f * f - 2)) & m;
    } else {  // eta is small here: up to 4 bits, 1/f mod 2^4 = f + 8 [f = 3 mod 4 ... bit 2 of f + 1]
      limit = ((int)eta + 1) > i ? i : ((int)eta + 1);
      m = (UINT64_MAX >> (64 - limit)) & 15U;
      w = f + (((f + 1) & 4) << 1);
      w = ((uint64_t)0 - w) * g & m;
    }
    g += f * w;
    q += u * w;
    r += v * w;
  }
  t->u = (int64_t)u; t->v = (int64_t)v; t->q = (int64_t)q; t->r = (int64_t)r;
  return eta;
}

static inline void update_fg(int len, S62* f, S62* g, const T2x2* t) {
  const int64_t u = t->u, v = t->v, q = t->q, r = t->r;
  int64_t fi = f->v[0], gi = g->v[0];
  i128 cf = (i128)u * fi + (i128)v * gi;
  i128 cg = (i128)q * fi + (i128)r * gi;
  cf >>= 62;
  cg >>= 62;
  for (int i = 1; i < len; i++) {
    fi = f->v[i];
    gi = g->v[i];
    cf += (i128)u * fi + (i128)v * gi;
    cg += (i128)q * fi + (i128)r * gi;
    f->v[i - 1] = (int64_t)((uint64_t)cf & M62);
    cf >>= 62;
    g->v[i - 1] = (int64_t)((uint64_t)cg & M62);
    cg >>= 62;
  }
  f->v[len - 1] = (int64_t)cf;
  g->v[len - 1] = (int64_t)cg;
}

// (d, e) <- t * (d, e) / 2^62 mod M, keeping both in (-2M, M)
static inline void update_de(S62* d, S62* e, const T2x2* t, const Modulus* mod) {
  const int64_t u = t->u, v = t->v, q = t->q, r = t->r;
  const int64_t sd = d->v[4] >> 63, se = e->v[4] >> 63;
  int64_t md = (u & sd) + (v & se);
  int64_t me = (q & sd) + (r & se);
  i128 cd = (i128)u * d->v[0] + (i128)v * e->v[0];
  i128 ce = (i128)q * d->v[0] + (i128)r * e->v[0];
  md -= (int64_t)((mod->m_inv62 * (uint64_t)cd + (uint64_t)md) & M62);
  me -= (int64_t)((mod->m_inv62 * (uint64_t)ce + (uint64_t)me) & M62);
  cd += (i128)mod->m.v[0] * md;
  ce += (i128)mod->m.v[0] * me;
  cd >>= 62;
  ce >>= 62;
  for (int i = 1; i < 5; i++) {
    cd += (i128)u * d->v[i] + (i128)v * e->v[i] + (i128)mod->m.v[i] * md;
    ce += (i128)q * d->v[i] + (i128)r * e->v[i] + (i128)mod->m.v[i] * me;
    d->v[i - 1] = (int64_t)((uint64_t)cd & M62);
    cd >>= 62;
    e->v[i - 1] = (int64_t)((uint64_t)ce & M62);
    ce >>= 62;
  }
  d->v[4] = (int64_t)cd;
  e->v[4] = (int64_t)ce;
}

// out = x^-1 mod M for 0 < x < M (plain integers, little-endian u64 x 4); x = 0 gives 0.
static inline void inverse(const uint64_t x[4], const Modulus& mod, uint64_t out[4]) {
  S62 d = {{0, 0, 0, 0, 0}}, e = {{1, 0, 0, 0, 0}};
  S62 f = mod.m, g = to_s62(x);
  int len = 5;
  int64_t eta = -1;
  if ((x[0] | x[1] | x[2] | x[3]) == 0) {
    out[0] = out[1] = out[2] = out[3] = 0;
    return;
  }
  for (;;) {
    T2x2 t;
    eta = divsteps_62_var(eta, (uint64_t)f.v[0], (uint64_t)g.v[0], &t);
    update_de(&d, &e, &t, &mod);
    update_fg(len, &f, &g, &t);
    if (g.v[0] == 0) {
      int64_t c = 0;
      for (int j = 1; j < len; j++) c |= g.v[j];
      if (c == 0) break;
    }
    const int64_t fn = f.v[len - 1], gn = g.v[len - 1];
    int64_t cond = ((int64_t)len - 2) >> 63;
    cond |= fn ^ (fn >> 63);
    cond |= gn ^ (gn >> 63);
    if (cond == 0) {  // top limbs are pure sign: fold them into the limb below
      f.v[len - 2] |= (int64_t)((uint64_t)fn << 62);
      g.v[len - 2] |= (int64_t)((uint64_t)gn << 62);
      --len;
    }
  }
  // f = +-1 and d * x = f (mod M): result = sign(f) * d, brought into [0, M)
  const bool neg = f.v[len - 1] < 0;
  i128 c = 0;
  int64_t w[5];
  for (int i = 0; i < 5; i++) {  // w = +-d, limbs renormalised
    c += neg ? -(i128)d.v[i] : (i128)d.v[i];
    w[i] = (i < 4) ? (int64_t)((uint64_t)c & M62) : (int64_t)c;
    if (i < 4) c >>= 62;
  }
  for (int pass = 0; pass < 3; pass++) {
    if (w[4] < 0) {  // += M
      i128 cc = 0;
      for (int i = 0; i < 5; i++) {
        cc += (i128)w[i] + mod.m.v[i];
        w[i] = (i < 4) ? (int64_t)((uint64_t)cc & M62) : (int64_t)cc;
        if (i < 4) cc >>= 62;
      }
    } else {
      bool ge = true;  // w >= M ?
      for (int i = 4; i >= 0; i--) {
        if (w[i] > mod.m.v[i]) break;
        if (w[i] < mod.m.v[i]) { ge = false; break; }
      }
      if (!ge) break;
      i128 cc = 0;
      for (int i = 0; i < 5; i++) {
        cc += (i128)w[i] - mod.m.v[i];
        w[i] = (i < 4) ? (int64_t)((uint64_t)cc & M62) : (int64_t)cc;
        if (i < 4) cc >>= 62;
      }
    }
  }
  S62 res = {{w[0], w[1], w[2], w[3], w[4]}};
  from_s62(res, out);
}

}  // namespace modinv
}  // namespace h2agg
