// BN254 G1 (y^2 = x^3 + 3) point arithmetic on the device.
//
// Affine layout = halo2curves `G1Affine {x, y}` (64 B, Montgomery Fq, identity = (0,0));
// Jacobian layout = halo2curves `G1 {x, y, z}` (96 B, identity z = 0)  -- SURVEY.md App. A.
// Buckets are kept in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, identity ZZ = 0):
// mixed add 8M+2S, full add 12M+2S, double 6M+4S(+small), no inversions until the very end; every Y3 = a*b - c*d
// is ONE dual-product Montgomery reduction (fp_mul_sub2).
// Formulas: EFD "shortw/xyzz" madd-2008-s, add-2008-s, dbl-2008-s-1, mdbl-2008-s-1 (a = 0).
#pragma once
#include "bn254_field.cuh"

namespace h2agg {

struct alignas(16) G1Affine {
  Fq x, y;
  __device__ __forceinline__ bool is_identity() const { return x.is_zero() && y.is_zero(); }
  __device__ __forceinline__ static G1Affine load_nc(const void* p) {
    G1Affine r;
    r.x = Fq::load_nc(p);
    r.y = Fq::load_nc(reinterpret_cast<const uint8_t*>(p) + 32);
    return r;
  }
};

struct alignas(16) G1Xyzz {
  Fq x, y, zz, zzz;
  __device__ __forceinline__ static G1Xyzz identity() {
    G1Xyzz r;
    r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
    return r;
  }
  __device__ __forceinline__ bool is_identity() const { return zz.is_zero(); }
  __device__ __forceinline__ static G1Xyzz load(const void* p) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
    G1Xyzz r;
    r.x = Fq::load(b); r.y = Fq::load(b + 32); r.zz = Fq::load(b + 64); r.zzz = Fq::load(b + 96);
    return r;
  }
  __device__ __forceinline__ void store(void* p) const {
    uint8_t* b = reinterpret_cast<uint8_t*>(p);
    x.store(b); y.store(b + 32); zz.store(b + 64); zzz.store(b + 96);
  }
};

__device__ __forceinline__ G1Xyzz xyzz_from_affine(const G1Affine& a) {
  G1Xyzz r;
  if (a.is_identity()) return G1Xyzz::identity();
  r.x = a.x; r.y = a.y; r.zz = Fq::one(); r.zzz = Fq::one();
  return r;
}

// 2*(affine)  (mdbl-2008-s-1)
__device__ __forceinline__ G1Xyzz xyzz_mdbl(const G1Affine& a) {
  if (a.is_identity()) return G1Xyzz::identity();
  G1Xyzz r;
  Fq u = fp_dbl(a.y);
  Fq v = fp_sqr(u);
  Fq w = u * v;
  Fq s = a.x * v;
  Fq xx = fp_sqr(a.x);
  Fq m = fp_dbl(xx) + xx;
  r.x = fp_sqr(m) - fp_dbl(s);
  r.y = fp_mul_sub2(m, s - r.x, w, a.y);
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2*p  (dbl-2008-s-1, a = 0)
__device__ __forceinline__ G1Xyzz xyzz_dbl(const G1Xyzz& p) {
  if (p.is_identity()) return p;
  G1Xyzz r;
  Fq u = fp_dbl(p.y);
  Fq v = fp_sqr(u);
  Fq w = u * v;
  Fq s = p.x * v;
  Fq xx = fp_sqr(p.x);
  Fq m = fp_dbl(xx) + xx;
  r.x = fp_sqr(m) - fp_dbl(s);
  r.y = fp_mul_sub2(m, s - r.x, w, p.y);
  r.zz = v * p.zz;
  r.zzz = w * p.zzz;
  return r;
}

// acc += (x2, y2) affine, y2 already sign-adjusted by the caller.  (madd-2008-s)
// Handles acc == identity, acc == q (doubling) and acc == -q (-> identity) exactly.
__device__ __forceinline__ void xyzz_madd(G1Xyzz& acc, const G1Affine& q) {
  if (q.is_identity()) return;
  if (acc.is_identity()) {
    acc.x = q.x; acc.y = q.y; acc.zz = Fq::one(); acc.zzz = Fq::one();
    return;
  }
  Fq u2 = q.x * acc.zz;
  Fq s2 = q.y * acc.zzz;
  Fq p = u2 - acc.x;
  Fq r = s2 - acc.y;
  if (p.is_zero()) {
    if (r.is_zero()) acc = xyzz_mdbl(q);
    else acc = G1Xyzz::identity();
    return;
  }
  Fq pp = fp_sqr(p);
  Fq ppp = p * pp;
  Fq qq = acc.x * pp;
  Fq x3 = fp_sqr(r) - ppp - fp_dbl(qq);
  acc.y = fp_mul_sub2(r, qq - x3, acc.y, ppp);  // one reduction for both products
  acc.x = x3;
  acc.zz = acc.zz * pp;
  acc.zzz = acc.zzz * ppp;
}

// acc += b  (add-2008-s), complete with the same special cases
__device__ __forceinline__ void xyzz_add(G1Xyzz& acc, const G1Xyzz& b) {
  if (b.is_identity()) return;
  if (acc.is_identity()) { acc = b; return; }
  Fq u1 = acc.x * b.zz;
  Fq u2 = b.x * acc.zz;
  Fq s1 = acc.y * b.zzz;
  Fq s2 = b.y * acc.zzz;
  Fq p = u2 - u1;
  Fq r = s2 - s1;
  if (p.is_zero()) {
    if (r.is_zero()) acc = xyzz_dbl(acc);
    else acc = G1Xyzz::identity();
    return;
  }
  Fq pp = fp_sqr(p);
  Fq ppp = p * pp;
  Fq qq = u1 * pp;
  Fq x3 = fp_sqr(r) - ppp - fp_dbl(qq);
  acc.y = fp_mul_sub2(r, qq - x3, s1, ppp);
  acc.x = x3;
  acc.zz = acc.zz * b.zz * pp;
  acc.zzz = acc.zzz * b.zzz * ppp;
}

// XYZZ -> affine (one inversion); identity -> (0,0)
static __device__ __noinline__ G1Affine xyzz_to_affine(const G1Xyzz& p) {
  G1Affine a;
  if (p.is_identity()) { a.x = Fq::zero(); a.y = Fq::zero(); return a; }
  // 1/ZZZ, then 1/ZZ = ZZZ^-2 * ZZ^2  (ZZ^3 = ZZZ^2)
  Fq izzz = fp_inv(p.zzz);
  Fq izz = fp_sqr(izzz) * fp_sqr(p.zz);
  a.x = p.x * izz;
  a.y = p.y * izzz;
  return a;
}

}  // namespace h2agg
