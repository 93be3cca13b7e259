// BN254 Fr / Fq arithmetic for sm_100a: 8 x u32 limbs, Montgomery form (R = 2^256).
//
// Memory layout is exactly halo2curves 0.2.1's `Fr`/`Fq` = [u64;4] little-endian limbs in
// Montgomery form (reference type used at
// halo2-snark-aggregator-circuit/src/verify_circuit.rs:53-55), so buffers cross the C ABI
// without conversion: 4 x u64 LE == 8 x u32 LE.
//
// Multiplication is the generated even/odd-column PTX schedule (gen/mont_mul_bn254.inc,
// produced and self-checked by tools/gen_mont_ptx.py): 128 IMAD.WIDE + 8 IMAD per product.
// Define H2AGG_PORTABLE_MUL to get a plain-C CIOS instead (used to A/B the PTX on device).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "gen/mont_mul_bn254.inc"

namespace h2agg {

struct FrTag {
  // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
  __host__ __device__ static constexpr uint32_t P(int i) {
    constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return v[i];
  }
  // R mod r  (Montgomery one)
  __host__ __device__ static constexpr uint32_t ONE(int i) {
    constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return v[i];
  }
  // R^2 mod r
  __host__ __device__ static constexpr uint32_t R2(int i) {
    constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return v[i];
  }
  static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
  static constexpr bool IS_FR = true;
};

struct FqTag {
  // p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
  __host__ __device__ static constexpr uint32_t P(int i) {
    constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return v[i];
  }
  __host__ __device__ static constexpr uint32_t ONE(int i) {
    constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return v[i];
  }
  __host__ __device__ static constexpr uint32_t R2(int i) {
    constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return v[i];
  }
  static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
  static constexpr bool IS_FR = false;
};

template <class T>
struct alignas(16) Fp {
  uint32_t v[8];

  __device__ __forceinline__ static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  __device__ __forceinline__ static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = T::ONE(i);
    return r;
  }
  __device__ __forceinline__ static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = T::R2(i);
    return r;
  }
  __device__ __forceinline__ bool is_zero() const {
    uint32_t o = v[0];
#pragma unroll
    for (int i = 1; i < 8; i++) o |= v[i];
    return o == 0;
  }
  __device__ __forceinline__ bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }

  // global / shared memory access as two 16-byte vectors (pointer must be 16-byte aligned)
  __device__ __forceinline__ static Fp load(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ static Fp load_nc(const void* p) {  // read-only path
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ void store(void* p) const {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v[0], v[1], v[2], v[3]);
    q[1] = make_uint4(v[4], v[5], v[6], v[7]);
  }
  __device__ __forceinline__ uint4 lo4() const { return make_uint4(v[0], v[1], v[2], v[3]); }
  __device__ __forceinline__ uint4 hi4() const { return make_uint4(v[4], v[5], v[6], v[7]); }
  __device__ __forceinline__ static Fp from_halves(uint4 a, uint4 b) {
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
};

// r = a - p if a >= p else a          (a < 2p on entry)
template <class T>
__device__ __forceinline__ void fp_reduce_once(uint32_t* a) {
  uint32_t t[8], borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
        "=r"(borrow)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "n"(T::P(0)), "n"(T::P(1)), "n"(T::P(2)), "n"(T::P(3)), "n"(T::P(4)), "n"(T::P(5)), "n"(T::P(6)),
        "n"(T::P(7)));
  // borrow == 0xffffffff when a < p
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = borrow ? a[i] : t[i];
}

template <class T>
__device__ __forceinline__ Fp<T> operator+(const Fp<T>& a, const Fp<T>& b) {
  Fp<T> r;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  fp_reduce_once<T>(r.v);  // both < p < 2^254, so the sum fits 256 bits and is < 2p
  return r;
}

template <class T>
__device__ __forceinline__ Fp<T> operator-(const Fp<T>& a, const Fp<T>& b) {
  uint32_t t[8], borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
        "=r"(borrow)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  Fp<T> r;
  // add back p under the borrow mask
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7])
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "r"(borrow & T::P(0)), "r"(borrow & T::P(1)), "r"(borrow & T::P(2)), "r"(borrow & T::P(3)),
        "r"(borrow & T::P(4)), "r"(borrow & T::P(5)), "r"(borrow & T::P(6)), "r"(borrow & T::P(7)));
  return r;
}

template <class T>
__device__ __forceinline__ Fp<T> fp_neg(const Fp<T>& a) {
  return Fp<T>::zero() - a;  // 0 - 0 = 0 stays canonical
}

template <class T>
__device__ __forceinline__ Fp<T> fp_dbl(const Fp<T>& a) {
  return a + a;
}

#ifdef H2AGG_PORTABLE_MUL
template <class T>
__device__ __forceinline__ Fp<T> operator*(const Fp<T>& a, const Fp<T>& b) {
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint64_t s = (uint64_t)a.v[j] * b.v[i] + t[j] + c;
      t[j] = (uint32_t)s;
      c = s >> 32;
    }
    uint64_t s = (uint64_t)t[8] + c;
    t[8] = (uint32_t)s;
    t[9] = (uint32_t)(s >> 32);
    uint32_t m = t[0] * T::INV;
    c = ((uint64_t)m * T::P(0) + t[0]) >> 32;
#pragma unroll
    for (int j = 1; j < 8; j++) {
      uint64_t s2 = (uint64_t)m * T::P(j) + t[j] + c;
      t[j - 1] = (uint32_t)s2;
      c = s2 >> 32;
    }
    s = (uint64_t)t[8] + c;
    t[7] = (uint32_t)s;
    t[8] = t[9] + (uint32_t)(s >> 32);
  }
  Fp<T> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
  fp_reduce_once<T>(r.v);
  return r;
}
#else
template <class T>
__device__ __forceinline__ Fp<T> operator*(const Fp<T>& a, const Fp<T>& b) {
  Fp<T> r;
#ifdef H2AGG_KARATSUBA
  if constexpr (T::IS_FR) {
    H2AGG_MONT_MUL_FR_K(r.v, a.v, b.v);
  } else {
    H2AGG_MONT_MUL_FQ_K(r.v, a.v, b.v);
  }
#else
  if constexpr (T::IS_FR) {
    H2AGG_MONT_MUL_FR(r.v, a.v, b.v);
  } else {
    H2AGG_MONT_MUL_FQ(r.v, a.v, b.v);
  }
#endif
  fp_reduce_once<T>(r.v);
  return r;
}
#endif

// a*b + c*d with ONE Montgomery reduction (generated dual-product schedule: 200 wide MADs + 8 instead of 2 x 136).
// Inputs < p < 2^254 keep every row below 2^288, the result below 1.5 p.
template <class T>
__device__ __forceinline__ Fp<T> fp_mul_add2(const Fp<T>& a, const Fp<T>& b, const Fp<T>& c, const Fp<T>& d) {
#ifdef H2AGG_PORTABLE_MUL
  return a * b + c * d;
#else
  Fp<T> r;
#ifdef H2AGG_KARATSUBA
  if constexpr (T::IS_FR) {
    H2AGG_MONT_MUL_ADD2_FR_K(r.v, a.v, b.v, c.v, d.v);
  } else {
    H2AGG_MONT_MUL_ADD2_FQ_K(r.v, a.v, b.v, c.v, d.v);
  }
#else
  if constexpr (T::IS_FR) {
    H2AGG_MONT_MUL_ADD2_FR(r.v, a.v, b.v, c.v, d.v);
  } else {
    H2AGG_MONT_MUL_ADD2_FQ(r.v, a.v, b.v, c.v, d.v);
  }
#endif
  fp_reduce_once<T>(r.v);
  return r;
#endif
}
// a*b - c*d
template <class T>
__device__ __forceinline__ Fp<T> fp_mul_sub2(const Fp<T>& a, const Fp<T>& b, const Fp<T>& c, const Fp<T>& d) {
  return fp_mul_add2(a, b, fp_neg(c), d);
}

// a^2 with the generated squaring schedule: 36 + 72 multiplier instructions instead of 136 (a must be < 2^254,
// which every fully reduced element is)
template <class T>
__device__ __forceinline__ Fp<T> fp_sqr(const Fp<T>& a) {
#ifdef H2AGG_PORTABLE_MUL
  return a * a;
#else
  Fp<T> r;
  if constexpr (T::IS_FR) {
    H2AGG_MONT_SQR_FR(r.v, a.v);
  } else {
    H2AGG_MONT_SQR_FQ(r.v, a.v);
  }
  fp_reduce_once<T>(r.v);
  return r;
#endif
}

// Montgomery -> canonical integer (still 8 x u32 LE)
template <class T>
__device__ __forceinline__ Fp<T> fp_from_mont(const Fp<T>& a) {
  Fp<T> o = Fp<T>::zero();
  o.v[0] = 1;
  return a * o;
}
template <class T>
__device__ __forceinline__ Fp<T> fp_to_mont(const Fp<T>& a) {
  return a * Fp<T>::r2();
}

// a^e for a 256-bit exponent given as 8 LE words (variable time; e is public)
template <class T>
__device__ __noinline__ Fp<T> fp_pow(const Fp<T>& a, const uint32_t* e) {
  Fp<T> r = Fp<T>::one();
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) r = fp_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) {
      r = started ? r * a : a;
      started = true;
    }
  }
  return r;
}

// a^(p-2); inverse(0) = 0
template <class T>
__device__ __noinline__ Fp<T> fp_inv(const Fp<T>& a) {
  uint32_t e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = T::P(i);
  e[0] -= 2;  // low limb of both moduli is >= 2
  return fp_pow(a, e);
}

// a^e for a small integer exponent
template <class T>
__device__ __forceinline__ Fp<T> fp_pow_u64(const Fp<T>& a, uint64_t e) {
  Fp<T> r = Fp<T>::one();
  Fp<T> base = a;
  while (e) {
    if (e & 1) r = r * base;
    e >>= 1;
    if (e) base = fp_sqr(base);
  }
  return r;
}

using Fr = Fp<FrTag>;
using Fq = Fp<FqTag>;

}  // namespace h2agg
