// W1/W2 (and the rows of W3-W5 that ride on them): batched wrong-field integer witness expansion.
//
// Replaces the value computation of halo2-ecc-circuit-lib's FiveColumnIntegerChip
// (/root/reference/halo2-ecc-circuit-lib/src/five/integer_chip.rs) as driven by EccChipOps
// (chips/ecc_chip.rs) inside the aggregation circuit's synthesize
// (halo2-snark-aggregator-circuit/src/verify_circuit.rs:242-504).  During proving halo2 keeps only
// the advice cells, so the output is exactly the 5 advice columns (Montgomery Fr, the layout
// halo2 holds in memory); fixed cells and copy constraints belong to keygen.
//
// One thread per op record (witness_ops.h).  The records arrive in recording order; a first kernel groups
// their indices by opcode (the host knows the per-opcode counts, so this is one scatter with warp-aggregated
// cursors), and every recipe then runs as its own kernel over its index range: own register budget, no
// divergence.  The three inversions of `is_zero` are not
// done per thread (three 254-step Fermat ladders each): a first kernel writes the values to invert, the
// batched inversion of scan.cu inverts all of them at once, a second kernel writes the rows.
// The heavy recipe is MULEQ: 560-bit product, exact division by p through
// p^-1 mod 2^288, 68-bit limbs, 17-bit range chunks, the limb-product chain, the carry words
// v0/v1 (computed in Fr, as the reference does), natives -- 29-31 rows x 5 cells per record.
#include "bn254_field.cuh"
#include "ctx.hpp"
#include "witness_ops.h"

namespace h2agg {

struct Cols {
  Fr* c[5];
  uint32_t n_rows;
};

__device__ __forceinline__ Fr fr_from_words(const uint32_t* w, int n) {  // canonical integer (< r) -> Montgomery
  Fr a = Fr::zero();
  for (int i = 0; i < n && i < 8; i++) a.v[i] = w[i];
  return fp_to_mont(a);
}
__device__ __forceinline__ Fr fr_from_u128(uint64_t lo, uint64_t hi) {
  Fr a = Fr::zero();
  a.v[0] = (uint32_t)lo; a.v[1] = (uint32_t)(lo >> 32); a.v[2] = (uint32_t)hi; a.v[3] = (uint32_t)(hi >> 32);
  return fp_to_mont(a);
}
__device__ __forceinline__ Fr fr_from_u64(uint64_t x) { return fr_from_u128(x, 0); }

__device__ __forceinline__ void put(const Cols& o, uint32_t row, int col, const Fr& v) {
  if (row < o.n_rows) v.store(o.c[col] + row);
}
__device__ __forceinline__ void put_row(const Cols& o, uint32_t row, const Fr& a0, const Fr& a1, const Fr& a2,
                                        const Fr& a3, const Fr& a4) {
  put(o, row, 0, a0); put(o, row, 1, a1); put(o, row, 2, a2); put(o, row, 3, a3); put(o, row, 4, a4);
}

// constants (see oracle/py/ecc_chip_ref.py::Helper; recomputed in tests/test_witness_cpu.py)
__device__ const uint32_t W_PINV288[9] = {0x1b799c77u, 0x782df87du, 0xe1359536u, 0x6121829au, 0xe7cc257fu,
                                          0x2750342fu, 0x6e777394u, 0x0a85dd48u, 0x5b52d390u};  // p^-1 mod 2^288
__device__ const uint32_t W_NEGW[4][3] = {{0x278302b9u, 0xc3df73e9u, 0x00000002u},   // limbs of 2^272 - p
                                          {0xe978e357u, 0x2687e956u, 0x0000000au},
                                          {0x497e7ea7u, 0xd647afbau, 0x0000000fu},
                                          {0x18d1ece5u, 0xfffcf9bbu, 0x0000000fu}};
__device__ const uint32_t W_P_LIMB0[3] = {0xd87cfd47u, 0x3c208c16u, 0x0000000du};          // p mod 2^68
__device__ const uint32_t W_NATIVE[8] = {0xe87cfd46u, 0xf83e9682u, 0xeeb859fbu, 0x6f4d8248u, 0, 0, 0, 0};  // p mod r
__device__ const uint32_t W_INV_2_136[8] = {0x0766f9ddu, 0x568bea8eu, 0x219532a9u, 0xa31a140fu,
                                            0xcea9b991u, 0x1a908db2u, 0xe8acfaedu, 0x1b7c016fu};  // 2^-136 mod r

__device__ __forceinline__ Fr limb_exp(int i) {  // 2^(68 i) mod r, Montgomery
  Fr a = Fr::zero();
  a.v[(68 * i) >> 5] = 1u << ((68 * i) & 31);
  return fp_to_mont(a);
}

// 68-bit limb (lo 64 bits, hi 4 bits) helpers -------------------------------------------------
struct L68 { uint64_t lo; uint32_t hi; };

__device__ __forceinline__ uint32_t chunk17(const L68& l, int i) {  // bits [17 i, 17 i + 17)
  int sh = 17 * i;
  uint64_t v = (sh < 64) ? (l.lo >> sh) : 0;
  if (sh + 17 > 64) v |= (uint64_t)l.hi << (64 - sh);
  return (uint32_t)v & 0x1ffffu;
}
__device__ __forceinline__ Fr l68_to_fr(const L68& l) { return fr_from_u128(l.lo, l.hi); }

// [c3,c2,c1,c0,n] (non-leading) or [top chunks..., 0..., n] for a leading limb with `nchunks` chunks
__device__ __forceinline__ void put_limb_row(const Cols& o, uint32_t row, const L68& l, int nchunks) {
  Fr cells[5];
#pragma unroll
  for (int k = 0; k < 4; k++) cells[k] = (k < nchunks) ? fr_from_u64(chunk17(l, nchunks - 1 - k)) : Fr::zero();
  cells[4] = l68_to_fr(l);
  put_row(o, row, cells[0], cells[1], cells[2], cells[3], cells[4]);
}

// split a little-endian word array into four 68-bit limbs (top limb takes what is left, < 2^68)
__device__ __forceinline__ void split_limbs(const uint32_t* w, int nwords, L68* out) {
  for (int i = 0; i < 4; i++) {
    int bit = 68 * i, wi = bit >> 5, sh = bit & 31;
    uint64_t parts[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; k++) parts[k] = (wi + k < nwords) ? w[wi + k] : 0;
    // 96+ bits starting at word wi, shifted right by sh
    unsigned __int128 v = (unsigned __int128)parts[0] | ((unsigned __int128)parts[1] << 32) |
                          ((unsigned __int128)parts[2] << 64) | ((unsigned __int128)parts[3] << 96);
    v >>= sh;
    out[i].lo = (uint64_t)v;
    out[i].hi = (uint32_t)(v >> 64) & (i < 3 ? 0xfu : 0xffffffffu);
  }
}

// bn = sum limb_i * 2^(68 i) for 128-bit limbs -> 10 words
__device__ __forceinline__ void limbs_to_bn(const uint64_t* l /*8 u64*/, uint32_t* out /*10*/) {
  for (int i = 0; i < 10; i++) out[i] = 0;
  for (int i = 0; i < 4; i++) {
    uint32_t w[4] = {(uint32_t)l[2 * i], (uint32_t)(l[2 * i] >> 32), (uint32_t)l[2 * i + 1], (uint32_t)(l[2 * i + 1] >> 32)};
    int bit = 68 * i, wi = bit >> 5, sh = bit & 31;
    uint64_t carry = 0;
    for (int k = 0; k < 5 && wi + k < 10; k++) {
      uint64_t piece = 0;
      if (k < 4) piece |= ((uint64_t)w[k] << sh) & 0xffffffffull;
      if (k > 0 && sh) piece |= (uint64_t)w[k - 1] >> (32 - sh);
      uint64_t s = (uint64_t)out[wi + k] + piece + carry;
      out[wi + k] = (uint32_t)s;
      carry = s >> 32;
    }
    for (int k = wi + 5; carry && k < 10; k++) {
      uint64_t s = (uint64_t)out[k] + carry;
      out[k] = (uint32_t)s;
      carry = s >> 32;
    }
  }
}

__device__ __forceinline__ Fr limb128_to_fr(const uint64_t* l) { return fr_from_u128(l[0], l[1]); }

__device__ __forceinline__ Fr native_of(const Fr* l /*4 limbs as Fr*/) {
  return l[0] + l[1] * limb_exp(1) + l[2] * limb_exp(2) + l[3] * limb_exp(3);
}

// assign_w rows of a canonical integer given as four 68-bit limbs (MS limb first)   :447-464
__device__ __forceinline__ void put_assign_w(const Cols& o, uint32_t row, const L68* f) {
  put_limb_row(o, row + 0, f[3], 3);  // w_ceil leading limb: 50 bits = 3 chunks
  put_limb_row(o, row + 1, f[2], 4);
  put_limb_row(o, row + 2, f[1], 4);
  put_limb_row(o, row + 3, f[0], 4);
}

__device__ __forceinline__ void l68_from_u128(const uint64_t* l, L68& out) { out.lo = l[0]; out.hi = (uint32_t)l[1]; }

// (is_zero condition, inverse) rows of BaseGateOps::invert, b = a^-1 (0 for a = 0)       gates/base_gate.rs:439-476
__device__ __forceinline__ Fr put_invert(const Cols& o, uint32_t row, const Fr& a, const Fr& b) {
  Fr c = Fr::one() - a * b;
  Fr z = Fr::zero();
  put_row(o, row, a, c, z, z, z);
  put_row(o, row + 1, a, b, c, z, z);
  return c;
}

__device__ void expand_muleq(const WitnessOp& op, const Cols& o) {
  const bool cx = op.flags & 1, cy = op.flags & 2, cz = op.flags & 4, sq = op.flags & 8, fresh_y = op.flags & 16;
  const uint64_t* X = op.v;
  const uint64_t* Y = sq ? op.v : op.v + 8;
  const uint64_t* Z = op.v + 16;
  uint32_t row = op.row;
  const Fr zero = Fr::zero();

  // d = (bn(x) * bn(y) - bn(z)) / p, exactly, via p^-1 mod 2^288 (d < 2^272)
  uint32_t bx[10], by[10], bz[10], t[9];
  limbs_to_bn(X, bx);
  limbs_to_bn(Y, by);
  limbs_to_bn(Z, bz);
  {
    uint32_t prod[9];
    for (int i = 0; i < 9; i++) prod[i] = 0;
    for (int i = 0; i < 9; i++) {
      uint64_t carry = 0;
      for (int j = 0; i + j < 9; j++) {
        uint64_t s = (uint64_t)bx[i] * by[j] + prod[i + j] + carry;
        prod[i + j] = (uint32_t)s;
        carry = s >> 32;
      }
    }
    uint64_t borrow = 0;
    for (int i = 0; i < 9; i++) {
      uint64_t s = (uint64_t)prod[i] - bz[i] - borrow;
      t[i] = (uint32_t)s;
      borrow = (s >> 32) & 1;
    }
  }
  uint32_t dw[9];
  for (int i = 0; i < 9; i++) dw[i] = 0;
  for (int i = 0; i < 9; i++) {
    uint64_t carry = 0;
    for (int j = 0; i + j < 9; j++) {
      uint64_t s = (uint64_t)t[i] * W_PINV288[j] + dw[i + j] + carry;
      dw[i + j] = (uint32_t)s;
      carry = s >> 32;
    }
  }
  L68 d[4];
  split_limbs(dw, 9, d);

  // rows 0-3: assign_w of the fresh integer (rem for mul/square, c for div)
  {
    L68 f[4];
    const uint64_t* F = fresh_y ? Y : Z;
    for (int i = 0; i < 4; i++) l68_from_u128(F + 2 * i, f[i]);
    put_assign_w(o, row, f);
    row += 4;
  }
  // rows 4-7: assign_d (leading limb 67 bits = 4 chunks)                                 :428-445
  put_limb_row(o, row + 0, d[3], 4);
  put_limb_row(o, row + 1, d[2], 4);
  put_limb_row(o, row + 2, d[1], 4);
  put_limb_row(o, row + 3, d[0], 4);
  row += 4;

  Fr xf[4], yf[4], zf[4], df[4], negw[4];
  for (int i = 0; i < 4; i++) {
    xf[i] = limb128_to_fr(X + 2 * i);
    yf[i] = limb128_to_fr(Y + 2 * i);
    zf[i] = limb128_to_fr(Z + 2 * i);
    df[i] = l68_to_fr(d[i]);
    Fr n = Fr::zero();
    n.v[0] = W_NEGW[i][0]; n.v[1] = W_NEGW[i][1]; n.v[2] = W_NEGW[i][2];
    negw[i] = fp_to_mont(n);
  }
  // rows 8-17: limb products, mul_add then mul_add2 chain                                :155-172, five/base_gate.rs:110-128
  Fr l[4];
  for (int pos = 0; pos < 4; pos++) {
    Fr acc = xf[0] * yf[pos] + df[0] * negw[pos];
    put_row(o, row++, xf[0], yf[pos], df[0], acc, zero);
    for (int i = 1; i <= pos; i++) {
      Fr nxt = xf[i] * yf[pos - i] + df[i] * negw[pos - i] + acc;
      put_row(o, row++, xf[i], yf[pos - i], df[i], acc, nxt);
      acc = nxt;
    }
    l[pos] = acc;
  }
  // carries, computed in Fr like the reference                                           :191-206
  const Fr e1 = limb_exp(1), e2 = limb_exp(2);
  Fr inv136;
  for (int i = 0; i < 8; i++) inv136.v[i] = W_INV_2_136[i];
  inv136 = fp_to_mont(inv136);
  Fr u0 = (l[1] - zf[1]) * e1 + l[0] - zf[0] + e2;
  Fr v0 = u0 * inv136;
  Fr v0c = fp_from_mont(v0);
  L68 v0l, v0h;
  v0l.lo = (uint64_t)v0c.v[0] | ((uint64_t)v0c.v[1] << 32);
  v0l.hi = v0c.v[2] & 0xfu;
  {
    unsigned __int128 hi = ((unsigned __int128)v0c.v[2] | ((unsigned __int128)v0c.v[3] << 32) |
                            ((unsigned __int128)v0c.v[4] << 64) | ((unsigned __int128)v0c.v[5] << 96)) >> 4;
    v0h.lo = (uint64_t)hi;
    v0h.hi = (uint32_t)(hi >> 64);
  }
  Fr u1 = v0 - Fr::one() + l[2] - zf[2] + (l[3] - zf[3]) * e1;
  Fr v1 = u1 * inv136;
  Fr v1c = fp_from_mont(v1);
  L68 v1l, v1h;
  v1l.lo = (uint64_t)v1c.v[0] | ((uint64_t)v1c.v[1] << 32);
  v1l.hi = v1c.v[2] & 0xfu;
  {
    unsigned __int128 hi = ((unsigned __int128)v1c.v[2] | ((unsigned __int128)v1c.v[3] << 32) |
                            ((unsigned __int128)v1c.v[4] << 64) | ((unsigned __int128)v1c.v[5] << 96)) >> 4;
    v1h.lo = (uint64_t)hi;
    v1h.hi = (uint32_t)(hi >> 64);
  }
  // rows 18-21: v0_h (n_floor leading: 49 bits = 3 chunks), v0_l, v1_h, v1_l              :208-211
  put_limb_row(o, row++, v0h, 3);
  put_limb_row(o, row++, v0l, 4);
  put_limb_row(o, row++, v1h, 3);
  put_limb_row(o, row++, v1l, 4);
  Fr v0hf = l68_to_fr(v0h), v0lf = l68_to_fr(v0l), v1hf = l68_to_fr(v1h), v1lf = l68_to_fr(v1l);
  // rows 22-25                                                                            :213-249
  put_row(o, row++, u0, l[0], l[1], zf[0], zf[1]);
  put_row(o, row++, u0, v0lf, v0hf, zero, zero);
  Fr u1s = l[2] + l[3] * e1 - zf[2] - zf[3] * e1;
  put_row(o, row++, u1s, l[2], l[3], zf[2], zf[3]);
  put_row(o, row++, u1s, v0lf, v0hf, v1lf, v1hf);
  // natives                                                                               :254-320
  Fr nx = native_of(xf), ny = sq ? nx : native_of(yf), nd = native_of(df), nz = native_of(zf);
  if (!cx) put_row(o, row++, nx, xf[0], xf[1], xf[2], xf[3]);
  if (!sq && !cy) put_row(o, row++, ny, yf[0], yf[1], yf[2], yf[3]);
  put_row(o, row++, nd, df[0], df[1], df[2], df[3]);
  if (!cz) put_row(o, row++, nz, zf[0], zf[1], zf[2], zf[3]);
  put_row(o, row++, nx, ny, nd, nz, zero);
}

__device__ void expand_reduce(const WitnessOp& op, const Cols& o) {
  const bool ca = op.flags & 1;
  const uint64_t* A = op.v;
  const uint64_t* Rm = op.v + 8;
  uint32_t row = op.row;
  const Fr zero = Fr::zero();
  uint32_t ba[10], br[10];
  limbs_to_bn(A, ba);
  limbs_to_bn(Rm, br);
  // d = (a - rem) / p < 2^17: one word of the exact quotient is enough
  uint32_t dd = (ba[0] - br[0]) * W_PINV288[0];
  // u = d * p_0 + rem_0 + 64 * 2^68 - a_0 ; v = u >> 68                                   :537-543
  unsigned __int128 p0 = (unsigned __int128)W_P_LIMB0[0] | ((unsigned __int128)W_P_LIMB0[1] << 32) | ((unsigned __int128)W_P_LIMB0[2] << 64);
  unsigned __int128 r0 = (unsigned __int128)Rm[0] | ((unsigned __int128)Rm[1] << 64);
  unsigned __int128 a0 = (unsigned __int128)A[0] | ((unsigned __int128)A[1] << 64);
  unsigned __int128 u = (unsigned __int128)dd * p0 + r0 + ((unsigned __int128)64 << 68) - a0;
  uint64_t v = (uint64_t)(u >> 68);
  L68 f[4];
  for (int i = 0; i < 4; i++) l68_from_u128(Rm + 2 * i, f[i]);
  put_assign_w(o, row, f);
  row += 4;
  Fr df = fr_from_u64(dd), vf = fr_from_u64(v);
  put_row(o, row++, df, vf, zero, zero, zero);
  Fr rf[4], af[4];
  for (int i = 0; i < 4; i++) { rf[i] = limb128_to_fr(Rm + 2 * i); af[i] = limb128_to_fr(A + 2 * i); }
  Fr nr = native_of(rf), na = native_of(af);
  put_row(o, row++, nr, rf[0], rf[1], rf[2], rf[3]);
  if (!ca) put_row(o, row++, na, af[0], af[1], af[2], af[3]);
  put_row(o, row++, na, df, nr, zero, zero);
  put_row(o, row++, df, rf[0], af[0], vf, zero);
}

// the three values is_zero inverts: the limb sum, native(a) - (p mod r), limb_0 - p_0       :53-102
__device__ __forceinline__ void iszero_values(const WitnessOp& op, Fr* af, Fr& s, Fr& na, Fr& nd, Fr& ld) {
  for (int i = 0; i < 4; i++) af[i] = limb128_to_fr(op.v + 2 * i);
  s = af[0] + af[1] + af[2] + af[3];
  na = native_of(af);
  Fr wn;
  for (int i = 0; i < 8; i++) wn.v[i] = W_NATIVE[i];
  wn = fp_to_mont(wn);
  nd = na - wn;
  Fr p0 = Fr::zero();
  p0.v[0] = W_P_LIMB0[0]; p0.v[1] = W_P_LIMB0[1]; p0.v[2] = W_P_LIMB0[2];
  p0 = fp_to_mont(p0);
  ld = af[0] - p0;
}

__global__ void __launch_bounds__(128) witness_iszero_values_kernel(const WitnessOp* __restrict__ ops, const uint32_t* __restrict__ idx,
                                                                    uint32_t n_ops, Fr* __restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ops) return;
  Fr af[4], s, na, nd, ld;
  iszero_values(ops[idx[i]], af, s, na, nd, ld);
  s.store(vals + 3 * (size_t)i);
  nd.store(vals + 3 * (size_t)i + 1);
  ld.store(vals + 3 * (size_t)i + 2);
}

// inv = the batch-inverted values of witness_iszero_values_kernel (zeros stay zero, as BaseGateOps::invert wants)
__global__ void __launch_bounds__(128) witness_iszero_rows_kernel(const WitnessOp* __restrict__ ops, const uint32_t* __restrict__ idx,
                                                                  uint32_t n_ops, const Fr* __restrict__ inv, Cols o) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ops) return;
  const WitnessOp& op = ops[idx[i]];
  const bool ca = op.flags & 1;
  uint32_t row = op.row;
  const Fr zero = Fr::zero();
  Fr af[4], s, na, nd, ld;
  iszero_values(op, af, s, na, nd, ld);
  // is_pure_zero                                                                          :53-66
  put_row(o, row++, s, af[0], af[1], af[2], af[3]);
  Fr c1 = put_invert(o, row, s, Fr::load(inv + 3 * (size_t)i));
  row += 2;
  // is_pure_w_modulus                                                                     :68-102
  if (!ca) put_row(o, row++, na, af[0], af[1], af[2], af[3]);
  put_row(o, row++, nd, na, zero, zero, zero);
  Fr c2 = put_invert(o, row, nd, Fr::load(inv + 3 * (size_t)i + 1));
  row += 2;
  put_row(o, row++, ld, af[0], zero, zero, zero);
  Fr c3 = put_invert(o, row, ld, Fr::load(inv + 3 * (size_t)i + 2));
  row += 2;
  Fr cand = c2 * c3;
  put_row(o, row++, c2, c3, cand, zero, zero);
  Fr cor = c1 + cand - c1 * cand;
  put_row(o, row++, c1, cand, cor, zero, zero);
}

// idx[off[opc] + j] = index of the j-th record with opcode opc (any order within an opcode).  One atomic per opcode
// present in a warp: the lanes with the same opcode elect a leader, which reserves their slots.
__global__ void __launch_bounds__(256) witness_group_kernel(const WitnessOp* __restrict__ ops, uint32_t n_ops, uint32_t* cursor /*[WOP_COUNT], preset to the offsets*/,
                                                            uint32_t* __restrict__ idx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n_ops;
  const uint32_t opc = live ? ops[i].opcode : 0xffffffffu;
  const unsigned peers = __match_any_sync(0xffffffffu, opc);
  if (!live || opc >= WOP_COUNT) return;
  const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(cursor + opc, (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  idx[base + __popc(peers & ((1u << lane) - 1))] = i;
}

template <uint32_t OPC>
__global__ void __launch_bounds__(128) witness_expand_kernel(const WitnessOp* __restrict__ ops, const uint32_t* __restrict__ idx, uint32_t n_ops, Cols o) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ops) return;
  const WitnessOp& op = ops[idx[i]];
  if (OPC == WOP_RAW128) {
    for (uint32_t r = 0; r < op.aux; r++) {
      const uint64_t* c = op.v + 10 * r;
      put_row(o, op.row + r, fr_from_u128(c[0], c[1]), fr_from_u128(c[2], c[3]), fr_from_u128(c[4], c[5]),
              fr_from_u128(c[6], c[7]), fr_from_u128(c[8], c[9]));
    }
  } else if (OPC == WOP_RAW256) {
    Fr cells[5];
    for (int c = 0; c < 5; c++) {
      Fr a;
      for (int k = 0; k < 4; k++) { a.v[2 * k] = (uint32_t)op.v[4 * c + k]; a.v[2 * k + 1] = (uint32_t)(op.v[4 * c + k] >> 32); }
      cells[c] = fp_to_mont(a);
    }
    put_row(o, op.row, cells[0], cells[1], cells[2], cells[3], cells[4]);
  } else if (OPC == WOP_NATIVE) {
    Fr af[4];
    for (int k = 0; k < 4; k++) af[k] = limb128_to_fr(op.v + 2 * k);
    put_row(o, op.row, native_of(af), af[0], af[1], af[2], af[3]);
  } else if (OPC == WOP_REDUCE) {
    expand_reduce(op, o);
  } else if (OPC == WOP_MULEQ) {
    expand_muleq(op, o);
  }
}

template <uint32_t OPC>
static void launch_expand(h2agg_ctx* ctx, const WitnessOp* ops, const uint32_t* idx, size_t n, const Cols& o) {
  if (!n) return;
  witness_expand_kernel<OPC><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ops, idx, (uint32_t)n, o);
  ctx->launches++;
}

// d_ops: n records on the device in recording order, counts[opc] of each opcode (known to the recorder); d_cols: 5
// device columns of n_rows Fr (zero-filled here: halo2 leaves unassigned advice cells at zero)
int witness_expand_dev(h2agg_ctx* ctx, const void* d_ops, const size_t* counts, void* const d_cols[5], size_t n_rows) {
  Cols o;
  for (int c = 0; c < 5; c++) {
    o.c[c] = (Fr*)d_cols[c];
    H2AGG_CUDA(ctx, cudaMemsetAsync(d_cols[c], 0, n_rows * 32, ctx->stream));
  }
  o.n_rows = (uint32_t)n_rows;
  size_t n_ops = 0;
  uint32_t off[WOP_COUNT];
  for (uint32_t k = 0; k < WOP_COUNT; k++) {
    off[k] = (uint32_t)n_ops;
    n_ops += counts[k];
  }
  if (!n_ops) return 0;
  const size_t niz = counts[WOP_ISZERO];
  // scratch: [ index array n_ops x 4 | cursors | values is_zero inverts ]
  const size_t o_cur = (n_ops * 4 + 255) & ~(size_t)255, o_vals = o_cur + 256;
  int rc = ensure(ctx, ctx->wit_ws, o_vals + niz * 3 * sizeof(Fr) + 256);
  if (rc) return rc;
  uint32_t* idx = (uint32_t*)ctx->wit_ws.p;
  uint32_t* cursor = (uint32_t*)((uint8_t*)ctx->wit_ws.p + o_cur);
  Fr* vals = (Fr*)((uint8_t*)ctx->wit_ws.p + o_vals);
  const WitnessOp* ops = (const WitnessOp*)d_ops;
  ScopedKernelTimer tk(ctx, KC_WITNESS, ctx->stream);
  H2AGG_CUDA(ctx, cudaMemcpyAsync(cursor, off, sizeof(off), cudaMemcpyHostToDevice, ctx->stream));   // (pageable 24 B: staged at once)
  witness_group_kernel<<<(unsigned)((n_ops + 255) / 256), 256, 0, ctx->stream>>>(ops, (uint32_t)n_ops, cursor, idx);
  ctx->launches++;
  launch_expand<WOP_MULEQ>(ctx, ops, idx + off[WOP_MULEQ], counts[WOP_MULEQ], o);
  launch_expand<WOP_REDUCE>(ctx, ops, idx + off[WOP_REDUCE], counts[WOP_REDUCE], o);
  launch_expand<WOP_RAW128>(ctx, ops, idx + off[WOP_RAW128], counts[WOP_RAW128], o);
  launch_expand<WOP_RAW256>(ctx, ops, idx + off[WOP_RAW256], counts[WOP_RAW256], o);
  launch_expand<WOP_NATIVE>(ctx, ops, idx + off[WOP_NATIVE], counts[WOP_NATIVE], o);
  if (niz) {
    witness_iszero_values_kernel<<<(unsigned)((niz + 127) / 128), 128, 0, ctx->stream>>>(ops, idx + off[WOP_ISZERO], (uint32_t)niz, vals);
    ctx->launches++;
    rc = batch_invert_dev(ctx, vals, niz * 3);
    if (rc) return rc;
    witness_iszero_rows_kernel<<<(unsigned)((niz + 127) / 128), 128, 0, ctx->stream>>>(ops, idx + off[WOP_ISZERO], (uint32_t)niz, vals, o);
    ctx->launches++;
  }
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace h2agg
