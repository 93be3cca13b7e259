// N4 (data formats either side of the path): the G1 point codec of the stage files.
//
// `ParamsKZG::read` / `write` (halo2_proofs poly/kzg/commitment.rs, external crate) store g and g_lagrange as COMPRESSED
// points -- halo2curves 0.2.1 `G1Affine::to_bytes` / `from_bytes`: 32-byte little-endian x with the
// parity of y in bit 7 of byte 31, the identity as 32 zero bytes.  The reference reads the 2 x 2^k points of
// `verify_circuit.params` at the start of every verify_run (halo2-snark-aggregator-circuit/src/fs.rs:109-115,
// halo2-snark-aggregator-sdk/src/lib.rs:129-148) and of HALO2_PARAMS_k in get_params_cached (verify_circuit.rs:701-731);
// decompression is one square root in Fq per point (a 254-bit exponentiation: p = 3 mod 4, y = (x^3 + 3)^((p+1)/4)),
// i.e. 8.4 M of them at k = 22 -- seconds of host time per run, milliseconds here.  The inner proofs' transcript points
// (halo2-snark-aggregator-api/src/systems/halo2/transcript.rs:63-65) use the same encoding.
#include "../../include/h2agg.h"
#include "bn254_field.cuh"
#include "ctx.hpp"
#include <cstring>

namespace h2agg {

// (p + 1) / 4, little-endian 32-bit words
__device__ __constant__ uint32_t FQ_SQRT_EXP[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u,
                                                   0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};

__global__ void __launch_bounds__(128) g1_decompress_kernel(const uint32_t* __restrict__ in, Fq* __restrict__ out, size_t n,
                                                             uint32_t* bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x;
#pragma unroll
  for (int j = 0; j < 8; j++) x.v[j] = in[8 * i + j];
  const uint32_t ysign = x.v[7] >> 31;
  x.v[7] &= 0x7fffffffu;
  bool ge = true;  // x >= p: Fq::from_bytes returns None
#pragma unroll
  for (int j = 7; j >= 0; j--) {
    if (x.v[j] != FqTag::P(j)) { ge = x.v[j] > FqTag::P(j); break; }
  }
  if (x.is_zero() && !ysign) {  // the identity
    Fq::zero().store(out + 2 * i);
    Fq::zero().store(out + 2 * i + 1);
    return;
  }
  Fq xm = fp_to_mont(x);
  Fq three = Fq::one() + Fq::one() + Fq::one();
  Fq a = fp_sqr(xm) * xm + three;
  uint32_t e[8];
#pragma unroll
  for (int j = 0; j < 8; j++) e[j] = FQ_SQRT_EXP[j];
  Fq y = fp_pow(a, e);
  if (ge || !(fp_sqr(y) == a)) {
    atomicAdd(bad, 1u);
    Fq::zero().store(out + 2 * i);
    Fq::zero().store(out + 2 * i + 1);
    return;
  }
  const uint32_t parity = fp_from_mont(y).v[0] & 1u;
  if (parity != ysign) y = fp_neg(y);
  xm.store(out + 2 * i);
  y.store(out + 2 * i + 1);
}

__global__ void __launch_bounds__(256) g1_compress_kernel(const Fq* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x = Fq::load(in + 2 * i), y = Fq::load(in + 2 * i + 1);
  Fq xc = Fq::zero();
  if (!(x.is_zero() && y.is_zero())) {
    xc = fp_from_mont(x);
    xc.v[7] |= (fp_from_mont(y).v[0] & 1u) << 31;
  }
#pragma unroll
  for (int j = 0; j < 8; j++) out[8 * i + j] = xc.v[j];
}

}  // namespace h2agg

using namespace h2agg;

static int decompress_dev(h2agg_ctx* ctx, const void* d_in, void* d_out, size_t n, uint32_t* nbad_out) {
  uint32_t* bad = (uint32_t*)((uint8_t*)ctx->small.p + 8192 + 64);
  H2AGG_CUDA(ctx, cudaMemsetAsync(bad, 0, 4, ctx->stream));
  g1_decompress_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t*)d_in, (Fq*)d_out, n, bad);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  H2AGG_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ctx->pinned + 8192 + 64, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(nbad_out, (uint8_t*)ctx->pinned + 8192 + 64, 4);
  return 0;
}

extern "C" {

int h2agg_g1_decompress_dev(h2agg_ctx* ctx, const void* d_in, void* d_out_affine, size_t n) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if ((!d_in || !d_out_affine) && n) { ctx->last_error = "g1_decompress: null argument"; return 1; }
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  uint32_t nbad = 0;
  int rc = decompress_dev(ctx, d_in, d_out_affine, n, &nbad);
  if (rc) return rc;
  if (nbad) {
    ctx->last_error = "g1_decompress: " + std::to_string(nbad) + " encoding(s) are not points of the curve (G1Affine::from_bytes returns None)";
    return 4;
  }
  return 0;
}

int h2agg_g1_decompress(h2agg_ctx* ctx, const uint8_t* in, uint64_t* out_affine, size_t n) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if ((!in || !out_affine) && n) { ctx->last_error = "g1_decompress: null argument"; return 1; }
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure(ctx, ctx->io_a, n * 32);
  if (rc) return rc;
  if ((rc = ensure(ctx, ctx->io_b, n * 64))) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, in, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t nbad = 0;
  if ((rc = decompress_dev(ctx, ctx->io_a.p, ctx->io_b.p, n, &nbad))) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(out_affine, ctx->io_b.p, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (nbad) {
    ctx->last_error = "g1_decompress: " + std::to_string(nbad) + " encoding(s) are not points of the curve (G1Affine::from_bytes returns None)";
    return 4;
  }
  return 0;
}

int h2agg_g1_compress(h2agg_ctx* ctx, const uint64_t* affine, uint8_t* out, size_t n) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if ((!affine || !out) && n) { ctx->last_error = "g1_compress: null argument"; return 1; }
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure(ctx, ctx->io_a, n * 32);
  if (rc) return rc;
  if ((rc = ensure(ctx, ctx->io_b, n * 64))) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, affine, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  g1_compress_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const Fq*)ctx->io_b.p, (uint32_t*)ctx->io_a.p, n);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  H2AGG_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_a.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
