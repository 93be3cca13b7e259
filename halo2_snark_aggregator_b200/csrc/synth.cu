// Deterministic synthetic inputs generated directly in HBM (bench / test helper).
// Bit-identical to oracle_gen_scalars / oracle_gen_bases (oracle/cpu_halo2.cpp) and to
// oracle/py/bn254_ref.py::gen_scalar / gen_base: counter-based splitmix64 streams, scalars by
// rejection below r, bases by try-and-increment on y^2 = x^3 + 3 with the even root.
#include "../../include/h2agg.h"
#include "bn254_g1.cuh"
#include "ctx.hpp"

namespace h2agg {

__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t stream_seed(uint64_t seed, uint64_t index) {
  uint64_t s = seed ^ (index * 0xd1342543de82ef95ULL + 0x2545f4914f6cdd1dULL);
  splitmix64(s);
  return s;
}
template <class T>
__device__ __forceinline__ bool geq_mod(const uint32_t* v) {
  for (int i = 7; i >= 0; i--) {
    if (v[i] > T::P(i)) return true;
    if (v[i] < T::P(i)) return false;
  }
  return true;
}
template <class T>
__device__ __forceinline__ Fp<T> draw_below(uint64_t& s) {
  Fp<T> r;
  for (;;) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint64_t w = splitmix64(s);
      r.v[2 * i] = (uint32_t)w;
      r.v[2 * i + 1] = (uint32_t)(w >> 32);
    }
    r.v[7] &= 0x3fffffffu;
    if (!geq_mod<T>(r.v)) return r;
  }
}

__global__ void synth_scalars_kernel(uint64_t seed, int kind, uint64_t first, uint64_t n, Fr* out) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint64_t s = stream_seed(seed, first + k);
  uint64_t sel = splitmix64(s) % 100;
  Fr c = Fr::zero();
  if (kind == 0) {
    c = draw_below<FrTag>(s);
  } else if (kind == 1) {
    if (sel < 70) c.v[0] = (uint32_t)(splitmix64(s) & 0x1ffff);
    else if (sel < 80) c.v[0] = (uint32_t)(splitmix64(s) & 1);
  } else if (kind == 2) {
    if (sel < 50) {
      uint64_t lo = splitmix64(s), hi = splitmix64(s) & 0xf;
      c.v[0] = (uint32_t)lo; c.v[1] = (uint32_t)(lo >> 32); c.v[2] = (uint32_t)hi;
    } else if (sel < 70) {
      c = draw_below<FrTag>(s);
    }
  } else {
    c.v[0] = (uint32_t)(splitmix64(s) & 0x1ffff);
  }
  fp_to_mont(c).store(out + k);
}

__global__ void synth_bases_kernel(uint64_t seed, uint64_t first, uint64_t n, uint8_t* out) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint64_t s = stream_seed(seed, first + k);
  Fq x = fp_to_mont(draw_below<FqTag>(s));
  Fq three = Fq::one() + Fq::one() + Fq::one();
  // (p + 1) / 4
  const uint32_t e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
  for (;;) {
    Fq rhs = fp_sqr(x) * x + three;
    Fq y = fp_pow(rhs, e);
    if (fp_sqr(y) == rhs && !y.is_zero()) {
      if (fp_from_mont(y).v[0] & 1) y = fp_neg(y);
      x.store(out + k * 64);
      y.store(out + k * 64 + 32);
      return;
    }
    x = x + Fq::one();
  }
}

}  // namespace h2agg

using namespace h2agg;

extern "C" {

int h2agg_synth_scalars_dev(h2agg_ctx* ctx, uint64_t seed, int kind, uint64_t first, uint64_t n, void* d_out) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_out || kind < 0 || kind > 3) { ctx->last_error = "synth_scalars: bad argument"; return 1; }
  if (!n) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  synth_scalars_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(seed, kind, first, n, (Fr*)d_out);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

int h2agg_synth_bases_dev(h2agg_ctx* ctx, uint64_t seed, uint64_t first, uint64_t n, void* d_out) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!d_out) { ctx->last_error = "synth_bases: bad argument"; return 1; }
  if (!n) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  synth_bases_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(seed, first, n, (uint8_t*)d_out);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  return 0;
}

int h2agg_memcpy_d2h(h2agg_ctx* ctx, void* dst, const void* d_src, size_t bytes) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = bg_join(ctx)) return rc;   // a host read is a synchronisation point: deferred transforms included
  H2AGG_CUDA(ctx, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
int h2agg_memcpy_h2d(h2agg_ctx* ctx, void* d_dst, const void* src, size_t bytes) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  H2AGG_CUDA(ctx, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
int h2agg_dev_alloc(h2agg_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  H2AGG_CUDA(ctx, cudaMalloc(out, bytes));
  return 0;
}
int h2agg_dev_free(h2agg_ctx* ctx, void* p) {
  if (!ctx) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  H2AGG_CUDA(ctx, cudaFree(p));
  return 0;
}

}  // extern "C"
