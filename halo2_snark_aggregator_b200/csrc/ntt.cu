// K2/K3: NTT / iNTT / coset transforms over BN254 Fr for sm_100a.
//
// Replaces halo2_proofs `best_fft` and `EvaluationDomain::{ifft, coeff_to_extended,
// extended_to_coeff}` (external crate, restated in SURVEY.md App. B2-B3) as reached from the
// reference's `create_proof` call at halo2-snark-aggregator-circuit/src/verify_circuit.rs:986
// and `keygen_pk` at :974.  Natural order in, natural order out, like the CPU function.
//
// Decomposition: N = R_1 * R_2 * ... * R_T (T <= 4, each R_t = 2^s_t, s_t <= 8, or one pass of
// up to 2^11).  Pass t transforms digit n_t -> k_t on shared-memory tiles and multiplies by the
// inter-pass twiddle w_N^(P_t * j * k_t); layout between passes is [k_1]..[k_t][n_{t+1}]..[n_T]
// so every pass but the last reads and writes the same positions (in-place safe) in runs of
// C x 32 B, and the last pass stores transposed (k_1 fastest) in runs of G x 32 B.
// One HBM round trip per pass; the coset scaling (zeta^(i mod 3)), the zero padding n -> 4n and
// the 1/n of the inverse transform are fused into the first load / last store (K3).
// Twiddles are never streamed from HBM: a two-level table (2 x 2^(k/2) entries, L2-resident)
// gives w^e with at most one extra multiplication, stage twiddles sit in shared memory.
#include "bn254_field.cuh"
#include "ctx.hpp"
#include <cstring>

namespace h2agg {

static constexpr int NTT_THREADS = 256;
static constexpr uint32_t NTT_TILE_LOG = 11;  // 2048 elements = 64 KiB of shared memory per CTA

struct NttPassArgs {
  const uint4* src;
  uint4* dst;
  const Fr* t_lo;
  const Fr* t_hi;
  const Fr* t_full;  // w^e for every e < N (first pass only; may be null)
  unsigned long long src_n, dst_n;
  uint32_t log_n, lo_bits;
  uint32_t s;       // log2 radix of this pass
  uint32_t log_l;   // log2 of product of later radices
  uint32_t log_p;   // log2 of product of earlier radices
  uint32_t cbits;   // log2 columns per tile
  uint32_t log_r1;  // log2 radix of pass 1 (last pass only)
  uint32_t nmid;    // number of middle passes (last pass only)
  uint32_t mid_log[2];
  uint32_t npass;
  uint32_t has_in, has_out;
  uint32_t zskip;   // top `zskip` bits of the row index are zero for every non-zero input row (first pass)
  Fr in3[3];
  Fr out3[3];
};

__device__ __forceinline__ Fr get_tw(const Fr* __restrict__ t_lo, const Fr* __restrict__ t_hi, uint32_t lo_bits,
                                     uint32_t e) {
  uint32_t lo = e & ((1u << lo_bits) - 1), hi = e >> lo_bits;
  if (hi == 0) return Fr::load_nc(t_lo + lo);
  if (lo == 0) return Fr::load_nc(t_hi + hi);
  return Fr::load_nc(t_lo + lo) * Fr::load_nc(t_hi + hi);
}

__global__ void ntt_gen_full_table(const Fr* __restrict__ t_lo, const Fr* __restrict__ t_hi, uint32_t lo_bits, uint32_t log_n,
                                   Fr* __restrict__ full) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (1u << log_n)) return;
  get_tw(t_lo, t_hi, lo_bits, e).store(full + e);
}

__global__ void ntt_gen_tables(Fr omega, uint32_t lo_bits, uint32_t hi_bits, Fr* t_lo, Fr* t_hi) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nlo = 1u << lo_bits, nhi = 1u << hi_bits;
  if (i < nlo) {
    fp_pow_u64(omega, i).store(t_lo + i);
  } else if (i < nlo + nhi) {
    uint32_t j = i - nlo;
    fp_pow_u64(omega, (uint64_t)j << lo_bits).store(t_hi + j);
  }
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: the tile load of the last pass ------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

template <bool LAST>
__global__ void __launch_bounds__(NTT_THREADS, 3) ntt_pass_kernel(const __grid_constant__ NttPassArgs p) {
  extern __shared__ uint4 smem[];
  const uint32_t s = p.s, R = 1u << s, cbits = p.cbits, C = 1u << cbits, E = R << cbits;
  // Last pass of a multi-pass transform: every column of the tile is ONE contiguous run of R x 32 bytes in HBM, so the
  // tile comes in as C bulk copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) issued by one thread: no
  // LDG -> register -> STS round trip, no address arithmetic per element.  The elements then sit in shared memory as
  // they do in HBM (32-byte elements, lo half | hi half), column c at c * (2R + 1) 16-byte units: the odd pitch keeps
  // the eight columns a quarter-warp reads on distinct bank groups.  Other passes keep two 16-byte planes.
  const bool bulk = LAST && p.npass > 1;
  const uint32_t rs = LAST ? (bulk ? 2u : 1u) : C;
  const uint32_t cs = LAST ? (bulk ? 2 * R + 1 : R + 1) : 1u;
  const uint32_t plane = LAST ? C * (R + 1) : E;
  uint4* xl = smem;
  uint4* xh = bulk ? smem + 1 : smem + plane;
  uint4* twl = smem + 2 * plane;  // stage twiddles w_R^m, m < R/2 (two 16-byte planes)
  uint4* twh = twl + (R >> 1);
  __shared__ __align__(8) uint64_t tile_bar;
  const uint32_t tid = threadIdx.x;
  const unsigned long long tile = blockIdx.x;

  // ---- tile coordinates
  unsigned long long base;  // position of (row 0, col 0)
  uint32_t j0 = 0;          // first column index j (non-last) / first k_1 (last)
  uint32_t log_q = 0, rest = 0;
  if (!LAST) {
    uint32_t tiles_per_a = 1u << (p.log_l - cbits);
    unsigned long long a = tile >> (p.log_l - cbits);
    j0 = (uint32_t)(tile & (tiles_per_a - 1)) << cbits;
    base = (a << (s + p.log_l)) + j0;
  } else {
    // A = k_1 * Q + rest, Q = P_T / R_1; tile -> (k1 group, rest)
    log_q = p.log_p - p.log_r1;
    rest = (uint32_t)(tile & ((1ull << log_q) - 1));
    j0 = (uint32_t)(tile >> log_q) << cbits;
    base = 0;
  }

  // ---- load.  Non-last passes: rows land bit-reversed and DIT stages follow (natural order out).
  // Last pass: rows land in natural order (contiguous, conflict-free) and DIF stages follow; the
  // store then reads row brev(k).  Rows r >= R >> zskip are known to be zero (zero-padded coset
  // input): they are neither loaded nor computed -- after the bit reversal the first `zskip` DIT
  // stages degenerate to copies, so each loaded value is replicated 2^zskip times instead.
  const uint32_t zs = LAST ? 0u : p.zskip;
  if (bulk && tid == 0) {   // the copies are in flight while the CTA stages its twiddles
    mbar_init(&tile_bar, 1);
    mbar_arrive_expect_tx(&tile_bar, E * 32);
    for (uint32_t c = 0; c < C; c++) {
      const unsigned long long pos0 = ((((unsigned long long)(j0 + c)) << log_q) + rest) << s;   // row 0 of column c
      bulk_g2s(smem + (size_t)c * cs, p.src + 2 * pos0, R * 32, &tile_bar);
    }
  }
  // ---- stage twiddles: w_R^m = w_N^(m * N/R)
  for (uint32_t m = tid; m < (R >> 1); m += NTT_THREADS) {
    Fr w = get_tw(p.t_lo, p.t_hi, p.lo_bits, m << (p.log_n - s));
    twl[m] = w.lo4();
    twh[m] = w.hi4();
  }
  if (bulk) {
    __syncthreads();          // the barrier's initialisation is visible to every waiter
    mbar_wait(&tile_bar, 0);
  } else {
#pragma unroll 4
  for (uint32_t idx = tid; idx < (E >> zs); idx += NTT_THREADS) {
    uint32_t r, c;
    unsigned long long pos;
    if (!LAST) {
      c = idx & (C - 1);
      r = idx >> cbits;
      pos = base + ((unsigned long long)r << p.log_l) + c;
    } else {
      r = idx & (R - 1);
      c = idx >> s;
      pos = (((((unsigned long long)(j0 + c)) << log_q) + rest) << s) + r;
    }
    Fr v;
    if (pos < p.src_n) {
      v = Fr::from_halves(p.src[2 * pos], p.src[2 * pos + 1]);
      if (p.has_in) {
        uint32_t m3 = (uint32_t)(pos % 3);
        if (m3) v = v * p.in3[m3];
      }
    } else {
      v = Fr::zero();
    }
    if (!LAST) {
      uint32_t rr = __brev(r) >> (32 - s);
      for (uint32_t t = 0; t < (1u << zs); t++) {
        uint32_t si = (rr + t) * rs + c * cs;
        xl[si] = v.lo4();
        xh[si] = v.hi4();
      }
    } else {
      uint32_t si = r * rs + c * cs;
      xl[si] = v.lo4();
      xh[si] = v.hi4();
    }
  }
  }

  // ---- butterfly stages in shared memory, two at a time (radix-4 in registers): a thread owns the
  // four rows {i, i+h, i+2h, i+3h} of one column, so every second shared-memory round trip and
  // __syncthreads() disappears and four independent products are in flight per thread.
  // Thread -> (j, g, c) with c fastest and the twiddle index j slowest: a warp shares its twiddles and the
  // w = 1 cases (j == 0) skip their products warp-uniformly.  An odd leftover stage runs radix-2.
  uint32_t st = zs;
  if ((s - zs) & 1) {  // single radix-2 stage first
    __syncthreads();
    const uint32_t lh = LAST ? (s - 1 - st) : st;
    const uint32_t h = 1u << lh;
    const uint32_t gbits = s - 1 - lh;
#pragma unroll 2
    for (uint32_t bb = tid; bb < (E >> 1); bb += NTT_THREADS) {
      uint32_t c = bb & (C - 1);
      uint32_t q = bb >> cbits;
      uint32_t g = q & ((1u << gbits) - 1);
      uint32_t j = q >> gbits;
      uint32_t i = (g << (lh + 1)) + j;
      uint32_t i0 = i * rs + c * cs, i1 = (i + h) * rs + c * cs;
      Fr x0 = Fr::from_halves(xl[i0], xh[i0]);
      Fr x1 = Fr::from_halves(xl[i1], xh[i1]);
      Fr y0, y1;
      if (!LAST) {
        if (j) {
          uint32_t m = j << gbits;
          x1 = x1 * Fr::from_halves(twl[m], twh[m]);
        }
        y0 = x0 + x1;
        y1 = x0 - x1;
      } else {
        y0 = x0 + x1;
        y1 = x0 - x1;
        if (j) {
          uint32_t m = j << gbits;
          y1 = y1 * Fr::from_halves(twl[m], twh[m]);
        }
      }
      xl[i0] = y0.lo4(); xh[i0] = y0.hi4();
      xl[i1] = y1.lo4(); xh[i1] = y1.hi4();
    }
    st++;
  }
  for (; st < s; st += 2) {
    __syncthreads();
    // DIT: strides h then 2h with lh = st; DIF (last pass): strides 2h then h with lh = s - 2 - st
    const uint32_t lh = LAST ? (s - 2 - st) : st;
    const uint32_t h = 1u << lh;
    const uint32_t gbits = s - 2 - lh;  // log2 of the number of groups R / 4h
    for (uint32_t bb = tid; bb < (E >> 2); bb += NTT_THREADS) {
      const uint32_t c = bb & (C - 1);
      const uint32_t q = bb >> cbits;
      const uint32_t g = q & ((1u << gbits) - 1);
      const uint32_t j = q >> gbits;  // < h
      const uint32_t i = (g << (lh + 2)) + j;
      const uint32_t i0 = i * rs + c * cs, i1 = (i + h) * rs + c * cs, i2 = (i + 2 * h) * rs + c * cs,
                     i3 = (i + 3 * h) * rs + c * cs;
      Fr x0 = Fr::from_halves(xl[i0], xh[i0]);
      Fr x1 = Fr::from_halves(xl[i1], xh[i1]);
      Fr x2 = Fr::from_halves(xl[i2], xh[i2]);
      Fr x3 = Fr::from_halves(xl[i3], xh[i3]);
      // twiddles: w1 = w_R^(j R/2h) (stride h), w2 = w_R^(j R/4h), w3 = w_R^((j+h) R/4h) (stride 2h)
      const uint32_t m1 = j << (gbits + 1), m2 = j << gbits, m3 = (j + h) << gbits;
      const Fr w3 = Fr::from_halves(twl[m3], twh[m3]);
      Fr y0, y1, y2, y3;
      if (!LAST) {
        if (j) {
          const Fr w1 = Fr::from_halves(twl[m1], twh[m1]);
          x1 = x1 * w1;
          x3 = x3 * w1;
        }
        Fr a0 = x0 + x1, a1 = x0 - x1, a2 = x2 + x3, a3 = x2 - x3;
        if (j) a2 = a2 * Fr::from_halves(twl[m2], twh[m2]);
        a3 = a3 * w3;
        y0 = a0 + a2; y2 = a0 - a2; y1 = a1 + a3; y3 = a1 - a3;
      } else {
        Fr b0 = x0 + x2, b2 = x0 - x2, b1 = x1 + x3, b3 = x1 - x3;
        if (j) b2 = b2 * Fr::from_halves(twl[m2], twh[m2]);
        b3 = b3 * w3;
        y0 = b0 + b1; y1 = b0 - b1; y2 = b2 + b3; y3 = b2 - b3;
        if (j) {
          const Fr w1 = Fr::from_halves(twl[m1], twh[m1]);
          y1 = y1 * w1;
          y3 = y3 * w1;
        }
      }
      xl[i0] = y0.lo4(); xh[i0] = y0.hi4();
      xl[i1] = y1.lo4(); xh[i1] = y1.hi4();
      xl[i2] = y2.lo4(); xh[i2] = y2.hi4();
      xl[i3] = y3.lo4(); xh[i3] = y3.hi4();
    }
  }
  __syncthreads();

  // ---- store
  if (!LAST) {
    for (uint32_t idx = tid; idx < E; idx += NTT_THREADS) {
      uint32_t c = idx & (C - 1), k = idx >> cbits;
      unsigned long long pos = base + ((unsigned long long)k << p.log_l) + c;
      uint32_t si = k * rs + c * cs;
      Fr v = Fr::from_halves(xl[si], xh[si]);
      uint32_t jj = j0 + c;
      if (k && jj) {
        // w_N^(P_t * j * k_t); j*k_t < M_t so the exponent is < N
        uint32_t e = (uint32_t)(((unsigned long long)jj * k) << p.log_p);
        // first pass: e ranges over the whole domain -> one 32-byte gather from the full table instead of
        // T_lo * T_hi (HBM bytes are cheap here, multiplier cycles are not); later passes hit T_hi directly
        v = v * (p.t_full ? Fr::load_nc(p.t_full + e) : get_tw(p.t_lo, p.t_hi, p.lo_bits, e));
      }
      p.dst[2 * pos] = v.lo4();
      p.dst[2 * pos + 1] = v.hi4();
    }
  } else {
    // revmix of the middle digits of `rest`
    uint32_t mid = 0;
    {
      uint32_t consumed = 0, outshift = 0;
      for (uint32_t i = 0; i < p.nmid; i++) consumed += p.mid_log[i];
      for (uint32_t i = 0; i < p.nmid; i++) {
        consumed -= p.mid_log[i];
        uint32_t d = (rest >> consumed) & ((1u << p.mid_log[i]) - 1);
        mid |= d << outshift;
        outshift += p.mid_log[i];
      }
    }
    unsigned long long obase = (p.npass == 1) ? 0ull : ((unsigned long long)j0 + ((unsigned long long)mid << p.log_r1));
    for (uint32_t idx = tid; idx < E; idx += NTT_THREADS) {
      uint32_t c = idx & (C - 1), k = idx >> cbits;
      unsigned long long pos = obase + c + ((unsigned long long)k << p.log_p);
      if (pos >= p.dst_n) continue;
      uint32_t si = (__brev(k) >> (32 - s)) * rs + c * cs;
      Fr v = Fr::from_halves(xl[si], xh[si]);
      if (p.has_out) v = v * p.out3[(uint32_t)(pos % 3)];
      p.dst[2 * pos] = v.lo4();
      p.dst[2 * pos + 1] = v.hi4();
    }
  }
}

int plan_passes(const h2agg_ctx* ctx, uint32_t log_n, uint32_t* s);

static int get_tables(h2agg_ctx* ctx, const uint64_t* omega, uint32_t log_n, TwiddleTable** out) {
  ctx->tick++;
  for (auto& t : ctx->tw) {
    if (t.log_n == log_n && memcmp(t.omega, omega, 32) == 0) {
      t.last_use = ctx->tick;
      *out = &t;
      return 0;
    }
  }
  if (ctx->tw.size() >= 12) {  // evict least recently used
    size_t v = 0;
    for (size_t i = 1; i < ctx->tw.size(); i++)
      if (ctx->tw[i].last_use < ctx->tw[v].last_use) v = i;
    // lanes may still be reading the victim: drain the device, not just the main stream
    H2AGG_CUDA(ctx, cudaDeviceSynchronize());
    cudaFree(ctx->tw[v].lo);
    cudaFree(ctx->tw[v].hi);
    cudaFree(ctx->tw[v].full);
    ctx->tw.erase(ctx->tw.begin() + v);
  }
  TwiddleTable t;
  memcpy(t.omega, omega, 32);
  t.log_n = log_n;
  // lo_bits = log2(R_1): every pass after the first needs w^(P_t x) with P_t a multiple of R_1,
  // i.e. a direct T_hi hit; only the first pass pays one multiplication to combine T_lo * T_hi.
  {
    uint32_t sp[4];
    int T = plan_passes(ctx, log_n, sp);
    t.lo_bits = (T == 1) ? (log_n + 1) / 2 : sp[0];
  }
  uint32_t hi_bits = log_n - t.lo_bits;
  H2AGG_CUDA(ctx, cudaMalloc(&t.lo, sizeof(Fr) << t.lo_bits));
  H2AGG_CUDA(ctx, cudaMalloc(&t.hi, sizeof(Fr) << hi_bits));
  Fr w;
  memcpy(w.v, omega, 32);
  uint32_t total = (1u << t.lo_bits) + (1u << hi_bits);
  ntt_gen_tables<<<(total + 127) / 128, 128, 0, ctx->stream>>>(w, t.lo_bits, hi_bits, (Fr*)t.lo, (Fr*)t.hi);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  // full table for the first pass of multi-pass transforms (N x 32 B: 128 MiB at 2^22, 512 MiB at 2^24)
  t.full = nullptr;
  if (log_n > NTT_TILE_LOG && log_n <= 26 && ctx->ntt_full_tables) {
    if (cudaMalloc(&t.full, sizeof(Fr) << log_n) == cudaSuccess) {
      ntt_gen_full_table<<<(1u << log_n) / 256, 256, 0, ctx->stream>>>((const Fr*)t.lo, (const Fr*)t.hi, t.lo_bits, log_n, (Fr*)t.full);
      ctx->launches++;
      H2AGG_CUDA(ctx, cudaGetLastError());
    } else {
      cudaGetLastError();
      t.full = nullptr;  // out of memory: the two-level tables still work
    }
  }
  t.last_use = ctx->tick;
  ctx->tw.push_back(t);
  *out = &ctx->tw.back();
  return 0;
}

// Create (or touch) the cached tables of (omega, log_n) on ctx->stream.  Callers that run transforms on lane streams
// call this BEFORE they fork the lanes off ctx->stream, so the generation kernels are ordered before every lane.
int ntt_warm_tables(h2agg_ctx* ctx, const uint64_t* omega, uint32_t log_n) {
  if (log_n == 0 || log_n > 28) return 0;
  TwiddleTable* tw;
  return get_tables(ctx, omega, log_n, &tw);
}

// split log_n into pass radices
int plan_passes(const h2agg_ctx* ctx, uint32_t log_n, uint32_t* s) {
  // ntt_radix_cap (test hook, default 8) lowers the largest per-pass radix so the 2/3/4-pass code paths
  // can be exercised at sizes the CPU oracle finishes in seconds
  const uint32_t cap = ctx->ntt_radix_cap;
  if (log_n <= (cap >= 8 ? NTT_TILE_LOG : cap)) {
    s[0] = log_n;
    return 1;
  }
  int T = (log_n + cap - 1) / cap;
  uint32_t base = log_n / T, extra = log_n % T;
  for (int t = 0; t < T; t++) s[t] = base + (t < (int)extra ? 1 : 0);
  return T;
}

int ntt_run(h2agg_ctx* ctx, const void* d_src, void* d_dst, const NttOpts& o, cudaStream_t st_in, DevBuf* tmpbuf_in) {
  cudaStream_t st = st_in ? st_in : ctx->stream;
  DevBuf& tmpbuf = tmpbuf_in ? *tmpbuf_in : ctx->ntt_tmp;
  if (o.log_n > 28) {
    ctx->last_error = "ntt: log_n exceeds the 2-adicity of BN254 Fr (28)";
    return 1;
  }
  const size_t N = (size_t)1 << o.log_n;
  if (o.src_n > N || o.dst_n > N) {
    ctx->last_error = "ntt: src_n/dst_n exceed the domain";
    return 1;
  }
  if (o.log_n == 0) {
    if (d_src != d_dst && o.dst_n) H2AGG_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, 32, cudaMemcpyDeviceToDevice, st));
    if (o.out_scale3 || o.in_coset3) {
      // a 1-point transform is the identity; scaling a single element is not needed by any caller
      if (o.out_scale3) {
        ctx->last_error = "ntt: log_n = 0 with scaling is not supported";
        return 1;
      }
    }
    return 0;
  }
  TwiddleTable* tw;
  int rc = get_tables(ctx, o.omega, o.log_n, &tw);
  if (rc) return rc;

  uint32_t s[4];
  int T = plan_passes(ctx, o.log_n, s);
  if (T > 4) {
    ctx->last_error = "ntt: more than 4 passes needed (lower the size or raise the radix cap)";
    return 1;
  }
  const void* cur = d_src;
  void* tmp = nullptr;
  if (T > 1) {
    rc = ensure(ctx, tmpbuf, N * 32);
    if (rc) return rc;
    tmp = tmpbuf.p;
  }
  if (!ctx->ntt_attr_set) {
    H2AGG_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    H2AGG_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    ctx->ntt_attr_set = true;
  }

  uint32_t log_p = 0;
  for (int t = 0; t < T; t++) {
    NttPassArgs a;
    memset(&a, 0, sizeof(a));
    bool last = (t == T - 1);
    a.src = (const uint4*)cur;
    a.dst = (uint4*)(last ? d_dst : tmp);
    a.t_lo = (const Fr*)tw->lo;
    a.t_hi = (const Fr*)tw->hi;
    a.t_full = (t == 0 && !last) ? (const Fr*)tw->full : nullptr;
    a.log_n = o.log_n;
    a.lo_bits = tw->lo_bits;
    a.s = s[t];
    a.log_p = log_p;
    a.log_l = o.log_n - log_p - s[t];
    a.npass = T;
    a.src_n = (t == 0) ? o.src_n : N;
    a.dst_n = last ? o.dst_n : N;
    if (t == 0 && !last) {
      uint32_t z = 0;
      while (z < s[0] && (o.src_n << (z + 1)) <= N) z++;
      a.zskip = z;
    }
    if (t == 0 && o.in_coset3) {
      a.has_in = 1;
      memcpy(a.in3, o.in_coset3, 96);
    }
    if (last && o.out_scale3) {
      a.has_out = 1;
      memcpy(a.out3, o.out_scale3, 96);
    }
    ScopedKernelTimer tk(ctx, KC_NTT_PASS, st);
    uint32_t cb = NTT_TILE_LOG - s[t];
    size_t smem;
    unsigned long long tiles;
    if (!last) {
      if (cb > a.log_l) cb = a.log_l;
      a.cbits = cb;
      tiles = (unsigned long long)N >> (s[t] + cb);
      smem = ((size_t)2 << (s[t] + cb)) * 16 + ((size_t)1 << s[t]) * 16;
      ntt_pass_kernel<false><<<(unsigned)tiles, NTT_THREADS, smem, st>>>(a);
    } else {
      a.log_r1 = (T == 1) ? 0 : s[0];
      if (cb > a.log_r1) cb = a.log_r1;
      if (T == 1) cb = 0;
      a.cbits = cb;
      a.nmid = (T >= 2) ? (uint32_t)(T - 2) : 0;
      for (uint32_t i = 0; i < a.nmid; i++) a.mid_log[i] = s[1 + i];
      tiles = (unsigned long long)N >> (s[t] + cb);
      smem = ((size_t)2 << cb) * (((size_t)1 << s[t]) + 1) * 16 + ((size_t)1 << s[t]) * 16;
      ntt_pass_kernel<true><<<(unsigned)tiles, NTT_THREADS, smem, st>>>(a);
    }
    ctx->launches++;
    H2AGG_CUDA(ctx, cudaGetLastError());
    log_p += s[t];
    cur = tmp;
  }
  return 0;
}

}  // namespace h2agg
