// Library context: device, stream, scratch arenas, cached twiddle tables, registered SRS.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace h2agg {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct TwiddleTable {
  uint64_t omega[4];
  uint32_t log_n;
  uint32_t lo_bits;  // T_lo has 2^lo_bits entries (omega^i), T_hi has 2^(log_n-lo_bits) (omega^(i<<lo_bits))
  void* lo = nullptr;
  void* hi = nullptr;
  void* full = nullptr;  // omega^e, e < 2^log_n (first-pass twiddles), optional
  uint64_t last_use = 0;
};

struct Srs {
  const void* d_bases = nullptr;  // n x 64 B affine, Montgomery
  size_t n = 0;
  bool owned = false;
  void* d_table = nullptr;        // table mode: table_nwin rows of n affine points, row w = 2^(c w) * bases
  int table_c = 0, table_nwin = 0;
};

// what an MSM call runs against
struct MsmBases {
  const void* d_bases = nullptr;
  const void* d_table = nullptr;
  int table_c = 0, table_nwin = 0;
  size_t srs_n = 0;
};

// A lane = one auxiliary stream with its own MSM workspace and staging buffer.  Batches alternate
// lanes so the latency-bound tail of MSM i (window sums, Horner, inversion) and the H2D copy of
// column i+1 overlap with the IMAD-bound bucket accumulation of its neighbour.
struct Lane {
  cudaStream_t st = nullptr;
  DevBuf ws;        // MSM workspace
  DevBuf io;        // staging for host-pointer batches
  DevBuf io_out;    // NTT batches: result staging (coset transforms are out of place)
  DevBuf ntt_tmp;   // NTT ping-pong buffer of this lane
  DevBuf scan_ws;   // grand-product scratch of this lane (batched lookup products)
  DevBuf args_ws;   // numerators / denominators of this lane
  cudaEvent_t done = nullptr;
  cudaEvent_t up = nullptr;   // "this lane's column is in HBM": lets a second lane start the column's transforms
};
// Measured on B200 (k = 18 / 22 schedule): 3 lanes 47.7 / 337.5 ms, 6 lanes 40.2 / 331.7, 12 lanes 38.7 / 331.7 -- the
// small sizes are bound by the ~20-launch latency chain of an MSM, which more lanes overlap; buffers are allocated
// lazily per lane actually used.
static constexpr int N_LANES = 8;

}  // namespace h2agg

namespace h2agg {
// device of the first context created in this process (-1: none yet): the witness recorder, which has no context of its
// own, page-locks its record chunks only once a device is known to exist
extern std::atomic<int> g_any_device;
}  // namespace h2agg

struct h2agg_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::recursive_mutex mu;  // the reference calls best_fft from rayon workers: serialise entry
  std::string last_error;
  // scratch
  h2agg::DevBuf ntt_tmp;      // ping-pong buffer for multi-pass NTT
  h2agg::DevBuf io_a, io_b;   // staging for host-pointer entry points
  h2agg::DevBuf msm_ws;       // MSM workspace (digits, sort, buckets) of the main stream
  h2agg::Lane lanes[h2agg::N_LANES];
  cudaEvent_t fork_ev = nullptr;
  h2agg::DevBuf small;        // small constants / results
  h2agg::DevBuf poly_ws;      // recursion levels of eval_polynomial / kate_division
  h2agg::DevBuf poly_many_ws; // eval_polynomials: two level buffers for 64 polynomials at a time
  h2agg::DevBuf scan_ws;      // batch_invert / grand_product scratch
  h2agg::DevBuf wit_ws;       // witness expansion: index array grouping the records by opcode + the values is_zero inverts
  h2agg::DevBuf args_ws;      // lookup / permutation products: numerators and denominators
  h2agg::DevBuf args_meta;    // compress_expressions: device copies of expression lists (four slots)
  int args_flip = 0;
  h2agg::DevBuf sort_ws;      // sort_fr / permute_expression_pair: key ping-pong buffers, histograms, flags
  h2agg::DevBuf quot_ws;      // evaluate_h: device copy of plan / column pointers / constants (two halves)
  h2agg::DevBuf quot_tw;      // evaluate_h: omega_ext^i two-level table
  uint64_t quot_omega[4] = {0, 0, 0, 0};
  uint32_t quot_ext_k = 0;
  int quot_flip = 0;
  // deferred transforms (h2agg_set_defer_transforms): the NTTs of a commit round run on a background stream
  bool defer_transforms = false;
  bool bg_pending = false;          // something was enqueued on bg_stream since the last join
  cudaStream_t bg_stream = nullptr;
  cudaEvent_t bg_ev = nullptr;
  h2agg::DevBuf bg_ntt_tmp;         // the background stream's own NTT ping-pong buffer
  void* pinned = nullptr;     // pinned host bounce buffer for tiny results
  size_t pinned_cap = 0;
  std::vector<h2agg::TwiddleTable> tw;
  uint64_t tick = 0;
  std::unordered_map<uint64_t, h2agg::Srs> srs;
  uint64_t next_srs = 1;
  int sm_count = 148;
  bool ntt_attr_set = false;
  uint32_t ntt_radix_cap = 8;   // largest log2 radix of a pass (test hook: 4..8)
  bool ntt_full_tables = true;  // trade N x 32 B of HBM per (omega, k) for one product per element in the first pass  // dynamic shared-memory opt-in done for this device
  // counters (claimed in bench.py as gpu_launches)
  uint64_t launches = 0;
  // MSM tuning (0 = auto; a forced width also forces plain mode)
  int msm_window_bits = 0;
  bool srs_precompute = true;  // build the 2^(c w) P table when an SRS is registered
  // per-kernel-class device timing (CUDA events on ctx->stream), enabled by h2agg_kernel_timing
  bool timing = false;
  struct Timed { cudaEvent_t a, b; int cls; };
  std::vector<Timed> timed;
  std::vector<cudaEvent_t> ev_pool;
};

namespace h2agg {

#define H2AGG_CUDA(ctx, call)                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      (ctx)->last_error = std::string(#call) + ": " + cudaGetErrorString(e_) + " @" __FILE__ ":" + \
                          std::to_string(__LINE__);                                           \
      return 2;                                                                               \
    }                                                                                         \
  } while (0)

inline int ensure(h2agg_ctx* ctx, DevBuf& b, size_t bytes) {
  if (b.cap >= bytes) return 0;
  if (b.p) {
    H2AGG_CUDA(ctx, cudaDeviceSynchronize());
    H2AGG_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + (bytes >> 3);  // slack so growing sweeps do not realloc every size
  if (cudaMalloc(&b.p, want) != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    H2AGG_CUDA(ctx, cudaMalloc(&b.p, want));
  }
  b.cap = want;
  return 0;
}

// kernel classes for the timing hook
enum KernelClass { KC_MSM_ACCUMULATE = 0, KC_MSM_DIGITS = 1, KC_MSM_REDUCE = 2, KC_NTT_PASS = 3, KC_MSM_TOTAL = 4, KC_WITNESS = 5, KC_QUOTIENT = 6, KC_COUNT = 7 };

struct ScopedKernelTimer {
  h2agg_ctx* ctx;
  h2agg_ctx::Timed t;
  bool on;
  cudaStream_t st;
  ScopedKernelTimer(h2agg_ctx* c, int cls, cudaStream_t s = nullptr) : ctx(c), on(c->timing), st(s ? s : c->stream) {
    if (!on) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!ctx->ev_pool.empty()) { e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    t.a = get(); t.b = get(); t.cls = cls;
    cudaEventRecord(t.a, st);
  }
  ~ScopedKernelTimer() {
    if (!on) return;
    cudaEventRecord(t.b, st);
    ctx->timed.push_back(t);
  }
};

// Make ctx->stream wait for everything enqueued on the background stream (no-op when nothing is pending).
int bg_join(h2agg_ctx* ctx);
// Background stream, created on first use; `after` (optional) is an event it must wait for.
int bg_stream_get(h2agg_ctx* ctx, cudaStream_t* out);

// ---- internal device-pointer API (all asynchronous on ctx->stream) -------------------------
// NTT over Fr. `src` has src_n valid elements (rest of the 2^log_n domain is implicit zero),
// result lands in `dst` (may equal src), first dst_n elements written.
struct NttOpts {
  const uint64_t* omega;            // host, 4 limbs, Montgomery
  uint32_t log_n;
  size_t src_n, dst_n;
  const uint64_t* in_coset3;        // host, 3x4 limbs: multiply input i by in_coset3[i%3]   (or null)
  const uint64_t* out_scale3;       // host, 3x4 limbs: multiply output i by out_scale3[i%3] (or null)
};
int ntt_run(h2agg_ctx* ctx, const void* d_src, void* d_dst, const NttOpts& o, cudaStream_t st = nullptr,
            DevBuf* tmp = nullptr);

// Make sure the cached twiddle tables of (omega, log_n) exist; generation is enqueued on ctx->stream.
int ntt_warm_tables(h2agg_ctx* ctx, const uint64_t* omega, uint32_t log_n);

// Fork the lanes off ctx->stream and join them back.  The guard's destructor covers the error returns in between:
// work already enqueued on the lanes (copies into caller buffers included) is drained before the entry point returns.
struct LaneFork {
  h2agg_ctx* ctx;
  bool joined = false;
  explicit LaneFork(h2agg_ctx* c) : ctx(c) {}
  int fork() {
    H2AGG_CUDA(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));
    for (int l = 0; l < N_LANES; l++) H2AGG_CUDA(ctx, cudaStreamWaitEvent(ctx->lanes[l].st, ctx->fork_ev, 0));
    return 0;
  }
  int join() {
    joined = true;
    for (int l = 0; l < N_LANES; l++) {
      H2AGG_CUDA(ctx, cudaEventRecord(ctx->lanes[l].done, ctx->lanes[l].st));
      H2AGG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lanes[l].done, 0));
    }
    return 0;
  }
  ~LaneFork() {
    if (joined) return;
    for (int l = 0; l < N_LANES; l++)
      if (ctx->lanes[l].st) cudaStreamSynchronize(ctx->lanes[l].st);
  }
};

// MSM over G1: d_scalars n x 32 B (Montgomery Fr), d_bases n x 64 B affine.
// Writes affine (64 B) + jacobian (96 B, z = 1 or 0) to d_out (160 B, device).
// Windows [win_begin, win_end) only (pass 0, -1 for all): partial = sum_w 2^(c w) B_w.
// normalize = false leaves the XYZZ result (128 B) in the slot: the caller normalises a whole round with g1_normalize.
int msm_run(h2agg_ctx* ctx, cudaStream_t st, DevBuf& ws, const MsmBases& bases, const void* d_scalars, size_t n,
            void* d_out160, int win_begin, int win_end, bool normalize = true);
// XYZZ results in 160-byte slots -> affine + normalised Jacobian, one inversion for all n.
int g1_normalize(h2agg_ctx* ctx, cudaStream_t st, void* d_out160s, size_t n);
int msm_build_srs_table(h2agg_ctx* ctx, Srs& s);
int msm_table_config(size_t srs_n, int* c, int* nwin);
// d_ops: records in recording order; counts[opc] = how many carry opcode opc
int witness_expand_dev(h2agg_ctx* ctx, const void* d_ops, const size_t* counts, void* const d_cols[5], size_t n_rows);
int batch_invert_dev(h2agg_ctx* ctx, void* d_a, size_t n);
// n_cols MSMs against the same bases, alternating lanes; joins back into ctx->stream.
// Columns are device pointers, or host pointers when `host_cols` (then staged through the lanes).
// `win_begins` / `win_ends` (optional, n_cols entries): a window range per column instead of one for the batch.
int msm_run_batch(h2agg_ctx* ctx, const MsmBases& bases, const void* const* cols, size_t n_cols, size_t n,
                  uint8_t* d_out160s, bool host_cols, int win_begin = 0, int win_end = -1, const int* win_begins = nullptr,
                  const int* win_ends = nullptr);
int lanes_init(h2agg_ctx* ctx);
// sum of m affine-or-jacobian(96 B) points -> d_out160
int g1_sum_jacobian(h2agg_ctx* ctx, const void* d_points96, size_t m, void* d_out160, size_t stride = 96, size_t n_out = 1);
int msm_window_config(size_t n, int forced_c, int* c, int* nwin);

}  // namespace h2agg
