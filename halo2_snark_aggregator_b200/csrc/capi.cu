// C ABI (include/h2agg.h) over the CUDA path.  No CPU fallback lives here: every entry point
// either runs the sm_100a kernels or returns an error.
#include "../../include/h2agg.h"
#include "bn254_g1.cuh"
#include "ctx.hpp"
#include <cstring>
#include <new>

using namespace h2agg;

static thread_local std::string g_init_error;

namespace h2agg {
std::atomic<int> g_any_device{-1};

int bg_stream_get(h2agg_ctx* ctx, cudaStream_t* out) {
  if (!ctx->bg_stream) {
    int least = 0, greatest = 0;
    H2AGG_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
    H2AGG_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->bg_stream, cudaStreamNonBlocking, least));
    H2AGG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->bg_ev, cudaEventDisableTiming));
  }
  *out = ctx->bg_stream;
  return 0;
}

int bg_join(h2agg_ctx* ctx) {
  if (!ctx->bg_pending || !ctx->bg_stream) return 0;
  H2AGG_CUDA(ctx, cudaEventRecord(ctx->bg_ev, ctx->bg_stream));
  H2AGG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->bg_ev, 0));
  ctx->bg_pending = false;
  return 0;
}
}  // namespace h2agg

#define LOCK(ctx) std::lock_guard<std::recursive_mutex> lock_((ctx)->mu)
#define CHECK_ARG(ctx, cond, msg) \
  do {                            \
    if (!(cond)) {                \
      (ctx)->last_error = (msg);  \
      return 1;                   \
    }                             \
  } while (0)

extern "C" {

const char* h2agg_version(void) { return "h2agg-b200 0.1 (sm_100a)"; }

int h2agg_init(int device_id, h2agg_ctx** out) {
  if (!out) return 1;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_init_error = std::string("h2agg_init: no CUDA device (") + cudaGetErrorString(e) + ")";
    cudaGetLastError();
    return 3;
  }
  if (device_id < 0 || device_id >= ndev) {
    g_init_error = "h2agg_init: device id out of range";
    return 1;
  }
  if ((e = cudaSetDevice(device_id)) != cudaSuccess) {
    g_init_error = std::string("h2agg_init: cudaSetDevice: ") + cudaGetErrorString(e);
    return 2;
  }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) {
    g_init_error = std::string("h2agg_init: ") + cudaGetErrorString(e);
    return 2;
  }
  if (prop.major != 10) {
    g_init_error = "h2agg_init: this library is built for sm_100a (Blackwell B200) only; found sm_" +
                   std::to_string(prop.major) + std::to_string(prop.minor);
    return 3;
  }
  h2agg_ctx* ctx = new (std::nothrow) h2agg_ctx();
  if (!ctx) return 2;
  ctx->device = device_id;
  ctx->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_init_error = std::string("h2agg_init: cudaStreamCreate: ") + cudaGetErrorString(e);
    delete ctx;
    return 2;
  }
  ctx->own_stream = true;
  ctx->pinned_cap = 1 << 16;
  if ((e = cudaMallocHost(&ctx->pinned, ctx->pinned_cap)) != cudaSuccess) {
    g_init_error = std::string("h2agg_init: cudaMallocHost: ") + cudaGetErrorString(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 2;
  }
  if (ensure(ctx, ctx->small, 1 << 16)) {
    g_init_error = ctx->last_error;
    h2agg_destroy(ctx);
    return 2;
  }
  int none = -1;
  h2agg::g_any_device.compare_exchange_strong(none, device_id);
  *out = ctx;
  return 0;
}

void h2agg_destroy(h2agg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& t : ctx->tw) {
    cudaFree(t.lo);
    cudaFree(t.hi);
    cudaFree(t.full);
  }
  for (auto& kv : ctx->srs) {
    if (kv.second.owned) cudaFree(const_cast<void*>(kv.second.d_bases));
    if (kv.second.d_table) cudaFree(kv.second.d_table);
  }
  cudaFree(ctx->ntt_tmp.p);
  cudaFree(ctx->io_a.p);
  cudaFree(ctx->wit_ws.p);
  if (ctx->bg_stream) {
    cudaStreamSynchronize(ctx->bg_stream);
    cudaStreamDestroy(ctx->bg_stream);
    cudaEventDestroy(ctx->bg_ev);
  }
  cudaFree(ctx->bg_ntt_tmp.p);
  cudaFree(ctx->io_b.p);
  cudaFree(ctx->msm_ws.p);
  for (int i = 0; i < N_LANES; i++) {
    if (ctx->lanes[i].st) cudaStreamSynchronize(ctx->lanes[i].st);
    cudaFree(ctx->lanes[i].ws.p);
    cudaFree(ctx->lanes[i].io.p);
    cudaFree(ctx->lanes[i].io_out.p);
    cudaFree(ctx->lanes[i].ntt_tmp.p);
    cudaFree(ctx->lanes[i].scan_ws.p);
    cudaFree(ctx->lanes[i].args_ws.p);
    if (ctx->lanes[i].done) cudaEventDestroy(ctx->lanes[i].done);
    if (ctx->lanes[i].up) cudaEventDestroy(ctx->lanes[i].up);
    if (ctx->lanes[i].st) cudaStreamDestroy(ctx->lanes[i].st);
  }
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  for (auto& t : ctx->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  cudaFree(ctx->small.p);
  cudaFree(ctx->poly_ws.p);
  cudaFree(ctx->poly_many_ws.p);
  cudaFree(ctx->scan_ws.p);
  cudaFree(ctx->sort_ws.p);
  cudaFree(ctx->args_ws.p);
  cudaFree(ctx->args_meta.p);
  cudaFree(ctx->quot_ws.p);
  cudaFree(ctx->quot_tw.p);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* h2agg_last_error(h2agg_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_init_error.c_str(); }

int h2agg_set_stream(h2agg_ctx* ctx, void* cuda_stream) {
  if (!ctx) return 1;
  LOCK(ctx);
  // No synchronisation: like any CUDA stream switch, ordering between the old and the new stream is the
  // caller's business (events).  The context's own stream is drained once before it is dropped.
  if (ctx->own_stream) {
    H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
  }
  ctx->stream = (cudaStream_t)cuda_stream;
  return 0;
}

int h2agg_synchronize(h2agg_ctx* ctx) {
  if (!ctx) return 1;
  LOCK(ctx);
  int rc = bg_join(ctx);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int h2agg_set_defer_transforms(h2agg_ctx* ctx, int enable) {
  if (!ctx) return 1;
  LOCK(ctx);
  ctx->defer_transforms = enable != 0;
  return 0;
}

int h2agg_transforms_join(h2agg_ctx* ctx) {
  if (!ctx) return 1;
  LOCK(ctx);
  return bg_join(ctx);
}

uint64_t h2agg_launch_count(h2agg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int h2agg_kernel_timing(h2agg_ctx* ctx, int enable) {
  if (!ctx) return 1;
  LOCK(ctx);
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& t : ctx->timed) { ctx->ev_pool.push_back(t.a); ctx->ev_pool.push_back(t.b); }
  ctx->timed.clear();
  ctx->timing = enable != 0;
  return 0;
}

int h2agg_kernel_times(h2agg_ctx* ctx, double* ms_per_class, uint64_t* count_per_class, int n_classes) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, ms_per_class && count_per_class && n_classes >= KC_COUNT, "kernel_times: need >= 7 classes");
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n_classes; i++) { ms_per_class[i] = 0; count_per_class[i] = 0; }
  for (auto& t : ctx->timed) {
    float ms = 0;
    H2AGG_CUDA(ctx, cudaEventElapsedTime(&ms, t.a, t.b));
    ms_per_class[t.cls] += ms;
    count_per_class[t.cls]++;
    ctx->ev_pool.push_back(t.a);
    ctx->ev_pool.push_back(t.b);
  }
  ctx->timed.clear();
  return 0;
}

int h2agg_host_register(h2agg_ctx* ctx, const void* p, size_t bytes) {
  if (!ctx) return 1;
  LOCK(ctx);
  H2AGG_CUDA(ctx, cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault));
  return 0;
}
int h2agg_host_unregister(h2agg_ctx* ctx, const void* p) {
  if (!ctx) return 1;
  LOCK(ctx);
  H2AGG_CUDA(ctx, cudaHostUnregister(const_cast<void*>(p)));
  return 0;
}

int h2agg_set_msm_window(h2agg_ctx* ctx, int c_bits) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, c_bits == 0 || (c_bits >= 2 && c_bits <= 20), "msm window must be 0 or in [2, 20]");
  ctx->msm_window_bits = c_bits;
  return 0;
}

int h2agg_set_ntt_radix_cap(h2agg_ctx* ctx, int log2_radix) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, log2_radix >= 2 && log2_radix <= 11, "ntt radix cap must be in [2, 11]");
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& t : ctx->tw) {  // the table split follows the plan: drop cached tables
    cudaFree(t.lo);
    cudaFree(t.hi);
    cudaFree(t.full);
  }
  ctx->tw.clear();
  ctx->ntt_radix_cap = (uint32_t)log2_radix;
  return 0;
}

int h2agg_set_srs_precompute(h2agg_ctx* ctx, int enable) {
  if (!ctx) return 1;
  LOCK(ctx);
  ctx->srs_precompute = enable != 0;
  return 0;
}

int h2agg_srs_config(h2agg_ctx* ctx, uint64_t srs_id, int* table_mode, int* c_bits, int* n_windows) {
  if (!ctx || !table_mode || !c_bits || !n_windows) return 1;
  LOCK(ctx);
  auto it = ctx->srs.find(srs_id);
  CHECK_ARG(ctx, it != ctx->srs.end(), "srs_config: unknown srs id");
  const bool table = it->second.d_table != nullptr && ctx->msm_window_bits == 0;
  *table_mode = table ? 1 : 0;
  if (table) {
    *c_bits = it->second.table_c;
    *n_windows = it->second.table_nwin;
    return 0;
  }
  return msm_window_config(it->second.n, ctx->msm_window_bits, c_bits, n_windows);
}

int h2agg_msm_config(h2agg_ctx* ctx, size_t n, int* c_bits, int* n_windows) {
  if (!ctx || !c_bits || !n_windows) return 1;
  LOCK(ctx);
  return msm_window_config(n ? n : 1, ctx->msm_window_bits, c_bits, n_windows);
}

// ---- SRS ----------------------------------------------------------------------------------------
int h2agg_srs_register(h2agg_ctx* ctx, const uint64_t* bases, size_t n, uint64_t* out_id) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, bases && out_id && n > 0, "srs_register: null argument or n == 0");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  void* d = nullptr;
  H2AGG_CUDA(ctx, cudaMalloc(&d, n * 64));
  cudaError_t e = cudaMemcpyAsync(d, bases, n * 64, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    cudaFree(d);
    ctx->last_error = std::string("srs_register: ") + cudaGetErrorString(e);
    return 2;
  }
  Srs s;
  s.d_bases = d;
  s.n = n;
  s.owned = true;
  if (ctx->srs_precompute) {
    int rc = msm_build_srs_table(ctx, s);
    if (rc) { cudaFree(d); return rc; }
  }
  *out_id = ctx->next_srs++;
  ctx->srs[*out_id] = s;
  return 0;
}

int h2agg_srs_register_dev(h2agg_ctx* ctx, const void* d_bases, size_t n, uint64_t* out_id) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_bases && out_id && n > 0, "srs_register_dev: null argument or n == 0");
  Srs s;
  s.d_bases = d_bases;
  s.n = n;
  s.owned = false;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->srs_precompute) {
    int rc = msm_build_srs_table(ctx, s);
    if (rc) return rc;
  }
  *out_id = ctx->next_srs++;
  ctx->srs[*out_id] = s;
  return 0;
}

int h2agg_srs_release(h2agg_ctx* ctx, uint64_t id) {
  if (!ctx) return 1;
  LOCK(ctx);
  auto it = ctx->srs.find(id);
  CHECK_ARG(ctx, it != ctx->srs.end(), "srs_release: unknown srs id");
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (it->second.owned) cudaFree(const_cast<void*>(it->second.d_bases));
  if (it->second.d_table) cudaFree(it->second.d_table);
  ctx->srs.erase(it);
  return 0;
}

// resolve what an MSM of n pairs runs against (uploads host bases to io_b if needed)
static int resolve_bases(h2agg_ctx* ctx, uint64_t srs_id, const void* host_bases, const void* dev_bases, size_t n,
                         MsmBases* out) {
  *out = MsmBases();
  if (srs_id) {
    auto it = ctx->srs.find(srs_id);
    CHECK_ARG(ctx, it != ctx->srs.end(), "msm: unknown srs id");
    CHECK_ARG(ctx, it->second.n >= n, "msm: more scalars than registered bases");
    out->d_bases = it->second.d_bases;
    out->d_table = it->second.d_table;
    out->table_c = it->second.table_c;
    out->table_nwin = it->second.table_nwin;
    out->srs_n = it->second.n;
    return 0;
  }
  out->srs_n = n;
  if (dev_bases) {
    out->d_bases = dev_bases;
    return 0;
  }
  CHECK_ARG(ctx, host_bases || n == 0, "msm: no bases given (srs_id == 0 and bases == NULL)");
  int rc = ensure(ctx, ctx->io_b, n * 64 + 64);
  if (rc) return rc;
  if (n) H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, host_bases, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  out->d_bases = ctx->io_b.p;
  return 0;
}

// ---- MSM ----------------------------------------------------------------------------------------
int h2agg_msm_g1_windows(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* bases, const uint64_t* scalars, size_t n,
                         int win_begin, int win_end, uint64_t out_jacobian[12]) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, out_jacobian, "msm: out is NULL");
  CHECK_ARG(ctx, scalars || n == 0, "msm: scalars is NULL");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  MsmBases d_bases;
  int rc = resolve_bases(ctx, srs_id, bases, nullptr, n, &d_bases);
  if (rc) return rc;
  rc = ensure(ctx, ctx->io_a, n * 32 + 64);
  if (rc) return rc;
  if (n) H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  rc = msm_run(ctx, ctx->stream, ctx->msm_ws, d_bases, ctx->io_a.p, n, ctx->small.p, win_begin, win_end);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->small.p, 160, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out_jacobian, (uint8_t*)ctx->pinned + 64, 96);
  return 0;
}

int h2agg_msm_g1(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* bases, const uint64_t* scalars, size_t n,
                 uint64_t out_jacobian[12]) {
  return h2agg_msm_g1_windows(ctx, srs_id, bases, scalars, n, 0, -1, out_jacobian);
}

int h2agg_msm_g1_windows_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases_in, const void* d_scalars, size_t n,
                             int win_begin, int win_end, void* d_out160) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_out160 && (d_scalars || n == 0), "msm_dev: null argument");
  CHECK_ARG(ctx, srs_id || d_bases_in || n == 0, "msm_dev: no bases");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  MsmBases d_bases;
  int rc = resolve_bases(ctx, srs_id, nullptr, d_bases_in, n, &d_bases);
  if (rc) return rc;
  return msm_run(ctx, ctx->stream, ctx->msm_ws, d_bases, d_scalars, n, d_out160, win_begin, win_end);
}

int h2agg_msm_g1_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases, const void* d_scalars, size_t n,
                     void* d_out160) {
  return h2agg_msm_g1_windows_dev(ctx, srs_id, d_bases, d_scalars, n, 0, -1, d_out160);
}

int h2agg_msm_g1_batch(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* const* cols, size_t n_cols, size_t n,
                       uint64_t* out_affine) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, srs_id != 0, "msm_batch: needs a registered srs");
  CHECK_ARG(ctx, cols && out_affine, "msm_batch: null argument");
  CHECK_ARG(ctx, n_cols * 160 <= ctx->small.cap && n_cols * 160 <= ctx->pinned_cap, "msm_batch: too many columns");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  MsmBases d_bases;
  int rc = resolve_bases(ctx, srs_id, nullptr, nullptr, n, &d_bases);
  if (rc) return rc;
  // lanes: column i+1 crosses PCIe while column i is in the bucket kernels
  for (size_t i = 0; i < n_cols; i++) CHECK_ARG(ctx, cols[i], "msm_batch: null column");
  rc = msm_run_batch(ctx, d_bases, (const void* const*)cols, n_cols, n, (uint8_t*)ctx->small.p, true);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->small.p, n_cols * 160, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < n_cols; i++) memcpy(out_affine + i * 8, (uint8_t*)ctx->pinned + i * 160, 64);
  return 0;
}

int h2agg_msm_g1_batch_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases_in, const void* const* d_cols,
                           size_t n_cols, size_t n, void* d_out160s) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_cols && d_out160s, "msm_batch_dev: null argument");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  MsmBases d_bases;
  int rc = resolve_bases(ctx, srs_id, nullptr, d_bases_in, n, &d_bases);
  if (rc) return rc;
  return msm_run_batch(ctx, d_bases, d_cols, n_cols, n, (uint8_t*)d_out160s, false);
}

int h2agg_msm_g1_batch_windows_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases_in, const void* const* d_cols,
                                   size_t n_cols, size_t n, int win_begin, int win_end, void* d_out160s) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_cols && d_out160s, "msm_batch_windows_dev: null argument");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  MsmBases d_bases;
  int rc = resolve_bases(ctx, srs_id, nullptr, d_bases_in, n, &d_bases);
  if (rc) return rc;
  return msm_run_batch(ctx, d_bases, d_cols, n_cols, n, (uint8_t*)d_out160s, false, win_begin, win_end);
}

int h2agg_msm_g1_batch_ranges_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* d_bases_in, const void* const* d_cols,
                                  size_t n_cols, size_t n, const int* win_begins, const int* win_ends, void* d_out160s) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_cols && d_out160s && win_begins && win_ends, "msm_batch_ranges_dev: null argument");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  MsmBases d_bases;
  int rc = resolve_bases(ctx, srs_id, nullptr, d_bases_in, n, &d_bases);
  if (rc) return rc;
  return msm_run_batch(ctx, d_bases, d_cols, n_cols, n, (uint8_t*)d_out160s, false, 0, -1, win_begins, win_ends);
}

int h2agg_g1_sum_dev(h2agg_ctx* ctx, const void* d_points, size_t m, size_t stride_bytes, size_t n_out, void* d_out160s) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_points && d_out160s && m > 0 && n_out > 0, "g1_sum_dev: bad argument");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  return g1_sum_jacobian(ctx, d_points, m, d_out160s, stride_bytes, n_out);
}

int h2agg_g1_sum(h2agg_ctx* ctx, const uint64_t* pts, size_t m, uint64_t out_jacobian[12]) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, pts && out_jacobian && m > 0 && m * 96 + 256 <= ctx->small.cap, "g1_sum: bad argument");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  uint8_t* d_in = (uint8_t*)ctx->small.p + 256;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(d_in, pts, m * 96, cudaMemcpyHostToDevice, ctx->stream));
  int rc = g1_sum_jacobian(ctx, d_in, m, ctx->small.p);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->small.p, 160, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out_jacobian, (uint8_t*)ctx->pinned + 64, 96);
  return 0;
}

// ---- NTT ----------------------------------------------------------------------------------------
static int ntt_host(h2agg_ctx* ctx, const uint64_t* src, size_t src_n, uint64_t* dst, size_t dst_n, NttOpts o) {
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t N = (size_t)1 << o.log_n;
  int rc = ensure(ctx, ctx->io_a, N * 32);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, src, src_n * 32, cudaMemcpyHostToDevice, ctx->stream));
  o.src_n = src_n;
  o.dst_n = dst_n;
  rc = ntt_run(ctx, ctx->io_a.p, ctx->io_a.p, o);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(dst, ctx->io_a.p, dst_n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int h2agg_ntt_fr(h2agg_ctx* ctx, uint64_t* a, const uint64_t omega[4], uint32_t log_n) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, a && omega, "ntt: null argument");
  CHECK_ARG(ctx, log_n <= 28, "ntt: log_n > 28");
  NttOpts o{omega, log_n, 0, 0, nullptr, nullptr};
  size_t N = (size_t)1 << log_n;
  return ntt_host(ctx, a, N, a, N, o);
}

int h2agg_intt_fr(h2agg_ctx* ctx, uint64_t* a, const uint64_t omega_inv[4], const uint64_t n_inv[4], uint32_t log_n) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, a && omega_inv && n_inv, "intt: null argument");
  CHECK_ARG(ctx, log_n >= 1 && log_n <= 28, "intt: log_n out of range");
  uint64_t s3[12];
  for (int i = 0; i < 3; i++) memcpy(s3 + 4 * i, n_inv, 32);
  NttOpts o{omega_inv, log_n, 0, 0, nullptr, s3};
  size_t N = (size_t)1 << log_n;
  return ntt_host(ctx, a, N, a, N, o);
}

int h2agg_ntt_fr_dev(h2agg_ctx* ctx, void* d_a, const uint64_t omega[4], const uint64_t* scale, uint32_t log_n) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_a && omega, "ntt_dev: null argument");
  CHECK_ARG(ctx, log_n <= 28 && (log_n >= 1 || !scale), "ntt_dev: log_n out of range");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  uint64_t s3[12];
  if (scale)
    for (int i = 0; i < 3; i++) memcpy(s3 + 4 * i, scale, 32);
  size_t N = (size_t)1 << log_n;
  NttOpts o{omega, log_n, N, N, nullptr, scale ? s3 : nullptr};
  return ntt_run(ctx, d_a, d_a, o);
}

// Batched host-pointer transforms: column i runs on lane i % N_LANES, so the H2D copy of column i+1,
// the passes of column i and the D2H copy of column i-1 overlap (PCIe is full duplex).
// in_n / out_n elements per column; src/dst host pointers (dst may equal src).
static int ntt_host_batch(h2agg_ctx* ctx, const uint64_t* const* src, uint64_t* const* dst, size_t n_cols, size_t in_n,
                          size_t out_n, NttOpts o) {
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = lanes_init(ctx);
  if (rc) return rc;
  const size_t N = (size_t)1 << o.log_n;
  o.src_n = in_n;
  o.dst_n = out_n;
  // the (cached) twiddle tables are generated on the main stream: before the lanes fork off it
  if ((rc = ntt_warm_tables(ctx, o.omega, o.log_n))) return rc;
  LaneFork lf(ctx);
  if ((rc = lf.fork())) return rc;
  for (size_t i = 0; i < n_cols; i++) {
    Lane& ln = ctx->lanes[i % N_LANES];
    if ((rc = ensure(ctx, ln.io, in_n * 32))) return rc;
    if ((rc = ensure(ctx, ln.io_out, N * 32))) return rc;
    H2AGG_CUDA(ctx, cudaMemcpyAsync(ln.io.p, src[i], in_n * 32, cudaMemcpyHostToDevice, ln.st));
    if ((rc = ntt_run(ctx, ln.io.p, ln.io_out.p, o, ln.st, &ln.ntt_tmp))) return rc;
    H2AGG_CUDA(ctx, cudaMemcpyAsync(dst[i], ln.io_out.p, out_n * 32, cudaMemcpyDeviceToHost, ln.st));
  }
  if ((rc = lf.join())) return rc;
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int h2agg_intt_fr_batch(h2agg_ctx* ctx, uint64_t* const* cols, size_t n_cols, const uint64_t omega_inv[4],
                        const uint64_t n_inv[4], uint32_t log_n) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, cols && omega_inv && n_inv, "intt_batch: null argument");
  CHECK_ARG(ctx, log_n >= 1 && log_n <= 28, "intt_batch: log_n out of range");
  uint64_t s3[12];
  for (int i = 0; i < 3; i++) memcpy(s3 + 4 * i, n_inv, 32);
  NttOpts o{omega_inv, log_n, 0, 0, nullptr, s3};
  size_t N = (size_t)1 << log_n;
  return ntt_host_batch(ctx, (const uint64_t* const*)cols, cols, n_cols, N, N, o);
}

// Montgomery products of a handful of field elements on the device (zeta powers etc.)
__global__ void coset_consts_kernel(Fr zeta, Fr scale, int inverse, Fr* out3) {
  if (threadIdx.x || blockIdx.x) return;
  Fr z2 = zeta * zeta;
  if (!inverse) {
    Fr::one().store(out3);
    zeta.store(out3 + 1);
    z2.store(out3 + 2);
  } else {  // scale * zeta^-(i mod 3); zeta^-1 = zeta^2
    scale.store(out3);
    (scale * z2).store(out3 + 1);
    (scale * zeta).store(out3 + 2);
  }
}

static int coset_consts(h2agg_ctx* ctx, const uint64_t zeta[4], const uint64_t* scale, int inverse, uint64_t out3[12]) {
  Fr z, s;
  memcpy(z.v, zeta, 32);
  if (scale) memcpy(s.v, scale, 32); else memset(s.v, 0, 32);
  Fr* d = (Fr*)((uint8_t*)ctx->small.p + 4096);
  coset_consts_kernel<<<1, 32, 0, ctx->stream>>>(z, s, inverse, d);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  H2AGG_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ctx->pinned + 4096, d, 96, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out3, (uint8_t*)ctx->pinned + 4096, 96);
  return 0;
}

int h2agg_coeff_to_extended_dev(h2agg_ctx* ctx, const void* d_coeffs, uint32_t k, uint32_t ext_k,
                                const uint64_t zeta[4], const uint64_t omega_ext[4], void* d_out) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_coeffs && d_out && zeta && omega_ext, "coeff_to_extended: null argument");
  CHECK_ARG(ctx, ext_k >= k && ext_k >= 1 && ext_k <= 28, "coeff_to_extended: bad k / ext_k");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  uint64_t in3[12];
  int rc = coset_consts(ctx, zeta, nullptr, 0, in3);
  if (rc) return rc;
  NttOpts o{omega_ext, ext_k, (size_t)1 << k, (size_t)1 << ext_k, in3, nullptr};
  return ntt_run(ctx, d_coeffs, d_out, o);
}

int h2agg_coeff_to_extended(h2agg_ctx* ctx, const uint64_t* coeffs, uint32_t k, uint32_t ext_k, const uint64_t zeta[4],
                            const uint64_t omega_ext[4], uint64_t* out) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, coeffs && out && zeta && omega_ext, "coeff_to_extended: null argument");
  CHECK_ARG(ctx, ext_k >= k && ext_k >= 1 && ext_k <= 28, "coeff_to_extended: bad k / ext_k");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  uint64_t in3[12];
  int rc = coset_consts(ctx, zeta, nullptr, 0, in3);
  if (rc) return rc;
  NttOpts o{omega_ext, ext_k, 0, 0, in3, nullptr};
  return ntt_host(ctx, coeffs, (size_t)1 << k, out, (size_t)1 << ext_k, o);
}

int h2agg_coeff_to_extended_batch(h2agg_ctx* ctx, const uint64_t* const* coeff_cols, uint64_t* const* out_cols,
                                  size_t n_cols, uint32_t k, uint32_t ext_k, const uint64_t zeta[4],
                                  const uint64_t omega_ext[4]) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, coeff_cols && out_cols && zeta && omega_ext, "coeff_to_extended_batch: null argument");
  CHECK_ARG(ctx, ext_k >= k && ext_k >= 1 && ext_k <= 28, "coeff_to_extended_batch: bad k / ext_k");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  uint64_t in3[12];
  int rc = coset_consts(ctx, zeta, nullptr, 0, in3);
  if (rc) return rc;
  NttOpts o{omega_ext, ext_k, 0, 0, in3, nullptr};
  return ntt_host_batch(ctx, coeff_cols, out_cols, n_cols, (size_t)1 << k, (size_t)1 << ext_k, o);
}

// One commit round fused per column (see h2agg.h): upload once; MSM, iNTT and (optionally) the coset NTT run
// back to back on the column's lane; the D2H copies of lane A overlap the kernels of lane B and the upload of lane C.
// `resident`: coeff_out / ext_out are caller-owned DEVICE buffers the transforms write straight into (no D2H);
// otherwise host buffers filled through the lane's staging buffers.
static int commit_round_impl(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* const* lagrange_cols, size_t n_cols, uint32_t k,
                             const uint64_t omega_inv[4], const uint64_t n_inv[4], uint64_t* out_affine,
                             uint64_t* const* coeff_out, uint32_t ext_k, const uint64_t zeta[4], const uint64_t omega_ext[4],
                             uint64_t* const* ext_out, bool resident, void* const* d_lagrange_keep = nullptr,
                             bool src_on_device = false) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, srs_id != 0 && lagrange_cols && out_affine && omega_inv && n_inv, "commit_round: null argument");
  CHECK_ARG(ctx, k >= 1 && k <= 28, "commit_round: k out of range");
  CHECK_ARG(ctx, !ext_out || (zeta && omega_ext && ext_k >= k && ext_k <= 28 && coeff_out), "commit_round: extended output needs zeta, omega_ext and coefficients");
  CHECK_ARG(ctx, n_cols * 160 <= ctx->small.cap && n_cols * 160 <= ctx->pinned_cap, "commit_round: too many columns");
  if (n_cols == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << k;
  MsmBases bases;
  int rc = resolve_bases(ctx, srs_id, nullptr, nullptr, n, &bases);
  if (rc) return rc;
  if ((rc = lanes_init(ctx))) return rc;
  uint64_t s3[12], in3[12];
  for (int i = 0; i < 3; i++) memcpy(s3 + 4 * i, n_inv, 32);
  if (ext_out && (rc = coset_consts(ctx, zeta, nullptr, 0, in3))) return rc;
  NttOpts oi{omega_inv, k, n, n, nullptr, s3};
  NttOpts oe{omega_ext, ext_k, n, (size_t)1 << ext_k, in3, nullptr};
  // the (cached) twiddle tables are generated on the main stream: make sure both exist before the lanes fork off
  // it, whichever columns ask for transforms (coeff_out / ext_out may hold NULL entries anywhere)
  if (coeff_out && (rc = ntt_warm_tables(ctx, omega_inv, k))) return rc;
  if (ext_out && (rc = ntt_warm_tables(ctx, omega_ext, ext_k))) return rc;
  LaneFork lf(ctx);
  if ((rc = lf.fork())) return rc;
  // Few columns (a rank's share of a round on a multi-GPU prover): the MSM and the transforms of a column only share
  // their INPUT, so with device-resident outputs they run on two lanes -- the ~20-kernel latency chain of the MSM beside
  // the pipe-bound NTT passes instead of in front of them.
  // Deferred transforms: the commitments gate the next Fiat-Shamir challenge, the coefficient / extended forms are only
  // needed in the quotient and evaluation stages -- with h2agg_set_defer_transforms the NTT passes go to the low-priority
  // background stream (after the column is in HBM) and this call returns as soon as the MSMs are done; they then fill the
  // latency-bound stretches of the following rounds.  h2agg_transforms_join orders the main stream after them.
  const bool defer = resident && coeff_out && ctx->defer_transforms;
  cudaStream_t bg = nullptr;
  if (defer) {
    if ((rc = bg_stream_get(ctx, &bg))) return rc;
    H2AGG_CUDA(ctx, cudaStreamWaitEvent(bg, ctx->fork_ev, 0));   // everything the main stream holds so far
    ctx->bg_pending = true;
  }
  const bool split = !defer && resident && coeff_out && 2 * n_cols <= (size_t)N_LANES;
  for (size_t i = 0; i < n_cols; i++) {
    Lane& ln = ctx->lanes[i % N_LANES];
    CHECK_ARG(ctx, lagrange_cols[i], "commit_round: null column");
    if (!src_on_device && !(d_lagrange_keep && d_lagrange_keep[i]) && (rc = ensure(ctx, ln.io, n * 32 + 64))) return rc;
    if (!resident && ext_out && ext_out[i] && (rc = ensure(ctx, ln.io_out, ((size_t)1 << ext_k) * 32))) return rc;
    // the Lagrange column: already in HBM, uploaded into the caller's resident buffer, or staged in the lane
    void* col = ln.io.p;
    if (src_on_device) {
      col = const_cast<uint64_t*>(lagrange_cols[i]);
    } else {
      if (d_lagrange_keep && d_lagrange_keep[i]) col = d_lagrange_keep[i];
      H2AGG_CUDA(ctx, cudaMemcpyAsync(col, lagrange_cols[i], n * 32, cudaMemcpyHostToDevice, ln.st));
    }
    // where the transforms of column i run (an in-place lagrange_to_coeff must stay behind the MSM that reads the column)
    Lane& ln_t = (split && coeff_out[i] && (const void*)coeff_out[i] != (const void*)col) ? ctx->lanes[n_cols + i] : ln;
    if (&ln_t != &ln) {
      H2AGG_CUDA(ctx, cudaEventRecord(ln.up, ln.st));
      H2AGG_CUDA(ctx, cudaStreamWaitEvent(ln_t.st, ln.up, 0));
    }
    if ((rc = msm_run(ctx, ln.st, ln.ws, bases, col, n, (uint8_t*)ctx->small.p + i * 160, 0, -1, false))) return rc;
    if (coeff_out && coeff_out[i]) {
      cudaStream_t st = ln_t.st;
      if (defer) {
        H2AGG_CUDA(ctx, cudaEventRecord(ln.up, ln.st));           // the column is in HBM (lanes reuse `up`: stream order keeps it safe)
        H2AGG_CUDA(ctx, cudaStreamWaitEvent(bg, ln.up, 0));
        if ((rc = ntt_run(ctx, col, coeff_out[i], oi, bg, &ctx->bg_ntt_tmp))) return rc;
        if (ext_out && ext_out[i] && (rc = ntt_run(ctx, coeff_out[i], ext_out[i], oe, bg, &ctx->bg_ntt_tmp))) return rc;
      } else if (resident) {
        if ((rc = ntt_run(ctx, col, coeff_out[i], oi, st, &ln_t.ntt_tmp))) return rc;
        if (ext_out && ext_out[i] && (rc = ntt_run(ctx, coeff_out[i], ext_out[i], oe, st, &ln_t.ntt_tmp))) return rc;
      } else {
        if ((rc = ntt_run(ctx, ln.io.p, ln.io.p, oi, st, &ln.ntt_tmp))) return rc;
        H2AGG_CUDA(ctx, cudaMemcpyAsync(coeff_out[i], ln.io.p, n * 32, cudaMemcpyDeviceToHost, st));
        if (ext_out && ext_out[i]) {
          if ((rc = ntt_run(ctx, ln.io.p, ln.io_out.p, oe, st, &ln.ntt_tmp))) return rc;
          H2AGG_CUDA(ctx, cudaMemcpyAsync(ext_out[i], ln.io_out.p, ((size_t)32) << ext_k, cudaMemcpyDeviceToHost, st));
        }
      }
    }
  }
  if ((rc = lf.join())) return rc;
  if ((rc = g1_normalize(ctx, ctx->stream, ctx->small.p, n_cols))) return rc;   // ONE inversion for the whole round
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->small.p, n_cols * 160, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < n_cols; i++) memcpy(out_affine + i * 8, (uint8_t*)ctx->pinned + i * 160, 64);
  return 0;
}

int h2agg_commit_round(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* const* lagrange_cols, size_t n_cols, uint32_t k,
                       const uint64_t omega_inv[4], const uint64_t n_inv[4], uint64_t* out_affine,
                       uint64_t* const* coeff_out, uint32_t ext_k, const uint64_t zeta[4], const uint64_t omega_ext[4],
                       uint64_t* const* ext_out) {
  return commit_round_impl(ctx, srs_id, lagrange_cols, n_cols, k, omega_inv, n_inv, out_affine, coeff_out, ext_k, zeta,
                           omega_ext, ext_out, false);
}

int h2agg_commit_round_resident(h2agg_ctx* ctx, uint64_t srs_id, const uint64_t* const* lagrange_cols, size_t n_cols,
                                uint32_t k, const uint64_t omega_inv[4], const uint64_t n_inv[4], uint64_t* out_affine,
                                void* const* d_lagrange_out, void* const* d_coeff_out, uint32_t ext_k,
                                const uint64_t zeta[4], const uint64_t omega_ext[4], void* const* d_ext_out) {
  if (ctx && !d_coeff_out) {
    LOCK(ctx);
    ctx->last_error = "commit_round_resident: d_coeff_out is NULL (use h2agg_msm_g1_batch for commitments only)";
    return 1;
  }
  return commit_round_impl(ctx, srs_id, lagrange_cols, n_cols, k, omega_inv, n_inv, out_affine,
                           (uint64_t* const*)d_coeff_out, ext_k, zeta, omega_ext, (uint64_t* const*)d_ext_out, true,
                           d_lagrange_out, false);
}

int h2agg_commit_round_dev(h2agg_ctx* ctx, uint64_t srs_id, const void* const* d_lagrange_cols, size_t n_cols, uint32_t k,
                           const uint64_t omega_inv[4], const uint64_t n_inv[4], uint64_t* out_affine,
                           void* const* d_coeff_out, uint32_t ext_k, const uint64_t zeta[4], const uint64_t omega_ext[4],
                           void* const* d_ext_out) {
  if (ctx && !d_coeff_out) {
    LOCK(ctx);
    ctx->last_error = "commit_round_dev: d_coeff_out is NULL (use h2agg_msm_g1_batch_dev for commitments only)";
    return 1;
  }
  return commit_round_impl(ctx, srs_id, (const uint64_t* const*)d_lagrange_cols, n_cols, k, omega_inv, n_inv, out_affine,
                           (uint64_t* const*)d_coeff_out, ext_k, zeta, omega_ext, (uint64_t* const*)d_ext_out, true, nullptr,
                           true);
}

// The transform half of a commit round alone (columns whose commitment is computed elsewhere, e.g. window-sharded over
// other GPUs): lagrange_to_coeff (+ coeff_to_extended), out of place, on the background stream when transforms are
// deferred, else on the main stream.
int h2agg_transforms_dev(h2agg_ctx* ctx, const void* const* d_lagrange_cols, size_t n_cols, uint32_t k,
                         const uint64_t omega_inv[4], const uint64_t n_inv[4], void* const* d_coeff_out, uint32_t ext_k,
                         const uint64_t zeta[4], const uint64_t omega_ext[4], void* const* d_ext_out) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_lagrange_cols && d_coeff_out && omega_inv && n_inv, "transforms_dev: null argument");
  CHECK_ARG(ctx, k >= 1 && k <= 28, "transforms_dev: k out of range");
  CHECK_ARG(ctx, !d_ext_out || (zeta && omega_ext && ext_k >= k && ext_k <= 28), "transforms_dev: extended output needs zeta, omega_ext");
  if (n_cols == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << k;
  int rc;
  uint64_t s3[12], in3[12];
  for (int i = 0; i < 3; i++) memcpy(s3 + 4 * i, n_inv, 32);
  if (d_ext_out && (rc = coset_consts(ctx, zeta, nullptr, 0, in3))) return rc;
  NttOpts oi{omega_inv, k, n, n, nullptr, s3};
  NttOpts oe{omega_ext, ext_k, n, (size_t)1 << ext_k, in3, nullptr};
  if ((rc = ntt_warm_tables(ctx, omega_inv, k))) return rc;
  if (d_ext_out && (rc = ntt_warm_tables(ctx, omega_ext, ext_k))) return rc;
  cudaStream_t st = ctx->stream;
  DevBuf* tmp = &ctx->ntt_tmp;
  if (ctx->defer_transforms) {
    if ((rc = lanes_init(ctx))) return rc;   // (creates fork_ev)
    if ((rc = bg_stream_get(ctx, &st))) return rc;
    H2AGG_CUDA(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));
    H2AGG_CUDA(ctx, cudaStreamWaitEvent(st, ctx->fork_ev, 0));
    ctx->bg_pending = true;
    tmp = &ctx->bg_ntt_tmp;
  }
  for (size_t i = 0; i < n_cols; i++) {
    CHECK_ARG(ctx, d_lagrange_cols[i] && d_coeff_out[i], "transforms_dev: null column");
    if ((rc = ntt_run(ctx, d_lagrange_cols[i], d_coeff_out[i], oi, st, tmp))) return rc;
    if (d_ext_out && d_ext_out[i] && (rc = ntt_run(ctx, d_coeff_out[i], d_ext_out[i], oe, st, tmp))) return rc;
  }
  return 0;
}

int h2agg_extended_to_coeff_dev(h2agg_ctx* ctx, void* d_a, uint32_t ext_k, const uint64_t omega_ext_inv[4],
                                const uint64_t ext_n_inv[4], const uint64_t zeta[4], size_t out_len) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, d_a && omega_ext_inv && ext_n_inv && zeta, "extended_to_coeff: null argument");
  CHECK_ARG(ctx, ext_k >= 1 && ext_k <= 28 && out_len <= ((size_t)1 << ext_k), "extended_to_coeff: bad size");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  uint64_t out3[12];
  int rc = coset_consts(ctx, zeta, ext_n_inv, 1, out3);
  if (rc) return rc;
  NttOpts o{omega_ext_inv, ext_k, (size_t)1 << ext_k, out_len, nullptr, out3};
  return ntt_run(ctx, d_a, d_a, o);
}

int h2agg_extended_to_coeff(h2agg_ctx* ctx, uint64_t* a, uint32_t ext_k, const uint64_t omega_ext_inv[4],
                            const uint64_t ext_n_inv[4], const uint64_t zeta[4], size_t out_len) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, a && omega_ext_inv && ext_n_inv && zeta, "extended_to_coeff: null argument");
  CHECK_ARG(ctx, ext_k >= 1 && ext_k <= 28 && out_len <= ((size_t)1 << ext_k), "extended_to_coeff: bad size");
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  uint64_t out3[12];
  int rc = coset_consts(ctx, zeta, ext_n_inv, 1, out3);
  if (rc) return rc;
  NttOpts o{omega_ext_inv, ext_k, 0, 0, nullptr, out3};
  return ntt_host(ctx, a, (size_t)1 << ext_k, a, out_len, o);
}

}  // extern "C"

// ---- field helpers (tests) ----------------------------------------------------------------------
template <class F>
__global__ void field_op_kernel(int op, const F* a, const F* b, F* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = F::load(a + i);
  F r;
  if (op == 2) r = fp_inv(x);
  else {
    F y = F::load(b + i);
    r = (op == 0) ? x + y : (op == 1) ? x - y : x * y;
  }
  r.store(out + i);
}

// PrimeField::to_repr / from_repr in bulk: Montgomery limbs <-> 32-byte little-endian canonical integers.
// from_repr counts the inputs that are not < r (Rust returns None for those; the callers unwrap()).
__global__ void fr_repr_kernel(int from_repr, const Fr* in, Fr* out, size_t n, uint32_t* bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr x = Fr::load(in + i);
  if (!from_repr) {
    fp_from_mont(x).store(out + i);
    return;
  }
  bool ge = true;  // x >= r ?
#pragma unroll
  for (int j = 7; j >= 0; j--) {
    if (x.v[j] != FrTag::P(j)) { ge = x.v[j] > FrTag::P(j); break; }
  }
  if (ge) atomicAdd(bad, 1u);
  fp_to_mont(x).store(out + i);
}

extern "C" {

int h2agg_fr_repr(h2agg_ctx* ctx, int from_repr, const void* in, void* out, size_t n) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, (in && out) || n == 0, "fr_repr: null argument");
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure(ctx, ctx->io_a, n * 32);
  if (rc) return rc;
  uint32_t* bad = (uint32_t*)((uint8_t*)ctx->small.p + 8192);
  H2AGG_CUDA(ctx, cudaMemsetAsync(bad, 0, 4, ctx->stream));
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, in, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  fr_repr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(from_repr, (const Fr*)ctx->io_a.p, (Fr*)ctx->io_a.p, n, bad);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  H2AGG_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_a.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ctx->pinned + 8192, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint32_t nbad;
  memcpy(&nbad, (uint8_t*)ctx->pinned + 8192, 4);
  if (from_repr && nbad) {
    ctx->last_error = "fr_repr: " + std::to_string(nbad) + " value(s) are not canonical (>= r): PrimeField::from_repr returns None";
    return 4;
  }
  return 0;
}

int h2agg_field_op(h2agg_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  if (!ctx) return 1;
  LOCK(ctx);
  CHECK_ARG(ctx, a && out && (b || op == 2) && (field == 0 || field == 1) && op >= 0 && op <= 3, "field_op: bad argument");
  if (n == 0) return 0;
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure(ctx, ctx->io_a, n * 32);
  if (rc) return rc;
  rc = ensure(ctx, ctx->io_b, n * 32);
  if (rc) return rc;
  H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (b) H2AGG_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  unsigned grid = (unsigned)((n + 127) / 128);
  if (field == 0)
    field_op_kernel<Fr><<<grid, 128, 0, ctx->stream>>>(op, (const Fr*)ctx->io_a.p, (const Fr*)ctx->io_b.p, (Fr*)ctx->io_a.p, n);
  else
    field_op_kernel<Fq><<<grid, 128, 0, ctx->stream>>>(op, (const Fq*)ctx->io_a.p, (const Fq*)ctx->io_b.p, (Fq*)ctx->io_a.p, n);
  ctx->launches++;
  H2AGG_CUDA(ctx, cudaGetLastError());
  H2AGG_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_a.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int h2agg_field_mul(h2agg_ctx* ctx, int field, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  return h2agg_field_op(ctx, field, 3, a, b, out, n);
}

}  // extern "C"
