// Host side of the witness path: a *recording* implementation of the reference's chip surface.
//
// The reference already has three implementations of its plugin traits ArithCommonChip /
// ArithEccChip / ArithFieldChip (halo2-snark-aggregator-api/src/arith/{common,ecc,field}.rs): Mock,
// Circuit (halo2-snark-aggregator-circuit/src/chips/ecc_chip.rs:28-133) and Solidity (a recording
// context, halo2-snark-aggregator-solidity/src/lib.rs:194,293).  This is the fourth: it walks the
// same strictly sequential chip-op chain as EccChipOps / FiveColumnIntegerChip
// (halo2-ecc-circuit-lib/src/chips/ecc_chip.rs, five/integer_chip.rs), but instead of pushing
// ~150 cells per integer op through `dyn Region` it
//   * keeps only the skeleton needed to continue the chain: limb values (small integers), the
//     value mod p, `overflows`, the native/curvature caches and the row offset -- so the row
//     layout (including the state-dependent `reduce`s, :583-593) is reproduced exactly;
//   * appends one 256-byte record per row recipe; the B200 kernel (witness.cu) computes the
//     quotients, carries, range chunks, natives and inverses and writes the 5 advice columns.
// No 256-bit division, no inversion except the one `div` needs natively, happens on the host.
#include "../../include/h2agg.h"
#include "ctx.hpp"
#include "witness_ops.h"
#include "host_modinv.hpp"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;

namespace h2agg {
namespace wit {

// ---- host Fq (4 x 64 Montgomery) -------------------------------------------------------------
struct Fq {
  u64 v[4];
};
static const u64 QP[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 QR2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
static const u64 QINV = 0x87d20782e4866389ULL;

static inline bool q_geq(const u64* a) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > QP[i]) return true;
    if (a[i] < QP[i]) return false;
  }
  return true;
}
static inline void q_subp(u64* a) {
  u64 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - QP[i] - b;
    a[i] = (u64)d;
    b = (u64)(d >> 64) & 1;
  }
}
static inline Fq q_add(const Fq& a, const Fq& b) {
  Fq r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.v[i] + b.v[i];
    r.v[i] = (u64)c;
    c >>= 64;
  }
  if (q_geq(r.v)) q_subp(r.v);
  return r;
}
static inline Fq q_sub(const Fq& a, const Fq& b) {
  Fq r;
  u64 bo = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a.v[i] - b.v[i] - bo;
    r.v[i] = (u64)d;
    bo = (u64)(d >> 64) & 1;
  }
  if (bo) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)r.v[i] + QP[i];
      r.v[i] = (u64)c;
      c >>= 64;
    }
  }
  return r;
}
static inline Fq q_mul(const Fq& a, const Fq& b) {
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
    u64 m = t[0] * QINV;
    c = ((u128)m * QP[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * QP[j] + t[j];
      t[j - 1] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (u64)c;
    t[4] = t[5] + (u64)(c >> 64);
  }
  Fq r{{t[0], t[1], t[2], t[3]}};
  if (t[4] || q_geq(r.v)) q_subp(r.v);
  return r;
}
static inline Fq q_zero() { return Fq{{0, 0, 0, 0}}; }
static inline bool q_is_zero(const Fq& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
static inline Fq q_from_canon(const u64* c) {
  Fq a{{c[0], c[1], c[2], c[3]}};
  Fq r2{{QR2[0], QR2[1], QR2[2], QR2[3]}};
  return q_mul(a, r2);
}
static inline void q_to_canon(const Fq& a, u64* c) {
  Fq one{{1, 0, 0, 0}};
  Fq r = q_mul(a, one);
  memcpy(c, r.v, 32);
}
static inline Fq q_small(u64 x) {
  u64 c[4] = {x, 0, 0, 0};
  return q_from_canon(c);
}
// Inverse in Montgomery form (0 -> 0): the safegcd inverse (host_modinv.hpp) of the Montgomery residue aR is (aR)^-1;
// one Montgomery product with R^3 turns that into a^-1 R.  `div` needs one native inverse per curve operation.
static const u64 QR3[4] = {0xb1cd6dafda1530dfULL, 0x62f210e6a7283db6ULL, 0xef7f0b0c0ada0afbULL, 0x20fd6e902d592544ULL};
static const modinv::Modulus Q_MODULUS = modinv::make_modulus(QP);
static Fq q_inv(const Fq& a) {
  if (q_is_zero(a)) return q_zero();
  Fq x;
  modinv::inverse(a.v, Q_MODULUS, x.v);
  Fq r3{{QR3[0], QR3[1], QR3[2], QR3[3]}};
  return q_mul(x, r3);
}


// ---- host Fr on CANONICAL 256-bit values (native scalars of the ScalarChip, gates/base_gate.rs BaseGateOps) -----
// AssignedValue<Fr> handles carry canonical integers (that is what decompose_scalar and the RAW256 records consume);
// a product is two Montgomery steps: mont(mont(a, b), R^2) = a b.
namespace fr {
static const u64 RP[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 RR2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
static const u64 RINV = 0xc2e1f593efffffffULL;
struct V {
  u64 v[4];
};
static inline V zero() { return V{{0, 0, 0, 0}}; }
static inline V small(u64 x) { return V{{x, 0, 0, 0}}; }
static inline V from_u128(u128 x) { return V{{(u64)x, (u64)(x >> 64), 0, 0}}; }
static inline bool is_zero(const V& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
static inline bool geq_p(const u64* a) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > RP[i]) return true;
    if (a[i] < RP[i]) return false;
  }
  return true;
}
static inline void sub_p(u64* a) {
  u64 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - RP[i] - b;
    a[i] = (u64)d;
    b = (u64)(d >> 64) & 1;
  }
}
static inline V add(const V& a, const V& b) {
  V r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.v[i] + b.v[i];
    r.v[i] = (u64)c;
    c >>= 64;
  }
  if (geq_p(r.v)) sub_p(r.v);
  return r;
}
static inline V sub(const V& a, const V& b) {
  V r;
  u64 bo = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a.v[i] - b.v[i] - bo;
    r.v[i] = (u64)d;
    bo = (u64)(d >> 64) & 1;
  }
  if (bo) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)r.v[i] + RP[i];
      r.v[i] = (u64)c;
      c >>= 64;
    }
  }
  return r;
}
static inline V neg(const V& a) { return sub(zero(), a); }
static inline V mont(const V& a, const V& b) {
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
    u64 m = t[0] * RINV;
    c = ((u128)m * RP[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * RP[j] + t[j];
      t[j - 1] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (u64)c;
    t[4] = t[5] + (u64)(c >> 64);
  }
  V r{{t[0], t[1], t[2], t[3]}};
  if (t[4] || geq_p(r.v)) sub_p(r.v);
  return r;
}
static inline V mul(const V& a, const V& b) { return mont(mont(a, b), V{{RR2[0], RR2[1], RR2[2], RR2[3]}}); }
// a^-1 mod r on the canonical value (0 -> 0)
static const modinv::Modulus R_MODULUS = modinv::make_modulus(RP);
static V inv(const V& a) {
  V x;
  modinv::inverse(a.v, R_MODULUS, x.v);
  return x;
}
static inline bool canonical(const u64* a) { return !geq_p(a); }
}  // namespace fr

// ---- assigned objects (value semantics: a C++ copy is a Rust `.clone()`) ---------------------------
struct Cell {
  uint32_t row = 0;
  uint8_t col = 0;
};
struct HInt {  // AssignedInteger, chips/integer_chip.rs:12-48
  u128 limb[4];
  Cell cell[4];
  uint32_t overflows = 0;
  bool native_cached = false;
  Cell native_cell;  // where the cached native lives (column 0 of its sum row), valid when native_cached
  Fq w;  // value mod p (Montgomery)
};
struct HCond {  // AssignedCondition / small AssignedValue
  u64 value = 0;
  Cell cell;
};
struct HCurv {
  HInt v;
  HCond z;
};
struct HPoint {  // AssignedPoint, chips/ecc_chip.rs:23-58
  HInt x, y;
  HCond z;
  bool has_curv = false;
  HCurv curv;
};
struct HScalar {  // AssignedValue<Fr>: canonical 256-bit
  u64 v[4];
  Cell cell;
};

static const u128 LIMB_MASK = (((u128)1) << 68) - 1;
static const int OVERFLOW_LIMIT = 64, OVERFLOW_THRESHOLD = 32;


// ---- op-record storage: fixed-size chunks, page-locked when a device context exists ------------------------------
// Records are only ever appended and later streamed to the GPU, so they live in 2 MiB chunks (one section of a
// multi_exp recorded on its own thread is ~1.7 MB of records: one chunk) carved from page-locked
// slabs (cudaHostAlloc, portable) once some h2agg_ctx exists in the process -- the H2D copy of a 2^21-row witness is
// ~110 MB and runs at PCIe speed from pinned memory, at a fraction of it from pageable memory.  Without a device
// (recording only: the CPU tests) chunks are plain heap memory.  Chunks go back to a process-wide pool.
static const uint32_t CHUNK_OPS = 8192;
static const size_t SLAB_CHUNKS = 16;  // 32 MiB per cudaHostAlloc
struct ChunkPool {
  std::mutex mu;
  std::vector<WitnessOp*> free_pinned, free_plain;
  size_t pinned_bytes = 0;
  size_t plain_keep = 2048;  // heap chunks kept for reuse (4 GiB): a fresh chunk costs ~0.5 ms of page faults
  size_t pinned_cap = 1024ull << 20;
  ChunkPool() {
    if (const char* e = getenv("H2AGG_WIT_PINNED_MB")) pinned_cap = (size_t)atoll(e) << 20;
  }
  WitnessOp* get(bool* pinned) {
    {
      std::lock_guard<std::mutex> l(mu);
      if (free_pinned.empty() && g_any_device.load() >= 0 && pinned_bytes + SLAB_CHUNKS * CHUNK_OPS * sizeof(WitnessOp) <= pinned_cap) {
        void* slab = nullptr;
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(g_any_device.load());
        cudaError_t e = cudaHostAlloc(&slab, SLAB_CHUNKS * CHUNK_OPS * sizeof(WitnessOp), cudaHostAllocPortable);
        if (prev >= 0) cudaSetDevice(prev);
        if (e == cudaSuccess) {
          pinned_bytes += SLAB_CHUNKS * CHUNK_OPS * sizeof(WitnessOp);
          for (size_t i = 0; i < SLAB_CHUNKS; i++) free_pinned.push_back((WitnessOp*)slab + i * CHUNK_OPS);
        } else {
          cudaGetLastError();
          pinned_cap = 0;  // do not try again
        }
      }
      if (!free_pinned.empty()) {
        WitnessOp* c = free_pinned.back();
        free_pinned.pop_back();
        *pinned = true;
        return c;
      }
    }
    *pinned = false;
    {
      std::lock_guard<std::mutex> l(mu);
      if (!free_plain.empty()) {
        WitnessOp* c = free_plain.back();
        free_plain.pop_back();
        return c;
      }
    }
    void* m = nullptr;
    if (posix_memalign(&m, 4096, CHUNK_OPS * sizeof(WitnessOp)) != 0) throw std::bad_alloc();
    return (WitnessOp*)m;
  }
  void put(WitnessOp* c, bool pinned) {
    std::lock_guard<std::mutex> l(mu);
    if (pinned) free_pinned.push_back(c);
    else if (free_plain.size() < plain_keep) free_plain.push_back(c);
    else free(c);
  }
};
static ChunkPool& chunk_pool() {
  static ChunkPool* p = new ChunkPool();  // never destroyed: chunks may be returned during static teardown
  return *p;
}

struct OpChunk {
  WitnessOp* p = nullptr;
  uint32_t n = 0;
  bool pinned = false;
};
struct OpStore {  // records in arrival order; only the per-opcode COUNTS are kept apart (the device groups the records)
  std::vector<OpChunk> chunks;
  size_t count[WOP_COUNT] = {0, 0, 0, 0, 0, 0};
  OpStore() = default;
  OpStore(const OpStore&) = delete;
  OpStore& operator=(const OpStore&) = delete;
  ~OpStore() { clear(); }
  void clear() {
    for (auto& c : chunks) chunk_pool().put(c.p, c.pinned);
    chunks.clear();
    for (uint32_t k = 0; k < WOP_COUNT; k++) count[k] = 0;
  }
  size_t size() const {
    size_t t = 0;
    for (uint32_t k = 0; k < WOP_COUNT; k++) t += count[k];
    return t;
  }
  void push_back(const WitnessOp& op) {
    if (chunks.empty() || chunks.back().n == CHUNK_OPS) {
      OpChunk c;
      c.p = chunk_pool().get(&c.pinned);
      chunks.push_back(c);
    }
    OpChunk& c = chunks.back();
    c.p[c.n++] = op;
    count[op.opcode]++;
  }
  void add_to_rows(uint32_t delta) {  // rows recorded relative to a tag become absolute (mod 2^32 arithmetic)
    for (auto& c : chunks)
      for (uint32_t i = 0; i < c.n; i++) c.p[i].row += delta;
  }
  void absorb(OpStore& o) {  // take over the chunks of a child store (no copies)
    for (auto& c : o.chunks) chunks.push_back(c);
    o.chunks.clear();
    for (uint32_t k = 0; k < WOP_COUNT; k++) {
      count[k] += o.count[k];
      o.count[k] = 0;
    }
  }
};

// ---- a few host threads for the independent sections of a multi_exp ------------------------------------------------
static std::atomic<unsigned> g_wit_threads{0};  // 0 = not decided yet
static unsigned wit_threads() {
  unsigned n = g_wit_threads.load();
  if (n) return n;
  unsigned t = std::thread::hardware_concurrency();
  if (t == 0) t = 1;
  if (t > 16) t = 16;
  if (const char* e = getenv("H2AGG_WIT_THREADS")) {
    int v = atoi(e);
    if (v >= 1) t = (unsigned)v;
  }
  g_wit_threads.store(t);
  return t;
}
template <class F>
static void parallel_for(size_t n_tasks, F&& task) {
  const unsigned nt = (unsigned)std::min<size_t>(wit_threads(), n_tasks);
  if (nt <= 1) {
    for (size_t i = 0; i < n_tasks; i++) task(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::exception_ptr err;
  std::mutex err_mu;
  auto worker = [&] {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= n_tasks) return;
      try {
        task(i);
      } catch (...) {
        std::lock_guard<std::mutex> l(err_mu);
        if (!err) err = std::current_exception();
        next.store(n_tasks);
        return;
      }
    }
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; t++) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  if (err) std::rethrow_exception(err);
}

// Rows recorded by a child recorder (an independent section of a multi_exp, recorded on its own thread) are relative
// to this tag until the section is stitched into the parent at its final row offset.
static const uint32_t REL_TAG = 0x80000000u;

struct Recorder {
  OpStore ops;
  uint32_t offset = 0;
  // pending RAW128 record
  bool raw_open = false;
  WitnessOp raw;
  // constants
  u128 p_limbs[4];
  std::vector<HPoint> points;
  std::vector<HScalar> scalars;

  Recorder() {
    // limbs of p
    u64 c[4] = {QP[0], QP[1], QP[2], QP[3]};
    canon_to_limbs(c, p_limbs);
  }

  static void canon_to_limbs(const u64* c, u128* out) {
    // 256-bit little-endian -> four 68-bit limbs
    u128 lo = (u128)c[0] | ((u128)c[1] << 64);
    u128 hi = (u128)c[2] | ((u128)c[3] << 64);
    out[0] = lo & LIMB_MASK;
    out[1] = ((lo >> 68) | (hi << 60)) & LIMB_MASK;
    out[2] = (hi >> 8) & LIMB_MASK;
    out[3] = hi >> 76;
  }

  // ---- record emission -----------------------------------------------------------------------
  void flush_raw() {
    if (raw_open) {
      ops.push_back(raw);
      raw_open = false;
    }
  }
  uint32_t raw_row(const u128 c[5]) {  // one advice row of small values
    if (raw_open && raw.aux == 3) flush_raw();
    if (!raw_open) {
      memset(&raw, 0, sizeof(raw));
      raw.opcode = WOP_RAW128;
      raw.row = offset;
      raw_open = true;
    }
    uint32_t r = raw.aux++;
    for (int k = 0; k < 5; k++) {
      raw.v[10 * r + 2 * k] = (u64)c[k];
      raw.v[10 * r + 2 * k + 1] = (u64)(c[k] >> 64);
    }
    return offset++;
  }
  uint32_t row5(u128 a0, u128 a1 = 0, u128 a2 = 0, u128 a3 = 0, u128 a4 = 0) {
    u128 c[5] = {a0, a1, a2, a3, a4};
    return raw_row(c);
  }
  uint32_t raw256_row(const u64 cells[5][4]) {
    flush_raw();
    WitnessOp op;
    memset(&op, 0, sizeof(op));
    op.opcode = WOP_RAW256;
    op.row = offset;
    for (int c = 0; c < 5; c++) memcpy(op.v + 4 * c, cells[c], 32);
    ops.push_back(op);
    return offset++;
  }
  static void put_limbs(u64* dst, const HInt& a) {
    for (int i = 0; i < 4; i++) {
      dst[2 * i] = (u64)a.limb[i];
      dst[2 * i + 1] = (u64)(a.limb[i] >> 64);
    }
  }

  // ---- base gate (gates/base_gate.rs, five/base_gate.rs) ----------------------------------------
  HCond bg_assign_constant(u64 v) {
    HCond c;
    c.value = v;
    c.cell = Cell{row5(v), 0};
    return c;
  }
  HCond bg_mul(const HCond& a, const HCond& b) {  // :302-324 -> cells[2]
    HCond c;
    c.value = a.value * b.value;
    c.cell = Cell{row5(a.value, b.value, c.value), 2};
    return c;
  }
  HCond bg_not(const HCond& a) {  // sum_with_constant([(a,-1)], 1) -> [1-a, a]
    HCond c;
    c.value = 1 - a.value;
    c.cell = Cell{row5(c.value, a.value), 0};
    return c;
  }
  HCond bg_or(const HCond& a, const HCond& b) {
    HCond c;
    c.value = a.value | b.value;
    c.cell = Cell{row5(a.value, b.value, c.value), 2};
    return c;
  }
  HCond bg_xnor(const HCond& a, const HCond& b) {
    HCond c;
    c.value = (a.value == b.value) ? 1 : 0;
    c.cell = Cell{row5(a.value, b.value, c.value), 2};
    return c;
  }
  HCond bg_bisec(const HCond& cond, const HCond& a, const HCond& b) {  // five/base_gate.rs:82-108 -> cells[4]
    HCond c;
    c.value = cond.value ? a.value : b.value;
    c.cell = Cell{row5(cond.value, a.value, cond.value, b.value, c.value), 4};
    return c;
  }
  void bg_assert_constant(const HCond& a) { row5(a.value); }
  void bg_assert_bit(u64 a) { row5(a, a); }

  // ---- integer chip (five/integer_chip.rs) ---------------------------------------------------------
  static u64 chunk(u128 l, int i) { return (u64)(l >> (17 * i)) & 0x1ffff; }
  uint32_t limb_row(u128 n, int nchunks) {  // assign_{nonleading, *_leading}_limb: [c_top..c0, 0.., n]
    u128 c[5] = {0, 0, 0, 0, n};
    for (int k = 0; k < nchunks; k++) c[k] = chunk(n, nchunks - 1 - k);
    return raw_row(c);
  }
  HInt assign_w_limbs(const u128 limbs[4], const Fq& w) {  // :447-464, most significant limb first
    HInt a;
    a.w = w;
    a.overflows = 0;
    a.native_cached = false;
    const int nch[4] = {4, 4, 4, 3};
    for (int i = 3; i >= 0; i--) {
      a.limb[i] = limbs[i];
      a.cell[i] = Cell{limb_row(limbs[i], nch[i]), 4};
    }
    return a;
  }
  HInt assign_w(const Fq& w) {
    u64 c[4];
    q_to_canon(w, c);
    u128 l[4];
    canon_to_limbs(c, l);
    return assign_w_limbs(l, w);
  }
  HInt int_assign_constant(const Fq& w) {  // :784-794
    u64 c[4];
    q_to_canon(w, c);
    HInt a;
    canon_to_limbs(c, a.limb);
    a.w = w;
    for (int i = 0; i < 4; i++) a.cell[i] = Cell{row5(a.limb[i]), 0};
    return a;
  }
  void native(HInt& a) {  // :595-621
    if (a.native_cached) return;
    flush_raw();
    WitnessOp op;
    memset(&op, 0, sizeof(op));
    op.opcode = WOP_NATIVE;
    op.row = offset;
    put_limbs(op.v, a);
    ops.push_back(op);
    a.native_cell = Cell{offset, 0};
    offset += 1;
    a.native_cached = true;
  }
  void find_w_modulus_ceil(const HInt& a, u128 out[4]) {  // :31-51
    // smallest multiple of p that is >= (overflows + 1) * 2^254, re-limbed with borrowed headroom
    // n = ceil(((ov+1) << 254) / p): p > 2^253 so n is in [ov+1, 2(ov+1)]
    u64 k = a.overflows + 1;
    // compute n*p for candidates with 320-bit arithmetic (5 x u64)
    auto mul_small = [&](u64 n, u64* out5) {
      u128 c = 0;
      for (int i = 0; i < 4; i++) {
        c += (u128)QP[i] * n;
        out5[i] = (u64)c;
        c >>= 64;
      }
      out5[4] = (u64)c;
    };
    u64 target[5] = {0, 0, 0, k << 62, k >> 2};  // k * 2^254
    u64 n = k, up[5];
    for (;; n++) {
      mul_small(n, up);
      bool ge = true;
      for (int i = 4; i >= 0; i--) {
        if (up[i] > target[i]) break;
        if (up[i] < target[i]) { ge = false; break; }
      }
      if (ge) break;
    }
    // limbs[i] = upper mod 2^68 + k * 2^68 ; upper = (upper - limbs[i]) / 2^68
    // work on a signed-free representation: upper as 5 x u64, subtraction never underflows overall
    auto low68 = [&](const u64* x) { return ((u128)x[0] | ((u128)x[1] << 64)) & LIMB_MASK; };
    auto shr68 = [&](u64* x) {
      for (int i = 0; i < 5; i++) {
        u128 lo = (i + 1 < 5) ? x[i + 1] : 0, hi = (i + 2 < 5) ? x[i + 2] : 0;
        u128 v = (lo >> 4) | (hi << 60);
        x[i] = (u64)v;
      }
    };
    for (int i = 0; i < 3; i++) {
      u128 rem = low68(up) + (u128)k * (((u128)1) << 68);
      // upper = (upper - rem) / 2^68  = (upper >> 68) - k   (the low 68 bits cancel)
      shr68(up);
      u128 borrow = k;
      for (int j = 0; j < 5 && borrow; j++) {
        u128 d = (u128)up[j] - (u64)borrow;
        up[j] = (u64)d;
        borrow = (d >> 64) & 1;
      }
      out[i] = rem;
    }
    out[3] = (u128)up[0] | ((u128)up[1] << 64);
  }
  void reduce(HInt& a) {  // :483-581
    if (a.overflows == 0) return;
    if (a.overflows >= (uint32_t)OVERFLOW_LIMIT) throw std::runtime_error("integer overflow limit exceeded");
    flush_raw();
    u64 c[4];
    q_to_canon(a.w, c);
    u128 rl[4];
    canon_to_limbs(c, rl);
    WitnessOp op;
    memset(&op, 0, sizeof(op));
    op.opcode = WOP_REDUCE;
    op.row = offset;
    op.flags = a.native_cached ? 1u : 0u;
    put_limbs(op.v, a);
    HInt rem;
    rem.w = a.w;
    for (int i = 0; i < 4; i++) rem.limb[i] = rl[i];
    put_limbs(op.v + 8, rem);
    ops.push_back(op);
    const uint32_t r0 = offset;
    // rows: assign_w(rem) 4 (limb 3 at r0 .. limb 0 at r0+3), [d,v], native(rem), [native(a)], 2 checks
    for (int i = 0; i < 4; i++) rem.cell[i] = Cell{r0 + (uint32_t)(3 - i), 4};
    offset += 4 + 1 + 1 + (a.native_cached ? 0 : 1) + 2;
    rem.overflows = 0;
    rem.native_cached = true;
    rem.native_cell = Cell{r0 + 5, 0};  // native(rem) follows the [d, v] row
    a = rem;
  }
  void conditionally_reduce(HInt& a) {
    if (a.overflows >= (uint32_t)OVERFLOW_THRESHOLD) reduce(a);
  }
  HInt int_add(const HInt& a, const HInt& b) {  // :641-658 rows [s, a_i, b_i]
    HInt r;
    for (int i = 0; i < 4; i++) {
      r.limb[i] = a.limb[i] + b.limb[i];
      r.cell[i] = Cell{row5(r.limb[i], a.limb[i], b.limb[i]), 0};
    }
    r.overflows = a.overflows + b.overflows + 1;
    r.w = q_add(a.w, b.w);
    conditionally_reduce(r);
    return r;
  }
  HInt int_sub(const HInt& a, const HInt& b) {  // :660-683
    u128 up[4];
    find_w_modulus_ceil(b, up);
    HInt r;
    for (int i = 0; i < 4; i++) {
      r.limb[i] = a.limb[i] + up[i] - b.limb[i];
      r.cell[i] = Cell{row5(r.limb[i], a.limb[i], b.limb[i]), 0};
    }
    r.overflows = a.overflows + (b.overflows + 1) + 1;
    r.w = q_sub(a.w, b.w);
    conditionally_reduce(r);
    return r;
  }
  HInt int_neg(const HInt& a) {  // :685-707
    u128 up[4];
    find_w_modulus_ceil(a, up);
    HInt r;
    for (int i = 0; i < 4; i++) {
      r.limb[i] = up[i] - a.limb[i];
      r.cell[i] = Cell{row5(r.limb[i], a.limb[i]), 0};
    }
    r.overflows = a.overflows + 1;
    r.w = q_sub(q_zero(), a.w);
    conditionally_reduce(r);
    return r;
  }
  HInt int_mul_small_constant(HInt& a, u64 b) {  // :808-835
    if (a.overflows * b >= (u64)OVERFLOW_LIMIT) reduce(a);
    HInt r;
    for (int i = 0; i < 4; i++) {
      r.limb[i] = a.limb[i] * b;
      r.cell[i] = Cell{row5(r.limb[i], a.limb[i]), 0};
    }
    r.overflows = a.overflows * (uint32_t)b;
    r.w = q_mul(a.w, q_small(b));
    conditionally_reduce(r);
    return r;
  }
  HInt int_bisec(const HCond& cond, const HInt& a, const HInt& b) {  // :845-866
    HInt r;
    for (int i = 0; i < 4; i++) {
      r.limb[i] = cond.value ? a.limb[i] : b.limb[i];
      r.cell[i] = Cell{row5(cond.value, a.limb[i], cond.value, b.limb[i], r.limb[i]), 4};
    }
    r.overflows = a.overflows > b.overflows ? a.overflows : b.overflows;
    r.w = cond.value ? a.w : b.w;
    return r;
  }
  // x * y = d * p + z with `fresh` (y or z) assign_w'd here; sets the native caches like the
  // reference's &mut borrows do                                                          :104-320
  void mul_eq(HInt& x, HInt* y /*null = square*/, HInt& z, bool fresh_is_y) {
    flush_raw();
    WitnessOp op;
    memset(&op, 0, sizeof(op));
    op.opcode = WOP_MULEQ;
    op.row = offset;
    const bool sq = (y == nullptr);
    op.flags = (x.native_cached ? 1u : 0u) | ((!sq && y->native_cached) ? 2u : 0u) | (z.native_cached ? 4u : 0u) |
               (sq ? 8u : 0u) | (fresh_is_y ? 16u : 0u);
    put_limbs(op.v, x);
    if (!sq) put_limbs(op.v + 8, *y);
    put_limbs(op.v + 16, z);
    ops.push_back(op);
    HInt& fresh = fresh_is_y ? *y : z;
    for (int i = 0; i < 4; i++) fresh.cell[i] = Cell{offset + (uint32_t)(3 - i), 4};
    uint32_t rows = 4 + 4 + 10 + 4 + 4;
    // natives in the order _mul_equation_on_native asks for them: a, b, (d), rem   :254-320
    if (!x.native_cached) x.native_cell = Cell{offset + rows++, 0};
    if (!sq && !y->native_cached) y->native_cell = Cell{offset + rows++, 0};
    rows += 1;
    if (!z.native_cached) z.native_cell = Cell{offset + rows++, 0};
    rows += 1;
    offset += rows;
    x.native_cached = true;
    if (!sq) y->native_cached = true;
    z.native_cached = true;
  }
  HInt fresh_int(const Fq& w) {  // canonical limbs, rows assigned by the following mul_eq
    HInt a;
    u64 c[4];
    q_to_canon(w, c);
    canon_to_limbs(c, a.limb);
    a.w = w;
    a.overflows = 0;
    a.native_cached = false;
    return a;
  }
  HInt int_mul(HInt& a, HInt& b) {  // :709-726
    if (a.overflows >= (uint32_t)OVERFLOW_LIMIT || b.overflows >= (uint32_t)OVERFLOW_LIMIT) throw std::runtime_error("mul: overflow");
    HInt rem = fresh_int(q_mul(a.w, b.w));
    if (&a == &b) mul_eq(a, nullptr, rem, false);
    else mul_eq(a, &b, rem, false);
    return rem;
  }
  HInt int_square(HInt& a) {  // :728-743
    HInt rem = fresh_int(q_mul(a.w, a.w));
    mul_eq(a, nullptr, rem, false);
    return rem;
  }
  HCond int_is_zero(HInt& a) {  // :796-806
    reduce(a);
    flush_raw();
    WitnessOp op;
    memset(&op, 0, sizeof(op));
    op.opcode = WOP_ISZERO;
    op.row = offset;
    op.flags = a.native_cached ? 1u : 0u;
    put_limbs(op.v, a);
    ops.push_back(op);
    uint32_t rows = 1 + 2 + (a.native_cached ? 0 : 1) + 1 + 2 + 1 + 2 + 1 + 1;
    if (!a.native_cached) a.native_cell = Cell{offset + 3, 0};  // after is_pure_zero's sum + 2 inversion rows
    offset += rows;
    a.native_cached = true;
    HCond c;
    c.value = q_is_zero(a.w) ? 1 : 0;  // a reduced value is canonical, so "== p" never holds
    c.cell = Cell{offset - 1, 2};
    return c;
  }
  HCond int_is_equal(HInt& a, HInt& b) {
    HInt diff = int_sub(a, b);
    return int_is_zero(diff);
  }
  HCond int_div(HInt& a, HInt& b, HInt* out_c) {  // :745-782
    HCond is_b_zero = int_is_zero(b);
    HCond a_coeff = bg_not(is_b_zero);
    reduce(a);
    HInt a2;
    for (int i = 0; i < 4; i++) {
      a2.limb[i] = a_coeff.value ? a.limb[i] : 0;
      a2.cell[i] = Cell{row5(a.limb[i], a_coeff.value, a2.limb[i]), 2};
    }
    a2.overflows = a.overflows;
    a2.native_cached = false;
    a2.w = a_coeff.value ? a.w : q_zero();
    Fq c = q_is_zero(b.w) ? q_zero() : q_mul(q_inv(b.w), a2.w);
    HInt ci = fresh_int(c);
    mul_eq(b, &ci, a2, true);
    *out_c = ci;
    return is_b_zero;
  }

  // ---- ecc chip (chips/ecc_chip.rs) ------------------------------------------------------------------
  HCurv& curvature(HPoint& a) {  // :280-307
    if (!a.has_curv) {
      HInt x_square = int_square(a.x);
      HInt numerator = int_mul_small_constant(x_square, 3);
      HInt denominator = int_mul_small_constant(a.y, 2);
      HCurv c;
      c.z = int_div(numerator, denominator, &c.v);
      a.curv = c;
      a.has_curv = true;
    }
    return a.curv;
  }
  HCurv bisec_curvature(const HCond& cond, const HCurv& a, const HCurv& b) {
    HCurv r;
    r.v = int_bisec(cond, a.v, b.v);
    r.z = bg_bisec(cond, a.z, b.z);
    return r;
  }
  HPoint bisec_point(const HCond& cond, const HPoint& a, const HPoint& b) {
    HPoint r;
    r.x = int_bisec(cond, a.x, b.x);
    r.y = int_bisec(cond, a.y, b.y);
    r.z = bg_bisec(cond, a.z, b.z);
    return r;
  }
  HPoint bisec_point_with_curvature(const HCond& cond, HPoint& a, HPoint& b) {
    HPoint r;
    r.x = int_bisec(cond, a.x, b.x);
    r.y = int_bisec(cond, a.y, b.y);
    r.z = bg_bisec(cond, a.z, b.z);
    HCurv& ca = curvature(a);
    HCurv& cb = curvature(b);
    r.curv = bisec_curvature(cond, ca, cb);
    r.has_curv = true;
    return r;
  }
  HPoint lambda_to_point(HCurv& lambda, const HPoint& a, const HPoint& b) {  // :356-382
    HInt& l = lambda.v;
    HInt l_square = int_square(l);
    HInt t = int_sub(l_square, a.x);
    HInt cx = int_sub(t, b.x);
    HInt t2 = int_sub(a.x, cx);
    HInt t3 = int_mul(t2, l);
    HInt cy = int_sub(t3, a.y);
    HPoint p;
    p.x = cx; p.y = cy; p.z = lambda.z;
    return p;
  }
  HPoint ecc_add(HPoint& a, const HPoint& b) {  // :383-408
    HInt diff_x = int_sub(a.x, b.x);
    HInt diff_y = int_sub(a.y, b.y);
    HCurv tangent;
    HCond x_eq = int_div(diff_y, diff_x, &tangent.v);
    HCond y_eq = int_is_zero(diff_y);
    HCond eq = bg_mul(x_eq, y_eq);
    tangent.z = x_eq;
    HCurv& curv = curvature(a);
    HCurv lambda = bisec_curvature(eq, curv, tangent);
    HPoint p = lambda_to_point(lambda, a, b);
    p = bisec_point(a.z, b, p);
    p = bisec_point(b.z, a, p);
    return p;
  }
  HPoint ecc_double(HPoint& a) {  // :409-419
    HCurv c = curvature(a);  // clone
    HPoint p = lambda_to_point(c, a, a);
    p.z = bg_bisec(a.z, a.z, p.z);
    return p;
  }
  HPoint assign_identity() {  // :517-527
    HInt zero = int_assign_constant(q_zero());
    HCond one = bg_assign_constant(1);
    HPoint p;
    p.x = zero; p.y = zero; p.z = one;
    p.curv.v = zero; p.curv.z = one;
    p.has_curv = true;
    return p;
  }
  // `y_over_x`: the caller already holds y * x^-1 (constant_mul batches the inversions of its whole table)
  HPoint assign_constant_point_with_curvature(bool identity, const Fq& x, const Fq& y, const Fq* y_over_x = nullptr) {  // :438-472
    Fq xv = identity ? q_zero() : x, yv = identity ? q_zero() : y;
    HPoint p;
    p.curv.v = int_assign_constant(q_is_zero(xv) ? q_zero() : (y_over_x ? *y_over_x : q_mul(yv, q_inv(xv))));
    p.curv.z = bg_assign_constant(q_is_zero(xv) ? 1 : 0);
    p.x = int_assign_constant(xv);
    p.y = int_assign_constant(yv);
    p.z = bg_assign_constant(identity ? 1 : 0);
    p.has_curv = true;
    return p;
  }
  HPoint assign_constant_point(bool identity, const Fq& x, const Fq& y) {  // :420-437
    HPoint p;
    p.x = int_assign_constant(identity ? q_zero() : x);
    p.y = int_assign_constant(identity ? q_zero() : y);
    p.z = bg_assign_constant(identity ? 1 : 0);
    return p;
  }
  HPoint assign_point(bool identity, const Fq& x, const Fq& y) {  // :473-500 (on-curve check)
    HPoint p;
    p.x = assign_w(identity ? q_zero() : x);
    p.y = assign_w(identity ? q_zero() : y);
    HCond z;
    z.value = identity ? 1 : 0;
    z.cell = Cell{row5(z.value), 0};  // base_gate.assign
    p.z = z;
    HInt b = int_assign_constant(q_small(3));
    HInt y2 = int_square(p.y);
    HInt x2 = int_square(p.x);
    HInt x3 = int_mul(x2, p.x);
    HInt right = int_add(x3, b);
    HCond eq = int_is_equal(y2, right);
    HCond eq_or_identity = bg_or(eq, z);
    if (!eq_or_identity.value) throw std::runtime_error("assign_point: point is not on the curve");
    bg_assert_constant(eq_or_identity);
    return p;
  }
  HPoint ecc_neg(const HPoint& a) {
    HPoint r;
    r.x = a.x;
    r.y = int_neg(a.y);
    r.z = a.z;
    return r;
  }
  HPoint ecc_sub(HPoint& a, const HPoint& b) {
    HPoint nb = ecc_neg(b);
    return ecc_add(a, nb);
  }
  HPoint ecc_reduce(HPoint& a) {  // :569-580
    reduce(a.x);
    reduce(a.y);
    HPoint id = assign_identity();
    return bisec_point(a.z, id, a);
  }
  // NativeEccChip::decompose_scalar, chips/native_ecc_chip.rs:42-132 -> windows big-endian, bits little-endian
  std::vector<std::vector<HCond>> decompose_scalar(const HScalar& s, int window) {
    const int windows = (254 - 1 + window) / window;
    std::vector<std::vector<HCond>> ret;
    u64 cur[4] = {s.v[0], s.v[1], s.v[2], s.v[3]};
    for (int w = 0; w < windows; w++) {
      u64 cells[5][4];
      memset(cells, 0, sizeof(cells));
      std::vector<HCond> bits(window);
      for (int i = 0; i < window; i++) {
        bits[i].value = (cur[0] >> i) & 1;
        cells[i][0] = bits[i].value;
      }
      memcpy(cells[4], cur, 32);
      uint32_t row = raw256_row(cells);
      for (int i = 0; i < window; i++) bits[i].cell = Cell{row, (uint8_t)i};
      ret.push_back(bits);
      // cur >>= window
      for (int k = 0; k < 4; k++) cur[k] = (cur[k] >> window) | (k + 1 < 4 ? cur[k + 1] << (64 - window) : 0);
    }
    std::vector<std::vector<HCond>> be(ret.rbegin(), ret.rend());
    for (auto& win : be)
      for (auto& bit : win) bg_assert_bit(bit.value);
    return be;
  }
  HPoint pick_candidate(const std::vector<HPoint>& candidates, const std::vector<HCond>& bits_le) {
    bool all_curv = candidates.size() == 16 && bits_le.size() == 4;
    for (const HPoint& c : candidates) all_curv = all_curv && c.has_curv;
    if (all_curv) {
      // every candidate already carries its curvature (shamir's tables): bisec_point_with_curvature mutates nothing,
      // so the reference's clone of the table is not needed -- halve in place over a small array
      HPoint buf[8];
      for (int k = 0; k < 8; k++)
        buf[k] = bisec_point_with_curvature(bits_le[0], const_cast<HPoint&>(candidates[2 * k + 1]), const_cast<HPoint&>(candidates[2 * k]));
      for (int level = 1, len = 4; level < 4; level++, len >>= 1)
        for (int k = 0; k < len; k++) buf[k] = bisec_point_with_curvature(bits_le[level], buf[2 * k + 1], buf[2 * k]);
      return buf[0];
    }
    std::vector<HPoint> curr = candidates;  // clone: curvature rows a candidate acquires here are dropped with it
    for (const HCond& bit : bits_le) {
      std::vector<HPoint> next;
      next.reserve(curr.size() / 2);
      for (size_t k = 0; k + 1 < curr.size(); k += 2) next.push_back(bisec_point_with_curvature(bit, curr[k + 1], curr[k]));
      curr.swap(next);
    }
    return curr[0];
  }
  HPoint ecc_mul(HPoint& a, const HScalar& s) {  // :86-138
    auto windows = decompose_scalar(s, 4);
    std::vector<HPoint> cands;
    cands.push_back(assign_identity());
    cands.push_back(a);
    for (int i = 2; i < 16; i++) {
      HPoint ai = ecc_add(cands[i - 1], a);
      cands.push_back(ai);
    }
    HPoint acc = pick_candidate(cands, windows[0]);
    for (size_t w = 1; w < windows.size(); w++) {
      for (int k = 0; k < 4; k++) acc = ecc_double(acc);
      HPoint curr = pick_candidate(cands, windows[w]);
      acc = ecc_add(curr, acc);
    }
    return acc;
  }
  // -- sections recorded on their own threads: a child starts at REL_TAG, the parent stitches it in at its offset
  static void fix_cell(Cell& c, uint32_t delta) {
    if (c.row & REL_TAG) c.row += delta;
  }
  static void fix_int(HInt& a, uint32_t delta) {
    for (int i = 0; i < 4; i++) fix_cell(a.cell[i], delta);
    fix_cell(a.native_cell, delta);
  }
  static void fix_point(HPoint& p, uint32_t delta) {
    fix_int(p.x, delta);
    fix_int(p.y, delta);
    fix_cell(p.z.cell, delta);
    fix_int(p.curv.v, delta);
    fix_cell(p.curv.z.cell, delta);
  }
  void begin_child() {
    offset = REL_TAG;
  }
  uint32_t child_rows() const { return offset - REL_TAG; }

  // EccChipOps::shamir (:139-244).  The row layout is the reference's, row for row; the ORDER OF RECORDING is not: the
  // candidate table of each point, and the inner sum of each window (which does not depend on the accumulator -- the
  // reference itself relies on every inner round having the same size, :193-226), are independent of one another, so
  // they are recorded concurrently into child recorders and stitched in at their row offsets.  Only the 63 x (4
  // doublings + 1 addition) of the accumulator chain stay sequential.
  HPoint ecc_shamir(std::vector<HPoint>& pts, const std::vector<HScalar>& scalars) {
    std::vector<std::vector<std::vector<HCond>>> windows;
    for (auto& s : scalars) windows.push_back(decompose_scalar(s, 4));
    HPoint identity = assign_identity();
    flush_raw();
    const size_t np = pts.size();
    std::vector<std::vector<HPoint>> pc(np);
    {
      std::vector<Recorder> kids(np);
      parallel_for(np, [&](size_t pi) {
        Recorder& r = kids[pi];
        r.begin_child();
        std::vector<HPoint>& cands = pc[pi];
        cands.reserve(16);
        cands.push_back(identity);
        cands.push_back(pts[pi]);
        for (int i = 2; i < 16; i++) {
          HPoint ai = r.ecc_add(cands[i - 1], pts[pi]);
          r.curvature(ai);
          cands.push_back(ai);
        }
        r.flush_raw();
      });
      std::vector<uint32_t> delta(np);
      for (size_t pi = 0; pi < np; pi++) {
        delta[pi] = offset - REL_TAG;
        offset += kids[pi].child_rows();
      }
      parallel_for(np, [&](size_t pi) {
        kids[pi].ops.add_to_rows(delta[pi]);
        for (HPoint& c : pc[pi]) fix_point(c, delta[pi]);
      });
      for (size_t pi = 0; pi < np; pi++) ops.absorb(kids[pi].ops);
    }
    const size_t nw = windows.empty() ? 0 : windows[0].size();
    std::vector<Recorder> kids(nw);
    std::vector<HPoint> inner(nw);
    parallel_for(nw, [&](size_t wi) {
      Recorder& r = kids[wi];
      r.begin_child();
      bool have_inner = false;
      HPoint in;
      for (size_t pi = 0; pi < np; pi++) {
        HPoint ci = r.pick_candidate(pc[pi], windows[pi][wi]);
        if (!have_inner) { in = ci; have_inner = true; }
        else in = r.ecc_add(ci, in);
      }
      r.flush_raw();
      inner[wi] = in;
    });
    std::vector<uint32_t> delta(nw);
    bool have_acc = false;
    HPoint acc;
    for (size_t wi = 0; wi < nw; wi++) {
      flush_raw();
      delta[wi] = offset - REL_TAG;
      offset += kids[wi].child_rows();
      fix_point(inner[wi], delta[wi]);
      if (!have_acc) { acc = inner[wi]; have_acc = true; }
      else {
        for (int k = 0; k < 4; k++) acc = ecc_double(acc);
        acc = ecc_add(inner[wi], acc);
      }
    }
    flush_raw();
    parallel_for(nw, [&](size_t wi) { kids[wi].ops.add_to_rows(delta[wi]); });
    for (size_t wi = 0; wi < nw; wi++) ops.absorb(kids[wi].ops);
    return acc;
  }

  // ---- ScalarChip = BaseGateOps on native Fr values (halo2-snark-aggregator-circuit/src/chips/scalar_chip.rs:17-127
  // over gates/base_gate.rs:193-511).  Every recipe is one or more RAW256 rows of canonical cells; the value the chain
  // continues with is computed here (a handful of host products per row -- the native-field part of an aggregation
  // witness is ~5 % of its rows, the wrong-field part goes through the expansion kernel).
  HScalar sc_cell(const fr::V& v, uint32_t row, uint8_t col) {
    HScalar s;
    memcpy(s.v, v.v, 32);
    s.cell = Cell{row, col};
    return s;
  }
  static fr::V sv(const HScalar& s) { return fr::V{{s.v[0], s.v[1], s.v[2], s.v[3]}}; }
  uint32_t sc_row(const fr::V* c0, const fr::V* c1 = nullptr, const fr::V* c2 = nullptr, const fr::V* c3 = nullptr,
                  const fr::V* c4 = nullptr) {
    u64 cells[5][4];
    memset(cells, 0, sizeof(cells));
    const fr::V* c[5] = {c0, c1, c2, c3, c4};
    for (int i = 0; i < 5; i++)
      if (c[i]) memcpy(cells[i], c[i]->v, 32);
    return raw256_row(cells);
  }
  HScalar sc_assign(const fr::V& v) {  // BaseGateOps::assign / assign_constant: advice [v, 0, 0, 0, 0]   :499-511
    return sc_cell(v, sc_row(&v), 0);
  }
  // BaseGateOps::sum_with_constant (:193-262): sum_i coeff_i * elem_i + constant, chained over rows through next_coeff
  HScalar sc_sum_with_constant(const std::vector<std::pair<HScalar, fr::V>>& elems, const fr::V& constant) {
    // sum_i coeff_i * elem_i on canonical values: the Montgomery products of a line are summed first and brought back
    // from the R^-1 domain once (n + 1 Montgomery steps per line instead of 2 n); a coefficient 1 costs none
    static const fr::V R2{{fr::RR2[0], fr::RR2[1], fr::RR2[2], fr::RR2[3]}};
    auto is_one = [](const fr::V& c) { return c.v[0] == 1 && (c.v[1] | c.v[2] | c.v[3]) == 0; };
    const size_t columns = 5;
    bool have_acc = false;
    fr::V acc = fr::zero();
    size_t curr = 0;
    while (elems.size() - curr + (have_acc ? 1 : 0) + 1 > columns) {
      const size_t line_len = columns - (have_acc ? 1 : 0);
      fr::V plain = fr::zero(), scaled = fr::zero();   // sum of the coefficient-1 terms / of mont(elem, coeff)
      fr::V cells[5] = {fr::zero(), fr::zero(), fr::zero(), fr::zero(), fr::zero()};
      for (size_t i = 0; i < line_len; i++) {
        cells[i] = sv(elems[curr + i].first);
        if (is_one(elems[curr + i].second)) plain = fr::add(plain, cells[i]);
        else scaled = fr::add(scaled, fr::mont(cells[i], elems[curr + i].second));
      }
      if (have_acc) cells[4] = acc;  // one_line_with_last_base: the running sum rides in the last column
      sc_row(&cells[0], &cells[1], &cells[2], &cells[3], &cells[4]);
      curr += line_len;
      acc = fr::add(acc, fr::add(plain, fr::mont(scaled, R2)));
      have_acc = true;
    }
    fr::V plain = fr::add(constant, acc), scaled = fr::zero();
    fr::V cells[5] = {fr::zero(), fr::zero(), fr::zero(), fr::zero(), fr::zero()};
    size_t k = 1;
    for (size_t i = curr; i < elems.size(); i++, k++) {
      cells[k] = sv(elems[i].first);
      if (is_one(elems[i].second)) plain = fr::add(plain, cells[k]);
      else scaled = fr::add(scaled, fr::mont(cells[k], elems[i].second));
    }
    fr::V sum = fr::add(plain, fr::mont(scaled, R2));
    cells[0] = sum;
    if (have_acc) cells[4] = acc;
    return sc_cell(sum, sc_row(&cells[0], &cells[1], &cells[2], &cells[3], &cells[4]), 0);
  }
  HScalar sc_add(const HScalar& a, const HScalar& b) { return sc_sum_with_constant({{a, fr::small(1)}, {b, fr::small(1)}}, fr::zero()); }
  HScalar sc_sub(const HScalar& a, const HScalar& b) {
    return sc_sum_with_constant({{a, fr::small(1)}, {b, fr::neg(fr::small(1))}}, fr::zero());
  }
  HScalar sc_mul(const HScalar& a, const HScalar& b) {  // :302-324 [a, b, c] -> cells[2]
    fr::V x = sv(a), y = sv(b), c = fr::mul(x, y);
    return sc_cell(c, sc_row(&x, &y, &c), 2);
  }
  HScalar sc_div_unsafe(const HScalar& a, const HScalar& b) {  // :478-497 [b, c, a] -> cells[1]; b = 0 panics in Rust
    fr::V x = sv(a), y = sv(b);
    if (fr::is_zero(y)) throw std::runtime_error("div_unsafe: division by zero (the reference unwraps the inverse)");
    fr::V c = fr::mul(fr::inv(y), x);
    return sc_cell(c, sc_row(&y, &c, &x), 1);
  }
  HScalar sc_mul_add_constant(const HScalar& a, const HScalar& b, const fr::V& c) {  // :326-349 [a, b, d] -> cells[2]
    fr::V x = sv(a), y = sv(b), d = fr::add(fr::mul(x, y), c);
    return sc_cell(d, sc_row(&x, &y, &d), 2);
  }
  // value of IntegerChip::native(a) = sum_i limb_i * 2^(68 i) in Fr (limbs may carry overflow)   five/integer_chip.rs:595-621
  static fr::V native_value(const HInt& a) {
    static const fr::V E[4] = {fr::V{{1, 0, 0, 0}}, fr::V{{0, 0x10, 0, 0}}, fr::V{{0, 0, 0x100, 0}}, fr::V{{0, 0, 0, 0x1000}}};
    fr::V acc = fr::zero();
    for (int i = 0; i < 4; i++) acc = fr::add(acc, fr::mul(fr::from_u128(a.limb[i]), E[i]));
    return acc;
  }
  // PoseidonEncodeChip::encode_point (chips/encode_chip.rs:18-33): native(x), native(y) on CLONES of the coordinates
  // (the caches the clones acquire are dropped, so encoding the same point again costs the rows again)
  void encode_point(const HPoint& p, HScalar out[2]) {
    HInt px = p.x, py = p.y;
    native(px);
    out[0] = sc_cell(native_value(px), px.native_cell.row, px.native_cell.col);
    native(py);
    out[1] = sc_cell(native_value(py), py.native_cell.row, py.native_cell.col);
  }
  HScalar int_get_last_bit(const HInt& a) {  // five/integer_chip.rs:874-901 (a reduced: limb 0 < 2^68)
    const u128 l0 = a.limb[0];
    const u128 d = l0 >> 1;
    const u64 bit = (u64)(l0 & 1);
    limb_row(d, 4);                            // assign_nonleading_limb(l0 / 2)
    uint32_t r = row5(d, bit, l0);             // 2 d + bit - l0 = 0
    bg_assert_bit(bit);
    return sc_cell(fr::small(bit), r, 1);
  }
  // Halo2VerifierCircuits::synthesize, second region (verify_circuit.rs:264-368): reduce the four coordinates, take the
  // parity bits of the y's, pack x (and the bit) into two 136-bit halves per point -> the four cells bound to the
  // instance column by constrain_instance.
  void expose_final_pair(HPoint p[2], HScalar out[4]) {
    static const fr::V E1{{0, 0x10, 0, 0}}, E2{{0, 0, 0x100, 0}};
    for (int i = 0; i < 2; i++) {
      reduce(p[i].x);
      reduce(p[i].y);
    }
    HScalar bit[2];
    for (int i = 0; i < 2; i++) bit[i] = int_get_last_bit(p[i].y);
    for (int i = 0; i < 2; i++) {
      const HInt& x = p[i].x;
      HScalar l[4];
      for (int j = 0; j < 4; j++) l[j] = sc_cell(fr::from_u128(x.limb[j]), x.cell[j].row, x.cell[j].col);
      out[2 * i] = sc_sum_with_constant({{l[0], fr::small(1)}, {l[1], E1}}, fr::zero());
      out[2 * i + 1] = sc_sum_with_constant({{l[2], fr::small(1)}, {l[3], E1}, {bit[i], E2}}, fr::zero());
    }
  }
  // EccChipOps::assert_equal (chips/ecc_chip.rs:528-548), used for the `coherent` commitments (verify_circuit.rs:487-493)
  void ecc_assert_equal(HPoint& a, HPoint& b) {
    HCond eq_x = int_is_equal(a.x, b.x);
    HCond eq_y = int_is_equal(a.y, b.y);
    HCond eq_z = bg_xnor(eq_x, eq_y);
    HCond eq_xy = bg_mul(eq_x, eq_y);
    HCond eq_xyz = bg_mul(eq_xy, eq_z);
    HCond both = bg_mul(a.z, b.z);
    HCond eq = bg_or(eq_xyz, both);
    if (!eq.value) throw std::runtime_error("assert_equal: the points differ");
    bg_assert_constant(eq);
  }
};

// a constant point of constant_mul's table (host, Fq Montgomery)
struct Aff {
  Fq x, y;
  bool inf;
};
// ---- the constant table of EccChipOps::constant_mul (chips/ecc_chip.rs:245-279) ---------------------------------
// Per 2-bit window j the chip assigns the constants B_j, 2 B_j, 3 B_j with B_(j+1) = 4 B_j, each with its "curvature"
// y / x.  Done with affine additions that is six field inversions per window on a sequential chain (127 windows); here the
// chain runs in Jacobian coordinates and all 3 x 127 points are brought back with ONE inversion (Montgomery's trick), a
// second batch gives every y / x.  Same field elements: affine coordinates are unique.
struct Jac {
  Fq X, Y, Z;
};
static Jac jac_dbl(const Jac& p) {  // a = 0: dbl-2009-l
  Fq A = q_mul(p.X, p.X), B = q_mul(p.Y, p.Y), C = q_mul(B, B);
  Fq t = q_add(p.X, B);
  Fq D = q_sub(q_sub(q_mul(t, t), A), C);
  D = q_add(D, D);
  Fq E = q_add(q_add(A, A), A), F = q_mul(E, E);
  Jac r;
  r.X = q_sub(F, q_add(D, D));
  Fq C8 = q_add(C, C);
  C8 = q_add(C8, C8);
  C8 = q_add(C8, C8);
  r.Y = q_sub(q_mul(E, q_sub(D, r.X)), C8);
  Fq yz = q_mul(p.Y, p.Z);
  r.Z = q_add(yz, yz);
  return r;
}
static Jac jac_add(const Jac& p, const Jac& q) {  // distinct, finite points: add-2007-bl
  Fq Z1Z1 = q_mul(p.Z, p.Z), Z2Z2 = q_mul(q.Z, q.Z);
  Fq U1 = q_mul(p.X, Z2Z2), U2 = q_mul(q.X, Z1Z1);
  Fq S1 = q_mul(q_mul(p.Y, q.Z), Z2Z2), S2 = q_mul(q_mul(q.Y, p.Z), Z1Z1);
  Fq H = q_sub(U2, U1);
  Fq I = q_add(H, H);
  I = q_mul(I, I);
  Fq J = q_mul(H, I);
  Fq rr = q_sub(S2, S1);
  rr = q_add(rr, rr);
  Fq V = q_mul(U1, I);
  Jac r;
  r.X = q_sub(q_sub(q_mul(rr, rr), J), q_add(V, V));
  Fq S1J = q_mul(S1, J);
  r.Y = q_sub(q_mul(rr, q_sub(V, r.X)), q_add(S1J, S1J));
  Fq zs = q_add(p.Z, q.Z);
  r.Z = q_mul(q_sub(q_sub(q_mul(zs, zs), Z1Z1), Z2Z2), H);
  return r;
}
static void q_batch_inv(std::vector<Fq>& v) {  // in place; zeros stay zero
  std::vector<Fq> pre(v.size());
  Fq run = q_small(1);
  for (size_t i = 0; i < v.size(); i++) {
    pre[i] = run;
    if (!q_is_zero(v[i])) run = q_mul(run, v[i]);
  }
  Fq inv = q_inv(run);
  for (size_t i = v.size(); i-- > 0;) {
    if (q_is_zero(v[i])) continue;
    Fq t = q_mul(inv, pre[i]);
    inv = q_mul(inv, v[i]);
    v[i] = t;
  }
}
struct ConstWindow {
  Aff p[3];     // B, 2B, 3B
  Fq lam[3];    // y / x of each (0 when x = 0)
};
static std::vector<ConstWindow> constant_mul_table(const Aff& base, size_t n_windows) {
  std::vector<ConstWindow> out(n_windows);
  if (base.inf || n_windows == 0) {
    for (auto& w : out)
      for (int t = 0; t < 3; t++) {
        w.p[t] = Aff{q_zero(), q_zero(), true};
        w.lam[t] = q_zero();
      }
    return out;
  }
  // BN254 G1 has prime order and no point with y = 0: k * B for k = 4^j * {1, 2, 3} is never the identity, and the
  // operands of every addition below are distinct
  std::vector<Jac> pts(3 * n_windows);
  Jac b{base.x, base.y, q_small(1)};
  for (size_t j = 0; j < n_windows; j++) {
    Jac b2 = jac_dbl(b), b3 = jac_add(b2, b);
    pts[3 * j] = b;
    pts[3 * j + 1] = b2;
    pts[3 * j + 2] = b3;
    b = jac_dbl(b2);
  }
  std::vector<Fq> zi(pts.size());
  for (size_t i = 0; i < pts.size(); i++) zi[i] = pts[i].Z;
  q_batch_inv(zi);
  std::vector<Fq> xi(pts.size());
  for (size_t i = 0; i < pts.size(); i++) {
    Fq z2 = q_mul(zi[i], zi[i]);
    Aff a{q_mul(pts[i].X, z2), q_mul(pts[i].Y, q_mul(z2, zi[i])), false};
    out[i / 3].p[i % 3] = a;
    xi[i] = a.x;
  }
  q_batch_inv(xi);
  for (size_t i = 0; i < pts.size(); i++) out[i / 3].lam[i % 3] = q_mul(out[i / 3].p[i % 3].y, xi[i]);
  return out;
}

}  // namespace wit
}  // namespace h2agg

using namespace h2agg;
using namespace h2agg::wit;

struct h2agg_witness {
  Recorder rec;
  std::string err;
};

static bool is_identity_affine(const uint64_t* xy) {
  for (int i = 0; i < 8; i++)
    if (xy[i]) return false;
  return true;
}
static Fq fq_from_mont_words(const uint64_t* w) { return Fq{{w[0], w[1], w[2], w[3]}}; }

#define WIT_TRY(w, ...)               \
  try {                               \
    __VA_ARGS__;                      \
  } catch (const std::exception& e) { \
    (w)->err = e.what();              \
    return -1;                        \
  }

extern "C" {

h2agg_witness* h2agg_wit_new(void) { return new h2agg_witness(); }
int h2agg_wit_set_threads(int n) {
  const int prev = (int)wit_threads();
  if (n >= 1) g_wit_threads.store((unsigned)(n > 256 ? 256 : n));
  return prev;
}
void h2agg_wit_free(h2agg_witness* w) { delete w; }
const char* h2agg_wit_error(h2agg_witness* w) { return w ? w->err.c_str() : ""; }
uint64_t h2agg_wit_rows(h2agg_witness* w) { return w->rec.offset; }
uint64_t h2agg_wit_ops(h2agg_witness* w) { w->rec.flush_raw(); return w->rec.ops.size(); }

// points are affine Montgomery (x, y), (0,0) = identity; returns a point handle
int64_t h2agg_wit_assign_point(h2agg_witness* w, const uint64_t xy[8]) {
  WIT_TRY(w, {
    w->rec.points.push_back(w->rec.assign_point(is_identity_affine(xy), fq_from_mont_words(xy), fq_from_mont_words(xy + 4)));
  });
  return (int64_t)w->rec.points.size() - 1;
}
int64_t h2agg_wit_assign_constant_point(h2agg_witness* w, const uint64_t xy[8]) {
  WIT_TRY(w, {
    w->rec.points.push_back(w->rec.assign_constant_point(is_identity_affine(xy), fq_from_mont_words(xy), fq_from_mont_words(xy + 4)));
  });
  return (int64_t)w->rec.points.size() - 1;
}
// scalar: Montgomery Fr (4 limbs) as the chips hold it; assigned with BaseGateOps::assign (1 row)
int64_t h2agg_wit_assign_scalar(h2agg_witness* w, const uint64_t s_canonical[4]) {
  if (!s_canonical || !fr::canonical(s_canonical)) {
    w->err = "assign_scalar: NULL or not a canonical value < r";
    return -1;
  }
  HScalar s;
  memcpy(s.v, s_canonical, 32);
  uint64_t cells[5][4];
  memset(cells, 0, sizeof(cells));
  memcpy(cells[0], s_canonical, 32);
  s.cell = Cell{w->rec.raw256_row(cells), 0};
  w->rec.scalars.push_back(s);
  return (int64_t)w->rec.scalars.size() - 1;
}
// The trait-level operations take their operands the way the reference's adapter does
// (halo2-snark-aggregator-circuit/src/chips/ecc_chip.rs:34-52, 99-131): `add` runs on `a.clone()` and `b.clone()`,
// `sub` on `a.clone()`, `scalar_mul` on `rhs.clone()`, `normalize` on `v.clone()`, `multi_exp` on a moved Vec.  The
// curvature / native caches an operation fills in are therefore DROPPED with the clone: a handle keeps the caches it had
// when it was created, and a later operation on it pays for the curvature rows again -- part of the row layout.
int64_t h2agg_wit_ecc_add(h2agg_witness* w, int64_t a, int64_t b) {
  WIT_TRY(w, {
    HPoint aa = w->rec.points.at(a);
    HPoint bb = w->rec.points.at(b);
    HPoint r = w->rec.ecc_add(aa, bb);
    w->rec.points.push_back(r);
  });
  return (int64_t)w->rec.points.size() - 1;
}
int64_t h2agg_wit_ecc_sub(h2agg_witness* w, int64_t a, int64_t b) {
  WIT_TRY(w, {
    HPoint aa = w->rec.points.at(a);
    HPoint bb = w->rec.points.at(b);
    HPoint r = w->rec.ecc_sub(aa, bb);
    w->rec.points.push_back(r);
  });
  return (int64_t)w->rec.points.size() - 1;
}
int64_t h2agg_wit_ecc_double(h2agg_witness* w, int64_t a) {
  WIT_TRY(w, {
    HPoint aa = w->rec.points.at(a);
    HPoint r = w->rec.ecc_double(aa);
    w->rec.points.push_back(r);
  });
  return (int64_t)w->rec.points.size() - 1;
}
int64_t h2agg_wit_ecc_reduce(h2agg_witness* w, int64_t a) {
  WIT_TRY(w, {
    HPoint aa = w->rec.points.at(a);
    HPoint r = w->rec.ecc_reduce(aa);
    w->rec.points.push_back(r);
  });
  return (int64_t)w->rec.points.size() - 1;
}
// ArithEccChip::scalar_mul -> EccChipOps::mul
int64_t h2agg_wit_ecc_mul(h2agg_witness* w, int64_t a, int64_t s) {
  WIT_TRY(w, {
    HPoint aa = w->rec.points.at(a);
    HPoint r = w->rec.ecc_mul(aa, w->rec.scalars.at(s));
    w->rec.points.push_back(r);
  });
  return (int64_t)w->rec.points.size() - 1;
}
// ArithEccChip::multi_exp -> EccChipOps::shamir (halo2-snark-aggregator-circuit/src/chips/ecc_chip.rs:125-132)
int64_t h2agg_wit_ecc_shamir(h2agg_witness* w, const int64_t* pts, const int64_t* scalars, size_t n) {
  if (n == 0 || !pts || !scalars) {  // the reference indexes windows_in_be[0]: a multi_exp of nothing panics there
    w->err = "multi_exp: no points";
    return -1;
  }
  WIT_TRY(w, {
    std::vector<HPoint> p;
    std::vector<HScalar> s;
    for (size_t i = 0; i < n; i++) {
      p.push_back(w->rec.points.at(pts[i]));
      s.push_back(w->rec.scalars.at(scalars[i]));
    }
    HPoint r = w->rec.ecc_shamir(p, s);
    w->rec.points.push_back(r);
  });
  return (int64_t)w->rec.points.size() - 1;
}
// ArithEccChip::scalar_mul_constant -> EccChipOps::constant_mul (2-bit windows over constant multiples)
int64_t h2agg_wit_ecc_constant_mul(h2agg_witness* w, const uint64_t base_xy[8], int64_t s) {
  WIT_TRY(w, {
    Recorder& r = w->rec;
    auto bits_be = r.decompose_scalar(r.scalars.at(s), 2);
    HPoint identity = r.assign_constant_point_with_curvature(true, q_zero(), q_zero());
    Aff base{fq_from_mont_words(base_xy), fq_from_mont_words(base_xy + 4), is_identity_affine(base_xy)};
    const std::vector<ConstWindow> table = constant_mul_table(base, bits_be.size());   // all inversions batched
    bool have = false;
    HPoint acc;
    size_t j = 0;
    for (auto it = bits_be.rbegin(); it != bits_be.rend(); ++it, ++j) {
      const ConstWindow& cw = table[j];   // B_j, 2 B_j, 3 B_j with B_(j+1) = 4 B_j
      HPoint c01 = r.assign_constant_point_with_curvature(cw.p[1].inf, cw.p[1].x, cw.p[1].y, &cw.lam[1]);
      HPoint c10 = r.assign_constant_point_with_curvature(cw.p[0].inf, cw.p[0].x, cw.p[0].y, &cw.lam[0]);
      HPoint c11 = r.assign_constant_point_with_curvature(cw.p[2].inf, cw.p[2].x, cw.p[2].y, &cw.lam[2]);
      HPoint c0 = r.bisec_point_with_curvature((*it)[0], c10, identity);
      HPoint c1 = r.bisec_point_with_curvature((*it)[0], c11, c01);
      HPoint slot = r.bisec_point_with_curvature((*it)[1], c1, c0);
      if (!have) { acc = slot; have = true; }
      else acc = r.ecc_add(slot, acc);
    }
    r.points.push_back(acc);
  });
  return (int64_t)w->rec.points.size() - 1;
}

// ---- ArithFieldChip: ScalarChip over the base gate (halo2-snark-aggregator-circuit/src/chips/scalar_chip.rs:17-127,
//      trait at halo2-snark-aggregator-api/src/arith/field.rs:6-105, common.rs:3-42).  Scalars are CANONICAL < r. -----
static bool scalar_arg_ok(h2agg_witness* w, const uint64_t* v) {
  if (!v || !fr::canonical(v)) {
    w->err = "scalar argument is NULL or not a canonical value < r";
    return false;
  }
  return true;
}
static int64_t push_scalar(h2agg_witness* w, const HScalar& s) {
  w->rec.scalars.push_back(s);
  return (int64_t)w->rec.scalars.size() - 1;
}
int64_t h2agg_wit_field_assign_const(h2agg_witness* w, const uint64_t c_canonical[4]) {  // assign_const / _zero / _one
  if (!scalar_arg_ok(w, c_canonical)) return -1;
  return push_scalar(w, w->rec.sc_assign(fr::V{{c_canonical[0], c_canonical[1], c_canonical[2], c_canonical[3]}}));
}
int64_t h2agg_wit_field_add(h2agg_witness* w, int64_t a, int64_t b) {
  WIT_TRY(w, { return push_scalar(w, w->rec.sc_add(w->rec.scalars.at(a), w->rec.scalars.at(b))); });
}
int64_t h2agg_wit_field_sub(h2agg_witness* w, int64_t a, int64_t b) {
  WIT_TRY(w, { return push_scalar(w, w->rec.sc_sub(w->rec.scalars.at(a), w->rec.scalars.at(b))); });
}
int64_t h2agg_wit_field_mul(h2agg_witness* w, int64_t a, int64_t b) {
  WIT_TRY(w, { return push_scalar(w, w->rec.sc_mul(w->rec.scalars.at(a), w->rec.scalars.at(b))); });
}
int64_t h2agg_wit_field_square(h2agg_witness* w, int64_t a) { return h2agg_wit_field_mul(w, a, a); }  // scalar_chip.rs:101-107
int64_t h2agg_wit_field_div(h2agg_witness* w, int64_t a, int64_t b) {  // div_unsafe
  WIT_TRY(w, { return push_scalar(w, w->rec.sc_div_unsafe(w->rec.scalars.at(a), w->rec.scalars.at(b))); });
}
int64_t h2agg_wit_field_sum_with_coeff_and_constant(h2agg_witness* w, const int64_t* elems, const uint64_t* coeffs_canonical,
                                                    size_t n, const uint64_t constant_canonical[4]) {
  if (!scalar_arg_ok(w, constant_canonical)) return -1;
  WIT_TRY(w, {
    std::vector<std::pair<HScalar, fr::V>> e;
    for (size_t i = 0; i < n; i++) {
      if (!scalar_arg_ok(w, coeffs_canonical + 4 * i)) return -1;
      e.push_back({w->rec.scalars.at(elems[i]), fr::V{{coeffs_canonical[4 * i], coeffs_canonical[4 * i + 1], coeffs_canonical[4 * i + 2], coeffs_canonical[4 * i + 3]}}});
    }
    return push_scalar(w, w->rec.sc_sum_with_constant(e, fr::V{{constant_canonical[0], constant_canonical[1], constant_canonical[2], constant_canonical[3]}}));
  });
}
int64_t h2agg_wit_field_mul_add_constant(h2agg_witness* w, int64_t a, int64_t b, const uint64_t c_canonical[4]) {
  if (!scalar_arg_ok(w, c_canonical)) return -1;
  WIT_TRY(w, {
    return push_scalar(w, w->rec.sc_mul_add_constant(w->rec.scalars.at(a), w->rec.scalars.at(b),
                                                     fr::V{{c_canonical[0], c_canonical[1], c_canonical[2], c_canonical[3]}}));
  });
}
int h2agg_wit_scalar_value(h2agg_witness* w, int64_t h, uint64_t out_canonical[4]) {  // to_value
  if (h < 0 || (size_t)h >= w->rec.scalars.size()) return -1;
  memcpy(out_canonical, w->rec.scalars[h].v, 32);
  return 0;
}
int h2agg_wit_scalar_cell(h2agg_witness* w, int64_t h, uint32_t* column, uint32_t* row) {  // AssignedValue.cell
  if (h < 0 || (size_t)h >= w->rec.scalars.size()) return -1;
  *column = w->rec.scalars[h].cell.col;
  *row = w->rec.scalars[h].cell.row;
  return 0;
}

// ---- Encode: PoseidonEncodeChip (halo2-snark-aggregator-circuit/src/chips/encode_chip.rs:14-51) --------------------
int h2agg_wit_encode_point(h2agg_witness* w, int64_t point, int64_t out_natives[2]) {
  WIT_TRY(w, {
    HScalar n2[2];
    w->rec.encode_point(w->rec.points.at(point), n2);
    out_natives[0] = push_scalar(w, n2[0]);
    out_natives[1] = push_scalar(w, n2[1]);
  });
  return 0;
}

// ---- what Halo2VerifierCircuits::synthesize does around the chips (verify_circuit.rs:264-368, 487-496) ---------------
int64_t h2agg_wit_ecc_assign_identity(h2agg_witness* w) {  // ArithCommonChip::assign_zero of the EccChip
  WIT_TRY(w, { w->rec.points.push_back(w->rec.assign_identity()); });
  return (int64_t)w->rec.points.size() - 1;
}
int h2agg_wit_ecc_assert_equal(h2agg_witness* w, int64_t a, int64_t b) {
  WIT_TRY(w, {
    HPoint aa = w->rec.points.at(a);   // `&mut commits[..].clone()` on the left, in place on the right (:488-492)
    w->rec.ecc_assert_equal(aa, w->rec.points.at(b));
  });
  return 0;
}
int h2agg_wit_assert_not_identity(h2agg_witness* w, int64_t point) {  // base_gate.assert_false(&p.z)
  WIT_TRY(w, {
    const HPoint& p = w->rec.points.at(point);
    if (p.z.value) throw std::runtime_error("assert_false: the point is the identity");
    w->rec.bg_assert_constant(p.z);
  });
  return 0;
}
int h2agg_wit_expose_final_pair(h2agg_witness* w, int64_t w_x, int64_t w_g, int64_t out_cells[4]) {
  WIT_TRY(w, {
    HPoint p[2] = {w->rec.points.at(w_x), w->rec.points.at(w_g)};
    HScalar out[4];
    w->rec.expose_final_pair(p, out);
    for (int i = 0; i < 4; i++) out_cells[i] = push_scalar(w, out[i]);
  });
  return 0;
}
// value of a point handle: canonical affine (x mod p, y mod p) as Montgomery limbs + identity flag
int h2agg_wit_point_value(h2agg_witness* w, int64_t h, uint64_t out_xy[8], int* is_identity) {
  if (h < 0 || (size_t)h >= w->rec.points.size()) return -1;
  const HPoint& p = w->rec.points[h];
  memcpy(out_xy, p.x.w.v, 32);
  memcpy(out_xy + 4, p.y.w.v, 32);
  *is_identity = (int)p.z.value;
  return 0;
}

}  // extern "C"

// record chunks -> one contiguous device array in arrival order (each record names its own rows; the device groups
// them by opcode through an index array, witness.cu)
static int upload_ops(h2agg_ctx* ctx, const OpStore& ops) {
  size_t at = 0;
  for (const OpChunk& c : ops.chunks) {
    if (!c.n) continue;
    H2AGG_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ctx->io_a.p + at * sizeof(WitnessOp), c.p, (size_t)c.n * sizeof(WitnessOp), cudaMemcpyHostToDevice, ctx->stream));
    at += c.n;
  }
  return 0;
}

extern "C" {

// Run the expansion kernel over everything recorded so far: 5 advice columns of n_rows Fr each
// (host pointers; rows beyond the recorded offset stay zero like unassigned halo2 cells).
int h2agg_witness_expand(h2agg_ctx* ctx, h2agg_witness* w, uint64_t* const advice_cols[5], size_t n_rows) {
  if (!ctx || !w) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  w->rec.flush_raw();
  if (n_rows < w->rec.offset) {
    ctx->last_error = "witness_expand: n_rows is smaller than the recorded layout";
    return 1;
  }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n_ops = w->rec.ops.size();
  int rc = ensure(ctx, ctx->io_a, n_ops * sizeof(WitnessOp) + 256);
  if (rc) return rc;
  rc = ensure(ctx, ctx->io_b, n_rows * 32 * 5);
  if (rc) return rc;
  rc = upload_ops(ctx, w->rec.ops);
  if (rc) return rc;
  void* cols[5];
  for (int c = 0; c < 5; c++) cols[c] = (uint8_t*)ctx->io_b.p + (size_t)c * n_rows * 32;
  rc = witness_expand_dev(ctx, ctx->io_a.p, w->rec.ops.count, cols, n_rows);
  if (rc) return rc;
  for (int c = 0; c < 5; c++) H2AGG_CUDA(ctx, cudaMemcpyAsync(advice_cols[c], cols[c], n_rows * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2AGG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Same, leaving the columns in HBM (they feed commit_lagrange directly): d_cols = 5 device pointers.
int h2agg_witness_expand_dev(h2agg_ctx* ctx, h2agg_witness* w, void* const d_cols[5], size_t n_rows) {
  if (!ctx || !w) return 1;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  w->rec.flush_raw();
  if (n_rows < w->rec.offset) {
    ctx->last_error = "witness_expand: n_rows is smaller than the recorded layout";
    return 1;
  }
  H2AGG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n_ops = w->rec.ops.size();
  int rc = ensure(ctx, ctx->io_a, n_ops * sizeof(WitnessOp) + 256);
  if (rc) return rc;
  rc = upload_ops(ctx, w->rec.ops);
  if (rc) return rc;
  return witness_expand_dev(ctx, ctx->io_a.p, w->rec.ops.count, d_cols, n_rows);
}

}  // extern "C"
