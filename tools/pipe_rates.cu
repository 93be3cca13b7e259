// Issue-rate microbenchmark for the instruction kinds a 256-bit modular multiplication can be built from on
// sm_100a: which pipe is the ceiling, and can two of them be kept busy at once?  (DESIGN.md "multiplier roofline")
// Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu && ./pipe_rates
// Every loop body was checked with cuobjdump -sass to be the instruction named (8 independent chains per thread).
// Combined mixes are WARP-SPECIALISED (even warps run one loop, odd warps the other), so both loops stay clean and
// the SM sees the two instruction streams side by side: if the pipes are independent the pair finishes in
// max(t_a, t_b) / 2 of the single-mix times, if they share an issue port in (t_a + t_b) / 2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

enum Mix { IMAD_WIDE, IMAD_LO, IMAD_HI, DFMA, DFMA_RZ, IADD64, LOP3, FFMA, N_SINGLE };
static const char* NAMES[] = {"IMAD.WIDE.U32 (64-bit accumulate)", "IMAD (mad.lo.u32)", "IMAD.HI.U32", "DFMA", "DFMA.RZ",
                              "IADD3 + IADD3.X/IMAD.X (64-bit add = 2 instr)", "LOP3", "FFMA"};

template <int MIX>
__device__ __forceinline__ uint64_t body(uint32_t a, uint32_t b, double da, double db) {
  uint64_t acc[NACC];
  double dacc[NACC];
  uint32_t lo[NACC], hi[NACC], m[NACC];
  float facc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) {
    acc[i] = threadIdx.x + i;
    dacc[i] = (double)(threadIdx.x + i);
    lo[i] = threadIdx.x * 3 + i;
    hi[i] = threadIdx.x * 5 + i;
    m[i] = a * (2 * i + 3) + threadIdx.x;  // per-accumulator multiplicand: no product is shared
    facc[i] = (float)(threadIdx.x + i);
  }
  float fa = (float)da, fb = (float)db;
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    a += b;  // loop-variant multiplier: no product can be hoisted
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (MIX == IMAD_WIDE) acc[i] = (uint64_t)m[i] * a + acc[i];
      if (MIX == IMAD_LO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(a), "r"(b));
      if (MIX == IMAD_HI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(a), "r"(b));
      if (MIX == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dacc[i]) : "d"(da), "d"(db));
      if (MIX == DFMA_RZ) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(dacc[i]) : "d"(da), "d"(db));
      if (MIX == IADD64) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %0;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a));
      if (MIX == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[i]) : "r"(a), "r"(b));
      if (MIX == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(facc[i]) : "f"(fa), "f"(fb));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i] + (uint64_t)__double_as_longlong(dacc[i]) + lo[i] + hi[i] + (uint64_t)__float_as_uint(facc[i]);
  return s;
}

template <int MIX>
__global__ void __launch_bounds__(256) rate_kernel(uint64_t* out, uint32_t a0, uint32_t b0, double da, double db, long long* cycles) {
  long long t0 = clock64();
  uint64_t s = body<MIX>(a0 + threadIdx.x, b0 + blockIdx.x, da, db);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MA, int MB>
__global__ void __launch_bounds__(256) pair_kernel(uint64_t* out, uint32_t a0, uint32_t b0, double da, double db, long long* cycles) {
  long long t0 = clock64();
  uint64_t s;
  if ((threadIdx.x >> 5) & 1) s = body<MB>(a0 + threadIdx.x, b0 + blockIdx.x, da, db);
  else s = body<MA>(a0 + threadIdx.x, b0 + blockIdx.x, da, db);
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static const int BLOCKS_PER_SM = 4, THREADS = 256;  // 32 warps per SM: every scheduler has 8 warps to pick from

template <class K>
static double launch(K kern, int sms, uint64_t* d_out, long long* d_cyc, long long* h_cyc, float* ms) {
  const int grid = sms * BLOCKS_PER_SM;
  kern<<<grid, THREADS>>>(d_out, 12345u, 678u, 1.000000001, 0.999999999, d_cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<grid, THREADS>>>(d_out, 12345u, 678u, 1.000000001, 0.999999999, d_cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(ms, e0, e1);
  cudaMemcpy(h_cyc, d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double cyc = 0;
  for (int i = 0; i < grid; i++) cyc += (double)h_cyc[i];
  return cyc / grid;
}

static double g_single_cyc[N_SINGLE];

template <int MIX>
static void run(int sms, uint64_t* d_out, long long* d_cyc, long long* h_cyc) {
  float ms;
  double cyc = launch(rate_kernel<MIX>, sms, d_out, d_cyc, h_cyc, &ms);
  g_single_cyc[MIX] = cyc;
  // all BLOCKS_PER_SM blocks of an SM are resident together, so the SM retired this many thread-level ops in `cyc` clocks
  double ops_sm = (double)BLOCKS_PER_SM * THREADS * ITERS * NACC;
  printf("{\"mix\": \"%s\", \"thread_ops_per_clk_per_sm\": %.2f, \"cycles\": %.0f, \"ms\": %.4f, \"G_ops_per_s\": %.1f}\n", NAMES[MIX],
         ops_sm / cyc, cyc, ms, ops_sm * sms / ms * 1e-6);
}

template <int MA, int MB>
static void run_pair(int sms, uint64_t* d_out, long long* d_cyc, long long* h_cyc) {
  float ms;
  double cyc = launch(pair_kernel<MA, MB>, sms, d_out, d_cyc, h_cyc, &ms);
  double ta = g_single_cyc[MA] / 2, tb = g_single_cyc[MB] / 2;  // half the warps run each loop
  printf("{\"pair\": \"%s | %s\", \"cycles\": %.0f, \"independent_pipes_would_be\": %.0f, \"shared_port_would_be\": %.0f, \"ms\": %.4f}\n",
         NAMES[MA], NAMES[MB], cyc, ta > tb ? ta : tb, ta + tb, ms);
}

int main() {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { fprintf(stderr, "no device\n"); return 1; }
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  uint64_t* d_out;
  long long *d_cyc, *h_cyc;
  cudaMalloc(&d_out, (size_t)sms * BLOCKS_PER_SM * THREADS * 8);
  cudaMalloc(&d_cyc, (size_t)sms * BLOCKS_PER_SM * 8);
  h_cyc = (long long*)malloc((size_t)sms * BLOCKS_PER_SM * 8);
  run<IMAD_WIDE>(sms, d_out, d_cyc, h_cyc);
  run<IMAD_LO>(sms, d_out, d_cyc, h_cyc);
  run<IMAD_HI>(sms, d_out, d_cyc, h_cyc);
  run<DFMA>(sms, d_out, d_cyc, h_cyc);
  run<DFMA_RZ>(sms, d_out, d_cyc, h_cyc);
  run<IADD64>(sms, d_out, d_cyc, h_cyc);
  run<LOP3>(sms, d_out, d_cyc, h_cyc);
  run<FFMA>(sms, d_out, d_cyc, h_cyc);
  run_pair<IMAD_WIDE, DFMA>(sms, d_out, d_cyc, h_cyc);
  run_pair<IMAD_WIDE, IADD64>(sms, d_out, d_cyc, h_cyc);
  run_pair<DFMA, IADD64>(sms, d_out, d_cyc, h_cyc);
  run_pair<IMAD_WIDE, FFMA>(sms, d_out, d_cyc, h_cyc);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 2;
}
