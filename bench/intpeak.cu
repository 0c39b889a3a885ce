// intpeak.cu -- integer / DPX issue-rate microbenchmark for sm_100a.
//
// Measures L = sustained lane-instructions per clock per SM for the instruction
// classes the Smith-Waterman/Gotoh recurrence is made of (SURVEY.md section 8d:
// GCUPS_roof = N_SM * f_clk * L * V / 7).  Every kernel runs ILP independent
// dependency chains per thread; we sweep warps/SM so both the latency-bound and
// the throughput-bound regime are visible.  Output: one JSON object per line.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o intpeak intpeak.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

enum Op {
  OP_VIMNMX3_S16X2_RELU = 0, OP_VIADDMNMX_S16X2, OP_VIADD_16X2, OP_VIMNMX_S16X2, OP_PRMT,
  OP_LOP3, OP_IMAD, OP_IADD32, OP_VIMNMX3_S32_RELU, OP_VIADDMNMX_S32, OP_HFMA2, OP_HMNMX2,
  OP_MIX_VIMNMX3_IMAD, OP_MIX_VIMNMX3_VIADD16, OP_MIX_VIMNMX3_PRMT, OP_MIX_VIADDMNMX_VIADD16,
  OP_MIX_VIMNMX3_HFMA2, OP_MIX_VIADD16_IMAD, OP_MIX_PRMT_IMAD, OP_MIX_VIADDMNMX_IMAD,
  OP_SHFL, OP_LDS, OP_CELL_AFFINE, OP_CELL_LINEAR, OP_CELL_AFFINE_S32,
  OP_VIMNMX_S16X2_RELU, OP_VIADDMNMX_S16X2_RELU, OP_MIX_VIMNMX_VIMNMX3, OP_MIX_VIMNMX_VIADD16, OP_MIX_VIMNMX_PRMT,
  OP_MIX_ROW_LINEAR, OP_MIX_A3_B1, OP_MIX_A2_MNMX2, OP_COUNT
};

static const char* op_name[OP_COUNT] = {
  "VIMNMX3.S16x2.RELU", "VIADDMNMX.S16x2", "VIADD.16x2", "VIMNMX.S16x2", "PRMT",
  "LOP3", "IMAD", "IADD32", "VIMNMX3.S32.RELU", "VIADDMNMX.S32", "HFMA2", "HMNMX2",
  "mix VIMNMX3+IMAD", "mix VIMNMX3+VIADD16", "mix VIMNMX3+PRMT", "mix VIADDMNMX+VIADD16",
  "mix VIMNMX3+HFMA2", "mix VIADD16+IMAD", "mix PRMT+IMAD", "mix VIADDMNMX+IMAD",
  "SHFL", "LDS", "cell affine s16x2 (6.5 instr)", "cell linear s16x2 (4.5 instr)", "cell affine s32 (6.5 instr)",
  "VIMNMX.S16x2.RELU", "VIADDMNMX.S16x2.RELU", "mix VIMNMX+VIMNMX3", "mix VIMNMX+VIADD16", "mix VIMNMX+PRMT",
  "mix PRMT+VIADDMNMX.RELU+VIMNMX+VIADD16 (linear row)", "mix PRMT+VIADDMNMX+VIMNMX3+VIADD16", "mix PRMT+VIADDMNMX+2xVIMNMX"
};

__device__ __forceinline__ uint32_t hfma2_u(uint32_t x, uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t hmnmx2_u(uint32_t x, uint32_t a) {
  uint32_t r; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a)); return r;
}
__device__ __forceinline__ uint32_t imad_u(uint32_t x, uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t lop3_u(uint32_t x, uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t iadd_u(uint32_t x, uint32_t a) {
  uint32_t r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a)); return r;
}

// one "op call" on chain k; returns new chain value.  instr_per_call[] below must match.
template <int OP>
__device__ __forceinline__ uint32_t apply(uint32_t x, uint32_t a, uint32_t b, int k, uint32_t* sm) {
  if constexpr (OP == OP_VIMNMX3_S16X2_RELU) return __vimax3_s16x2_relu(x, a, b);
  else if constexpr (OP == OP_VIADDMNMX_S16X2) return __viaddmax_s16x2(x, a, b);
  else if constexpr (OP == OP_VIADD_16X2) return __vadd2(x, a);
  else if constexpr (OP == OP_VIMNMX_S16X2) return __vmaxs2(x, a);
  else if constexpr (OP == OP_PRMT) return __byte_perm(x, a, b);
  else if constexpr (OP == OP_LOP3) return lop3_u(x, a, b);
  else if constexpr (OP == OP_IMAD) return imad_u(x, a, b);
  else if constexpr (OP == OP_IADD32) return iadd_u(x, a);
  else if constexpr (OP == OP_VIMNMX3_S32_RELU) return (uint32_t)__vimax3_s32_relu((int)x, (int)a, (int)b);
  else if constexpr (OP == OP_VIADDMNMX_S32) return (uint32_t)__viaddmax_s32((int)x, (int)a, (int)b);
  else if constexpr (OP == OP_HFMA2) return hfma2_u(x, a, b);
  else if constexpr (OP == OP_HMNMX2) return hmnmx2_u(x, a);
  else if constexpr (OP == OP_MIX_VIMNMX3_IMAD) return (k & 1) ? imad_u(x, a, b) : __vimax3_s16x2_relu(x, a, b);
  else if constexpr (OP == OP_MIX_VIMNMX3_VIADD16) return (k & 1) ? __vadd2(x, a) : __vimax3_s16x2_relu(x, a, b);
  else if constexpr (OP == OP_MIX_VIMNMX3_PRMT) return (k & 1) ? __byte_perm(x, a, b) : __vimax3_s16x2_relu(x, a, b);
  else if constexpr (OP == OP_MIX_VIADDMNMX_VIADD16) return (k & 1) ? __vadd2(x, a) : __viaddmax_s16x2(x, a, b);
  else if constexpr (OP == OP_MIX_VIMNMX3_HFMA2) return (k & 1) ? hfma2_u(x, a, b) : __vimax3_s16x2_relu(x, a, b);
  else if constexpr (OP == OP_MIX_VIADD16_IMAD) return (k & 1) ? imad_u(x, a, b) : __vadd2(x, a);
  else if constexpr (OP == OP_MIX_PRMT_IMAD) return (k & 1) ? imad_u(x, a, b) : __byte_perm(x, a, b);
  else if constexpr (OP == OP_MIX_VIADDMNMX_IMAD) return (k & 1) ? imad_u(x, a, b) : __viaddmax_s16x2(x, a, b);
  else if constexpr (OP == OP_VIMNMX_S16X2_RELU) return __vimax_s16x2_relu(x, a);
  else if constexpr (OP == OP_VIADDMNMX_S16X2_RELU) return __viaddmax_s16x2_relu(x, a, b);
  else if constexpr (OP == OP_MIX_VIMNMX_VIMNMX3) return (k & 1) ? __vmaxs2(x, a) : __vimax3_s16x2_relu(x, a, b);
  else if constexpr (OP == OP_MIX_VIMNMX_VIADD16) return (k & 1) ? __vmaxs2(x, a) : __vadd2(x, a);
  else if constexpr (OP == OP_MIX_VIMNMX_PRMT) return (k & 1) ? __vmaxs2(x, a) : __byte_perm(x, a, b);
  else if constexpr (OP == OP_MIX_ROW_LINEAR)
    return (k & 3) == 0 ? __byte_perm(x, a, b) : ((k & 3) == 1 ? __viaddmax_s16x2_relu(x, a, b) : ((k & 3) == 2 ? __vmaxs2(x, a) : __vadd2(x, a)));
  else if constexpr (OP == OP_MIX_A3_B1)
    return (k & 3) == 0 ? __byte_perm(x, a, b) : ((k & 3) == 1 ? __viaddmax_s16x2(x, a, b) : ((k & 3) == 2 ? __vimax3_s16x2_relu(x, a, b) : __vadd2(x, a)));
  else if constexpr (OP == OP_MIX_A2_MNMX2)
    return (k & 3) == 0 ? __byte_perm(x, a, b) : ((k & 3) == 1 ? __viaddmax_s16x2(x, a, b) : __vmaxs2(x, a));
  else if constexpr (OP == OP_SHFL) return __shfl_sync(0xffffffffu, x, (int)(a & 31));
  else if constexpr (OP == OP_LDS) return sm[(x + a) & 1023];
  else return x;
}

template <int OP, int ILP>
__global__ void __launch_bounds__(1024) k_chain(uint32_t* out, uint32_t a, uint32_t b, int iters, long long* cyc) {
  __shared__ uint32_t sm[1024];
  sm[threadIdx.x & 1023] = threadIdx.x * 7u & 1023u;
  __syncthreads();
  uint32_t x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 0x10001u + k;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int k = 0; k < ILP; ++k) x[k] = apply<OP>(x[k], a, b, k, sm);
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s ^= x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t0; cyc[2 * blockIdx.x + 1] = t1; }
}

// Realistic recurrence bodies: CH independent cell-vector chains per thread, each chain is
// a column of cells updated once per "step" (H,E state in registers, F carried down the chain).
template <int MODE, int CH>
__global__ void __launch_bounds__(1024) k_cell(uint32_t* out, uint32_t tlo, uint32_t thi, uint32_t nopen, uint32_t next,
                                               int iters, long long* cyc) {
  uint32_t Ho[CH], E[CH], sel[CH];
  uint32_t best = 0, F = 0, diag = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) { Ho[k] = 0; E[k] = 0; sel[k] = out[(threadIdx.x + k) & 1023]; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t t1v = tlo + (uint32_t)(it * 4 + u) * 0x01010101u, t2v = thi ^ (uint32_t)(it * 4 + u);
      F = diag;  // top input of this step (register move)
      uint32_t d_in = diag;
#pragma unroll
      for (int k = 0; k < CH; k += 2) {
        uint32_t h0, h1;
        {
          uint32_t s = __byte_perm(t1v, t2v, sel[k]);
          uint32_t old = Ho[k];
          if (MODE == 2) {
            int d = (int)d_in + (int)s;
            E[k] = (uint32_t)__viaddmax_s32((int)E[k], (int)next, (int)old);
            F = (uint32_t)__viaddmax_s32((int)F, (int)next, (int)(k ? Ho[k - 1] : d_in));
            h0 = (uint32_t)__vimax3_s32_relu(d, (int)E[k], (int)F);
            Ho[k] = h0 + nopen;
          } else if (MODE == 0) {
            uint32_t d = __vadd2(d_in, s);
            E[k] = __viaddmax_s16x2(E[k], next, old);
            F = __viaddmax_s16x2(F, next, k ? Ho[k - 1] : d_in);
            h0 = __vimax3_s16x2_relu(d, E[k], F);
            Ho[k] = __vadd2(h0, nopen);
          } else {
            uint32_t d = __vadd2(d_in, s);
            h0 = __vimax3_s16x2_relu(d, old, k ? Ho[k - 1] : d_in);
            Ho[k] = __vadd2(h0, nopen);
          }
          d_in = old;
        }
        {
          uint32_t s = __byte_perm(t1v, t2v, sel[k + 1]);
          uint32_t old = Ho[k + 1];
          if (MODE == 2) {
            int d = (int)d_in + (int)s;
            E[k + 1] = (uint32_t)__viaddmax_s32((int)E[k + 1], (int)next, (int)old);
            F = (uint32_t)__viaddmax_s32((int)F, (int)next, (int)Ho[k]);
            h1 = (uint32_t)__vimax3_s32_relu(d, (int)E[k + 1], (int)F);
            Ho[k + 1] = h1 + nopen;
          } else if (MODE == 0) {
            uint32_t d = __vadd2(d_in, s);
            E[k + 1] = __viaddmax_s16x2(E[k + 1], next, old);
            F = __viaddmax_s16x2(F, next, Ho[k]);
            h1 = __vimax3_s16x2_relu(d, E[k + 1], F);
            Ho[k + 1] = __vadd2(h1, nopen);
          } else {
            uint32_t d = __vadd2(d_in, s);
            h1 = __vimax3_s16x2_relu(d, old, Ho[k]);
            Ho[k + 1] = __vadd2(h1, nopen);
          }
          d_in = old;
        }
        if (MODE == 2) best = (uint32_t)__vimax3_s32((int)best, (int)h0, (int)h1);
        else best = __vimax3_s16x2(best, h0, h1);
      }
      diag = Ho[CH - 1] ^ u;
    }
  }
  long long t1 = clock64();
  uint32_t s = best ^ F;
#pragma unroll
  for (int k = 0; k < CH; ++k) s ^= Ho[k] ^ E[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t0; cyc[2 * blockIdx.x + 1] = t1; }
}

struct Result { double lane_per_clk_sm; double ms; long long cycles; };

static int g_sms = 148;
static uint32_t* d_out; static long long* d_cyc;

template <typename F>
static Result run(F launch, int blocks_per_sm, int threads, double instr_per_thread) {
  int blocks = g_sms * blocks_per_sm;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(blocks, threads); CK(cudaDeviceSynchronize());     // warm-up
  CK(cudaEventRecord(e0)); launch(blocks, threads); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> c(2 * blocks);
  CK(cudaMemcpy(c.data(), d_cyc, sizeof(long long) * 2 * blocks, cudaMemcpyDeviceToHost));
  // per-block duration (clock64 is per-SM); all blocks are co-resident so the median duration is the per-SM time
  std::vector<long long> d(blocks);
  for (int i = 0; i < blocks; ++i) d[i] = c[2 * i + 1] - c[2 * i];
  std::sort(d.begin(), d.end());
  long long cyc = d[blocks / 2];
  Result r; r.cycles = cyc; r.ms = ms;
  r.lane_per_clk_sm = instr_per_thread * threads * blocks_per_sm / (double)cyc;
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  return r;
}

template <int OP, int ILP>
static void bench_chain(FILE* f, int iters) {
  const int cfgs[][2] = {{1, 128}, {1, 256}, {1, 512}, {1, 1024}, {2, 1024}};  // blocks/SM, threads
  for (auto& c : cfgs) {
    double ipt = (double)iters * 16 * ILP;
    Result r = run([&](int b, int t) { k_chain<OP, ILP><<<b, t>>>(d_out, 0x00030001u, 0x00010002u, iters, d_cyc); }, c[0], c[1], ipt);
    fprintf(f, "{\"kind\":\"chain\",\"op\":\"%s\",\"ilp\":%d,\"warps_per_sm\":%d,\"lane_instr_per_clk_per_sm\":%.2f,\"cycles\":%lld,\"ms\":%.4f}\n",
            op_name[OP], ILP, c[0] * c[1] / 32, r.lane_per_clk_sm, r.cycles, r.ms);
    fflush(f);
  }
}

template <int MODE, int CH>
static void bench_cell(FILE* f, int iters) {
  const int cfgs[][2] = {{1, 128}, {1, 256}, {1, 512}, {1, 1024}};
  for (auto& c : cfgs) {
    double cells_vec = (double)iters * 4 * CH;          // cell-vectors per thread
    double ipc = (MODE == 1) ? 4.5 : 6.5;               // algorithmic instructions per cell-vector in this body
    Result r = run([&](int b, int t) { k_cell<MODE, CH><<<b, t>>>(d_out, 0x01ff01ffu, 0xff01ff01u, 0xffffffffu, 0xffffffffu, iters, d_cyc); },
                   c[0], c[1], cells_vec);
    int op = MODE == 0 ? OP_CELL_AFFINE : (MODE == 1 ? OP_CELL_LINEAR : OP_CELL_AFFINE_S32);
    fprintf(f, "{\"kind\":\"cell\",\"op\":\"%s\",\"rows\":%d,\"warps_per_sm\":%d,\"cellvec_per_clk_per_sm\":%.3f,\"lane_instr_per_clk_per_sm\":%.2f,\"cycles\":%lld,\"ms\":%.4f}\n",
            op_name[op], CH, c[0] * c[1] / 32, r.lane_per_clk_sm, r.lane_per_clk_sm * ipc, r.cycles, r.ms);
    fflush(f);
  }
}

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "gpurun_out/intpeak.jsonl";
  FILE* f = fopen(path, "w");
  if (!f) { f = stdout; }
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  g_sms = p.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  fprintf(f, "{\"kind\":\"device\",\"name\":\"%s\",\"sms\":%d,\"cc\":\"%d.%d\",\"max_clock_mhz\":%d}\n", p.name, g_sms, p.major, p.minor, clk_khz / 1000);
  CK(cudaMalloc(&d_out, sizeof(uint32_t) * g_sms * 2 * 1024));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 2 * g_sms * 2));
  const int IT = 2000;
#define CH1(OP) bench_chain<OP, 1>(f, IT); bench_chain<OP, 2>(f, IT); bench_chain<OP, 8>(f, IT);
#define CH8(OP) bench_chain<OP, 8>(f, IT);
  CH1(OP_VIMNMX3_S16X2_RELU) CH1(OP_VIADDMNMX_S16X2) CH8(OP_VIADD_16X2) CH8(OP_VIMNMX_S16X2) CH8(OP_PRMT)
  CH8(OP_LOP3) CH8(OP_IMAD) CH8(OP_IADD32) CH8(OP_VIMNMX3_S32_RELU) CH8(OP_VIADDMNMX_S32) CH8(OP_HFMA2) CH8(OP_HMNMX2)
  CH8(OP_MIX_VIMNMX3_IMAD) CH8(OP_MIX_VIMNMX3_VIADD16) CH8(OP_MIX_VIMNMX3_PRMT) CH8(OP_MIX_VIADDMNMX_VIADD16)
  CH8(OP_MIX_VIMNMX3_HFMA2) CH8(OP_MIX_VIADD16_IMAD) CH8(OP_MIX_PRMT_IMAD) CH8(OP_MIX_VIADDMNMX_IMAD)
  CH8(OP_SHFL) CH8(OP_LDS)
  CH8(OP_VIMNMX_S16X2_RELU) CH8(OP_VIADDMNMX_S16X2_RELU) CH8(OP_MIX_VIMNMX_VIMNMX3) CH8(OP_MIX_VIMNMX_VIADD16)
  CH8(OP_MIX_VIMNMX_PRMT) CH8(OP_MIX_ROW_LINEAR) CH8(OP_MIX_A3_B1) CH8(OP_MIX_A2_MNMX2)
  bench_cell<0, 2>(f, IT); bench_cell<0, 4>(f, IT); bench_cell<0, 8>(f, IT); bench_cell<0, 16>(f, IT);
  bench_cell<1, 4>(f, IT); bench_cell<1, 8>(f, IT); bench_cell<1, 16>(f, IT);
  bench_cell<2, 4>(f, IT); bench_cell<2, 8>(f, IT);
  if (f != stdout) fclose(f);
  printf("intpeak done -> %s\n", path);
  return 0;
}
