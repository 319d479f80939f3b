// Micro-benchmarks behind DESIGN.md's instruction-bound ceiling: issue rates of the instructions the simulate loop is
// made of (IMAD.WIDE.U32, IMAD.HI, IMAD, LOP3, DFMA) and of Philox4x32-10 itself with 1 and 2 independent blocks per
// thread, on one 1024-thread CTA per SM (the persistent kernel's shape) and with 2 CTAs of 512... Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/ubench_int.bin tools/ubench/ubench_int.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define M0 0xD2511F53u
#define M1 0xCD9E8D57u

template <int OP, int CH>
__global__ void __launch_bounds__(1024) k_op(int iters, uint32_t seed, uint32_t *out) {
  uint32_t x[CH];
  uint64_t p[CH];
  double d[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    x[c] = seed + threadIdx.x * 7919u + c * 104729u + blockIdx.x;
    p[c] = x[c];
    d[c] = 1.0 + 1e-9 * x[c];
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (OP == 0) {  // IMAD.WIDE.U32 with 64-bit accumulate: p = lo(p) * M0 + p
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(p[c]) : "r"((uint32_t)p[c]), "r"(M0));
      } else if (OP == 1) {  // IMAD.HI.U32
        asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(x[c]) : "r"(M0));
      } else if (OP == 2) {  // IMAD (32-bit)
        asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(x[c]) : "r"(M0));
      } else if (OP == 3) {  // LOP3
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(seed), "r"(i));
      } else if (OP == 4) {  // DFMA
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(0.999999), "d"(1e-7));
      } else if (OP == 5) {  // mul.wide (no accumulate) + xor of halves: 1 IMAD.WIDE + 1 LOP3
        uint64_t q;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(q) : "r"(x[c]), "r"(M0));
        x[c] = (uint32_t)q ^ (uint32_t)(q >> 32);
      } else if (OP == 6) {  // IADD3
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(seed));
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc ^= x[c] ^ (uint32_t)p[c] ^ (uint32_t)(p[c] >> 32) ^ (uint32_t)__double_as_longlong(d[c]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
  c0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
  c1 = (uint32_t)p1;
  c2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
  c3 = (uint32_t)p0;
}

template <int ILP, int ROUNDS>
__global__ void __launch_bounds__(1024) k_philox(int iters, uint32_t k0, uint32_t k1, uint32_t *out) {
  uint32_t acc = 0;
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    uint32_t c[ILP][4];
#pragma unroll
    for (int b = 0; b < ILP; ++b) {
      c[b][0] = i * ILP + b;
      c[b][1] = gid;
      c[b][2] = 0;
      c[b][3] = 1u << 28;
    }
    uint32_t a0 = k0, a1 = k1;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
#pragma unroll
      for (int b = 0; b < ILP; ++b) philox_round(c[b][0], c[b][1], c[b][2], c[b][3], a0, a1);
      a0 += 0x9E3779B9u;
      a1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int b = 0; b < ILP; ++b) acc ^= c[b][0] ^ c[b][1] ^ c[b][2] ^ c[b][3];
  }
  out[gid] = acc;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  uint32_t *out;
  cudaMalloc(&out, sizeof(uint32_t) * sms * 4 * 1024);
  printf("device %s, %d SMs, %.0f MHz (nominal max; rates below assume it)\n", pr.name, sms, clk / 1e6);
  const int iters = 4096;
  const char *names[] = {"IMAD.WIDE.U32 (mad.wide, 64-bit acc)", "IMAD.HI.U32", "IMAD (32-bit)", "LOP3", "DFMA", "mul.wide + LOP3", "IADD"};
#define RUN_OP(OP, CH, THREADS, BLOCKS_PER_SM)                                                                          \
  {                                                                                                                     \
    float ms = time_ms([&] { k_op<OP, CH><<<sms * BLOCKS_PER_SM, THREADS>>>(iters, 12345u, out); });                   \
    double winst = (double)iters * CH * (THREADS / 32) * BLOCKS_PER_SM;                                                 \
    double cyc = ms * 1e-3 * clk;                                                                                       \
    printf("%-40s chains=%d warps/SM=%2d: %.3f warp-instr/clk/SMSP  (rt %.2f cycles)\n", names[OP], CH,                  \
           THREADS / 32 * BLOCKS_PER_SM, winst / cyc / 4.0, cyc * 4.0 / winst);                                         \
  }
  RUN_OP(0, 8, 1024, 1) RUN_OP(0, 2, 1024, 1) RUN_OP(1, 8, 1024, 1) RUN_OP(2, 8, 1024, 1) RUN_OP(3, 8, 1024, 1)
  RUN_OP(4, 8, 1024, 1) RUN_OP(5, 8, 1024, 1) RUN_OP(5, 2, 1024, 1) RUN_OP(6, 8, 1024, 1)
#define RUN_PH(ILP, ROUNDS, THREADS, BLOCKS_PER_SM)                                                                     \
  {                                                                                                                     \
    const int it2 = 2048;                                                                                               \
    float ms = time_ms([&] { k_philox<ILP, ROUNDS><<<sms * BLOCKS_PER_SM, THREADS>>>(it2, 1234u, 0u, out); });         \
    double blocks = (double)it2 * ILP * THREADS * BLOCKS_PER_SM * sms;                                                  \
    double cyc = ms * 1e-3 * clk;                                                                                       \
    printf("Philox4x32-%d ILP=%d threads=%d x%d/SM: %.1f G blocks/s, %.1f cycles per warp-block per SMSP\n", ROUNDS, ILP, \
           THREADS, BLOCKS_PER_SM, blocks / (ms * 1e-3) / 1e9, cyc * 4.0 / (blocks / sms / 32.0));                       \
  }
  RUN_PH(1, 10, 1024, 1) RUN_PH(2, 10, 1024, 1) RUN_PH(4, 10, 1024, 1) RUN_PH(1, 10, 512, 1) RUN_PH(2, 10, 512, 1)
  RUN_PH(1, 10, 1024, 2) RUN_PH(2, 10, 1024, 2) RUN_PH(1, 7, 1024, 1) RUN_PH(2, 7, 1024, 1)
  cudaFree(out);
  return 0;
}
