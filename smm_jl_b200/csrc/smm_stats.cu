// smm_stats.cu -- device-side reductions over the SoA trace for the accepted-only statistics of a chain
// (mean / median / CI, AlgoBGP.jl:174-188: params(c; accepted_only = true) -> mean, median, quantile) and for
// summary(c) (AlgoBGP.jl:197-206), so that printing a summary of a 1024-chain x 1000-iteration run does not ship the
// whole trace (240 MB) to the host.  Not on the hot path: one launch per call, any trace length.
#include <cuda_runtime.h>

#include "smm_device.cuh"

namespace smm {

constexpr int kStatsThreads = 256;

// ascending bitonic sort of v[0 .. n2) (n2 a power of two) by the whole CTA; v in shared or global memory
__device__ void cta_bitonic_sort(double *v, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const double a = v[i], b = v[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            v[i] = b;
            v[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// grid (P, L): accepted values of parameter k of local chain c over iterations [it_lo, it_hi], sorted;
// count[c], mean[c][k], quant[c][k][q] = type-7 quantile (Julia's `quantile` and numpy's default: linear
// interpolation between the order statistics around (n-1) p)
__global__ void __launch_bounds__(kStatsThreads) accepted_stats_kernel(const uint8_t *t_acc, const double *t_params, int L,
                                                                       int P, int it_lo, int it_hi, const double *probs,
                                                                       int n_probs, long long *count, double *mean,
                                                                       double *quant, double *scratch, int cap2,
                                                                       int smem_cap) {
  extern __shared__ double s_val[];
  __shared__ int s_n;
  __shared__ double s_red[kStatsThreads];
  const int k = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
  if (tid == 0) s_n = 0;
  __syncthreads();
  // pass 1: how many accepted (decides where the values live)
  int mine = 0;
  for (int it = it_lo + tid; it <= it_hi; it += blockDim.x) mine += t_acc[(size_t)(it - 1) * L + c] != 0;
  atomicAdd(&s_n, mine);
  __syncthreads();
  const int n = s_n;
  __syncthreads();
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  double *v = n2 <= smem_cap ? s_val : scratch + ((size_t)c * P + k) * cap2;
  if (tid == 0) s_n = 0;
  __syncthreads();
  // pass 2: gather (any order: the values are sorted next), pad with +inf
  for (int it = it_lo + tid; it <= it_hi; it += blockDim.x)
    if (t_acc[(size_t)(it - 1) * L + c]) v[atomicAdd(&s_n, 1)] = t_params[((size_t)(it - 1) * L + c) * P + k];
  for (int i = n + tid; i < n2; i += blockDim.x) v[i] = __longlong_as_double(0x7FF0000000000000ll);
  __syncthreads();
  cta_bitonic_sort(v, n2);
  // mean: fixed tree over the sorted values (independent of the gather order)
  double acc = 0.0;
  for (int i = tid; i < n; i += blockDim.x) acc += v[i];
  s_red[tid] = acc;
  __syncthreads();
  for (int s = kStatsThreads / 2; s > 0; s >>= 1) {
    if (tid < s) s_red[tid] += s_red[tid + s];
    __syncthreads();
  }
  if (tid == 0) {
    if (k == 0) count[c] = n;
    mean[(size_t)c * P + k] = n > 0 ? s_red[0] / (double)n : __longlong_as_double(0x7FF8000000000000ll);
  }
  for (int q = tid; q < n_probs; q += blockDim.x) {
    double out = __longlong_as_double(0x7FF8000000000000ll);
    if (n > 0) {
      const double h = (double)(n - 1) * probs[q];
      int lo = (int)floor(h);
      lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
      const int hi = lo + 1 < n ? lo + 1 : lo;
      const double g = h - (double)lo;
      out = v[lo] + g * (v[hi] - v[lo]);
    }
    quant[((size_t)c * P + k) * n_probs + q] = out;
  }
}

// grid L: summary(c) (AlgoBGP.jl:197-206) -- iterations with an exchange, the partner met most often (smallest id on
// ties), best_val at the last completed iteration
__global__ void __launch_bounds__(kStatsThreads) chain_summary_kernel(const int *t_exch, const double *t_best, int L, int N,
                                                                      int iter, long long *n_exchanged, int *most_with,
                                                                      double *best_val) {
  extern __shared__ int s_hist[];  // [N + 1]
  __shared__ int s_cnt;
  __shared__ unsigned long long s_best;
  const int c = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i <= N; i += blockDim.x) s_hist[i] = 0;
  if (tid == 0) {
    s_cnt = 0;
    s_best = 0ull;
  }
  __syncthreads();
  int mine = 0;
  for (int it = 1 + tid; it <= iter; it += blockDim.x) {
    const int p = t_exch[(size_t)(it - 1) * L + c];
    if (p != 0) {
      ++mine;
      if (p > 0 && p <= N) atomicAdd(&s_hist[p], 1);
    }
  }
  atomicAdd(&s_cnt, mine);
  __syncthreads();
  // argmax with the smallest partner id on ties: key = count << 32 | (0xFFFFFFFF - id)
  unsigned long long bestk = 0ull;
  for (int i = 1 + tid; i <= N; i += blockDim.x)
    if (s_hist[i] > 0) {
      const unsigned long long key = ((unsigned long long)s_hist[i] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
      bestk = key > bestk ? key : bestk;
    }
  atomicMax(&s_best, bestk);
  __syncthreads();
  if (tid == 0) {
    n_exchanged[c] = s_cnt;
    most_with[c] = s_best ? (int)(0xFFFFFFFFu - (unsigned)(s_best & 0xFFFFFFFFull)) : 0;
    best_val[c] = iter >= 1 ? t_best[(size_t)(iter - 1) * L + c] : __longlong_as_double(0x7FF0000000000000ll);
  }
}

constexpr int kStatsSmemCap = 8192;  // doubles of dynamic shared memory for the sort (64 KB)

int stats_smem_cap() { return kStatsSmemCap; }

cudaError_t launch_accepted_stats(const DevState &st, int L, int P, int it_lo, int it_hi, const double *probs, int n_probs,
                                  long long *count, double *mean, double *quant, double *scratch, int cap2,
                                  cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(accepted_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(double) * kStatsSmemCap));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  accepted_stats_kernel<<<dim3(P, L), kStatsThreads, sizeof(double) * kStatsSmemCap, s>>>(
      st.t_acc, st.t_params, L, P, it_lo, it_hi, probs, n_probs, count, mean, quant, scratch, cap2, kStatsSmemCap);
  return cudaGetLastError();
}

cudaError_t launch_chain_summary(const DevState &st, int L, int N, int iter, long long *n_exchanged, int *most_with,
                                 double *best_val, cudaStream_t s) {
  chain_summary_kernel<<<L, kStatsThreads, sizeof(int) * (size_t)(N + 1), s>>>(st.t_exch, st.t_best, L, N, iter, n_exchanged,
                                                                             most_with, best_val);
  return cudaGetLastError();
}

}  // namespace smm
