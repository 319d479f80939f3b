// smm_device.cuh -- device-side data layout of the BGP hot path (shared by the kernels and the host API).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/smm_b200.h"
#include "../../include/smm_stream.h"

namespace smm {

constexpr int kEvalThreads = 128;   // CTA size of the evaluation kernel
constexpr int kExchThreads = 512;   // CTA size of the exchange kernel
constexpr int kPairThreads = 256;   // CTA size of the pair-schedule kernel
constexpr int kMaxSplit = 64;       // CTAs cooperating on one evaluation
constexpr int kPairChunk = 128;     // iterations of pair schedules precomputed per launch

// Last-accepted record of a chain, one row of R = 3 + P + M doubles:
//   [0] value  [1] prob  [2] status (as double)  [3..3+P) params  [3+P..3+P+M) simMoments
__host__ __device__ inline int rec_len(int P, int M) { return 3 + P + M; }

// device error flags (sticky, OR-ed into DevState::err)
constexpr int kErrNegative = 1;
constexpr int kErrExhausted = 2;

struct DevProblem {
  int P, M, S, obj, noseed;
  int N, L, chain0, max_iter, world;
  int sigma_update_steps, smpl_iters, batch_size;
  int panel_T, panel_N, panel_K;
  double sigma_adjust_by, slow_seconds;
  uint64_t seed_sim, seed_algo;
  const double *lb, *ub, *init, *data, *w;  // [P] x3, [M] x2
  const double *acc_tuner, *min_improve;    // [N]
};

struct DevState {
  // per local chain
  double *sigma, *accept_rate;
  int *n_noex, *n_acc;         // iterations without exchange / accepted among them (set_acceptRate!)
  double *la_cur;              // [L][R] last accepted record = proposal centre for the next iteration
  double *la_pub;              // [L][R] record published to the exchange step
  double *la_all;              // [N][R] gathered records (== la_pub when world == 1)
  // trace, [max_iter][L] (+[P], +[M])
  double *t_value, *t_prob, *t_curr, *t_best, *t_params, *t_mom;
  uint8_t *t_acc;
  int *t_status, *t_exch, *t_bestid;
  // evaluation scratch
  double *partials;            // [L][n_split][2*kPartLen]
  unsigned *arrive;            // [L]
  // exchange schedule for iterations [sched_iter0, sched_iter0 + kPairChunk)
  int *sched_ij;               // [kPairChunk][n_s][2], level order
  int *sched_off;              // [kPairChunk][n_s + 1]
  int *sched_nlev;             // [kPairChunk]
  // diagnostics
  int *err;
  unsigned long long *counters;  // [0] accepted [1] swaps [2] proposal attempts
};

}  // namespace smm
