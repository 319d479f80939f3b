// smm_device.cuh -- device-side data layout of the BGP hot path (shared by the kernels and the host API).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/smm_b200.h"
#include "../../include/smm_stream.h"

namespace smm {

constexpr int kEvalThreads = 128;   // CTA size of the evaluation kernel (multi-launch mode)
constexpr int kPersistThreads = 768;   // CTA size of the persistent kernel: one CTA per SM (24 warps, 80 registers)
constexpr int kExchThreads = 512;   // CTA size of the stand-alone exchange kernel
constexpr int kPairThreads = 256;   // CTA size of the pair-schedule kernel
constexpr int kMaxSplit = 64;       // CTAs cooperating on one evaluation (multi-launch mode)
constexpr int kPairChunk = 128;     // iterations of pair schedules precomputed per launch
constexpr int kMaxWorld = 8;        // GPUs of one box
constexpr int kPanelThreads = 128;  // CTA size of the panel simulation kernel (4 warps, one individual per thread)
constexpr int kPanelMaxK = 16;      // regressors of the dynamic-panel objective (P = 2K+4 <= 36)

// pooled raw sums of the dynamic-panel objective (see smm_panel.cuh): 14 for y, 6 per regressor
__host__ __device__ inline int panel_na(int K) { return 14 + 6 * K; }

// Philox blocks per simulated row of a ziggurat objective: three draws per block (include/smm_stream.h)
__host__ __device__ inline int zig_blocks(int S) { return (S + SMM_ZIG_PER_BLOCK - 1) / SMM_ZIG_PER_BLOCK; }

// Last-accepted record of a chain, one row of R = 3 + P + M doubles:
//   [0] value  [1] prob  [2] status (as double)  [3..3+P) params  [3+P..3+P+M) simMoments
__host__ __device__ inline int rec_len(int P, int M) { return 3 + P + M; }

// Chains are dealt to the ranks round robin: local chain c of rank r is global chain c * world + r (0-based).  The
// temperature ladder runs along the global id (AlgoBGP.jl:508), and a hot chain's truncated proposal needs more
// rejection attempts than a cold one's, so contiguous blocks would make the last rank the straggler of every iteration.
// (global_chain / gather_slot below.)

// device error flags (sticky, OR-ed into DevState::err)
constexpr int kErrNegative = 1;
constexpr int kErrExhausted = 2;
constexpr int kErrTimeout = 4;

struct DevProblem {
  int P, M, S, obj, noseed;
  int N, L, max_iter, world, rank;
  int sigma_update_steps, smpl_iters, batch_size;
  int panel_T, panel_N, panel_K;
  double sigma_adjust_by, slow_seconds;
  double magic_sum, magic_sq;               // 1.5 * 2^(52-F): fixed-point rounding constants of the accumulators
  double scale_sum, scale_sq;               // 2^-F
  double pan_hi_scale, pan_hi_inv;          // panel: 2^Fhi, 2^-Fhi   (two-word fixed point of the per-individual sums)
  double pan_lo_scale, pan_lo_inv;          //        2^Flo, 2^-Flo   (Flo = Fhi + 40)
  uint64_t seed_sim, seed_algo;
  uint32_t rk_sim0[10], rk_sim1[10];        // Philox round keys of seed_sim (constant-bank operands)
  const double *lb, *ub, *init, *data, *w;  // [P] x3, [M] x2
  const double *acc_tuner, *min_improve;    // [N]
};

__host__ __device__ inline int global_chain(const DevProblem &pb, int c) { return c * pb.world + pb.rank; }
// slot of global chain g in a rank-major gather buffer (ncclAllGather of the ranks' [L] records, exchange_mode 0)
__host__ __device__ inline int gather_slot(const DevProblem &pb, int g) { return (g % pb.world) * pb.L + g / pb.world; }

// exchange_mode 3: 8-byte words per chain in the LL table: doubles {value, sigma, params[P]}, two words each (hi, lo)
__host__ __device__ inline int ll_words(int P) { return 2 * (P + 2); }

struct GridBarrier {
  unsigned arrive;
  unsigned gen;
};

struct DevState {
  // per local chain
  double *sigma, *accept_rate;
  int *n_noex, *n_acc;         // iterations without exchange / accepted among them (set_acceptRate!)
  double *la_cur;              // [L][R] last accepted record = proposal centre for the next iteration
  double *la_pub;              // [L][R] record published to the exchange step
  double *la_all;              // gathered records: [N][R] (multi-launch) or [2][N][R] by iteration parity (fused)
  double *val_all;             // [2][N] compact values of la_all (fused mode) | [N] u64 completion tags (exchange_mode 2)
  unsigned *applied;           // [L] last exchange iteration applied to the chain's state (exchange_mode 2)
  double *pp;                  // [L][P] proposals of the current iteration (persistent kernel)
  // trace, [max_iter][L] (+[P], +[M])
  double *t_value, *t_prob, *t_curr, *t_best, *t_params, *t_mom;
  uint8_t *t_acc;
  int *t_status, *t_exch, *t_bestid;
  // evaluation scratch
  double *partials;            // [L][max segments][part_len]
  unsigned *arrive;            // [L]
  unsigned *unit_ctr;          // work-queue head of the panel simulation kernel
  // exchange schedule for iterations [sched_iter0, sched_iter0 + kPairChunk)
  int *sched_ij;               // [kPairChunk][n_s][2], level order
  int *sched_off;              // [kPairChunk][n_s + 1]
  int *sched_nlev;             // [kPairChunk]
  // persistent kernel
  GridBarrier *bar;
  unsigned long long *sync_seq;            // cross-GPU sync sequence number (monotone over the handle's life)
  double *peer_la_all[kMaxWorld];          // every rank's la_all (IPC-mapped; [rank] is our own)
  double *peer_val_all[kMaxWorld];
  unsigned long long *peer_flags[kMaxWorld];  // every rank's flags[kMaxWorld]; we write slot [our rank]
  unsigned long long *flags;               // our own flags[kMaxWorld], written by the peers
  // exchange_mode 3: flag-in-data ("LL") table [2 parities][N][ll_words(P)] of {32 payload bits | 32-bit iteration tag}
  // words holding each chain's last-accepted value, its sigma and its last-accepted parameters (null in other modes)
  unsigned long long *ll;
  unsigned long long *peer_ll[kMaxWorld];
  // diagnostics
  int *err;
  unsigned long long *phase_ts;  // debug: per-CTA globaltimer stamps of the last iteration (or null)
  unsigned long long *counters;  // [0] accepted [1] swaps [2] proposal attempts [3] barrier spins
};

}  // namespace smm
