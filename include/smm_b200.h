/* smm_b200.h -- C ABI of libsmm_b200.so: the B200 (sm_100a) implementation of SMM.jl's
 * parallel-tempered BGP MCMC hot path.
 *
 * The reference has no FFI (it is pure Julia); each entry point below replaces a Julia-level
 * function of /root/reference/src/mopt and is what a Julia `ccall` shim (julia/SMMB200.jl,
 * INTEGRATION.md), Python ctypes (smm_jl_b200/_lib.py) or a C++ driver binds:
 *
 *   smm_bgp_create      <- MAlgoBGP(m::MProb, opts)            AlgoBGP.jl:505-538 (+ BGPChain :78-109)
 *   smm_bgp_step        <- run!(algo) / computeNextIteration!  AlgoAbstract.jl:27-76, AlgoBGP.jl:589-640
 *                          (proposal :424, evaluateObjective mprob.jl:175, objfunc_norm
 *                          ObjExamples.jl:59, doAcceptReject! :324, set_eval! :220,
 *                          exchangeMoves! :647, swap_ev_ij! :734)
 *   smm_bgp_read_trace  <- BGPChain fields / history(c)        AlgoBGP.jl:42-61,138-160
 *   smm_bgp_read_chain_state <- c.sigma, c.accept_rate         AlgoBGP.jl:253-257,381-390
 *   smm_bgp_eval_batch  <- evaluateObjective(m, p; noseed)     mprob.jl:175-205 (batched)
 *   smm_bgp_export_state / smm_bgp_import_state <- save / readMalgo / restart!
 *                                                              AlgoAbstract.jl:83-102, AlgoBGP.jl:804
 *
 * Conventions: plain pointers and sizes only; every function returns 0 or a negative SMM_E_* code
 * and leaves a message for smm_last_error() (thread local); no C++ exception crosses the boundary.
 * Host buffers are caller-owned and touched only during the call.  The library owns all device
 * memory, its CUDA stream and (world_size > 1) its NCCL communicator.  One process drives one GPU;
 * a multi-GPU run is world_size processes that each create a handle with the same config, their
 * own rank, and the same nccl_id (made by rank 0 with smm_nccl_unique_id and broadcast by the
 * host language: torch.distributed in Python, Distributed/MPI in Julia).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with SMM_E_CUDA.
 */
#ifndef SMM_B200_H
#define SMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMM_ABI_VERSION 1

/* error codes */
#define SMM_OK 0
#define SMM_E_ARG (-1)
#define SMM_E_CUDA (-2)
#define SMM_E_NCCL (-3)
#define SMM_E_UNSUPPORTED_SHAPE (-4)
#define SMM_E_NEGATIVE_OBJECTIVE (-5) /* AlgoBGP.jl:341 `error("AlgoBGP assumes ... non-negative")` */
#define SMM_E_SAMPLER_EXHAUSTED (-6)  /* AlgoBGP.jl:409 `error("no draw in support after ...")`   */
#define SMM_E_STATE (-7)

/* built-in objective functions (the device simulators selected by addEvalFunc!) */
#define SMM_OBJ_NORM 0      /* objfunc_norm      ObjExamples.jl:59-116  (P == M: row means)          */
#define SMM_OBJ_NORM_SLOW 1 /* objfunc_norm_slow ObjExamples.jl:124-184 (+ slow_seconds per eval)    */
#define SMM_OBJ_NORM_MV 2   /* means + sample variances, M == 2P (generalises
                               objfunc_norm2 ObjExamples.jl:191, which is broken upstream)           */
#define SMM_OBJ_PANEL 3     /* dynamic panel, P == 2K+4, M == 4K+8 (no upstream code)   */
#define SMM_OBJ_FAILS 4     /* Testobj_fails     ObjExamples.jl:27-32: every evaluation throws ->
                               status -2, value stays -1.0 (mprob.jl:183-186)                        */

#define SMM_MAX_PARAMS 64
#define SMM_MAX_MOMENTS 128
#define SMM_NCCL_ID_BYTES 128

typedef struct smm_bgp smm_bgp; /* opaque */

typedef struct smm_bgp_config {
  int32_t abi_version; /* SMM_ABI_VERSION */
  /* problem (MProb, mprob.jl:29-53) */
  int32_t n_params;      /* P: sampled parameters, in params_to_sample order                       */
  int32_t n_moments;     /* M: data moments, in moments order                                       */
  const double *lb;      /* [P] lower bounds                                                        */
  const double *ub;      /* [P] upper bounds                                                        */
  const double *init;    /* [P] initial_value                                                       */
  const double *data_mom; /* [M] data moment values                                                 */
  const double *data_w;   /* [M] data moment weights (divide, as an s.d.: ObjExamples.jl:97)        */
  /* objective */
  int32_t objective_id;  /* SMM_OBJ_*                                                               */
  int32_t n_sim;         /* S simulation draws per evaluation (10000 at ObjExamples.jl:76)          */
  uint64_t seed_sim;     /* 1234 at ObjExamples.jl:74                                               */
  int32_t noseed;        /* ev.options[:noseed] (ObjExamples.jl:71): fresh draws per evaluation     */
  double slow_seconds;   /* SMM_OBJ_NORM_SLOW: sleep(0.1) at ObjExamples.jl:130                     */
  int32_t panel_T;       /* SMM_OBJ_PANEL                                                           */
  int32_t panel_N;
  int32_t panel_K;
  /* algorithm (opts of MAlgoBGP, AlgoBGP.jl:505-538) */
  int32_t n_chains;            /* opts["N"] (total over all ranks)                                   */
  int32_t max_iter;            /* opts["maxiter"]: trace capacity                                    */
  const double *sigma0;        /* [N] sigma * temps[i]                                               */
  const double *acc_tuner;     /* [N] opts["acc_tuners"]                                             */
  const double *min_improve;   /* [N] opts["min_improve"]                                            */
  int32_t sigma_update_steps;  /* 10                                                                 */
  double sigma_adjust_by;      /* 0.01                                                               */
  int32_t smpl_iters;          /* 1000                                                               */
  int32_t batch_size;          /* P by default; must divide P (see DESIGN.md on the upstream bug)    */
  uint64_t seed_algo;          /* seeds Zprop, Uacc, Pairs                                           */
  /* placement */
  int32_t device;      /* CUDA device ordinal of this process                                        */
  int32_t world_size;  /* number of processes/GPUs sharing the chains (1, 2, 4, 8)                   */
  int32_t rank;        /* this process owns the chains rank, rank + world, rank + 2 world, ... (0-based global
                          ids, dealt round robin so that every rank holds the same mix of temperatures); its local
                          chain c -- column c of every trace view -- is global chain c * world_size + rank      */
  uint8_t nccl_id[SMM_NCCL_ID_BYTES]; /* from smm_nccl_unique_id on rank 0 (ignored if world == 1)   */
  int32_t exchange_mode; /* 0 = one launch per iteration (+ ncclAllGather + exchange kernel when world > 1);
                            1 = persistent cooperative kernel; with world > 1 the all-gather is fused into it
                            as peer stores over NVLink (CUDA IPC) and a flag exchange inside the grid barrier;
                            2 = the same persistent kernel without any grid barrier: every CTA waits for its rank's
                            completion counter, to which every CTA of every rank adds the chains it finished;
                            3 = as 2, but what the next iteration's critical path needs (value, sigma and last accepted
                            parameters of every chain) travels as flag-in-data words {32 payload bits | iteration tag},
                            so no system fence and no counter round trip sits between two iterations; the counter
                            only covers the full records the owners copy for swap_ev_ij! (fastest with world > 1)   */
  int32_t n_split;       /* mode 0: CTAs per chain evaluation; mode 1: cap on CTAs per SM; 0 = automatic      */
} smm_bgp_config;

/* Host-side SoA view of iterations [iter_lo, iter_hi] (1-based, inclusive) of the chains this rank
 * owns.  n = iter_hi - iter_lo + 1, L = local chains.  Row-major [n][L] (+[P] / [M]).  Any pointer
 * may be NULL to skip that column.  Mirrors the BGPChain vectors and the Eval fields history() reads. */
typedef struct smm_trace_view {
  double *value;       /* [n][L]     evals[it].value                                                */
  double *prob;        /* [n][L]     evals[it].prob                                                 */
  double *curr_val;    /* [n][L]     curr_val[it]                                                   */
  double *best_val;    /* [n][L]     best_val[it]                                                   */
  double *params;      /* [n][L][P]  evals[it].params                                               */
  double *sim_moments; /* [n][L][M]  evals[it].simMoments                                           */
  uint8_t *accepted;   /* [n][L]     accepted[it]                                                   */
  int32_t *status;     /* [n][L]     evals[it].status                                               */
  int32_t *exchanged;  /* [n][L]     exchanged[it] (1-based partner id, 0 = none)                   */
  int32_t *best_id;    /* [n][L]     best_id[it] (1-based iteration)                                */
} smm_trace_view;

typedef struct smm_counters {
  int64_t iterations;       /* iterations completed                                                  */
  int64_t evaluations;      /* objective evaluations done by this rank (iterations * local chains)  */
  int64_t kernel_launches;  /* this library's kernel launches (all kinds)                            */
  int64_t collectives;      /* NCCL collectives enqueued                                             */
  int64_t accepted;         /* Metropolis accepts over local chains                                  */
  int64_t swaps;            /* exchange moves that swapped (counted once per pair, all chains)       */
  int64_t proposal_attempts; /* rejection-loop attempts used by local chains                         */
} smm_counters;

int smm_abi_version(void);
const char *smm_last_error(void);
int smm_device_count(void);
int smm_nccl_unique_id(uint8_t out[SMM_NCCL_ID_BYTES]);

int smm_bgp_create(const smm_bgp_config *cfg, smm_bgp **out);
void smm_bgp_destroy(smm_bgp *h);
/* Process-wide caches behind create/destroy: CUDA streams and events per device; with world_size > 1 the NCCL
 * communicator of (device, world_size, rank) and the CUDA-IPC mapped exchange arena the peers store their records
 * into.  They are made by the first handle (ncclCommInitRank from cfg->nccl_id, ~1 s at 8 ranks) and handed to every
 * later handle of the same (device, world_size, rank), whose nccl_id is then not looked at -- all ranks of a job
 * create their handles in the same order, so the cache hits on every rank or on none.  smm_shutdown() ends them
 * (every handle must have been destroyed); the next smm_bgp_create starts over with a fresh nccl_id. */
void smm_shutdown(void);

/* run iterations i+1 .. i+n_iters; blocking (returns after the device finished and the sticky
 * device error flag was checked).  elapsed_ms (may be NULL) receives the CUDA-event time of the
 * region on the library's stream. */
int smm_bgp_step(smm_bgp *h, int32_t n_iters, float *elapsed_ms);
/* run!(algo) (AlgoAbstract.jl:27-76) with the trace streamed to the host: iterations i+1 .. i+n_iters are enqueued
 * in windows of `window` iterations (0 = default); the rows of a finished window are copied into host_out
 * (row 0 = iteration i+1; [n_iters][L] layout as in smm_bgp_read_trace; NULL = no read-back) on a second stream
 * while the next window computes.  Blocking.  Pinned host memory (smm_host_alloc, or the host language's own
 * pinned allocator) makes the copies asynchronous; pageable memory works but serialises them. */
int smm_bgp_run(smm_bgp *h, int32_t n_iters, int32_t window, const smm_trace_view *host_out, float *elapsed_ms);
int smm_host_alloc(int64_t nbytes, void **out); /* page-locked host memory for smm_bgp_run / read_trace */
void smm_host_free(void *p);
int smm_bgp_iteration(const smm_bgp *h); /* iterations completed so far (algo.i) */
int smm_bgp_local_chains(const smm_bgp *h);
void *smm_bgp_stream(smm_bgp *h); /* cudaStream_t the kernels run on */

int smm_bgp_read_trace(smm_bgp *h, int32_t iter_lo, int32_t iter_hi, const smm_trace_view *out);
int smm_bgp_read_chain_state(smm_bgp *h, double *sigma, double *accept_rate); /* [L] each */
int smm_bgp_get_counters(smm_bgp *h, smm_counters *out);

/* Accepted-only statistics of the parameter draws of every local chain over iterations [iter_lo, iter_hi], reduced on
 * the device (mean(c), median(c), CI(c) of AlgoBGP.jl:174-188 without shipping the trace to the host): count[L] accepted
 * iterations, mean[L][P], quantiles[L][P][n_probs] at the probabilities probs[n_probs] with the definition of Julia's
 * `quantile` (type 7: linear interpolation around (n-1) p).  A chain without accepted iterations gives NaN.  Any output
 * may be NULL.  smm_bgp_chain_summary: the columns of summary(c) (AlgoBGP.jl:197-206) that need the trace -- iterations
 * with an exchange, the partner exchanged with most often (1-based id, smallest on ties, 0 = none), best_val. */
int smm_bgp_accepted_stats(smm_bgp *h, int32_t iter_lo, int32_t iter_hi, const double *probs, int32_t n_probs,
                           int64_t *count, double *mean, double *quantiles);
int smm_bgp_chain_summary(smm_bgp *h, int64_t *n_exchanged, int32_t *exchanged_most_with, double *best_val);

/* batched bare objective: value[B], moments[B][M], status[B] at params[B][P] (host pointers).
 * noseed != 0 draws fresh shocks indexed by (entry b, rep0 + b). */
int smm_bgp_eval_batch(smm_bgp *h, const double *params, int32_t B, int32_t noseed, uint32_t rep0,
                       double *value, double *moments, int32_t *status);

/* checkpoint: the per-chain algorithm state (sigma, accept counters, last-accepted record) plus the
 * trace are enough to resume; all randomness is counter-indexed so there is no RNG state. */
int64_t smm_bgp_state_bytes(const smm_bgp *h);
int smm_bgp_export_state(smm_bgp *h, void *buf, int64_t nbytes);
int smm_bgp_import_state(smm_bgp *h, const void *buf, int64_t nbytes);

/* BGPChain.probs_acc[iter_lo..iter_hi] of a chain (0-based global id): the Uacc stream evaluated on the
 * host (a pure function of the seed; replaces `rand(n)` at AlgoBGP.jl:85).  out[iter_hi-iter_lo+1]. */
int smm_stream_acc_uniforms(uint64_t seed_algo, uint32_t chain, int32_t iter_lo, int32_t iter_hi, double *out);

/* per-kernel device timing: when enabled, smm_bgp_step brackets every launch with CUDA events on the
 * library's stream and accumulates the elapsed time per kernel kind (slows the step; off by default).
 * kinds: 0 = evaluation kernel (multi-launch) or persistent kernel, 1 = exchange kernel, 2 = pair-schedule
 * kernel, 3 = all-gather.  smm_bgp_kernel_times returns the number of BGP iterations the kind-0 launches
 * covered (>= 0; a persistent launch covers many), or a negative error code. */
int smm_bgp_set_profiling(smm_bgp *h, int32_t enabled);
int smm_bgp_kernel_times(smm_bgp *h, double ms_sum[4], int64_t launches[4]);

/* test/diagnostic entry points (device implementations of the stream definitions) */
int smm_debug_normals(int32_t device, uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int32_t n_pairs,
                      double *out /* [2*n_pairs] */);
/* same, through the ziggurat transform (smm_zig_triple: three normals per block): the simulator stream of the
 * MvNormal objectives */
int smm_debug_zig_normals(int32_t device, uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int32_t n_blocks,
                          double *out /* [3*n_blocks] */);
int smm_debug_pairs(smm_bgp *h, int32_t iter, int32_t *ij /* [n_pairs][2] in execution order */,
                    int32_t *level_offsets /* [n_pairs+1] */, int32_t *n_levels);
/* debug: per-CTA globaltimer stamps {start, after proposal, after simulate, end} of the last iteration;
 * needs SMM_PHASE_TS=1 in the environment at create time.  Returns n_split (>0) or an error. */
int smm_debug_phase_ts(smm_bgp *h, uint64_t *out, int64_t n);
/* n grid barriers back to back on one persistent-kernel-sized CTA per SM (variant 0 = the one exchange_mode 1 uses) */
int smm_debug_barrier_bench(smm_bgp *h, int32_t variant, int32_t n, float *elapsed_ms);
/* the simulate inner loop alone with this handle's keys/accumulators: blocks x threads CTAs, each thread
 * n_pairs_per_thread Philox blocks; dynamic != 0 uses the shared-counter unit distribution of the persistent kernel */
int smm_debug_sim_throughput(smm_bgp *h, int32_t n_pairs_per_thread, int32_t blocks, int32_t threads, int32_t dynamic,
                             float *elapsed_ms);
int smm_debug_rng_throughput(int32_t device, int64_t n_pairs_per_thread, int32_t blocks, int32_t threads,
                             float *elapsed_ms, double *checksum);

#ifdef __cplusplus
}
#endif
#endif /* SMM_B200_H */
