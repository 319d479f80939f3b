// smm_kernels.cu -- sm_100a kernels of the BGP hot path.
//
//   bgp_eval_kernel    one iteration of every local chain: truncated random-walk proposal
//                      (AlgoBGP.jl:424-471), model simulation + moments + weighted distance
//                      (ObjExamples.jl:59-116), Metropolis accept/reject + sigma adaptation
//                      (AlgoBGP.jl:324-392), trace bookkeeping (set_eval! :220-245).
//                      Grid (n_split, L): n_split CTAs share one evaluation; each thread owns one
//                      simulated dimension k and a strided set of Philox blocks, keeps its draws in
//                      registers, accumulates in fp64; the last CTA to arrive reduces the partials in
//                      a fixed order (deterministic) and finishes the chain.
//   bgp_pairs_kernel   Pairs[iter] for a chunk of iterations (AlgoBGP.jl:653-656) plus a level
//                      schedule: pairs that share no chain with an earlier unfinished pair get the
//                      same level, so the sequential pair loop (:662-691) runs level-parallel with
//                      identical results.  Data independent -> off the critical path.
//   bgp_exchange_kernel exchangeMoves! / swap_ev_ij! (AlgoBGP.jl:647-749) on the gathered records.
//   objective_kernel   batched bare objective (evaluateObjective, mprob.jl:175-205).
#include <cooperative_groups.h>

#include "smm_device.cuh"

namespace smm {

// ------------------------------------------------------------------------------------------------
// shared-memory scratch of the evaluation CTA
// ------------------------------------------------------------------------------------------------
struct EvalSmem {
  double pp[SMM_MAX_PARAMS];        // proposed parameter vector
  double mu01[SMM_MAX_PARAMS];      // centre in unit-cube coordinates
  double cand[2 * kEvalThreads];    // candidates of one round of attempts, [attempt][P]
  double red[2 * kEvalThreads];     // per-thread partial sums
  double tot[2 * SMM_MAX_PARAMS];   // totals over all splits
  double mom[SMM_MAX_MOMENTS];      // simulated moments
  smm_logent logtab[1 << SMM_LOG_BITS];
  unsigned char okf[2 * kEvalThreads];
  int first[SMM_MAX_PARAMS];        // per batch: first in-support attempt of this round
  int resolved[SMM_MAX_PARAMS];     // per batch: attempts used (0 = unresolved)
  int is_last;
  double value;
};

__device__ __forceinline__ void load_logtab(smm_logent *dst) {
  const smm_logent *src = smm_logtab();
  for (int i = threadIdx.x; i < (1 << SMM_LOG_BITS); i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// proposal(c) -- AlgoBGP.jl:424-471, mysample :400-410, mapto_01/ab mprob.jl:246-272.
// Attempts are counter-indexed, so a round evaluates blockDim/kp attempts at once and the lowest
// in-support attempt wins: identical to the reference's sequential loop.
// ------------------------------------------------------------------------------------------------
__device__ void block_proposal(const DevProblem &pb, const DevState &st, EvalSmem &sm, int c, int gc, int iter,
                               bool count) {
  const int P = pb.P, tid = threadIdx.x, nthr = blockDim.x;
  if (iter == 1) {
    for (int k = tid; k < P; k += nthr) sm.pp[k] = pb.init[k];
    __syncthreads();
    return;
  }
  const int R = rec_len(pb.P, pb.M);
  const double *la = st.la_cur + (size_t)c * R;
  const double sigma = st.sigma[c];
  const int kp = (P + 1) >> 1;
  const int A = nthr / kp;                // attempts per round
  const int bs = pb.batch_size, nb = P / bs;
  for (int k = tid; k < P; k += nthr) {
    sm.mu01[k] = __ddiv_rn(__dsub_rn(la[3 + k], pb.lb[k]), __dsub_rn(pb.ub[k], pb.lb[k]));
    sm.pp[k] = 0.0;  // pp = zero(mu01) (:445)
  }
  for (int b = tid; b < nb; b += nthr) sm.resolved[b] = 0;
  __syncthreads();
  int unresolved = nb;
  for (int base = 0; base < pb.smpl_iters && unresolved > 0; base += A) {
    for (int b = tid; b < nb; b += nthr) sm.first[b] = 0x7fffffff;
    if (tid < A * kp) {
      const int a_loc = tid / kp, kq = tid - a_loc * kp;
      const int a = base + a_loc;
      if (a < pb.smpl_iters) {
        double z0, z1;
        smm_normal_pair_tab(smm_prop_block(pb.seed_algo, (uint32_t)gc, (uint32_t)iter, (uint32_t)a, (uint32_t)kq),
                            sm.logtab, &z0, &z1);
        const int k0 = 2 * kq, k1 = k0 + 1;
        const double x0 = __dadd_rn(sm.mu01[k0], __dmul_rn(sigma, z0));
        sm.cand[a_loc * P + k0] = x0;
        sm.okf[a_loc * P + k0] = (x0 >= 0.0) && (x0 <= 1.0);
        if (k1 < P) {
          const double x1 = __dadd_rn(sm.mu01[k1], __dmul_rn(sigma, z1));
          sm.cand[a_loc * P + k1] = x1;
          sm.okf[a_loc * P + k1] = (x1 >= 0.0) && (x1 <= 1.0);
        }
      }
    }
    __syncthreads();
    for (int t = tid; t < A * nb; t += nthr) {
      const int a_loc = t / nb, b = t - a_loc * nb;
      if (sm.resolved[b] == 0 && base + a_loc < pb.smpl_iters) {
        bool ok = true;
        for (int k = b * bs; k < (b + 1) * bs; ++k) ok = ok && sm.okf[a_loc * P + k];
        if (ok) atomicMin(&sm.first[b], a_loc);
      }
    }
    __syncthreads();
    for (int k = tid; k < P; k += nthr) {
      const int b = k / bs;
      if (sm.resolved[b] == 0 && sm.first[b] != 0x7fffffff) sm.pp[k] = sm.cand[sm.first[b] * P + k];
    }
    __syncthreads();
    int still = 0;
    for (int b = 0; b < nb; ++b) {  // every thread computes the same count (nb <= 64)
      if (sm.resolved[b] == 0) {
        if (sm.first[b] != 0x7fffffff) {
          if (tid == 0) sm.resolved[b] = base + sm.first[b] + 1;
        } else {
          ++still;
        }
      }
    }
    unresolved = still;
    __syncthreads();
  }
  if (unresolved > 0) {
    // single batch: `error("no draw in support ...")` (:409) aborts the run -> sticky error flag;
    // several batches: the exception is logged and swallowed, pp[i] stays 0 (:447-451)
    if (nb == 1) {
      if (tid == 0) atomicOr(st.err, kErrExhausted);
      for (int k = tid; k < P; k += nthr) sm.pp[k] = sm.mu01[k];
    }
  }
  if (count && tid == 0) {
    unsigned long long att = 0;
    for (int b = 0; b < nb; ++b) att += sm.resolved[b] ? sm.resolved[b] : pb.smpl_iters;
    atomicAdd(&st.counters[2], att);
  }
  __syncthreads();
  for (int k = tid; k < P; k += nthr)
    sm.pp[k] = __dadd_rn(__dmul_rn(sm.pp[k], __dsub_rn(pb.ub[k], pb.lb[k])), pb.lb[k]);
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Simulation of the MvNormal objectives (ObjExamples.jl:76-79): X[k,s] = p_k + Z[k,s], reduced on the
// fly to sum_s X and sum_s X^2 per row.  Thread t owns row k = t % D and Philox blocks
// j = j0 + t/D, +lanes, ... of this CTA's share [j0, j1) of the ceil(S/2) blocks.
// Writes this CTA's partial sums [2][D] to `part`.
// ------------------------------------------------------------------------------------------------
__device__ void simulate_norm_partial(const DevProblem &pb, EvalSmem &sm, int split, int n_split, uint32_t uid,
                                      uint32_t rep, double *part) {
  const int D = pb.P, S = pb.S, tid = threadIdx.x;
  const int lanes = blockDim.x / D;
  const int n_blocks = (S + 1) >> 1;
  const int j0 = (int)(((long long)n_blocks * split) / n_split);
  const int j1 = (int)(((long long)n_blocks * (split + 1)) / n_split);
  double sum = 0.0, sq = 0.0;
  if (tid < lanes * D) {
    const int k = tid % D, ln = tid / D;
    const double p = sm.pp[k];
    for (int j = j0 + ln; j < j1; j += lanes) {
      double z0, z1;
      smm_normal_pair_tab(smm_sim_block(pb.seed_sim, (uint32_t)j, (uint32_t)k, pb.noseed, uid, rep), sm.logtab, &z0,
                          &z1);
      const double x0 = __dadd_rn(p, z0);
      sum = __dadd_rn(sum, x0);
      sq = __fma_rn(x0, x0, sq);
      if (2 * j + 1 < S) {
        const double x1 = __dadd_rn(p, z1);
        sum = __dadd_rn(sum, x1);
        sq = __fma_rn(x1, x1, sq);
      }
    }
  }
  sm.red[2 * tid] = sum;
  sm.red[2 * tid + 1] = sq;
  __syncthreads();
  if (tid < 2 * D) {
    const int k = tid % D, which = tid / D;
    double acc = 0.0;
    for (int ln = 0; ln < lanes; ++ln) acc = __dadd_rn(acc, sm.red[2 * (ln * D + k) + which]);
    part[which * D + k] = acc;
  }
}

// moments + weighted distance from the totals (thread 0..M-1 compute moments, thread 0 the value)
__device__ void finalize_norm(const DevProblem &pb, EvalSmem &sm, bool with_var) {
  const int D = pb.P, tid = threadIdx.x;
  const double S = (double)pb.S;
  if (tid < D) {
    const double mean = __ddiv_rn(sm.tot[tid], S);
    sm.mom[tid] = mean;
    if (with_var) {
      // sum (x - mean)^2 = sum x^2 - mean * sum x
      const double ss = __dsub_rn(sm.tot[D + tid], __dmul_rn(mean, sm.tot[tid]));
      sm.mom[D + tid] = __ddiv_rn(ss, S - 1.0);
    }
  }
  __syncthreads();
  if (tid == 0) {
    double acc = 0.0;
    for (int k = 0; k < pb.M; ++k) {
      const double d = __ddiv_rn(__dsub_rn(sm.mom[k], pb.data[k]), pb.w[k]);
      acc = __fma_rn(d, d, acc);
    }
    sm.value = __ddiv_rn(acc, (double)pb.M);
  }
  __syncthreads();
}

__device__ __forceinline__ void slow_spin(double seconds) {
  // objfunc_norm_slow: sleep(0.1) (ObjExamples.jl:130)
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const unsigned long long ns = (unsigned long long)(seconds * 1e9);
  do {
    __nanosleep(20000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < ns);
}

// Evaluate the objective at sm.pp cooperatively over n_split CTAs.  Returns true in the one CTA that
// arrives last; there sm.mom / sm.value / *status are complete.
__device__ bool objective_core(const DevProblem &pb, EvalSmem &sm, int split, int n_split, int part_len,
                               uint32_t uid, uint32_t rep, double *part_base, unsigned *arrive, int *status) {
  const int tid = threadIdx.x;
  if (pb.obj == SMM_OBJ_FAILS) {
    // the objective throws -> caught by evaluateObjective, status -2, value stays -1.0, no moments
    if (split != 0) return false;
    for (int k = tid; k < pb.M; k += blockDim.x) sm.mom[k] = __longlong_as_double(0x7ff8000000000000ll);
    if (tid == 0) sm.value = -1.0;
    *status = -2;
    __syncthreads();
    return true;
  }
  if (pb.obj == SMM_OBJ_NORM_SLOW) slow_spin(pb.slow_seconds);
  double *part = part_base + (size_t)split * part_len;
  simulate_norm_partial(pb, sm, split, n_split, uid, rep, part);
  const int n_tot = 2 * pb.P;
  if (n_split > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(arrive, 1u);
      sm.is_last = (prev == (unsigned)(n_split - 1));
      if (sm.is_last) *arrive = 0u;  // re-arm for the next iteration
    }
    __syncthreads();
    if (!sm.is_last) return false;
    __threadfence();
  } else {
    __syncthreads();
  }
  if (tid < n_tot) {
    double acc = 0.0;
    for (int s = 0; s < n_split; ++s) acc = __dadd_rn(acc, __ldcg(part_base + (size_t)s * part_len + tid));
    sm.tot[tid] = acc;
  }
  __syncthreads();
  finalize_norm(pb, sm, pb.obj == SMM_OBJ_NORM_MV);
  *status = 1;
  return true;
}

// ------------------------------------------------------------------------------------------------
// one BGP iteration for every local chain
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEvalThreads) bgp_eval_kernel(DevProblem pb, DevState st, int iter, int n_split,
                                                                int part_len) {
  __shared__ EvalSmem sm;
  const int c = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  const int gc = pb.chain0 + c;
  const int P = pb.P, M = pb.M, R = rec_len(P, M), L = pb.L;
  load_logtab(sm.logtab);
  __syncthreads();
  block_proposal(pb, st, sm, c, gc, iter, split == 0);

  int ev_status = -1;
  const bool last = objective_core(pb, sm, split, n_split, part_len, (uint32_t)gc, (uint32_t)iter,
                                   st.partials + (size_t)c * n_split * part_len, st.arrive + c, &ev_status);
  if (!last) return;

  // ---- doAcceptReject! (AlgoBGP.jl:324-392) + set_eval! (:220-245), one thread ------------------
  double *la = st.la_cur + (size_t)c * R;
  double *pub = st.la_pub + (size_t)c * R;
  const size_t slot = (size_t)(iter - 1) * L + c;
  __shared__ int s_acc;
  __shared__ double s_prob;
  __shared__ int s_status;
  if (tid == 0) {
    const double value = sm.value;
    double prob;
    int accepted, status = ev_status;
    if (iter == 1) {
      prob = 1.0;
      accepted = 1;
      status = 1;
    } else {
      const double old_value = la[0];
      if (status < 0) {
        prob = 0.0;
        accepted = 0;
      } else {
        if (!(value >= 0.0)) atomicOr(st.err, kErrNegative);  // `error(...)` upstream (:341)
        const double e = exp(__dmul_rn(pb.acc_tuner[gc], __dsub_rn(old_value, value)));
        prob = isnan(e) ? e : (e < 1.0 ? e : 1.0);  // minimum([1.0, e]) propagates NaN
        if (!isfinite(prob)) {
          prob = 0.0;
          accepted = 0;
          status = -1;
        } else if (!isfinite(old_value)) {
          prob = 1.0;
          accepted = 1;
        } else {
          status = 1;
          accepted = prob > smm_acc_uniform(pb.seed_algo, (uint32_t)gc, (uint32_t)iter);
        }
      }
    }
    // set_acceptRate! (:253-257): this iteration has exchanged == 0 at this point
    const int n_noex = st.n_noex[c] + 1, n_acc = st.n_acc[c] + accepted;
    st.n_noex[c] = n_noex;
    st.n_acc[c] = n_acc;
    const double rate = __ddiv_rn((double)n_acc, (double)n_noex);
    st.accept_rate[c] = rate;
    if (iter > 1 && iter % pb.sigma_update_steps == 0) {
      const double s = st.sigma[c];
      st.sigma[c] = rate > 0.234 ? __dmul_rn(s, __dadd_rn(1.0, pb.sigma_adjust_by))
                                 : __dmul_rn(s, __dsub_rn(1.0, pb.sigma_adjust_by));
    }
    // set_eval!
    double curr, best;
    int best_id;
    if (iter == 1) {
      curr = value;
      best = value;
      best_id = 1;
    } else {
      const size_t prev = slot - L;
      curr = accepted ? value : st.t_curr[prev];
      const double bprev = st.t_best[prev];
      if (value < bprev) {
        best = value;
        best_id = iter;
      } else {
        best = bprev;
        best_id = st.t_bestid[prev];
      }
    }
    st.t_value[slot] = value;
    st.t_prob[slot] = prob;
    st.t_curr[slot] = curr;
    st.t_best[slot] = best;
    st.t_acc[slot] = (uint8_t)accepted;
    st.t_status[slot] = status;
    st.t_exch[slot] = 0;
    st.t_bestid[slot] = best_id;
    if (accepted) atomicAdd(&st.counters[0], 1ull);
    s_acc = accepted;
    s_prob = prob;
    s_status = status;
  }
  __syncthreads();
  // trace rows + last-accepted record (coalesced over threads)
  for (int k = tid; k < P; k += blockDim.x) st.t_params[slot * P + k] = sm.pp[k];
  for (int k = tid; k < M; k += blockDim.x) st.t_mom[slot * M + k] = sm.mom[k];
  if (s_acc) {
    if (tid == 0) {
      la[0] = sm.value;
      la[1] = s_prob;
      la[2] = (double)s_status;
    }
    for (int k = tid; k < P; k += blockDim.x) la[3 + k] = sm.pp[k];
    for (int k = tid; k < M; k += blockDim.x) la[3 + P + k] = sm.mom[k];
  }
  __syncthreads();
  for (int k = tid; k < R; k += blockDim.x) pub[k] = la[k];
}

// ------------------------------------------------------------------------------------------------
// batched bare objective: grid (n_split, B)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEvalThreads) objective_kernel(DevProblem pb, const double *params, int noseed,
                                                                 uint32_t rep0, int n_split, int part_len,
                                                                 double *partials, unsigned *arrive, double *value,
                                                                 double *moments, int *status) {
  __shared__ EvalSmem sm;
  const int b = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  load_logtab(sm.logtab);
  for (int k = tid; k < pb.P; k += blockDim.x) sm.pp[k] = params[(size_t)b * pb.P + k];
  __syncthreads();
  pb.noseed = noseed;
  int ev_status = -1;
  const bool last = objective_core(pb, sm, split, n_split, part_len, (uint32_t)b, rep0 + (uint32_t)b,
                                   partials + (size_t)b * n_split * part_len, arrive + b, &ev_status);
  if (!last) return;
  if (tid == 0) {
    value[b] = sm.value;
    status[b] = ev_status;
  }
  for (int k = tid; k < pb.M; k += blockDim.x) moments[(size_t)b * pb.M + k] = sm.mom[k];
}

// ------------------------------------------------------------------------------------------------
// Pairs[iter] + level schedule, one CTA per iteration.
// dynamic smem: cand[n_s] pi[n_s] pj[n_s] lvl[n_s] cnt[n_s+2] last[N]   (all 32-bit)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPairThreads) bgp_pairs_kernel(DevProblem pb, DevState st, int iter0, int n_iters,
                                                                 int n_s) {
  extern __shared__ unsigned smem_u[];
  const int N = pb.N, tid = threadIdx.x, nthr = blockDim.x;
  const int it_idx = blockIdx.x;
  if (it_idx >= n_iters) return;
  const int iter = iter0 + it_idx;
  unsigned *cand = smem_u;
  int *pi = (int *)(cand + n_s), *pj = pi + n_s, *lvl = pj + n_s, *cnt = lvl + n_s, *last = cnt + n_s + 2;
  __shared__ int s_min, s_found;
  const unsigned n_all = (unsigned)((unsigned long long)N * (N - 1) / 2);
  for (int t = tid; t < n_s; t += nthr) cand[t] = smm_pair_candidate(pb.seed_algo, (uint32_t)iter, (uint32_t)t, 0u, n_all);
  __syncthreads();
  // resolve duplicates in slot order: slot t keeps its first candidate not among slots < t
  int start = 0;
  for (;;) {
    if (tid == 0) s_min = 0x7fffffff;
    __syncthreads();
    for (int t = start + tid; t < n_s; t += nthr) {
      const unsigned v = cand[t];
      bool dup = false;
      for (int s = 0; s < t && !dup; ++s) dup = (cand[s] == v);
      if (dup) atomicMin(&s_min, t);
    }
    __syncthreads();
    const int tstar = s_min;
    if (tstar == 0x7fffffff) break;
    for (unsigned a = 1;; ++a) {
      const unsigned v = smm_pair_candidate(pb.seed_algo, (uint32_t)iter, (uint32_t)tstar, a, n_all);
      __syncthreads();
      if (tid == 0) s_found = 0;
      __syncthreads();
      bool dup = false;
      for (int s = tid; s < tstar; s += nthr) dup = dup || (cand[s] == v);
      if (dup) s_found = 1;
      __syncthreads();
      if (!s_found) {
        if (tid == 0) cand[tstar] = v;
        break;
      }
    }
    __syncthreads();
    start = tstar + 1;  // slots <= tstar are final
  }
  for (int t = tid; t < n_s; t += nthr) {
    uint32_t i, j;
    smm_pair_unrank(cand[t], &i, &j);
    pi[t] = (int)i;
    pj[t] = (int)j;
  }
  for (int i = tid; i < N; i += nthr) last[i] = 0;
  for (int i = tid; i < n_s + 2; i += nthr) cnt[i] = 0;
  __syncthreads();
  if (tid == 0) {
    int maxl = 0;
    for (int t = 0; t < n_s; ++t) {
      const int i = pi[t], j = pj[t];
      const int l = 1 + max(last[i], last[j]);
      last[i] = l;
      last[j] = l;
      lvl[t] = l;
      cnt[l] += 1;
      maxl = max(maxl, l);
    }
    // cnt[l] = #pairs of level l (1-based levels) -> off[l-1] = start of level l, off[maxl] = n_s
    int *off = st.sched_off + (size_t)it_idx * (n_s + 1);
    int run = 0;
    for (int l = 1; l <= maxl; ++l) {
      const int n_l = cnt[l];
      off[l - 1] = run;
      cnt[l] = run;
      run += n_l;
    }
    off[maxl] = run;
    st.sched_nlev[it_idx] = maxl;
    // stable placement (order inside a level is irrelevant: its pairs are disjoint)
    int *ij = st.sched_ij + (size_t)it_idx * n_s * 2;
    for (int t = 0; t < n_s; ++t) {
      const int l = lvl[t];
      const int pos = cnt[l]++;  // cnt[l] = running start of level l (levels are 1-based)
      ij[2 * pos] = pi[t];
      ij[2 * pos + 1] = pj[t];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// exchangeMoves! (AlgoBGP.jl:647-691) + swap_ev_ij! (:734-749): replicated on every rank over the
// gathered last-accepted records; each rank rewrites slot `iter` of the chains it owns.
// dynamic smem: val[N] (double) own[N] exch[N] (int)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kExchThreads) bgp_exchange_kernel(DevProblem pb, DevState st, int iter,
                                                                    int sched_idx, int n_s) {
  extern __shared__ double smem_d[];
  const int N = pb.N, tid = threadIdx.x, nthr = blockDim.x;
  const int P = pb.P, M = pb.M, R = rec_len(P, M), L = pb.L;
  double *val = smem_d;
  int *own = (int *)(val + N), *exch = own + N;
  for (int i = tid; i < N; i += nthr) {
    val[i] = st.la_all[(size_t)i * R];
    own[i] = i;
    exch[i] = 0;
  }
  __syncthreads();
  const int *ij = st.sched_ij + (size_t)sched_idx * n_s * 2;
  const int *off = st.sched_off + (size_t)sched_idx * (n_s + 1);
  const int nlev = st.sched_nlev[sched_idx];
  unsigned n_swaps = 0;
  for (int l = 0; l < nlev; ++l) {
    const int lo = off[l], hi = off[l + 1];
    for (int t = lo + tid; t < hi; t += nthr) {
      const int i = ij[2 * t], j = ij[2 * t + 1];
      const double vi = val[i], vj = val[j];
      if (__dsub_rn(vi, vj) > pb.min_improve[i]) {  // dist_fun(evi.value, evj.value) > min_improve[i]
        val[i] = vj;
        val[j] = vi;
        const int oi = own[i];
        own[i] = own[j];
        own[j] = oi;
        exch[i] = j + 1;
        exch[j] = i + 1;
        ++n_swaps;
      }
    }
    __syncthreads();
  }
  if (pb.chain0 == 0 && n_swaps) atomicAdd(&st.counters[1], (unsigned long long)n_swaps);
  // rewrite the chains this rank owns that took part in a swap: set_eval!(ci, ej) + set_exchanged!
  const size_t row = (size_t)(iter - 1) * L;
  for (int c = tid / 32; c < L; c += nthr / 32) {  // one warp per chain
    const int lane = tid & 31, gc = pb.chain0 + c;
    const int partner = exch[gc];
    if (partner == 0) continue;
    const double *src = st.la_all + (size_t)own[gc] * R;
    double *la = st.la_cur + (size_t)c * R;
    for (int k = lane; k < R; k += 32) la[k] = src[k];
    const size_t slot = row + c;
    for (int k = lane; k < P; k += 32) st.t_params[slot * P + k] = src[3 + k];
    for (int k = lane; k < M; k += 32) st.t_mom[slot * M + k] = src[3 + P + k];
    if (lane == 0) {
      const double value = src[0];
      // this iteration no longer counts towards the acceptance rate (exchanged != 0)
      st.n_noex[c] -= 1;
      st.n_acc[c] -= (int)st.t_acc[slot];
      st.t_value[slot] = value;
      st.t_prob[slot] = src[1];
      st.t_status[slot] = (int)src[2];
      st.t_acc[slot] = 1;  // the swapped-in eval is an accepted one
      st.t_curr[slot] = value;
      const double bprev = st.t_best[slot - L];
      if (value < bprev) {
        st.t_best[slot] = value;
        st.t_bestid[slot] = iter;
      } else {
        st.t_best[slot] = bprev;
        st.t_bestid[slot] = st.t_bestid[slot - L];
      }
      st.t_exch[slot] = partner;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------------
__global__ void debug_normals_kernel(uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int n_pairs, double *out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pairs) return;
  double z0, z1;
  smm_normal_pair(smm_philox4x32_10((uint32_t)j, k, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32)), &z0, &z1);
  out[2 * j] = z0;
  out[2 * j + 1] = z1;
}

// RNG-only roofline: Philox + Box-Muller + the two accumulations, nothing else
__global__ void __launch_bounds__(kEvalThreads) rng_throughput_kernel(long long n_per_thread, double *out) {
  __shared__ smm_logent tab[1 << SMM_LOG_BITS];
  load_logtab(tab);
  __syncthreads();
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  double sum = 0.0, sq = 0.0;
  for (long long j = 0; j < n_per_thread; ++j) {
    double z0, z1;
    smm_normal_pair_tab(smm_philox4x32_10((uint32_t)j, gid, 0u, 0u, 1234u, 0u), tab, &z0, &z1);
    sum = __dadd_rn(sum, z0);
    sq = __fma_rn(z0, z0, sq);
    sum = __dadd_rn(sum, z1);
    sq = __fma_rn(z1, z1, sq);
  }
  out[2 * gid] = sum;
  out[2 * gid + 1] = sq;
}

// ------------------------------------------------------------------------------------------------
// launchers (called from smm_api.cu)
// ------------------------------------------------------------------------------------------------
size_t pairs_smem_bytes(int N, int n_s) { return sizeof(unsigned) * ((size_t)5 * n_s + 2 + N); }
size_t exch_smem_bytes(int N) { return sizeof(double) * (size_t)N + sizeof(int) * 2 * (size_t)N; }

cudaError_t configure_kernels(int N, int n_s) {
  cudaError_t e = cudaFuncSetAttribute(bgp_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)pairs_smem_bytes(N, n_s));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(bgp_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)exch_smem_bytes(N));
}

int eval_max_blocks_per_sm() {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bgp_eval_kernel, kEvalThreads, 0);
  return n;
}

void launch_eval(const DevProblem &pb, const DevState &st, int iter, int n_split, int part_len, cudaStream_t s) {
  dim3 grid(n_split, pb.L);
  bgp_eval_kernel<<<grid, kEvalThreads, 0, s>>>(pb, st, iter, n_split, part_len);
}
void launch_pairs(const DevProblem &pb, const DevState &st, int iter0, int n_iters, int n_s, cudaStream_t s) {
  bgp_pairs_kernel<<<n_iters, kPairThreads, pairs_smem_bytes(pb.N, n_s), s>>>(pb, st, iter0, n_iters, n_s);
}
void launch_exchange(const DevProblem &pb, const DevState &st, int iter, int sched_idx, int n_s, cudaStream_t s) {
  bgp_exchange_kernel<<<1, kExchThreads, exch_smem_bytes(pb.N), s>>>(pb, st, iter, sched_idx, n_s);
}
void launch_objective(const DevProblem &pb, const double *params, int B, int noseed, uint32_t rep0, int n_split,
                      int part_len, double *partials, unsigned *arrive, double *value, double *moments, int *status,
                      cudaStream_t s) {
  dim3 grid(n_split, B);
  objective_kernel<<<grid, kEvalThreads, 0, s>>>(pb, params, noseed, rep0, n_split, part_len, partials, arrive, value,
                                                 moments, status);
}
void launch_debug_normals(uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int n_pairs, double *out,
                          cudaStream_t s) {
  debug_normals_kernel<<<(n_pairs + 255) / 256, 256, 0, s>>>(seed, k, c2, c3, n_pairs, out);
}
void launch_rng_throughput(long long n_per_thread, int blocks, double *out, cudaStream_t s) {
  rng_throughput_kernel<<<blocks, kEvalThreads, 0, s>>>(n_per_thread, out);
}

}  // namespace smm
