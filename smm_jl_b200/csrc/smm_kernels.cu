// smm_kernels.cu -- sm_100a kernels of the BGP hot path.
//
// Two ways to run an iteration, same device functions, bit-identical results:
//
//  (A) multi-launch (exchange_mode 0):
//      bgp_eval_kernel      grid (n_split, L): proposal -> simulate -> [last CTA of the chain] moments,
//                           distance, accept/reject, trace.
//      [ncclAllGather of the last-accepted records when world > 1]
//      bgp_exchange_kernel  exchangeMoves! on the gathered records.
//
//  (B) persistent (exchange_mode 1): bgp_persistent_kernel, one cooperative launch for up to kPairChunk
//      iterations, one CTA set resident on every SM, two grid barriers per iteration:
//        [owner CTAs: exchange of iteration i-1 (replicated), proposal of iteration i]      -- barrier --
//        [all CTAs: an equal share of the flattened (chain, draw) space; the CTA that completes a
//         chain finishes it (moments, distance, accept/reject, trace) and, with world > 1, stores the
//         chain's record straight into every peer GPU's gather buffer over NVLink]         -- barrier,
//         folded with a cross-GPU flag exchange: the all-gather costs no launch and no extra barrier --
//
// Reference lines: proposal AlgoBGP.jl:424-471 (mysample :400-410, mapto_01/ab mprob.jl:246-272);
// objfunc_norm ObjExamples.jl:59-116; doAcceptReject! AlgoBGP.jl:324-392; set_eval! :220-245;
// set_acceptRate! :253-257; exchangeMoves! :647-691; swap_ev_ij! :734-749; pair sample :653-656.
//
// Each thread owns one simulated dimension k and a strided set of Philox blocks, keeps its draws in
// registers and accumulates in fp64; partial sums meet in shared memory, then (across CTAs) in a small
// global buffer that the last-arriving CTA reduces in a fixed order, so results are deterministic.
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
  double tot[2 * SMM_MAX_PARAMS];   // totals over all segments
  double mom[SMM_MAX_MOMENTS];      // simulated moments
  smm_logent logtab[1 << SMM_LOG_BITS];
  unsigned char okf[2 * kEvalThreads];
  int first[SMM_MAX_PARAMS];        // per batch: first in-support attempt of this round
  int resolved[SMM_MAX_PARAMS];     // per batch: attempts used (0 = unresolved)
  int is_last;
  int acc, status;
  double value, prob;
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PHASE_STAMP(slot, i)                                                                  \
  do {                                                                                        \
    if (st.phase_ts && threadIdx.x == 0) st.phase_ts[(size_t)(slot)*4 + (i)] = gtimer();      \
  } while (0)

__device__ __forceinline__ void load_logtab(smm_logent *dst) {
  const smm_logent *src = smm_logtab();
  for (int i = threadIdx.x; i < (1 << SMM_LOG_BITS); i += blockDim.x) dst[i] = src[i];
}

// Philox4x32-10 with the round keys of seed_sim taken from the kernel-parameter constant bank
__device__ __forceinline__ smm_u32x4 philox_sim(const DevProblem &pb, uint32_t c0, uint32_t c1, uint32_t c2,
                                                uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)SMM_PHILOX_M0 * c0;
    const uint64_t p1 = (uint64_t)SMM_PHILOX_M1 * c2;
    c0 = (uint32_t)(p1 >> 32) ^ c1 ^ pb.rk_sim0[r];
    c1 = (uint32_t)p1;
    c2 = (uint32_t)(p0 >> 32) ^ c3 ^ pb.rk_sim1[r];
    c3 = (uint32_t)p0;
  }
  smm_u32x4 out;
  out.x = c0;
  out.y = c1;
  out.z = c2;
  out.w = c3;
  return out;
}

// ------------------------------------------------------------------------------------------------
// proposal(c) -- AlgoBGP.jl:424-471.  Attempts are counter-indexed, so a round evaluates
// blockDim/kp attempts at once and the lowest in-support attempt wins: identical to the reference's
// sequential rejection loop.  Result in sm.pp.
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
  const double sigma = __ldcg(st.sigma + c);
  const int kp = (P + 1) >> 1;
  const int A = nthr / kp;  // attempts per round
  const int bs = pb.batch_size, nb = P / bs;
  for (int k = tid; k < P; k += nthr) {
    sm.mu01[k] = __ddiv_rn(__dsub_rn(__ldcg(la + 3 + k), pb.lb[k]), __dsub_rn(pb.ub[k], pb.lb[k]));
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
    int still = 0;
    for (int b = 0; b < nb; ++b)  // every thread computes the same count (nb <= 64)
      if (sm.resolved[b] == 0 && sm.first[b] == 0x7fffffff) ++still;
    unresolved = still;
    __syncthreads();
    for (int b = tid; b < nb; b += nthr)
      if (sm.resolved[b] == 0 && sm.first[b] != 0x7fffffff) sm.resolved[b] = base + sm.first[b] + 1;
    __syncthreads();
  }
  if (unresolved > 0 && nb == 1) {
    // single batch: `error("no draw in support ...")` (:409) aborts the run -> sticky error flag;
    // (several batches: the exception is logged and swallowed, pp[i] stays 0, :447-451)
    if (tid == 0) atomicOr(st.err, kErrExhausted);
    for (int k = tid; k < P; k += nthr) sm.pp[k] = sm.mu01[k];
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
// fly to sum_s X and sum_s X^2 per row, for Philox blocks j in [j0, j1) of this evaluation.
// Thread t owns row k = t % D and blocks j0 + t/D, +lanes, ...   Writes partial sums [2][D] to `part`.
//
// ORDER-INVARIANT ACCUMULATION.  Every term is rounded once to a fixed-point grid (x -> x + M with
// M = 1.5 * 2^(52-F): the low mantissa bits of the sum are round(x * 2^F)) and the 64-bit patterns are
// added as integers, so the totals do not depend on how draws are split over threads, CTAs, launches or
// GPUs: chains with equal parameters get bit-equal values (the exchange step compares values, ties must
// stay ties as in the sequential reference), and 1-GPU and N-GPU runs agree to the bit.  Grid error per
// term <= 2^-(F+1) (F chosen by the host from S and the parameter box; ~2^-45 for the C2 shapes).
// ------------------------------------------------------------------------------------------------
__device__ void simulate_norm_segment(const DevProblem &pb, EvalSmem &sm, int j0, int j1, uint32_t uid, uint32_t rep,
                                      double *part) {
  const int D = pb.P, S = pb.S, tid = threadIdx.x;
  const int lanes = blockDim.x / D;
  const int n_full = S >> 1;  // blocks whose two normals are both used
  const uint32_t c2 = pb.noseed ? uid : 0u;
  const uint32_t c3 = (SMM_STREAM_SIM << 28) | (pb.noseed ? (rep & SMM_ITER_MASK) : 0u);
  const double Msum = pb.magic_sum, Msq = pb.magic_sq;
  unsigned long long isum = 0ull, isq = 0ull;  // sums of raw bit patterns; the n*bits(M) offset leaves at the end
  if (tid < lanes * D) {
    const int k = tid % D, ln = tid / D;
    const double p = sm.pp[k];
    const int jend = j1 < n_full ? j1 : n_full;
    for (int j = j0 + ln; j < jend; j += lanes) {
      double z0, z1;
      smm_normal_pair_tab(philox_sim(pb, (uint32_t)j, (uint32_t)k, c2, c3), sm.logtab, &z0, &z1);
      const double x0 = __dadd_rn(p, z0);
      const double x1 = __dadd_rn(p, z1);
      isum += (unsigned long long)__double_as_longlong(__dadd_rn(x0, Msum));
      isq += (unsigned long long)__double_as_longlong(__fma_rn(x0, x0, Msq));
      isum += (unsigned long long)__double_as_longlong(__dadd_rn(x1, Msum));
      isq += (unsigned long long)__double_as_longlong(__fma_rn(x1, x1, Msq));
    }
    if ((S & 1) && ln == 0 && j0 <= n_full && n_full < j1) {  // odd S: the last block contributes one draw
      double z0, z1;
      smm_normal_pair_tab(philox_sim(pb, (uint32_t)n_full, (uint32_t)k, c2, c3), sm.logtab, &z0, &z1);
      const double x0 = __dadd_rn(p, z0);
      isum += (unsigned long long)__double_as_longlong(__dadd_rn(x0, Msum));
      isq += (unsigned long long)__double_as_longlong(__fma_rn(x0, x0, Msq));
    }
  }
  unsigned long long *red = (unsigned long long *)sm.red;
  red[2 * tid] = isum;
  red[2 * tid + 1] = isq;
  __syncthreads();
  if (tid < 2 * D) {
    const int k = tid % D, which = tid / D;
    unsigned long long acc = 0ull;
    for (int ln = 0; ln < lanes; ++ln) acc += red[2 * (ln * D + k) + which];
    ((unsigned long long *)part)[which * D + k] = acc;
  }
  __syncthreads();
}

// sum the n_seg partials of an evaluation (integer adds: exact), then moments + weighted distance
__device__ void reduce_and_finalize(const DevProblem &pb, EvalSmem &sm, const double *part_base, int n_seg,
                                    int part_len) {
  const int D = pb.P, tid = threadIdx.x;
  if (pb.obj == SMM_OBJ_FAILS) {
    // the objective throws -> caught by evaluateObjective: status -2, value stays -1.0, no moments
    for (int k = tid; k < pb.M; k += blockDim.x) sm.mom[k] = __longlong_as_double(0x7ff8000000000000ll);
    if (tid == 0) {
      sm.value = -1.0;
      sm.status = -2;
    }
    __syncthreads();
    return;
  }
  if (tid < 2 * D) {
    unsigned long long acc = 0ull;
    const unsigned long long *pu = (const unsigned long long *)part_base;
    for (int s = 0; s < n_seg; ++s) acc += __ldcg(pu + (size_t)s * part_len + tid);
    // remove S copies of bits(M) (mod 2^64: exact), then fixed point -> double
    const bool is_sq = tid >= D;
    const unsigned long long mb = (unsigned long long)__double_as_longlong(is_sq ? pb.magic_sq : pb.magic_sum);
    const long long fixed = (long long)(acc - (unsigned long long)pb.S * mb);
    sm.tot[tid] = __dmul_rn((double)fixed, is_sq ? pb.scale_sq : pb.scale_sum);
  }
  __syncthreads();
  const double S = (double)pb.S;
  if (tid < D) {
    const double mean = __ddiv_rn(sm.tot[tid], S);
    sm.mom[tid] = mean;
    if (pb.obj == SMM_OBJ_NORM_MV) {
      // sum (x - mean)^2 = sum x^2 - mean * sum x
      const double ss = __dsub_rn(sm.tot[D + tid], __dmul_rn(mean, sm.tot[tid]));
      sm.mom[D + tid] = __ddiv_rn(ss, S - 1.0);
    }
  }
  __syncthreads();
  if (tid < 32) {  // value = mean_k ((sim_k - data_k) / w_k)^2, divisions in parallel, summed in moment order
    double acc = 0.0;
    for (int k0 = 0; k0 < pb.M; k0 += 32) {
      const int k = k0 + tid;
      double d2 = 0.0;
      if (k < pb.M) {
        const double d = __ddiv_rn(__dsub_rn(sm.mom[k], pb.data[k]), pb.w[k]);
        d2 = __dmul_rn(d, d);
      }
      const int n = pb.M - k0 < 32 ? pb.M - k0 : 32;
      for (int l = 0; l < n; ++l) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, d2, l));
    }
    if (tid == 0) {
      sm.value = __ddiv_rn(acc, (double)pb.M);
      sm.status = 1;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void slow_spin(double seconds) {
  // objfunc_norm_slow: sleep(0.1) (ObjExamples.jl:130)
  const unsigned long long t0 = gtimer();
  const unsigned long long ns = (unsigned long long)(seconds * 1e9);
  while (gtimer() - t0 < ns) __nanosleep(20000);
}

// ------------------------------------------------------------------------------------------------
// doAcceptReject! (AlgoBGP.jl:324-392) + set_eval! (:220-245) for chain c with the evaluation in
// sm.value / sm.status / sm.mom / sm.pp.  Writes the trace slot, the last-accepted record (la_cur),
// the published record (la_pub) and -- fused multi-GPU mode -- the record and its value into every
// rank's gather buffer (peer stores over NVLink).
// ------------------------------------------------------------------------------------------------
__device__ void accept_and_store(const DevProblem &pb, const DevState &st, EvalSmem &sm, int c, int gc, int iter,
                                 bool fused) {
  const int tid = threadIdx.x;
  const int P = pb.P, M = pb.M, R = rec_len(P, M), L = pb.L;
  double *la = st.la_cur + (size_t)c * R;
  double *pub = st.la_pub + (size_t)c * R;
  const size_t slot = (size_t)(iter - 1) * L + c;
  if (tid == 0) {
    const double value = sm.value;
    double prob;
    int accepted, status = sm.status;
    if (iter == 1) {
      prob = 1.0;
      accepted = 1;
      status = 1;
    } else {
      const double old_value = __ldcg(la);
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
    const int n_noex = __ldcg(st.n_noex + c) + 1, n_acc = __ldcg(st.n_acc + c) + accepted;
    st.n_noex[c] = n_noex;
    st.n_acc[c] = n_acc;
    const double rate = __ddiv_rn((double)n_acc, (double)n_noex);
    st.accept_rate[c] = rate;
    if (iter > 1 && iter % pb.sigma_update_steps == 0) {
      const double s = __ldcg(st.sigma + c);
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
      curr = accepted ? value : __ldcg(st.t_curr + prev);
      const double bprev = __ldcg(st.t_best + prev);
      if (value < bprev) {
        best = value;
        best_id = iter;
      } else {
        best = bprev;
        best_id = __ldcg(st.t_bestid + prev);
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
    sm.acc = accepted;
    sm.prob = prob;
    sm.status = status;
  }
  __syncthreads();
  // trace rows + last-accepted record (coalesced over threads)
  for (int k = tid; k < P; k += blockDim.x) st.t_params[slot * P + k] = sm.pp[k];
  for (int k = tid; k < M; k += blockDim.x) st.t_mom[slot * M + k] = sm.mom[k];
  const int par = fused ? (iter & 1) : 0;
  for (int k = tid; k < R; k += blockDim.x) {
    double v;
    if (sm.acc) {
      v = k == 0 ? sm.value : k == 1 ? sm.prob : k == 2 ? (double)sm.status : k < 3 + P ? sm.pp[k - 3] : sm.mom[k - 3 - P];
      la[k] = v;
    } else {
      v = __ldcg(la + k);
    }
    pub[k] = v;
    if (fused) {
      // the all-gather: one coalesced row per peer, straight into its gather buffer
      for (int r = 0; r < pb.world; ++r) st.peer_la_all[r][((size_t)par * pb.N + gc) * R + k] = v;
      if (k == 0)
        for (int r = 0; r < pb.world; ++r) st.peer_val_all[r][(size_t)par * pb.N + gc] = v;
    } else if (k == 0) {
      st.val_all[gc] = v;  // compact copy of the values for the exchange step
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// exchangeMoves! (AlgoBGP.jl:647-691): the sequential pair loop, run level-parallel over the schedule of
// iteration `iter` (pairs inside a level share no chain).  val/own/exch live in shared memory.
// ------------------------------------------------------------------------------------------------
__device__ unsigned exchange_levels(const DevProblem &pb, const DevState &st, int sched_idx, int n_s, double *val,
                                    unsigned short *own, unsigned short *exch) {
  const int tid = threadIdx.x, nthr = blockDim.x;
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
        const unsigned short oi = own[i];
        own[i] = own[j];
        own[j] = oi;
        exch[i] = (unsigned short)(j + 1);
        exch[j] = (unsigned short)(i + 1);
        ++n_swaps;
      }
    }
    __syncthreads();
  }
  return n_swaps;
}

// swap_ev_ij! (:734-749) for one chain this rank owns: set_eval!(ci, ej) + set_exchanged!(ci, j).
// One warp; `src` is the record that ended up on this chain.
__device__ void exchange_apply_chain(const DevProblem &pb, const DevState &st, int iter, int c, int partner,
                                     const double *src, int lane) {
  const int P = pb.P, M = pb.M, R = rec_len(P, M), L = pb.L;
  double *la = st.la_cur + (size_t)c * R;
  const size_t slot = (size_t)(iter - 1) * L + c;
  for (int k = lane; k < R; k += 32) {
    const double v = __ldcg(src + k);
    la[k] = v;
    if (k >= 3 && k < 3 + P) st.t_params[slot * P + (k - 3)] = v;
    if (k >= 3 + P) st.t_mom[slot * M + (k - 3 - P)] = v;
  }
  if (lane == 0) {
    const double value = __ldcg(src);
    // this iteration no longer counts towards the acceptance rate (exchanged != 0)
    st.n_noex[c] = __ldcg(st.n_noex + c) - 1;
    st.n_acc[c] = __ldcg(st.n_acc + c) - (int)__ldcg(st.t_acc + slot);
    st.t_value[slot] = value;
    st.t_prob[slot] = __ldcg(src + 1);
    st.t_status[slot] = (int)__ldcg(src + 2);
    st.t_acc[slot] = 1;  // the swapped-in eval is an accepted one
    st.t_curr[slot] = value;
    const double bprev = __ldcg(st.t_best + slot - L);
    if (value < bprev) {
      st.t_best[slot] = value;
      st.t_bestid[slot] = iter;
    } else {
      st.t_best[slot] = bprev;
      st.t_bestid[slot] = __ldcg(st.t_bestid + slot - L);
    }
    st.t_exch[slot] = partner;
  }
}

// ------------------------------------------------------------------------------------------------
// (A) multi-launch mode
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEvalThreads) bgp_eval_kernel(DevProblem pb, DevState st, int iter, int n_split,
                                                                int part_len) {
  __shared__ EvalSmem sm;
  const int c = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  const int gc = pb.chain0 + c;
  const size_t stamp = (size_t)c * n_split + split;
  PHASE_STAMP(stamp, 0);
  load_logtab(sm.logtab);
  __syncthreads();
  block_proposal(pb, st, sm, c, gc, iter, split == 0);
  PHASE_STAMP(stamp, 1);
  double *part_base = st.partials + (size_t)c * n_split * part_len;
  if (pb.obj == SMM_OBJ_FAILS) {
    if (split != 0) return;
  } else {
    if (pb.obj == SMM_OBJ_NORM_SLOW) slow_spin(pb.slow_seconds);
    const int nb = (pb.S + 1) >> 1;
    const int j0 = (int)(((long long)nb * split) / n_split), j1 = (int)(((long long)nb * (split + 1)) / n_split);
    simulate_norm_segment(pb, sm, j0, j1, (uint32_t)gc, (uint32_t)iter, part_base + (size_t)split * part_len);
    PHASE_STAMP(stamp, 2);
    if (n_split > 1) {
      if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(st.arrive + c, 1u);
        sm.is_last = (prev == (unsigned)(n_split - 1));
        if (sm.is_last) st.arrive[c] = 0u;  // re-arm for the next iteration
        __threadfence();
      }
      __syncthreads();
      if (!sm.is_last) return;
    }
  }
  reduce_and_finalize(pb, sm, part_base, n_split, part_len);
  accept_and_store(pb, st, sm, c, gc, iter, false);
  PHASE_STAMP(stamp, 3);
}

// dynamic smem: val[N] (double) own[N] exch[N] (u16)
__global__ void __launch_bounds__(kExchThreads) bgp_exchange_kernel(DevProblem pb, DevState st, int iter,
                                                                    int sched_idx, int n_s) {
  extern __shared__ double smem_d[];
  const int N = pb.N, tid = threadIdx.x, nthr = blockDim.x;
  const int R = rec_len(pb.P, pb.M), L = pb.L;
  double *val = smem_d;
  unsigned short *own = (unsigned short *)(val + N), *exch = own + N;
  for (int i = tid; i < N; i += nthr) {
    val[i] = st.la_all[(size_t)i * R];
    own[i] = (unsigned short)i;
    exch[i] = 0;
  }
  __syncthreads();
  const unsigned n_swaps = exchange_levels(pb, st, sched_idx, n_s, val, own, exch);
  if (pb.chain0 == 0 && n_swaps) atomicAdd(&st.counters[1], (unsigned long long)n_swaps);
  for (int c = tid / 32; c < L; c += nthr / 32) {  // one warp per chain
    const int gc = pb.chain0 + c;
    const int partner = exch[gc];
    if (partner == 0) continue;
    exchange_apply_chain(pb, st, iter, c, partner, st.la_all + (size_t)own[gc] * R, tid & 31);
  }
}

// ------------------------------------------------------------------------------------------------
// (B) persistent mode
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) { return *(const volatile unsigned *)p; }
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  return *(const volatile unsigned long long *)p;
}

constexpr unsigned long long kSpinTimeoutNs = 4000000000ull;  // 4 s: a stuck peer becomes an error, not a hang

// Grid barrier (CTA 0 is the master).  With `cross`, the master also exchanges sequence flags with every
// peer GPU before releasing, so the records stored into our gather buffer by the peers are complete.
// Returns false if the kernel must abort (timeout somewhere).
__device__ bool grid_barrier(const DevProblem &pb, const DevState &st, unsigned &gen, bool cross,
                             unsigned long long &seq) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    bool ok = true;
    const unsigned G = gridDim.x;
    const unsigned target = ++gen;
    if (cross) {
      ++seq;
      __threadfence_system();  // this CTA's peer stores (all threads, ordered by the bar.sync above)
    } else {
      __threadfence();
    }
    if (blockIdx.x == 0) {
      const unsigned long long t0 = gtimer();
      unsigned spins = 0;
      while (ld_volatile_u32(&st.bar->arrive) < G - 1) {
        if ((++spins & 1023u) == 0 &&
            (gtimer() - t0 > kSpinTimeoutNs || (*(volatile int *)st.err & kErrTimeout))) {
          atomicOr(st.err, kErrTimeout);
          ok = false;
          break;
        }
      }
      st.bar->arrive = 0;
      if (cross && ok) {
        __threadfence_system();
        for (int r = 0; r < pb.world; ++r) *(volatile unsigned long long *)(st.peer_flags[r] + pb.rank) = seq;
        for (int r = 0; r < pb.world && ok; ++r) {
          spins = 0;
          while (ld_volatile_u64(st.flags + r) < seq) {
            if ((++spins & 1023u) == 0 && gtimer() - t0 > kSpinTimeoutNs) {
              atomicOr(st.err, kErrTimeout);
              ok = false;
              break;
            }
          }
        }
        __threadfence_system();
      }
      __threadfence();
      *(volatile unsigned *)&st.bar->gen = target;
    } else {
      atomicAdd(&st.bar->arrive, 1u);
      const unsigned long long t0 = gtimer();
      unsigned spins = 0;
      while ((int)(ld_volatile_u32(&st.bar->gen) - target) < 0) {
        if ((++spins & 1023u) == 0 && gtimer() - t0 > 2 * kSpinTimeoutNs) {
          atomicOr(st.err, kErrTimeout);
          break;
        }
      }
    }
    __threadfence();
    if (*(volatile int *)st.err & kErrTimeout) ok = false;
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// block that holds flattened index x when T items are split evenly over G blocks: [floor(b*T/G), floor((b+1)*T/G))
__device__ __forceinline__ long long block_of(long long x, long long T, long long G) {
  return ((x + 1) * G + T - 1) / T - 1;
}

// exchange of iteration `pit` for the chains CTA b owns (replicated computation of the pair loop)
__device__ void persistent_exchange(const DevProblem &pb, const DevState &st, int pit, int sched_iter0, int n_s,
                                    bool fused, double *val, unsigned short *own, unsigned short *exch) {
  const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
  const int N = pb.N, L = pb.L, R = rec_len(pb.P, pb.M);
  const int par = fused ? (pit & 1) : 0;
  const double *la_all = st.la_all + (size_t)par * N * R;
  const double *val_all = st.val_all + (size_t)par * N;
  for (int i = tid; i < N; i += blockDim.x) {
    val[i] = __ldcg(val_all + i);
    own[i] = (unsigned short)i;
    exch[i] = 0;
  }
  __syncthreads();
  const unsigned n_swaps = exchange_levels(pb, st, pit - sched_iter0, n_s, val, own, exch);
  if (b == 0 && pb.chain0 == 0 && n_swaps) atomicAdd(&st.counters[1], (unsigned long long)n_swaps);
  for (int c = b; c < L; c += G) {
    const int gc = pb.chain0 + c;
    const int partner = exch[gc];
    if (partner != 0 && tid < 32) exchange_apply_chain(pb, st, pit, c, partner, la_all + (size_t)own[gc] * R, tid);
  }
  __syncthreads();
}

// dynamic smem: val[N] (double) own[N] exch[N] (u16)
__global__ void __launch_bounds__(kEvalThreads, 8) bgp_persistent_kernel(DevProblem pb, DevState st, int iter0,
                                                                      int n_iters, int sched_iter0, int n_s,
                                                                      int part_len, int max_seg) {
  __shared__ EvalSmem sm;
  extern __shared__ double smem_d[];
  const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x;
  const int N = pb.N, L = pb.L, P = pb.P, R = rec_len(pb.P, pb.M);
  const bool fused = pb.world > 1;
  double *val = smem_d;
  unsigned short *own = (unsigned short *)(val + N), *exch = own + N;
  load_logtab(sm.logtab);
  unsigned gen = ld_volatile_u32(&st.bar->gen);
  unsigned long long seq = ld_volatile_u64(st.sync_seq);
  __syncthreads();
  const int nb = (pb.S + 1) >> 1;
  const long long T = (long long)L * nb;
  const long long Gw = T < G ? T : G;  // CTAs that take a share of the draw space (all of them unless T is tiny)
  const long long lo = b < Gw ? (T * b) / Gw : 0, hi = b < Gw ? (T * (b + 1)) / Gw : 0;
  const bool owner = b < L;  // owner CTAs handle chains b, b + G, ...

  for (int it = iter0; it < iter0 + n_iters; ++it) {
    // ---- exchange of iteration it-1 (AlgoBGP.jl:637), then this iteration's proposals ----
    if (owner) {
      if (N > 1 && it - 1 >= 2 && it > iter0) persistent_exchange(pb, st, it - 1, sched_iter0, n_s, fused, val, own, exch);
      for (int c = b; c < L; c += G) {
        block_proposal(pb, st, sm, c, pb.chain0 + c, it, true);
        for (int k = tid; k < P; k += blockDim.x) st.pp[(size_t)c * P + k] = sm.pp[k];
        __syncthreads();
      }
    }
    if (!grid_barrier(pb, st, gen, false, seq)) return;

    // ---- an equal share of the flattened (chain, Philox block) space; finish the chains we complete ----
    PHASE_STAMP(b * 2 + (it & 1), 0);
    if (pb.obj == SMM_OBJ_NORM_SLOW) slow_spin(pb.slow_seconds);
    for (long long x = lo; x < hi;) {
      const int c = (int)(x / nb);
      const long long cbase = (long long)c * nb;
      const int j0 = (int)(x - cbase);
      const long long xe = (cbase + nb < hi) ? cbase + nb : hi;
      const int j1 = (int)(xe - cbase);
      const int b_first = (int)block_of(cbase, T, Gw), b_last = (int)block_of(cbase + nb - 1, T, Gw);
      const int n_seg = b_last - b_first + 1;
      double *part_base = st.partials + (size_t)c * max_seg * part_len;
      for (int k = tid; k < P; k += blockDim.x) sm.pp[k] = __ldcg(st.pp + (size_t)c * P + k);
      __syncthreads();
      if (pb.obj != SMM_OBJ_FAILS)
        simulate_norm_segment(pb, sm, j0, j1, (uint32_t)(pb.chain0 + c), (uint32_t)it,
                              part_base + (size_t)(b - b_first) * part_len);
      PHASE_STAMP(b * 2 + (it & 1), 1);
      if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(st.arrive + c, 1u);
        sm.is_last = (prev == (unsigned)(n_seg - 1));
        if (sm.is_last) st.arrive[c] = 0u;
        __threadfence();
      }
      __syncthreads();
      if (sm.is_last) {
        reduce_and_finalize(pb, sm, part_base, n_seg, part_len);
        accept_and_store(pb, st, sm, c, pb.chain0 + c, it, fused);
      }
      __syncthreads();
      x = xe;
    }
    PHASE_STAMP(b * 2 + (it & 1), 2);
    if (!grid_barrier(pb, st, gen, fused, seq)) return;
    PHASE_STAMP(b * 2 + (it & 1), 3);
  }
  // ---- exchange of the last iteration of this launch ----
  const int pit = iter0 + n_iters - 1;
  if (owner && N > 1 && pit >= 2) persistent_exchange(pb, st, pit, sched_iter0, n_s, fused, val, own, exch);
  if (b == 0 && tid == 0) *st.sync_seq = seq;
}

// ------------------------------------------------------------------------------------------------
// batched bare objective: grid (n_split, B)  (evaluateObjective, mprob.jl:175-205)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEvalThreads) objective_kernel(DevProblem pb, const double *params, int noseed,
                                                                 uint32_t rep0, int n_split, int part_len,
                                                                 double *partials, unsigned *arrive, double *value,
                                                                 double *moments, int *status) {
  __shared__ EvalSmem sm;
  const int bi = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  load_logtab(sm.logtab);
  for (int k = tid; k < pb.P; k += blockDim.x) sm.pp[k] = params[(size_t)bi * pb.P + k];
  __syncthreads();
  pb.noseed = noseed;
  double *part_base = partials + (size_t)bi * n_split * part_len;
  if (pb.obj == SMM_OBJ_FAILS) {
    if (split != 0) return;
  } else {
    if (pb.obj == SMM_OBJ_NORM_SLOW) slow_spin(pb.slow_seconds);
    const int nb = (pb.S + 1) >> 1;
    const int j0 = (int)(((long long)nb * split) / n_split), j1 = (int)(((long long)nb * (split + 1)) / n_split);
    simulate_norm_segment(pb, sm, j0, j1, (uint32_t)bi, rep0 + (uint32_t)bi, part_base + (size_t)split * part_len);
    if (n_split > 1) {
      if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(arrive + bi, 1u);
        sm.is_last = (prev == (unsigned)(n_split - 1));
        if (sm.is_last) arrive[bi] = 0u;
        __threadfence();
      }
      __syncthreads();
      if (!sm.is_last) return;
    }
  }
  reduce_and_finalize(pb, sm, part_base, n_split, part_len);
  if (tid == 0) {
    value[bi] = sm.value;
    status[bi] = sm.status;
  }
  for (int k = tid; k < pb.M; k += blockDim.x) moments[(size_t)bi * pb.M + k] = sm.mom[k];
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
      const int pos = cnt[l]++;
      ij[2 * pos] = pi[t];
      ij[2 * pos + 1] = pj[t];
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
size_t exch_smem_bytes(int N) { return sizeof(double) * (size_t)N + sizeof(unsigned short) * 2 * (size_t)N; }

cudaError_t configure_kernels(int N, int n_s) {
  cudaError_t e = cudaFuncSetAttribute(bgp_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)pairs_smem_bytes(N, n_s));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(bgp_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exch_smem_bytes(N));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(bgp_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)exch_smem_bytes(N));
}

int eval_max_blocks_per_sm() {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bgp_eval_kernel, kEvalThreads, 0);
  return n;
}
int persistent_max_blocks_per_sm(int N) {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bgp_persistent_kernel, kEvalThreads, exch_smem_bytes(N));
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
cudaError_t launch_persistent(const DevProblem &pb, const DevState &st, int iter0, int n_iters, int sched_iter0,
                              int n_s, int part_len, int max_seg, int grid, cudaStream_t s) {
  DevProblem pbc = pb;
  DevState stc = st;
  void *args[] = {&pbc, &stc, &iter0, &n_iters, &sched_iter0, &n_s, &part_len, &max_seg};
  return cudaLaunchCooperativeKernel((void *)bgp_persistent_kernel, dim3(grid), dim3(kEvalThreads), args,
                                     exch_smem_bytes(pb.N), s);
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
