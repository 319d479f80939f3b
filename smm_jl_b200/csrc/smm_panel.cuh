// smm_panel.cuh -- the dynamic-panel objective (BASELINE config C4).  Included by smm_kernels.cu
// inside namespace smm, after the shared device functions (group_proposal, group_distance, group_accept_store).
//
// There is no upstream code for this simulator; the spec is oracle/smm_oracle.cpp::objfunc_panel (a user model
// written against evaluateObjective, mprob.jl:175-188, with the objfunc_norm template ObjExamples.jl:59-116):
//   theta = (rho, beta[K], phi[K], sigma_alpha, sigma_eps, mu0);  individual i owns row i of the Zsim stream:
//   normal 0 = a_i, 1..K = initial regressor shocks, then per period t = 1..T the K regressor shocks and eps_t;
//   alpha = fma(sigma_alpha, a, mu0); y_0 = alpha/(1-rho); x_k0 = eta/sqrt(1-phi_k^2);
//   x_kt = fma(phi_k, x_k,t-1, eta_kt); y_t = fma(sigma_eps, eps_t, fma(rho, y_t-1, alpha) + sum_k fma(beta_k x_kt));
//   4K+8 moments pooled over (i, t = 1..T): mean y, autocov_y lag 0..6, cov(y_t, x_kt), cov(y_t, x_k,t-1),
//   autocov_xk lag 1, var x_k.
//
// Execution: ONE THREAD PER INDIVIDUAL.  The T-step recurrence, its shocks (Philox + Box-Muller) and the
// individual's 14 + 6K raw sums (sequential in t: a fixed order) live in registers; nothing of the panel is ever
// stored (the oracle materialises y[i,t] and x[k,i,t]: 2*8*(K+1)*N*T = 36 MB per evaluation at C4).
// Work distribution: warps pull units of 32 individuals from a global queue over the flattened (evaluation,
// individual) space, so the grid is one resident wave whatever the chain count.  Pooling over individuals is
// ORDER INVARIANT: each per-individual sum is split exactly into two fixed-point words (grids 2^-Fhi and
// 2^-(Fhi+40), sized by the host from the parameter box) that are added as 64-bit integers -- in lane-owned
// registers, then with global atomics when a warp leaves an evaluation.  Chains with equal parameters therefore
// get bit-equal values however the units were scheduled (the exchange step compares values), and 1-GPU and
// N-GPU runs agree to the bit.  The warp that completes an evaluation turns the sums into centred moments
// (the algebra is in panel_finalize), then distance + doAcceptReject! + set_eval! as for the other objectives.

struct PanelWork {
  const double *params;       // [n_eval][P]
  unsigned long long *acc;    // [n_eval][2 * NA] fixed-point words (hi, lo) per raw sum; zero on entry
  unsigned *done;             // [n_eval] units finished; zero on entry
  unsigned *unit_ctr;         // queue head; zero on entry
  int n_eval, units_per_eval;
  int noseed;
  uint32_t uid0, rep0;        // evaluation e draws from stream (uid0 + uid_stride * e, rep0 + rep_stride * e) when noseed
  uint32_t uid_stride, rep_stride;
  int iter;                   // chain mode: BGP iteration (>= 1); batch mode: 0
  double *value, *moments;    // batch mode outputs
  int *status;
};

// the normals of one stream row, in order
struct NormalRow {
  uint32_t j, row, c2, c3;
  double pz;   // second normal of the last block, not yet consumed
  bool pend;
  __device__ __forceinline__ void block(const DevProblem &pb, const smm_logent *tab, double &z0, double &z1) {
    smm_normal_pair_tab(philox_sim(pb, j, row, c2, c3), tab, &z0, &z1);
    ++j;
  }
  // NEED normals into zt[0..NEED) with static indices; PEND = a pending normal comes first
  template <int NEED, bool PEND>
  __device__ __forceinline__ void fill_static(const DevProblem &pb, const smm_logent *tab, double *zt) {
    constexpr int first = PEND ? 1 : 0;
    constexpr int rem = NEED - first, full = rem / 2;
    if (PEND) zt[0] = pz;
#pragma unroll
    for (int b = 0; b < full; ++b) block(pb, tab, zt[first + 2 * b], zt[first + 2 * b + 1]);
    if (rem & 1) {
      block(pb, tab, zt[first + 2 * full], pz);
      pend = true;
    } else {
      pend = false;
    }
  }
  template <int NEED>
  __device__ __forceinline__ void fill(const DevProblem &pb, const smm_logent *tab, double *zt) {
    if (pend)  // warp uniform: every lane is at the same position of its row
      fill_static<NEED, true>(pb, tab, zt);
    else
      fill_static<NEED, false>(pb, tab, zt);
  }
  __device__ void fill_dynamic(const DevProblem &pb, const smm_logent *tab, double *zt, int need) {
    int n = 0;
    if (pend) {
      zt[n++] = pz;
      pend = false;
    }
    while (n < need) {
      double z0, z1;
      block(pb, tab, z0, z1);
      zt[n++] = z0;
      if (n < need) {
        zt[n++] = z1;
      } else {
        pz = z1;
        pend = true;
      }
    }
  }
};

// Simulate individual `row` and write its NA raw sums to out[a * ostride], a = 0..NA-1:
//   0        Sy   = sum_{t=1..T} y_t
//   1 + l    Syy_l = sum_{t>=max(l,1)} y_t y_{t-l}                    l = 0..6
//   7 + l    HG_l = sum_{t>=l} y_t + sum_{t>=l} y_{t-l}              l = 1..6   (for the centring)
//   14 + k          Sx_k   = sum x_kt          14 + K + k   Syx_k  = sum y_t x_kt
//   14 + 2K + k     Syxl_k = sum y_t x_k,t-1   14 + 3K + k  Sxxl_k = sum x_kt x_k,t-1
//   14 + 4K + k     Sxx_k  = sum x_kt^2        14 + 5K + k  XL_k   = sum x_k,t-1
// KT > 0: K known at compile time (everything in registers); KT == 0: run-time K <= kPanelMaxK (local memory).
// SACC: the 5K per-regressor sums are accumulated IN PLACE in `out` (the warp's shared staging tile, one conflict-free
// column per lane) instead of in registers -- 80 registers less at K = 8, which is what lets a third CTA live on the SM.
template <int KT, bool SACC = false>
__device__ __forceinline__ void panel_individual(const DevProblem &pb, const smm_logent *tab, const double *th, int Krt,
                                                 int T, uint32_t row, uint32_t c2, uint32_t c3, double *out,
                                                 int ostride) {
  constexpr int KM = KT > 0 ? KT : kPanelMaxK;
  constexpr int KA = SACC ? 1 : KM;  // register copies of the per-regressor sums (unused with SACC)
  const int K = KT > 0 ? KT : Krt;
  const double rho = th[0];
  const double *beta = th + 1, *phi = th + 1 + K;
  const double sig_a = th[1 + 2 * K], sig_e = th[2 + 2 * K], mu0 = th[3 + 2 * K];
  NormalRow nr{0u, row, c2, c3, 0.0, false};
  double zt[KM + 1];
  double x[KM], xp[KM], x0[KM], sx[KA], syx[KA], syxl[KA], sxxl[KA], sxx[KA];
  if constexpr (KT > 0)
    nr.template fill<KM + 1>(pb, tab, zt);
  else
    nr.fill_dynamic(pb, tab, zt, K + 1);
  const double alpha = __fma_rn(sig_a, zt[0], mu0);
  const double y0 = __ddiv_rn(alpha, __dsub_rn(1.0, rho));
#pragma unroll
  for (int k = 0; k < K; ++k) {
    x[k] = __ddiv_rn(zt[1 + k], __dsqrt_rn(__fma_rn(-phi[k], phi[k], 1.0)));
    x0[k] = x[k];
    if constexpr (SACC) {
      out[(14 + k) * ostride] = out[(14 + K + k) * ostride] = out[(14 + 2 * K + k) * ostride] =
          out[(14 + 3 * K + k) * ostride] = out[(14 + 4 * K + k) * ostride] = 0.0;
    } else {
      sx[k] = syx[k] = syxl[k] = sxxl[k] = sxx[k] = 0.0;
    }
  }
  double yl[7] = {y0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // y_{t-1}, y_{t-2}, ...: zero where t-l < 0
  double sy = 0.0, syy[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double hp[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // hp[l-2] = sum_{t=1..l-1} y_t, l = 2..6
#pragma unroll 1
  for (int t = 1; t <= T; ++t) {
    if constexpr (KT > 0)
      nr.template fill<KM + 1>(pb, tab, zt);
    else
      nr.fill_dynamic(pb, tab, zt, K + 1);
    double xb = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      xp[k] = x[k];
      x[k] = __fma_rn(phi[k], xp[k], zt[k]);
      xb = __fma_rn(beta[k], x[k], xb);
    }
    const double y = __fma_rn(sig_e, zt[K], __dadd_rn(__fma_rn(rho, yl[0], alpha), xb));
    sy = __dadd_rn(sy, y);
    syy[0] = __fma_rn(y, y, syy[0]);
#pragma unroll
    for (int l = 1; l <= 6; ++l) syy[l] = __fma_rn(y, yl[l - 1], syy[l]);
#pragma unroll
    for (int l = 2; l <= 6; ++l) hp[l - 2] = __dadd_rn(hp[l - 2], t < l ? y : 0.0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if constexpr (SACC) {
        double *o = out + (14 + k) * ostride;
        o[0] = __dadd_rn(o[0], x[k]);
        o[K * ostride] = __fma_rn(y, x[k], o[K * ostride]);
        o[2 * K * ostride] = __fma_rn(y, xp[k], o[2 * K * ostride]);
        o[3 * K * ostride] = __fma_rn(x[k], xp[k], o[3 * K * ostride]);
        o[4 * K * ostride] = __fma_rn(x[k], x[k], o[4 * K * ostride]);
      } else {
        sx[k] = __dadd_rn(sx[k], x[k]);
        sxx[k] = __fma_rn(x[k], x[k], sxx[k]);
        sxxl[k] = __fma_rn(x[k], xp[k], sxxl[k]);
        syx[k] = __fma_rn(y, x[k], syx[k]);
        syxl[k] = __fma_rn(y, xp[k], syxl[k]);
      }
    }
#pragma unroll
    for (int l = 6; l >= 1; --l) yl[l] = yl[l - 1];
    yl[0] = y;
  }
  // yl[0..6] = y_T, y_{T-1}, ..., y_{T-6}
  out[0] = sy;
#pragma unroll
  for (int l = 0; l <= 6; ++l) out[(1 + l) * ostride] = syy[l];
  {
    // HG_l = (Sy - sum_{t<l} y_t) + (y_0 + Sy - sum_{s>T-l} y_s)
    const double base = __dadd_rn(__dadd_rn(sy, sy), y0);
    double tail = 0.0;
#pragma unroll
    for (int l = 1; l <= 6; ++l) {
      tail = __dadd_rn(tail, yl[l - 1]);
      const double head = l >= 2 ? hp[l - 2] : 0.0;
      out[(7 + l) * ostride] = __dsub_rn(base, __dadd_rn(head, tail));
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if constexpr (SACC) {
      out[(14 + 5 * K + k) * ostride] = __dsub_rn(__dadd_rn(x0[k], out[(14 + k) * ostride]), x[k]);
    } else {
      out[(14 + k) * ostride] = sx[k];
      out[(14 + K + k) * ostride] = syx[k];
      out[(14 + 2 * K + k) * ostride] = syxl[k];
      out[(14 + 3 * K + k) * ostride] = sxxl[k];
      out[(14 + 4 * K + k) * ostride] = sxx[k];
      out[(14 + 5 * K + k) * ostride] = __dsub_rn(__dadd_rn(x0[k], sx[k]), x[k]);
    }
  }
}

// raw sums (fs.tot[NA]) -> the 4K+8 centred moments (fs.mom), one warp.  With n = N_ind * T, my = Sy/n, mx = Sx/n:
//   sum (y_t-my)(y_{t-l}-my) = Syy_l - my HG_l + cnt_l my^2     (cnt_l = N_ind (T - max(l-1, 0)) terms)
//   sum (y_t-my)(x_t-mx)     = Syx - my Sx          sum (y_t-my)(x_{t-1}-mx) = Syxl - my XL
//   sum (x_t-mx)(x_{t-1}-mx) = Sxxl - mx XL         sum (x_t-mx)^2           = Sxx - mx Sx
__device__ void panel_moments(const DevProblem &pb, const FinScratch &fs, int lane) {
  const int K = pb.panel_K, T = pb.panel_T, M = pb.M;
  const double NI = (double)pb.panel_N;
  const double n = __dmul_rn(NI, (double)T);
  const double *tot = fs.tot;
  const double my = __ddiv_rn(tot[0], n);
  for (int m = lane; m < M; m += 32) {
    double v;
    if (m == 0) {
      v = my;
    } else if (m < 8) {
      const int l = m - 1;
      const double cnt = __dmul_rn(NI, (double)(T - (l > 1 ? l - 1 : 0)));
      const double hg = l == 0 ? __dadd_rn(tot[0], tot[0]) : tot[7 + l];
      const double c = __dadd_rn(__dsub_rn(tot[1 + l], __dmul_rn(my, hg)), __dmul_rn(cnt, __dmul_rn(my, my)));
      v = __ddiv_rn(c, n);
    } else {
      const int q = (m - 8) / K, k = (m - 8) - q * K;
      const double sxk = tot[14 + k], xl = tot[14 + 5 * K + k];
      const double mx = __ddiv_rn(sxk, n);
      double c;
      if (q == 0)
        c = __dsub_rn(tot[14 + K + k], __dmul_rn(my, sxk));
      else if (q == 1)
        c = __dsub_rn(tot[14 + 2 * K + k], __dmul_rn(my, xl));
      else if (q == 2)
        c = __dsub_rn(tot[14 + 3 * K + k], __dmul_rn(mx, xl));
      else
        c = __dsub_rn(tot[14 + 4 * K + k], __dmul_rn(mx, sxk));
      v = __ddiv_rn(c, n);
    }
    fs.mom[m] = v;
  }
  __syncwarp();
}

// split v exactly into round(v 2^Fhi) and round((v - hi 2^-Fhi) 2^Flo)
__device__ __forceinline__ void panel_split(const DevProblem &pb, double v, unsigned long long &hi,
                                            unsigned long long &lo) {
  const long long h = __double2ll_rn(__dmul_rn(v, pb.pan_hi_scale));
  const double r = __fma_rn(-(double)h, pb.pan_hi_inv, v);  // exact
  hi += (unsigned long long)h;
  lo += (unsigned long long)__double2ll_rn(__dmul_rn(r, pb.pan_lo_scale));
}

constexpr int kPanelStageStride = 33;  // doubles per raw sum in a warp's staging tile (32 lanes + 1: conflict free)

// dynamic smem per warp: stage[NA][33] doubles; reused as {theta[P] | tot[NA] | mom[M] | value[2] | flags[2]} when a
// warp finalises an evaluation.  Per warp additionally theta[P] of the current evaluation.
__host__ __device__ inline size_t panel_warp_smem_doubles(int K, int P) {
  return (size_t)panel_na(K) * kPanelStageStride + (size_t)P;
}

template <int KT, int MINB, bool SACC = false>
__global__ void __launch_bounds__(kPanelThreads, MINB)
    panel_sim_kernel(DevProblem pb, DevState st, PanelWork w) {
  constexpr int KM = KT > 0 ? KT : kPanelMaxK;
  constexpr int NAM = 14 + 6 * KM;        // raw sums (compile-time bound)
  constexpr int QM = (NAM + 31) / 32;     // raw sums owned per lane
  __shared__ smm_logent logtab[1 << SMM_LOG_BITS];
  extern __shared__ double smem_d[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = KT > 0 ? KT : pb.panel_K, T = pb.panel_T, NI = pb.panel_N, P = pb.P, M = pb.M;
  const int NA = panel_na(K);
  double *stage = smem_d + (size_t)warp * panel_warp_smem_doubles(K, P);
  double *theta = stage + (size_t)NA * kPanelStageStride;
  load_logtab(logtab);
  __syncthreads();
  const int total = w.n_eval * w.units_per_eval;
  unsigned long long ahi[QM], alo[QM];
#pragma unroll
  for (int q = 0; q < QM; ++q) ahi[q] = alo[q] = 0ull;
  int cur = -1, units_cur = 0;
  uint32_t c2 = 0u, c3 = SMM_STREAM_SIM << 28;
  int next = 0;
  if (lane == 0) next = (int)atomicAdd(w.unit_ctr, 1u);
  next = __shfl_sync(0xffffffffu, next, 0);
  for (;;) {
    const int u = next;
    const int e = u < total ? u / w.units_per_eval : -1;
    if (e != cur) {
      if (cur >= 0) {
        // leave evaluation `cur`: add this warp's exact sums to the evaluation's, count its units
        unsigned long long *acc = w.acc + (size_t)cur * 2 * NA;
#pragma unroll
        for (int q = 0; q < QM; ++q) {
          const int a = lane + 32 * q;
          if (a < NA) {
            atomicAdd(acc + 2 * a, ahi[q]);
            atomicAdd(acc + 2 * a + 1, alo[q]);
          }
          ahi[q] = alo[q] = 0ull;
        }
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          const unsigned prev = atom_acq_rel_gpu(w.done + cur, (unsigned)units_cur);
          last = (prev + (unsigned)units_cur == (unsigned)w.units_per_eval);
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          // this warp completed the evaluation: totals -> moments -> distance -> (chain mode) accept/reject + trace
          double *f = stage;
          const FinScratch fs{theta, f, f + NA, f + NA + M, (int *)(f + NA + M + 2)};
          for (int a = lane; a < NA; a += 32) {
            const long long hi = (long long)__ldcg(acc + 2 * a), lo = (long long)__ldcg(acc + 2 * a + 1);
            fs.tot[a] = __fma_rn((double)lo, pb.pan_lo_inv, __dmul_rn((double)hi, pb.pan_hi_inv));
          }
          __syncwarp();
          panel_moments(pb, fs, lane);
          const Grp gw{lane, 32, 0};
          group_distance(pb, gw, fs);
          if (w.iter > 0) {
            group_accept_store(pb, st, gw, fs, cur, global_chain(pb, cur), w.iter, false);
          } else {
            if (lane == 0) {
              w.value[cur] = fs.value[0];
              w.status[cur] = fs.flags[1];
            }
            for (int m = lane; m < M; m += 32) w.moments[(size_t)cur * M + m] = fs.mom[m];
          }
          __syncwarp();
        }
      }
      cur = e;
      units_cur = 0;
      if (e >= 0) {
        for (int k = lane; k < P; k += 32) theta[k] = __ldcg(w.params + (size_t)e * P + k);
        c2 = w.noseed ? w.uid0 + w.uid_stride * (uint32_t)e : 0u;
        c3 = (SMM_STREAM_SIM << 28) | (w.noseed ? ((w.rep0 + w.rep_stride * (uint32_t)e) & SMM_ITER_MASK) : 0u);
        __syncwarp();
      }
    }
    if (e < 0) break;
    if (lane == 0) next = (int)atomicAdd(w.unit_ctr, 1u);  // prefetch: the latency hides behind the simulation
    ++units_cur;
    const int i = (u - e * w.units_per_eval) * 32 + lane;
    if (i < NI) {
      panel_individual<KT, SACC>(pb, logtab, theta, K, T, (uint32_t)i, c2, c3, stage + lane, kPanelStageStride);
    } else {
      for (int a = 0; a < NA; ++a) stage[a * kPanelStageStride + lane] = 0.0;
    }
    __syncwarp();
    // pool the 32 individuals: lane l owns raw sums l, l + 32, ...
#pragma unroll
    for (int q = 0; q < QM; ++q) {
      const int a = lane + 32 * q;
      if (a < NA) {
        const double *row = stage + (size_t)a * kPanelStageStride;
#pragma unroll 8
        for (int s = 0; s < 32; ++s) panel_split(pb, row[s], ahi[q], alo[q]);
      }
    }
    __syncwarp();
    next = __shfl_sync(0xffffffffu, next, 0);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K = 8, EIGHT LANES PER INDIVIDUAL (one regressor per lane): the thread-per-individual kernel above needs 255
// registers (two CTAs of four warps per SM, issue slots 45 % busy: fixed-latency stalls with two warps per scheduler).
// Here a warp simulates 4 individuals at a time, lane (g, k) = individual g, regressor k:
//   * normals: the 8 lanes of an individual compute 8 consecutive Philox blocks of its stream row per round (16
//     Box-Muller normals) into a 32-entry shared ring; a period consumes 9 (eta_0..7 from slot k, eps from slot 8);
//   * recurrence: x_k in lane k; sum_k beta_k x_k by a three-step butterfly over the 8 lanes (every lane gets the same
//     bits; NOT the oracle's sequential fma chain: values agree to ~1e-15 relative, inside the 1e-6 contract and the
//     1e-9 the tests ask for); y is computed redundantly by the 8 lanes;
//   * moments: lane k owns its regressor's six sums; lane l = 0..6 owns Syy_l (its y_{t-l} arrives through a shuffle
//     delay line: one shfl_up per period instead of seven registers) and HG_l; lane 7 owns Sy;
//   * pooling: every raw sum is split exactly into the two fixed-point words and added to the warp's shared 64-bit
//     accumulators (red.shared), which go to the evaluation's global accumulators when the warp leaves it.
// 80-108 registers per thread: 16-24 warps per SM.  Same work queue, same exact pooling, same finalisation as above.
// MEASURED (profiles/panel_variants_r2.txt, ncu prof_panel_v7): issue slots 67 % busy instead of 45 %, but 760 M warp
// instructions per launch instead of 382 M -- the recurrence costs ~116 instructions per warp-period for 4 individuals
// (shuffles, selects, ring addressing, redundant y) against ~180 thread-instructions per individual-period in the
// register-resident kernel -- so C4 runs at 1.02 ms per iteration instead of 0.78 ms.  Kept as SMM_PANEL_VARIANT=6..8
// for the record; the thread-per-individual kernel (variant 2) stays the default.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kLaneK = 8;

// per warp: ring[4][32] | theta[P] (padded) | fin scratch tot[NA] mom[M] value[2] flags[2] | acc u64 [2 NA]
__host__ __device__ inline size_t panel_lanes_warp_doubles(int P, int M) {
  const int NA = panel_na(kLaneK);
  return 128 + (size_t)((P + 1) & ~1) + (size_t)NA + (size_t)M + 4 + 2 * (size_t)NA;
}

template <int MINB>
__global__ void __launch_bounds__(kPanelThreads, MINB) panel_lanes_kernel(DevProblem pb, DevState st, PanelWork w) {
  constexpr int K = kLaneK, NA = 14 + 6 * K;
  __shared__ smm_logent logtab[1 << SMM_LOG_BITS];
  extern __shared__ double smem_d[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = lane & 7, grp = lane >> 3;
  const int T = pb.panel_T, NI = pb.panel_N, P = pb.P, M = pb.M;
  double *base = smem_d + (size_t)warp * panel_lanes_warp_doubles(P, M);
  double *ring = base + grp * 32;
  double *theta = base + 128;
  double *fin = theta + ((P + 1) & ~1);
  unsigned long long *sacc = (unsigned long long *)(fin + NA + M + 4);  // [2 NA] (hi, lo) per raw sum
  load_logtab(logtab);
  __syncthreads();
  const int total = w.n_eval * w.units_per_eval;
  int cur = -1, units_cur = 0;
  uint32_t c2 = 0u, c3 = SMM_STREAM_SIM << 28;
  double rho = 0.0, beta = 0.0, phi = 0.0, sig_a = 0.0, sig_e = 0.0, mu0 = 0.0, x0scale = 0.0;
  int next = 0;
  if (lane == 0) next = (int)atomicAdd(w.unit_ctr, 1u);
  next = __shfl_sync(0xffffffffu, next, 0);
  for (;;) {
    const int u = next;
    const int e = u < total ? u / w.units_per_eval : -1;
    if (e != cur) {
      if (cur >= 0) {
        // leave evaluation `cur`: the warp's exact sums go to the evaluation's, its units are counted
        unsigned long long *acc = w.acc + (size_t)cur * 2 * NA;
        __syncwarp();
        for (int a = lane; a < 2 * NA; a += 32) atomicAdd(acc + a, sacc[a]);
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          const unsigned prev = atom_acq_rel_gpu(w.done + cur, (unsigned)units_cur);
          last = (prev + (unsigned)units_cur == (unsigned)w.units_per_eval);
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          const FinScratch fs{theta, fin, fin + NA, fin + NA + M, (int *)(fin + NA + M + 2)};
          for (int a = lane; a < NA; a += 32) {
            const long long hi = (long long)__ldcg(acc + 2 * a), lo = (long long)__ldcg(acc + 2 * a + 1);
            fs.tot[a] = __fma_rn((double)lo, pb.pan_lo_inv, __dmul_rn((double)hi, pb.pan_hi_inv));
          }
          __syncwarp();
          panel_moments(pb, fs, lane);
          const Grp gw{lane, 32, 0};
          group_distance(pb, gw, fs);
          if (w.iter > 0) {
            group_accept_store(pb, st, gw, fs, cur, global_chain(pb, cur), w.iter, false);
          } else {
            if (lane == 0) {
              w.value[cur] = fs.value[0];
              w.status[cur] = fs.flags[1];
            }
            for (int m = lane; m < M; m += 32) w.moments[(size_t)cur * M + m] = fs.mom[m];
          }
          __syncwarp();
        }
      }
      cur = e;
      units_cur = 0;
      if (e >= 0) {
        for (int q = lane; q < P; q += 32) theta[q] = __ldcg(w.params + (size_t)e * P + q);
        for (int a = lane; a < 2 * NA; a += 32) sacc[a] = 0ull;
        c2 = w.noseed ? w.uid0 + w.uid_stride * (uint32_t)e : 0u;
        c3 = (SMM_STREAM_SIM << 28) | (w.noseed ? ((w.rep0 + w.rep_stride * (uint32_t)e) & SMM_ITER_MASK) : 0u);
        __syncwarp();
        rho = theta[0];
        beta = theta[1 + k];
        phi = theta[1 + K + k];
        sig_a = theta[1 + 2 * K];
        sig_e = theta[2 + 2 * K];
        mu0 = theta[3 + 2 * K];
        x0scale = __dsqrt_rn(__fma_rn(-phi, phi, 1.0));
      }
    }
    if (e < 0) break;
    if (lane == 0) next = (int)atomicAdd(w.unit_ctr, 1u);  // prefetch: the latency hides behind the simulation
    ++units_cur;
    const int i_unit = (u - e * w.units_per_eval) * 32;
#pragma unroll 1
    for (int pass = 0; pass < 8; ++pass) {
      const int i = i_unit + pass * 4 + grp;
      if (i_unit + pass * 4 >= NI) break;  // warp uniform: nobody left in this unit
      const bool valid = i < NI;
      const uint32_t row = valid ? (uint32_t)i : 0u;
      // ---- normals: rounds of 8 blocks per individual into the ring ----
      int avail = 0, rd = 0, wr = 0;
      uint32_t jblk = 0u;
      auto produce = [&]() {
        __syncwarp();  // everybody has read the slots that are overwritten now
        double z0, z1;
        smm_normal_pair_tab(philox_sim(pb, jblk + (uint32_t)k, row, c2, c3), logtab, &z0, &z1);
        ring[(wr + 2 * k) & 31] = z0;
        ring[(wr + 2 * k + 1) & 31] = z1;
        __syncwarp();
        jblk += 8u;
        wr = (wr + 16) & 31;
        avail += 16;
      };
      produce();
      const double a_i = ring[rd & 31], e_init = ring[(rd + 1 + k) & 31];
      rd = (rd + 9) & 31;
      avail -= 9;
      const double alpha = __fma_rn(sig_a, a_i, mu0);
      const double y0 = __ddiv_rn(alpha, __dsub_rn(1.0, rho));
      double x = __ddiv_rn(e_init, x0scale);
      const double x0 = x;
      double sx = 0.0, sxx = 0.0, sxxl = 0.0, syx = 0.0, syxl = 0.0;
      double sy = 0.0, syy = 0.0, head = 0.0, tail = 0.0;
      double d = k == 1 ? y0 : 0.0;  // lane l >= 1: y_{t-l} (0 where t - l < 0)
      double yl = y0;
#pragma unroll 1
      for (int t = 1; t <= T; ++t) {
        if (avail < 9) produce();
        const double eta = ring[(rd + k) & 31], eps = ring[(rd + 8) & 31];
        rd = (rd + 9) & 31;
        avail -= 9;
        const double xp = x;
        x = __fma_rn(phi, xp, eta);
        double bx = __dmul_rn(beta, x);
        bx = __dadd_rn(bx, __shfl_xor_sync(0xffffffffu, bx, 1));
        bx = __dadd_rn(bx, __shfl_xor_sync(0xffffffffu, bx, 2));
        bx = __dadd_rn(bx, __shfl_xor_sync(0xffffffffu, bx, 4));
        const double y = __fma_rn(sig_e, eps, __dadd_rn(__fma_rn(rho, yl, alpha), bx));
        sy = __dadd_rn(sy, y);
        syy = __fma_rn(y, k == 0 ? y : d, syy);
        head = __dadd_rn(head, t < k ? y : 0.0);
        tail = __dadd_rn(tail, t > T - k ? y : 0.0);
        sx = __dadd_rn(sx, x);
        sxx = __fma_rn(x, x, sxx);
        sxxl = __fma_rn(x, xp, sxxl);
        syx = __fma_rn(y, x, syx);
        syxl = __fma_rn(y, xp, syxl);
        const double dn = __shfl_up_sync(0xffffffffu, d, 1, 8);
        d = k == 1 ? y : dn;
        yl = y;
      }
      if (valid) {
        // lane k: its regressor's six sums; lane l <= 6: Syy_l (and HG_l for l >= 1); lane 7: Sy
        auto pool = [&](int a, double v) {
          unsigned long long hi = 0ull, lo = 0ull;
          panel_split(pb, v, hi, lo);
          atomicAdd(sacc + 2 * a, hi);
          atomicAdd(sacc + 2 * a + 1, lo);
        };
        pool(14 + k, sx);
        pool(14 + K + k, syx);
        pool(14 + 2 * K + k, syxl);
        pool(14 + 3 * K + k, sxxl);
        pool(14 + 4 * K + k, sxx);
        pool(14 + 5 * K + k, __dsub_rn(__dadd_rn(x0, sx), x));
        if (k == 7) {
          pool(0, sy);
        } else {
          pool(1 + k, syy);
          if (k >= 1) pool(7 + k, __dsub_rn(__dadd_rn(__dadd_rn(sy, sy), y0), __dadd_rn(head, tail)));
        }
      }
    }
    __syncwarp();
    next = __shfl_sync(0xffffffffu, next, 0);
  }
}

// proposals of iteration `iter` for every local chain (one CTA per chain) -> st.pp; re-arms the panel work queue
__global__ void __launch_bounds__(kEvalThreads) bgp_propose_kernel(DevProblem pb, DevState st, int iter, int zero_len) {
  __shared__ EvalSmem sm;
  const int c = blockIdx.x, tid = threadIdx.x;
  const Grp g{tid, (int)blockDim.x, 0};
  load_logtab(sm.logtab);
  __syncthreads();
  group_proposal(pb, st, g, prop_scratch(sm), c, global_chain(pb, c), iter, true);
  for (int k = tid; k < pb.P; k += blockDim.x) st.pp[(size_t)c * pb.P + k] = sm.pp[k];
  unsigned long long *acc = (unsigned long long *)st.partials + (size_t)c * zero_len;
  for (int a = tid; a < zero_len; a += blockDim.x) acc[a] = 0ull;
  if (tid == 0) {
    st.arrive[c] = 0u;
    if (c == 0) *st.unit_ctr = 0u;
  }
}

static bool panel_is_lanes(int K, int variant) { return K == kLaneK && variant >= 6; }
size_t panel_smem_bytes(int K, int P, int variant) {
  if (panel_is_lanes(K, variant)) return sizeof(double) * panel_lanes_warp_doubles(P, 4 * K + 8) * (kPanelThreads / 32);
  return sizeof(double) * panel_warp_smem_doubles(K, P) * (kPanelThreads / 32);
}

// K = 8 (the C4 shape) has register-resident instantiations; `variant` picks the register budget (= CTAs per SM the
// compiler must make room for) and where the 5K per-regressor sums live:
//   2: 255 registers, sums in registers      3: 168 registers, sums in registers (spills)
//   4: 168 registers, sums in the shared staging tile (3 CTAs per SM)      5: 128 registers, shared sums (4 CTAs)
//   6 / 7 / 8: panel_lanes_kernel (eight lanes per individual) with room for 4 / 6 / 8 CTAs per SM
typedef void (*PanelKernel)(DevProblem, DevState, PanelWork);
static PanelKernel panel_kernel(int K, int variant) {
  if (K == 8) {
    switch (variant) {
      case 3: return panel_sim_kernel<8, 3>;
      case 4: return panel_sim_kernel<8, 3, true>;
      case 5: return panel_sim_kernel<8, 4, true>;
      case 6: return panel_lanes_kernel<4>;
      case 7: return panel_lanes_kernel<6>;
      case 8: return panel_lanes_kernel<8>;
      default: return panel_sim_kernel<8, 2>;
    }
  }
  return panel_sim_kernel<0, 1>;
}

cudaError_t configure_panel(int K, int P, int variant) {
  return cudaFuncSetAttribute(panel_kernel(K, variant), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)panel_smem_bytes(K, P, variant));
}
int panel_max_blocks_per_sm(int K, int P, int variant) {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, panel_kernel(K, variant), kPanelThreads, panel_smem_bytes(K, P, variant));
  return n;
}
void launch_propose(const DevProblem &pb, const DevState &st, int iter, int zero_len, cudaStream_t s) {
  bgp_propose_kernel<<<pb.L, kEvalThreads, 0, s>>>(pb, st, iter, zero_len);
}
// chain mode (iter >= 1): evaluations = the local chains' proposals in st.pp, sums in st.partials / st.arrive
void launch_panel_chains(const DevProblem &pb, const DevState &st, int iter, int grid, int variant, cudaStream_t s) {
  PanelWork w{};
  w.params = st.pp;
  w.acc = (unsigned long long *)st.partials;
  w.done = st.arrive;
  w.unit_ctr = st.unit_ctr;
  w.n_eval = pb.L;
  w.units_per_eval = (pb.panel_N + 31) / 32;
  w.noseed = pb.noseed;
  w.uid0 = (uint32_t)pb.rank;  // evaluation e = local chain e = global chain e * world + rank
  w.uid_stride = (uint32_t)pb.world;
  w.rep0 = (uint32_t)iter;
  w.rep_stride = 0u;
  w.iter = iter;
  panel_kernel(pb.panel_K, variant)<<<grid, kPanelThreads, panel_smem_bytes(pb.panel_K, pb.P, variant), s>>>(pb, st, w);
}
// batch mode: bare objective at params[B][P]
void launch_panel_batch(const DevProblem &pb, const DevState &st, const double *params, int B, int noseed, uint32_t uid0,
                        uint32_t rep0, unsigned long long *acc, unsigned *done, unsigned *unit_ctr, double *value,
                        double *moments, int *status, int grid, int variant, cudaStream_t s) {
  PanelWork w{};
  w.params = params;
  w.acc = acc;
  w.done = done;
  w.unit_ctr = unit_ctr;
  w.n_eval = B;
  w.units_per_eval = (pb.panel_N + 31) / 32;
  w.noseed = noseed;
  w.uid0 = uid0;
  w.uid_stride = 1u;
  w.rep0 = rep0;
  w.rep_stride = 1u;
  w.iter = 0;
  w.value = value;
  w.moments = moments;
  w.status = status;
  panel_kernel(pb.panel_K, variant)<<<grid, kPanelThreads, panel_smem_bytes(pb.panel_K, pb.P, variant), s>>>(pb, st, w);
}
