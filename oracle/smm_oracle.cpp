// smm_oracle.cpp -- CPU ORACLE for the BGP hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library; smm_jl_b200 (the product) never does and has no CPU fallback.
//
// What it is: a plain, sequential C++ restatement of the reference's algorithm, structured like the
// reference (one Eval record per evaluation, one BGPChain object per chain, O(iter) findlast and
// accept-rate recomputation, the simulated draw matrix materialised and then reduced), with the
// reference's unseedable randomness replaced by the injected counter-indexed streams of
// include/smm_stream.h.  Each function cites the reference lines it follows.
//
// PARITY UNPINNED against a real Julia run: the reference draws from RandomDevice() and Julia's
// global RNG (src/SMM.jl:59-60, AlgoBGP.jl:85,404,656, ObjExamples.jl:74), has no golden vectors
// for this path, and Julia is not installable here.  What IS pinned: the
// reference's portable behavioural tests (tests/test_oracle_behaviour.py ports
// test/test_BGPchain.jl:95-144, test/test_objfunc.jl:22-29, test/test_algoBGP.jl:30-38,57-121),
// the Philox known-answer vectors, and an independent numpy re-derivation (oracle/oracle_np.py).
//
// Build: oracle/Makefile (g++ -O2 -mfma -ffp-contract=off).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../include/smm_b200.h"
#include "../include/smm_stream.h"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

const double kNaN = std::numeric_limits<double>::quiet_NaN();
const double kInf = std::numeric_limits<double>::infinity();

// ---- Eval (src/mopt/Eval.jl:20-31; constructor defaults :82-106) ---------------------------------
struct Eval {
  double value = -1.0;
  int status = -1;
  double prob = 0.0;
  bool accepted = false;
  std::vector<double> params;      // P, in params_to_sample order
  std::vector<double> simMoments;  // M (NaN = never set, the reference's empty dict)
};

// ---- MProb (src/mopt/mprob.jl:29-53) -------------------------------------------------------------
struct MProb {
  int P = 0, M = 0;
  std::vector<double> lb, ub, init, data, w;
  int objective = 0;
  int S = 10000;
  uint64_t seed_sim = 1234;
  int noseed = 0;
  double slow_seconds = 0.0;
  int panel_T = 0, panel_N = 0, panel_K = 0;
};

// ---- the simulator streams -----------------------------------------------------------------------
// Z[k, s] for s in [0, S).  The transform is part of the simulator: ziggurat for the MvNormal objectives (the
// algorithm behind Julia's randn; three normals per Philox block: Z[k, 3j + t] = draw t of block (j, row k),
// smm_zig_triple), Box-Muller for the dynamic panel (two per block: Z[k, 2j + t], smm_normal_pair).
void fill_normals_row(const MProb &m, uint32_t k, int S, uint32_t uid, uint32_t rep, double *out) {
  if (m.objective != SMM_OBJ_PANEL) {
    for (int j = 0; 3 * j < S; ++j) {
      double z[3];
      smm_zig_triple(smm_sim_block(m.seed_sim, (uint32_t)j, k, m.noseed, uid, rep), z);
      for (int t = 0; t < 3 && 3 * j + t < S; ++t) out[3 * j + t] = z[t];
    }
    return;
  }
  for (int j = 0; 2 * j < S; ++j) {
    double z0, z1;
    smm_normal_pair(smm_sim_block(m.seed_sim, (uint32_t)j, k, m.noseed, uid, rep), &z0, &z1);
    out[2 * j] = z0;
    if (2 * j + 1 < S) out[2 * j + 1] = z1;
  }
}

// value = mean_k ((sim_k - data_k) / w_k)^2   (ObjExamples.jl:90-101; the weight divides)
double weighted_distance(const MProb &m, const std::vector<double> &sim) {
  double acc = 0.0;
  for (int k = 0; k < m.M; ++k) {
    double d = (sim[k] - m.data[k]) / m.w[k];
    acc += d * d;
  }
  return acc / (double)m.M;
}

// objfunc_norm (ObjExamples.jl:59-116): X = rand(MvNormal(mu, I), ns) is an nm x ns matrix filled
// column by column, X[k,s] = mu[k] + Z[k,s]; simM = mean(X, dims=2) accumulates along s.
void objfunc_norm(const MProb &m, Eval &ev, uint32_t uid, uint32_t rep) {
  const int D = m.P, S = m.S;
  std::vector<double> X((size_t)D * S), zrow(S);
  for (int k = 0; k < D; ++k) {
    fill_normals_row(m, (uint32_t)k, S, uid, rep, zrow.data());
    for (int s = 0; s < S; ++s) X[(size_t)k + (size_t)D * s] = ev.params[k] + zrow[s];
  }
  std::vector<double> simM(D, 0.0);
  for (int s = 0; s < S; ++s)
    for (int k = 0; k < D; ++k) simM[k] += X[(size_t)k + (size_t)D * s];
  for (int k = 0; k < D; ++k) simM[k] /= (double)S;
  ev.simMoments = simM;
  ev.value = weighted_distance(m, simM);
  ev.status = 1;
}

// norm_mv (generalises objfunc_norm2, ObjExamples.jl:191-249): the same draw matrix,
// moments 1..D = row means, D+1..2D = row sample variances (Julia var: two-pass, n-1).
void objfunc_norm_mv(const MProb &m, Eval &ev, uint32_t uid, uint32_t rep) {
  const int D = m.P, S = m.S;
  std::vector<double> X((size_t)D * S), zrow(S);
  for (int k = 0; k < D; ++k) {
    fill_normals_row(m, (uint32_t)k, S, uid, rep, zrow.data());
    for (int s = 0; s < S; ++s) X[(size_t)k + (size_t)D * s] = ev.params[k] + zrow[s];
  }
  std::vector<double> sim(2 * D, 0.0);
  for (int s = 0; s < S; ++s)
    for (int k = 0; k < D; ++k) sim[k] += X[(size_t)k + (size_t)D * s];
  for (int k = 0; k < D; ++k) sim[k] /= (double)S;
  for (int s = 0; s < S; ++s)
    for (int k = 0; k < D; ++k) {
      double d = X[(size_t)k + (size_t)D * s] - sim[k];
      sim[D + k] += d * d;
    }
  for (int k = 0; k < D; ++k) sim[D + k] /= (double)(S - 1);
  ev.simMoments = sim;
  ev.value = weighted_distance(m, sim);
  ev.status = 1;
}

// Dynamic panel (SURVEY.md 8d "C4"; no upstream code -- the spec is ours and is restated in
// DESIGN.md).  theta = (rho, beta[K], phi[K], sigma_alpha, sigma_eps, mu0).  Shocks of individual i
// live in stream row i: normal index 0 = a_i, 1..K = x-initial eta_k, then for t = 1..T the K
// regressor shocks eta_{k,t} followed by eps_t  =>  1 + K + T*(K+1) normals.
void objfunc_panel(const MProb &m, Eval &ev, uint32_t uid, uint32_t rep) {
  const int K = m.panel_K, T = m.panel_T, NI = m.panel_N;
  const double *th = ev.params.data();
  const double rho = th[0];
  const double *beta = th + 1, *phi = th + 1 + K;
  const double sig_a = th[1 + 2 * K], sig_e = th[2 + 2 * K], mu0 = th[3 + 2 * K];
  const int nz = 1 + K + T * (K + 1);
  // materialise the panel like a user model would: y[i,t], x[k,i,t], t = 0..T
  std::vector<double> y((size_t)NI * (T + 1)), x((size_t)K * NI * (T + 1)), z(nz + 1);
  for (int i = 0; i < NI; ++i) {
    fill_normals_row(m, (uint32_t)i, nz, uid, rep, z.data());
    // the recurrences are DEFINED with fused multiply-adds (the spec is ours; Julia's `fma`), so that the
    // device's DFMA and this loop round identically
    const double alpha = std::fma(sig_a, z[0], mu0);
    double yc = alpha / (1.0 - rho);
    std::vector<double> xc(K);
    for (int k = 0; k < K; ++k) {
      xc[k] = z[1 + k] / std::sqrt(std::fma(-phi[k], phi[k], 1.0));
      x[((size_t)k * NI + i) * (T + 1)] = xc[k];
    }
    y[(size_t)i * (T + 1)] = yc;
    for (int t = 1; t <= T; ++t) {
      const double *zt = z.data() + 1 + K + (t - 1) * (K + 1);
      double xb = 0.0;
      for (int k = 0; k < K; ++k) {
        xc[k] = std::fma(phi[k], xc[k], zt[k]);
        x[((size_t)k * NI + i) * (T + 1) + t] = xc[k];
        xb = std::fma(beta[k], xc[k], xb);
      }
      yc = std::fma(sig_e, zt[K], std::fma(rho, yc, alpha) + xb);
      y[(size_t)i * (T + 1) + t] = yc;
    }
  }
  // moments pooled over (i, t = 1..T); lagged terms use t-l >= 0 (t = 0 is the initial condition)
  auto Y = [&](int i, int t) { return y[(size_t)i * (T + 1) + t]; };
  auto Xk = [&](int k, int i, int t) { return x[((size_t)k * NI + i) * (T + 1) + t]; };
  const double n = (double)NI * (double)T;
  std::vector<double> sim(4 * K + 8, 0.0);
  double my = 0.0;
  for (int i = 0; i < NI; ++i)
    for (int t = 1; t <= T; ++t) my += Y(i, t);
  my /= n;
  std::vector<double> mx(K, 0.0);
  for (int k = 0; k < K; ++k) {
    for (int i = 0; i < NI; ++i)
      for (int t = 1; t <= T; ++t) mx[k] += Xk(k, i, t);
    mx[k] /= n;
  }
  int o = 0;
  sim[o++] = my;
  {  // var y and autocovariances lag 1..6: E[(y_t - my)(y_{t-l} - my)] over t = 1..T, all i
    for (int l = 0; l <= 6; ++l) {
      double a = 0.0;
      for (int i = 0; i < NI; ++i)
        for (int t = 1; t <= T; ++t) {
          int tl = t - l;
          if (tl < 0) continue;
          a += (Y(i, t) - my) * (Y(i, tl) - my);
        }
      sim[o++] = a / n;
    }
  }
  for (int k = 0; k < K; ++k) {  // cov(y_t, x_kt)
    double a = 0.0;
    for (int i = 0; i < NI; ++i)
      for (int t = 1; t <= T; ++t) a += (Y(i, t) - my) * (Xk(k, i, t) - mx[k]);
    sim[o++] = a / n;
  }
  for (int k = 0; k < K; ++k) {  // cov(y_t, x_k,t-1)
    double a = 0.0;
    for (int i = 0; i < NI; ++i)
      for (int t = 1; t <= T; ++t) a += (Y(i, t) - my) * (Xk(k, i, t - 1) - mx[k]);
    sim[o++] = a / n;
  }
  for (int k = 0; k < K; ++k) {  // autocov x_k lag 1
    double a = 0.0;
    for (int i = 0; i < NI; ++i)
      for (int t = 1; t <= T; ++t) a += (Xk(k, i, t) - mx[k]) * (Xk(k, i, t - 1) - mx[k]);
    sim[o++] = a / n;
  }
  for (int k = 0; k < K; ++k) {  // var x_k
    double a = 0.0;
    for (int i = 0; i < NI; ++i)
      for (int t = 1; t <= T; ++t) a += (Xk(k, i, t) - mx[k]) * (Xk(k, i, t) - mx[k]);
    sim[o++] = a / n;
  }
  ev.simMoments = sim;
  ev.value = weighted_distance(m, sim);
  ev.status = 1;
}

// evaluateObjective (mprob.jl:175-188): build the record, call the objective, exceptions -> -2
Eval evaluateObjective(const MProb &m, const std::vector<double> &p, uint32_t uid, uint32_t rep) {
  Eval ev;
  ev.params = p;
  ev.simMoments.assign(m.M, kNaN);
  switch (m.objective) {
    case SMM_OBJ_NORM:
      objfunc_norm(m, ev, uid, rep);
      break;
    case SMM_OBJ_NORM_SLOW:  // ObjExamples.jl:130 sleep, then the same body
      std::this_thread::sleep_for(std::chrono::duration<double>(m.slow_seconds));
      objfunc_norm(m, ev, uid, rep);
      break;
    case SMM_OBJ_NORM_MV:
      objfunc_norm_mv(m, ev, uid, rep);
      break;
    case SMM_OBJ_PANEL:
      objfunc_panel(m, ev, uid, rep);
      break;
    case SMM_OBJ_FAILS:  // Testobj_fails throws (ObjExamples.jl:27-32) -> caught, status = -2
    default:
      ev.status = -2;
      break;
  }
  return ev;
}

// ---- BGPChain (AlgoBGP.jl:42-110) --------------------------------------------------------------
struct BGPChain {
  std::vector<Eval> evals;
  std::vector<int> best_id;
  std::vector<double> best_val, curr_val, probs_acc;
  std::vector<char> accepted;
  std::vector<int> exchanged;
  int id = 0;  // 1-based
  int iter = 0;
  double accept_rate = 0.0, acc_tuner = 2.0, sigma = 0.5;
  int sigma_update_steps = 10;
  double sigma_adjust_by = 0.01;
  int smpl_iters = 1000;
  double min_improve = 0.0;
  std::vector<std::pair<int, int>> batches;  // [lo, hi) 0-based
  long attempts = 0;
};

// BGPChain constructor (AlgoBGP.jl:78-109).  probs_acc = rand(n) becomes the Uacc stream.
BGPChain make_chain(int id, int n, const MProb &m, double sig, int upd, double upd_by, int smpl_iters,
                    double min_improve, double acc_tuner, int batch_size, uint64_t seed_algo) {
  BGPChain c;
  c.evals.resize(n);
  c.best_val.assign(n, kInf);
  c.best_id.assign(n, -1);
  c.curr_val.assign(n, kInf);
  c.probs_acc.resize(n);
  for (int it = 1; it <= n; ++it) c.probs_acc[it - 1] = smm_acc_uniform(seed_algo, (uint32_t)(id - 1), (uint32_t)it);
  c.accepted.assign(n, 0);
  c.exchanged.assign(n, 0);
  c.id = id;
  c.iter = 0;
  c.acc_tuner = acc_tuner;
  c.sigma = sig;
  // batches (AlgoBGP.jl:95-103); only np % batch_size == 0 is accepted (checked by the caller)
  for (int lo = 0; lo < m.P; lo += batch_size) c.batches.push_back({lo, lo + batch_size});
  c.sigma_update_steps = upd;
  c.sigma_adjust_by = upd_by;
  c.smpl_iters = smpl_iters;
  c.min_improve = min_improve;
  return c;
}

// lastAccepted (AlgoBGP.jl:209-215): findlast(c.accepted[1:c.iter]); returns 1-based index
int lastAccepted(const BGPChain &c) {
  if (c.iter == 1) return 1;
  for (int it = c.iter; it >= 1; --it)
    if (c.accepted[it - 1]) return it;
  return 0;  // `nothing` upstream; cannot happen because iteration 1 is always accepted
}
const Eval &getLastAccepted(const BGPChain &c) { return c.evals[lastAccepted(c) - 1]; }

// set_eval! (AlgoBGP.jl:220-245)
void set_eval(BGPChain &c, const Eval &ev) {
  const int i = c.iter - 1;
  c.evals[i] = ev;  // deepcopy
  c.accepted[i] = ev.accepted;
  if (c.iter == 1) {
    c.best_val[i] = ev.value;
    c.curr_val[i] = ev.value;
    c.best_id[i] = c.iter;
  } else {
    c.curr_val[i] = ev.accepted ? ev.value : c.curr_val[i - 1];
    if (ev.value < c.best_val[i - 1]) {
      c.best_val[i] = ev.value;
      c.best_id[i] = c.iter;
    } else {
      c.best_val[i] = c.best_val[i - 1];
      c.best_id[i] = c.best_id[i - 1];
    }
  }
}

// set_acceptRate! (AlgoBGP.jl:253-257): mean(accepted[1:iter][exchanged[1:iter] .== 0])
void set_acceptRate(BGPChain &c) {
  long n = 0, a = 0;
  for (int it = 0; it < c.iter; ++it)
    if (c.exchanged[it] == 0) {
      ++n;
      a += c.accepted[it] ? 1 : 0;
    }
  c.accept_rate = n ? (double)a / (double)n : kNaN;  // mean of an empty vector is NaN in Julia
}

struct FatalError {
  int code;
  std::string msg;
};

// mysample (AlgoBGP.jl:400-410) for parameter indices [lo, hi): redraw until every coordinate is in
// [0,1]; attempt a uses Zprop[chain, iter, a, k].  Returns false when smpl_iters are exhausted.
bool mysample(BGPChain &c, uint64_t seed_algo, const std::vector<double> &mu01, int lo, int hi,
              std::vector<double> &out) {
  for (int a = 0; a < c.smpl_iters; ++a) {
    ++c.attempts;
    bool ok = true;
    for (int k = lo; k < hi; ++k) {
      double z0, z1;
      smm_normal_pair(smm_prop_block(seed_algo, (uint32_t)(c.id - 1), (uint32_t)c.iter, (uint32_t)a, (uint32_t)(k >> 1)),
                      &z0, &z1);
      const double z = (k & 1) ? z1 : z0;
      const double x = mu01[k] + c.sigma * z;  // MvNormal(mu01, sigma): isotropic, sigma is the s.d.
      out[k] = x;
      if (!(x >= 0.0) || !(x <= 1.0)) ok = false;
    }
    if (ok) return true;
  }
  return false;
}

// proposal (AlgoBGP.jl:424-471) with mapto_01 / mapto_ab (mprob.jl:246-272)
std::vector<double> proposal(BGPChain &c, const MProb &m, uint64_t seed_algo) {
  if (c.iter == 1) return m.init;
  const Eval &ev_old = getLastAccepted(c);
  std::vector<double> mu01(m.P), pp(m.P, 0.0);
  for (int k = 0; k < m.P; ++k) mu01[k] = (ev_old.params[k] - m.lb[k]) / (m.ub[k] - m.lb[k]);
  if (c.batches.size() == 1) {
    if (!mysample(c, seed_algo, mu01, 0, m.P, pp))
      throw FatalError{SMM_E_SAMPLER_EXHAUSTED, "no draw in support after smpl_iters trials"};
  } else {
    for (auto &b : c.batches) {
      std::vector<double> tmp(m.P, 0.0);
      if (mysample(c, seed_algo, mu01, b.first, b.second, tmp)) {
        for (int k = b.first; k < b.second; ++k) pp[k] = tmp[k];
      }  // else: exception logged and swallowed upstream (:447-451), pp[i] stays 0
    }
  }
  std::vector<double> newp(m.P);
  for (int k = 0; k < m.P; ++k) newp[k] = pp[k] * (m.ub[k] - m.lb[k]) + m.lb[k];
  return newp;
}

// Julia: minimum([1.0, x]) propagates NaN
double julia_min1(double x) { return std::isnan(x) ? x : (x < 1.0 ? x : 1.0); }

// doAcceptReject! (AlgoBGP.jl:324-392)
void doAcceptReject(BGPChain &c, Eval &ev) {
  const int i = c.iter - 1;
  if (c.iter == 1) {
    ev.prob = 1.0;
    ev.accepted = true;
    ev.status = 1;
    c.accepted[i] = ev.accepted;
    set_acceptRate(c);
  } else {
    const Eval &old = getLastAccepted(c);
    if (ev.status < 0) {
      ev.prob = 0.0;
      ev.accepted = false;
    } else {
      if (!(ev.value >= 0.0))
        throw FatalError{SMM_E_NEGATIVE_OBJECTIVE, "AlgoBGP assumes that your objective function returns a non-negative number."};
      ev.prob = julia_min1(std::exp(c.acc_tuner * (old.value - ev.value)));
      if (!std::isfinite(ev.prob)) {
        ev.prob = 0.0;
        ev.accepted = false;
        ev.status = -1;
      } else if (!std::isfinite(old.value)) {
        ev.prob = 1.0;
        ev.accepted = true;
      } else {
        ev.status = 1;
        ev.accepted = ev.prob > c.probs_acc[i];
      }
    }
    c.accepted[i] = ev.accepted;
    set_acceptRate(c);
    if (c.iter % c.sigma_update_steps == 0) {
      if (c.accept_rate > 0.234)
        c.sigma = c.sigma * (1.0 + c.sigma_adjust_by);
      else
        c.sigma = c.sigma * (1.0 - c.sigma_adjust_by);
    }
  }
}

// ---- MAlgoBGP (AlgoBGP.jl:497-539) -----------------------------------------------------------
struct MAlgoBGP {
  MProb m;
  int N = 0, i = 0, maxiter = 0;
  uint64_t seed_algo = 0;
  std::vector<BGPChain> chains;
  long swaps = 0;
};

// Pairs[iter]: `sample(props, N < 3 ? N-1 : N, replace=false)` (AlgoBGP.jl:653-656) restated as a
// self-avoiding sample over pair ranks: slot t keeps its first candidate not already chosen.
std::vector<std::pair<int, int>> sample_pairs(uint64_t seed_algo, int iter, int N) {
  const uint32_t n_all = (uint32_t)((uint64_t)N * (N - 1) / 2);
  const int n_s = N < 3 ? N - 1 : N;
  std::vector<uint32_t> chosen;
  std::vector<std::pair<int, int>> out;
  for (int t = 0; t < n_s; ++t) {
    uint32_t a = 0, q;
    do {
      q = smm_pair_candidate(seed_algo, (uint32_t)iter, (uint32_t)t, a++, n_all);
    } while (std::find(chosen.begin(), chosen.end(), q) != chosen.end());
    chosen.push_back(q);
    uint32_t i, j;
    smm_pair_unrank(q, &i, &j);
    out.push_back({(int)i, (int)j});
  }
  return out;
}

// swap_ev_ij! (AlgoBGP.jl:734-749)
void swap_ev_ij(MAlgoBGP &algo, int i, int j) {
  BGPChain &ci = algo.chains[i], &cj = algo.chains[j];
  const Eval ei = getLastAccepted(ci);  // copies: upstream holds references to the old objects
  const Eval ej = getLastAccepted(cj);
  set_eval(ci, ej);
  set_eval(cj, ei);
  ci.exchanged[ci.iter - 1] = j + 1;
  cj.exchanged[cj.iter - 1] = i + 1;
  ++algo.swaps;
}

// exchangeMoves! (AlgoBGP.jl:647-691), dist_fun = `-` (:537)
void exchangeMoves(MAlgoBGP &algo) {
  for (auto &p : sample_pairs(algo.seed_algo, algo.i, algo.N)) {
    const int i = p.first, j = p.second;
    const Eval &evi = getLastAccepted(algo.chains[i]);
    const Eval &evj = getLastAccepted(algo.chains[j]);
    if (evi.value - evj.value > algo.chains[i].min_improve) swap_ev_ij(algo, i, j);
  }
}

// computeNextIteration! (AlgoBGP.jl:589-640), written in the `parallel` form (:596-605): proposals on
// the master, objective evaluations farmed out (threads here, pmap upstream), accept/reject on the
// master.  The serial form (:614) computes the same thing chain by chain.
void computeNextIteration(MAlgoBGP &algo, int n_threads) {
  const int N = algo.N;
  for (auto &c : algo.chains) c.iter += 1;
  std::vector<std::vector<double>> pps(N);
  for (int c = 0; c < N; ++c) pps[c] = proposal(algo.chains[c], algo.m, algo.seed_algo);
  std::vector<Eval> evs(N);
  if (n_threads <= 1) {
    for (int c = 0; c < N; ++c) evs[c] = evaluateObjective(algo.m, pps[c], (uint32_t)c, (uint32_t)algo.i);
  } else {
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
      pool.emplace_back([&] {
        for (;;) {
          int c = next.fetch_add(1);
          if (c >= N) break;
          evs[c] = evaluateObjective(algo.m, pps[c], (uint32_t)c, (uint32_t)algo.i);
        }
      });
    for (auto &th : pool) th.join();
  }
  for (int c = 0; c < N; ++c) {  // next_acceptreject (AlgoBGP.jl:305-316)
    doAcceptReject(algo.chains[c], evs[c]);
    set_eval(algo.chains[c], evs[c]);
  }
  if (algo.i >= 2 && N > 1) exchangeMoves(algo);
}

int check_config(const smm_bgp_config *cfg) {
  if (!cfg) return fail(SMM_E_ARG, "null config");
  if (cfg->abi_version != SMM_ABI_VERSION) return fail(SMM_E_ARG, "abi_version mismatch");
  if (cfg->n_params < 1 || cfg->n_params > SMM_MAX_PARAMS) return fail(SMM_E_ARG, "n_params out of range");
  if (cfg->n_moments < 1 || cfg->n_moments > SMM_MAX_MOMENTS) return fail(SMM_E_ARG, "n_moments out of range");
  if (cfg->n_chains < 1 || cfg->max_iter < 1) return fail(SMM_E_ARG, "n_chains / max_iter must be positive");
  if (cfg->batch_size < 1 || cfg->n_params % cfg->batch_size != 0)
    return fail(SMM_E_UNSUPPORTED_SHAPE, "batch_size must divide n_params (upstream's batches are ill-formed otherwise, AlgoBGP.jl:95-103)");
  const int P = cfg->n_params, M = cfg->n_moments;
  switch (cfg->objective_id) {
    case SMM_OBJ_NORM:
    case SMM_OBJ_NORM_SLOW:
      if (P != M) return fail(SMM_E_UNSUPPORTED_SHAPE, "objfunc_norm needs n_params == n_moments (ObjExamples.jl:77-78)");
      break;
    case SMM_OBJ_NORM_MV:
      if (M != 2 * P) return fail(SMM_E_UNSUPPORTED_SHAPE, "norm_mv needs n_moments == 2*n_params");
      break;
    case SMM_OBJ_PANEL:
      if (cfg->panel_K < 1 || P != 2 * cfg->panel_K + 4 || M != 4 * cfg->panel_K + 8 || cfg->panel_T < 7 || cfg->panel_N < 1)
        return fail(SMM_E_UNSUPPORTED_SHAPE, "panel needs P == 2K+4, M == 4K+8, T >= 7");
      break;
    case SMM_OBJ_FAILS:
      break;
    default:
      return fail(SMM_E_ARG, "unknown objective_id");
  }
  if (cfg->n_sim < 2) return fail(SMM_E_ARG, "n_sim must be >= 2");
  for (int k = 0; k < P; ++k)
    if (!(cfg->ub[k] > cfg->lb[k])) return fail(SMM_E_ARG, "need ub > lb (mprob.jl:82)");
  return 0;
}

MProb make_mprob(const smm_bgp_config *cfg) {
  MProb m;
  m.P = cfg->n_params;
  m.M = cfg->n_moments;
  m.lb.assign(cfg->lb, cfg->lb + m.P);
  m.ub.assign(cfg->ub, cfg->ub + m.P);
  m.init.assign(cfg->init, cfg->init + m.P);
  m.data.assign(cfg->data_mom, cfg->data_mom + m.M);
  m.w.assign(cfg->data_w, cfg->data_w + m.M);
  m.objective = cfg->objective_id;
  m.S = cfg->n_sim;
  m.seed_sim = cfg->seed_sim;
  m.noseed = cfg->noseed;
  m.slow_seconds = cfg->slow_seconds;
  m.panel_T = cfg->panel_T;
  m.panel_N = cfg->panel_N;
  m.panel_K = cfg->panel_K;
  return m;
}

}  // namespace

extern "C" {

const char *smm_oracle_last_error(void) { return g_err.c_str(); }

// run!(algo) for n_iters iterations from a fresh MAlgoBGP (AlgoAbstract.jl:27-76).  `out` receives
// [n_iters][N] rows in the layout of smm_trace_view; sigma / accept_rate / counters are per chain.
int smm_oracle_run(const smm_bgp_config *cfg, int n_iters, int n_threads, const smm_trace_view *out,
                   double *sigma, double *accept_rate, int64_t *swaps, int64_t *attempts) {
  if (int rc = check_config(cfg)) return rc;
  if (n_iters < 1 || n_iters > cfg->max_iter) return fail(SMM_E_ARG, "n_iters must be in [1, max_iter]");
  MAlgoBGP algo;
  algo.m = make_mprob(cfg);
  algo.N = cfg->n_chains;
  algo.maxiter = cfg->max_iter;
  algo.seed_algo = cfg->seed_algo;
  for (int c = 0; c < algo.N; ++c)
    algo.chains.push_back(make_chain(c + 1, cfg->max_iter, algo.m, cfg->sigma0[c], cfg->sigma_update_steps,
                                     cfg->sigma_adjust_by, cfg->smpl_iters, cfg->min_improve[c], cfg->acc_tuner[c],
                                     cfg->batch_size, cfg->seed_algo));
  try {
    for (int i = 1; i <= n_iters; ++i) {
      algo.i = i;
      computeNextIteration(algo, n_threads);
    }
  } catch (const FatalError &e) {
    return fail(e.code, e.msg);
  }
  const int N = algo.N, P = algo.m.P, M = algo.m.M;
  if (out) {
    for (int it = 0; it < n_iters; ++it)
      for (int c = 0; c < N; ++c) {
        const BGPChain &ch = algo.chains[c];
        const Eval &ev = ch.evals[it];
        const size_t r = (size_t)it * N + c;
        if (out->value) out->value[r] = ev.value;
        if (out->prob) out->prob[r] = ev.prob;
        if (out->curr_val) out->curr_val[r] = ch.curr_val[it];
        if (out->best_val) out->best_val[r] = ch.best_val[it];
        if (out->params)
          for (int k = 0; k < P; ++k) out->params[r * P + k] = ev.params[k];
        if (out->sim_moments)
          for (int k = 0; k < M; ++k) out->sim_moments[r * M + k] = ev.simMoments[k];
        if (out->accepted) out->accepted[r] = ch.accepted[it] ? 1 : 0;
        if (out->status) out->status[r] = ev.status;
        if (out->exchanged) out->exchanged[r] = ch.exchanged[it];
        if (out->best_id) out->best_id[r] = ch.best_id[it];
      }
  }
  long att = 0;
  for (int c = 0; c < N; ++c) {
    if (sigma) sigma[c] = algo.chains[c].sigma;
    if (accept_rate) accept_rate[c] = algo.chains[c].accept_rate;
    att += algo.chains[c].attempts;
  }
  if (swaps) *swaps = algo.swaps;
  if (attempts) *attempts = att;
  return 0;
}

// evaluateObjective(m, p; noseed) for B parameter vectors (mprob.jl:175-188)
int smm_oracle_eval_batch(const smm_bgp_config *cfg, const double *params, int B, int noseed, uint32_t rep0,
                          int n_threads, double *value, double *moments, int32_t *status) {
  if (int rc = check_config(cfg)) return rc;
  MProb m = make_mprob(cfg);
  m.noseed = noseed;
  const int P = m.P, M = m.M;
  auto one = [&](int b) {
    std::vector<double> p(params + (size_t)b * P, params + (size_t)(b + 1) * P);
    Eval ev = evaluateObjective(m, p, (uint32_t)b, rep0 + (uint32_t)b);
    if (value) value[b] = ev.value;
    if (status) status[b] = ev.status;
    if (moments)
      for (int k = 0; k < M; ++k) moments[(size_t)b * M + k] = ev.simMoments[k];
  };
  if (n_threads <= 1) {
    for (int b = 0; b < B; ++b) one(b);
  } else {
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
      pool.emplace_back([&] {
        for (;;) {
          int b = next.fetch_add(1);
          if (b >= B) break;
          one(b);
        }
      });
    for (auto &th : pool) th.join();
  }
  return 0;
}

// the Pairs stream in sampling order (0-based i<j); returns the number of pairs
int smm_oracle_pairs(uint64_t seed_algo, int iter, int N, int32_t *ij) {
  auto ps = sample_pairs(seed_algo, iter, N);
  for (size_t t = 0; t < ps.size(); ++t) {
    ij[2 * t] = ps[t].first;
    ij[2 * t + 1] = ps[t].second;
  }
  return (int)ps.size();
}

// raw stream access for the stream tests
void smm_oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
  smm_u32x4 r = smm_philox4x32_10(c0, c1, c2, c3, k0, k1);
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
  out[3] = r.w;
}
void smm_oracle_normals(uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int n_pairs, double *out) {
  for (int j = 0; j < n_pairs; ++j)
    smm_normal_pair(smm_philox4x32_10((uint32_t)j, k, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32)), &out[2 * j],
                    &out[2 * j + 1]);
}
void smm_oracle_normal_from_words(uint32_t x, uint32_t y, uint32_t z, uint32_t w, double *out) {
  smm_u32x4 r = {x, y, z, w};
  smm_normal_pair(r, &out[0], &out[1]);
}
// out[3 * n_blocks]: the three ziggurat normals of Philox blocks (j, k, c2, c3), j < n_blocks
void smm_oracle_zig_normals(uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int n_blocks, double *out) {
  for (int j = 0; j < n_blocks; ++j)
    smm_zig_triple(smm_philox4x32_10((uint32_t)j, k, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32)), &out[3 * j]);
}
// one ziggurat normal from its 32-bit uniform and 10-bit select field; *slow = 1 when the fast path rejected the candidate
double smm_oracle_zig_from_words(uint32_t u, uint32_t sel, int *slow) {
  int ok;
  const double z = smm_zig_fast(u, sel, smm_zigtab(), &ok);
  if (slow) *slow = !ok;
  return ok ? z : smm_zig_slow(u, sel, smm_zigtab(), smm_logtab());
}
uint32_t smm_oracle_zig_select(uint32_t w, int t) { return smm_zig_select(w, t); }
double smm_oracle_exp_neg(double t) { return smm_exp_neg(t); }

// "Reference-speed proxy" (BASELINE.md section 3, B-proxy): the objective's data flow exactly as the reference has it
// (ObjExamples.jl:76-101: materialise the D x S Float64 draw matrix, then reduce it to means and variances) with the
// kind of generator Julia's randn is -- xoshiro256++ state in registers + a ziggurat (one 64-bit word per normal), libm exp/log in the
// rare branch.  NOT stream compatible with anything: it only answers "how fast would the CPU path be if its normals
// were as cheap as Julia's", so that the reported CPU baseline is not handicapped by the counter-based streams.
// Returns evaluations per second summed over n_threads threads (each thread evaluates n_evals_per_thread times).
double smm_oracle_proxy_rate(int D, int S, int n_evals_per_thread, int n_threads) {
  if (D < 1 || S < 2 || n_evals_per_thread < 1 || n_threads < 1) return 0.0;
  auto worker = [&](int tid, double *sink) {
    // xoshiro256++ with its state in locals (registers); the ziggurat's rare branch is kept out of line
    uint64_t s0 = 0x9E3779B97F4A7C15ull * (uint64_t)(tid + 1), s1 = 0xBF58476D1CE4E5B9ull, s2 = 0x94D049BB133111EBull,
             s3 = 0x2545F4914F6CDD1Dull + (uint64_t)tid;
#define SMM_XOSHIRO_NEXT(r)                                         \
  do {                                                              \
    const uint64_t sum_ = s0 + s3, t_ = s1 << 17;                   \
    (r) = ((sum_ << 23) | (sum_ >> 41)) + s0;                       \
    s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t_;               \
    s3 = (s3 << 45) | (s3 >> 19);                                   \
  } while (0)
    const smm_zigent *tab = smm_zigtab();
    std::vector<double> X((size_t)D * S), sim(2 * D), p(D);
    double acc = 0.0;
    const size_t n_draws = (size_t)D * S;
    for (int e = 0; e < n_evals_per_thread; ++e) {
      for (int k = 0; k < D; ++k) p[k] = 0.01 * (k + e % 7);
      // rand(MvNormal(mu, I), S): the D x S matrix filled in memory order, X[k, s] = mu[k] + z
      int k = 0;
      for (size_t q = 0; q < n_draws; ++q) {
        double z;
        for (;;) {
          uint64_t r;
          SMM_XOSHIRO_NEXT(r);
          const uint32_t i = (uint32_t)(r >> 32) & (SMM_ZIG_LAYERS - 1u);
          const double x = (double)(uint32_t)r * smm_bits_to_double(tab[i]);
          z = (r >> 63) ? -x : x;
          if (__builtin_expect((uint32_t)r < ((uint32_t)tab[i] << 20), 1)) break;
          if (i == 0) {
            if (x < SMM_ZIG_R) break;
            for (;;) {
              uint64_t a, b;
              SMM_XOSHIRO_NEXT(a);
              SMM_XOSHIRO_NEXT(b);
              const double xt = -std::log(((double)(a >> 11) + 0.5) * 0x1.0p-53) / SMM_ZIG_R;
              const double yt = -std::log(((double)(b >> 11) + 0.5) * 0x1.0p-53);
              if (yt + yt > xt * xt) {
                z = (r >> 63) ? -(SMM_ZIG_R + xt) : SMM_ZIG_R + xt;
                break;
              }
            }
            break;
          }
          uint64_t c;
          SMM_XOSHIRO_NEXT(c);
          const double f_lo = SMM_ZIGF_HOST[i], f_hi = SMM_ZIGF_HOST[i + 1];
          if (f_lo + ((double)(c >> 11) * 0x1.0p-53) * (f_hi - f_lo) < std::exp(-0.5 * x * x)) break;
        }
        X[q] = p[k] + z;
        if (++k == D) k = 0;
      }
#undef SMM_XOSHIRO_NEXT
      std::fill(sim.begin(), sim.end(), 0.0);
      for (int s = 0; s < S; ++s)
        for (int kk = 0; kk < D; ++kk) sim[kk] += X[(size_t)kk + (size_t)D * s];
      for (int kk = 0; kk < D; ++kk) sim[kk] /= (double)S;
      for (int s = 0; s < S; ++s)
        for (int kk = 0; kk < D; ++kk) {
          const double d = X[(size_t)kk + (size_t)D * s] - sim[kk];
          sim[D + kk] += d * d;
        }
      for (int kk = 0; kk < D; ++kk) acc += sim[kk] + sim[D + kk] / (double)(S - 1);
    }
    *sink = acc;
  };
  std::vector<double> sinks(n_threads, 0.0);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker, t, &sinks[t]);
  for (auto &th : pool) th.join();
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double chk = 0.0;
  for (double v : sinks) chk += v;
  if (!(chk == chk)) return 0.0;  // keeps the work observable
  return (double)n_threads * n_evals_per_thread / dt;
}
double smm_oracle_neglog01(double u) { return smm_neglog01(u, smm_logtab()); }
double smm_oracle_acc_uniform(uint64_t seed, uint32_t chain, uint32_t iter) { return smm_acc_uniform(seed, chain, iter); }
void smm_oracle_pair_unrank(uint32_t q, uint32_t *i, uint32_t *j) { smm_pair_unrank(q, i, j); }

// One accept/reject decision on a hand-built chain: the hook the reference's own unit tests use
// (test/test_BGPchain.jl:95-144).  iter >= 2 compares against `old_value`.
int smm_oracle_accept_reject(int iter, double old_value, double new_value, int new_status, double acc_tuner,
                             double u, double *prob, int *accepted, int *status) {
  BGPChain c;
  const int n = iter < 2 ? 2 : iter;
  c.evals.resize(n);
  c.best_val.assign(n, kInf);
  c.best_id.assign(n, -1);
  c.curr_val.assign(n, kInf);
  c.probs_acc.assign(n, u);
  c.accepted.assign(n, 0);
  c.exchanged.assign(n, 0);
  c.acc_tuner = acc_tuner;
  c.sigma_update_steps = 1 << 30;
  c.iter = 1;
  Eval e0;
  e0.value = old_value;
  e0.accepted = true;
  e0.status = 1;
  set_eval(c, e0);
  c.iter = iter;
  Eval ev;
  ev.value = new_value;
  ev.status = new_status;
  try {
    doAcceptReject(c, ev);
  } catch (const FatalError &e) {
    return fail(e.code, e.msg);
  }
  *prob = ev.prob;
  *accepted = ev.accepted ? 1 : 0;
  *status = ev.status;
  return 0;
}

}  // extern "C"
