"""The BASELINE.json configurations as `BGPConfig`s (shapes from /root/reference/src/mopt/Examples.jl
and SURVEY.md section 8d).  Pure host-side data; used by bench.py and the tests."""
from __future__ import annotations

import functools

import numpy as np

from ._abi import BGPConfig, SMM_OBJ_NORM, SMM_OBJ_NORM_MV, SMM_OBJ_NORM_SLOW, SMM_OBJ_PANEL


@functools.lru_cache(maxsize=64)
def _ladder(n_chains: int, maxtemp: float) -> np.ndarray:
    t = np.ones(1) if n_chains == 1 else np.linspace(1.0, maxtemp, n_chains)
    t.setflags(write=False)          # shared between callers
    return t


def temperature_ladder(n_chains: int, maxtemp: float) -> np.ndarray:
    """`temps = range(1.0, stop=maxtemp, length=N)` (AlgoBGP.jl:508); a single chain has no ladder (:526).
    (Read-only, cached: the constructor of a short run is inside the timed region.)"""
    return _ladder(int(n_chains), float(maxtemp))


def c1_serial_normal(niter: int = 200, slow: bool = False, slow_seconds: float = 0.1, **kw) -> BGPConfig:
    """`SMM.serialNormal(2, niter)` (Examples.jl:118-153 -> snorm_impl :373-446): 3 chains, 2 params."""
    n = 3
    temps = temperature_ladder(n, 5.0)
    args = dict(
        lb=[-3.0, -20.0], ub=[3.0, 20.0], init=[0.2, -0.2],
        data_mom=[-1.0, 10.0], data_w=[1.0, 1.0],
        n_chains=n, max_iter=niter,
        sigma0=0.05 * temps, acc_tuner=[20.0, 2.0, 1.0], min_improve=[0.0] * n,
        objective_id=SMM_OBJ_NORM_SLOW if slow else SMM_OBJ_NORM, slow_seconds=slow_seconds,
        n_sim=10000, seed_sim=1234, seed_algo=12,
    )
    args.update(kw)
    return BGPConfig(**args)


def mvnormal(n_chains: int = 256, niter: int = 1000, n_params: int = 8, **kw) -> BGPConfig:
    """C2/C3: MvNormal SMM with 2P moments (P means + P variances), S = 10 000 draws per evaluation."""
    P = n_params
    base_means = [-1.0, 1.0, 0.5, -0.5, 0.7, -0.7, 0.3, -0.3]
    means = [base_means[k % 8] for k in range(P)]
    temps = temperature_ladder(n_chains, 5.0)
    tuners = np.geomspace(20.0, 1.0, n_chains) if n_chains > 1 else np.array([20.0])
    args = dict(
        lb=[-3.0] * P, ub=[3.0] * P, init=[0.2 * (-1.0) ** (k + 1) for k in range(P)],
        data_mom=means + [1.0] * P, data_w=[1.0] * (2 * P),
        n_chains=n_chains, max_iter=niter,
        sigma0=0.05 * temps, acc_tuner=tuners, min_improve=[0.0] * n_chains,
        objective_id=SMM_OBJ_NORM_MV, n_sim=10000, seed_sim=1234, seed_algo=20261017,
        # upstream's default of 1000 rejection attempts (AlgoBGP.jl:520) makes the reference algorithm itself
        # abort at this scale: a hot chain in a corner of the 8-dim box has P(in support) ~ 0.5^8 per attempt,
        # and with 256+ chains x 1000 iterations some proposal exhausts 1000 attempts (-> `error`, :409).
        smpl_iters=100000,
    )
    args.update(kw)
    return BGPConfig(**args)


def normal_means(n_chains: int = 256, niter: int = 1000, n_params: int = 8, **kw) -> BGPConfig:
    """`objfunc_norm` itself (P == M, means only) at the C2 chain count, for a like-for-like."""
    cfg = mvnormal(n_chains, niter, n_params, **kw)
    P = n_params
    cfg.data_mom = list(cfg.data_mom[:P])
    cfg.data_w = [1.0] * P
    cfg.objective_id = kw.get("objective_id", SMM_OBJ_NORM)
    return cfg


def slow_normal(n_chains: int = 64, niter: int = 2000, slow_seconds: float = 0.1, **kw) -> BGPConfig:
    """C5: `objfunc_norm_slow` (C1's problem + slow_seconds per evaluation) on n_chains chains."""
    temps = temperature_ladder(n_chains, 5.0)
    args = dict(
        lb=[-3.0, -20.0], ub=[3.0, 20.0], init=[0.2, -0.2],
        data_mom=[-1.0, 10.0], data_w=[1.0, 1.0],
        n_chains=n_chains, max_iter=niter,
        sigma0=0.05 * temps, acc_tuner=np.geomspace(20.0, 1.0, n_chains), min_improve=[0.0] * n_chains,
        objective_id=SMM_OBJ_NORM_SLOW, slow_seconds=slow_seconds, n_sim=10000, seed_sim=1234, seed_algo=12,
    )
    args.update(kw)
    return BGPConfig(**args)


def panel_box(K: int = 8):
    """C4 parameter box: theta = (rho, beta[K], phi[K], sigma_alpha, sigma_eps, mu0)."""
    lb = [0.0] + [-2.0] * K + [0.0] * K + [0.1, 0.1, -2.0]
    ub = [0.95] + [2.0] * K + [0.95] * K + [2.0, 2.0, 2.0]
    return np.array(lb), np.array(ub)


def dynamic_panel(n_chains: int = 512, niter: int = 200, K: int = 8, T: int = 50, n_ind: int = 5000,
                  data_mom=None, **kw) -> BGPConfig:
    """C4: dynamic-panel SMM, P = 2K+4 params, M = 4K+8 moments, panel of n_ind individuals x T periods.

    `data_mom` are the moments the chains are matched to: the simulation at theta0 = mid-box with sim seed 4321
    (`panel_data_moments`); weights = max(|data|, 0.1).  Without them the config carries placeholders (zeros,
    unit weights) and is only good for bare objective evaluations."""
    lb, ub = panel_box(K)
    M = 4 * K + 8
    dm = np.zeros(M) if data_mom is None else np.asarray(data_mom, dtype=np.float64)
    w = np.ones(M) if data_mom is None else np.maximum(np.abs(dm), 0.1)
    temps = temperature_ladder(n_chains, 5.0)
    tuners = np.geomspace(20.0, 1.0, n_chains) if n_chains > 1 else np.array([20.0])
    args = dict(
        lb=lb, ub=ub, init=lb + 0.4 * (ub - lb), data_mom=dm, data_w=w,
        n_chains=n_chains, max_iter=niter,
        sigma0=0.02 * temps, acc_tuner=tuners, min_improve=[0.0] * n_chains,
        objective_id=SMM_OBJ_PANEL, panel_K=K, panel_T=T, panel_N=n_ind,
        n_sim=2, seed_sim=1234, seed_algo=20261017, smpl_iters=100000, exchange_mode=0,
    )
    args.update(kw)
    return BGPConfig(**args)


def panel_data_moments(evaluate, K: int = 8, T: int = 50, n_ind: int = 5000) -> np.ndarray:
    """Moments of the panel simulated at theta0 = mid-box with sim seed 4321.  `evaluate(cfg, params)`
    returns (value, moments, status) of the bare objective -- the GPU's `BGPHandle.eval_batch` in the product
    (`panel_data_moments_gpu`), the oracle's in the CPU tests."""
    lb, ub = panel_box(K)
    cfg = dynamic_panel(1, 1, K, T, n_ind, seed_sim=4321)
    _, mom, status = evaluate(cfg, 0.5 * (lb + ub))
    assert int(np.asarray(status).reshape(-1)[0]) == 1
    return np.asarray(mom, dtype=np.float64).reshape(-1)


def panel_data_moments_gpu(K: int = 8, T: int = 50, n_ind: int = 5000, device: int = 0) -> np.ndarray:
    from . import _lib

    def ev(cfg, params):
        cfg.device = device
        with _lib.BGPHandle(cfg) as h:
            return h.eval_batch(params)
    return panel_data_moments(ev, K, T, n_ind)
