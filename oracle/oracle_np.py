"""Independent numpy/pure-Python re-derivation of the oracle -- TEST INFRASTRUCTURE.

Purpose: pin oracle/smm_oracle.cpp (and the shared stream header it includes) against a second
implementation that shares no code with it: Philox4x32-10 written in numpy integer arithmetic, the
ziggurat re-derived with mpmath tables and libm exp/log, the Box-Muller transform evaluated with libm
(np.log / np.cos / np.sin, so both also check the accuracy of the header's polynomial kernels), and the BGP algorithm written as plain Python loops following the same
reference lines (AlgoBGP.jl:209-257, 272-471, 589-749; ObjExamples.jl:59-116; mprob.jl:246-272).
Small cases only.
"""
from __future__ import annotations

import math

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
STREAM_SIM, STREAM_PROP, STREAM_ACC, STREAM_PAIR = 1, 2, 3, 4


def philox(c0, c1, c2, c3, k0, k1):
    """vectorised over numpy uint64 arrays holding 32-bit values"""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    return c0, c1, c2, c3


def normal_pairs(x, y, z, w):
    """the transform of smm_stream.h, evaluated with libm in float64"""
    x, y, z, w = [np.asarray(v, dtype=np.uint64) for v in (x, y, z, w)]
    A = ((x << np.uint64(20)) | (y >> np.uint64(12))) | np.uint64(1)
    u1 = (np.float64(2 ** 52) - A.astype(np.float64)) / np.float64(2 ** 52)
    B = (z << np.uint64(20)) | (w >> np.uint64(12))
    g = (B & np.uint64(2 ** 49 - 1)).astype(np.float64) / np.float64(2 ** 49)
    rad = np.sqrt(-2.0 * np.log(u1))
    c, s = np.cos(np.pi / 4 * g), np.sin(np.pi / 4 * g)
    swap = ((z >> np.uint64(29)) & np.uint64(1)).astype(bool)
    c, s = np.where(swap, s, c), np.where(swap, c, s)
    c = np.where(((z >> np.uint64(30)) & np.uint64(1)).astype(bool), -c, c)
    s = np.where(((z >> np.uint64(31)) & np.uint64(1)).astype(bool), -s, s)
    return rad * c, rad * s


_ZIG = None
ZIG_LAYERS = 512


def zig_tables():
    """The 512-layer Marsaglia-Tsang ziggurat re-derived here with mpmath from its defining equations (equal-area
    layers under exp(-x^2/2), closing condition solved for R by bisection) -- not read from the C header -- followed by
    the documented table rule of include/smm_stream.h: going from the top layer down, W'[i] = double(W[i]) with its
    low 12 mantissa bits replaced by KH[i], the largest k <= 4096 W'[i+1] / W[i] with k / 4096 <= W'[i+1] / W'[i];
    F[i] = exp(-W'[i]^2 / 2).  Returns (W', KH, F, R' = W'[1])."""
    global _ZIG
    if _ZIG is None:
        import struct
        from fractions import Fraction
        import mpmath as mp
        NZ = ZIG_LAYERS
        with mp.workprec(160):
            f = lambda x: mp.exp(-x * x / 2)

            def build(R):
                V = R * f(R) + mp.sqrt(mp.pi / 2) * mp.erfc(R / mp.sqrt(2))
                xs = [V / f(R), R]
                for i in range(1, NZ - 1):
                    y = V / xs[i] + f(xs[i])
                    if y >= 1:
                        return None, xs
                    xs.append(mp.sqrt(-2 * mp.log(y)))
                return xs[NZ - 1] * (1 - f(xs[NZ - 1])) - V, xs

            lo, hi = mp.mpf("3.2"), mp.mpf("4.5")
            for _ in range(300):
                mid = (lo + hi) / 2
                res, xs = build(mid)
                if res is None or res < 0:
                    lo = mid
                else:
                    hi = mid
            _, xs = build((lo + hi) / 2)
            ideal = [float(x) for x in xs] + [0.0]
            Wp, KH = [0.0] * (NZ + 1), [0] * NZ
            for i in range(NZ - 1, -1, -1):
                base = struct.unpack("<Q", struct.pack("<d", ideal[i]))[0] & ~0xFFF
                k = int(mp.floor(4096 * mp.mpf(Wp[i + 1]) / xs[i]))
                while True:
                    cand = struct.unpack("<d", struct.pack("<Q", base | k))[0]
                    if Fraction(k, 4096) <= Fraction(Wp[i + 1]) / Fraction(cand):
                        break
                    k -= 1
                Wp[i], KH[i] = cand, k
            F = np.array([float(f(mp.mpf(x))) for x in Wp])
        _ZIG = (np.array(Wp), np.array(KH, dtype=np.uint64), F, Wp[1])
    return _ZIG


ZIG_TAG, ZIG_KEY0, ZIG_KEY1 = 0x5A494732, 0x736D6D5A, 0x69676733


def _u52(a, b):
    return ((int(a) << 20) | (int(b) >> 12))


def zig_select(w, t):
    """the 10-bit select field [sign:1][layer:9] of draw t in the block's fourth word"""
    w = int(w)
    return ((w >> 3) if t == 0 else (w >> 13) if t == 1 else ((w >> 23) | (w << 9))) & 0x3FF


def zig_slow(u0, s0):
    """the ziggurat's rare branch for one candidate (wedge test / base strip / tail / retry), with libm exp and log"""
    W, KH, F, R = zig_tables()
    u, s, n = int(u0), int(s0), 0
    while True:
        i = s & 0x1FF
        sgn = -1.0 if (s >> 9) else 1.0
        x = u * (W[i] / 2.0 ** 32)
        if (u >> 20) < int(KH[i]):
            return sgn * x
        if i == 0:
            if x < R:
                return sgn * x
            while True:
                n += 1
                r = [int(v) for v in philox(u0, s0, n, ZIG_TAG, ZIG_KEY0, ZIG_KEY1)]
                u1 = (2 ** 52 - (_u52(r[0], r[1]) | 1)) / 2.0 ** 52
                u2 = (2 ** 52 - (_u52(r[2], r[3]) | 1)) / 2.0 ** 52
                xt, yt = -math.log(u1) / R, -math.log(u2)
                if yt + yt > xt * xt:
                    return sgn * (R + xt)
        n += 1
        r = [int(v) for v in philox(u0, s0, n, ZIG_TAG, ZIG_KEY0, ZIG_KEY1)]
        uw = _u52(r[2], r[3]) / 2.0 ** 52
        if F[i] + uw * (F[i + 1] - F[i]) < math.exp(-0.5 * x * x):
            return sgn * x
        u, s = r[0], r[1] & 0x3FF


def zig_normals(u, s):
    """ziggurat normals from arrays of 32-bit uniform words and 10-bit select fields: numpy fast path, Python loop for
    the ~0.8 % rest"""
    W, KH, _, _ = zig_tables()
    u, s = np.asarray(u, dtype=np.uint64), np.asarray(s, dtype=np.uint64)
    i = (s & np.uint64(0x1FF)).astype(np.int64)
    z = np.where((s >> np.uint64(9)).astype(bool), -1.0, 1.0) * (u.astype(np.float64) * (W[i] / 2.0 ** 32))
    slow = np.nonzero((u >> np.uint64(20)) >= KH[i])[0]
    for t in slow:
        z[t] = zig_slow(u[t], s[t])
    return z


def sim_normals(seed_sim, k, S, noseed=0, uid=0, rep=0, transform="zig"):
    per = 3 if transform == "zig" else 2
    nb = (S + per - 1) // per
    j = np.arange(nb, dtype=np.uint64)
    c2 = uid if noseed else 0
    c3 = (STREAM_SIM << 28) | ((rep & 0x0FFFFFFF) if noseed else 0)
    r = philox(j, k, c2, c3, seed_sim & MASK, seed_sim >> 32)
    if transform == "zig":      # MvNormal objectives: three draws per block
        w = r[3]
        sel = [(w >> np.uint64(3)) & np.uint64(0x3FF), (w >> np.uint64(13)) & np.uint64(0x3FF),
               ((w >> np.uint64(23)) | (w << np.uint64(9))) & np.uint64(0x3FF)]
        zs = [zig_normals(r[t], sel[t]) for t in range(3)]
        return np.stack(zs, axis=1).reshape(-1)[:S]
    z0, z1 = normal_pairs(*r)  # dynamic panel
    return np.stack([z0, z1], axis=1).reshape(-1)[:S]


def prop_normal(seed_algo, chain, it, attempt, k):
    r = philox(attempt, k >> 1, chain, (STREAM_PROP << 28) | it, seed_algo & MASK, seed_algo >> 32)
    z0, z1 = normal_pairs(*r)
    return float(z1) if (k & 1) else float(z0)


def acc_uniform(seed_algo, chain, it):
    x, y, _, _ = philox(0, 0, chain, (STREAM_ACC << 28) | it, seed_algo & MASK, seed_algo >> 32)
    return ((int(x) << 20) | (int(y) >> 12)) / 2.0 ** 52


def pair_sample(seed_algo, it, N):
    n_all = N * (N - 1) // 2
    n_s = N - 1 if N < 3 else N
    props = [(i, j) for j in range(N) for i in range(N) if i < j]   # i fastest (AlgoBGP.jl:653)
    chosen = []
    for t in range(n_s):
        a = 0
        while True:
            x, y, _, _ = philox(t, a, 0, (STREAM_PAIR << 28) | it, seed_algo & MASK, seed_algo >> 32)
            q = (((int(x) << 32) | int(y)) * n_all) >> 64
            a += 1
            if q not in chosen:
                break
        chosen.append(q)
    return [props[q] for q in chosen]


def objective(cfg, p, uid=0, rep=0, noseed=None):
    """objfunc_norm / norm_mv with the draw matrix materialised (ObjExamples.jl:76-101)"""
    P, M, S = cfg.n_params, cfg.n_moments, cfg.n_sim
    noseed = cfg.noseed if noseed is None else noseed
    X = np.stack([p[k] + sim_normals(cfg.seed_sim, k, S, noseed, uid, rep) for k in range(P)])
    sim = X.mean(axis=1)
    if M == 2 * P:
        sim = np.concatenate([sim, X.var(axis=1, ddof=1)])
    d = (sim - np.asarray(cfg.data_mom, float)) / np.asarray(cfg.data_w, float)
    return float(np.mean(d * d)), sim


def panel_objective(cfg, p, uid=0, rep=0, noseed=None):
    """the dynamic-panel objective (oracle/smm_oracle.cpp::objfunc_panel) re-derived with numpy:
    the panel is materialised as arrays y[i,t], x[k,i,t] and every moment is a plain centred numpy average.
    (numpy has no fma, so this agrees with the C++ oracle to rounding, not to the bit.)"""
    K, T, NI = cfg.panel_K, cfg.panel_T, cfg.panel_N
    noseed = cfg.noseed if noseed is None else noseed
    p = np.asarray(p, float)
    rho, beta, phi = p[0], p[1:1 + K], p[1 + K:1 + 2 * K]
    sig_a, sig_e, mu0 = p[1 + 2 * K], p[2 + 2 * K], p[3 + 2 * K]
    nz = 1 + K + T * (K + 1)
    Z = np.stack([sim_normals(cfg.seed_sim, i, nz, noseed, uid, rep, transform="bm") for i in range(NI)])   # [NI][nz]
    alpha = mu0 + sig_a * Z[:, 0]
    y = np.zeros((NI, T + 1))
    x = np.zeros((K, NI, T + 1))
    y[:, 0] = alpha / (1.0 - rho)
    x[:, :, 0] = (Z[:, 1:1 + K] / np.sqrt(1.0 - phi * phi)).T
    for t in range(1, T + 1):
        zt = Z[:, 1 + K + (t - 1) * (K + 1): 1 + K + t * (K + 1)]
        x[:, :, t] = phi[:, None] * x[:, :, t - 1] + zt[:, :K].T
        y[:, t] = alpha + rho * y[:, t - 1] + (beta[:, None] * x[:, :, t]).sum(axis=0) + sig_e * zt[:, K]
    n = NI * T
    my = y[:, 1:].sum() / n
    mx = x[:, :, 1:].sum(axis=(1, 2)) / n
    sim = [my]
    for l in range(7):
        lo = max(l, 1)
        sim.append(((y[:, lo:] - my) * (y[:, lo - l:T + 1 - l] - my)).sum() / n)
    yc = y[:, 1:] - my
    xc = x - mx[:, None, None]
    sim += [(yc * xc[k, :, 1:]).sum() / n for k in range(K)]
    sim += [(yc * xc[k, :, :-1]).sum() / n for k in range(K)]
    sim += [(xc[k, :, 1:] * xc[k, :, :-1]).sum() / n for k in range(K)]
    sim += [(xc[k, :, 1:] ** 2).sum() / n for k in range(K)]
    sim = np.array(sim)
    d = (sim - np.asarray(cfg.data_mom, float)) / np.asarray(cfg.data_w, float)
    return float(np.mean(d * d)), sim


def run(cfg, n_iters):
    """run!(MAlgoBGP) as plain loops; returns dict of arrays shaped [n_iters][N](...)"""
    N, P, M = cfg.n_chains, cfg.n_params, cfg.n_moments
    lb, ub = np.asarray(cfg.lb, float), np.asarray(cfg.ub, float)
    sigma = np.asarray(cfg.sigma0, float).copy()
    tun, mi = np.asarray(cfg.acc_tuner, float), np.asarray(cfg.min_improve, float)
    bs = cfg.batch_size or P
    out = dict(value=np.zeros((n_iters, N)), prob=np.zeros((n_iters, N)), curr_val=np.zeros((n_iters, N)),
               best_val=np.zeros((n_iters, N)), params=np.zeros((n_iters, N, P)), sim_moments=np.zeros((n_iters, N, M)),
               accepted=np.zeros((n_iters, N), np.uint8), status=np.zeros((n_iters, N), np.int32),
               exchanged=np.zeros((n_iters, N), np.int32), best_id=np.zeros((n_iters, N), np.int32))
    evals = [[None] * n_iters for _ in range(N)]   # (value, prob, status, params, moments)

    def last_accepted(c, it):
        for t in range(it, 0, -1):
            if out["accepted"][t - 1, c]:
                return evals[c][t - 1]
        raise AssertionError

    def set_eval(c, it, ev, accepted):
        evals[c][it - 1] = ev
        i = it - 1
        out["value"][i, c], out["prob"][i, c], out["status"][i, c] = ev[0], ev[1], ev[2]
        out["params"][i, c], out["sim_moments"][i, c] = ev[3], ev[4]
        out["accepted"][i, c] = accepted
        if it == 1:
            out["best_val"][i, c] = out["curr_val"][i, c] = ev[0]
            out["best_id"][i, c] = 1
        else:
            out["curr_val"][i, c] = ev[0] if accepted else out["curr_val"][i - 1, c]
            if ev[0] < out["best_val"][i - 1, c]:
                out["best_val"][i, c], out["best_id"][i, c] = ev[0], it
            else:
                out["best_val"][i, c], out["best_id"][i, c] = out["best_val"][i - 1, c], out["best_id"][i - 1, c]

    for it in range(1, n_iters + 1):
        for c in range(N):
            if it == 1:
                pp = np.asarray(cfg.init, float).copy()
            else:
                old = last_accepted(c, it - 1)
                mu01 = (old[3] - lb) / (ub - lb)
                x01 = np.zeros(P)
                for b0 in range(0, P, bs):
                    for a in range(cfg.smpl_iters):
                        cand = np.array([mu01[k] + sigma[c] * prop_normal(cfg.seed_algo, c, it, a, k) for k in range(b0, b0 + bs)])
                        if np.all(cand >= 0) and np.all(cand <= 1):
                            x01[b0:b0 + bs] = cand
                            break
                pp = x01 * (ub - lb) + lb
            value, sim = (panel_objective if cfg.panel_K else objective)(cfg, pp, c, it)
            status = 1
            if it == 1:
                prob, accepted = 1.0, True
            else:
                old = last_accepted(c, it - 1)
                e = math.exp(tun[c] * (old[0] - value))
                prob = min(1.0, e)
                accepted = prob > acc_uniform(cfg.seed_algo, c, it)
            set_eval(c, it, (value, prob, status, pp, sim), accepted)
            if it > 1:
                noex = out["exchanged"][:it, c] == 0
                rate = out["accepted"][:it, c][noex].mean()
                if it % cfg.sigma_update_steps == 0:
                    sigma[c] = sigma[c] * (1.0 + cfg.sigma_adjust_by) if rate > 0.234 else sigma[c] * (1.0 - cfg.sigma_adjust_by)
        if it >= 2 and N > 1:
            for (i, j) in pair_sample(cfg.seed_algo, it, N):
                ei, ej = last_accepted(i, it), last_accepted(j, it)
                if ei[0] - ej[0] > mi[i]:
                    set_eval(i, it, ej, True)
                    set_eval(j, it, ei, True)
                    out["exchanged"][it - 1, i], out["exchanged"][it - 1, j] = j + 1, i + 1
    out["sigma"] = sigma
    return out
