# ObjB200.jl -- Julia CPU objective functions on the counter-indexed streams (include/smm_stream.h, SMMStreams.jl).
#
# SURVEY.md section 0, fact 3: every objective exists three times with identical results -- CUDA kernel
# (smm_jl_b200/csrc), C++ oracle (oracle/smm_oracle.cpp) and Julia CPU function (this file).  Registered with
# `addEvalFunc!(m, f)` exactly like the reference's own examples (mprob.jl:159-161; template ObjExamples.jl:59-116):
# on the stock `MAlgoBGP` they run on the CPU, on the B200 backend (`opts["backend"] = :b200`, AlgoBGPB200.jl) the
# same function object selects the device simulator with the same id.
#
#   objfunc_norm_b200   SMM_OBJ_NORM    objfunc_norm (ObjExamples.jl:59-116) with Zsim instead of Random.seed!(1234)
#   objfunc_norm_mv     SMM_OBJ_NORM_MV P means + P sample variances of the same draw matrix (M == 2P)
#   objfunc_panel       SMM_OBJ_PANEL   dynamic panel, P == 2K+4, M == 4K+8 (SURVEY.md 8d; no upstream code)
#
# Options: keyword arguments (`evaluateObjective` splats `m.objfunc_opts` into the call, mprob.jl:182) or entries of
# `ev.options` (where `evaluateObjective(m, p; noseed=true)` puts :noseed, mprob.jl:177-179): :noseed (fresh draws,
# indexed by :uid and :rep), :seed (1234), :n_sim (10000), :panel_T, :panel_N, :panel_K.
#
# Include inside module SMM after ObjExamples.jl:   include("mopt/SMMStreams.jl"); include("mopt/ObjB200.jl")
# NOT EXECUTED in the build environment (no Julia in the image).

using .SMMStreams

_opt(ev::Eval, kw, k::Symbol, default) = haskey(kw, k) ? kw[k] : get(ev.options, k, default)

"value = mean_k ((sim_k - data_k) / w_k)^2 in moment order (ObjExamples.jl:90-101; the weight divides)"
function _weighted_distance!(ev::Eval, sim::Vector{Float64})
    acc = 0.0
    simMoments = Dict{Symbol,Float64}()
    i = 0
    for (k, mom) in dataMomentd(ev)
        i += 1
        simMoments[k] = sim[i]
        w = haskey(dataMomentWd(ev), k) ? dataMomentW(ev, k) : 1.0
        d = (sim[i] - mom) / w
        acc += d * d
    end
    setValue!(ev, acc / i)
    setMoments!(ev, simMoments)
    ev.status = 1
    return ev
end

"the D x S draw matrix X[k,s] = mu[k] + Zsim[k,s], materialised as the reference does (ObjExamples.jl:78)"
function _draw_matrix(ev::Eval, kw, mu::Vector{Float64})
    S = _opt(ev, kw, :n_sim, 10000)
    seed = UInt64(_opt(ev, kw, :seed, 1234))
    noseed = _opt(ev, kw, :noseed, false)
    uid, rep = _opt(ev, kw, :uid, 0), _opt(ev, kw, :rep, 0)
    X = Matrix{Float64}(undef, length(mu), S)
    for k in 1:length(mu)
        X[k, :] .= mu[k] .+ sim_normals(seed, k - 1, S; noseed = noseed, uid = uid, rep = rep)
    end
    return X
end

"sequential row sums along s, the order the oracle uses (any order agrees to ~1e-16 relative)"
function _row_means(X::Matrix{Float64})
    D, S = size(X)
    m = zeros(D)
    for s in 1:S, k in 1:D
        m[k] += X[k, s]
    end
    return m ./ S
end

function objfunc_norm_b200(ev::Eval; kw...)
    start(ev)
    mu = collect(values(ev.params))
    length(mu) == length(ev.dataMoments) || error("objfunc_norm needs #params == #moments (ObjExamples.jl:77-78)")
    _weighted_distance!(ev, _row_means(_draw_matrix(ev, kw, mu)))
    finish(ev)
    return ev
end

function objfunc_norm_mv(ev::Eval; kw...)
    start(ev)
    mu = collect(values(ev.params))
    D = length(mu)
    length(ev.dataMoments) == 2D || error("objfunc_norm_mv needs #moments == 2 #params")
    X = _draw_matrix(ev, kw, mu)
    S = size(X, 2)
    means = _row_means(X)
    vars = zeros(D)
    for s in 1:S, k in 1:D
        d = X[k, s] - means[k]
        vars[k] += d * d
    end
    _weighted_distance!(ev, vcat(means, vars ./ (S - 1)))          # two-pass, n-1 (Julia's var)
    finish(ev)
    return ev
end

"""
    objfunc_panel(ev)

theta = (rho, beta[1:K], phi[1:K], sigma_alpha, sigma_eps, mu0).  Individual i reads stream row i-1 (Box-Muller):
normal 0 = a_i, 1..K = initial x shocks, then for t = 1..T the K regressor shocks followed by eps_t.  The recurrences
are DEFINED with `fma` (as in oracle/smm_oracle.cpp::objfunc_panel and the device's DFMA).  Moments (4K+8), pooled over
(i, t = 1..T): mean y; var y; autocov_y lags 1-6; cov(y_t, x_kt) x K; cov(y_t, x_k,t-1) x K; autocov_xk lag 1 x K;
var x_k x K.
"""
function objfunc_panel(ev::Eval; kw...)
    start(ev)
    th = collect(values(ev.params))
    K, T, NI = _opt(ev, kw, :panel_K, 8), _opt(ev, kw, :panel_T, 50), _opt(ev, kw, :panel_N, 5000)
    length(th) == 2K + 4 || error("objfunc_panel needs 2K+4 parameters")
    seed = UInt64(_opt(ev, kw, :seed, 1234))
    noseed = _opt(ev, kw, :noseed, false)
    uid, rep = _opt(ev, kw, :uid, 0), _opt(ev, kw, :rep, 0)
    rho = th[1]; beta = th[2:K+1]; phi = th[K+2:2K+1]
    sig_a, sig_e, mu0 = th[2K+2], th[2K+3], th[2K+4]
    nz = 1 + K + T * (K + 1)
    y = zeros(NI, T + 1)                  # y[i, t+1], t = 0..T
    x = zeros(K, NI, T + 1)
    for i in 1:NI
        z = sim_normals(seed, i - 1, nz; noseed = noseed, uid = uid, rep = rep, transform = :bm)
        alpha = fma(sig_a, z[1], mu0)
        yc = alpha / (1.0 - rho)
        xc = [z[1 + k] / sqrt(fma(-phi[k], phi[k], 1.0)) for k in 1:K]
        x[:, i, 1] .= xc
        y[i, 1] = yc
        for t in 1:T
            o = 1 + K + (t - 1) * (K + 1)            # zt[k] = z[o + k], k = 1..K+1
            xb = 0.0
            for k in 1:K
                xc[k] = fma(phi[k], xc[k], z[o + k])
                x[k, i, t + 1] = xc[k]
                xb = fma(beta[k], xc[k], xb)
            end
            yc = fma(sig_e, z[o + K + 1], fma(rho, yc, alpha) + xb)
            y[i, t + 1] = yc
        end
    end
    n = Float64(NI) * Float64(T)
    my = sum(y[:, 2:end]) / n
    mx = [sum(x[k, :, 2:end]) / n for k in 1:K]
    sim = Float64[my]
    for l in 0:6
        a = 0.0
        for i in 1:NI, t in 1:T
            t - l < 0 && continue
            a += (y[i, t + 1] - my) * (y[i, t - l + 1] - my)
        end
        push!(sim, a / n)
    end
    for k in 1:K
        push!(sim, sum((y[i, t + 1] - my) * (x[k, i, t + 1] - mx[k]) for i in 1:NI, t in 1:T) / n)
    end
    for k in 1:K
        push!(sim, sum((y[i, t + 1] - my) * (x[k, i, t] - mx[k]) for i in 1:NI, t in 1:T) / n)
    end
    for k in 1:K
        push!(sim, sum((x[k, i, t + 1] - mx[k]) * (x[k, i, t] - mx[k]) for i in 1:NI, t in 1:T) / n)
    end
    for k in 1:K
        push!(sim, sum((x[k, i, t + 1] - mx[k])^2 for i in 1:NI, t in 1:T) / n)
    end
    _weighted_distance!(ev, sim)
    finish(ev)
    return ev
end
