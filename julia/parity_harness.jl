# parity_harness.jl -- run the REAL SMM.jl BGP iteration on injected random streams and dump its trace.
#
#     python tools/dump_streams.py                                         # writes tests/golden/julia/<case>/ (committed)
#     julia --project=<SMM.jl checkout> julia/parity_harness.jl tests/golden/julia/c1_serial_normal [--check-port]
#     python -m pytest tests/test_julia_parity.py                          # oracle (and, -m gpu, the CUDA path) vs <case>/julia_trace/
#
# Why: the reference draws from Julia's global RNG and from an unseedable RandomDevice (SMM.jl:59-60), so no seed makes
# two runs comparable.  Parity is therefore defined on the reference's algorithm as a function of four injected streams
# (include/smm_stream.h).  This script replaces ONLY the four randomness sites and leaves every line of the algorithm --
# proposal / mapto_01 / mapto_ab, evaluateObjective, doAcceptReject!, set_eval!, set_acceptRate!, sigma adaptation,
# exchangeMoves!, swap_ev_ij! -- to the package:
#
#   site                                                   upstream                          here
#   BGPChain.probs_acc   AlgoBGP.jl:85                     rand(n)                           overwritten with Uacc[chain, :] after construction
#   mysample             AlgoBGP.jl:400-410                rand(RAND, d), RAND=RandomDevice  same loop, rand(PROP_RNG, d): randn -> Zprop[chain, iter, attempt, k]
#   exchangeMoves!       AlgoBGP.jl:656                    sample(props, n, replace=false)   a method of `sample` for Vector{Tuple{Int,Int}} -> Pairs[iter]
#   objective            ObjExamples.jl:74-79              Random.seed!(1234); rand(MvNormal) a user objective (the package's own plug-in point,
#                                                                                            mprob.jl:159) with the body of objfunc_norm and X = mu .+ Zsim
#
# Output, <case>/julia_trace/: value prob curr_val best_val (f64, [I][N] in C order), params [I][N][P], sim_moments
# [I][N][M], accepted (u8), status exchanged best_id (i32), sigma accept_rate (f64 [N]), meta.txt (Julia / package versions).
#
# NOT EXECUTED in the build environment (no Julia in the image).  Tested syntax-level only; the stream files and the
# comparison side (tests/test_julia_parity.py) are exercised by the CPU test-suite against the oracle's own trace.

using SMM
using Random, Statistics
using OrderedCollections, DataFrames
const Distributions = SMM.Distributions
const StatsBase = Distributions.StatsBase

dir = ARGS[1]
check_port = "--check-port" in ARGS

# ---- stream files --------------------------------------------------------------------------------------------------
meta = Dict{String,String}()
for ln in eachline(joinpath(dir, "meta.txt"))
    k, v = split(ln, "="; limit = 2)
    meta[k] = v
end
geti(k) = parse(Int, meta[k]); getf(k) = parse(Float64, meta[k])
N, P, M, S, I, A, NS = geti("n_chains"), geti("n_params"), geti("n_moments"), geti("n_sim"), geti("n_iter"), geti("n_attempts"), geti("n_pairs")
readf(name, dims...) = (a = Array{Float64}(undef, dims...); read!(joinpath(dir, name), a); a)
lb, ub, init = readf("lb.f64", P), readf("ub.f64", P), readf("init.f64", P)
data_mom, data_w = readf("data_mom.f64", M), readf("data_w.f64", M)
acc_tuner, min_improve = readf("acc_tuner.f64", N), readf("min_improve.f64", N)
const ZSIM = readf("zsim.f64", S, P)                  # C [P][S]        -> Julia (S, P)
const ZPROP = readf("zprop.f64", P, A, I, N)          # C [N][I][A][P]  -> Julia (P, A, I, N)
const UACC = readf("uacc.f64", I, N)                  # C [N][I]        -> Julia (I, N)
const PAIRS = (a = Array{Int32}(undef, 2, NS, I); read!(joinpath(dir, "pairs.i32"), a); a)   # C [I][n_s][2]

# ---- injection context ---------------------------------------------------------------------------------------------
mutable struct Ctx
    iter::Int          # iteration being computed (= algo.i)
    calls::Int         # mysample calls so far in this iteration: call = chain * n_batches + batch (serial map over chains)
    n_batches::Int
    max_attempt::Int
end
const CTX = Ctx(0, 0, P ÷ geti("batch_size"), 0)

"an RNG whose randn is the next element of Zprop[chain, iter, attempt, k0 + 0, 1, ...]"
mutable struct PropRNG <: Random.AbstractRNG
    chain::Int; iter::Int; attempt::Int; k::Int        # all 1-based indices into ZPROP
end
const PROP_RNG = PropRNG(1, 1, 1, 1)
function Base.randn(r::PropRNG, ::Type{Float64} = Float64)
    z = ZPROP[r.k, r.attempt, r.iter, r.chain]
    r.k += 1
    return z
end
function Random.randn!(r::PropRNG, a::AbstractArray{Float64})      # rand(rng, ::MvNormal) fills with randn! and unwhitens
    for i in eachindex(a)
        a[i] = randn(r, Float64)
    end
    return a
end

# mysample (AlgoBGP.jl:400-410): the same rejection loop; the draw comes from PROP_RNG instead of RAND
@eval SMM function mysample(d::Distributions.MultivariateDistribution, lb::Float64, ub::Float64, iters::Int)
    ctx, rng = Main.CTX, Main.PROP_RNG
    chain, batch = divrem(ctx.calls, ctx.n_batches)
    ctx.calls += 1
    for i in 1:iters
        i <= size(Main.ZPROP, 2) || error("parity harness: more than $(size(Main.ZPROP, 2)) attempts needed; dump more (tools/dump_streams.py N_ATTEMPTS)")
        rng.chain, rng.iter, rng.attempt, rng.k = chain + 1, ctx.iter, i, batch * length(d) + 1
        x = rand(rng, d)
        ctx.max_attempt = max(ctx.max_attempt, i)
        if all(x .>= lb) && all(x .<= ub)
            return x
        end
    end
    error("no draw in support after $iters trials. increase either opts[smpl_iters] or opts[bound_prob].")
end

# the pair sample of exchangeMoves! (AlgoBGP.jl:656): Pairs[iter], in the stream's order
function StatsBase.sample(a::Vector{Tuple{Int,Int}}, n::Integer; replace::Bool = true, ordered::Bool = false)
    n == size(PAIRS, 2) || error("parity harness: expected $(size(PAIRS, 2)) pairs, asked for $n")
    return [(Int(PAIRS[1, t, CTX.iter]), Int(PAIRS[2, t, CTX.iter])) for t in 1:n]
end

# the objective: the body of objfunc_norm (ObjExamples.jl:59-116) with X = mu .+ Zsim instead of rand(MvNormal(mu, I), ns)
function objfunc_inject(ev::SMM.Eval; kw...)
    SMM.start(ev)
    mu = collect(values(ev.params))
    X = mu .+ permutedims(ZSIM)                                     # (P, S): X[k, s] = mu[k] + Zsim[k, s]
    simM = vec(mean(X, dims = 2))
    if meta["objective"] == "norm_mv"                               # moments P+1..2P: row sample variances (n-1)
        simM = vcat(simM, vec(var(X, dims = 2)))
    end
    v = Dict{Symbol,Float64}()
    simMoments = Dict{Symbol,Float64}()
    i = 0
    for (k, mom) in SMM.dataMomentd(ev)
        i += 1
        simMoments[k] = simM[i]
        v[k] = ((simMoments[k] .- mom) ./ SMM.dataMomentW(ev, k)) .^ 2
    end
    SMM.setValue!(ev, mean(collect(values(v))))
    SMM.setMoments!(ev, simMoments)
    ev.status = 1
    SMM.finish(ev)
    return ev
end

# ---- the problem and the algorithm: stock SMM.jl calls ------------------------------------------------------------------
pb = OrderedDict("p$k" => [init[k], lb[k], ub[k]] for k in 1:P)
moms = DataFrame(name = ["m$k" for k in 1:M], value = data_mom, weight = data_w)
mprob = SMM.MProb()
SMM.addSampledParam!(mprob, pb)
SMM.addMoment!(mprob, moms)
SMM.addEvalFunc!(mprob, objfunc_inject)
opts = Dict("N" => N, "maxiter" => I, "maxtemp" => getf("maxtemp"), "sigma" => getf("sigma"),
            "sigma_update_steps" => geti("sigma_update_steps"), "sigma_adjust_by" => getf("sigma_adjust_by"),
            "smpl_iters" => geti("smpl_iters"), "batch_size" => geti("batch_size"), "parallel" => false,
            "min_improve" => min_improve, "acc_tuners" => acc_tuner, "animate" => false)
MA = SMM.MAlgoBGP(mprob, opts)
for c in MA.chains
    c.probs_acc .= UACC[:, c.id]                                    # BGPChain.probs_acc = Uacc[chain, :]
end
for it in 1:I                                                        # run!'s loop (AlgoAbstract.jl:38-45) without the progress bar
    MA.i = it
    CTX.iter, CTX.calls = it, 0
    SMM.computeNextIteration!(MA)
end
println("ran $I iterations of $N chains; most attempts used by one proposal: $(CTX.max_attempt)")

# ---- dump the trace ------------------------------------------------------------------------------------------------------
out = joinpath(dir, "julia_trace"); mkpath(out)
pn = collect(keys(mprob.params_to_sample)); mn = collect(keys(mprob.moments))
value = zeros(N, I); prob = zeros(N, I); curr = zeros(N, I); best = zeros(N, I)
params = zeros(P, N, I); smom = fill(NaN, M, N, I)
acc = zeros(UInt8, N, I); status = zeros(Int32, N, I); exch = zeros(Int32, N, I); bid = zeros(Int32, N, I)
for (ic, c) in enumerate(MA.chains), it in 1:I
    ev = c.evals[it]
    value[ic, it] = ev.value; prob[ic, it] = ev.prob; status[ic, it] = ev.status
    curr[ic, it] = c.curr_val[it]; best[ic, it] = c.best_val[it]; bid[ic, it] = c.best_id[it]
    acc[ic, it] = c.accepted[it]; exch[ic, it] = c.exchanged[it]
    for (j, k) in enumerate(pn); params[j, ic, it] = ev.params[k]; end
    for (j, k) in enumerate(mn); haskey(ev.simMoments, k) && (smom[j, ic, it] = ev.simMoments[k]); end
end
for (name, a) in (("value.f64", value), ("prob.f64", prob), ("curr_val.f64", curr), ("best_val.f64", best), ("params.f64", params),
                  ("sim_moments.f64", smom), ("accepted.u8", acc), ("status.i32", status), ("exchanged.i32", exch), ("best_id.i32", bid),
                  ("sigma.f64", Float64[c.sigma for c in MA.chains]), ("accept_rate.f64", Float64[c.accept_rate for c in MA.chains]))
    write(joinpath(out, name), a)
end
open(joinpath(out, "meta.txt"), "w") do f
    println(f, "julia=$(VERSION)")
    println(f, "smm=$(pkgdir(SMM))")
    println(f, "case=$(meta["case"])")
    println(f, "max_attempt=$(CTX.max_attempt)")
end
println("wrote $out")

# ---- optional: the Julia port of the stream definitions against the dumped values -----------------------------------------
if check_port
    include(joinpath(@__DIR__, "SMMStreams.jl"))
    using .SMMStreams
    sa, ss = UInt64(parse(Int, meta["seed_algo"])), UInt64(parse(Int, meta["seed_sim"]))
    bad = 0
    for k in 1:P
        bad += count(SMMStreams.sim_normals(ss, k - 1, S) .!== ZSIM[:, k])
    end
    for c in 1:N, it in 2:I, a in 1:A, k in 1:P
        bad += SMMStreams.prop_normal(sa, c - 1, it, a - 1, k - 1) !== ZPROP[k, a, it, c]
    end
    for c in 1:N, it in 1:I
        bad += SMMStreams.acc_uniform(sa, c - 1, it) !== UACC[it, c]
    end
    for it in 2:I
        bad += SMMStreams.pair_sample(sa, it, N) != [(Int(PAIRS[1, t, it]), Int(PAIRS[2, t, it])) for t in 1:NS]
    end
    println(bad == 0 ? "SMMStreams.jl reproduces every dumped stream element bit for bit" : "SMMStreams.jl: $bad mismatches")
    bad == 0 || exit(1)
end
