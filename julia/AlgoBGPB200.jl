# AlgoBGPB200.jl -- the reference-side binding of libsmm_b200.so.
#
# Drop this file into SMM.jl's src/mopt/ and `include("mopt/AlgoBGPB200.jl")` after AlgoBGP.jl in
# src/SMM.jl (the reference's algorithm plug-in point is a subtype of MAlgo that implements
# computeNextIteration!, README.md:105-107, AlgoAbstract.jl:8,45).  User code keeps calling
# addSampledParam!/addMoment!/addEvalFunc! and run!; only the constructor changes:
#
#     MA = MAlgoBGPB200(mprob, opts)      # instead of MAlgoBGP(mprob, opts)
#     run!(MA); summary(MA); history(MA.chains[1])
#
# NOT EXECUTED in the build environment (no Julia in the image): written against include/smm_b200.h,
# whose struct layout tests/test_abi_and_host.py pins with a C compiler.  Host code stays Julia; the
# only foreign calls are the `ccall`s below (no CUDA.jl needed: the library owns device, stream, memory).

const LIBSMM_B200 = get(ENV, "SMM_B200_LIB", "libsmm_b200.so")
const SMM_ABI_VERSION = Int32(1)

# objective ids (include/smm_b200.h)
const SMM_OBJ = Dict{Function,Int32}(objfunc_norm => 0, objfunc_norm_slow => 1)
# objfunc_norm_mv / objfunc_panel / Testobj_fails: ids 2 / 3 / 4 (define Julia CPU versions with the
# streams of include/smm_stream.h to use them on both sides)

# mirrors `struct smm_bgp_config` field by field
struct SmmBgpConfig
    abi_version::Int32
    n_params::Int32
    n_moments::Int32
    lb::Ptr{Cdouble}
    ub::Ptr{Cdouble}
    init::Ptr{Cdouble}
    data_mom::Ptr{Cdouble}
    data_w::Ptr{Cdouble}
    objective_id::Int32
    n_sim::Int32
    seed_sim::UInt64
    noseed::Int32
    slow_seconds::Cdouble
    panel_T::Int32
    panel_N::Int32
    panel_K::Int32
    n_chains::Int32
    max_iter::Int32
    sigma0::Ptr{Cdouble}
    acc_tuner::Ptr{Cdouble}
    min_improve::Ptr{Cdouble}
    sigma_update_steps::Int32
    sigma_adjust_by::Cdouble
    smpl_iters::Int32
    batch_size::Int32
    seed_algo::UInt64
    device::Int32
    world_size::Int32
    rank::Int32
    nccl_id::NTuple{128,UInt8}
    exchange_mode::Int32
    n_split::Int32
end

# mirrors `struct smm_trace_view`
struct SmmTraceView
    value::Ptr{Cdouble}
    prob::Ptr{Cdouble}
    curr_val::Ptr{Cdouble}
    best_val::Ptr{Cdouble}
    params::Ptr{Cdouble}
    sim_moments::Ptr{Cdouble}
    accepted::Ptr{UInt8}
    status::Ptr{Int32}
    exchanged::Ptr{Int32}
    best_id::Ptr{Int32}
end

smm_last_error() = unsafe_string(ccall((:smm_last_error, LIBSMM_B200), Cstring, ()))
smm_check(rc) = rc == 0 ? nothing : error("libsmm_b200: $(smm_last_error()) (code $rc)")

mutable struct MAlgoBGPB200 <: MAlgo
    m::MProb
    opts::Dict
    i::Int
    chains::Array{BGPChain}
    anim::Plots.Animation
    dist_fun::Function
    handle::Ptr{Cvoid}
    synced::Int          # iterations already copied into `chains`

    function MAlgoBGPB200(m::MProb, opts::Dict)
        haskey(opts, "dist_fun") && error("dist_fun: only the default `-` runs on the device")
        haskey(SMM_OBJ, m.objfunc) || error("$(m.objfunc) has no device simulator (no CPU fallback in the B200 path)")
        collect(keys(m.initial_value)) == collect(keys(m.params_to_sample)) ||
            error("every parameter must be sampled (proposal broadcasts params against lb/ub, AlgoBGP.jl:430-436)")
        N = opts["N"]; n = opts["maxiter"]
        np = length(m.params_to_sample)
        temps = N > 1 ? collect(range(1.0, stop = opts["maxtemp"], length = N)) : [1.0]        # AlgoBGP.jl:508
        sigma0 = get(opts, "sigma", 0.05) .* temps
        tuners = Float64.(get(opts, "acc_tuners", [2.0 for j in 1:N]))
        minimp = Float64.(get(opts, "min_improve", [0.5 for j in 1:N]))
        lb = Float64[v[:lb] for (k, v) in m.params_to_sample]
        ub = Float64[v[:ub] for (k, v) in m.params_to_sample]
        init = Float64[m.initial_value[k] for k in keys(m.params_to_sample)]
        dm = Float64[v[:value] for (k, v) in m.moments]
        dw = Float64[v[:weight] for (k, v) in m.moments]
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve lb ub init dm dw sigma0 tuners minimp begin
            cfg = SmmBgpConfig(SMM_ABI_VERSION, np, length(dm), pointer(lb), pointer(ub), pointer(init), pointer(dm), pointer(dw),
                SMM_OBJ[m.objfunc], get(m.objfunc_opts, :n_sim, 10000), UInt64(get(m.objfunc_opts, :seed, 1234)),
                Int32(get(m.objfunc_opts, :noseed, false)), 0.1, 0, 0, 0,
                N, n, pointer(sigma0), pointer(tuners), pointer(minimp),
                get(opts, "sigma_update_steps", 10), get(opts, "sigma_adjust_by", 0.01), get(opts, "smpl_iters", 1000),
                get(opts, "batch_size", np), UInt64(get(opts, "seed", 20261017)),
                get(opts, "device", 0), 1, 0, ntuple(_ -> 0x00, 128), get(opts, "exchange_mode", 2), 0)
            smm_check(ccall((:smm_bgp_create, LIBSMM_B200), Cint, (Ref{SmmBgpConfig}, Ref{Ptr{Cvoid}}), cfg, h))
        end
        # the chain objects the rest of the package (summary, history, plotting) reads; probs_acc is the Uacc stream
        chains = BGPChain[BGPChain(i, n, m = m, sig = sigma0[i], upd = get(opts, "sigma_update_steps", 10),
                                   upd_by = get(opts, "sigma_adjust_by", 0.01), smpl_iters = get(opts, "smpl_iters", 1000),
                                   min_improve = minimp[i], acc_tuner = tuners[i], batch_size = get(opts, "batch_size", np)) for i in 1:N]
        for c in chains
            smm_check(ccall((:smm_stream_acc_uniforms, LIBSMM_B200), Cint, (UInt64, UInt32, Int32, Int32, Ptr{Cdouble}),
                            UInt64(get(opts, "seed", 20261017)), c.id - 1, 1, n, c.probs_acc))
        end
        this = new(m, opts, 0, chains, Animation(), -, h[], 0)
        finalizer(a -> (a.handle == C_NULL || ccall((:smm_bgp_destroy, LIBSMM_B200), Cvoid, (Ptr{Cvoid},), a.handle); a.handle = C_NULL), this)
        return this
    end
end

"copy iterations synced+1..algo.i of the device trace into the BGPChain objects"
function materialize!(algo::MAlgoBGPB200)
    lo, hi = algo.synced + 1, algo.i
    hi >= lo || return algo
    n = hi - lo + 1; N = algo.opts["N"]
    np = length(algo.m.params_to_sample); nm = length(algo.m.moments)
    value = zeros(N, n); prob = zeros(N, n); curr = zeros(N, n); best = zeros(N, n)       # C row-major [n][N] == Julia (N, n)
    pars = zeros(np, N, n); moms = zeros(nm, N, n)
    acc = zeros(UInt8, N, n); status = zeros(Int32, N, n); exch = zeros(Int32, N, n); bid = zeros(Int32, N, n)
    GC.@preserve value prob curr best pars moms acc status exch bid begin
        v = SmmTraceView(pointer(value), pointer(prob), pointer(curr), pointer(best), pointer(pars), pointer(moms),
                         pointer(acc), pointer(status), pointer(exch), pointer(bid))
        smm_check(ccall((:smm_bgp_read_trace, LIBSMM_B200), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{SmmTraceView}), algo.handle, lo, hi, v))
    end
    sig = zeros(N); ar = zeros(N)
    smm_check(ccall((:smm_bgp_read_chain_state, LIBSMM_B200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), algo.handle, sig, ar))
    pnames = collect(keys(algo.m.params_to_sample)); mnames = collect(keys(algo.m.moments))
    for (ic, c) in enumerate(algo.chains)
        for t in 1:n
            it = lo + t - 1
            ev = Eval(algo.m, OrderedDict(zip(pnames, pars[:, ic, t])))
            ev.value = value[ic, t]; ev.prob = prob[ic, t]; ev.status = status[ic, t]; ev.accepted = acc[ic, t] != 0
            status[ic, t] >= 0 && setMoments!(ev, mnames, moms[:, ic, t])
            c.evals[it] = ev
            c.accepted[it] = ev.accepted; c.exchanged[it] = exch[ic, t]
            c.curr_val[it] = curr[ic, t]; c.best_val[it] = best[ic, t]; c.best_id[it] = bid[ic, t]
        end
        c.iter = hi; c.sigma = sig[ic]; c.accept_rate = ar[ic]
    end
    algo.synced = hi
    return algo
end

"computeNextIteration!(algo) (AlgoBGP.jl:589-640): the whole iteration -- proposals, objective, accept/reject, exchange -- on the device"
function computeNextIteration!(algo::MAlgoBGPB200)
    smm_check(ccall((:smm_bgp_step, LIBSMM_B200), Cint, (Ptr{Cvoid}, Int32, Ptr{Cfloat}), algo.handle, 1, C_NULL))
    materialize!(algo)       # keeps `run!`'s per-iteration contract (algo.i already set by run!, AlgoAbstract.jl:42)
end

"run!(algo): all iterations in one call (one persistent kernel launch per 128 iterations), then one trace read-back"
function run!(algo::MAlgoBGPB200)
    t0 = time()
    n = algo["maxiter"] - algo.i
    smm_check(ccall((:smm_bgp_step, LIBSMM_B200), Cint, (Ptr{Cvoid}, Int32, Ptr{Cfloat}), algo.handle, n, C_NULL))
    algo.i = algo["maxiter"]
    materialize!(algo)
    algo.opts["time"] = round((time() - t0) / 60, digits = 1)
    haskey(algo.opts, "filename") && save(algo, algo.opts["filename"])
    return algo
end

"""
    evaluateObjectiveBatch(m::MProb, plist; noseed=false, rep0=0, device=0)

Many `evaluateObjective(m, p)` calls (mprob.jl:175-205) as ONE `smm_bgp_eval_batch` launch: what the loops of
`doSlices` / `optSlices` (slices.jl:114-290) and `FD_gradient` / `getSigma` (econometrics.jl:29-145) need.
`plist` is a vector of parameter dicts; returns a vector of `Eval`s (status -2 and no moments where the objective
failed, exactly like the CPU path).  With `noseed=true` entry `b` draws its own shocks, indexed by `(b, rep0 + b)`.
"""
function evaluateObjectiveBatch(m::MProb, plist::Vector; noseed::Bool = false, rep0::Integer = 0, device::Integer = 0)
    haskey(SMM_OBJ, m.objfunc) || error("$(m.objfunc) has no device simulator (no CPU fallback in the B200 path)")
    pnames = collect(keys(m.params_to_sample)); mnames = collect(keys(m.moments))
    np = length(pnames); nm = length(mnames); B = length(plist)
    # a one-chain, one-iteration handle carries the problem definition (bounds, data moments, weights, simulator)
    algo = MAlgoBGPB200(m, Dict("N" => 1, "maxiter" => 1, "maxtemp" => 1, "device" => device, "exchange_mode" => 0))
    P = Float64[plist[b][pnames[k]] for k in 1:np, b in 1:B]            # column b = entry b  (C row-major [B][P])
    value = zeros(B); moms = zeros(nm, B); status = zeros(Int32, B)
    smm_check(ccall((:smm_bgp_eval_batch, LIBSMM_B200), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Int32, Int32, UInt32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                    algo.handle, P, B, noseed ? 1 : 0, rep0, value, moms, status))
    finalize(algo)
    evs = Eval[]
    for b in 1:B
        ev = Eval(m, OrderedDict(zip(pnames, P[:, b])))
        ev.value = value[b]; ev.status = status[b]
        status[b] >= 0 && setMoments!(ev, mnames, moms[:, b])
        noseed && (ev.options[:noseed] = true)
        push!(evs, ev)
    end
    return evs
end
