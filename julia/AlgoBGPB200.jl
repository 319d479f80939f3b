# AlgoBGPB200.jl -- the reference-side binding of libsmm_b200.so.
#
# Drop julia/{SMMStreams.jl, smm_stream_tables.jl, ObjB200.jl, AlgoBGPB200.jl} into SMM.jl's src/mopt/ and add, after
# `include("mopt/AlgoBGP.jl")` in src/SMM.jl:
#
#     include("mopt/SMMStreams.jl"); include("mopt/ObjB200.jl"); include("mopt/AlgoBGPB200.jl")
#
# The reference's algorithm plug-in point is a subtype of MAlgo that implements computeNextIteration!
# (README.md:105-107, AlgoAbstract.jl:8,45).  User code keeps calling addSampledParam! / addMoment! / addEvalFunc! /
# run! / summary / history; the backend is chosen by ONE opts key:
#
#     opts["backend"] = :b200
#     MA = SMM.bgp_algorithm(mprob, opts)          # MAlgoBGPB200 if opts["backend"] == :b200, else the stock MAlgoBGP
#
# (or, with the three-line hook of INTEGRATION.md section 2 inside MAlgoBGP's constructor, literally `MAlgoBGP(m, opts)`).
#
# Host code stays Julia; the only foreign calls are the `ccall`s below -- no CUDA.jl: the library owns device, stream,
# memory and (world_size > 1) its NCCL communicator.  There is no CPU fallback: an objective without a device
# simulator raises.  NOT EXECUTED in the build environment (no Julia in the image): written against
# include/smm_b200.h, whose struct layout tests/test_abi_and_host.py pins with a C compiler, and checked field by
# field against it by tests/test_julia_files.py.

const LIBSMM_B200 = get(ENV, "SMM_B200_LIB", "libsmm_b200.so")
const SMM_ABI_VERSION = Int32(1)
const SMM_E_UNSUPPORTED_SHAPE = -4
const SMM_NCCL_ID_BYTES = 128

# objective ids (include/smm_b200.h SMM_OBJ_*): the Julia function registered with addEvalFunc! selects the device
# simulator that computes the same thing (ObjB200.jl holds the CPU twins of ids 0, 2, 3 on the shared streams)
const SMM_OBJ = Dict{Function,Int32}(
    objfunc_norm => 0, objfunc_norm_b200 => 0,      # ObjExamples.jl:59-116
    objfunc_norm_slow => 1,                         # ObjExamples.jl:124-184
    objfunc_norm_mv => 2,                           # means + variances (C2 / C3)
    objfunc_panel => 3,                             # dynamic panel (C4)
    Testobj_fails => 4)                             # ObjExamples.jl:27-32: every evaluation throws -> status -2

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

"the job's NCCL id: call on rank 0, send the 128 bytes to the other ranks (Distributed / MPI), pass as opts[\"nccl_id\"]"
function smm_nccl_unique_id()
    id = zeros(UInt8, SMM_NCCL_ID_BYTES)
    smm_check(ccall((:smm_nccl_unique_id, LIBSMM_B200), Cint, (Ptr{UInt8},), id))
    return id
end

"SoA copy of the device trace for iterations lo..hi of this rank's chains: Julia arrays are (L, n) = C's [n][L]"
struct B200Trace
    lo::Int
    value::Matrix{Float64}; prob::Matrix{Float64}; curr_val::Matrix{Float64}; best_val::Matrix{Float64}
    params::Array{Float64,3}; sim_moments::Array{Float64,3}            # (P, L, n), (M, L, n)
    accepted::Matrix{UInt8}; status::Matrix{Int32}; exchanged::Matrix{Int32}; best_id::Matrix{Int32}
end

mutable struct MAlgoBGPB200 <: MAlgo
    m::MProb
    opts::Dict
    i::Int
    chains::Array{BGPChain}      # this rank's chains: ids rank + 1, rank + 1 + world, ... (round robin; all of them when world_size == 1)
    anim::Plots.Animation
    dist_fun::Function
    handle::Ptr{Cvoid}
    synced::Int                  # iterations already copied into `chains`
    exchange_mode::Int
    world::Int                   # local chain i (1-based) is global chain (i - 1) * world + rank + 1

    function MAlgoBGPB200(m::MProb, opts::Dict)
        haskey(opts, "dist_fun") && opts["dist_fun"] !== (-) && error("dist_fun: only the default `-` runs on the device")
        haskey(SMM_OBJ, m.objfunc) || error("$(m.objfunc) has no device simulator (no CPU fallback in the B200 path)")
        collect(keys(m.initial_value)) == collect(keys(m.params_to_sample)) ||
            error("every parameter must be sampled (proposal broadcasts params against lb/ub, AlgoBGP.jl:430-436)")
        N = opts["N"]; n = opts["maxiter"]
        np = length(m.params_to_sample)
        world = get(opts, "world_size", 1); rank = get(opts, "rank", 0)
        N % world == 0 || error("N must be a multiple of world_size")
        L = N ÷ world
        gid = [(i - 1) * world + rank + 1 for i in 1:L]          # global ids of this rank's chains
        temps = N > 1 ? collect(range(1.0, stop = Float64(get(opts, "maxtemp", 1.0)), length = N)) : [1.0]   # AlgoBGP.jl:508
        sigma0 = get(opts, "sigma", 0.05) .* temps
        tuners = Float64.(get(opts, "acc_tuners", [2.0 for j in 1:N]))
        minimp = Float64.(get(opts, "min_improve", [0.5 for j in 1:N]))
        lb = Float64[v[:lb] for (k, v) in m.params_to_sample]
        ub = Float64[v[:ub] for (k, v) in m.params_to_sample]
        init = Float64[m.initial_value[k] for k in keys(m.params_to_sample)]
        dm = Float64[v[:value] for (k, v) in m.moments]
        dw = Float64[v[:weight] for (k, v) in m.moments]
        oo = m.objfunc_opts
        obj = SMM_OBJ[m.objfunc]
        idb = get(opts, "nccl_id", zeros(UInt8, SMM_NCCL_ID_BYTES))
        world == 1 || length(idb) == SMM_NCCL_ID_BYTES || error("opts[\"nccl_id\"]: 128 bytes from smm_nccl_unique_id() on rank 0")
        # exchange_mode: the fastest mode the shape allows (2 = barrier-free persistent kernel on one GPU, 3 = the same
        # with the flag-in-data hand-over on several; 0 = one launch per iteration for the panel objective and for more
        # than 32 parameters); a persistent mode that does not fit the shape falls back to 0 unless the user asked for
        # it explicitly (same rule as smm_jl_b200/api.py)
        explicit = haskey(opts, "exchange_mode")
        mode = explicit ? opts["exchange_mode"] : ((obj == 3 || np > 32) ? 0 : (world > 1 ? 3 : 2))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve lb ub init dm dw sigma0 tuners minimp begin
            mkcfg(md) = SmmBgpConfig(SMM_ABI_VERSION, np, length(dm), pointer(lb), pointer(ub), pointer(init), pointer(dm), pointer(dw),
                obj, get(oo, :n_sim, 10000), UInt64(get(oo, :seed, 1234)), Int32(get(oo, :noseed, false)),
                Float64(get(oo, :slow_seconds, 0.1)), get(oo, :panel_T, 0), get(oo, :panel_N, 0), get(oo, :panel_K, 0),
                N, n, pointer(sigma0), pointer(tuners), pointer(minimp),
                get(opts, "sigma_update_steps", 10), get(opts, "sigma_adjust_by", 0.01), get(opts, "smpl_iters", 1000),
                get(opts, "batch_size", np), UInt64(get(opts, "seed", 20261017)),
                get(opts, "device", 0), world, rank, ntuple(i -> idb[i], SMM_NCCL_ID_BYTES), md, get(opts, "n_split", 0))
            rc = ccall((:smm_bgp_create, LIBSMM_B200), Cint, (Ref{SmmBgpConfig}, Ref{Ptr{Cvoid}}), mkcfg(mode), h)
            if rc == SMM_E_UNSUPPORTED_SHAPE && !explicit && mode != 0
                mode = 0
                rc = ccall((:smm_bgp_create, LIBSMM_B200), Cint, (Ref{SmmBgpConfig}, Ref{Ptr{Cvoid}}), mkcfg(mode), h)
            end
            smm_check(rc)
        end
        # the chain objects the rest of the package (summary, history, plotting) reads; probs_acc is the Uacc stream
        chains = BGPChain[BGPChain(gid[i], n, m = m, sig = sigma0[gid[i]], upd = get(opts, "sigma_update_steps", 10),
                                   upd_by = get(opts, "sigma_adjust_by", 0.01), smpl_iters = get(opts, "smpl_iters", 1000),
                                   min_improve = minimp[gid[i]], acc_tuner = tuners[gid[i]],
                                   batch_size = get(opts, "batch_size", np)) for i in 1:L]
        for c in chains
            smm_check(ccall((:smm_stream_acc_uniforms, LIBSMM_B200), Cint, (UInt64, UInt32, Int32, Int32, Ptr{Cdouble}),
                            UInt64(get(opts, "seed", 20261017)), c.id - 1, 1, n, c.probs_acc))
        end
        this = new(m, opts, 0, chains, Animation(), -, h[], 0, mode, world)
        finalizer(a -> (a.handle == C_NULL || ccall((:smm_bgp_destroy, LIBSMM_B200), Cvoid, (Ptr{Cvoid},), a.handle); a.handle = C_NULL), this)
        return this
    end
end

"`MAlgoBGP(m, opts)` or its B200 twin, chosen by `opts[\"backend\"]` (:cpu by default)"
bgp_algorithm(m::MProb, opts::Dict) = get(opts, "backend", :cpu) == :b200 ? MAlgoBGPB200(m, opts) : MAlgoBGP(m, opts)

"read iterations lo..hi of this rank's chains from the device (one D2H per column)"
function read_trace(algo::MAlgoBGPB200, lo::Int, hi::Int)
    n = hi - lo + 1; L = length(algo.chains)
    np = length(algo.m.params_to_sample); nm = length(algo.m.moments)
    t = B200Trace(lo, zeros(L, n), zeros(L, n), zeros(L, n), zeros(L, n), zeros(np, L, n), zeros(nm, L, n),
                  zeros(UInt8, L, n), zeros(Int32, L, n), zeros(Int32, L, n), zeros(Int32, L, n))
    GC.@preserve t begin
        v = SmmTraceView(pointer(t.value), pointer(t.prob), pointer(t.curr_val), pointer(t.best_val), pointer(t.params),
                         pointer(t.sim_moments), pointer(t.accepted), pointer(t.status), pointer(t.exchanged), pointer(t.best_id))
        smm_check(ccall((:smm_bgp_read_trace, LIBSMM_B200), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{SmmTraceView}), algo.handle, lo, hi, v))
    end
    return t
end

"""
    materialize!(algo; evals = N * maxiter <= 200_000)

Copy iterations synced+1..algo.i of the device trace into the BGPChain vectors (`accepted`, `exchanged`, `curr_val`,
`best_val`, `best_id`, `sigma`, `accept_rate`).  `Eval` objects -- four dicts each, the reference's allocation hot
spot (SURVEY.md section 7, "trace memory") -- are built only when `evals` is true (default: small runs) or later, for
the slots someone asks for, by `eval_at(algo, chain, iter)`.
"""
function materialize!(algo::MAlgoBGPB200; evals::Bool = get(algo.opts, "materialize_evals", algo.opts["N"] * algo.opts["maxiter"] <= 200_000))
    lo, hi = algo.synced + 1, algo.i
    hi >= lo || return algo
    t = read_trace(algo, lo, hi)
    L = length(algo.chains)
    sig = zeros(L); ar = zeros(L)
    smm_check(ccall((:smm_bgp_read_chain_state, LIBSMM_B200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), algo.handle, sig, ar))
    for (ic, c) in enumerate(algo.chains)
        for k in 1:(hi - lo + 1)
            it = lo + k - 1
            c.accepted[it] = t.accepted[ic, k] != 0; c.exchanged[it] = t.exchanged[ic, k]
            c.curr_val[it] = t.curr_val[ic, k]; c.best_val[it] = t.best_val[ic, k]; c.best_id[it] = t.best_id[ic, k]
            evals && (c.evals[it] = _make_eval(algo, t, ic, k))
        end
        c.iter = hi; c.sigma = sig[ic]; c.accept_rate = ar[ic]
    end
    algo.synced = hi
    return algo
end

function _make_eval(algo::MAlgoBGPB200, t::B200Trace, ic::Int, k::Int)
    pnames = collect(keys(algo.m.params_to_sample)); mnames = collect(keys(algo.m.moments))
    ev = Eval(algo.m, OrderedDict(zip(pnames, t.params[:, ic, k])))
    ev.value = t.value[ic, k]; ev.prob = t.prob[ic, k]; ev.status = t.status[ic, k]; ev.accepted = t.accepted[ic, k] != 0
    t.status[ic, k] >= 0 && setMoments!(ev, Dict(zip(mnames, t.sim_moments[:, ic, k])))
    return ev
end

"the Eval of (local chain index, iteration), built on demand from the device trace (lazy `c.evals[iter]`)"
function eval_at(algo::MAlgoBGPB200, chain::Int, iter::Int)
    c = algo.chains[chain]
    isassigned(c.evals, iter) && iter <= algo.synced && c.evals[iter].status != -1 && return c.evals[iter]
    c.evals[iter] = _make_eval(algo, read_trace(algo, iter, iter), chain, 1)
    return c.evals[iter]
end

"history(algo, chain): the DataFrame of history(c::BGPChain) (AlgoBGP.jl:138-160) straight from the SoA trace -- no Eval objects"
function history(algo::MAlgoBGPB200, chain::Int)
    t = read_trace(algo, 1, algo.i)
    d = DataFrame(iter = 1:algo.i, value = t.value[chain, :], accepted = t.accepted[chain, :] .!= 0, curr_val = t.curr_val[chain, :],
                  best_val = t.best_val[chain, :], prob = t.prob[chain, :], exchanged = Int.(t.exchanged[chain, :]))
    for (j, k) in enumerate(keys(algo.m.params_to_sample))
        d[!, k] = t.params[j, chain, :]
    end
    return d
end

"""
    accepted_stats(algo, probs) -> (count[L], mean[P, L], quantiles[length(probs), P, L])

Accepted-only statistics of every local chain reduced on the device (`smm_bgp_accepted_stats`): what `mean(c)`,
`median(c)`, `CI(c)` (AlgoBGP.jl:174-188) compute from `params(c)`, without reading the trace back.
"""
function accepted_stats(algo::MAlgoBGPB200, probs::Vector{Float64})
    L = length(algo.chains); np = length(algo.m.params_to_sample); nq = length(probs)
    cnt = zeros(Int64, L); mu = zeros(np, L); q = zeros(max(nq, 1), np, L)
    smm_check(ccall((:smm_bgp_accepted_stats, LIBSMM_B200), Cint,
                    (Ptr{Cvoid}, Int32, Int32, Ptr{Cdouble}, Int32, Ptr{Int64}, Ptr{Cdouble}, Ptr{Cdouble}),
                    algo.handle, 1, algo.i, probs, nq, cnt, mu, q))
    return cnt, mu, q[1:nq, :, :]
end
_named(algo::MAlgoBGPB200, v) = Dict(zip(collect(keys(algo.m.params_to_sample)), v))
mean(algo::MAlgoBGPB200) = [_named(algo, accepted_stats(algo, Float64[])[2][:, c]) for c in 1:length(algo.chains)]
median(algo::MAlgoBGPB200) = (q = accepted_stats(algo, [0.5])[3]; [_named(algo, q[1, :, c]) for c in 1:length(algo.chains)])
function CI(algo::MAlgoBGPB200; level = 0.95)
    q = accepted_stats(algo, [(1 - level) / 2, 1 - (1 - level) / 2])[3]
    return [_named(algo, [q[:, k, c] for k in 1:size(q, 2)]) for c in 1:length(algo.chains)]
end

"summary(algo) (AlgoBGP.jl:541-550) from device-side reductions (`smm_bgp_chain_summary`): no trace read-back"
function summary(algo::MAlgoBGPB200)
    L = length(algo.chains)
    nx = zeros(Int64, L); mw = zeros(Int32, L); bv = zeros(L); sig = zeros(L); ar = zeros(L)
    smm_check(ccall((:smm_bgp_chain_summary, LIBSMM_B200), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}, Ptr{Cdouble}), algo.handle, nx, mw, bv))
    smm_check(ccall((:smm_bgp_read_chain_state, LIBSMM_B200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), algo.handle, sig, ar))
    return DataFrame(id = [c.id for c in algo.chains], acc_rate = ar, perc_exchanged = 100 .* nx ./ algo["maxiter"],
                     exchanged_most_with = Int.(mw), best_val = bv)
end

"computeNextIteration!(algo) (AlgoBGP.jl:589-640): the whole iteration -- proposals, objective, accept/reject, exchange -- on the device"
function computeNextIteration!(algo::MAlgoBGPB200)
    smm_check(ccall((:smm_bgp_step, LIBSMM_B200), Cint, (Ptr{Cvoid}, Int32, Ptr{Cfloat}), algo.handle, 1, C_NULL))
    materialize!(algo)       # keeps `run!`'s per-iteration contract (algo.i already set by run!, AlgoAbstract.jl:42)
end

"""
    run!(algo::MAlgoBGPB200)

run!(algo) (AlgoAbstract.jl:27-76): every remaining iteration in one `smm_bgp_step` call (one persistent kernel
launch per 128 iterations) -- or in chunks of `opts["save_frequency"]` iterations with a `save` after each chunk when
`opts["filename"]` is set, as upstream (AlgoAbstract.jl:48-73) -- then one trace read-back.
"""
function run!(algo::MAlgoBGPB200)
    t0 = time()
    sf = get(algo.opts, "save_frequency", 0); fn = get(algo.opts, "filename", "")
    chunk = (sf > 0 && fn != "") ? sf : algo["maxiter"]
    while algo.i < algo["maxiter"]
        n = min(chunk, algo["maxiter"] - algo.i)
        smm_check(ccall((:smm_bgp_step, LIBSMM_B200), Cint, (Ptr{Cvoid}, Int32, Ptr{Cfloat}), algo.handle, n, C_NULL))
        algo.i += n
        if sf > 0 && fn != "" && algo.i % sf == 0
            save(algo, fn)
        end
    end
    materialize!(algo)
    algo.opts["time"] = round((time() - t0) / 60, digits = 1)
    fn != "" && save(algo, fn)
    return algo
end

"device checkpoint (smm_bgp_export_state): sigma, accept counters, last-accepted records, the trace; no RNG state exists"
function export_state(algo::MAlgoBGPB200)
    nb = ccall((:smm_bgp_state_bytes, LIBSMM_B200), Int64, (Ptr{Cvoid},), algo.handle)
    buf = zeros(UInt8, nb)
    smm_check(ccall((:smm_bgp_export_state, LIBSMM_B200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), algo.handle, buf, nb))
    return buf
end

"save(algo, filename) (AlgoAbstract.jl:83-88): problem + opts + the device checkpoint; one file per rank when world_size > 1"
function save(algo::MAlgoBGPB200, filename::AbstractString)
    world = get(algo.opts, "world_size", 1)
    fn = world == 1 ? filename : "$(filename).rank$(get(algo.opts, "rank", 0))of$(world)"
    opts = filter(kv -> kv.first != "nccl_id", algo.opts)
    JLD2.jldsave(fn; m = algo.m, opts = opts, i = algo.i, state = algo.i > 0 ? export_state(algo) : UInt8[])
end

"readMalgo for a B200 checkpoint (AlgoAbstract.jl:95-102); `placement` = Dict of device / world_size / rank / nccl_id of the new job"
function readMalgoB200(filename::AbstractString; placement::Dict = Dict())
    world = get(placement, "world_size", 1)
    fn = world == 1 ? filename : "$(filename).rank$(get(placement, "rank", 0))of$(world)"
    d = JLD2.load(fn)
    algo = MAlgoBGPB200(d["m"], merge(d["opts"], placement))
    if d["i"] > 0
        st = d["state"]
        smm_check(ccall((:smm_bgp_import_state, LIBSMM_B200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), algo.handle, st, length(st)))
        algo.i = d["i"]
        materialize!(algo)
    end
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
        status[b] >= 0 && setMoments!(ev, Dict(zip(mnames, moms[:, b])))
        noseed && (ev.options[:noseed] = true)
        push!(evs, ev)
    end
    return evs
end
