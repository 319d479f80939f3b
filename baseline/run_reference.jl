# run_reference.jl -- time the REAL SMM.jl (floswald/SMM.jl) on the shapes bench.py uses, for anyone who has
# Julia (the build image does not: BASELINE.md section 2).  Prints objective-evaluations per second.
#
#     julia -p 8 baseline/run_reference.jl 256 100        # chains, iterations; workers mirror pmap
#
# objfunc_norm needs #params == #moments (ObjExamples.jl:77-78), so this runs the P = M = 8 "means only"
# variant of C2 (bench.py's norm_mv adds the 8 variances; same 80 000 normal draws per evaluation).
using Distributed
@everywhere using SMM
using DataStructures, DataFrames, Random

nchains = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 256
niter = length(ARGS) > 1 ? parse(Int, ARGS[2]) : 100
means = [-1.0, 1.0, 0.5, -0.5, 0.7, -0.7, 0.3, -0.3]
pb = OrderedDict("p$k" => [0.2 * (-1.0)^k, -3, 3] for k in 1:8)
moms = DataFrame(name = ["mu$k" for k in 1:8], value = means, weight = ones(8))
m = MProb(); addSampledParam!(m, pb); addMoment!(m, moms); addEvalFunc!(m, SMM.objfunc_norm)
opts = Dict("N" => nchains, "maxiter" => niter, "maxtemp" => 5, "sigma" => 0.05, "smpl_iters" => 100000,
            "sigma_update_steps" => 10, "sigma_adjust_by" => 0.01, "parallel" => nworkers() > 1,
            "min_improve" => zeros(nchains), "acc_tuners" => exp.(range(log(20.0), stop = log(1.0), length = nchains)))
MA = MAlgoBGP(m, opts)
t = @elapsed SMM.run!(MA)
println("SMM.jl reference: $(nchains) chains x $(niter) iterations in $(round(t, digits = 2)) s = ",
        round(nchains * niter / t, digits = 1), " objective-evals/s on $(nworkers()) worker(s)")
