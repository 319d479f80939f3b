# test_AlgoBGPB200.jl -- the reference's own BGP tests (test/test_algoBGP.jl, test/test_objfunc.jl, test/test_slices.jl)
# restated for the B200 backend: drop into SMM.jl's test/ next to them and `include` it from runtests.jl on a machine
# with a B200 and libsmm_b200.so on the loader path (ENV["SMM_B200_LIB"]).  Same assertions as upstream where upstream
# has them; in addition the CPU twins of the device objectives (ObjB200.jl) are compared with the device.
#
# NOT EXECUTED in the build environment (no Julia in the image); tests/test_gpu_api.py runs the same assertions through
# the Python mirror of this surface.

@testset "AlgoBGP on B200" begin

	pb   = OrderedDict("p1" => [0.2, -3, 3], "p2" => [-0.2, -20, 20])            # Examples.jl:387-388
	moms = DataFrame(name = ["mu1", "mu2"], value = [-1.0, 10.0], weight = ones(2))
	mprob = MProb()
	addSampledParam!(mprob, pb)
	addMoment!(mprob, moms)
	addEvalFunc!(mprob, SMM.objfunc_norm)
	opts = Dict("N" => 3, "maxiter" => 20, "maxtemp" => 5, "smpl_iters" => 1000, "parallel" => false,
	            "min_improve" => [0.0 for i in 1:3], "acc_tuners" => [20; 2; 1.0], "backend" => :b200)

	@testset "Constructor" begin                                                    # test_algoBGP.jl:10-24
		MA = SMM.bgp_algorithm(mprob, opts)
		@test isa(MA, MAlgo) == true
		@test isa(MA, SMM.MAlgoBGPB200) == true
		@test isa(MA.m, MProb) == true
		@test MA.i == 0
		@test length(MA.chains) == 3
		for ix = 1:length(MA.chains)
			@test isa(MA.chains[ix], SMM.BGPChain)
		end
		@test isa(SMM.bgp_algorithm(mprob, delete!(copy(opts), "backend")), MAlgoBGP)   # default backend: the stock algorithm
	end

	@testset "serialNormal() shape runs" begin                                      # test_algoBGP.jl:26-34
		o = SMM.bgp_algorithm(mprob, opts)
		SMM.run!(o)
		h = SMM.history(o.chains[1])
		@test isa(h, DataFrame)
		@test nrow(h) == 20
		@test ncol(h) == 9
		@test o.i == 20
		@test SMM.history(o, 1)[!, :value] == h[!, :value]                        # straight from the SoA trace, no Eval objects
		s = SMM.summary(o)
		@test nrow(s) == 3 && all(s[!, :acc_rate] .>= 0)
	end

	@testset "Can we recover the mean of a Normal?" begin                           # test_algoBGP.jl:57-121 (tolerance 1.0)
		o = SMM.bgp_algorithm(mprob, merge(opts, Dict("maxiter" => 200)))
		SMM.run!(o)
		med = SMM.median(o)[1]                                                      # reduced on the device
		@test abs(med[:p1] - (-1.0)) < 1.0
		@test abs(med[:p2] - 10.0) < 1.0
		@test med == SMM.median(o.chains[1])                                        # same as from the chain object
	end

	@testset "accept/reject bookkeeping is upstream's" begin                        # test_BGPchain.jl:95-144 on a device run
		o = SMM.bgp_algorithm(mprob, opts); SMM.run!(o)
		for c in o.chains
			@test c.accepted[1] == true && c.evals[1].prob == 1.0                   # iteration 1 is always accepted (AlgoBGP.jl:326-331)
			for it in 2:20
				if c.exchanged[it] == 0
					@test c.accepted[it] == (c.evals[it].prob > c.probs_acc[it])    # strict >, pre-drawn uniform (:363-367)
					@test c.curr_val[it] == (c.accepted[it] ? c.evals[it].value : c.curr_val[it-1])
				end
				@test c.best_val[it] <= c.best_val[it-1]
			end
		end
	end

	@testset "device objective == Julia CPU twin" begin                             # the three-times-identical contract (SURVEY.md 0.3)
		m2 = MProb()
		addSampledParam!(m2, OrderedDict("p$k" => [0.2 * (-1.0)^k, -3, 3] for k in 1:4))
		addMoment!(m2, DataFrame(name = ["m$k" for k in 1:8], value = vcat([-1.0, 1.0, 0.5, -0.5], ones(4)), weight = ones(8)))
		addEvalFunc!(m2, SMM.objfunc_norm_mv)
		p = OrderedDict(:p1 => 0.3, :p2 => -0.4, :p3 => 1.1, :p4 => -2.0)            # parameter names are Symbols (mprob.jl:81-84)
		cpu = SMM.evaluateObjective(m2, p)                                          # ObjB200.jl on the host
		gpu = SMM.evaluateObjectiveBatch(m2, [p])[1]                                # objective_kernel on the device
		@test cpu.status == gpu.status == 1
		@test isapprox(cpu.value, gpu.value, rtol = 1e-9)
		for k in keys(cpu.simMoments)
			@test isapprox(cpu.simMoments[k], gpu.simMoments[k], rtol = 1e-9, atol = 1e-12)
		end
	end

	@testset "failing objective" begin                                              # test_slices.jl:39-60, mprob.jl:181-186
		mf = MProb(); addSampledParam!(mf, pb); addMoment!(mf, moms); addEvalFunc!(mf, SMM.Testobj_fails)
		ev = SMM.evaluateObjectiveBatch(mf, [OrderedDict(:p1 => 0.0, :p2 => 0.0)])[1]
		@test ev.status == -2 && ev.value == -1.0 && length(ev.simMoments) == 0
	end

	@testset "save / read / continue" begin                                         # test_AlgoAbstract.jl:36-62
		fn = tempname() * ".jld2"
		o = SMM.bgp_algorithm(mprob, merge(opts, Dict("maxiter" => 10, "filename" => fn, "save_frequency" => 5)))
		SMM.run!(o)
		o2 = SMM.readMalgoB200(fn)
		@test o2.i == o.i == 10
		for (a, b) in zip(o.chains, o2.chains)
			@test a.best_val == b.best_val && a.accepted == b.accepted && a.exchanged == b.exchanged && a.sigma == b.sigma
		end
	end
end
