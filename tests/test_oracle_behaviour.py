"""The reference's own portable behavioural tests, run against the oracle:
test/test_BGPchain.jl:95-144 (accept/reject), test/test_objfunc.jl:22-29 (objfunc_norm),
test/test_algoBGP.jl:30-38 (history shape), :57-121 and :123-193 (recover the mean),
plus the README's serialNormal(2,200) run (README.md:41-53)."""
import numpy as np
import pytest

from smm_jl_b200 import configs
from smm_jl_b200._abi import BGPConfig, SMM_OBJ_FAILS, SMM_OBJ_NORM, SMM_E_NEGATIVE_OBJECTIVE, SMM_E_UNSUPPORTED_SHAPE


def test_initial_period_accepts_everything(oracle):
    # test_BGPchain.jl:95-111
    prob, acc, st = oracle.accept_reject(1, 0.0, 123.0, -2, 2.0, 0.999)
    assert prob == 1.0 and acc and st == 1


def test_accept_reject_rules(oracle):
    # test_BGPchain.jl:113-144: old value 1.5; bad = 2.0 -> prob < 1, accepted iff prob > probs_acc[iter]
    prob, acc, st = oracle.accept_reject(2, 1.5, 2.0, 1, 2.0, 0.3)
    assert prob == pytest.approx(np.exp(2.0 * (1.5 - 2.0))) and prob < 1 and st == 1
    assert acc == (prob > 0.3)
    prob2, acc2, _ = oracle.accept_reject(2, 1.5, 2.0, 1, 2.0, 0.9)
    assert prob2 == prob and not acc2
    # strict inequality: prob == u rejects
    _, acc3, _ = oracle.accept_reject(2, 1.5, 2.0, 1, 2.0, prob)
    assert not acc3
    # good = 1.0 -> prob == 1.0, accepted whatever the uniform
    prob, acc, st = oracle.accept_reject(2, 2.0, 1.0, 1, 2.0, 0.999999)
    assert prob == 1.0 and acc and st == 1


def test_failed_and_nonfinite_evaluations(oracle):
    # AlgoBGP.jl:336-338: status < 0 -> prob 0, rejected, status kept
    assert oracle.accept_reject(2, 1.0, -1.0, -2, 2.0, 0.0) == (0.0, False, -2)
    # :350-353: non-finite prob (Inf - Inf) -> reject, status -1
    assert oracle.accept_reject(2, np.inf, np.inf, 1, 2.0, 0.0) == (0.0, False, -1)
    # :355-359: old value not finite -> accept with prob 1
    assert oracle.accept_reject(2, np.inf, 3.0, 1, 2.0, 0.99) == (1.0, True, 1)
    # new value +Inf: prob = exp(-Inf) = 0 -> finite -> rejected, status 1
    assert oracle.accept_reject(2, 1.0, np.inf, 1, 2.0, 0.5) == (0.0, False, 1)
    # :341 negative objective is a hard error
    with pytest.raises(oracle.OracleError) as e:
        oracle.accept_reject(2, 1.0, -0.5, 1, 2.0, 0.5)
    assert e.value.code == SMM_E_NEGATIVE_OBJECTIVE


def test_objfunc_norm_at_zero(oracle):
    # test_objfunc.jl:22-29: p = (0,0), data (0,0): simulated means within 0.1
    cfg = BGPConfig(lb=[-3, -3], ub=[3, 3], init=[0, 0], data_mom=[0, 0], data_w=[1, 1], n_chains=1, max_iter=1,
                    sigma0=[0.05], acc_tuner=[2.0], min_improve=[0.0], objective_id=SMM_OBJ_NORM)
    v, m, s = oracle.eval_batch(cfg, [[0.0, 0.0]])
    assert s[0] == 1 and np.all(np.abs(m[0]) < 0.1) and v[0] == pytest.approx(np.mean(m[0] ** 2))
    # common random numbers: sim = p + zbar, an affine function of the parameters
    v2, m2, _ = oracle.eval_batch(cfg, [[1.0, -2.0]])
    np.testing.assert_allclose(m2[0] - m[0], [1.0, -2.0], rtol=0, atol=1e-12)
    # noseed draws fresh shocks
    _, m3, _ = oracle.eval_batch(cfg, [[0.0, 0.0]], noseed=1, rep0=5)
    assert not np.allclose(m3, m)


def test_serial_normal_run_shape_and_first_iteration(oracle):
    # test_algoBGP.jl:30-38: serialNormal(2,20): history is 20 x 9 -> 20 rows, 7 + 2 param columns
    cfg = configs.c1_serial_normal(20)
    r = oracle.run(cfg, 20)
    tr = r.trace
    assert tr.value.shape == (20, 3) and tr.params.shape == (20, 3, 2)
    # iteration 1: the initial value, accepted with prob 1, best = current = value (AlgoBGP.jl:327-333, 224-227)
    np.testing.assert_array_equal(tr.params[0], np.tile([0.2, -0.2], (3, 1)))
    assert (tr.accepted[0] == 1).all() and (tr.prob[0] == 1.0).all() and (tr.status[0] == 1).all()
    np.testing.assert_array_equal(tr.best_val[0], tr.value[0])
    np.testing.assert_array_equal(tr.curr_val[0], tr.value[0])
    assert (tr.best_id[0] == 1).all() and (tr.exchanged[0] == 0).all()
    # iteration 2 differs from the initial value (test_BGPchain.jl:46-64)
    assert not np.array_equal(tr.params[1], tr.params[0])
    # all proposals inside the bounds
    assert (tr.params[..., 0] >= -3).all() and (tr.params[..., 0] <= 3).all()
    assert (tr.params[..., 1] >= -20).all() and (tr.params[..., 1] <= 20).all()


def test_bookkeeping_invariants(oracle):
    cfg = configs.c1_serial_normal(150)
    tr = oracle.run(cfg, 150).trace
    n = tr.n
    for c in range(3):
        # best_val is the running minimum of value (set_eval!, :233-240) and best_id points at it
        np.testing.assert_array_equal(tr.best_val[:, c], np.minimum.accumulate(tr.value[:, c]))
        for it in range(n):
            assert tr.value[tr.best_id[it, c] - 1, c] == tr.best_val[it, c]
        # curr_val follows accepted values
        for it in range(1, n):
            want = tr.value[it, c] if tr.accepted[it, c] else tr.curr_val[it - 1, c]
            assert tr.curr_val[it, c] == want
    # an exchange at iteration t stores partner ids symmetrically and swaps accepted records
    its, cs = np.nonzero(tr.exchanged)
    assert len(its) > 0
    for it, c in zip(its, cs):
        assert it >= 1 and tr.accepted[it, c] == 1
    # no exchange before iteration 2 (AlgoBGP.jl:637)
    assert (tr.exchanged[0] == 0).all()


def test_readme_run_recovers_the_mean(oracle):
    # README.md:41-53 / test_algoBGP.jl:57-121: after 200 iterations the median of chain 1 is within
    # tolerance of the truth; acceptance rates of the order of 0.1; best value ~1e-3
    cfg = configs.c1_serial_normal(200)
    r = oracle.run(cfg, 200)
    tr = r.trace
    acc = tr.accepted[:, 0] == 1
    med = np.median(tr.params[acc, 0, :], axis=0)
    assert abs(med[0] - (-1.0)) < 1.0 and abs(med[1] - 10.0) < 1.0
    assert tr.best_val[-1, 0] < 0.05
    assert (r.accept_rate > 0.02).all() and (r.accept_rate < 0.6).all()
    # sigma adapted every 10 iterations by 1% (AlgoBGP.jl:381-390): 20 updates (iterations 10..200)
    ratio = r.sigma / np.asarray(cfg.sigma0)
    k = np.log(ratio) / np.log(1.01)
    assert np.all(np.abs(ratio - 1) < 0.25)


def test_two_chain_recovery_with_batches(oracle):
    # test_algoBGP.jl:123-193: 2 chains, batch_size = 1, never adapt sigma, tol 0.7... (we use 1.0: other streams)
    cfg = BGPConfig(lb=[-3, -2], ub=[3, 2], init=[0.2, -0.2], data_mom=[-1.0, 1.0], data_w=[1, 1], n_chains=2,
                    max_iter=200, sigma0=0.05 * np.array([1.0, 5.0]), acc_tuner=[5.0, 1.0], min_improve=[0.0, 0.0],
                    objective_id=SMM_OBJ_NORM, sigma_update_steps=201, batch_size=1, seed_algo=1234)
    r = oracle.run(cfg, 200)
    tr = r.trace
    med = np.median(tr.params[tr.accepted[:, 0] == 1, 0, :], axis=0)
    assert abs(med[0] + 1.0) < 1.0 and abs(med[1] - 1.0) < 1.0
    np.testing.assert_array_equal(r.sigma, cfg.sigma0)      # never adapted


def test_failing_objective_run(oracle):
    # Testobj_fails (ObjExamples.jl:27-32) through run!: iteration 1 is force-accepted with status 1
    # (AlgoBGP.jl:327-331), later evaluations are rejected with prob 0 and status -2 (:336-338)
    cfg = configs.c1_serial_normal(10, objective_id=SMM_OBJ_FAILS)
    tr = oracle.run(cfg, 10).trace
    assert (tr.status[0] == 1).all() and (tr.value == -1.0).all()
    noex = tr.exchanged[1:] == 0
    assert (tr.status[1:][noex] == -2).all() and (tr.prob[1:][noex] == 0).all() and (tr.accepted[1:][noex] == 0).all()
    assert np.isnan(tr.sim_moments).all()


def test_unsupported_shapes_are_rejected(oracle):
    cfg = configs.mvnormal(4, 4, n_params=5, batch_size=2)   # test_chain2's shape: np % batch_size != 0
    with pytest.raises(oracle.OracleError) as e:
        oracle.run(cfg, 2)
    assert e.value.code == SMM_E_UNSUPPORTED_SHAPE
    cfg = configs.c1_serial_normal(5, data_mom=[1.0, 2.0, 3.0], data_w=[1, 1, 1])  # objfunc_norm needs P == M
    with pytest.raises(oracle.OracleError):
        oracle.run(cfg, 2)


def test_single_chain_has_no_exchange(oracle):
    cfg = configs.mvnormal(1, 30)
    tr = oracle.run(cfg, 30).trace
    assert (tr.exchanged == 0).all()


def test_threads_do_not_change_results(oracle):
    cfg = configs.mvnormal(8, 12)
    a, b = oracle.run(cfg, 12, n_threads=1).trace, oracle.run(cfg, 12, n_threads=4).trace
    for f in a.FLOAT_FIELDS + a.INT_FIELDS:
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True)
