"""Probabilistic registration on the device (gingr_b200/csrc/mcmc.cuh) against the oracle's literal restatement of
api/sampling/*: evaluators, the informed transition density (the library's SVD-free form vs the literal SVD form),
and whole Metropolis-Hastings chains step by step (device and oracle consume the same Philox stream, so generator
choice, proposals and accept / reject decisions must coincide).  Tolerances: log values 1e-9 relative, transition
densities 1e-6 relative, states as in test_update_gpu.py (1e-6)."""
import dataclasses

import numpy as np
import pytest

from test_update_gpu import _compare, _problem, _to_api_state

pytestmark = pytest.mark.gpu


def _setup(ctx, oracle, algo, M=150, N=180, r=16, seed=0, **cfgkw):
    from gingr_b200 import api
    m, target, tt = _problem(oracle, M, N, r, seed)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    if algo == "icp":
        reg = api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration(maxIterations=40, initialSigma=2.0, endSigma=0.5, **cfgkw))
        oalgo = oracle.IcpAlgorithm(oracle.IcpConfig(max_iterations=40, initial_sigma=2.0, end_sigma=0.5))
    else:
        reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(maxIterations=40, w=0.1))
        oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(max_iterations=40, w=0.1), literal=False)
    ost = oalgo.initialize(oracle.initial_state(m, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
    diag = float(np.linalg.norm(m.ref.max(0) - m.ref.min(0)))
    return reg, oalgo, ost, diag, (dm, dt)


def _settings(api, oracle, **kw):
    o = oracle.McmcSettings(**kw)
    g = api.ProbabilisticSettings(uncertainty=o.uncertainty, mode=o.mode, randomMixture=o.random_mixture,
                                  modelPointIds=o.model_ids, targetPointIds=o.target_ids, rotationSdev=o.rot_sdev,
                                  translationSdev=o.trans_sdev, shapeSteps=o.shape_steps)
    return g, o


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("subset", [False, True])
def test_evaluators_match_oracle(ctx, oracle, mode, subset):
    from gingr_b200 import api
    reg, oalgo, ost, diag, keep = _setup(ctx, oracle, "icp")
    for _ in range(2):
        ost = oracle.propose(oalgo, ost)
    kw = dict(uncertainty=1.7, mode=mode)
    if subset:
        kw.update(model_ids=np.arange(0, 150, 4), target_ids=np.arange(1, 180, 5))
    gs, os_ = _settings(api, oracle, **kw)
    reg.configureProbabilistic(gs)
    prior, dist = reg.logValue(_to_api_state(ost, api))
    oprior, odist = oracle.log_value(os_, ost)
    assert abs(prior - oprior) <= 1e-12 * abs(oprior)
    assert abs(dist - odist) <= 1e-9 * abs(odist)
    reg.close()


@pytest.mark.parametrize("algo", ["icp", "cpd"])
@pytest.mark.parametrize("step_length", [1.0, 0.5])
def test_informed_transition_density_matches_literal_svd_form(ctx, oracle, algo, step_length):
    from gingr_b200 import api
    reg, oalgo, ost, diag, keep = _setup(ctx, oracle, algo)
    ost = dataclasses.replace(ost, step_length=step_length)
    gs, os_ = _settings(api, oracle)
    reg.configureProbabilistic(gs)
    for _ in range(2):
        ost = oracle.propose(oalgo, ost)
    nxt = oracle.propose(oalgo, ost, True, seed=3)
    lit_fw = oracle.log_transition_informed(oalgo, ost, nxt)
    lit_bw = oracle.log_transition_informed(oalgo, nxt, ost)
    fw = reg.logTransitionProbability(_to_api_state(ost, api), _to_api_state(nxt, api))
    bw = reg.logTransitionProbability(_to_api_state(nxt, api), _to_api_state(ost, api))
    assert np.isfinite(lit_fw) and np.isfinite(lit_bw)
    assert abs(fw - lit_fw) <= 1e-6 * abs(lit_fw)
    assert abs(bw - lit_bw) <= 1e-6 * abs(lit_bw)
    reg.close()


def test_transition_density_is_minus_infinity_when_the_posterior_fails(ctx, oracle):
    from gingr_b200 import api
    reg, oalgo, ost, diag, keep = _setup(ctx, oracle, "cpd")
    gs, os_ = _settings(api, oracle)
    reg.configureProbabilistic(gs)
    bad = dataclasses.replace(oracle.propose(oalgo, ost), sigma2=float("nan"))
    assert oracle.log_transition_informed(oalgo, bad, ost) == -np.inf
    assert reg.logTransitionProbability(_to_api_state(bad, api), _to_api_state(ost, api)) == -np.inf
    reg.close()


@pytest.mark.parametrize("algo,rho,mode", [("icp", 0.5, 0), ("icp", 0.0, 2), ("cpd", 0.6, 0), ("icp", 1.0, 1)])
def test_mh_chain_matches_oracle_step_by_step(ctx, oracle, algo, rho, mode):
    from gingr_b200 import api
    reg, oalgo, ost, diag, keep = _setup(ctx, oracle, algo, M=120, N=140, r=12)
    gs, os_ = _settings(api, oracle, uncertainty=1.5, random_mixture=rho, mode=mode)
    reg.configureProbabilistic(gs)
    gst = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    _compare(gst, ost, diag)
    lp = oracle.log_value(os_, ost)
    seed, nsteps = 12345, 14
    leaves, n_acc = [], 0
    for k in range(nsteps):
        reg.mcmcChain(1, seed)
        ost, lp, info = oracle.mcmc_step(oalgo, os_, ost, lp, k, seed)
        values, counts = reg.mcmcStats()
        assert counts[0] == k + 1
        assert counts[1] == info["leaf"], (k, counts[1], info)
        # log values of the proposal, transition densities, decision
        lpp = info["lp_prop"]
        assert abs(values[2] - lpp[0]) <= 1e-6 * abs(lpp[0]) and abs(values[3] - lpp[1]) <= 1e-6 * abs(lpp[1])
        for got, want in ((values[4], info["t_fw"]), (values[5], info["t_bw"]), (values[9], info["fw"]), (values[10], info["bw"])):
            assert (got == want) if not np.isfinite(want) else abs(got - want) <= 1e-6 * abs(want), (k, got, want)
        assert abs(values[6] - info["u_accept"]) < 1e-15
        assert bool(counts[2]) == info["accept"], (k, values[7], info)
        n_acc += int(info["accept"])
        leaves.append(info["leaf"])
        _compare(reg.downloadState(), ost, diag)
        assert abs(values[0] - lp[0]) <= 1e-6 * abs(lp[0]) and abs(values[1] - lp[1]) <= 1e-6 * abs(lp[1])
    values, counts = reg.mcmcStats()
    assert counts[3] == n_acc and sum(counts[8:18]) == nsteps and sum(counts[18:28]) == n_acc
    if rho == 0.0:
        assert set(leaves) == {0}
    if rho == 1.0:
        assert 0 not in leaves
    best = reg.mcmcBest()
    assert np.all(np.isfinite(best.fit)) and values[8] >= values[0] + values[1] - 1e-9 * abs(values[8])
    reg.close()


def test_chain_in_one_call_equals_single_steps_and_batch(ctx, oracle):
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 100, 110, 10, seed=2)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    gs = api.ProbabilisticSettings(uncertainty=1.5, randomMixture=0.5)

    def chain():
        reg = api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration(maxIterations=40, initialSigma=2.0, endSigma=0.5))
        reg.configureProbabilistic(gs)
        reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
        return reg
    a, b = chain(), chain()
    for _ in range(10):
        a.mcmcChain(1, 77)
    b.mcmcChain(10, 77)
    sa, sb = a.downloadState(), b.downloadState()
    assert np.array_equal(sa.fit, sb.fit) and np.array_equal(sa.modelParameters.shape, sb.modelParameters.shape)
    assert np.array_equal(a.mcmcStats()[1], b.mcmcStats()[1])
    # batched replicas: chain k of the batch == a solo chain with seed + k
    chains = [chain() for _ in range(5)]
    api.mcmc_batch(chains, 8, seed=100)
    ctx.synchronize()
    for k, c in enumerate(chains):
        solo = chain()
        solo.mcmcChain(8, 100 + k)
        s1, s2 = c.downloadState(), solo.downloadState()
        assert np.array_equal(s1.fit, s2.fit) and np.array_equal(c.mcmcStats()[1], solo.mcmcStats()[1])
        solo.close()
    fits = [c.downloadState().fit for c in chains]
    assert not np.array_equal(fits[0], fits[1])      # different seeds, different chains
    for c in chains + [a, b]:
        c.close()


@pytest.mark.parametrize("algo,step_length", [("icp", 1.0), ("icp", 0.5)])
def test_many_chains_run_as_one_batched_kernel_sequence(ctx, oracle, algo, step_length):
    """gingr_mcmc_batch with many chains: one launch per step of the kernel sequence serves all chains (batch.cuh,
    blockIdx.z = chain).  Same chains as solo runs with seed + k -- the Gram of a batched chain runs on fewer CTAs (another
    summation order), so values agree to rounding (1e-9 relative) and the decisions coincide."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 100, 110, 10, seed=3)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    gs = api.ProbabilisticSettings(uncertainty=1.5, randomMixture=0.5)

    def chain():
        reg = api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration(maxIterations=40, initialSigma=2.0, endSigma=0.5))
        reg.configureProbabilistic(gs)
        pars = api.ModelFittingParameters(1.0, np.zeros(3), (0.0, 0.0, 0.0), np.zeros(dm.rank))
        reg.initializeState(general=api.GeneralRegistrationState(pars, np.zeros((dm.M, 3)), globalTransformation=api.RIGID_TRANSFORMS,
                                                                 stepLength=step_length))
        return reg
    n, iters = 48, 6
    chains = [chain() for _ in range(n)]
    api.mcmc_batch(chains, 1, seed=500)
    ctx.synchronize()
    l0 = ctx.launch_count
    api.mcmc_batch(chains, iters - 1, seed=500)
    ctx.synchronize()
    per_step = (ctx.launch_count - l0) / (iters - 1)
    assert per_step < 100, f"{per_step} launches per MH step of {n} chains: the batched sequence did not run"
    for k in (0, 1, 7, n - 1):
        solo = chain()
        solo.mcmcChain(iters, 500 + k)
        v1, c1 = chains[k].mcmcStats()
        v2, c2 = solo.mcmcStats()
        assert np.array_equal(c1, c2), (k, c1, c2)
        np.testing.assert_allclose(v1, v2, rtol=1e-9, atol=1e-9)
        s1, s2 = chains[k].downloadState(), solo.downloadState()
        np.testing.assert_allclose(s1.fit, s2.fit, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(s1.modelParameters.shape, s2.modelParameters.shape, rtol=1e-8, atol=1e-9)
        b1, b2 = chains[k].mcmcBest(), solo.mcmcBest()
        np.testing.assert_allclose(b1.fit, b2.fit, rtol=1e-9, atol=1e-9)
        solo.close()
    # the plan is reused; a different seed rebuilds it
    api.mcmc_batch(chains, 2, seed=900)
    ctx.synchronize()
    assert all(np.all(np.isfinite(c.downloadState().fit)) for c in chains[:4])
    for c in chains:
        c.close()


def test_run_probabilistic_returns_best_sample(ctx, oracle):
    from gingr_b200 import api
    reg, oalgo, ost, diag, keep = _setup(ctx, oracle, "icp", M=100, N=110, r=10)
    gs = api.ProbabilisticSettings(uncertainty=1.5, randomMixture=0.3)
    st0 = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    best = reg.runProbabilistic(st0, gs, seed=5)
    values, counts = reg.mcmcStats()
    assert counts[0] == reg.config.maxIterations - 1
    assert best.status == api.STATUS_MAX_ITERATION and np.all(np.isfinite(best.fit))
    lp_best = sum(reg.logValue(best))
    lp_init = sum(reg.logValue(st0))
    assert abs(lp_best - values[8]) <= 1e-9 * abs(values[8]) and lp_best >= lp_init
    reg.close()


def test_chain_restarts_after_the_state_changed_outside_it(ctx, oracle):
    """initializeState / update between two chains: the second chain must start from the new state (fresh posterior, log
    values, best sample), exactly like a chain on a fresh registration."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 100, 110, 10, seed=2)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    gs = api.ProbabilisticSettings(uncertainty=1.5, randomMixture=0.5)
    cfg = api.IcpConfiguration(maxIterations=40, initialSigma=2.0, endSigma=0.5)
    a = api.IcpRegistration(ctx, dm, dt, cfg)
    a.configureProbabilistic(gs)
    st0 = a.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    a.mcmcChain(6, 5)
    st1 = a.propose(a.propose(st0))                      # deterministic updates in between (host state in / out)
    a.mcmcChain(6, 9)
    b = api.IcpRegistration(ctx, dm, dt, cfg)
    b.configureProbabilistic(gs)
    b.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    b.propose(b.propose(st0))
    b.mcmcChain(6, 9)
    sa, sb = a.downloadState(), b.downloadState()
    assert np.array_equal(sa.fit, sb.fit) and np.array_equal(sa.modelParameters.shape, sb.modelParameters.shape)
    va, ca = a.mcmcStats()
    vb, cb = b.mcmcStats()
    assert np.array_equal(va[:2], vb[:2]) and va[8] == vb[8]          # current and best log values
    a.close(); b.close(); dm.close(); dt.close()


def test_mcmc_argument_errors(ctx, oracle):
    from gingr_b200 import api
    reg, oalgo, ost, diag, keep = _setup(ctx, oracle, "icp", M=60, N=70, r=6)
    with pytest.raises(api.GingrError):
        reg.mcmcChain(1, 0)                                          # not configured
    with pytest.raises(api.GingrError):
        reg.configureProbabilistic(api.ProbabilisticSettings(uncertainty=0.0))
    with pytest.raises(api.GingrError):
        reg.configureProbabilistic(api.ProbabilisticSettings(randomMixture=1.5))
    with pytest.raises(api.GingrError):
        reg.configureProbabilistic(api.ProbabilisticSettings(modelPointIds=np.array([0, 999])))
    reg.configureProbabilistic(api.ProbabilisticSettings())
    with pytest.raises(api.GingrError):
        reg.mcmcChain(1, 0)                                          # no device-resident state yet
    reg.close()


def test_logged_probabilistic_run_is_the_same_chain(ctx, oracle, tmp_path):
    """run(..., acceptRejectLogger, callBackLogger, probabilisticSettings) (GingrAlgorithm.scala:115-175): the logged,
    step-by-step run makes the decisions of the one-call device chain; the JSON log carries one record per sample
    (the initial state first, accepted), generator names of Generator.DefaultRandom, and the proposal's log values."""
    from gingr_b200 import api, io
    gs = api.ProbabilisticSettings(uncertainty=1.5, randomMixture=0.5)
    reg, _, _, _, keep = _setup(ctx, oracle, "icp", M=100, N=110, r=10)
    st0 = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    plain = reg.runProbabilistic(st0, gs, seed=9, iterations=25)
    _, counts_plain = reg.mcmcStats()
    logger = io.JSONStateLogger(path=str(tmp_path / "log.json"))
    seen = []
    logged = reg.runProbabilistic(st0, gs, seed=9, iterations=25, acceptRejectLogger=logger, callBackLogger=seen.append)
    values, counts = reg.mcmcStats()
    assert np.array_equal(counts, counts_plain)
    assert np.array_equal(logged.fit, plain.fit) and np.array_equal(logged.modelParameters.shape, plain.modelParameters.shape)
    assert logger.totalSamples == 26 and len(seen) == 26
    assert logger.accepted == 1 + counts[3] and logger.rejected == 25 - counts[3]
    names = reg.generatorNames(gs)
    assert names == ["ICP", "RotationYaw-0.01", "RotationPitch-0.01", "RotationRoll-0.01", "TranslationX-0.1", "TranslationY-0.1",
                     "TranslationZ-0.1", "RandomShape-1.0", "RandomShape-0.1", "RandomShape-0.01"]
    for leaf, name in enumerate(names):
        recs = [r for r in logger.log[1:] if r.name == name]
        assert len(recs) == counts[8 + leaf] and sum(r.status for r in recs) == counts[18 + leaf]
    # accepted records hold the state the chain moved to; its log values are the evaluators' at that state
    last = max(k for k, r in enumerate(logger.log) if r.status)
    cur = seen[-1]
    assert logger.log[last].modelParameters == [float(v) for v in cur.modelParameters.shape]
    pr, di = reg.logValue(cur)
    lv = logger.log[last].logvalue
    assert abs(lv["Prior"] - pr) <= 1e-9 * abs(pr) and abs(lv["Distance"] - di) <= 1e-9 * abs(di)
    assert abs(lv["product"] - (lv["Prior"] + lv["Distance"])) == 0.0
    logger.write()
    back = io.JSONStateLogger.load(logger.path)
    assert len(back) == 26 and back[0].status and back[0].index == 0
    reg.close()
