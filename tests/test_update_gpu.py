"""Full-iteration parity: gingr_update through the C ABI against the oracle's literal restatement of
GingrAlgorithm.update, per iteration and after a whole run.  Tolerances (north-star): coefficients and fitted
vertices 1e-6 relative (coefficients relative to |alpha|_inf, vertices to the bounding-box diagonal), sigma2
1e-6 relative, status codes equal."""
import dataclasses

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _problem(oracle, M, N, r, seed, with_tri=True):
    from gingr_b200 import synthetic
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, seed + 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, seed)
    return oracle.Gpmm(ref, mean, basis, var, tri), target, tt


def _to_api_state(ost, api):
    p = ost.params
    return api.GeneralRegistrationState(
        api.ModelFittingParameters(p.scale, np.asarray(p.translation, float), tuple(p.euler), p.shape.copy()),
        ost.fit.copy(), ost.sigma2, ost.global_transformation, ost.step_length, "", ost.iteration, ost.status)


def _compare(gst, ost, diag, tol=1e-6):
    a_scale = max(np.max(np.abs(ost.params.shape)), 1e-12)
    assert np.max(np.abs(gst.modelParameters.shape - ost.params.shape)) < tol * a_scale
    assert np.max(np.abs(gst.fit - ost.fit)) < tol * diag
    assert abs(gst.sigma2 - ost.sigma2) <= tol * abs(ost.sigma2)
    assert np.max(np.abs(np.asarray(gst.modelParameters.translation) - ost.params.translation)) < tol * diag
    assert np.max(np.abs(np.asarray(gst.modelParameters.euler) - np.asarray(ost.params.euler))) < tol
    assert abs(gst.modelParameters.scale - ost.params.scale) < tol
    assert gst.status == ost.status and gst.iteration == ost.iteration


@pytest.mark.parametrize("gt", ["none", "rigid", "similarity"])
@pytest.mark.parametrize("w", [0.0, 0.1])
def test_cpd_iterations_match_oracle(ctx, oracle, gt, w):
    from gingr_b200 import api
    gtc = {"none": oracle.NO_TRANSFORMS, "rigid": oracle.RIGID_TRANSFORMS, "similarity": oracle.SIMILARITY_TRANSFORMS}[gt]
    m, target, tt = _problem(oracle, 100, 100, 50, seed=0)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    cfg = api.CpdConfiguration(maxIterations=100, w=w)
    reg = api.CpdRegistration(ctx, dm, dt, cfg)
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(max_iterations=100, w=w))
    ost = oalgo.initialize(oracle.initial_state(m, target, tt, global_transformation=gtc))
    gst = reg.initializeState(globalTransformation=gtc)
    assert abs(gst.sigma2 - ost.sigma2) < 1e-12 * ost.sigma2
    assert np.max(np.abs(gst.fit - ost.fit)) < 1e-10
    for it in range(6):
        # feed the oracle's state to the GPU each iteration (per-iteration parity, no drift accumulation)
        gst = reg.propose(_to_api_state(ost, api))
        ost = oracle.propose(oalgo, ost)
        _compare(gst, ost, diag)
    reg.close()


def test_cpd_chained_run_matches_oracle(ctx, oracle):
    """A whole deterministic run (GingrAlgorithm.run): maxIterations - 1 proposals, states chained on each side."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 100, 100, 50, seed=1)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(maxIterations=30))
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(max_iterations=30))
    ofinal = oracle.run(oalgo, oracle.initial_state(m, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
    gfinal = reg.run(reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS))
    _compare(gfinal, ofinal, diag, tol=1e-6)
    assert gfinal.iteration == ofinal.iteration and gfinal.status == ofinal.status
    assert gfinal.status in (api.STATUS_MAX_ITERATION, api.STATUS_CONVERGED)
    # device-resident chaining gives the same end state as host round trips
    st0 = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    reg.updateChain(gfinal.iteration)
    chained = reg.downloadState()
    assert np.max(np.abs(chained.fit - gfinal.fit)) < 1e-9 * diag
    assert chained.iteration == gfinal.iteration
    reg.close()


def test_cpd_democpd_setting(ctx, oracle):
    """DemoCPD.scala:21-23: initialSigma = 1, NoTransforms, femur-like scale: kernel values down to 1e-150."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 100, 100, 50, seed=2)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(maxIterations=100, initialSigma=1.0))
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(max_iterations=100, initial_sigma=1.0))
    ost = oalgo.initialize(oracle.initial_state(m, target, tt, global_transformation=oracle.NO_TRANSFORMS))
    gst = reg.initializeState(globalTransformation=api.NO_TRANSFORMS)
    assert gst.sigma2 == 1.0
    for it in range(5):
        gst = reg.propose(_to_api_state(ost, api))
        ost = oracle.propose(oalgo, ost)
        assert gst.status == ost.status
        if ost.status == oracle.STATUS_NONE:
            _compare(gst, ost, diag)
    reg.close()


def test_cpd_with_landmarks(ctx, oracle):
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 120, 140, 40, seed=3)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    rng = np.random.default_rng(4)
    lm_pid = np.array([3, 40, 77, 111], dtype=np.int32)
    lm_pts = m.ref[lm_pid] + rng.normal(scale=2.0, size=(4, 3)) + 5.0
    A = rng.normal(size=(4, 3, 3))
    lm_cov = A @ np.transpose(A, (0, 2, 1)) + np.eye(3)
    lm_cov[0] = np.eye(3)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(maxIterations=50, w=0.05))
    reg.setLandmarks(lm_pid, lm_pts, lm_cov)
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(max_iterations=50, w=0.05))
    ost = oalgo.initialize(oracle.initial_state(m, target, tt, global_transformation=oracle.RIGID_TRANSFORMS,
                                                landmarks=oracle.Landmarks(lm_pid, lm_pts, lm_cov)))
    for it in range(4):
        gst = reg.propose(_to_api_state(ost, api))
        ost = oracle.propose(oalgo, ost)
        _compare(gst, ost, diag)
    reg.close()


def test_cpd_step_length_and_initial_pose(ctx, oracle):
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 90, 110, 30, seed=5)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    R0, t0 = oracle.euler_to_matrix(0.05, -0.03, 0.02), np.array([2.0, -1.0, 3.0])
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration())
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig())
    ost = oalgo.initialize(oracle.initial_state(m, target, tt, global_transformation=oracle.RIGID_TRANSFORMS, R0=R0, t0=t0))
    ost = dataclasses.replace(ost, step_length=0.5)
    gst0 = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS, rotation=R0, translation=t0)
    assert np.max(np.abs(gst0.fit - ost.fit)) < 1e-10
    for it in range(4):
        gst = reg.propose(_to_api_state(ost, api))
        ost = oracle.propose(oalgo, ost)
        _compare(gst, ost, diag)
    reg.close()


def test_cpd_model_flexibility_error_status(ctx, oracle):
    """A target point far from every model point with w = 0 and a small sigma2: the column underflows, P has
    NaNs, the posterior fails; iteration 0 returns the state unchanged, later iterations set
    ModelFlexibilityError (GingrAlgorithm.scala:194-208)."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 60, 60, 20, seed=6)
    target = target.copy()
    target[0] = [5000.0, 0.0, 0.0]
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target)
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(initialSigma=0.5))
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(initial_sigma=0.5))
    ost = oalgo.initialize(oracle.initial_state(m, target, None, global_transformation=oracle.NO_TRANSFORMS))
    g1 = reg.propose(_to_api_state(ost, api))
    o1 = oracle.propose(oalgo, ost)
    assert o1.status == oracle.STATUS_NONE and g1.status == api.STATUS_NONE          # iteration 0: unchanged
    assert np.array_equal(g1.modelParameters.shape, ost.params.shape) and g1.iteration == 1
    g2 = reg.propose(_to_api_state(o1, api))
    o2 = oracle.propose(oalgo, o1)
    assert o2.status == oracle.STATUS_MODEL_FLEXIBILITY_ERROR
    assert g2.status == api.STATUS_MODEL_FLEXIBILITY_ERROR
    reg.close()


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("method", ["triangular", "along_normal", "pointcloud"])
def test_icp_iterations_match_oracle(ctx, oracle, method, reverse):
    """All three ICPCorrespondenceMethods (ICP.scala:29-33) in both directions (reverseCorrespondenceDirection,
    ICP.scala:46-49): the reversed direction observes target points at template vertices, several per vertex."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 200, 260, 40, seed=7)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    om = {"triangular": oracle.METHOD_TRIANGULAR, "along_normal": oracle.METHOD_ALONG_NORMAL,
          "pointcloud": oracle.METHOD_POINTCLOUD}[method]
    gm = {"triangular": api.TRIANGULAR_CLOSEST_POINT, "along_normal": api.ALONG_NORMAL_CLOSEST_POINT,
          "pointcloud": api.POINTCLOUD_CLOSEST_POINT}[method]
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    reg = api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration(maxIterations=100, initialSigma=2.0, endSigma=0.01,
                                                                correspondenceMethod=gm,
                                                                reverseCorrespondenceDirection=reverse))
    oalgo = oracle.IcpAlgorithm(oracle.IcpConfig(max_iterations=100, initial_sigma=2.0, end_sigma=0.01, method=om,
                                                 reverse=reverse))
    ost = oalgo.initialize(oracle.initial_state(m, target, tt, global_transformation=oracle.NO_TRANSFORMS))
    gst = reg.initializeState(globalTransformation=api.NO_TRANSFORMS)
    assert gst.sigma2 == 2.0
    for it in range(5):
        gst = reg.propose(_to_api_state(ost, api))
        ost = oracle.propose(oalgo, ost)
        _compare(gst, ost, diag)
    reg.close()


def test_icp_run_demoicp_setting(ctx, oracle):
    """DemoICP.scala:22-24: maxIterations = 100 -> 99 updates, sigma2 1 -> 1, NoTransforms."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 100, 100, 50, seed=8)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    reg = api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration(maxIterations=100, initialSigma=1.0, endSigma=1.0))
    oalgo = oracle.IcpAlgorithm(oracle.IcpConfig(max_iterations=100, initial_sigma=1.0, end_sigma=1.0))
    ofinal = oracle.run(oalgo, oracle.initial_state(m, target, tt, global_transformation=oracle.NO_TRANSFORMS))
    gfinal = reg.run(reg.initializeState(globalTransformation=api.NO_TRANSFORMS))
    assert gfinal.iteration == 99 and ofinal.iteration == 99
    _compare(gfinal, ofinal, diag, tol=1e-6)
    reg.close()


def test_update_resumes_from_device_state(ctx, oracle):
    """Feeding back exactly the state the previous call returned reuses the device-resident fit; feeding
    anything else re-seeds from (alpha, pose, sigma2).  Both give the same result."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 80, 90, 25, seed=9)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration())
    s0 = reg.initializeState()
    s1 = reg.propose(s0)
    s2 = reg.propose(s1)                       # resumed
    reg2 = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration())
    t2 = reg2.propose(dataclasses.replace(s1))  # fresh handle: re-seeded
    assert np.array_equal(s2.fit, t2.fit) and np.array_equal(s2.modelParameters.shape, t2.modelParameters.shape)
    reg.close()
    reg2.close()


@pytest.mark.parametrize("algo", ["cpd", "icp"])
def test_probabilistic_update_matches_oracle(ctx, oracle, algo):
    """update(current, probabilistic = true): posterior.sample() (GingrAlgorithm.scala:211) with the documented
    Philox/Box-Muller stream -- coefficients c + L^-T z.  Same tolerances as the deterministic branch."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 150, 170, 30, seed=11)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    if algo == "cpd":
        reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(w=0.05))
        oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(w=0.05))
    else:
        reg = api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration(initialSigma=3.0, endSigma=1.0))
        oalgo = oracle.IcpAlgorithm(oracle.IcpConfig(initial_sigma=3.0, end_sigma=1.0))
    ost = oalgo.initialize(oracle.initial_state(m, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
    det = oracle.propose(oalgo, ost)
    for it in range(4):
        seed = 0x1234567890ABCDEF + it
        gst = reg.propose(_to_api_state(ost, api), probabilistic=True, seed=seed)
        ost = oracle.propose(oalgo, ost, probabilistic=True, seed=seed)
        _compare(gst, ost, diag)
        assert gst.generatedBy == "Stochastic"
    # the sample is not the mean
    assert np.max(np.abs(ost.params.shape - det.params.shape)) > 1e-3
    reg.close()


def test_probabilistic_chain_is_reproducible_and_seed_dependent(ctx, oracle):
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 120, 130, 24, seed=12)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    outs = []
    for seed in (7, 7, 8):
        reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(w=0.1))
        reg.initializeState()
        reg.updateChainSampled(5, seed)
        outs.append(reg.downloadState())
        reg.close()
    assert np.array_equal(outs[0].fit, outs[1].fit) and outs[0].iteration == 5
    assert not np.array_equal(outs[0].fit, outs[2].fit)
    # against the oracle, step by step with the device's iteration counter as the Philox counter
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(w=0.1))
    ost = oalgo.initialize(oracle.initial_state(m, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
    for _ in range(5):
        ost = oracle.propose(oalgo, ost, probabilistic=True, seed=7)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    assert np.max(np.abs(outs[0].fit - ost.fit)) < 1e-6 * diag


def test_probabilistic_retry_counter(ctx, oracle):
    """A failing posterior in the probabilistic branch returns the state unchanged retryCounter (= 10) times and
    only then sets ModelFlexibilityError (GingrAlgorithm.scala:69-70, :195-202)."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 60, 60, 20, seed=6)
    target = target.copy()
    target[0] = [5000.0, 0.0, 0.0]
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target)
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(initialSigma=0.5))
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(initial_sigma=0.5))
    ost = oalgo.initialize(oracle.initial_state(m, target, None, global_transformation=oracle.NO_TRANSFORMS))
    gst = _to_api_state(ost, api)
    statuses_g, statuses_o = [], []
    for it in range(13):
        gst = reg.propose(gst, probabilistic=True, seed=it)
        ost = oracle.propose(oalgo, ost, probabilistic=True, seed=it)
        statuses_g.append(gst.status)
        statuses_o.append(ost.status)
    assert statuses_g == statuses_o
    assert statuses_g[:11] == [api.STATUS_NONE] * 11 and statuses_g[11] == api.STATUS_MODEL_FLEXIBILITY_ERROR
    reg.close()


def test_large_batch_runs_as_one_kernel_sequence(ctx, oracle):
    """gingr_update_batch with many ICP chains: one launch per step of the iteration's kernel sequence serves all chains
    (batch.cuh).  A batched chain's Gram runs on fewer CTAs (another summation order): 1e-9 against the solo chains."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 100, 100, 30, seed=14)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    n, iters = 64, 3
    cfg = api.IcpConfiguration(initialSigma=2.0, endSigma=1.0)
    chains = [api.IcpRegistration(ctx, dm, dt, cfg) for _ in range(n)]
    for k, c in enumerate(chains):
        c.initializeState(translation=np.array([0.01 * k, 0.0, 0.0]))       # different states, same kernel sequence
    l0 = ctx.launch_count
    api.update_batch(chains, iters)
    ctx.synchronize()
    assert (ctx.launch_count - l0) / iters < 100, "the batched sequence did not run"
    for k in (0, 5, n - 1):
        solo = api.IcpRegistration(ctx, dm, dt, cfg)
        solo.initializeState(translation=np.array([0.01 * k, 0.0, 0.0]))
        solo.updateChain(iters)
        a, b = chains[k].downloadState(), solo.downloadState()
        assert a.iteration == iters == b.iteration
        np.testing.assert_allclose(a.fit, b.fit, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(a.modelParameters.shape, b.modelParameters.shape, rtol=1e-8, atol=1e-9)
        solo.close()
    for c in chains:
        c.close()


def test_large_cpd_batch_runs_as_one_kernel_sequence(ctx, oracle):
    """The same for CPD registrations: E-step sweeps, Gram, factorisation and the update of 48 registrations as one batched
    sequence (1e-9 against the solo runs: the batched Gram runs on fewer CTAs)."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 120, 150, 20, seed=15)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    n, iters = 48, 3
    cfg = api.CpdConfiguration(maxIterations=40, w=0.1)
    chains = [api.CpdRegistration(ctx, dm, dt, cfg) for _ in range(n)]
    for k, c in enumerate(chains):
        c.initializeState(translation=np.array([0.02 * k, 0.0, 0.0]))
    l0 = ctx.launch_count
    api.update_batch(chains, iters)
    ctx.synchronize()
    assert (ctx.launch_count - l0) / iters < 100, "the batched sequence did not run"
    for k in (0, 9, n - 1):
        solo = api.CpdRegistration(ctx, dm, dt, cfg)
        solo.initializeState(translation=np.array([0.02 * k, 0.0, 0.0]))
        solo.updateChain(iters)
        a, b = chains[k].downloadState(), solo.downloadState()
        assert a.iteration == iters == b.iteration
        np.testing.assert_allclose(a.fit, b.fit, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(a.sigma2, b.sigma2, rtol=1e-9)
        np.testing.assert_allclose(a.modelParameters.shape, b.modelParameters.shape, rtol=1e-8, atol=1e-9)
        solo.close()
    for c in chains:
        c.close()


def test_batched_chains_equal_independent_chains(ctx, oracle):
    """gingr_update_batch (BASELINE config 5: batched MCMC chains, replicas only): n chains sharing model and target,
    each with its own state and Philox stream, give exactly what the same chains give one by one."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 100, 100, 30, seed=13)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    n, iters, seed = 24, 4, 1000
    cfg = api.IcpConfiguration(initialSigma=2.0, endSigma=1.0)
    chains = [api.IcpRegistration(ctx, dm, dt, cfg) for _ in range(n)]
    for c in chains:
        c.initializeState()
    api.update_batch(chains, iters, probabilistic=True, seed=seed)
    got = [c.downloadState() for c in chains]
    for k in (0, 1, 7, 23):
        solo = api.IcpRegistration(ctx, dm, dt, cfg)
        solo.initializeState()
        solo.updateChainSampled(iters, seed + k)
        ref = solo.downloadState()
        assert np.array_equal(got[k].fit, ref.fit) and got[k].iteration == iters
        assert np.array_equal(got[k].modelParameters.shape, ref.modelParameters.shape)
        solo.close()
    assert not np.array_equal(got[0].fit, got[1].fit)          # different seeds, different chains
    # and against the oracle for one chain
    oalgo = oracle.IcpAlgorithm(oracle.IcpConfig(initial_sigma=2.0, end_sigma=1.0))
    ost = oalgo.initialize(oracle.initial_state(m, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
    for _ in range(iters):
        ost = oracle.propose(oalgo, ost, probabilistic=True, seed=seed + 7)
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    assert np.max(np.abs(got[7].fit - ost.fit)) < 1e-6 * diag
    for c in chains:
        c.close()
