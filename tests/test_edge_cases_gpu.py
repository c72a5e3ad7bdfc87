"""Edge cases through the C ABI: degenerate sizes, ragged shapes (nothing a multiple of a tile), empty observation
sets, bad arguments and non-finite inputs.  Each case is checked against the oracle or against the documented error
convention of include/gingr_cuda.h (0 OK, 1 MODEL_FLEXIBILITY, negative = error with text)."""
import dataclasses

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_model(M, r, seed):
    rng = np.random.default_rng(seed)
    ref = rng.normal(scale=40.0, size=(M, 3))
    basis = rng.normal(size=(3 * M, r))
    basis /= np.linalg.norm(basis, axis=0, keepdims=True)
    var = 50.0 * 0.9 ** np.arange(r)
    mean = rng.normal(scale=0.5, size=3 * M)
    return ref, mean, basis, var


@pytest.mark.parametrize("M,N,r", [(1, 1, 1), (2, 3, 1), (3, 2, 5), (33, 65, 9), (70, 31, 67), (129, 257, 130)])
def test_cpd_ragged_sizes_match_oracle(ctx, oracle, M, N, r):
    """Sizes that are multiples of no tile (rows of Phi % 32, r % 8, r % 64, r % 128, M % 256, N % 1024 all != 0),
    including single points and rank 1; r > 3M (rank-deficient Gram) included."""
    from gingr_b200 import api
    ref, mean, basis, var = _rand_model(M, r, 3)
    target = np.random.default_rng(4).normal(scale=40.0, size=(N, 3))
    m = oracle.Gpmm(ref, mean, basis, var, None)
    dm = api.Model(ctx, ref, mean, basis, var)
    dt = api.Target(ctx, target)
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(w=0.2))
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(w=0.2))
    ost = oalgo.initialize(oracle.initial_state(m, target, None, global_transformation=oracle.NO_TRANSFORMS))
    gst = reg.initializeState(globalTransformation=api.NO_TRANSFORMS)
    assert abs(gst.sigma2 - ost.sigma2) <= 1e-12 * max(ost.sigma2, 1e-300)
    diag = max(np.linalg.norm(ref.max(0) - ref.min(0)), 1.0)
    for _ in range(3):
        gst = reg.propose(gst)
        ost = oracle.propose(oalgo, ost)
        assert gst.status == ost.status
        if ost.status == oracle.STATUS_NONE:
            assert np.max(np.abs(gst.fit - ost.fit)) < 1e-6 * diag
            assert np.max(np.abs(gst.modelParameters.shape - ost.params.shape)) < 1e-6 * max(np.max(np.abs(ost.params.shape)), 1e-9)
    reg.close()


def test_icp_without_any_accepted_pair_returns_the_prior_mean(ctx, oracle):
    """Every correspondence rejected (w = 0: the template's normals point against the target's, ICP.scala:50 keeps
    no pair): the regression has no observation, M = I, the posterior mean is the prior mean."""
    from gingr_b200 import api, synthetic
    ref, tri = synthetic.sphere_mesh(120)
    mean, basis, var = synthetic.make_gpmm(ref, 20, 5)
    tv, tt = synthetic.sphere_mesh(150, radius=105.0)
    tt_flipped = tt[:, [0, 2, 1]]                                # inward normals: opposite to the template's
    m = oracle.Gpmm(ref, mean, basis, var, tri)
    dm = api.Model(ctx, ref, mean, basis, var, tri)
    dt = api.Target(ctx, tv, tt_flipped)
    reg = api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration(initialSigma=1.0, endSigma=1.0))
    oalgo = oracle.IcpAlgorithm(oracle.IcpConfig(initial_sigma=1.0, end_sigma=1.0))
    ost = oalgo.initialize(oracle.initial_state(m, tv, tt_flipped, global_transformation=oracle.RIGID_TRANSFORMS))
    pids, _ = oracle.icp_correspondence(oracle.METHOD_TRIANGULAR, False, ost.fit, tri, tv, tt_flipped)
    assert len(pids) == 0
    gst = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    gst = reg.propose(gst)
    ost = oracle.propose(oalgo, ost)
    assert gst.status == ost.status == oracle.STATUS_NONE
    assert np.max(np.abs(gst.modelParameters.shape)) < 1e-9 and np.max(np.abs(ost.params.shape)) < 1e-9
    assert np.max(np.abs(gst.fit - ost.fit)) < 1e-9
    reg.close()


def test_bad_arguments_are_errors_with_text(ctx):
    from gingr_b200 import api
    from gingr_b200._native import GingrError
    ref, mean, basis, var = _rand_model(10, 4, 0)
    with pytest.raises(GingrError) as e:
        api.Target(ctx, np.zeros((0, 3)))                        # empty target
    assert e.value.code == -1 and "bad argument" in str(e.value)
    with pytest.raises(GingrError):
        api.Model(ctx, ref, mean, basis, -var)                   # negative variance
    with pytest.raises(GingrError):
        api.Model(ctx, ref, mean, basis, var, np.array([[0, 1, 99]]))   # triangle index out of range
    dm = api.Model(ctx, ref, mean, basis, var)
    dt = api.Target(ctx, np.random.default_rng(0).normal(size=(7, 3)))
    with pytest.raises(GingrError):
        api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(w=1.0))   # w must be < 1
    with pytest.raises(GingrError):
        api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration())       # triangular flavour without triangles
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration())
    st = reg.initializeState()
    bad = dataclasses.replace(st, modelParameters=dataclasses.replace(st.modelParameters, shape=np.zeros(3)))
    with pytest.raises(GingrError):
        reg.update(bad)                                          # alpha of the wrong rank
    with pytest.raises(GingrError):
        reg.setLandmarks([42], np.zeros((1, 3)))                 # landmark vertex id out of range
    reg.close()


def test_non_finite_inputs_fail_like_the_reference(ctx, oracle):
    """NaN in the target: the reference's P is NaN everywhere, the posterior fails (state unchanged at iteration 0,
    ModelFlexibilityError afterwards; GingrAlgorithm.scala:194-208)."""
    from gingr_b200 import api
    ref, mean, basis, var = _rand_model(20, 6, 1)
    target = np.random.default_rng(2).normal(scale=40.0, size=(25, 3))
    target[3, 1] = np.nan
    dm = api.Model(ctx, ref, mean, basis, var)
    dt = api.Target(ctx, target)
    P1, Pt1, PX = api.cpd_estep(ctx, dt, ref, 10.0, 0.1)
    assert np.all(np.isnan(P1)) and np.all(np.isnan(PX))
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(initialSigma=10.0))
    s0 = reg.initializeState()
    s1 = reg.propose(s0)
    assert s1.status == api.STATUS_NONE and np.array_equal(s1.modelParameters.shape, s0.modelParameters.shape)
    s2 = reg.propose(s1)
    assert s2.status == api.STATUS_MODEL_FLEXIBILITY_ERROR
    reg.close()


def test_posterior_with_zero_observations_and_duplicates(ctx, oracle):
    from gingr_b200 import api
    ref, mean, basis, var = _rand_model(15, 5, 7)
    m = oracle.Gpmm(ref, mean, basis, var, None)
    dm = api.Model(ctx, ref, mean, basis, var)
    R, t = np.eye(3), np.zeros(3)
    c, mesh = api.posterior_mean(ctx, dm, R, t, np.zeros(0, dtype=np.int32), np.zeros((0, 3)), np.zeros(0))
    assert np.max(np.abs(c)) == 0.0
    assert np.max(np.abs(mesh - m.instance(np.zeros(5)))) < 1e-12
    pids = np.array([4, 4, 4, 9], dtype=np.int32)                 # the same vertex observed three times
    pts = ref[pids] + np.random.default_rng(1).normal(size=(4, 3))
    noise = np.array([0.5, 2.0, 1.0, 0.1])
    c, mesh = api.posterior_mean(ctx, dm, R, t, pids, pts, noise)
    cov = np.eye(3)[None] * noise[:, None, None]
    c_ref, _ = m.posterior_coefficients(pids, pts, cov)
    assert np.max(np.abs(c - c_ref)) < 1e-9 * max(np.max(np.abs(c_ref)), 1e-9)


def test_update_with_an_older_state_after_a_device_chain(ctx, oracle):
    """After gingr_update_chain the device no longer holds the state the host last saw: handing that older state to
    gingr_update must re-seed the device (not silently resume from the chain's end)."""
    from gingr_b200 import api, synthetic
    ref, tri = synthetic.sphere_mesh(120)
    mean, basis, var = synthetic.make_gpmm(ref, 12, 1)
    tv, tt = synthetic.sphere_mesh(130)
    target = synthetic.make_target(tv, 0)
    dm = api.Model(ctx, ref, mean, basis, var, tri)
    dt = api.Target(ctx, target, tt)
    a = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(w=0.1))
    b = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(w=0.1))
    st0 = a.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    b.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    a.updateChain(3)
    got = a.propose(st0)
    want = b.propose(st0)
    assert np.array_equal(got.fit, want.fit) and got.sigma2 == want.sigma2
    # and the downloaded end state of a chain can be continued
    a.updateChain(2)
    end = a.downloadState()
    nxt = a.propose(end)
    assert np.all(np.isfinite(nxt.fit)) and nxt.iteration == end.iteration + 1
    for x in (a, b, dm, dt):
        x.close()
