"""K3 parity: weighted Gram (DMMA) + Cholesky + mean evaluation through the C ABI against the oracle's
literal restatement of scalismo's regression (pinv via SVD).  Tolerance: 1e-8 relative on coefficients
(north-star bar: 1e-6) and 1e-9 of the bounding-box diagonal on vertices."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _model(oracle, M, r, seed=0, orthonormal=True):
    from gingr_b200 import synthetic
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, seed, orthonormal=orthonormal)
    mean = np.random.default_rng(seed + 7).normal(scale=0.5, size=3 * M)
    return oracle.Gpmm(ref, mean, basis, var, tri)


def _upload(ctx, m):
    from gingr_b200 import api
    return api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)


@pytest.mark.parametrize("M,r", [(60, 12), (100, 50), (300, 130), (1000, 257)])
def test_posterior_mean_isotropic(ctx, oracle, M, r):
    from gingr_b200 import api
    m = _model(oracle, M, r)
    dm = _upload(ctx, m)
    rng = np.random.default_rng(1)
    R, t = oracle.euler_to_matrix(0.2, -0.1, 0.3), np.array([3.0, -2.0, 1.0])
    posed = m.transform(R, t)
    obs = posed.instance(rng.normal(size=r)) + rng.normal(scale=0.3, size=(M, 3))
    var = rng.uniform(0.05, 5.0, size=M)
    pids = rng.permutation(M)[: max(3, (3 * M) // 4)].astype(np.int32)
    c, mesh = api.posterior_mean(ctx, dm, R, t, pids, obs[pids], var[pids])
    cov = np.eye(3)[None] * var[pids][:, None, None]
    c_ref, _ = posed.posterior_coefficients(pids, obs[pids], cov)
    assert rel_err(c, c_ref) < 1e-8
    diag = np.linalg.norm(m.ref.max(0) - m.ref.min(0))
    assert np.max(np.abs(mesh - posed.instance(c_ref))) < 1e-9 * diag


def test_posterior_mean_full_covariance_and_duplicates(ctx, oracle):
    from gingr_b200 import api
    M, r = 120, 40
    m = _model(oracle, M, r, seed=3)
    dm = _upload(ctx, m)
    rng = np.random.default_rng(2)
    R, t = oracle.euler_to_matrix(-0.3, 0.2, 0.1), np.array([1.0, 2.0, 3.0])
    posed = m.transform(R, t)
    pids = np.array([5, 17, 17, 60, 99, 3], dtype=np.int32)        # vertex 17 observed twice
    obs = posed.instance(rng.normal(size=r))[pids] + rng.normal(scale=0.5, size=(len(pids), 3))
    A = rng.normal(size=(len(pids), 3, 3))
    cov = A @ np.transpose(A, (0, 2, 1)) + 0.1 * np.eye(3)
    c, mesh = api.posterior_mean(ctx, dm, R, t, pids, obs, cov)
    c_ref, _ = posed.posterior_coefficients(pids, obs, cov)
    assert rel_err(c, c_ref) < 1e-8


def test_posterior_weights_span_many_orders(ctx, oracle):
    """CPD hands over variances sigma2/P1 with P1 down to 1e-100: no scaling tricks."""
    from gingr_b200 import api
    M, r = 200, 60
    m = _model(oracle, M, r, seed=4)
    dm = _upload(ctx, m)
    rng = np.random.default_rng(5)
    R, t = np.eye(3), np.zeros(3)
    obs = m.instance(rng.normal(size=r)) + rng.normal(scale=0.2, size=(M, 3))
    var = 10.0 ** rng.uniform(-3, 100, size=M)
    var[:20] = 10.0 ** rng.uniform(-4, -2, size=20)
    pids = np.arange(M, dtype=np.int32)
    c, _ = api.posterior_mean(ctx, dm, R, t, pids, obs, var)
    c_ref, _ = m.posterior_coefficients(pids, obs, np.eye(3)[None] * var[:, None, None])
    assert rel_err(c, c_ref) < 1e-8


def test_posterior_non_finite_noise_is_model_flexibility(ctx, oracle):
    from gingr_b200 import api
    m = _model(oracle, 60, 12)
    dm = _upload(ctx, m)
    pids = np.arange(60, dtype=np.int32)
    var = np.ones(60)
    var[7] = np.inf                                   # P1 = 0 in CPD.scala:125
    with pytest.raises(FloatingPointError):
        api.posterior_mean(ctx, dm, np.eye(3), np.zeros(3), pids, m.instance(np.zeros(12)), var)


@pytest.mark.parametrize("orthonormal", [True, False])
@pytest.mark.parametrize("M,r", [(80, 20), (400, 150)])
def test_coefficients(ctx, oracle, M, r, orthonormal):
    from gingr_b200 import api
    m = _model(oracle, M, r, seed=6, orthonormal=orthonormal)
    dm = _upload(ctx, m)
    rng = np.random.default_rng(7)
    R, t = oracle.euler_to_matrix(0.1, 0.2, -0.2), np.array([-1.0, 0.5, 2.0])
    posed = m.transform(R, t)
    mesh = posed.instance(rng.normal(size=r)) + rng.normal(scale=0.1, size=(M, 3))
    c = api.coefficients(ctx, dm, R, t, mesh)
    c_ref = posed.coefficients(mesh)
    assert rel_err(c, c_ref) < 1e-8


def test_coefficients_rank_deficient_decimated_model(ctx, oracle):
    """runDecimated models: rows copied from nearest vertices, r close to 3M, basis far from orthonormal
     -- the regression constants are badly conditioned; parity must still hold to 1e-6."""
    from gingr_b200 import api, synthetic
    M, r = 40, 100
    ref, tri = synthetic.sphere_mesh(M)
    big_ref, _ = synthetic.sphere_mesh(400)
    _, big_basis, var = synthetic.make_gpmm(big_ref, r, 11)
    nn = ((ref[:, None] - big_ref[None]) ** 2).sum(-1).argmin(1)
    rows = (3 * nn[:, None] + np.arange(3)[None]).reshape(-1)
    m = oracle.Gpmm(ref, np.zeros(3 * M), np.ascontiguousarray(big_basis[rows]), var * 50, tri)
    dm = _upload(ctx, m)
    rng = np.random.default_rng(8)
    mesh = m.instance(rng.normal(size=r))
    c = api.coefficients(ctx, dm, np.eye(3), np.zeros(3), mesh)
    c_ref = m.coefficients(mesh)
    # the coefficients themselves are ill-determined (r > rank); what the reference consumes is the instance
    diag = np.linalg.norm(ref.max(0) - ref.min(0))
    assert np.max(np.abs(m.instance(c) - m.instance(c_ref))) < 1e-6 * diag


def test_model_instance(ctx, oracle):
    from gingr_b200 import api, _native as nat
    import ctypes
    m = _model(oracle, 150, 33)
    dm = _upload(ctx, m)
    rng = np.random.default_rng(9)
    alpha = rng.normal(size=33)
    p = oracle.Params(1.3, np.array([1.0, -2.0, 0.5]), (0.3, -0.2, 0.4), alpha)
    ref_fit = oracle.model_instance_shape_pose_scale(m, p)
    st = api.GeneralRegistrationState(api.ModelFittingParameters(p.scale, p.translation, p.euler, alpha), None)
    pod, a = st.to_pod()
    fit = np.empty((150, 3))
    ctx.check(ctx._lib.gingr_model_instance(ctx.handle, dm.handle, ctypes.byref(pod), nat.as_dp(a), nat.as_dp(fit)))
    assert np.max(np.abs(fit - ref_fit)) < 1e-11


def test_gram_large_property(ctx):
    """Size-independent property at a bench-like shape: posterior of exact observations of instance(a) with tiny
    noise reproduces a (orthonormal basis): exercises the stream-K DMMA Gram with many tiles and segments."""
    from gingr_b200 import api, synthetic
    M, r = 6000, 700
    ref = synthetic.fibonacci_sphere(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 21)
    dm = api.Model(ctx, ref, mean, basis, var)
    rng = np.random.default_rng(10)
    a = rng.normal(size=r)
    obs = ref + (basis @ (np.sqrt(var) * a)).reshape(-1, 3)
    noise = np.full(M, 1e-8)
    c, mesh = api.posterior_mean(ctx, dm, np.eye(3), np.zeros(3), np.arange(M, dtype=np.int32), obs, noise)
    expect = a * var / (var + 1e-8)
    assert rel_err(c, expect) < 1e-9
    assert np.max(np.abs(mesh - obs)) < 1e-5


def test_gram_more_tiles_than_sms(ctx, oracle):
    """r = 2304 -> 171 lower-triangular tiles > 148 SMs: the schedule has a full round plus a main/tail round."""
    from gingr_b200 import api, synthetic
    M, r = 900, 2304
    ref = synthetic.fibonacci_sphere(M)
    rng = np.random.default_rng(31)
    basis = rng.normal(size=(3 * M, r)) / np.sqrt(3 * M)
    var = 50.0 * 0.999 ** np.arange(r)
    m = oracle.Gpmm(ref, np.zeros(3 * M), basis, var, None)
    dm = api.Model(ctx, ref, m.mean, basis, var)
    obs = ref + rng.normal(scale=1.0, size=(M, 3))
    noise = rng.uniform(0.5, 2.0, size=M)
    pids = np.arange(M, dtype=np.int32)
    c, mesh = api.posterior_mean(ctx, dm, np.eye(3), np.zeros(3), pids, obs, noise)
    c_ref, _ = m.posterior_coefficients(pids, obs, np.eye(3)[None] * noise[:, None, None])
    diag = np.linalg.norm(ref.max(0) - ref.min(0))
    assert np.max(np.abs(mesh - m.instance(c_ref))) < 1e-9 * diag


@pytest.mark.parametrize("M,r,full", [(60, 12, False), (300, 130, False), (120, 40, True), (700, 200, False)])
def test_posterior_covariance_at_the_mesh_points(ctx, oracle, M, r, full):
    """gingr_posterior_covariance: cov_i = R Phi_i D Mx^-1 D Phi_i^T R^T against the oracle's dense form
    Q'_i Minv Q'_i^T with Minv = pinv(Mx) of the literal regression (scalismo posterior; what
    helper/PosteriorHelper.scala:26-80 visualises).  Tolerance 1e-9 of the largest prior variance."""
    from gingr_b200 import api
    m = _model(oracle, M, r, seed=5)
    dm = _upload(ctx, m)
    rng = np.random.default_rng(8)
    R, t = oracle.euler_to_matrix(0.25, -0.15, 0.1), np.array([2.0, -1.0, 0.5])
    posed = m.transform(R, t)
    pids = rng.permutation(M)[: max(3, M // 2)].astype(np.int32)
    obs = posed.instance(rng.normal(size=r))[pids] + rng.normal(scale=0.3, size=(len(pids), 3))
    if full:
        A = rng.normal(size=(len(pids), 3, 3))
        noise = A @ np.transpose(A, (0, 2, 1)) + 0.1 * np.eye(3)
        cov_obs = noise
    else:
        noise = rng.uniform(0.05, 5.0, size=len(pids))
        cov_obs = np.eye(3)[None] * noise[:, None, None]
    got = api.posterior_covariance(ctx, dm, R, t, pids, obs, noise)
    _, Minv = posed.posterior_coefficients(pids, obs, cov_obs)
    Q = (posed.basis * np.sqrt(posed.variance)[None, :]).reshape(M, 3, r)
    ref = np.einsum("mak,kl,mbl->mab", Q, Minv, Q)
    prior = np.einsum("mak,mak->m", Q, Q).max()
    assert got.shape == (M, 3, 3)
    assert np.max(np.abs(got - ref)) < 1e-9 * prior
    assert np.max(np.abs(got - np.transpose(got, (0, 2, 1)))) < 1e-12 * prior     # symmetric
    # observed vertices are more certain than the prior, unobserved ones never less certain than observed neighbours' floor
    tr_post, tr_prior = np.trace(got, axis1=1, axis2=2), np.einsum("mak,mak->m", Q, Q)
    assert np.all(tr_post <= tr_prior * (1 + 1e-12)) and np.all(tr_post > 0)
