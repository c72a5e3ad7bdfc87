"""GPMM construction on the device (gingr_b200/csrc/gpmm.cuh) against the oracle's literal
approximateGPCholesky: rank equal, variances 1e-9 relative, covariance basis diag(variance) basis^T 1e-9 relative,
orthonormal columns; at scale the trace criterion itself."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(n, seed=0):
    from gingr_b200 import synthetic
    v, t = synthetic.sphere_mesh(n)
    return v + 0.5 * np.random.default_rng(seed).normal(size=v.shape), t


@pytest.mark.parametrize("M,sig,sc,tol,cap", [(60, [70.0, 25.0], [50.0, 10.0], 0.01, 0), (150, [70.0], [50.0], 0.01, 0),
                                              (150, [60.0], [30.0], 0.12, 0), (150, [60.0], [30.0], 0.16, 0),
                                              (120, [40.0], [20.0], 0.001, 50), (90, [500.0], [5.0], 0.3, 0)])
def test_gpmm_matches_literal_construction(ctx, oracle, M, sig, sc, tol, cap):
    from gingr_b200 import api
    ref, tri = _ref(M)
    om, L = oracle.approximate_gp_cholesky(ref, sig, sc, tol, cap or None)
    dm = api.Model.gaussianMixture(ctx, ref, tri, sig, sc, tol, cap)
    assert dm.rank == om.rank
    gref, mean, basis, var = dm.download()
    assert np.array_equal(gref, ref) and np.all(mean == 0.0)
    assert np.max(np.abs(var - om.variance)) <= 1e-9 * om.variance[0]
    assert np.all(np.diff(var) <= 0.0)
    assert np.max(np.abs(basis.T @ basis - np.eye(dm.rank))) < 1e-10
    lit = L @ L.T
    assert np.max(np.abs((basis * var) @ basis.T - lit)) <= 1e-9 * np.max(np.abs(lit))
    dm.close()


def test_gpmm_model_registers_like_the_uploaded_one(ctx, oracle):
    """A registration on the device-built model equals (1e-6) one on the same model uploaded from the host."""
    from gingr_b200 import api, synthetic
    ref, tri = _ref(300)
    tv, tt = synthetic.sphere_mesh(350)
    target = synthetic.make_target(tv, 0)
    dm = api.Model.gaussianMixture(ctx, ref, tri, [70.0], [50.0], 0.01)
    gref, mean, basis, var = dm.download()
    um = api.Model(ctx, gref, mean, basis, var, tri)
    dt = api.Target(ctx, target, tt)
    outs = []
    for model in (dm, um):
        reg = api.CpdRegistration(ctx, model, dt, api.CpdConfiguration(maxIterations=12, w=0.05))
        outs.append(reg.run(reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)))
        reg.close()
    diag = float(np.linalg.norm(ref.max(0) - ref.min(0)))
    assert np.max(np.abs(outs[0].fit - outs[1].fit)) < 1e-6 * diag
    assert abs(outs[0].sigma2 - outs[1].sigma2) <= 1e-6 * outs[1].sigma2
    # and the oracle agrees on the downloaded model
    om = oracle.Gpmm(gref, mean, np.ascontiguousarray(basis), var, tri)
    oalgo = oracle.CpdAlgorithm(oracle.CpdConfig(max_iterations=12, w=0.05), literal=False)
    ofinal = oracle.run(oalgo, oracle.initial_state(om, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
    assert np.max(np.abs(outs[0].fit - ofinal.fit)) < 1e-6 * diag
    dm.close(); um.close(); dt.close()


def test_gpmm_at_scale_meets_the_trace_criterion(ctx):
    from gingr_b200 import api
    M = 20000
    ref, tri = _ref(M)
    sig, sc, tol = 70.0, 50.0, 0.01
    dm = api.Model.gaussianMixture(ctx, ref, tri, [sig], [sc], tol)
    _, mean, basis, var = dm.download()
    tr = 3.0 * M * sc
    assert tr - var.sum() <= tol * tr * (1 + 1e-9)
    assert 30 <= dm.rank <= 3000
    assert np.max(np.abs(basis.T @ basis - np.eye(dm.rank))) < 1e-9
    # the diagonal of the low-rank covariance never exceeds the kernel's
    diag_cov = np.einsum("ij,j,ij->i", basis, var, basis)
    assert np.all(diag_cov <= sc * (1 + 1e-9)) and diag_cov.min() > 0.5 * sc
    dm.close()


def test_gpmm_argument_errors(ctx):
    from gingr_b200 import api
    ref, tri = _ref(30)
    with pytest.raises(api.GingrError):
        api.Model.gaussianMixture(ctx, ref, tri, [0.0], [1.0])
    with pytest.raises(api.GingrError):
        api.Model.gaussianMixture(ctx, ref, tri, [1.0] * 9, [1.0] * 9)
    with pytest.raises(api.GingrError):
        api.Model.gaussianMixture(ctx, ref, tri, [10.0], [1.0], relativeTolerance=1.5)      # nothing left


def test_model_cache_builds_once_then_reloads_identically(ctx, tmp_path):
    """DataSetLoader.model (DemoDatasetLoader.scala:40-53): first call builds on the device and writes the file, the
    second reads it; both models are bit-identical on the device, and a corrupt file is rebuilt."""
    import os
    from gingr_b200 import io
    ref, tri = _ref(120)
    a = io.load_or_create_gauss_model(ctx, str(tmp_path), "blob", ref, tri, scaling=50.0, sigma=70.0, decimate=120)
    path = os.path.join(str(tmp_path), "blob_dec-120_Gauss_50.0_70.0.h5.json")
    assert os.path.isfile(path)
    stamp = os.stat(path).st_mtime_ns
    b = io.load_or_create_gauss_model(ctx, str(tmp_path), "blob", ref, tri, scaling=50.0, sigma=70.0, decimate=120)
    assert os.stat(path).st_mtime_ns == stamp and b.rank == a.rank and b.T == a.T
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x, y)
    with open(path, "w") as f:
        f.write("{")
    c = io.load_or_create_gauss_model(ctx, str(tmp_path), "blob", ref, tri, scaling=50.0, sigma=70.0, decimate=120)
    assert c.rank == a.rank and io.read_statistical_model(path)[4].shape == (a.rank,)
    a.close(); b.close(); c.close()


def test_automatic_gaussian_is_the_two_kernel_mixture(ctx):
    """AutomaticGaussian (GPMMHelper.scala:119-129) = GaussianMixture((d/4, d/8), (d/8, d/16)), d = largest point distance."""
    from gingr_b200 import api
    ref, tri = _ref(200)
    d = max(float(np.linalg.norm(a - b)) for a in ref for b in ref)
    assert abs(api.maximum_point_distance(ref) - d) <= 1e-13 * d
    d = api.maximum_point_distance(ref)
    a = api.Model.automaticGaussian(ctx, ref, tri, 0.05)
    b = api.Model.gaussianMixture(ctx, ref, tri, [d / 4.0, d / 8.0], [d / 8.0, d / 16.0], 0.05)
    assert a.rank == b.rank and a.rank > 3
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x, y)
    a.close(); b.close()


def test_model_cache_with_the_automatic_mixture(ctx, tmp_path):
    import os
    from gingr_b200 import api, io
    ref, tri = _ref(90)
    a = io.load_or_create_model(ctx, str(tmp_path), "blob", ref, tri, api.GaussMixKernel(), relativeTolerance=0.05)
    assert os.path.isfile(os.path.join(str(tmp_path), "blob_dec-full_GaussMix_.h5.json"))
    b = io.load_or_create_model(ctx, str(tmp_path), "blob", ref, tri, api.GaussMixKernel(), relativeTolerance=0.05)
    assert a.rank == b.rank
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x, y)
    with pytest.raises(NotImplementedError):
        api.SimpleTriangleModels3D.create(ctx, ref, tri, object())
    a.close(); b.close()
