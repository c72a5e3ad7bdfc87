"""K1 parity: the CUDA E-step through the C ABI against the CPU oracle on the same seeded inputs.
Floating point: tolerance 1e-11 relative to the largest entry (north-star bar is 1e-6 on the final
coefficients / vertices; the kernel itself is held far tighter)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-11


def _cloud(n, seed, scale=50.0):
    return np.random.default_rng(seed).normal(scale=scale, size=(n, 3))


@pytest.mark.parametrize("M,N", [(1, 1), (1, 300), (300, 1), (100, 100), (257, 513), (1000, 777), (2049, 4097)])
@pytest.mark.parametrize("w", [0.0, 0.1])
def test_cpd_estep_matches_oracle(ctx, oracle, M, N, w):
    from gingr_b200 import api
    fit, tgt = _cloud(M, M), _cloud(N, N + 1)
    sigma2 = 400.0
    target = api.Target(ctx, tgt)
    P1, Pt1, PX = api.cpd_estep(ctx, target, fit, sigma2, w)
    if M * N <= 4_000_000:
        P = oracle.cpd_P(fit, tgt, sigma2, w)
        r1, rt1, rx = oracle.P_reductions(P, tgt)
    else:
        r1, rt1, rx = oracle.cpd_estep(fit, tgt, sigma2, w)
    assert rel_err(P1, r1) < TOL
    assert rel_err(Pt1, rt1) < TOL
    assert rel_err(PX, rx) < TOL
    target.close()


def test_cpd_estep_democpd_dynamic_range(ctx, oracle):
    """DemoCPD's sigma2 = 1 on femur-scale coordinates: column sums down to 1e-150."""
    from gingr_b200 import api
    rng = np.random.default_rng(12)
    fit = rng.uniform(-200, 200, size=(100, 3))
    tgt = fit[rng.permutation(100)] + rng.normal(scale=8.0, size=(100, 3))
    target = api.Target(ctx, tgt)
    P1, Pt1, PX = api.cpd_estep(ctx, target, fit, 1.0, 0.0)
    P = oracle.cpd_P(fit, tgt, 1.0, 0.0)
    r1, rt1, rx = oracle.P_reductions(P, tgt)
    assert np.all(np.isfinite(P1)) and np.all(np.isfinite(PX))
    # entry-wise relative check: values span 100+ orders of magnitude
    np.testing.assert_allclose(P1, r1, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(Pt1, rt1, rtol=1e-12)
    np.testing.assert_allclose(PX, rx, rtol=1e-9, atol=1e-300)


def test_cpd_estep_underflow_gives_same_failure_as_reference(ctx, oracle):
    """No log-sum-exp stabilisation, like the reference: a column whose kernel values all underflow gives
    0/0 = NaN with w = 0 (-> ModelFlexibilityError downstream)."""
    from gingr_b200 import api
    fit = np.array([[0.0, 0, 0], [1, 0, 0]])
    tgt = np.array([[0.5, 0, 0], [1000.0, 0, 0]])
    target = api.Target(ctx, tgt)
    P1, Pt1, PX = api.cpd_estep(ctx, target, fit, 0.01, 0.0)
    P = oracle.cpd_P(fit, tgt, 0.01, 0.0)
    r1, rt1, rx = oracle.P_reductions(P, tgt)
    assert np.isnan(rt1[1]) and np.isnan(Pt1[1])
    assert np.array_equal(np.isnan(P1), np.isnan(r1))


def test_gradual_underflow_band(ctx, oracle):
    """Kernel values in the subnormal range (2^-1074 .. 2^-1022) are kept, not flushed."""
    from gingr_b200 import api
    # exp(-d2/2) with d2/2 in [708.4, 744]: subnormal doubles
    d = np.sqrt(2 * np.linspace(700.0, 744.0, 45))
    tgt = np.stack([d, np.zeros_like(d), np.zeros_like(d)], axis=1)
    fit = np.zeros((1, 3))
    target = api.Target(ctx, tgt)
    # w > 0 keeps the denominators finite; sigma2 = 1
    P1, Pt1, PX = api.cpd_estep(ctx, target, fit, 1.0, 0.5)
    r1, rt1, rx = oracle.cpd_estep(fit, tgt, 1.0, 0.5)
    np.testing.assert_allclose(Pt1, rt1, rtol=1e-9, atol=1e-320)
    assert np.count_nonzero(Pt1) == np.count_nonzero(rt1)


def test_column_sums_to_one_property_full_size(ctx):
    """Size-independent property at a large size: with w = 0, Pt1_j = 1 and sum(P1) = N (CPD.scala:69-74)."""
    from gingr_b200 import api, synthetic
    fit = synthetic.fibonacci_sphere(20000)
    tgt = synthetic.make_target(synthetic.fibonacci_sphere(50000), 0)
    target = api.Target(ctx, tgt)
    P1, Pt1, PX = api.cpd_estep(ctx, target, fit, 25.0, 0.0)
    np.testing.assert_allclose(Pt1, 1.0, rtol=0, atol=1e-12)
    assert abs(P1.sum() - 50000) < 1e-6
    # PX/P1 is a convex combination of target points: inside the target's bounding box
    td = PX / P1[:, None]
    assert np.all(td >= tgt.min(0) - 1e-9) and np.all(td <= tgt.max(0) + 1e-9)


def test_estep_is_deterministic(ctx):
    from gingr_b200 import api
    fit, tgt = _cloud(700, 1), _cloud(1900, 2)
    target = api.Target(ctx, tgt)
    a = api.cpd_estep(ctx, target, fit, 300.0, 0.1)
    b = api.cpd_estep(ctx, target, fit, 300.0, 0.1)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("M,N", [(12, 17), (300, 500)])
def test_bcpd_estep_matches_oracle(ctx, oracle, M, N):
    from gingr_b200 import api
    y, x = _cloud(M, 13, 5.0), _cloud(N, 14, 5.0)
    rng = np.random.default_rng(15)
    sig = rng.uniform(0.1, 1.0, M)
    al = rng.uniform(0.5, 1.5, M) / M
    s2, s, w = 9.0, 1.1, 0.2
    target = api.Target(ctx, x)
    nu, nup, nhat, xhat = api.bcpd_estep(ctx, target, y, sig, al, s2, s, w)
    rnu, rnup, rnhat, rxhat = oracle.bcpd_estep(y, x, sig, al, s2, s, w)
    assert rel_err(nu, rnu) < TOL and rel_err(nup, rnup) < TOL
    assert abs(nhat - rnhat) < TOL * abs(rnhat)
    assert rel_err(xhat, rxhat) < 1e-10


def test_initial_sigma2(ctx, oracle):
    from gingr_b200 import api
    a, b = _cloud(230, 10) + 40.0, _cloud(290, 11) - 25.0
    target = api.Target(ctx, b)
    got = api.cpd_initial_sigma2(ctx, target, a)
    ref = oracle.cpd_initial_sigma2(a, b)
    assert abs(got - ref) < 1e-12 * ref


@pytest.mark.gpu
def test_gaussian_kernel_accuracy_over_the_whole_range(ctx):
    """The device 2^x (table + degree-5 tail, exp2_poly.cuh) against numpy's exp, isolated through a 1 x N E-step:
    Pt1_j = K_j / (K_j + c)  =>  K_j = c Pt1_j / (1 - Pt1_j).  Distances are multiples of 1/64, so d^2 is exact.
    Both sides round the scaled argument x = d^2 / (2 sigma2) once (the reference divides, CPD.scala:55-57; the
    kernel folds log2(e) / (2 sigma2) into one constant), so the error bound grows like |x| eps."""
    from gingr_b200 import api
    n = 13000
    sigma2, w = 1.0, 0.999
    d = np.arange(n) / 64.0 * 0.18                       # x = d^2 / 2 up to 668 -> K down to 1e-290
    d = np.round(d * 64) / 64
    target = np.zeros((n, 3))
    target[:, 0] = d
    fit = np.zeros((1, 3))
    tgt = api.Target(ctx, target)
    P1, Pt1, PX = api.cpd_estep(ctx, tgt, fit, sigma2, w)
    c = w / (1 - w) * (2 * np.pi * sigma2) ** 1.5 * 1 / n
    K = c * Pt1 / (1 - Pt1)
    x = (d * d) / (2 * sigma2)
    Kref = np.exp(-x)
    rel = np.abs(K - Kref) / Kref
    bound = (2.0 * x + 8.0) * 1.2e-16
    assert np.all(rel < bound), float(np.max(rel / bound))
    tgt.close()
