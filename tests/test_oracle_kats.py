"""Known-answer tests pinning the CPU oracle (CPU only).  The reference ships no golden vectors
(src/test/scala/DummyTest.scala.scala:1-3 -- PARITY UNPINNED), so these follow from the reference's
formulas alone."""
import numpy as np
import pytest

from conftest import rel_err


def _cloud(n, seed, scale=50.0):
    return np.random.default_rng(seed).normal(scale=scale, size=(n, 3))


def test_P_columns_sum_to_one_without_outliers(oracle):
    # w = 0  =>  c = 0, Pt1_j = 1, sum_i P1_i = N        (CPD.scala:69-74)
    fit, tgt = _cloud(37, 0), _cloud(53, 1)
    P = oracle.cpd_P(fit, tgt, 400.0, 0.0)
    P1, Pt1, PX = oracle.P_reductions(P, tgt)
    np.testing.assert_allclose(Pt1, 1.0, rtol=0, atol=1e-14)
    assert abs(P1.sum() - 53) < 1e-11


def test_outlier_term_matches_formula(oracle):
    fit, tgt = _cloud(20, 2), _cloud(30, 3)
    s2, w = 250.0, 0.3
    P = oracle.cpd_P(fit, tgt, s2, w)
    K = np.exp(-((tgt[None, :, :] - fit[:, None, :]) ** 2).sum(-1) / (2 * s2))
    c = w / (1 - w) * (2 * np.pi * s2) ** 1.5 * 20 / 30
    np.testing.assert_allclose(P, K / (K.sum(0, keepdims=True) + c), rtol=1e-13)


@pytest.mark.parametrize("w", [0.0, 0.1])
def test_streaming_estep_equals_literal(oracle, w):
    fit, tgt = _cloud(65, 4), _cloud(129, 5)
    P = oracle.cpd_P(fit, tgt, 300.0, w)
    P1, Pt1, PX = oracle.P_reductions(P, tgt)
    for fast in (False, True):
        q1, qt1, qx = oracle.cpd_estep(fit, tgt, 300.0, w, fast=fast)
        assert rel_err(q1, P1) < 1e-13 and rel_err(qt1, Pt1) < 1e-13 and rel_err(qx, PX) < 1e-13


def test_correspondence_is_PX_over_P1(oracle):
    fit, tgt = _cloud(40, 6), _cloud(60, 7)
    P = oracle.cpd_P(fit, tgt, 500.0, 0.05)
    P1, _, PX = oracle.P_reductions(P, tgt)
    td = oracle.cpd_correspondence(P, fit, tgt)
    assert rel_err(td, PX / P1[:, None]) < 1e-13


def test_sigma2_update_equals_brute_force(oracle):
    # CPD.scala:142-145 == sum_ij P_ij |x_j - ty_i|^2 / (3 Np)
    fit, tgt = _cloud(31, 8), _cloud(47, 9)
    P = oracle.cpd_P(fit, tgt, 700.0, 0.1)
    P1, Pt1, PX = oracle.P_reductions(P, tgt)
    s2 = oracle.cpd_sigma2_update(P1, Pt1, PX, tgt, fit)
    d2 = ((tgt[None, :, :] - fit[:, None, :]) ** 2).sum(-1)
    assert abs(s2 - (P * d2).sum() / (3 * P1.sum())) < 1e-9 * s2


def test_initial_sigma2(oracle):
    a, b = _cloud(23, 10), _cloud(29, 11)
    ref = ((b[None] - a[:, None]) ** 2).sum() / (3 * 23 * 29)
    assert abs(oracle.cpd_initial_sigma2(a, b) - ref) < 1e-12 * ref


def test_femur_scale_dynamic_range(oracle):
    # DemoCPD setting sigma2 = 1 on coordinates of +-200: column sums far below FP32 range, still finite
    rng = np.random.default_rng(12)
    fit = rng.uniform(-200, 200, size=(100, 3))
    tgt = fit[rng.permutation(100)] + rng.normal(scale=8.0, size=(100, 3))
    P = oracle.cpd_P(fit, tgt, 1.0, 0.0)
    assert np.all(np.isfinite(P))
    K = np.exp(-((tgt[None] - fit[:, None]) ** 2).sum(-1) / 2.0)
    assert K.sum(0).min() < 1e-45  # below FP32 range (1.4e-45)


def test_bcpd_P_matches_dense_formula(oracle):
    y, x = _cloud(12, 13, 5.0), _cloud(17, 14, 5.0)
    sig = np.random.default_rng(15).uniform(0.1, 1.0, 12)
    al = np.full(12, 1 / 12)
    s2, s, w = 9.0, 1.1, 0.2
    P = oracle.bcpd_P(y, x, sig, al, s2, s, w)
    d2 = ((x[None] - y[:, None]) ** 2).sum(-1)
    phi = (2 * np.pi * s2) ** -1.5 * np.exp(-d2 / (2 * s2)) * np.exp(-s / (2 * s2) * 3 * sig)[:, None] * al[:, None]
    Pinit = phi * (1 - w)
    den = Pinit.sum(0) * (1 - w) + w / 17
    np.testing.assert_allclose(P, Pinit / den, rtol=1e-12)


def test_nearest_vertex_lowest_index_on_ties(oracle):
    pts = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [2, 0, 0.0]])
    q = np.array([[1, 0, 0], [2, 0.1, 0], [0.2, 1.9, 0.0]])
    idx, d2 = oracle.nearest_vertex(q, pts)
    assert idx.tolist() == [0, 1, 2]
    assert d2[0] == 1.0


def test_nearest_vertex_matches_numpy(oracle):
    q, p = _cloud(200, 16), _cloud(333, 17)
    idx, d2 = oracle.nearest_vertex(q, p)
    D = ((q[:, None] - p[None]) ** 2).sum(-1)
    assert np.array_equal(idx, D.argmin(1))


def test_closest_on_triangle_regions(oracle):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.0]])
    tri = np.array([[0, 1, 2]], dtype=np.int32)
    q = np.array([[0.25, 0.25, 1.0], [-1, -1, 0], [2, 0, 0], [0, 3, 0], [0.5, -1, 0], [1, 1, 0], [-1, 0.5, 2]])
    cp, d2, ti = oracle.closest_on_surface(q, v, tri)
    expect = np.array([[0.25, 0.25, 0], [0, 0, 0], [1, 0, 0], [0, 1, 0], [0.5, 0, 0], [0.5, 0.5, 0], [0, 0.5, 0]])
    np.testing.assert_allclose(cp, expect, atol=1e-15)
    np.testing.assert_allclose(d2, ((q - expect) ** 2).sum(1), atol=1e-15)


def test_closest_on_surface_of_sphere_mesh(oracle):
    from gingr_b200 import synthetic
    v, tri = synthetic.sphere_mesh(300)
    q = synthetic.fibonacci_sphere(50, 130.0)
    cp, d2, ti = oracle.closest_on_surface(q, v, tri)
    # closest point lies on the returned triangle's plane and is no farther than any vertex
    assert np.all(np.sqrt(d2) <= np.sqrt(((q[:, None] - v[None]) ** 2).sum(-1)).min(1) + 1e-12)
    assert np.all(np.sqrt(d2) > 25.0)


def test_boundary_and_normals(oracle):
    from gingr_b200 import synthetic
    v, tri = synthetic.sphere_mesh(200)
    assert not oracle.boundary_vertices(200, tri).any()          # closed mesh
    n = oracle.vertex_normals(v, tri)
    assert np.all(np.einsum("ij,ij->i", n, v / 100.0) > 0.95)     # outward
    open_tri = tri[~np.any(tri == 0, axis=1)]                    # remove the fan around vertex 0
    b = oracle.boundary_vertices(200, open_tri)
    ring = np.unique(tri[np.any(tri == 0, axis=1)])
    assert set(np.nonzero(b)[0]) == set(ring) - {0}


def test_euler_round_trip(oracle):
    for ang in [(0.1, 0.2, 0.3), (-1.0, 0.5, 2.0), (0.0, 0.0, 0.0)]:
        R = oracle.euler_to_matrix(*ang)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-15)
        np.testing.assert_allclose(oracle.matrix_to_euler(R), ang, atol=1e-14)


def test_umeyama_recovers_planted_motion(oracle):
    X = _cloud(50, 18)
    R0 = oracle.euler_to_matrix(0.3, -0.2, 0.1)
    t0 = np.array([5.0, -3.0, 2.0])
    R, t, s = oracle.umeyama(X, X @ R0.T + t0, False)
    np.testing.assert_allclose(R, R0, atol=1e-12)
    np.testing.assert_allclose(t, t0, atol=1e-10)
    R, t, s = oracle.umeyama(X, 1.7 * (X @ R0.T) + t0, True)
    assert abs(s - 1.7) < 1e-12


def _small_model(oracle, M=60, r=12, seed=0):
    from gingr_b200 import synthetic
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, seed)
    return oracle.Gpmm(ref, mean, basis, var, tri)


def test_coefficients_of_instance_shrink(oracle):
    # orthonormal basis: coefficients(instance(a)) = a * l/(l + 1e-5)
    m = _small_model(oracle)
    a = np.random.default_rng(1).normal(size=m.rank)
    c = m.coefficients(m.instance(a))
    np.testing.assert_allclose(c, a * m.variance / (m.variance + 1e-5), rtol=1e-9)


def test_posterior_limits(oracle):
    m = _small_model(oracle)
    a = np.random.default_rng(2).normal(size=m.rank)
    obs = m.instance(a)
    pids = np.arange(m.M)
    big = np.eye(3)[None] * 1e12 * np.ones((m.M, 1, 1))
    c, _ = m.posterior_coefficients(pids, obs, big)
    assert np.abs(c).max() < 1e-6                                 # noise -> inf : prior mean
    tiny = np.eye(3)[None] * 1e-9 * np.ones((m.M, 1, 1))
    c, _ = m.posterior_coefficients(pids, obs, tiny)
    np.testing.assert_allclose(m.instance(c), obs, atol=1e-6)     # noise -> 0 : interpolates


def test_transform_commutes_with_instance(oracle):
    m = _small_model(oracle)
    a = np.random.default_rng(3).normal(size=m.rank)
    R, t = oracle.euler_to_matrix(0.2, 0.1, -0.3), np.array([1.0, 2.0, 3.0])
    np.testing.assert_allclose(m.transform(R, t).instance(a), m.instance(a) @ R.T + t, atol=1e-11)


def test_update_runs_and_reduces_distance(oracle):
    from gingr_b200 import synthetic
    m = _small_model(oracle, M=80, r=15)
    a_true = np.random.default_rng(4).normal(size=m.rank)
    target = m.instance(a_true)
    st = oracle.initial_state(m, target, global_transformation=oracle.NO_TRANSFORMS)
    algo = oracle.CpdAlgorithm(oracle.CpdConfig(max_iterations=15))
    out = oracle.run(algo, st)
    d0 = np.abs(st.fit - target).mean()
    d1 = np.abs(out.fit - target).mean()
    assert d1 < 0.2 * d0
    assert out.status in (oracle.STATUS_MAX_ITERATION, oracle.STATUS_CONVERGED)
    assert out.iteration >= 1


def test_update_literal_equals_streaming(oracle):
    m = _small_model(oracle, M=50, r=10)
    target = m.instance(np.random.default_rng(5).normal(size=m.rank))
    st = oracle.initial_state(m, target, global_transformation=oracle.RIGID_TRANSFORMS)
    a = oracle.CpdAlgorithm(oracle.CpdConfig(), literal=True)
    b = oracle.CpdAlgorithm(oracle.CpdConfig(), literal=False)
    sa = oracle.propose(a, a.initialize(st))
    sb = oracle.propose(b, b.initialize(st))
    assert rel_err(sa.params.shape, sb.params.shape) < 1e-10
    assert rel_err(sa.fit, sb.fit) < 1e-12
    assert abs(sa.sigma2 - sb.sigma2) < 1e-10 * sa.sigma2


def test_icp_update_runs(oracle):
    from gingr_b200 import synthetic
    m = _small_model(oracle, M=120, r=15)
    tv, tt = synthetic.sphere_mesh(150)
    target = synthetic.make_target(tv, 0, t=(1.0, 1.0, 1.0), euler=(0.01, 0.01, 0.01))
    st = oracle.initial_state(m, target, tt, global_transformation=oracle.NO_TRANSFORMS)
    algo = oracle.IcpAlgorithm(oracle.IcpConfig(max_iterations=5, initial_sigma=1.0, end_sigma=1.0))
    out = oracle.run(algo, st)
    assert out.iteration == 4 and out.status == oracle.STATUS_MAX_ITERATION
    assert np.all(np.isfinite(out.fit))


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for Philox4x32-10 (kat_vectors of the Random123 distribution)."""
    out = oracle.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(v[0]) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    out = oracle.philox4x32_10([f], [f], [f], [f], f, f)
    assert [int(v[0]) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = oracle.philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)
    assert [int(v[0]) for v in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_standard_normals_moments_and_counter_separation(oracle):
    z = oracle.standard_normals(200001, 99, 3)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01 and np.all(np.isfinite(z))
    assert abs(np.mean(z ** 4) - 3.0) < 0.1
    assert not np.array_equal(z[:100], oracle.standard_normals(100, 99, 4))
    assert not np.array_equal(z[:100], oracle.standard_normals(100, 98, 3))
    assert np.array_equal(z[:101], oracle.standard_normals(101, 99, 3))


def test_posterior_sample_has_covariance_minv(oracle):
    """c + L^-T z with Mx = L L^T has covariance L^-T L^-1 = Minv, the covariance of scalismo's posterior
    (SURVEY.md A3): checked exactly through the unit vectors z = e_k."""
    rng = np.random.default_rng(0)
    A = rng.normal(size=(40, 12))
    Mx = A.T @ A + np.eye(12)
    L = np.linalg.cholesky(Mx)
    S = np.linalg.solve(L.T, np.eye(12))
    np.testing.assert_allclose(S @ S.T, oracle.breeze_pinv(Mx), rtol=1e-10, atol=1e-12)


def test_posterior_equals_dense_gaussian_conditioning(oracle):
    """Independent of the low-rank algebra the oracle restates [A3]: the GP the model describes has covariance
    K = Phi diag(lambda) Phi^T over all 3M coordinates; conditioning it on noisy observations of some points with the
    textbook dense formula  mean' = mu + K_xo (K_oo + Sigma)^-1 (y - mu_o),  K' = K - K_xo (K_oo + Sigma)^-1 K_ox  must give
    the mean the oracle's coefficient vector produces and the covariance D Minv D spans (posterior_model's SVD form)."""
    rng = np.random.default_rng(11)
    m = _small_model(oracle, M=40, r=9, seed=3)
    pids = rng.choice(m.M, size=17, replace=False)
    A = rng.normal(size=(len(pids), 3, 3))
    cov = A @ np.transpose(A, (0, 2, 1)) + 0.05 * np.eye(3)[None]             # full 3x3 noise per observation
    y_pts = m.ref[pids] + m.mean.reshape(-1, 3)[pids] + rng.normal(size=(len(pids), 3)) * 3.0
    c, Minv = m.posterior_coefficients(pids, y_pts, cov)
    D = np.sqrt(m.variance)
    low_rank_mean = m.mean + m.basis @ (D * c)
    K = (m.basis * m.variance) @ m.basis.T
    rows = (3 * pids[:, None] + np.arange(3)[None, :]).ravel()
    S = np.zeros((len(rows), len(rows)))
    for k in range(len(pids)):
        S[3 * k:3 * k + 3, 3 * k:3 * k + 3] = cov[k]
    resid = (y_pts - m.ref[pids]).ravel() - m.mean[rows]
    G = np.linalg.solve(K[np.ix_(rows, rows)] + S, np.eye(len(rows)))
    dense_mean = m.mean + K[:, rows] @ (G @ resid)
    assert np.max(np.abs(low_rank_mean - dense_mean)) < 1e-9 * np.max(np.abs(dense_mean))
    dense_cov = K - K[:, rows] @ G @ K[rows, :]
    low_rank_cov = m.basis @ (D[:, None] * Minv * D[None, :]) @ m.basis.T
    assert np.max(np.abs(low_rank_cov - dense_cov)) < 1e-9 * np.max(np.abs(K))


def test_closest_on_triangle_against_constrained_least_squares(oracle):
    """Independent of the region walk: the closest point of a triangle is the minimiser of |a + u e1 + v e2 - q|^2 over
    u, v >= 0, u + v <= 1, a convex QP solved here by enumerating its KKT candidates in closed form (interior stationary
    point, the three edge projections clamped to [0, 1]) -- random, skinny and near-degenerate triangles."""
    rng = np.random.default_rng(21)
    worst = 0.0
    for trial in range(300):
        a, b, c = rng.normal(size=(3, 3)) * rng.choice([1.0, 1e-3, 50.0])
        if trial % 5 == 0:
            c = a + (b - a) * rng.uniform(0.2, 0.8) + 1e-4 * rng.normal(size=3)      # sliver
        q = rng.normal(size=3) * rng.choice([0.5, 5.0, 100.0])
        cands = []
        e1, e2 = b - a, c - a
        Gm = np.array([[e1 @ e1, e1 @ e2], [e1 @ e2, e2 @ e2]])
        if abs(np.linalg.det(Gm)) > 1e-300:
            u, v = np.linalg.solve(Gm, np.array([e1 @ (q - a), e2 @ (q - a)]))
            if u >= 0 and v >= 0 and u + v <= 1:
                cands.append(a + u * e1 + v * e2)
        for p0, p1 in ((a, b), (b, c), (c, a)):
            d = p1 - p0
            s = np.clip(d @ (q - p0) / (d @ d), 0.0, 1.0)
            cands.append(p0 + s * d)
        best = min(np.linalg.norm(q - x) for x in cands)
        cp, d2, _ = oracle.closest_on_surface(q[None], np.array([a, b, c]), np.array([[0, 1, 2]], dtype=np.int32))
        got = np.linalg.norm(q - cp[0])
        scale = max(best, np.linalg.norm(b - a), 1e-12)
        worst = max(worst, abs(got - best) / scale)
        assert abs(np.sqrt(d2[0]) - got) <= 1e-12 * scale
    assert worst < 1e-9, worst


def test_umeyama_is_the_least_squares_optimum_on_noisy_pairs(oracle):
    """Independent of the SVD recipe: on NOISY correspondences the rigid result must be the minimiser of
    sum |R x + t - y|^2 over rotations (checked against scipy's Kabsch implementation and by perturbation), and the
    similarity scale the minimiser of the same residual in s for the returned R."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(8)
    X = _cloud(80, 5)
    R0 = oracle.euler_to_matrix(-0.4, 0.25, 0.6)
    Y = 1.3 * (X @ R0.T) + np.array([2.0, -1.0, 0.5]) + 0.3 * rng.normal(size=X.shape)

    def resid(R, t, s):
        return float(np.sum((s * (X @ R.T) + t - Y) ** 2))
    R, t, s = oracle.umeyama(X, Y, False)
    assert s == 1.0 and abs(np.linalg.det(R) - 1.0) < 1e-12
    Rk, _ = Rotation.align_vectors(Y - Y.mean(0), X - X.mean(0))
    np.testing.assert_allclose(R, Rk.as_matrix(), atol=1e-9)
    np.testing.assert_allclose(t, Y.mean(0) - R @ X.mean(0), atol=1e-12)
    base = resid(R, t, 1.0)
    for _ in range(20):
        dR = Rotation.from_rotvec(1e-3 * rng.normal(size=3)).as_matrix()
        assert resid(dR @ R, Y.mean(0) - dR @ R @ X.mean(0), 1.0) >= base - 1e-9
    R, t, s = oracle.umeyama(X, Y, True)
    Xc, Yc = X - X.mean(0), Y - Y.mean(0)
    s_opt = np.sum(Yc * (Xc @ R.T)) / np.sum(Xc * Xc)                # d/ds of the residual = 0
    assert abs(s - s_opt) < 1e-12 and abs(s - 1.3) < 0.05
    np.testing.assert_allclose(t, Y.mean(0) - s * (R @ X.mean(0)), atol=1e-12)
    # reflection case: the data prefer an improper map; the result must still be a rotation
    Yr = X * np.array([1.0, 1.0, -1.0]) + 0.01 * rng.normal(size=X.shape)
    Rr, _, _ = oracle.umeyama(X, Yr, False)
    assert abs(np.linalg.det(Rr) - 1.0) < 1e-12


def test_line_mesh_nearest_against_linear_solves(oracle):
    """Independent of Moeller-Trumbore: for every triangle solve  a + u (b - a) + v (c - a) = p + s d  as a 3x3 linear
    system; an intersection is 0 <= u, v, u + v <= 1 (s of either sign: the line is infinite), the answer the nearest
    one that is not p itself.  Queries off the mesh, so the 'drop points == p' rule plays no role here."""
    from gingr_b200 import synthetic
    rng = np.random.default_rng(31)
    v, tri = synthetic.sphere_mesh(120)
    p = rng.normal(size=(40, 3)) * 60.0                         # inside, outside and far outside the radius-100 sphere
    p[::4] *= 4.0
    d = rng.normal(size=p.shape)
    dist, hit = oracle.line_mesh_nearest(p, d, v, tri)
    for i in range(len(p)):
        best, bh = np.inf, p[i]
        for t in tri:
            a, b, c = v[t]
            A = np.column_stack([b - a, c - a, -d[i]])
            if abs(np.linalg.det(A)) < 1e-12:
                continue
            u, w, s = np.linalg.solve(A, p[i] - a)
            if u >= 0 and w >= 0 and u + w <= 1:
                x = p[i] + s * d[i]
                if np.linalg.norm(x - p[i]) < best:
                    best, bh = np.linalg.norm(x - p[i]), x
        if np.isinf(best):
            assert np.isinf(dist[i])
        else:
            assert abs(dist[i] - best) <= 1e-9 * max(best, 1.0), (i, dist[i], best)
            assert np.linalg.norm(hit[i] - bh) <= 1e-8 * max(best, 1.0)
    inside = np.linalg.norm(p, axis=1) < 90.0
    assert inside.any() and np.all(np.isfinite(dist[inside]))      # a line through an interior point always hits a closed mesh
    assert np.isinf(dist).any()                                    # and some far lines miss it


@pytest.mark.parametrize("sigma2,w", [(25.0, 0.1), (1.0, 0.0), (400.0, 0.5)])
def test_estep_reductions_against_extended_precision(oracle, sigma2, w):
    """Independent evaluation of P1, P^T 1 and P X in 80-bit arithmetic (numpy longdouble) straight from the definition
    (CPD.scala:54-75), both oracle variants: the FP64 restatement is within 1e-12 relative of it."""
    if np.finfo(np.longdouble).eps >= np.finfo(np.float64).eps:
        pytest.skip("no extended precision on this platform")
    rng = np.random.default_rng(17)
    M, N = 70, 90
    fit = rng.normal(size=(M, 3)) * 20.0
    tgt = fit[rng.integers(0, M, N)] + rng.normal(size=(N, 3)) * np.sqrt(sigma2)
    L = np.longdouble
    f, t = fit.astype(L), tgt.astype(L)
    d2 = ((t[None, :, :] - f[:, None, :]) ** 2).sum(-1)
    K = np.exp(-d2 / (L(2) * L(sigma2)))
    c = (L(2) * L(np.pi) * L(sigma2)) ** (L(3) / L(2)) * L(w) / (L(1) - L(w)) * L(M) / L(N)
    P = K / (K.sum(0) + c)[None, :]
    P1, Pt1, PX = P.sum(1), P.sum(0), P @ t
    for fast in (False, True):
        g1, gt1, gx = oracle.cpd_estep(fit, tgt, sigma2, w, fast=fast)
        assert np.max(np.abs(g1 - P1) / P1.max()) < 1e-12
        assert np.max(np.abs(gt1 - Pt1)) < 1e-12
        assert np.max(np.abs(gx - PX)) < 1e-12 * float(np.max(np.abs(PX)))
