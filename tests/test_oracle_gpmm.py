"""CPU known-answer tests of the GPMM construction restated in oracle/ (no GPU): the scalar-factorisation shortcut that
gingr_b200/csrc/gpmm.cuh takes for K = Ks (x) I3 gives the rank and covariance of the literal 3M x 3M pivoted Cholesky
(scalismo approximateGPCholesky), incl. ranks that are not a multiple of three."""
import numpy as np
import pytest


def _ref(n, seed=0):
    from gingr_b200 import synthetic
    return synthetic.fibonacci_sphere(n) + 0.5 * np.random.default_rng(seed).normal(size=(n, 3))


def _structured_cov(oracle, ref, sig, sc, rel_tol, max_rank=None):
    rank, Ls = oracle.approximate_gp_cholesky_structured(ref, sig, sc, rel_tol, max_rank)
    M = len(ref)
    cov = np.zeros((3 * M, 3 * M))
    for d in range(3):
        kd = (rank + 2) // 3 if (rank % 3 == 0 or d < rank % 3) else rank // 3
        cov[d::3, d::3] = Ls[:, :kd] @ Ls[:, :kd].T
    return rank, cov


@pytest.mark.parametrize("rel_tol,max_rank", [(0.01, None), (0.05, None), (0.2, None), (0.001, 40), (0.03, None), (0.0, 31)])
def test_structured_factorisation_equals_literal(oracle, rel_tol, max_rank):
    ref = _ref(60)
    sig, sc = [70.0, 25.0], [50.0, 10.0]
    model, L = oracle.approximate_gp_cholesky(ref, sig, sc, rel_tol, max_rank)
    rank, cov = _structured_cov(oracle, ref, sig, sc, rel_tol, max_rank)
    assert rank == L.shape[1] == model.rank
    lit = L @ L.T
    assert np.max(np.abs(cov - lit)) < 1e-10 * np.max(np.abs(lit))
    # KL basis of the literal model: orthonormal, covariance preserved, variances sorted
    B = model.basis
    assert np.max(np.abs(B.T @ B - np.eye(rank))) < 1e-8
    assert np.max(np.abs((B * model.variance) @ B.T - lit)) < 1e-9 * np.max(np.abs(lit))
    assert np.all(np.diff(model.variance) <= 1e-12 * model.variance[0])


def test_ranks_not_multiple_of_three_occur(oracle):
    ref = _ref(50, 1)
    ranks = set()
    for tol in np.linspace(0.02, 0.5, 25):
        rank, _ = oracle.approximate_gp_cholesky_structured(ref, [60.0], [30.0], tol)
        model, L = oracle.approximate_gp_cholesky(ref, [60.0], [30.0], tol)
        assert rank == L.shape[1]
        ranks.add(rank % 3)
    assert ranks == {0, 1, 2}


def test_trace_criterion_and_kernel(oracle):
    ref = _ref(40, 2)
    Ks = oracle.gaussian_mixture_kernel_matrix(ref, [50.0], [20.0])
    assert abs(Ks[0, 0] - 20.0) < 1e-14 and abs(Ks[0, 1] - 20.0 * np.exp(-np.sum((ref[0] - ref[1]) ** 2) / 2500.0)) < 1e-12
    model, L = oracle.approximate_gp_cholesky(ref, [50.0], [20.0], 0.01)
    tr = 3 * np.trace(Ks)
    assert tr - np.sum(L ** 2) <= 0.01 * tr
    assert tr - np.sum(L[:, :-1] ** 2) > 0.01 * tr         # one column less would not have met the tolerance
    assert abs(np.sum(model.variance) - np.sum(L ** 2)) < 1e-9 * tr
