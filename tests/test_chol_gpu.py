"""K3b: the on-device r x r Cholesky solve (gingr_spd_solve = the factorisation + substitutions the iteration uses in
place of scalismo's `pinv(Mx) * rhs`) against numpy / LAPACK on the same matrices.  Sizes cover one tile,
ragged last tiles, right-hand sides that land in the last diagonal tile or in a tile row of their own, many extra rows
(the model constants use nrows = 2n) and the benchmark rank.  Tolerances: factor and solution 1e-12 relative (the
iteration's own bar is 1e-6; the factorisation is plain FP64)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spd(n, seed, cond=1e3):
    rng = np.random.default_rng(seed)
    Q = rng.normal(size=(n, n + 8))
    A = Q @ Q.T / (n + 8)
    A += np.eye(n) * (np.trace(A) / n / cond)
    return A


@pytest.mark.parametrize("n,nrhs", [(1, 1), (7, 1), (31, 2), (32, 1), (33, 1), (50, 1), (63, 1), (64, 1), (64, 0), (65, 1),
                                    (100, 1), (127, 3), (128, 1), (129, 1), (200, 200), (50, 50), (300, 1), (520, 7)])
def test_spd_solve_matches_lapack(ctx, n, nrhs):
    from gingr_b200 import api
    A = _spd(n, n)
    B = np.random.default_rng(n + 1).normal(size=(nrhs, n)) if nrhs else None
    out = api.spd_solve(ctx, A, B)
    L = np.linalg.cholesky(A)
    scale = np.abs(L).max()
    assert np.max(np.abs(out["L"] - L)) < 1e-12 * scale
    if nrhs:
        Y = np.linalg.solve(L, B.T).T
        assert np.max(np.abs(out["Y"] - Y)) < 1e-11 * max(1.0, np.abs(Y).max())
        x = np.linalg.solve(A, B[0])
        assert np.max(np.abs(out["x"] - x)) < 1e-10 * max(1.0, np.abs(x).max())


def test_spd_solve_benchmark_rank_and_repeatability(ctx):
    """r = 2000 (C4): 32 block columns, the right-hand side in the ragged last tile row; two launches in a row reuse the
    self-resetting tickets / flags and must give identical bits."""
    from gingr_b200 import api
    n = 2000
    A = _spd(n, 3, cond=1e6)
    b = np.random.default_rng(4).normal(size=(1, n))
    o1 = api.spd_solve(ctx, A, b, reps=3)
    o2 = api.spd_solve(ctx, A, b)
    assert np.array_equal(o1["L"], o2["L"]) and np.array_equal(o1["x"], o2["x"])
    L = np.linalg.cholesky(A)
    assert np.max(np.abs(o1["L"] - L)) < 1e-11 * np.abs(L).max()
    x = np.linalg.solve(A, b[0])
    assert np.max(np.abs(o1["x"] - x)) < 1e-8 * np.abs(x).max()       # cond 1e6
    res = A @ o1["x"] - b[0]
    assert np.max(np.abs(res)) < 1e-9 * np.abs(b).max()


def test_spd_solve_wide_dynamic_range(ctx):
    """Posterior matrices at CPD convergence: I + Q^T W Q with weights up to 1e10 (sigma2 -> 1e-10)."""
    from gingr_b200 import api
    rng = np.random.default_rng(11)
    n = 150
    Q = rng.normal(size=(400, n))
    w = 10.0 ** rng.uniform(-6, 10, size=400)
    A = np.eye(n) + (Q * w[:, None]).T @ Q
    b = rng.normal(size=(1, n)) * 1e8
    out = api.spd_solve(ctx, A, b)
    L = np.linalg.cholesky(A)
    assert np.max(np.abs(out["L"] - L)) < 1e-9 * np.abs(L).max()
    res = A @ out["x"] - b[0]
    assert np.max(np.abs(res)) < 1e-8 * np.abs(b).max()


@pytest.mark.parametrize("n", [40, 130])
def test_spd_solve_flags_indefinite_and_nonfinite(ctx, n):
    from gingr_b200 import api
    A = _spd(n, 5)
    A[n // 2, n // 2] = -1.0
    with pytest.raises(FloatingPointError):
        api.spd_solve(ctx, A, np.ones((1, n)))
    A = _spd(n, 6)
    A[n - 1, 0] = A[0, n - 1] = np.nan
    with pytest.raises(FloatingPointError):
        api.spd_solve(ctx, A, np.ones((1, n)))
    # and the context stays usable (no stuck flags after a failed factorisation)
    A = _spd(n, 7)
    out = api.spd_solve(ctx, A, np.ones((1, n)))
    assert np.max(np.abs(out["L"] - np.linalg.cholesky(A))) < 1e-12 * np.abs(out["L"]).max()
