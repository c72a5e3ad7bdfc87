"""gingr_b200/textbook_cpd.py: the reference's textbook CPD variants with the E-step on the device.  The device entry
points (gingr_cpd_estep, gingr_cpd_initial_sigma2: GPU parity in tests/test_estep_gpu.py) are replaced by the oracle's here,
and each variant is compared with a LITERAL restatement of the Scala statements that materialises P (RigidCPD.scala:90-137,
AffineCPD.scala:34-60, NonRigidCPD.scala:46-87)."""
import numpy as np
import pytest


def _install(monkeypatch, oracle):
    from gingr_b200 import api

    class FakeTarget:
        def __init__(self, ctx, pts, tri=None):
            self.points = np.asarray(pts, float)

        def close(self):
            pass
    monkeypatch.setattr(api, "Target", FakeTarget)
    monkeypatch.setattr(api, "cpd_estep", lambda ctx, tgt, fit, s2, w: oracle.cpd_estep(np.asarray(fit), tgt.points, s2, w))
    monkeypatch.setattr(api, "cpd_initial_sigma2", lambda ctx, tgt, pts: oracle.cpd_initial_sigma2(np.asarray(pts), tgt.points))


def _P(X, Y, sigma2, w):
    M, N = len(Y), len(X)
    K = np.exp(-((X[None, :, :] - Y[:, None, :]) ** 2).sum(-1) / (2 * sigma2))
    c = w / (1 - w) * (2.0 * np.pi * sigma2) ** 1.5 * (M / N)
    return K / (K.sum(0)[None, :] + c)


def _literal(kind, X, Y0, G, lam, w, max_iteration, tolerance=0.001):
    M, N = len(Y0), len(X)
    Y = Y0.copy()
    sigma2 = ((Y[:, None, :] - X[None, :, :]) ** 2).sum() / (3 * N * M)
    i, converged = 0, False
    while i < max_iteration and not converged:
        P = _P(X, Y, sigma2, w)
        P1, Pt1, Np = P.sum(1), P.sum(0), P.sum()
        if kind == "nonrigid":
            dinv = np.linalg.inv(np.diag(P1))
            W = np.linalg.solve(G + dinv * (lam * sigma2), dinv @ (P @ X) - Y)
            TY = Y + G @ W
            new = (Pt1 @ (X * X).sum(1) - 2 * np.sum(TY * (P @ X)) + P1 @ (TY * TY).sum(1)) / (Np * 3)
        else:
            muX = X.T @ P.T @ np.ones(M) / Np
            muY = Y.T @ P1 / Np
            Xh, Yh = X - muX, Y - muY
            A = Xh.T @ P.T @ Yh
            if kind == "rigid":
                u, _, v = np.linalg.svd(A)
                C = np.ones(3)
                C[2] = np.linalg.det(u @ v.T)
                R = u @ np.diag(C) @ v
                s = np.trace(A.T @ R) / np.trace(Yh.T @ np.diag(P1) @ Yh)
                new = (np.trace(Xh.T @ np.diag(Pt1) @ Xh) - s * np.trace(A.T @ R)) / (Np * 3)
                TY = s * Y @ R.T + (muX - s * R @ muY)
            else:
                B = A @ np.linalg.inv(Yh.T @ np.diag(P1) @ Yh)
                new = (np.trace(Xh.T @ np.diag(Pt1) @ Xh) - np.trace(A @ B.T)) / (Np * 3)
                TY = Y @ B.T + (muX - B @ muY)
        if abs(new - sigma2) < tolerance:
            converged = True
        else:
            i += 1
        Y, sigma2 = TY, new
    return Y, sigma2, i


@pytest.mark.parametrize("kind,w,iters", [("rigid", 0.0, 25), ("rigid", 0.2, 25), ("affine", 0.1, 25), ("nonrigid", 0.0, 8),
                                          ("nonrigid", 0.3, 8)])
def test_variants_equal_the_literal_statements(oracle, monkeypatch, kind, w, iters):
    import textbook_cpd
    _install(monkeypatch, oracle)
    rng = np.random.default_rng(5)
    Y0 = rng.normal(size=(40, 3)) * 2.0
    Rz = np.array([[np.cos(0.2), -np.sin(0.2), 0], [np.sin(0.2), np.cos(0.2), 0], [0, 0, 1.0]])
    base = np.concatenate([Y0, Y0[:15] + 0.05 * rng.normal(size=(15, 3))])
    X = 1.05 * (base @ Rz.T) + np.array([0.3, -0.2, 0.1]) + 0.05 * np.sin(base[:, [1, 2, 0]])
    cpd = textbook_cpd.CPDFactory(None, Y0, lambda_=2.0, beta=2.0, w=w)
    task = {"rigid": cpd.registerRigidly, "affine": cpd.registerAffine, "nonrigid": cpd.registerNonRigidly}[kind](X)
    got = task.Registration(iters)
    want, sigma2, n_it = _literal(kind, X, Y0, cpd.G, 2.0, w, iters)
    assert task.iterations == n_it and np.all(np.isfinite(want))
    assert abs(task.sigma2 - sigma2) <= 1e-8 * abs(sigma2)
    assert np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want))
    if w == 0.0:                    # the registration did something useful: the template moved onto the target
        d0 = np.sqrt(((X[:, None] - Y0[None]) ** 2).sum(-1).min(1)).mean()
        d1 = np.sqrt(((X[:, None] - got[None]) ** 2).sum(-1).min(1)).mean()
        assert d1 < 0.5 * d0


def test_factory_requirements_and_kernel_matrix(oracle, monkeypatch):
    import textbook_cpd
    _install(monkeypatch, oracle)
    Y = np.random.default_rng(0).normal(size=(12, 3))
    for bad in (dict(w=1.5), dict(w=-0.1), dict(beta=0.0), dict(lambda_=0.0)):
        with pytest.raises(ValueError):
            textbook_cpd.CPDFactory(None, Y, **bad)
    f = textbook_cpd.CPDFactory(None, Y)
    assert (f.lambda_, f.beta, f.w, f.M) == (2.0, 2.0, 0.0, 12)
    lit = np.array([[np.exp(-((a - b) ** 2).sum() / 8.0) for b in Y] for a in Y])
    assert np.max(np.abs(f.G - lit)) < 1e-14 and np.all(np.diag(f.G) == 1.0)
    out = textbook_cpd.RigidCPDRegistration(None, Y, Y + 0.5, max_iterations=30)
    assert np.max(np.abs(out - (Y + 0.5))) < 0.05
