"""gingr_b200/textbook_bcpd.py: the reference's BCPD with computeP and its reductions on the device.  The device entry
points (gingr_bcpd_estep, gingr_cpd_initial_sigma2: GPU parity in tests/test_estep_gpu.py) are replaced by the oracle's;
the iteration is compared with a LITERAL restatement of BCPD.scala:112-259 -- explicit P, Kronecker products and all."""
import numpy as np
import pytest


def _install(monkeypatch, oracle):
    from gingr_b200 import api

    class FakeTarget:
        def __init__(self, ctx, pts, tri=None):
            self.points = np.ascontiguousarray(np.asarray(pts, float))

        def close(self):
            pass
    monkeypatch.setattr(api, "Target", FakeTarget)
    monkeypatch.setattr(api, "bcpd_estep", lambda ctx, tgt, y, sm, al, s2, s, w: oracle.bcpd_estep(np.asarray(y), tgt.points, sm, al, s2, s, w))
    monkeypatch.setattr(api, "cpd_initial_sigma2", lambda ctx, tgt, pts: oracle.cpd_initial_sigma2(np.asarray(pts), tgt.points))


def _literal(oracle, Ypts, Xpts, w, lam, gamma, k, G, max_iteration, tolerance=1e-6):
    from scipy.special import digamma
    M, N, dim = len(Ypts), len(Xpts), 3
    GinvLambda = np.linalg.pinv(G) * lam
    X, Y = Xpts.reshape(-1), Ypts.reshape(-1)
    D1, Dm = np.ones((1, dim)), np.eye(dim)
    pars = dict(sigma=np.eye(M), s=1.0, R=np.eye(dim), t=np.zeros(dim),
                sigma2=gamma * ((Ypts[:, None] - Xpts[None]) ** 2).sum() / (dim * N * M), alpha=np.ones(M) / M)

    def vt(v, p):
        return (np.kron(np.eye(M), p["R"]) @ v + np.kron(np.ones((1, M)), p["t"][None, :]).reshape(-1)) * p["s"]

    def vit(v, p):
        tmp = v + np.kron(np.ones((1, M)), -p["t"][None, :]).reshape(-1)
        return np.kron(np.eye(M), np.linalg.pinv(p["R"])) @ (tmp * (1.0 / p["s"]))
    fit, i, converged = Ypts.copy(), 0, False
    while i < max_iteration and not converged:
        P = oracle.bcpd_P(fit, Xpts, np.diag(pars["sigma"]).copy(), pars["alpha"], pars["sigma2"], pars["s"], w)
        v, v_ = P.sum(1), P.sum(0)
        Nhat = v_.sum()
        Pk, vk, v_k = np.kron(P, Dm), np.kron(v[None, :], D1).reshape(-1), np.kron(v_[None, :], D1).reshape(-1)
        xhat = np.linalg.pinv(np.diag(vk)) @ Pk @ X
        xti = vit(xhat, pars)
        s2 = pars["s"] ** 2 / pars["sigma2"]
        Sigma = np.linalg.pinv(GinvLambda + np.diag(v) * s2)
        vhat = (np.kron(Sigma, Dm) * s2) @ np.diag(vk) @ (xti - Y)
        uhat = Y + vhat
        alpha = np.exp(digamma(k + v) - digamma(k * M + Nhat))
        seg = lambda a, m: a[m * dim:(m + 1) * dim]
        xMean = sum(seg(xhat, m) * v[m] for m in range(M)) / Nhat
        uMean = sum(seg(uhat, m) * v[m] for m in range(M)) / Nhat
        s2bar = sum(v[m] * Sigma[m, m] for m in range(M)) / Nhat
        Sxu = sum(np.outer((seg(xhat, m) - xMean) * v[m], seg(uhat, m) - uMean) for m in range(M)) / Nhat
        Suu = sum(np.outer((seg(uhat, m) - uMean) * v[m], seg(uhat, m) - uMean) + np.eye(dim) * s2bar for m in range(M)) / Nhat
        phi, _, psiT = np.linalg.svd(Sxu)
        dd = np.ones(dim)
        dd[-1] = np.linalg.det(phi @ psiT)
        R = phi @ np.diag(dd) @ psiT
        s = np.trace(R @ Sxu) / np.trace(Suu)
        t = xMean - (R * s) @ uMean
        newY = vt(Y + vhat, pars)
        new_sigma2 = (X @ np.diag(v_k) @ X - 2 * (X @ Pk.T @ newY) + newY @ np.diag(vk) @ newY + pars["sigma2"] * s2bar) / (Nhat * dim)
        new = dict(sigma=Sigma, s=s, R=R, t=t, sigma2=new_sigma2, alpha=alpha)
        if abs(new_sigma2 - pars["sigma2"]) < tolerance:
            converged = True
        else:
            fit, pars = newY.reshape(-1, 3), new
            i += 1
    return fit, pars, i


@pytest.mark.parametrize("w", [0.0, 0.2])
def test_bcpd_equals_the_literal_statements(oracle, monkeypatch, w):
    import textbook_bcpd
    _install(monkeypatch, oracle)
    rng = np.random.default_rng(9)
    Y = rng.normal(size=(24, 3)) * 2.0
    Rz = np.array([[np.cos(0.2), -np.sin(0.2), 0], [np.sin(0.2), np.cos(0.2), 0], [0, 0, 1.0]])
    base = np.concatenate([Y, Y[:8] + 0.05 * rng.normal(size=(8, 3))])
    X = 1.05 * (base @ Rz.T) + np.array([0.3, -0.2, 0.1]) + 0.05 * np.sin(base[:, [1, 2, 0]])
    G = textbook_bcpd.gaussian_kernel_matrix(Y, 3.0)
    assert np.max(np.abs(G - np.array([[np.exp(-((a - b) ** 2).sum() / 9.0) for b in Y] for a in Y]))) < 1e-14
    task = textbook_bcpd.BCPD(None, Y, X, w, 2.0, 1.0, 1.0, G)
    got = task.Registration(12)
    want, pars, iters = _literal(oracle, Y, X, w, 2.0, 1.0, 1.0, G, 12)
    assert task.iterations == iters and np.all(np.isfinite(want))
    assert np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want))
    assert abs(task.pars.sigma2 - pars["sigma2"]) <= 1e-8 * abs(pars["sigma2"]) and abs(task.pars.s - pars["s"]) < 1e-9
    assert np.max(np.abs(task.pars.R - pars["R"])) < 1e-9 and np.max(np.abs(task.pars.alpha - pars["alpha"])) < 1e-10
    d0 = np.sqrt(((X[:, None] - Y[None]) ** 2).sum(-1).min(1)).mean()
    d1 = np.sqrt(((X[:, None] - got[None]) ** 2).sum(-1).min(1)).mean()
    if w == 0.0:
        assert d1 < d0
    out = textbook_bcpd.BCPDRegistration(None, Y, X, G, w=w, max_iterations=12)
    assert np.array_equal(out, got)


def test_bcpd_requirements(oracle, monkeypatch):
    import textbook_bcpd
    _install(monkeypatch, oracle)
    Y = np.random.default_rng(0).normal(size=(6, 3))
    G = np.eye(6)
    for bad in (dict(w=1.2, lambda_=2.0, gamma=1.0), dict(w=0.0, lambda_=0.0, gamma=1.0), dict(w=0.0, lambda_=2.0, gamma=0.0)):
        with pytest.raises(ValueError):
            textbook_bcpd.BCPD(None, Y, Y, bad["w"], bad["lambda_"], bad["gamma"], 1.0, G)
    with pytest.raises(ValueError):
        textbook_bcpd.BCPD(None, Y, Y, 0.0, 2.0, 1.0, 1.0, np.eye(5))
