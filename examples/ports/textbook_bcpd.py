"""Bayesian Coherent Point Drift of the reference (other/algorithms/cpd/BCPD.scala, wrapper other/algorithms/BCPDRegistration.scala)
over the device E-step gingr_bcpd_estep.

computeP and its reductions (BCPD.scala:167-184, :200-209: nu, nu', N-hat, x-hat -- the M x N part, which the reference
materialises and even Kronecker-expands to 3M x 3N) run on the device; the rest of an iteration is the reference's
statement sequence on those reductions: the M x M pseudo-inverses of the local-deformation update on the host (cubic in M,
as in the reference), 3 x 3 algebra for the similarity transform.  Quirks kept literally: `s` (not s^2) in the exponent of
computeP and (1 - w) applied twice (inside the device kernel); the identity times sigma2bar added M times in Suu; the new
points are transformed with the PREVIOUS similarity parameters; on convergence the previous fit is returned."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

from gingr_b200 import api

DIM = 3


def gaussian_kernel_matrix(points, sigma: float) -> np.ndarray:
    """G_ij = GaussianKernel(sigma)(y_i, y_j) = exp(-|y_i - y_j|^2 / sigma^2) [scalismo's Gaussian kernel convention,
    SURVEY A7] -- the usual choice for BCPD's `kernel` argument; any PSD M x M matrix may be passed instead."""
    y = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    sq = np.sum(y * y, axis=1)
    d2 = np.maximum(sq[:, None] + sq[None, :] - 2.0 * (y @ y.T), 0.0)
    np.fill_diagonal(d2, 0.0)
    return np.exp(-d2 / (sigma * sigma))


@dataclass
class SimilarityTransformationParameters:
    """BCPD.scala:29-36."""
    sigma: np.ndarray      # [M, M]
    s: float
    R: np.ndarray          # [3, 3]
    t: np.ndarray          # [3]
    sigma2: float
    alpha: np.ndarray      # [M]


class BCPD:
    """BCPD.scala:42-260.  G: the M x M kernel matrix of the template points (the reference evaluates its PDKernel)."""

    def __init__(self, ctx: "api.Context", templatePoints, targetPoints, w: float, lambda_: float, gamma: float, k: float, G):
        if not (0.0 <= w <= 1.0) or not lambda_ > 0 or not gamma > 0:
            raise ValueError("requirement failed: 0 <= w <= 1, lambda > 0, gamma > 0")
        self.ctx, self.w, self.lambda_, self.gamma, self.k = ctx, float(w), float(lambda_), float(gamma), float(k)
        self.Y = np.ascontiguousarray(np.asarray(templatePoints, dtype=np.float64).reshape(-1, 3))
        self.X = np.ascontiguousarray(np.asarray(targetPoints, dtype=np.float64).reshape(-1, 3))
        self.M, self.N = self.Y.shape[0], self.X.shape[0]
        self.G = np.asarray(G, dtype=np.float64)
        if self.G.shape != (self.M, self.M):
            raise ValueError("G must be M x M")
        self.GinvLambda = np.linalg.pinv(self.G) * self.lambda_                      # :60-61
        self._dev_target = api.Target(ctx, self.X)
        self.iterations, self.converged, self.pars = 0, False, None

    def close(self):
        self._dev_target.close()

    def _transform(self, v: np.ndarray, pars: SimilarityTransformationParameters) -> np.ndarray:
        return (v @ pars.R.T + pars.t) * pars.s                                        # :150-154 s (R v + t)

    def _inv_transform(self, v: np.ndarray, pars: SimilarityTransformationParameters) -> np.ndarray:
        return ((v - pars.t) * (1.0 / pars.s)) @ np.linalg.pinv(pars.R).T              # :157-165

    def Iteration(self, Yhat: np.ndarray, pars: SimilarityTransformationParameters
                  ) -> Tuple[np.ndarray, SimilarityTransformationParameters]:
        """:194-259"""
        from scipy.special import digamma
        v, v_, Nhat, xhat = api.bcpd_estep(self.ctx, self._dev_target, Yhat, np.diag(pars.sigma).copy(), pars.alpha,
                                           pars.sigma2, pars.s, self.w)                # :196-207 on the device
        xhat_tinv = self._inv_transform(xhat, pars)
        s2divsigma2 = pars.s ** 2 / pars.sigma2                                        # :211-216 local deformations
        Sigma = np.linalg.pinv(self.GinvLambda + np.diag(v) * s2divsigma2)
        vhat = s2divsigma2 * (Sigma @ (v[:, None] * (xhat_tinv - self.Y)))
        uhat = self.Y + vhat
        alpha = np.exp(digamma(self.k + v) - digamma(self.k * self.M + Nhat))          # :218
        x_mean = (v[:, None] * xhat).sum(0) / Nhat                                     # :221-223
        u_mean = (v[:, None] * uhat).sum(0) / Nhat
        sigma2bar = float(np.sum(v * np.diag(Sigma)) / Nhat)
        Sxu = ((xhat - x_mean) * v[:, None]).T @ (uhat - u_mean) / Nhat                # :225-229
        Suu = (((uhat - u_mean) * v[:, None]).T @ (uhat - u_mean) + self.M * np.eye(DIM) * sigma2bar) / Nhat   # :231-234
        phi, _, psiT = np.linalg.svd(Sxu)                                              # :236-242
        d = np.ones(DIM)
        d[DIM - 1] = np.linalg.det(phi @ psiT)
        R = phi @ np.diag(d) @ psiT
        s = float(np.trace(R @ Sxu) / np.trace(Suu))
        t = x_mean - s * (R @ u_mean)
        new_yhat = self._transform(uhat, pars)                                         # :244 previous parameters
        sXX = float(np.sum(v_ * np.sum(self.X * self.X, axis=1)))                      # :246-251
        sXY = float(np.sum((v[:, None] * xhat) * new_yhat))                            # X^T (P (x) I)^T Yhat = sum_m (P X)_m . yhat_m
        sYY = float(np.sum(v * np.sum(new_yhat * new_yhat, axis=1)))
        sC = pars.sigma2 * sigma2bar
        new_sigma2 = (sXX - 2.0 * sXY + sYY + sC) / (Nhat * DIM)
        return new_yhat, SimilarityTransformationParameters(Sigma, s, R, t, new_sigma2, alpha)

    def Registration(self, max_iteration: int, tolerance: float = 0.000001) -> np.ndarray:
        """:112-146"""
        sigma2_init = self.gamma * api.cpd_initial_sigma2(self.ctx, self._dev_target, self.Y)
        pars = SimilarityTransformationParameters(np.eye(self.M), 1.0, np.eye(DIM), np.zeros(DIM), sigma2_init,
                                                  np.ones(self.M) / self.M)
        fit, i, converged = self.Y.copy(), 0, False
        while i < max_iteration and not converged:
            ty, new_pars = self.Iteration(fit, pars)
            if abs(new_pars.sigma2 - pars.sigma2) < tolerance:
                converged = True                                                       # the previous fit is kept (:131-134)
            else:
                fit, pars = ty, new_pars
                i += 1
        self.iterations, self.converged, self.pars = i, converged, pars
        return fit


def BCPDRegistration(ctx, template, target, G, w: float = 0.0, lambda_: float = 2.0, gamma: float = 1.0, k: float = 1.0,
                     max_iterations: int = 100) -> np.ndarray:
    """BCPDRegistration.register (other/algorithms/BCPDRegistration.scala:25-45): the warped template points."""
    task = BCPD(ctx, template, target, w, lambda_, gamma, k, G)
    try:
        return task.Registration(max_iterations)
    finally:
        task.close()
