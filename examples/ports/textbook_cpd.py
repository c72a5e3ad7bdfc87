"""The textbook Coherent Point Drift variants of the reference (other/algorithms/cpd/{CPDFactory, RigidCPD, AffineCPD,
NonRigidCPD}.scala, wrappers in other/algorithms/CPDRegistration.scala) over the device E-step.

Their Expectation is the formula of the GiNGR E-step (RigidCPD.scala:90-105), and every Maximization
consumes P only through P1 = P 1, P^T 1 and P X -- exactly what gingr_cpd_estep returns without ever forming the M x N
matrix the reference materialises.  So the O(M N) part runs on the device (K1) and the M-steps are the reference's
statements on those reductions: 3 x 3 algebra for the rigid (+ scale) and affine variants, the reference's M x M solve for
the non-rigid one (numpy on the host: cubic in M as in the reference, fine for the template sizes it is used with).
Quirks kept: RigidCPD estimates a scale as well; convergence is |sigma2_new - sigma2_old| < tolerance with the iteration
counter not advanced on the converging step (RigidCPD.scala:58-80)."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from gingr_b200 import api

DIM = 3


class RigidCPD:
    """other/algorithms/cpd/RigidCPD.scala:30-139 (rigid + isotropic scale)."""

    def __init__(self, targetPoints, cpd: "CPDFactory"):
        self.cpd = cpd
        self.target = np.ascontiguousarray(np.asarray(targetPoints, dtype=np.float64).reshape(-1, 3))
        self.N = self.target.shape[0]
        self._dev_target = api.Target(cpd.ctx, self.target)
        self.iterations = 0

    def close(self):
        self._dev_target.close()

    def initializeGaussianKernel(self, Y) -> float:
        """sum |y_m - x_n|^2 / (dim N M) (:44-55) -- the same reduction as computeInitialSigma2, on the device."""
        return api.cpd_initial_sigma2(self.cpd.ctx, self._dev_target, Y)

    def Expectation(self, Y, sigma2: float):
        """The reductions of P (:90-105): (P1 [M], P^T 1 [N], P X [M, 3])."""
        return api.cpd_estep(self.cpd.ctx, self._dev_target, Y, sigma2, self.cpd.w)

    def _moments(self, Y, P1, Pt1, PX):
        X = self.target
        Np = float(np.sum(P1))
        muX = X.T @ Pt1 / Np
        muY = Y.T @ P1 / Np
        A = PX.T @ Y - Np * np.outer(muX, muY)                       # Xhat^T P^T Yhat
        Yhat = Y - muY
        Xhat = X - muX
        return Np, muX, muY, A, Xhat, Yhat

    def Maximization(self, Y, P1, Pt1, PX, sigma2: float) -> Tuple[np.ndarray, float]:
        """:107-137"""
        Np, muX, muY, A, Xhat, Yhat = self._moments(Y, P1, Pt1, PX)
        u, _, vt = np.linalg.svd(A)
        C = np.ones(DIM)
        C[DIM - 1] = np.linalg.det(u @ vt.T)
        R = u @ np.diag(C) @ vt
        trAR = float(np.trace(A.T @ R))
        s = trAR / float(np.sum(P1 * np.sum(Yhat * Yhat, axis=1)))
        s1 = float(np.sum(Pt1 * np.sum(Xhat * Xhat, axis=1)))
        s2 = s * trAR
        t = muX - s * (R @ muY)
        self.last_transform = (s, R, t)
        return s * (Y @ R.T) + t, (s1 - s2) / (Np * DIM)

    def Iteration(self, Y, sigma2: float) -> Tuple[np.ndarray, float]:
        P1, Pt1, PX = self.Expectation(Y, sigma2)
        return self.Maximization(Y, P1, Pt1, PX, sigma2)

    def Registration(self, max_iteration: int, tolerance: float = 0.001) -> np.ndarray:
        """:57-82 -> the registered template points."""
        Y = self.cpd.template.copy()
        sigma2 = self.initializeGaussianKernel(Y)
        i, converged = 0, False
        while i < max_iteration and not converged:
            Y, new_sigma2 = self.Iteration(Y, sigma2)
            if abs(new_sigma2 - sigma2) < tolerance:
                converged = True
            else:
                i += 1
            sigma2 = new_sigma2
        self.iterations, self.sigma2, self.converged = i, sigma2, converged
        return Y


class AffineCPD(RigidCPD):
    """other/algorithms/cpd/AffineCPD.scala:28-61."""

    def Maximization(self, Y, P1, Pt1, PX, sigma2: float) -> Tuple[np.ndarray, float]:
        Np, muX, muY, A, Xhat, Yhat = self._moments(Y, P1, Pt1, PX)
        B = A @ np.linalg.inv(Yhat.T @ (Yhat * P1[:, None]))
        t = muX - B @ muY
        s1 = float(np.sum(Pt1 * np.sum(Xhat * Xhat, axis=1)))
        s2 = float(np.trace(A @ B.T))
        self.last_transform = (B, t)
        return Y @ B.T + t, (s1 - s2) / (Np * DIM)


class NonRigidCPD(RigidCPD):
    """other/algorithms/cpd/NonRigidCPD.scala:28-88: W = (G + lambda sigma2 diag(P1)^-1) \\ (diag(P1)^-1 P X - Y)."""

    def Maximization(self, Y, P1, Pt1, PX, sigma2: float) -> Tuple[np.ndarray, float]:
        X = self.target
        Np = float(np.sum(P1))
        with np.errstate(divide="ignore", invalid="ignore"):
            inv_p1 = 1.0 / P1                                          # inv(diag(P1)): infinite where a row of P vanished
            A = self.cpd.G + np.diag(inv_p1 * (self.cpd.lambda_ * sigma2))
            B = inv_p1[:, None] * PX - Y
            W = np.linalg.solve(A, B)
        TY = Y + self.cpd.G @ W
        xPx = float(Pt1 @ np.sum(X * X, axis=1))
        yPy = float(P1 @ np.sum(TY * TY, axis=1))
        trPXY = float(np.sum(TY * PX))
        return TY, (xPx - 2.0 * trPXY + yPy) / (Np * DIM)


class CPDFactory:
    """other/algorithms/cpd/CPDFactory.scala:29-79 (lambda = 2, beta = 2, w = 0; G_ij = exp(-|y_i - y_j|^2 / (2 beta^2)))."""

    def __init__(self, ctx: "api.Context", templatePoints, lambda_: float = 2.0, beta: float = 2.0, w: float = 0.0):
        if not (0.0 <= w <= 1.0) or not beta > 0 or not lambda_ > 0:
            raise ValueError("requirement failed: 0 <= w <= 1, beta > 0, lambda > 0")
        self.ctx, self.lambda_, self.beta, self.w = ctx, float(lambda_), float(beta), float(w)
        self.template = np.ascontiguousarray(np.asarray(templatePoints, dtype=np.float64).reshape(-1, 3))
        self.M = self.template.shape[0]
        self._G = None

    @property
    def G(self) -> np.ndarray:
        """The M x M kernel matrix (:52-66); built on first use -- only the non-rigid variant reads it."""
        if self._G is None:
            y = self.template
            sq = np.sum(y * y, axis=1)
            d2 = np.maximum(sq[:, None] + sq[None, :] - 2.0 * (y @ y.T), 0.0)
            np.fill_diagonal(d2, 0.0)
            self._G = np.exp(-d2 / (2.0 * self.beta ** 2))
        return self._G

    def registerRigidly(self, targetPoints) -> RigidCPD:
        return RigidCPD(targetPoints, self)

    def registerAffine(self, targetPoints) -> AffineCPD:
        return AffineCPD(targetPoints, self)

    def registerNonRigidly(self, targetPoints) -> NonRigidCPD:
        return NonRigidCPD(targetPoints, self)


def _register(kind: str, ctx, template, target, lambda_=2.0, beta=2.0, w=0.0, max_iterations=100) -> np.ndarray:
    """{Rigid, Affine, NonRigid}CPDRegistration.register (other/algorithms/CPDRegistration.scala:23-72): the warped
    template points."""
    cpd = CPDFactory(ctx, template, lambda_, beta, w)
    task = {"rigid": cpd.registerRigidly, "affine": cpd.registerAffine, "nonrigid": cpd.registerNonRigidly}[kind](target)
    try:
        return task.Registration(max_iterations)
    finally:
        task.close()


def RigidCPDRegistration(ctx, template, target, **kw) -> np.ndarray:
    return _register("rigid", ctx, template, target, **kw)


def AffineCPDRegistration(ctx, template, target, **kw) -> np.ndarray:
    return _register("affine", ctx, template, target, **kw)


def NonRigidCPDRegistration(ctx, template, target, **kw) -> np.ndarray:
    return _register("nonrigid", ctx, template, target, **kw)
