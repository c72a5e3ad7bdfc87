"""
fast.py -- the CPU restatement of `GingrAlgorithm.update` (CPD) in ALGORITHMIC-MINIMUM form, for problem sizes at which
the literal restatement (oracle.update: three full regressions with SVD pseudo-inverses on a re-posed 3M x r basis, as
the JVM executes them) takes minutes per iteration.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as oracle.py: only tests/, smoke() and bench.py's cpu_baseline /
`--impl reference` legs import it).  PARITY UNPINNED like oracle.py: it is pinned to oracle.update -- and through it to
the reference's Scala statements -- by tests/test_oracle_fast.py, which requires every state component of the two forms to
agree to 1e-9 on seeded problems (rigid / similarity / no transform, w = 0 and w > 0, step lengths, several iterations).

What is reduced, each identity being exact algebra on the statements of api/GingrAlgorithm.scala:192-254 and of scalismo's
regression:

  * model.transform(R, t) never forms the re-posed basis: Phi' = (I (x) R) Phi, so with Q = Phi diag(sqrt(lambda))
        Q'^T W Q' = Q^T W Q   for isotropic observation noise (CPD: sigma2 lambda / P1_i, CPD.scala:120-128), and
        Q'^T (y - m') = Q^T vec(((points - t) R) - (ref + mean))          (residuals rotated back instead)
  * pinv(Mx) (Breeze SVD) -> Cholesky solve: Mx = I + PSD
  * coefficients(mesh) (all M points, noise 1e-5): (G + 1e-5 I)^-1 Q^T u with G = Q^T Q computed once per model;
    coefficients(posterior.mean) = (G + 1e-5 I)^-1 G c  (no pass over the basis)
  * one streaming E-step per iteration (the sigma2 update uses the PRE-update fit)

One iteration = E-step (C, OpenMP) + one weighted Gram (BLAS dsyrk) + r x r Cholesky + five passes over Q.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import oracle


class FastCpdModel:
    """Per-model constants: Q = Phi sqrt(lambda), G = Q^T Q, the factor of G + 1e-5 I."""

    def __init__(self, model: "oracle.Gpmm"):
        import scipy.linalg as sla
        self.model = model
        self.Q = np.ascontiguousarray(model.basis * np.sqrt(model.variance)[None, :])
        self.rm = (model.ref + model.mean.reshape(-1, 3))                       # ref + mean, [M, 3]
        self.G = sla.blas.dsyrk(1.0, self.Q, trans=1, lower=1)
        self.G = np.tril(self.G) + np.tril(self.G, -1).T
        self.K0 = sla.cho_factor(self.G + 1e-5 * np.eye(model.rank), lower=True)

    def instance_unposed(self, alpha):
        """ref + mean + Phi (sqrt(lambda) alpha)"""
        return self.rm + (self.Q @ alpha).reshape(-1, 3)

    def coefficients(self, R, t, mesh):
        """model.transform(R, t).coefficients(mesh)   [A3], noise 1e-5 I3 on all M points"""
        import scipy.linalg as sla
        u = ((np.asarray(mesh) - t) @ R - self.rm).reshape(-1)
        return sla.cho_solve(self.K0, self.Q.T @ u)


def update(fm: FastCpdModel, algo: "oracle.CpdAlgorithm", st: "oracle.State") -> "oracle.State":
    """GingrAlgorithm.update (deterministic, CPD, no landmarks) -- the statements of oracle.update in reduced form."""
    import scipy.linalg as sla
    cfg = algo.config
    if cfg.use_landmark_correspondence and st.landmarks is not None and len(st.landmarks.pids) > 0:
        raise NotImplementedError("fast.update: landmark observations need the literal form (oracle.update)")
    R, t = st.params.rotation_matrix(), st.params.translation
    fail = dataclasses.replace(st, status=oracle.STATUS_MODEL_FLEXIBILITY_ERROR) if st.iteration > 0 else st
    # ---- computePosterior: E-step, observations, regression (GingrAlgorithm.scala:281-302) --------------------------
    P1, Pt1, PX = oracle.cpd_estep(st.fit, st.target, st.sigma2, cfg.w, fast=True)
    with np.errstate(all="ignore"):
        td = PX / P1[:, None]                                                 # CPD.scala:32-49
        wgt = P1 / (st.sigma2 * cfg.lam)                                      # 1 / (sigma2 lambda / P1_i), CPD.scala:120-128
    if not (np.all(np.isfinite(td)) and np.all(np.isfinite(wgt)) and np.all(wgt > 0)):
        return fail
    w3 = np.repeat(wgt, 3)
    Qw = fm.Q * np.sqrt(w3)[:, None]
    Mx = sla.blas.dsyrk(1.0, Qw, trans=1, lower=1)
    del Qw
    Mx[np.diag_indices_from(Mx)] += 1.0
    u = ((td - t) @ R - fm.rm).reshape(-1)
    rhs = fm.Q.T @ (w3 * u)
    try:
        c_post = sla.cho_solve(sla.cho_factor(Mx, lower=True, check_finite=True), rhs)
    except (np.linalg.LinAlgError, ValueError):
        return fail
    if not np.all(np.isfinite(c_post)):
        return fail
    # ---- alpha* = transformedModel.coefficients(posterior.mean) (:211-216) ------------------------------------------
    new_coeffs = sla.cho_solve(fm.K0, fm.G @ c_post)
    cur = st.params.shape
    combined = cur + (new_coeffs - cur) * st.step_length                      # :218-220
    newshape = fm.instance_unposed(combined) @ R.T + t                        # :222
    current_fit_no_transform = fm.instance_unposed(cur)                       # :224
    if st.global_transformation == oracle.SIMILARITY_TRANSFORMS:              # :227-231
        R2, t2, s2 = oracle.umeyama(current_fit_no_transform, newshape, True)
    elif st.global_transformation == oracle.RIGID_TRANSFORMS:
        R2, t2, s2 = oracle.umeyama(current_fit_no_transform, newshape, False)
    else:
        R2, t2, s2 = np.eye(3), np.zeros(3), 1.0
    alpha = fm.coefficients(R2, t2, newshape)                                 # :232-237
    if not np.all(np.isfinite(alpha)):
        return dataclasses.replace(st, status=oracle.STATUS_MODEL_FLEXIBILITY_ERROR)
    params = oracle.Params(float(s2), np.asarray(t2, dtype=float), oracle.matrix_to_euler(R2), alpha)   # :239-243
    sigma2 = oracle.cpd_sigma2_update(P1, Pt1, PX, st.target, st.fit)         # :245-246, pre-update fit
    return dataclasses.replace(st, params=params, sigma2=float(sigma2))


def propose(fm: FastCpdModel, algo: "oracle.CpdAlgorithm", st: "oracle.State") -> "oracle.State":
    """GingrGeneratorWrapper.propose (sampling/generators/GingrGeneratorWrapper.scala:28-39): update, refresh the fit,
    iteration += 1."""
    ns = update(fm, algo, st)
    p = ns.params
    fit = (fm.instance_unposed(p.shape) @ p.rotation_matrix().T + p.translation) * p.scale   # ModelFittingParameters.scala:130-143
    return dataclasses.replace(ns, fit=fit, iteration=ns.iteration + 1)
