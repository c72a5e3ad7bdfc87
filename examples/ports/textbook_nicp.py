"""Optimal-step non-rigid ICP of the reference (other/algorithms/icp/NonRigidOptimalStepICP.scala: N-ICP-T and N-ICP-A of
Amberg et al.) over the device closest-point search.

Each iteration takes the robust surface correspondences of the GiNGR ICP -- closest point on the target surface, 0 / 1
weight from the boundary / normal / self-intersection predicates, mean distance (ClosestPointRegistrator.scala:74-96) --
from gingr_icp_closest (triangular flavour: the exact K2 search, uniform grid at scale) and then solves the reference's
sparse least-squares system  [alpha M; W ...; beta landmarks] X = [0; W ...; beta ...]  on the host (scipy.sparse normal
equations; Breeze's `\\` on a tall CSC matrix is the same least-squares solution).  Quirks kept: the default stiffness
schedule is eleven times 10.0 (the scanLeft list is overwritten by `.map(_ => 1e1)`, :61-63); N-ICP-T writes its landmark
rows at column i, not at the landmark's vertex id (:174-175)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from gingr_b200 import api

DIM = 3
DEFAULT_ALPHA = [10.0] * 11          # NonRigidOptimalStepICP.scala:61-63
DEFAULT_BETA = DEFAULT_ALPHA


def _nearest(points: np.ndarray, q: np.ndarray) -> int:
    d = np.sum((points - q[None, :]) ** 2, axis=1)
    return int(np.argmin(d))                                         # lowest index on ties


class NonRigidOptimalStepICP:
    """:30-138.  Meshes are (points [n, 3], triangles [t, 3]); landmarks are io.Landmark lists (matched by id)."""

    def __init__(self, ctx: "api.Context", templateMesh, targetMesh, templateLandmarks=(), targetLandmarks=(), gamma: float = 1.0):
        import scipy.sparse as sp
        if gamma < 0:
            raise ValueError("requirement failed: gamma >= 0")
        self.ctx, self.gamma = ctx, float(gamma)
        self.template = np.ascontiguousarray(np.asarray(templateMesh[0], dtype=np.float64).reshape(-1, 3))
        self.triangles = np.ascontiguousarray(np.asarray(templateMesh[1], dtype=np.int32).reshape(-1, 3))
        tv = np.ascontiguousarray(np.asarray(targetMesh[0], dtype=np.float64).reshape(-1, 3))
        self.n = self.template.shape[0]
        tgt_by_id = {l.id: l for l in targetLandmarks}
        common = [l for l in templateLandmarks if l.id in tgt_by_id]                     # :44-53
        self.lmIdsOnTemplate = np.array([_nearest(self.template, np.asarray(l.point, float)) for l in common], dtype=np.int64)
        self.UL = np.array([tv[_nearest(tv, np.asarray(tgt_by_id[l.id].point, float))] for l in common], dtype=np.float64).reshape(-1, 3)
        t = np.sort(self.triangles.astype(np.int64), axis=1)                              # :65-73 unique sorted edges
        e = np.unique(np.concatenate([t[:, [0, 1]], t[:, [0, 2]], t[:, [1, 2]]]), axis=0)
        self.numOfEdges = e.shape[0]
        rows = np.repeat(np.arange(self.numOfEdges), 2)
        self.M = sp.csr_matrix((np.tile([1.0, -1.0], self.numOfEdges), (rows, e.reshape(-1))), shape=(self.numOfEdges, self.n))   # :75-84
        self._dev_target = api.Target(ctx, tv, np.asarray(targetMesh[1], dtype=np.int32).reshape(-1, 3))
        self.iterations = 0

    def close(self):
        self._dev_target.close()

    def getClosestPoints(self, template: np.ndarray) -> Tuple[np.ndarray, np.ndarray, float]:
        """:118-127 -> (corresponding points [n, 3], weights [n] in {0, 1}, mean distance), on the device."""
        _, cp, w, dist = api.icp_closest(self.ctx, self._dev_target, template, self.triangles, api.TRIANGULAR_CLOSEST_POINT)
        return cp, np.asarray(w, dtype=np.float64), dist

    @staticmethod
    def _least_squares(A, B) -> np.ndarray:
        import scipy.sparse.linalg as spla
        AtA = (A.T @ A).tocsc()
        return np.column_stack([spla.spsolve(AtA, np.asarray(A.T @ B[:, d]).reshape(-1)) for d in range(B.shape[1])])

    def Iteration(self, template: np.ndarray, alpha: float, beta: float):
        raise NotImplementedError

    def Registration(self, max_iteration: int, tolerance: float = 0.001, alpha: Optional[Sequence[float]] = None,
                     beta: Optional[Sequence[float]] = None) -> np.ndarray:
        """:86-116 -> the registered template points."""
        alpha = DEFAULT_ALPHA if alpha is None else list(alpha)
        beta = DEFAULT_BETA if beta is None else list(beta)
        if len(alpha) != len(beta):
            raise ValueError("requirement failed: alpha.length == beta.length")
        fit = self.template.copy()
        self.iterations = 0
        for a, b in zip(alpha, beta):
            dist, i = float("inf"), 0
            while i < max_iteration and dist >= tolerance:
                fit, dist = self.Iteration(fit, a, b)[:2]
                i += 1
                self.iterations += 1
            self.distance = dist
        return fit


class NonRigidOptimalStepICP_T(NonRigidOptimalStepICP):
    """N-ICP-T (:140-192): one translation per vertex."""

    def Iteration(self, template: np.ndarray, alpha: float, beta: float):
        import scipy.sparse as sp
        if alpha < 0 or beta < 0:
            raise ValueError("requirement failed: alpha, beta >= 0")
        cp, w, dist = self.getClosestPoints(template)
        L = len(self.lmIdsOnTemplate)
        W = sp.diags(w)
        VL = template[self.lmIdsOnTemplate].reshape(-1, 3)
        A3 = sp.csr_matrix((np.ones(L), (np.arange(L), np.arange(L))), shape=(L, self.n))     # :174-175 column i, as written
        A = sp.vstack([self.M * alpha, W, A3]).tocsr()
        B = np.vstack([np.zeros((self.numOfEdges, DIM)), w[:, None] * (cp - template), (self.UL - VL) * beta])
        X = self._least_squares(A, B)
        return template + X, dist, np.zeros((0, 3))


class NonRigidOptimalStepICP_A(NonRigidOptimalStepICP):
    """N-ICP-A (:198-284): one affine 3 x 4 transformation per vertex, stiffness on their differences (G = diag(1, 1, 1, gamma))."""

    def __init__(self, *args, **kw):
        import scipy.sparse as sp
        super().__init__(*args, **kw)
        self.kronMG = sp.kron(self.M, sp.diags([1.0, 1.0, 1.0, self.gamma])).tocsr()           # :211-212

    def _D(self, points: np.ndarray):
        import scipy.sparse as sp
        n = points.shape[0]
        rows = np.repeat(np.arange(n), 4)
        cols = np.arange(4 * n)
        vals = np.column_stack([points, np.ones(n)]).reshape(-1)
        return sp.csr_matrix((vals, (rows, cols)), shape=(n, 4 * n))                            # :216-231

    def _DL(self, template: np.ndarray):
        import scipy.sparse as sp
        L = len(self.lmIdsOnTemplate)
        rows = np.repeat(np.arange(L), 4)
        cols = (self.lmIdsOnTemplate[:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
        vals = np.column_stack([template[self.lmIdsOnTemplate].reshape(-1, 3), np.ones(L)]).reshape(-1)
        return sp.csr_matrix((vals, (rows, cols)), shape=(L, 4 * self.n))                       # :233-244

    def Iteration(self, template: np.ndarray, alpha: float, beta: float):
        import scipy.sparse as sp
        if alpha < 0 or beta < 0:
            raise ValueError("requirement failed: alpha, beta >= 0")
        cp, w, dist = self.getClosestPoints(template)
        w = w.copy()
        w[self.lmIdsOnTemplate] = 0.0                                                           # :257-260
        D, DL = self._D(template), self._DL(template)
        A = sp.vstack([self.kronMG * alpha, sp.diags(w) @ D, DL * beta]).tocsr()
        B = np.vstack([np.zeros((4 * self.numOfEdges, DIM)), w[:, None] * cp, self.UL * beta])
        X = self._least_squares(A, B)
        return np.asarray(D @ X), dist, np.asarray(DL @ X)
