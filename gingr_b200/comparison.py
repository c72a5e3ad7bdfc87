"""Mesh-to-mesh quality measures of the reference (api/helper/RegistrationComparison.scala:24-86), the last statement of
SimpleRegistrator.run (SimpleRegistrator.scala:156).  Every closest-point-on-surface query is the device's exact K2
search (api.icp_closest with the triangular flavour: closest point on the surface and the vertex nearest to it); the
few reductions over the M results are numpy.  scalismo's MeshMetrics.avgDistance / hausdorffDistance are the mean /
two-sided maximum of the same point-to-surface distances [scalismo-recalled]."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import api


def boundary_vertices(n_vertices: int, triangles) -> np.ndarray:
    """mesh.operations.pointIsOnBoundary for every vertex: end points of edges that belong to exactly one triangle."""
    t = np.asarray(triangles, dtype=np.int64).reshape(-1, 3)
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=0)
    e.sort(axis=1)
    key = e[:, 0] * np.int64(n_vertices) + e[:, 1]
    uniq, counts = np.unique(key, return_counts=True)
    lone = uniq[counts == 1]
    out = np.zeros(n_vertices, dtype=bool)
    out[lone // n_vertices] = True
    out[lone % n_vertices] = True
    return out


def _norms(d: np.ndarray) -> np.ndarray:
    return np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])      # EuclideanVector3D.norm


class RegistrationComparison:
    """Meshes are (points [n, 3], triangles [t, 3]) pairs; `ctx` is the api.Context the searches run on."""

    def __init__(self, ctx: "api.Context"):
        self.ctx = ctx

    def _closest(self, m1, m2) -> Tuple[np.ndarray, np.ndarray]:
        """(closest point on the surface of m2, id of the m2 vertex nearest to it) for every vertex of m1.  The
        triangular flavour of the K2 entry point also evaluates the reference's robustness predicates, which need the
        query mesh's triangles; its weights are not used here."""
        target = api.Target(self.ctx, m2[0], m2[1])
        try:
            idx, cp, _, _ = api.icp_closest(self.ctx, target, m1[0], m1[1], api.TRIANGULAR_CLOSEST_POINT)
        finally:
            target.close()
        return cp, idx

    def distances(self, m1, m2) -> np.ndarray:
        p = np.asarray(m1[0], dtype=np.float64).reshape(-1, 3)
        cp, _ = self._closest(m1, m2)
        return _norms(p - cp)

    def maxDistance(self, m1, m2) -> float:
        """:26-36 largest distance from a vertex of m1 to the surface of m2 (one-sided)."""
        return float(self.distances(m1, m2).max())

    def avgDistance(self, m1, m2) -> float:
        """MeshMetrics.avgDistance(m1, m2): mean distance from the vertices of m1 to the surface of m2."""
        return float(self.distances(m1, m2).mean())

    def hausdorffDistance(self, m1, m2) -> float:
        """MeshMetrics.hausdorffDistance: the larger of the two one-sided maxima."""
        return max(self.maxDistance(m1, m2), self.maxDistance(m2, m1))

    def evaluateReconstruction2GroundTruth(self, reconstruction, groundTruth) -> Tuple[float, float, float]:
        """:38-49 -> (average to surface, one-sided max, Hausdorff)."""
        d12 = self.distances(reconstruction, groundTruth)
        d21 = self.distances(groundTruth, reconstruction)
        return float(d12.mean()), float(d12.max()), float(max(d12.max(), d21.max()))

    def evaluateReconstruction2GroundTruthDouble(self, reconstruction, groundTruth) -> Tuple[float, float]:
        """:51-62 -> (mean of the two one-sided averages, Hausdorff)."""
        d12 = self.distances(reconstruction, groundTruth)
        d21 = self.distances(groundTruth, reconstruction)
        return float((d12.mean() + d21.mean()) / 2.0), float(max(d12.max(), d21.max()))

    def avgDistanceBoundaryAware(self, m1, m2) -> Tuple[float, float]:
        """:64-75 mean and max over the vertices of m1 whose closest surface point on m2 is not nearest to a boundary
        vertex of m2.  (nan, -inf) when every vertex is filtered out -- the reference divides by zero / takes the max of
        an empty sequence there."""
        p = np.asarray(m1[0], dtype=np.float64).reshape(-1, 3)
        cp, idx = self._closest(m1, m2)
        n2 = np.asarray(m2[0]).reshape(-1, 3).shape[0]
        keep = ~boundary_vertices(n2, m2[1])[idx]
        d = _norms(cp - p)[keep]
        if d.size == 0:
            return float("nan"), float("-inf")
        return float(d.sum() / d.size), float(d.max())

    def evaluateReconstruction2GroundTruthBoundaryAware(self, reconstruction, groundTruth) -> Tuple[float, float]:
        """:77-86 -> (mean of the two boundary-aware averages, larger of the two maxima); the reference prints them."""
        a1, m1 = self.avgDistanceBoundaryAware(reconstruction, groundTruth)
        a2, m2 = self.avgDistanceBoundaryAware(groundTruth, reconstruction)
        return (a1 + a2) / 2.0, max(m1, m2)
