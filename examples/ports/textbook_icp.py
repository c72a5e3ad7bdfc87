"""The textbook rigid ICP of the reference (other/algorithms/icp/{ICPFactory, RigidICP}.scala, other/utils/PoseRegistrator.scala,
wrapper other/algorithms/RigidICPRegistration.scala) over the device closest-point search.

An iteration is: nearest target VERTEX of every template point and the mean of those distances (RigidICP.scala:62-75) --
gingr_icp_closest with the point-cloud flavour, the exact K2 search (uniform grid at scale) --, then the rigid or similarity
landmark registration of the pairs (PoseRegistrator.scala:30-42) applied to the template: an O(M) reduction and a 3 x 3 SVD
on the host.  Quirks kept: the convergence test compares the mean distance with the previous one (initially 0), and the
iteration counter advances on the converging step too (RigidICP.scala:36-58)."""
from __future__ import annotations

import numpy as np

from gingr_b200 import api
from gingr_b200.template import umeyama


class RigidICP:
    """RigidICP.scala:25-83.  registrator: "rigid" (RigidRegistrator3D) or "similarity" (the reference's AffineRegistrator3D,
    which is scalismo's similarity3DLandmarkRegistration)."""

    def __init__(self, ctx: "api.Context", templatePoints, targetPoints, registrator: str = "rigid"):
        if registrator not in ("rigid", "similarity"):
            raise ValueError("registrator: 'rigid' or 'similarity'")
        self.ctx, self.registrator = ctx, registrator
        self.template = np.ascontiguousarray(np.asarray(templatePoints, dtype=np.float64).reshape(-1, 3))
        self._dev_target = api.Target(ctx, np.asarray(targetPoints, dtype=np.float64).reshape(-1, 3))
        self.iterations, self.distance, self.converged = 0, 0.0, False

    def close(self):
        self._dev_target.close()

    def attributeCorrespondences(self, template):
        """-> (closest target vertex per template point [M, 3], mean distance)  (:62-75)."""
        _, cp, _, mean_distance = api.icp_closest(self.ctx, self._dev_target, template, None, api.POINTCLOUD_CLOSEST_POINT)
        return cp, mean_distance

    def Iteration(self, template):
        """:77-83"""
        cp, distance = self.attributeCorrespondences(template)
        R, t, s = umeyama(template, cp, self.registrator == "similarity", euler_round_trip=False)
        return s * (template @ R.T) + t, distance

    def Registration(self, max_iteration: int, tolerance: float = 0.001) -> np.ndarray:
        """:33-60 -> the registered template points."""
        fit, last = self.template.copy(), 0.0
        i, converged = 0, False
        while i < max_iteration and not converged:
            TY, distance = self.Iteration(fit)
            if abs(distance - last) < tolerance:
                converged = True
            fit, last = TY, distance
            i += 1
        self.iterations, self.distance, self.converged = i, last, converged
        return fit


def RigidICPRegistration(ctx, template, target, max_iterations: int = 100, registrator: str = "rigid") -> np.ndarray:
    """RigidICPRegistration.register (other/algorithms/RigidICPRegistration.scala:23-45): the registered template points."""
    task = RigidICP(ctx, template, target, registrator)
    try:
        return task.Registration(max_iterations)
    finally:
        task.close()
