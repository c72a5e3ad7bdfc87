"""Euler-angle helpers in scalismo's convention (Rotation(phi, theta, psi, center): R = Rz(phi) Ry(theta) Rx(psi);
RotationSpace3D.rotMatrixToEulerAngles), used at api/GeneralRegistrationState.scala:83-87, :143 and
api/ModelFittingParameters.scala:40.  Host-side convenience only; the device has its own copy."""
import math

import numpy as np


def euler_to_matrix(phi: float, theta: float, psi: float) -> np.ndarray:
    cph, sph = math.cos(phi), math.sin(phi)
    cth, sth = math.cos(theta), math.sin(theta)
    cps, sps = math.cos(psi), math.sin(psi)
    return np.array([
        [cth * cph, sps * sth * cph - cps * sph, sps * sph + cps * sth * cph],
        [cth * sph, cps * cph + sps * sth * sph, cps * sth * sph - sps * cph],
        [-sth, sps * cth, cps * cth],
    ])


def matrix_to_euler(R: np.ndarray):
    if abs(abs(R[2, 0]) - 1) > 0.0001:
        theta = math.asin(-R[2, 0])
        ct = math.cos(theta)
        return math.atan2(R[1, 0] / ct, R[0, 0] / ct), theta, math.atan2(R[2, 1] / ct, R[2, 2] / ct)
    if abs(R[2, 0] + 1) < 0.0001:
        return 0.0, math.pi / 2.0, math.atan2(R[0, 1], R[0, 2])
    return 0.0, -math.pi / 2.0, math.atan2(-R[0, 1], -R[0, 2])
