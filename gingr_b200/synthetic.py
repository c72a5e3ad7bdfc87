"""Seeded synthetic inputs of the shapes named in BASELINE.json / SURVEY.md 8(d): Fibonacci-sphere meshes,
smoothly displaced targets and low-rank GPMMs.  There is no network for datasets; bench.py and the tests
use these generators (and the two femur STL fixtures under tests/golden where present)."""
from __future__ import annotations

import numpy as np


def fibonacci_sphere(n: int, radius: float = 100.0) -> np.ndarray:
    """n quasi-uniform points on a sphere."""
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = np.pi * (1 + 5 ** 0.5) * i
    return radius * np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)


def sphere_mesh(n: int, radius: float = 100.0):
    """Fibonacci sphere with its convex-hull triangulation (closed, consistently outward oriented)."""
    from scipy.spatial import ConvexHull
    pts = fibonacci_sphere(n, radius)
    hull = ConvexHull(pts)
    tri = hull.simplices.astype(np.int32)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    return pts, tri


def euler_matrix(phi, theta, psi):
    from .rotation import euler_to_matrix
    return euler_to_matrix(phi, theta, psi)


def smooth_displacement(points: np.ndarray, rng: np.random.Generator, n_waves: int = 8, amp: float = 3.0,
                        wavelength: float = 50.0) -> np.ndarray:
    """sum_k a_k sin(w_k . p + phi_k),  a_k ~ N(0, amp^2) per axis, |w_k| ~ 1/wavelength."""
    out = np.zeros_like(points)
    for _ in range(n_waves):
        w = rng.normal(size=3)
        w *= (1.0 / wavelength) / np.linalg.norm(w)
        a = rng.normal(scale=amp, size=3)
        ph = rng.uniform(0, 2 * np.pi)
        out += np.sin(points @ w + ph)[:, None] * a[None, :]
    return out


def make_target(points: np.ndarray, seed: int, t=(5.0, 5.0, 5.0), euler=(0.05, 0.05, 0.05)) -> np.ndarray:
    rng = np.random.default_rng(seed)
    R = euler_matrix(*euler)
    return (points + smooth_displacement(points, rng)) @ R.T + np.asarray(t)


def make_gpmm(ref: np.ndarray, rank: int, seed: int, lambda0: float = 100.0, decay: float = 0.995,
              kernel_sigma: float = 60.0, orthonormal: bool = True):
    """(mean = 0, basis [3M, r], variance [r]).  Basis columns: smooth vector fields built from Gaussian-kernel
    features at `rank` random centres (Nystrom style), QR-orthonormalised; for large M a seeded Gaussian matrix
    is used before QR.  variance_k = lambda0 * decay^k."""
    rng = np.random.default_rng(seed)
    M = ref.shape[0]
    r = int(rank)
    if M <= 20000 and r <= 600:
        nc = max(1, (r + 2) // 3)
        centres = ref[rng.choice(M, size=min(nc, M), replace=nc > M)]
        d2 = ((ref[:, None, :] - centres[None, :, :]) ** 2).sum(-1)
        feat = np.exp(-d2 / (2 * kernel_sigma ** 2))                      # [M, nc]
        B = np.zeros((3 * M, 3 * feat.shape[1]))
        for d in range(3):
            B[d::3, d::3] = feat
        B = B[:, :r] if B.shape[1] >= r else np.concatenate([B, rng.normal(size=(3 * M, r - B.shape[1]))], axis=1)
        B = B + 1e-3 * rng.normal(size=B.shape)
    else:
        B = rng.normal(size=(3 * M, r))
    if orthonormal:
        B, _ = np.linalg.qr(B)
    else:
        B /= np.linalg.norm(B, axis=0, keepdims=True)
    variance = lambda0 * decay ** np.arange(r)
    return np.zeros(3 * M), np.ascontiguousarray(B), variance


def workload(name: str, seed: int = 0):
    """Named workloads.  Returns dict(ref, tri, mean, basis, variance, target, target_tri)."""
    sizes = {
        "c1": (100, 100, 50), "c2": (100, 100, 50), "c3a": (100, 100, 50), "c3b": (500, 500, 100),
        "c3c": (1000, 1000, 100), "c4": (20000, 200000, 2000), "c4_small": (4000, 20000, 500),
    }
    M, N, r = sizes[name]
    ref, tri = sphere_mesh(M)
    tgt_pts, tgt_tri = sphere_mesh(N) if N <= 50000 else (fibonacci_sphere(N), None)
    target = make_target(tgt_pts, seed)
    mean, basis, variance = make_gpmm(ref, r, seed + 1)
    return dict(ref=ref, tri=tri, mean=mean, basis=basis, variance=variance, target=target, target_tri=tgt_tri)
