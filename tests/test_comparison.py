"""api/helper/RegistrationComparison.scala mirror (gingr_b200/comparison.py): the reductions around the device search,
checked on the CPU with the K2 entry point replaced by a brute-force stand-in (oracle closest-point-on-surface)."""
import numpy as np
import pytest


def _open_sheet(n=6):
    """(n x n) height-field sheet: has a boundary; triangles split each cell in two."""
    g = np.arange(n)
    x, y = np.meshgrid(g, g, indexing="ij")
    v = np.c_[x.ravel(), y.ravel(), 0.1 * np.sin(x.ravel() + y.ravel())].astype(np.float64)
    t = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j, (i + 1) * n + j, i * n + j + 1, (i + 1) * n + j + 1
            t += [[a, b, c], [b, d, c]]
    return v, np.array(t, dtype=np.int32)


def test_boundary_vertices_of_sheet_and_closed_mesh(oracle):
    from gingr_b200 import comparison, synthetic
    v, t = _open_sheet(6)
    b = comparison.boundary_vertices(len(v), t)
    want = np.array([(i in (0, 5)) or (j in (0, 5)) for i in range(6) for j in range(6)])
    assert np.array_equal(b, want)
    assert np.array_equal(b, oracle.boundary_vertices(len(v), t))
    sv, stri = synthetic.sphere_mesh(80)
    assert not comparison.boundary_vertices(len(sv), stri).any()
    # drop some triangles: the holes' rims become boundary, as the oracle's edge count says
    holed = np.delete(stri, [3, 17, 40], axis=0)
    assert np.array_equal(comparison.boundary_vertices(len(sv), holed), oracle.boundary_vertices(len(sv), holed))


def test_metrics_with_a_brute_force_search(oracle, monkeypatch):
    from gingr_b200 import api, comparison
    rng = np.random.default_rng(4)
    v2, t2 = _open_sheet(7)
    v1, t1 = _open_sheet(5)
    v1 = v1 * 1.2 + np.array([0.3, 0.2, 0.5]) + 0.05 * rng.normal(size=v1.shape)

    class FakeTarget:
        def __init__(self, ctx, pts, tri):
            self.pts, self.tri = np.asarray(pts, float), np.asarray(tri, np.int32)

        def close(self):
            pass

    def fake_icp_closest(ctx, target, points, tri, method):
        assert method == api.TRIANGULAR_CLOSEST_POINT and tri is not None
        cp = oracle.closest_on_surface(points, target.pts, target.tri)[0]
        idx = oracle.nearest_vertex(cp, target.pts)[0]
        return idx.astype(np.int32), cp, np.ones(len(cp), np.uint8), 0.0

    monkeypatch.setattr(api, "Target", FakeTarget)
    monkeypatch.setattr(api, "icp_closest", fake_icp_closest)
    rc = comparison.RegistrationComparison(None)
    m1, m2 = (v1, t1), (v2, t2)
    cp12 = oracle.closest_on_surface(v1, v2, t2)[0]
    cp21 = oracle.closest_on_surface(v2, v1, t1)[0]
    d12, d21 = np.linalg.norm(v1 - cp12, axis=1), np.linalg.norm(v2 - cp21, axis=1)
    assert abs(rc.avgDistance(m1, m2) - d12.mean()) < 1e-14 and abs(rc.maxDistance(m1, m2) - d12.max()) < 1e-14
    assert abs(rc.hausdorffDistance(m1, m2) - max(d12.max(), d21.max())) < 1e-14
    a, mx, h = rc.evaluateReconstruction2GroundTruth(m1, m2)
    assert (abs(a - d12.mean()), abs(mx - d12.max()), abs(h - max(d12.max(), d21.max()))) < (1e-14, 1e-14, 1e-14)
    a2, h2 = rc.evaluateReconstruction2GroundTruthDouble(m1, m2)
    assert abs(a2 - (d12.mean() + d21.mean()) / 2) < 1e-14 and h2 == h
    # boundary aware: literal loop of RegistrationComparison.scala:64-75
    def literal(pa, pb, tb):
        bnd = oracle.boundary_vertices(len(pb), tb)
        cp = oracle.closest_on_surface(pa, pb, tb)[0]
        ds = [np.linalg.norm(c - p) for p, c in zip(pa, cp) if not bnd[int(np.argmin(((pb - c) ** 2).sum(1)))]]
        return (sum(ds) / len(ds), max(ds)) if ds else (float("nan"), float("-inf"))
    l1, l2 = literal(v1, v2, t2), literal(v2, v1, t1)
    g1 = rc.avgDistanceBoundaryAware(m1, m2)
    assert abs(g1[0] - l1[0]) < 1e-13 and abs(g1[1] - l1[1]) < 1e-13 and g1[1] <= d12.max()
    avg, mx = rc.evaluateReconstruction2GroundTruthBoundaryAware(m1, m2)
    if np.isfinite(l2[0]):
        assert abs(avg - (l1[0] + l2[0]) / 2) < 1e-13 and abs(mx - max(l1[1], l2[1])) < 1e-13
    # everything filtered: a single triangle is all boundary
    tri_mesh = (np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]]), np.array([[0, 1, 2]], np.int32))
    assert np.isnan(rc.avgDistanceBoundaryAware(m1, tri_mesh)[0])
