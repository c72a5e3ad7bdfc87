"""gingr_b200/decimate.py (the stated stand-in for scalismo's decimate): size, validity, topology, determinism.  No GPU."""
import numpy as np
import pytest


def _edge_counts(t):
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]).astype(np.int64)
    e.sort(axis=1)
    return np.unique(e, axis=0, return_counts=True)[1]


def _sheet(n):
    g = np.arange(n)
    x, y = np.meshgrid(g, g, indexing="ij")
    v = np.c_[x.ravel(), y.ravel(), 0.3 * np.sin(0.7 * x.ravel()) * np.cos(0.5 * y.ravel())].astype(np.float64)
    t = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j, (i + 1) * n + j, i * n + j + 1, (i + 1) * n + j + 1
            t += [[a, b, c], [b, d, c]]
    return v, np.array(t, dtype=np.int32)


@pytest.mark.parametrize("M,n", [(2000, 100), (2000, 500), (5000, 1000), (300, 40)])
def test_closed_mesh_stays_a_closed_manifold_with_exactly_n_vertices(M, n):
    from gingr_b200 import decimate, synthetic
    v, t = synthetic.sphere_mesh(M)
    v = v * np.array([1.0, 0.7, 1.6])                                   # an ellipsoid: not every edge has the same length
    dv, dt = decimate.decimate(v, t, n)
    assert len(dv) == n
    idx = np.array([np.flatnonzero((v == q).all(1))[0] for q in dv])    # kept vertices are input vertices, in input order
    assert np.all(np.diff(idx) > 0)
    assert dt.dtype == np.int32 and dt.min() >= 0 and dt.max() < n and len(np.unique(dt)) == n
    assert np.all((dt[:, 0] != dt[:, 1]) & (dt[:, 1] != dt[:, 2]) & (dt[:, 0] != dt[:, 2]))
    assert np.all(_edge_counts(dt) == 2)                                # closed 2-manifold: every edge in exactly two triangles
    assert len(dt) == 2 * n - 4                                         # Euler: a sphere keeps genus 0
    # orientation preserved (the ellipsoid is convex: outward normals have a positive component along the centroid), no slivers
    c = dv[dt].mean(1)
    nrm = np.cross(dv[dt[:, 1]] - dv[dt[:, 0]], dv[dt[:, 2]] - dv[dt[:, 0]])
    assert np.all((nrm * c).sum(1) > 0)
    e2 = sum(((dv[dt[:, i]] - dv[dt[:, (i + 1) % 3]]) ** 2).sum(1) for i in range(3))
    assert np.min(2 * np.sqrt(3) * np.linalg.norm(nrm, axis=1) / e2) > 0.2
    dv2, dt2 = decimate.decimate(v, t, n)
    assert np.array_equal(dv, dv2) and np.array_equal(dt, dt2)          # deterministic


def test_open_sheet_keeps_its_boundary_and_shape():
    from gingr_b200 import decimate
    from gingr_b200.comparison import boundary_vertices
    v, t = _sheet(30)
    dv, dt = decimate.decimate(v, t, 150)
    assert len(dv) == 150
    cnt = _edge_counts(dt)
    assert set(np.unique(cnt)) <= {1, 2} and (cnt == 1).sum() > 0       # still a manifold with one boundary loop
    b = boundary_vertices(len(dv), dt)
    on_rim = (dv[:, 0] == 0) | (dv[:, 0] == 29) | (dv[:, 1] == 0) | (dv[:, 1] == 29)
    assert np.array_equal(b, on_rim)                                    # boundary vertices are exactly the kept rim vertices
    nz = np.cross(dv[dt[:, 1]] - dv[dt[:, 0]], dv[dt[:, 2]] - dv[dt[:, 0]])[:, 2]
    assert np.all(nz > 0)                                               # no fold-over of the height field
    area = 0.5 * np.linalg.norm(np.cross(dv[dt[:, 1]] - dv[dt[:, 0]], dv[dt[:, 2]] - dv[dt[:, 0]]), axis=1).sum()
    area0 = 0.5 * np.linalg.norm(np.cross(v[t[:, 1]] - v[t[:, 0]], v[t[:, 2]] - v[t[:, 0]]), axis=1).sum()
    assert 0.9 * area0 < area <= area0


def test_point_cloud_and_edge_cases():
    from gingr_b200 import decimate, synthetic
    v, t = synthetic.sphere_mesh(100)
    same_v, same_t = decimate.decimate(v, t, 100)
    assert np.array_equal(same_v, v) and np.array_equal(same_t, t)
    assert np.array_equal(decimate.decimate(v, t, 10 ** 6)[0], v)
    big = synthetic.fibonacci_sphere(3000)
    cloud_v, cloud_t = decimate.decimate(big, None, 300)
    assert cloud_t.shape == (0, 3) and abs(len(cloud_v) - 300) <= 30
    ids = decimate.decimate_points(big, 300)
    assert np.all(np.diff(ids) > 0) and np.array_equal(big[ids], cloud_v)
    one_v, one_t = decimate.decimate(np.ones((50, 3)), None, 5)
    assert one_v.shape == (1, 3) and one_t.shape == (0, 3)
    tetra_v = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    tetra_t = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]], np.int32)
    tv, tt = decimate.decimate(tetra_v, tetra_t, 2)                     # nothing smaller than a tetrahedron
    assert len(tv) == 4 and len(tt) == 4
    with pytest.raises(ValueError):
        decimate.decimate(v, t, 0)
    with pytest.raises(ValueError):
        decimate.decimate(v, t + 50, 10)


def test_torus_keeps_its_genus():
    """A closed surface of genus 1: V - E + F stays 0, every edge stays in two triangles, no fold-over against the input
    normals (checked through the signed volume, which must keep its sign and most of its size)."""
    from gingr_b200 import decimate
    nu, nv, R, r = 60, 24, 10.0, 3.0
    u, v = np.meshgrid(np.arange(nu) * 2 * np.pi / nu, np.arange(nv) * 2 * np.pi / nv, indexing="ij")
    pts = np.c_[((R + r * np.cos(v)) * np.cos(u)).ravel(), ((R + r * np.cos(v)) * np.sin(u)).ravel(), (r * np.sin(v)).ravel()]
    tri = []
    for i in range(nu):
        for j in range(nv):
            a, b = i * nv + j, ((i + 1) % nu) * nv + j
            c, d = i * nv + (j + 1) % nv, ((i + 1) % nu) * nv + (j + 1) % nv
            tri += [[a, b, c], [b, d, c]]
    tri = np.array(tri, dtype=np.int32)

    def volume(p, t):
        return float(np.einsum("ij,ij->i", p[t[:, 0]], np.cross(p[t[:, 1]], p[t[:, 2]])).sum() / 6.0)
    v0 = volume(pts, tri)
    for n in (600, 250):
        dv, dt = decimate.decimate(pts, tri, n)
        cnt = _edge_counts(dt)
        assert len(dv) == n and np.all(cnt == 2)
        assert len(dv) - len(cnt) + len(dt) == 0                       # Euler characteristic of a torus
        v1 = volume(dv, dt)
        assert np.sign(v1) == np.sign(v0) and 0.8 * abs(v0) < abs(v1) <= abs(v0) * 1.001


def test_arbitrary_triangle_soups_never_break_the_decimator():
    """Non-manifold edges, duplicated vertices, degenerate and repeated triangles, isolated vertices: the output is always a
    valid (possibly undecimated) mesh over a subset of the input vertices."""
    from gingr_b200 import decimate
    rng = np.random.default_rng(0)
    for trial in range(200):
        n, T = int(rng.integers(4, 40)), int(rng.integers(1, 80))
        v = rng.normal(size=(n, 3))
        if trial % 3 == 0:
            v = np.round(v, 0)
        t = rng.integers(0, n, size=(T, 3)).astype(np.int32)
        if trial % 4:
            t = t[(t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2])]
        if len(t) == 0:
            continue
        dv, dt = decimate.decimate(v, t, int(rng.integers(1, n + 2)))
        assert len(dv) <= n and np.all(np.isfinite(dv))
        assert dt.size == 0 or (dt.min() >= 0 and dt.max() < len(dv))
        assert all(any((v == q).all(1)) for q in dv)


def test_native_twin_equals_the_python_specification():
    """csrc_host/decimate.cpp against decimate._collapse: same kept vertices and the same triangles, bit for bit, on closed,
    open, genus-1 and arbitrary inputs."""
    from gingr_b200 import decimate, synthetic
    lib = decimate._native_lib()
    if lib is None:
        pytest.skip("no host C++ compiler")
    cases = []
    v, t = synthetic.sphere_mesh(1500)
    cases += [(v * np.array([1.0, 0.7, 1.6]), t, 200), (v, t, 37)]
    sv, st = _sheet(25)
    cases += [(sv, st, 120)]
    rng = np.random.default_rng(4)
    for _ in range(60):
        n = int(rng.integers(5, 40))
        tv = np.round(rng.normal(size=(n, 3)), int(rng.integers(0, 3)))
        tt = rng.integers(0, n, size=(int(rng.integers(2, 90)), 3)).astype(np.int32)
        tt = tt[(tt[:, 0] != tt[:, 1]) & (tt[:, 1] != tt[:, 2]) & (tt[:, 0] != tt[:, 2])]
        if len(tt):
            cases.append((tv, tt, int(rng.integers(1, n))))
    for pts, tri, n in cases:
        p = np.ascontiguousarray(pts, dtype=np.float64)
        tr = np.ascontiguousarray(tri, dtype=np.int32)
        a = decimate._collapse_native(lib, p, tr, n)
        b = decimate._collapse(p, tr, n)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (len(p), len(tr), n)
