"""K2 uniform-grid searches (gingr_b200/csrc/grid.cu) against the brute-force scans (closest.cu) on the same inputs:
every output must be BIT-IDENTICAL (the grid only selects candidates; winners are chosen by (value, lowest index) with
the shared arithmetic of closest_geom.cuh), and against the CPU oracle where it finishes in seconds.
GINGR_K2_GRID=0/1 forces the scan / the grid (read by the library on every call)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class grid_mode:
    def __init__(self, v):
        self.v = v

    def __enter__(self):
        self.old = os.environ.get("GINGR_K2_GRID")
        os.environ["GINGR_K2_GRID"] = str(self.v)

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("GINGR_K2_GRID", None)
        else:
            os.environ["GINGR_K2_GRID"] = self.old


def _both(fn):
    """fn() with the scan and with the grid (handles are created inside fn so the static grids follow the mode)."""
    with grid_mode(0):
        a = fn()
    with grid_mode(1):
        b = fn()
    return a, b


def _assert_identical(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        x, y = np.asarray(x), np.asarray(y)
        assert x.shape == y.shape
        assert np.array_equal(x, y, equal_nan=True), f"max diff {np.max(np.abs(x.astype(float) - y.astype(float)))}"


def _cloud(n, seed, scale=50.0):
    return np.random.default_rng(seed).normal(scale=scale, size=(n, 3))


@pytest.mark.parametrize("M,N", [(1, 1), (7, 900), (900, 7), (2000, 3000), (5000, 40000)])
def test_nn_grid_equals_scan_and_oracle(ctx, oracle, M, N):
    from gingr_b200 import api
    q, p = _cloud(M, 1), _cloud(N, 2)

    def run():
        t = api.Target(ctx, p)
        out = api.icp_closest(ctx, t, q, None, api.POINTCLOUD_CLOSEST_POINT)
        t.close()
        return out
    a, b = _both(run)
    _assert_identical(a, b)
    if M * N <= 2e7:
        ridx, _ = oracle.nearest_vertex(q, p)
        assert np.array_equal(b[0], ridx)


def test_nn_grid_ties_duplicates_and_far_queries(ctx, oracle):
    from gingr_b200 import api
    g = np.stack(np.meshgrid(np.arange(10.0), np.arange(10.0), np.arange(10.0), indexing="ij"), -1).reshape(-1, 3)
    p = np.concatenate([g, g[::-1], g])                      # every lattice point three times
    q = np.concatenate([g + 0.5,                             # equidistant to 8 corners
                        g[:50] * 1e4 + 1e5,                  # far outside the grid (fallback scan)
                        g[:50] - 300.0,
                        np.array([[4.5, 4.5, -1e3], [1e8, 0, 0], [4.5, -77.0, 4.5]])])

    def run():
        t = api.Target(ctx, p)
        out = api.icp_closest(ctx, t, q, None, api.POINTCLOUD_CLOSEST_POINT)
        t.close()
        return out
    a, b = _both(run)
    _assert_identical(a, b)
    ridx, _ = oracle.nearest_vertex(q, p)
    assert np.array_equal(b[0], ridx)
    assert b[0].max() < len(g)


@pytest.mark.parametrize("kind", ["one_point", "coincident", "planar", "line", "anisotropic"])
def test_nn_grid_degenerate_point_sets(ctx, oracle, kind):
    from gingr_b200 import api
    rng = np.random.default_rng(5)
    if kind == "one_point":
        p = np.array([[1.0, 2.0, 3.0]])
    elif kind == "coincident":
        p = np.tile([[1.0, 2.0, 3.0]], (500, 1))
    elif kind == "planar":
        p = np.c_[rng.uniform(-50, 50, (4000, 2)), np.zeros(4000)]
    elif kind == "line":
        p = np.c_[rng.uniform(-50, 50, 3000), np.zeros(3000), np.full(3000, 7.0)]
    else:
        p = rng.normal(size=(6000, 3)) * np.array([1000.0, 1.0, 1e-3])
    q = np.concatenate([p[:200] + rng.normal(scale=0.3, size=(min(200, len(p)), 3)), _cloud(300, 3, 80.0)])

    def run():
        t = api.Target(ctx, p)
        out = api.icp_closest(ctx, t, q, None, api.POINTCLOUD_CLOSEST_POINT)
        t.close()
        return out
    a, b = _both(run)
    _assert_identical(a, b)
    ridx, _ = oracle.nearest_vertex(q, p)
    assert np.array_equal(b[0], ridx)


def _mesh_pair(M, N, seed=3, radius=97.0):
    from gingr_b200 import synthetic
    tv, tt = synthetic.sphere_mesh(M, radius=radius)
    gv, gt = synthetic.sphere_mesh(N)
    gv = synthetic.make_target(gv, seed)
    return tv, tt, gv, gt


@pytest.mark.parametrize("method", ["TRIANGULAR_CLOSEST_POINT", "ALONG_NORMAL_CLOSEST_POINT"])
@pytest.mark.parametrize("M,N", [(300, 400), (3000, 2500), (20000, 30000)])
def test_mesh_flavours_grid_equals_scan(ctx, oracle, method, M, N):
    from gingr_b200 import api
    tv, tt, gv, gt = _mesh_pair(M, N)
    meth = getattr(api, method)

    def run():
        t = api.Target(ctx, gv, gt)
        out = api.icp_closest(ctx, t, tv, tt, meth)
        t.close()
        return out
    a, b = _both(run)
    _assert_identical(a, b)
    assert 0 < b[2].sum() <= M
    if M <= 3000:
        om = oracle.METHOD_TRIANGULAR if method.startswith("TRI") else oracle.METHOD_ALONG_NORMAL
        rcp, rw, rmd, ridx = oracle.closest_point_correspondence(om, tv, tt, gv, gt)
        assert np.array_equal(b[1], rcp) and np.array_equal(b[0], ridx) and np.array_equal(b[2].astype(float), rw)


def test_mesh_flavours_grid_open_folded_mesh(ctx):
    """Boundary vertices, opposite normals and self-intersections (w = 0 cases) agree between grid and scan."""
    from gingr_b200 import api, synthetic
    tv, tt = synthetic.sphere_mesh(4000, radius=95.0)
    gv, gt = synthetic.sphere_mesh(5000)
    gt_open = gt[~np.any(gt < 200, axis=1)]
    tv = tv.copy()
    cap = tv[:, 2] > 60.0
    tv[cap, 2] = 120.0 - tv[cap, 2] * 1.3                    # fold the cap inward through the template itself
    for meth in (api.TRIANGULAR_CLOSEST_POINT, api.ALONG_NORMAL_CLOSEST_POINT):
        def run():
            t = api.Target(ctx, gv, gt_open)
            out = api.icp_closest(ctx, t, tv, tt, meth)
            t.close()
            return out
        a, b = _both(run)
        _assert_identical(a, b)
        assert 0 < b[2].sum() < len(tv)


@pytest.mark.parametrize("method", ["POINTCLOUD_CLOSEST_POINT", "TRIANGULAR_CLOSEST_POINT", "ALONG_NORMAL_CLOSEST_POINT"])
def test_reversal_grid_equals_scan(ctx, method):
    from gingr_b200 import api
    tv, tt, gv, gt = _mesh_pair(2500, 3500)
    meth = getattr(api, method)

    def run():
        t = api.Target(ctx, gv, gt)
        out = api.icp_closest_reversal(ctx, t, tv, tt, meth)
        t.close()
        return out
    a, b = _both(run)
    _assert_identical(a, b)


def test_oversized_triangles_fall_back_to_the_scan(ctx):
    """A few triangles spanning the whole mesh overflow the per-triangle cell budget: the grid marks itself overflowed
    and every query scans all triangles -- still identical."""
    from gingr_b200 import api
    tv, tt, gv, gt = _mesh_pair(1500, 6000)
    gt = np.concatenate([gt, np.array([[0, len(gv) // 2, len(gv) - 1], [1, len(gv) // 3, len(gv) - 2]], dtype=np.int32)])

    def run():
        t = api.Target(ctx, gv, gt)
        out = api.icp_closest(ctx, t, tv, tt, api.TRIANGULAR_CLOSEST_POINT)
        t.close()
        return out
    a, b = _both(run)
    _assert_identical(a, b)


@pytest.mark.parametrize("method,reverse", [("TRIANGULAR_CLOSEST_POINT", False), ("POINTCLOUD_CLOSEST_POINT", False),
                                            ("ALONG_NORMAL_CLOSEST_POINT", False), ("TRIANGULAR_CLOSEST_POINT", True)])
def test_icp_update_with_grids_is_bit_identical(ctx, method, reverse):
    """Whole ICP iterations (captured graph, grids over the moving fit rebuilt inside it) with and without grids."""
    from gingr_b200 import api, synthetic
    M, N, r = 1500, 2000, 40
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, 0)

    def run():
        model = api.Model(ctx, ref, mean, basis, var, tri)
        tgt = api.Target(ctx, target, tt)
        cfg = api.IcpConfiguration(initialSigma=2.0, endSigma=0.5, reverseCorrespondenceDirection=reverse,
                                   correspondenceMethod=getattr(api, method))
        reg = api.IcpRegistration(ctx, model, tgt, cfg)
        st = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
        for _ in range(2):
            st = reg.propose(st)
        reg.updateChain(3)
        fin = reg.downloadState()
        out = (st.fit, st.modelParameters.shape, fin.fit, fin.modelParameters.shape, np.array([fin.sigma2, fin.status]))
        reg.close(); model.close(); tgt.close()
        return out
    a, b = _both(run)
    _assert_identical(a, b)
    assert np.all(np.isfinite(b[2]))


@pytest.mark.parametrize("method", ["TRIANGULAR_CLOSEST_POINT", "POINTCLOUD_CLOSEST_POINT"])
def test_reversed_fold_by_lists_equals_the_scan(ctx, method):
    """closestPointCorrespondenceReversal: the O(N + M) list form of the per-template-vertex fold (large problems) gives the
    bits of the O(M N) scan (GINGR_RFOLD_LIST forces either)."""
    from gingr_b200 import api, synthetic
    M, N, r = 1200, 2600, 24
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, 0)
    outs = []
    for forced in ("0", "1"):
        os.environ["GINGR_RFOLD_LIST"] = forced
        try:
            model = api.Model(ctx, ref, mean, basis, var, tri)
            tgt = api.Target(ctx, target, tt)
            cfg = api.IcpConfiguration(initialSigma=2.0, endSigma=0.5, reverseCorrespondenceDirection=True,
                                       correspondenceMethod=getattr(api, method))
            reg = api.IcpRegistration(ctx, model, tgt, cfg)
            reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
            reg.updateChain(3)
            st = reg.downloadState()
            outs.append((st.fit, st.modelParameters.shape, np.array([st.sigma2])))
            reg.close(); model.close(); tgt.close()
        finally:
            os.environ.pop("GINGR_RFOLD_LIST", None)
    _assert_identical(outs[0], outs[1])
    assert np.all(np.isfinite(outs[1][0]))
