"""K2 parity: closest-point correspondence through the C ABI against the O(M N) oracle scan.
Indices and 0/1 weights must be EXACT (ties -> lowest index); points are compared bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cloud(n, seed, scale=50.0):
    return np.random.default_rng(seed).normal(scale=scale, size=(n, 3))


@pytest.mark.parametrize("M,N", [(1, 1), (5, 700), (700, 5), (1000, 1000), (3000, 20011)])
def test_pointcloud_nearest_vertex_exact(ctx, oracle, M, N):
    from gingr_b200 import api
    q, p = _cloud(M, 1), _cloud(N, 2)
    target = api.Target(ctx, p)
    idx, cp, w, md = api.icp_closest(ctx, target, q, None, api.POINTCLOUD_CLOSEST_POINT)
    ridx, rd2 = oracle.nearest_vertex(q, p)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(cp, p[ridx])
    assert np.all(w == 1)
    assert abs(md - np.sqrt(rd2).mean()) < 1e-12 * max(md, 1e-300)


def test_ties_break_to_lowest_index(ctx, oracle):
    from gingr_b200 import api
    # integer lattice with duplicated points: many exact ties
    g = np.stack(np.meshgrid(np.arange(8.0), np.arange(8.0), np.arange(8.0), indexing="ij"), -1).reshape(-1, 3)
    p = np.concatenate([g, g[::-1], g])              # every point three times
    q = g + 0.5                                       # equidistant to 8 lattice corners
    target = api.Target(ctx, p)
    idx, cp, w, md = api.icp_closest(ctx, target, q, None, api.POINTCLOUD_CLOSEST_POINT)
    ridx, _ = oracle.nearest_vertex(q, p)
    assert np.array_equal(idx, ridx)
    assert idx.max() < len(g)                         # always the first copy


@pytest.mark.parametrize("M,N", [(100, 100), (500, 400), (1000, 1000)])
def test_triangular_closest_point_exact(ctx, oracle, M, N):
    from gingr_b200 import api, synthetic
    tv, tt = synthetic.sphere_mesh(M, radius=97.0)
    gv, gt = synthetic.sphere_mesh(N)
    gv = synthetic.make_target(gv, 3)
    target = api.Target(ctx, gv, gt)
    idx, cp, w, md = api.icp_closest(ctx, target, tv, tt, api.TRIANGULAR_CLOSEST_POINT)
    rcp, rw, rmd, ridx = oracle.closest_point_correspondence(oracle.METHOD_TRIANGULAR, tv, tt, gv, gt)
    assert np.array_equal(cp, rcp)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(w.astype(float), rw)
    assert abs(md - rmd) < 1e-12 * rmd


def test_triangular_weights_open_mesh_and_flipped_normals(ctx, oracle):
    """Boundary vertices, opposite normals and self-intersections all produce w = 0 somewhere."""
    from gingr_b200 import api, synthetic
    tv, tt = synthetic.sphere_mesh(400, radius=95.0)
    gv, gt = synthetic.sphere_mesh(500)
    gt_open = gt[~np.any(gt < 25, axis=1)]            # cut a hole -> boundary vertices
    # fold part of the template inward so that some closest-point segments cross the template itself
    tv2 = tv.copy()
    sel = tv2[:, 2] > 60
    tv2[sel, 2] = 120 - tv2[sel, 2] * 1.3
    target = api.Target(ctx, gv, gt_open)
    idx, cp, w, md = api.icp_closest(ctx, target, tv2, tt, api.TRIANGULAR_CLOSEST_POINT)
    rcp, rw, rmd, ridx = oracle.closest_point_correspondence(oracle.METHOD_TRIANGULAR, tv2, tt, gv, gt_open)
    assert np.array_equal(cp, rcp) and np.array_equal(idx, ridx)
    assert np.array_equal(w.astype(float), rw)
    assert 0 < w.sum() < len(w)


@pytest.mark.parametrize("M,N", [(100, 120), (600, 500)])
def test_along_normal_closest_point_exact(ctx, oracle, M, N):
    """ClosestPointAlongNormalTriangleMesh3D (ClosestPointRegistrator.scala:98-131): nearest intersection of the
    vertex-normal line with the target mesh, (p, 0.0) when it misses."""
    from gingr_b200 import api, synthetic
    tv, tt = synthetic.sphere_mesh(M, radius=90.0)
    gv, gt = synthetic.sphere_mesh(N)
    gv = synthetic.make_target(gv, 5)
    gt_open = gt[~np.any(gt < N // 6, axis=1)]        # a hole: some normal lines miss the target
    target = api.Target(ctx, gv, gt_open)
    idx, cp, w, md = api.icp_closest(ctx, target, tv, tt, api.ALONG_NORMAL_CLOSEST_POINT)
    rcp, rw, rmd, ridx = oracle.closest_point_correspondence(oracle.METHOD_ALONG_NORMAL, tv, tt, gv, gt_open)
    assert np.array_equal(cp, rcp)
    assert np.array_equal(w.astype(float), rw)
    keep = rw == 1.0
    assert np.array_equal(idx[keep], ridx[keep])
    assert abs(md - rmd) <= 1e-12 * rmd
    assert 0 < w.sum() < len(w)


@pytest.mark.parametrize("method", ["triangular", "along_normal", "pointcloud"])
def test_correspondence_reversal_exact(ctx, oracle, method):
    """closestPointCorrespondenceReversal (ClosestPointRegistrator.scala:34-45): search from the target, map back to
    the nearest template vertex."""
    from gingr_b200 import api, synthetic
    om = {"triangular": oracle.METHOD_TRIANGULAR, "along_normal": oracle.METHOD_ALONG_NORMAL,
          "pointcloud": oracle.METHOD_POINTCLOUD}[method]
    gm = {"triangular": api.TRIANGULAR_CLOSEST_POINT, "along_normal": api.ALONG_NORMAL_CLOSEST_POINT,
          "pointcloud": api.POINTCLOUD_CLOSEST_POINT}[method]
    tv, tt = synthetic.sphere_mesh(300, radius=94.0)
    gv, gt = synthetic.sphere_mesh(420)
    gv = synthetic.make_target(gv, 6)
    target = api.Target(ctx, gv, gt)
    tid, w, md = api.icp_closest_reversal(ctx, target, tv, tt, gm)
    # oracle: roles swapped, then template.findClosestPoint(p).id
    rcp, rw, rmd, _ = oracle.closest_point_correspondence(om, gv, gt, tv, tt)
    rtid, _ = oracle.nearest_vertex(rcp, tv)
    assert np.array_equal(w.astype(float), rw)
    assert np.array_equal(tid, rtid)
    assert abs(md - rmd) <= 1e-12 * max(rmd, 1e-300)
    pids, pts = oracle.icp_correspondence(om, True, tv, tt, gv, gt)
    assert np.array_equal(pids, tid[w == 1]) and np.array_equal(pts, gv[w == 1])
