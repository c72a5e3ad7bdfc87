"""gingr_b200/textbook_icp.py: the reference's textbook rigid ICP with the closest-point search on the device.  The device
entry point (gingr_icp_closest, point-cloud flavour: GPU parity in tests/test_closest_gpu.py) is replaced by the oracle's
nearest-vertex scan here; the loop is compared with a literal restatement of RigidICP.scala:33-83."""
import numpy as np
import pytest


def _install(monkeypatch, oracle):
    from gingr_b200 import api

    class FakeTarget:
        def __init__(self, ctx, pts, tri=None):
            self.points = np.ascontiguousarray(np.asarray(pts, float))

        def close(self):
            pass

    def icp_closest(ctx, target, pts, tri, method):
        assert method == api.POINTCLOUD_CLOSEST_POINT and tri is None
        idx, d2 = oracle.nearest_vertex(np.ascontiguousarray(pts), target.points)
        return idx, target.points[idx].copy(), np.ones(len(idx), np.uint8), float(np.sum(np.sqrt(d2)) / len(idx))
    monkeypatch.setattr(api, "Target", FakeTarget)
    monkeypatch.setattr(api, "icp_closest", icp_closest)


def _literal(oracle, template, target, similarity, max_iteration, tolerance=0.001):
    fit, last, i, converged = template.copy(), 0.0, 0, False
    while i < max_iteration and not converged:
        cps, dist = [], 0.0
        for p in fit:
            j = int(np.argmin(((target - p) ** 2).sum(1)))
            cps.append(target[j])
            dist += np.linalg.norm(p - target[j])
        dist /= len(fit)
        R, t, s = oracle.umeyama(fit, np.array(cps), similarity)      # the oracle's carries the Euler round trip: 1e-15 apart
        TY = s * (fit @ R.T) + t
        if abs(dist - last) < tolerance:
            converged = True
        fit, last = TY, dist
        i += 1
    return fit, last, i


@pytest.mark.parametrize("registrator", ["rigid", "similarity"])
def test_rigid_icp_equals_the_literal_loop(oracle, monkeypatch, registrator):
    import textbook_icp
    _install(monkeypatch, oracle)
    rng = np.random.default_rng(3)
    tpl = rng.normal(size=(120, 3)) * np.array([10.0, 6.0, 3.0])
    Rz = np.array([[np.cos(0.15), -np.sin(0.15), 0], [np.sin(0.15), np.cos(0.15), 0], [0, 0, 1.0]])
    scale = 1.08 if registrator == "similarity" else 1.0
    tgt = scale * (tpl @ Rz.T) + np.array([0.8, -0.5, 0.3])
    task = textbook_icp.RigidICP(None, tpl, tgt, registrator)
    got = task.Registration(60)
    want, dist, iters = _literal(oracle, tpl, tgt, registrator == "similarity", 60)
    assert task.iterations == iters and task.converged and abs(task.distance - dist) < 1e-9
    assert np.max(np.abs(got - want)) < 1e-9 * 10.0
    assert np.sqrt(((got - tgt) ** 2).sum(1)).mean() < 0.05 * np.sqrt(((tpl - tgt) ** 2).sum(1)).mean()
    out = textbook_icp.RigidICPRegistration(None, tpl, tgt, 60, registrator)
    assert np.array_equal(out, got)
    with pytest.raises(ValueError):
        textbook_icp.RigidICP(None, tpl, tgt, "affine")
