"""gingr_b200/textbook_nicp.py: the reference's optimal-step non-rigid ICP (N-ICP-T / N-ICP-A) with the robust surface
correspondences on the device.  The device entry point (gingr_icp_closest, triangular flavour: GPU parity in
tests/test_closest_gpu.py, tests/test_grid_gpu.py) is replaced by the oracle's closest_point_correspondence; every
iteration's sparse solve is compared with a dense literal least-squares restatement of the Scala statements."""
import numpy as np
import pytest


def _install(monkeypatch, oracle):
    from gingr_b200 import api

    class FakeTarget:
        def __init__(self, ctx, pts, tri=None):
            self.points, self.tri = np.ascontiguousarray(np.asarray(pts, float)), np.asarray(tri, np.int32)

        def close(self):
            pass

    def icp_closest(ctx, target, pts, tri, method):
        assert method == api.TRIANGULAR_CLOSEST_POINT
        out = oracle.closest_point_correspondence(oracle.METHOD_TRIANGULAR, np.ascontiguousarray(pts), tri, target.points, target.tri)
        cp, w, dist = out[0], out[1], out[2]
        return np.zeros(len(cp), np.int32), cp, np.asarray(w).astype(np.uint8), dist
    monkeypatch.setattr(api, "Target", FakeTarget)
    monkeypatch.setattr(api, "icp_closest", icp_closest)


def _meshes():
    from gingr_b200 import io, synthetic
    tv, tt = synthetic.sphere_mesh(60)
    sv, st = synthetic.sphere_mesh(80)
    sv = sv * np.array([1.1, 0.95, 1.05]) + np.array([2.0, -1.0, 1.5]) + 2.0 * np.sin(sv[:, [1, 2, 0]] / 40.0)
    tl = [io.Landmark("a", tv[5] + 0.3, None), io.Landmark("b", tv[30] - 0.2, None), io.Landmark("only", tv[1], None)]
    sl = [io.Landmark("b", sv[41] + 0.1, None), io.Landmark("a", sv[7] - 0.1, None)]
    return (tv, tt), (sv, st), tl, sl


def _edges(tri):
    t = np.sort(tri.astype(np.int64), axis=1)
    return np.unique(np.concatenate([t[:, [0, 1]], t[:, [0, 2]], t[:, [1, 2]]]), axis=0)


def _literal_T(task, template, cp, w, alpha, beta):
    n, E, L = task.n, task.numOfEdges, len(task.lmIdsOnTemplate)
    M = np.zeros((E, n))
    for i, (p, q) in enumerate(_edges(task.triangles)):
        M[i, p], M[i, q] = 1.0, -1.0
    W = np.diag(w)
    A3 = np.zeros((L, n))
    for i in range(L):
        A3[i, i] = 1.0
    A = np.vstack([M * alpha, W @ np.eye(n), A3])
    B = np.vstack([np.zeros((E, 3)), W @ (cp - template), (task.UL - template[task.lmIdsOnTemplate]) * beta])
    return template + np.linalg.lstsq(A, B, rcond=None)[0]


def _literal_A(task, template, cp, w, alpha, beta, gamma):
    n, E, L = task.n, task.numOfEdges, len(task.lmIdsOnTemplate)
    M = np.zeros((E, n))
    for i, (p, q) in enumerate(_edges(task.triangles)):
        M[i, p], M[i, q] = 1.0, -1.0
    w = w.copy()
    w[task.lmIdsOnTemplate] = 0.0
    D = np.zeros((n, 4 * n))
    for i in range(n):
        D[i, 4 * i:4 * i + 3], D[i, 4 * i + 3] = template[i], 1.0
    DL = np.zeros((L, 4 * n))
    for i, pid in enumerate(task.lmIdsOnTemplate):
        DL[i, 4 * pid:4 * pid + 3], DL[i, 4 * pid + 3] = template[pid], 1.0
    A = np.vstack([np.kron(M, np.diag([1.0, 1.0, 1.0, gamma])) * alpha, np.diag(w) @ D, DL * beta])
    B = np.vstack([np.zeros((4 * E, 3)), np.diag(w) @ cp, task.UL * beta])
    return D @ np.linalg.lstsq(A, B, rcond=None)[0]


@pytest.mark.parametrize("variant", ["T", "A"])
def test_nicp_iterations_equal_dense_least_squares(oracle, monkeypatch, variant):
    import textbook_nicp
    _install(monkeypatch, oracle)
    tpl, tgt, tl, sl = _meshes()
    cls = textbook_nicp.NonRigidOptimalStepICP_T if variant == "T" else textbook_nicp.NonRigidOptimalStepICP_A
    task = cls(None, tpl, tgt, tl, sl, gamma=0.7)
    assert list(task.lmIdsOnTemplate) == [5, 30] and task.UL.shape == (2, 3)      # matched by id, in template-landmark order
    assert np.array_equal(task.UL[0], tgt[0][7]) and np.array_equal(task.UL[1], tgt[0][41])
    assert task.numOfEdges == len(_edges(tpl[1])) == 3 * 60 - 6                   # closed genus-0 mesh: E = 3V - 6
    fit = tpl[0].copy()
    for alpha, beta in ((10.0, 10.0), (3.0, 1.0), (1.0, 0.0)):
        cp, w, dist = task.getClosestPoints(fit)
        got, d2 = task.Iteration(fit, alpha, beta)[:2]
        want = _literal_T(task, fit, cp, w, alpha, beta) if variant == "T" else _literal_A(task, fit, cp, w, alpha, beta, 0.7)
        assert d2 == dist and 0 < w.sum() <= len(w)
        assert np.max(np.abs(got - want)) < 1e-7 * 100.0, (variant, alpha)
        fit = got
    # the schedule: defaults are eleven times 10.0; the loop stops on the mean-distance tolerance
    assert textbook_nicp.DEFAULT_ALPHA == [10.0] * 11 == textbook_nicp.DEFAULT_BETA
    out = task.Registration(2, alpha=[10.0, 2.0], beta=[1.0, 1.0])
    assert task.iterations == 4 and out.shape == tpl[0].shape
    d0 = np.sqrt(((tpl[0][:, None] - tgt[0][None]) ** 2).sum(-1).min(1)).mean()
    d1 = np.sqrt(((out[:, None] - tgt[0][None]) ** 2).sum(-1).min(1)).mean()
    assert d1 < d0
    with pytest.raises(ValueError):
        task.Registration(1, alpha=[1.0], beta=[1.0, 2.0])
    with pytest.raises(ValueError):
        task.Iteration(fit, -1.0, 0.0)
    with pytest.raises(ValueError):
        cls(None, tpl, tgt, tl, sl, gamma=-1.0)
