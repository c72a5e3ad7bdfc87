"""oracle/fast.py (the algorithmic-minimum CPU form of update(), used where the literal restatement takes minutes per
iteration: the benchmark's reference arm and the C4-shaped parity test) must be the SAME function as oracle.update /
oracle.propose, the literal restatement of GingrAlgorithm.scala:192-254: every state component within 1e-9 (relative to
the mesh diagonal / max |alpha|) over several iterations, for each transformation type, with and without outliers and
with a damped step."""
import dataclasses

import numpy as np
import pytest


def _problem(oracle, M, N, r, seed):
    from gingr_b200 import synthetic
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, seed + 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, seed)
    return oracle.Gpmm(ref, mean, basis, var, tri), target, tt


@pytest.mark.parametrize("gt,w,step,literal", [("rigid", 0.1, 1.0, True), ("rigid", 0.0, 1.0, False), ("similarity", 0.2, 1.0, False),
                                               ("none", 0.1, 0.5, False)])
def test_fast_update_equals_the_literal_update(oracle, gt, w, step, literal):
    from oracle import fast
    om, target, tt = _problem(oracle, 260, 330, 24, seed=3)
    g = {"rigid": oracle.RIGID_TRANSFORMS, "similarity": oracle.SIMILARITY_TRANSFORMS, "none": oracle.NO_TRANSFORMS}[gt]
    algo = oracle.CpdAlgorithm(oracle.CpdConfig(w=w), literal=literal)
    st_l = algo.initialize(oracle.initial_state(om, target, tt, global_transformation=g))
    st_l = dataclasses.replace(st_l, step_length=step)
    st_f = st_l
    fm = fast.FastCpdModel(om)
    diag = float(np.linalg.norm(om.ref.max(0) - om.ref.min(0)))
    for _ in range(4):
        st_l = oracle.propose(algo, st_l)
        st_f = fast.propose(fm, algo, st_f)
        assert st_f.status == st_l.status and st_f.iteration == st_l.iteration
        assert np.max(np.abs(st_f.fit - st_l.fit)) < 1e-9 * diag
        assert np.max(np.abs(st_f.params.shape - st_l.params.shape)) < 1e-9 * max(1.0, np.max(np.abs(st_l.params.shape)))
        assert abs(st_f.sigma2 - st_l.sigma2) < 1e-9 * st_l.sigma2
        assert np.max(np.abs(st_f.params.translation - st_l.params.translation)) < 1e-9 * diag
        assert np.max(np.abs(np.array(st_f.params.euler) - np.array(st_l.params.euler))) < 1e-9
        assert abs(st_f.params.scale - st_l.params.scale) < 1e-9


def test_fast_update_reports_the_same_failure(oracle):
    """P1 = 0 rows (a target far away at a tiny sigma2) make the uncertainty infinite: the literal update returns the state
    unchanged at iteration 0 and ModelFlexibilityError later (GingrAlgorithm.scala:194-208); so must the reduced form."""
    from oracle import fast
    om, target, tt = _problem(oracle, 120, 150, 10, seed=1)
    algo = oracle.CpdAlgorithm(oracle.CpdConfig(w=0.0, initial_sigma=1e-9), literal=False)
    st = algo.initialize(oracle.initial_state(om, target + 1e4, tt))
    fm = fast.FastCpdModel(om)
    a, b = oracle.update(algo, st), fast.update(fm, algo, st)
    assert a.status == b.status == oracle.STATUS_NONE and np.array_equal(a.params.shape, b.params.shape)
    st1 = dataclasses.replace(st, iteration=3)
    a, b = oracle.update(algo, st1), fast.update(fm, algo, st1)
    assert a.status == b.status == oracle.STATUS_MODEL_FLEXIBILITY_ERROR
