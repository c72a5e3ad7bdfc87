"""gingr_b200/helper.py: the reference's log / posterior helpers (api/helper/LogHelper.scala, PosteriorHelper.scala,
JSONStateLogger companion), against literal restatements.  No GPU."""
import dataclasses

import numpy as np
import pytest


def _log(statuses, products):
    from gingr_b200 import io
    out = []
    for k, (s, p) in enumerate(zip(statuses, products)):
        out.append(io.JsonLogRecord(k, "CPD", {"Prior": p / 2, "Distance": p / 2, "product": p}, s,
                                    [0.1 * k, 1.0] if s else [], [1.0, 2.0, 3.0 + k] if s else [], [0.1, 0.2, 0.3] if s else [],
                                    [0.0, 0.0, 0.0] if s else [], 1.0, "2024-01-01 00:00:00"))
    return out


def test_samples_from_log_thinning_and_fallback_to_last_accepted():
    from gingr_b200 import helper
    st = [True, False, False, True, False, True, False, False, False, True, True, False]
    log = _log(st, [-float(k) for k in range(len(st))])
    got = helper.samples_from_log(log, takeEveryN=2, total=100, burnIn=1)
    # literal: indices 1, 3, 5, 7, 9, 11 -> last accepted at or before each
    assert [j for _, j in got] == [0, 3, 5, 5, 9, 10]
    assert all(r is log[j] for r, j in got)
    assert [j for _, j in helper.samples_from_log(log, takeEveryN=50, total=100, burnIn=0)] == [0]
    assert [j for _, j in helper.samples_from_log(log, takeEveryN=3, total=7, burnIn=2)] == [0, 5]      # range below min(len, total)
    assert helper.samples_from_log(log, burnIn=50) == []


def test_best_record_and_parameters():
    from gingr_b200 import helper
    log = _log([True, True, False, True], [-5.0, -1.0, -9.0, -1.0])
    assert helper.best_record(log) is log[3]                     # ties: sortBy(...).reverse.head = the last of the equal ones
    p = helper.record_to_parameters(log[3])
    assert p.scale == 1.0 and tuple(p.euler) == (0.1, 0.2, 0.3) and list(p.translation) == [1.0, 2.0, 6.0]
    assert list(p.shape) == [0.1 * 3, 1.0]
    with pytest.raises(ValueError):
        helper.record_to_parameters(log[2])                      # rejected record: `require` fails in the reference
    with pytest.raises(ValueError):
        helper.best_record([])


def test_log_samples_to_shapes_uses_model_instance():
    from gingr_b200 import helper
    log = _log([True, True], [-1.0, -2.0])

    class FakeModel:
        def instance(self, p):
            return np.full((4, 3), p.translation[2])
    shapes = helper.log_samples_to_shapes(FakeModel(), log)
    assert [s[0, 0] for s in shapes] == [3.0, 4.0]


def test_vertex_normals_match_the_oracle(oracle):
    from gingr_b200 import helper, synthetic
    v, t = synthetic.sphere_mesh(60)
    v = v + np.random.default_rng(0).normal(size=v.shape)
    assert np.max(np.abs(helper.vertex_normals(v, t) - oracle.vertex_normals(v, t))) < 1e-14


def test_distance_maps_against_literal_loops(oracle):
    from gingr_b200 import helper, synthetic
    rng = np.random.default_rng(1)
    v, t = synthetic.sphere_mesh(40)
    meshes = [v + rng.normal(size=v.shape) for _ in range(7)]
    tot = helper.distance_map_total(meshes)
    nor = helper.distance_map_normal(meshes, t)
    nor_ref = helper.distance_map_normal(meshes, t, reference=v, sumNormals=False)
    n = len(meshes)
    normals = [oracle.vertex_normals(m, t) for m in meshes]
    ref_normals = oracle.vertex_normals(v, t)
    for pid in (0, 13, 39):
        s = np.array([m[pid] for m in meshes])
        mean = s.sum(0) / n
        cov = sum(np.outer(x - mean, x - mean) for x in s) / (n - 1)
        assert abs(tot[pid] - np.trace(cov)) < 1e-12
        nn = sum(nm[pid] for nm in normals) / n
        assert abs(nor[pid] - sum(np.dot(nn, x - mean) ** 2 for x in s) / (n - 1)) < 1e-12
        assert abs(nor_ref[pid] - sum(np.dot(ref_normals[pid], x - mean) ** 2 for x in s) / (n - 1)) < 1e-12
    assert np.all(nor <= tot + 1e-12)            # |mean of unit normals| <= 1: the projected variance cannot exceed the trace
    with pytest.raises(ValueError):
        helper.distance_map_normal(meshes, t, sumNormals=False)


def test_simple_logger_callback_prints_every_nth_state(tmp_path):
    from gingr_b200 import api, helper, io
    lg = io.JSONStateLogger(path=str(tmp_path / "l.json"))
    p = api.ModelFittingParameters(1.0, np.zeros(3), (0.0, 0.0, 0.0), np.zeros(2))
    s = api.GeneralRegistrationState(p, np.zeros((3, 3)), generatedBy="ICP")
    lines = []

    class FakeComparison:
        def evaluateReconstruction2GroundTruthBoundaryAware(self, a, b):
            return 0.5, 2.0
    cb = helper.SimpleLogger(lg, printUpdateFrequency=3, comparison=FakeComparison(), fit_triangles=None, target_mesh=(None, None),
                             out=lines.append)
    for k in range(7):
        (lg.accept if k % 2 == 0 else lg.reject)(s, {"Prior": -1.0, "Distance": -1.0})
        cb(s)
    assert cb.counter == 7 and len(lines) == 6            # states 3 and 6: total, one generator, distance
    assert lines[0] == "Total accepted (3): " + repr(1.0 - 0.33) and lines[1] == "ICP: " + repr(2 / 3) and lines[2] == "average2surface: 0.5 max: 2.0"
    assert len(io.JSONStateLogger.load(lg.path)) == 6     # rewritten at state 6
