"""gingr_b200/template.py -- the reference's TemplateRegistration extension point (user closures for correspondence and
uncertainty inside GingrAlgorithm.update).  Its GPMM operations are kernel-level device entry points that have their own
GPU parity tests (tests/test_posterior_gpu.py); here they are replaced by oracle-backed stand-ins so that the statement
sequence, the landmark handling, the Procrustes step and the failure rules are checked against oracle.update on the CPU."""
import dataclasses

import numpy as np
import pytest


class _OracleAlgo:
    """A user-defined algorithm on the oracle side: nearest target vertex as correspondence, variance 1 + 0.01 pid."""
    def __init__(self, oracle, use_landmarks=True):
        self.oracle = oracle
        self.config = oracle.CpdConfig(use_landmark_correspondence=use_landmarks)

    def observations(self, st):
        idx, _ = self.oracle.nearest_vertex(st.fit, st.target)
        pids = np.arange(st.model.M, dtype=np.int32)[::2]
        cov = (1.0 + 0.01 * pids)[:, None, None] * np.eye(3)[None]
        return pids, st.target[idx[::2]], cov

    def update_sigma2(self, st):
        return st.sigma2 * 0.9


def _install_stand_ins(monkeypatch, oracle, om):
    from gingr_b200 import api

    def posterior_mean(ctx, model, R, t, pids, points, noise):
        noise = np.asarray(noise, dtype=float)
        cov = noise[:, None, None] * np.eye(3)[None] if noise.ndim == 1 else noise
        posed = om.transform(np.asarray(R), np.asarray(t))
        c, _ = posed.posterior_coefficients(np.asarray(pids), np.asarray(points), cov)
        if not np.all(np.isfinite(c)):
            raise FloatingPointError("posterior")
        return c, posed.instance(c)

    def coefficients(ctx, model, R, t, mesh):
        return om.transform(np.asarray(R), np.asarray(t)).coefficients(np.asarray(mesh))

    class FakeModel:
        M, rank = om.M, om.rank

        def instance(self, p):
            return oracle.model_instance_shape_pose_scale(om, oracle.Params(p.scale, np.asarray(p.translation, float), tuple(p.euler),
                                                                             np.asarray(p.shape, float)))
    monkeypatch.setattr(api, "posterior_mean", posterior_mean)
    monkeypatch.setattr(api, "coefficients", coefficients)
    return FakeModel()


def _problem(oracle, seed=0):
    from gingr_b200 import synthetic
    ref, tri = synthetic.sphere_mesh(80)
    mean, basis, var = synthetic.make_gpmm(ref, 10, seed)
    om = oracle.Gpmm(ref, mean, basis, var, tri)
    tv, tt = synthetic.sphere_mesh(90)
    return om, synthetic.make_target(tv, seed), tt


@pytest.mark.parametrize("gt_name,step,with_lm", [("RIGID_TRANSFORMS", 1.0, False), ("SIMILARITY_TRANSFORMS", 0.5, True),
                                                  ("NO_TRANSFORMS", 1.0, True)])
def test_template_iterations_equal_the_oracle(oracle, monkeypatch, gt_name, step, with_lm):
    from gingr_b200 import api, template
    om, target, tt = _problem(oracle)
    fm = _install_stand_ins(monkeypatch, oracle, om)
    gt = getattr(api, gt_name)
    lm = None
    if with_lm:
        lp = np.array([0, 7, 20], dtype=np.int32)                      # 0 and 20 are also correspondence ids: they must be replaced
        lpts = target[[3, 11, 40]]
        A = np.random.default_rng(1).normal(size=(3, 3, 3))
        lcov = A @ np.transpose(A, (0, 2, 1)) + 0.5 * np.eye(3)[None]
        lm = oracle.Landmarks(lp, lpts, lcov)
    oalgo = _OracleAlgo(oracle)
    ost = oracle.initial_state(om, target, tt, global_transformation=getattr(oracle, gt_name), landmarks=lm)
    ost = dataclasses.replace(ost, step_length=step)

    def correspondence(state):
        idx, _ = oracle.nearest_vertex(state.fit, target)              # any host logic may live here
        return np.arange(om.M, dtype=np.int32)[::2], target[idx[::2]]

    reg = template.TemplateRegistration(None, fm, None, template.TemplateConfiguration(maxIterations=5),
                                        getCorrespondence=correspondence, getUncertainty=lambda pids, state: 1.0 + 0.01 * pids,
                                        updateSigma2=lambda state: state.sigma2 * 0.9)
    if with_lm:
        reg.setLandmarks(lm.pids, lm.points, lm.cov)
    st = dataclasses.replace(reg.initializeState(globalTransformation=gt), stepLength=step)
    assert np.allclose(st.fit, ost.fit, atol=1e-12)
    for it in range(3):
        st = reg.propose(st)
        ost = oracle.propose(oalgo, ost)
        assert st.iteration == ost.iteration == it + 1 and st.status == ost.status == 0
        assert np.max(np.abs(st.modelParameters.shape - ost.params.shape)) < 1e-9 * max(1.0, np.max(np.abs(ost.params.shape)))
        assert np.max(np.abs(st.fit - ost.fit)) < 1e-9 * 200.0
        assert abs(st.modelParameters.scale - ost.params.scale) < 1e-12 and abs(st.sigma2 - ost.sigma2) < 1e-15
        assert np.allclose(st.modelParameters.euler, ost.params.euler, atol=1e-12)
    if gt_name == "NO_TRANSFORMS":
        assert st.modelParameters.scale == 1.0 and tuple(st.modelParameters.euler) == (0.0, 0.0, 0.0)
    if gt_name == "SIMILARITY_TRANSFORMS":
        assert st.modelParameters.scale != 1.0


def test_template_defaults_and_failure_rules(oracle, monkeypatch):
    from gingr_b200 import api, template
    om, target, tt = _problem(oracle, seed=1)
    fm = _install_stand_ins(monkeypatch, oracle, om)
    cfg = template.TemplateConfiguration()
    assert (cfg.maxIterations, cfg.threshold, cfg.useLandmarkCorrespondence) == (1, 1e-5, True)          # Template.scala:24-29
    assert cfg.converged(None, None, 1.0) is False and template.TemplateRegistration.name == "Template"
    # default closures: no correspondences -> the posterior cannot be formed; iteration 0 returns the state unchanged
    # (GingrAlgorithm.scala:206-208), a later iteration flags ModelFlexibilityError (:204)
    reg = template.TemplateRegistration(None, fm)
    st0 = reg.initializeState()
    assert reg.update(st0) is st0
    later = dataclasses.replace(st0, iteration=3)
    assert reg.update(later).status == api.STATUS_MODEL_FLEXIBILITY_ERROR
    # run(): maxIterations = 1 yields the initial state only, status MaxIteration
    out = reg.run(st0)
    assert out.status == api.STATUS_MAX_ITERATION and out.iteration == 0
    # non-finite uncertainty (e.g. sigma2 / 0) is the same failure
    reg2 = template.TemplateRegistration(None, fm, getCorrespondence=lambda s: (np.arange(5), target[:5]),
                                         getUncertainty=lambda pids, s: np.full(len(pids), np.inf))
    assert reg2.update(later).status == api.STATUS_MODEL_FLEXIBILITY_ERROR
    # a failing coefficients call is the same status (:247-252)
    def boom(*a):
        raise FloatingPointError("x")
    reg3 = template.TemplateRegistration(None, fm, getCorrespondence=lambda s: (np.arange(5), target[:5]))
    monkeypatch.setattr(api, "coefficients", boom)
    assert reg3.update(st0).status == api.STATUS_MODEL_FLEXIBILITY_ERROR
    with pytest.raises(ValueError):
        template.TemplateRegistration(None, fm, getCorrespondence=lambda s: (np.arange(4), target[:5])).update(st0)


def test_umeyama_equals_the_oracle(oracle):
    from gingr_b200 import template
    rng = np.random.default_rng(2)
    X = rng.normal(size=(60, 3)) * 30.0
    Y = 1.2 * (X @ oracle.euler_to_matrix(0.3, -0.1, 0.2).T) + np.array([4.0, 5.0, -6.0]) + rng.normal(size=X.shape)
    for sim in (False, True):
        R, t, s = template.umeyama(X, Y, sim)
        Ro, to, so = oracle.umeyama(X, Y, sim)
        assert np.array_equal(R, Ro) and np.array_equal(t, to) and s == so
