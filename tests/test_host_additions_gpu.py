"""GPU parity tests for the host-side callers of the device path (template plugin, mesh distances, SimpleRegistrator
options): each composes device entry points and is checked against the oracle / brute force."""
import dataclasses
import os

import numpy as np
import pytest

from test_update_gpu import _problem

pytestmark = pytest.mark.gpu


def test_template_registration_matches_oracle_update(ctx, oracle):
    from gingr_b200 import api, template
    om, target, tt = _problem(oracle, 150, 170, 14, seed=0)
    dm = api.Model(ctx, om.ref, om.mean, om.basis, om.variance, om.tri)

    class OAlgo:
        config = oracle.CpdConfig()

        def observations(self, st):
            idx, _ = oracle.nearest_vertex(st.fit, st.target)
            pids = np.arange(st.model.M, dtype=np.int32)[::2]
            return pids, st.target[idx[::2]], (1.0 + 0.01 * pids)[:, None, None] * np.eye(3)[None]

        def update_sigma2(self, st):
            return st.sigma2

    def correspondence(state):
        idx, _, _, _ = api.icp_closest(ctx, dt, state.fit, None, api.POINTCLOUD_CLOSEST_POINT)      # device search in the closure
        return np.arange(om.M, dtype=np.int32)[::2], target[idx[::2]]
    dt = api.Target(ctx, target, tt)
    reg = template.TemplateRegistration(ctx, dm, dt, template.TemplateConfiguration(maxIterations=4),
                                        getCorrespondence=correspondence, getUncertainty=lambda pids, s: 1.0 + 0.01 * pids)
    st = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    ost = oracle.initial_state(om, target, tt, global_transformation=oracle.RIGID_TRANSFORMS)
    oalgo = OAlgo()
    diag = float(np.linalg.norm(om.ref.max(0) - om.ref.min(0)))
    for _ in range(3):
        st = reg.propose(st)
        ost = oracle.propose(oalgo, ost)
        assert np.max(np.abs(st.fit - ost.fit)) < 1e-6 * diag
        assert np.max(np.abs(st.modelParameters.shape - ost.params.shape)) < 1e-6 * max(1.0, np.max(np.abs(ost.params.shape)))
    dm.close(); dt.close()


def test_registration_comparison_matches_brute_force(ctx, oracle):
    from gingr_b200 import comparison, synthetic
    v1, t1 = synthetic.sphere_mesh(150)
    v2, t2 = synthetic.sphere_mesh(190)
    v1 = v1 * 1.04 + np.array([1.0, 0.5, -0.5])
    rc = comparison.RegistrationComparison(ctx)
    cp12 = oracle.closest_on_surface(v1, v2, t2)[0]
    cp21 = oracle.closest_on_surface(v2, v1, t1)[0]
    d12, d21 = np.linalg.norm(v1 - cp12, axis=1), np.linalg.norm(v2 - cp21, axis=1)
    a, mx, h = rc.evaluateReconstruction2GroundTruth((v1, t1), (v2, t2))
    assert abs(a - d12.mean()) < 1e-12 and abs(mx - d12.max()) < 1e-12 and abs(h - max(d12.max(), d21.max())) < 1e-12
    avg, m = rc.evaluateReconstruction2GroundTruthBoundaryAware((v1, t1), (v2, t2))     # closed meshes: nothing filtered
    assert abs(avg - (d12.mean() + d21.mean()) / 2) < 1e-12 and abs(m - max(d12.max(), d21.max())) < 1e-12


def test_simple_registrator_options_on_the_device(ctx, oracle, tmp_path):
    from gingr_b200 import api, io
    om, target, tt = _problem(oracle, 400, 450, 16, seed=2)
    dm = api.Model(ctx, om.ref, om.mean, om.basis, om.variance, om.tri)
    dt = api.Target(ctx, target, tt)
    mlm = [io.Landmark("a", om.ref[10], None), io.Landmark("b", om.ref[200], None)]
    tlm = [io.Landmark("a", target[12], None), io.Landmark("b", target[220], None)]
    log = str(tmp_path / "chain.json")
    gi = api.GingrInterface(ctx, dm, dt, modelLandmarks=mlm, targetLandmarks=tlm, evaluatorUncertainty=2.0, evaluatedPoints=60,
                            logFileFittingParameters=log)
    det = gi.CPD(api.CpdConfiguration(maxIterations=10, w=0.05)).runDecimated(120, 150)
    assert det.fit.shape == (400, 3) and det.status in (api.STATUS_MAX_ITERATION, api.STATUS_CONVERGED)
    seen = []
    sr = gi.ICP(api.IcpConfiguration(maxIterations=15, initialSigma=2.0, endSigma=0.5))
    pro = sr.runDecimated(120, 150, generalState=det, probabilistic=True, callback=seen.append, seed=3)
    assert pro.fit.shape == (400, 3) and len(seen) == 15 and sr.jsonLogger.totalSamples == 15
    assert len(io.JSONStateLogger.load(log)) == 15
    dm.close(); dt.close()
