"""SimpleRegistrator.decimateState / runDecimated on the device: model.newReference with the
nearest-neighbour interpolator (rows gathered from the resident basis) and the multi-resolution hand-over of
examples/DemoMultiResolution.scala:39-47, against the oracle."""
import dataclasses

import numpy as np
import pytest

from test_update_gpu import _compare, _problem, _to_api_state

pytestmark = pytest.mark.gpu


def _decimated(n, radius=100.0):
    from gingr_b200 import synthetic
    return synthetic.sphere_mesh(n, radius=radius)


@pytest.mark.parametrize("M,M2", [(400, 150), (3000, 700), (20000, 9000)])
def test_new_reference_matches_oracle(ctx, oracle, M, M2):
    from gingr_b200 import api
    m, target, tt = _problem(oracle, M, 300, 24, seed=0)
    dv, dtri = _decimated(M2)
    dv = dv + 0.3 * np.random.default_rng(1).normal(size=dv.shape)       # off the old vertices
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dm2 = dm.newReference(dv, dtri)
    om2 = m.new_reference(dv, dtri)
    # instance(alpha) exposes reference, mean rows and basis rows of the new model at once
    alpha = np.random.default_rng(2).normal(size=24)
    pars = api.ModelFittingParameters(1.0, np.array([1.0, -2.0, 0.5]), (0.1, -0.2, 0.05), alpha)
    got = dm2.instance(pars)
    want = oracle.model_instance_shape_pose_scale(om2, oracle.Params(1.0, np.array([1.0, -2.0, 0.5]), (0.1, -0.2, 0.05), alpha))
    assert np.max(np.abs(got - want)) < 1e-11 * 100.0
    # one full update on the re-referenced (non-orthonormal) model: S / W0 were rebuilt
    dt = api.Target(ctx, target, tt)
    reg = api.IcpRegistration(ctx, dm2, dt, api.IcpConfiguration(initialSigma=2.0, endSigma=0.5))
    oalgo = oracle.IcpAlgorithm(oracle.IcpConfig(initial_sigma=2.0, end_sigma=0.5))
    if M2 <= 1000:
        ost = oalgo.initialize(oracle.initial_state(om2, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
        gst = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
        for _ in range(2):
            gst = reg.propose(gst)
            ost = oracle.propose(oalgo, ost)
        _compare(gst, ost, float(np.linalg.norm(dv.max(0) - dv.min(0))))
    reg.close(); dm2.close(); dm.close(); dt.close()


def test_multi_resolution_schedule_matches_oracle(ctx, oracle):
    """DemoMultiResolution: CPD coarse -> CPD finer -> ICP fine, handing (pose, scale, shape) over; sigma2 is
    re-initialised from the config at every level, the final fit is evaluated on the full model."""
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 600, 700, 20, seed=3)
    diag = float(np.linalg.norm(m.ref.max(0) - m.ref.min(0)))
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    dt = api.Target(ctx, target, tt)
    levels = [("cpd", 100, 120, 8), ("cpd", 250, 300, 6), ("icp", 500, 600, 6)]
    ogeneral, ggeneral = None, None
    for algo, nm, nt, iters in levels:
        dref, dtgt = _decimated(nm), _decimated(nt)
        dtgt = (dtgt[0] * 1.0 + (target.mean(0) - dtgt[0].mean(0)), dtgt[1])       # a coarse stand-in for the decimated target
        if algo == "cpd":
            cfg, ocfg = api.CpdConfiguration(maxIterations=iters, w=0.05), oracle.CpdConfig(max_iterations=iters, w=0.05)
            cls, oalgo = api.CpdRegistration, oracle.CpdAlgorithm(ocfg, literal=False)
        else:
            cfg = api.IcpConfiguration(maxIterations=iters, initialSigma=2.0, endSigma=0.5)
            ocfg = oracle.IcpConfig(max_iterations=iters, initial_sigma=2.0, end_sigma=0.5)
            cls, oalgo = api.IcpRegistration, oracle.IcpAlgorithm(ocfg)
        ggeneral = api.SimpleRegistrator(ctx, cls, cfg, dm, dt).runDecimated(dref, dtgt, ggeneral)
        # oracle: decimateState + run
        om2 = m.new_reference(*dref)
        if ogeneral is None:
            ost = oracle.initial_state(om2, dtgt[0], dtgt[1], global_transformation=oracle.RIGID_TRANSFORMS)
        else:
            ost = dataclasses.replace(ogeneral, model=om2, target=np.ascontiguousarray(dtgt[0]), target_tri=dtgt[1], iteration=0,
                                      status=oracle.STATUS_NONE,
                                      fit=oracle.model_instance_shape_pose_scale(om2, ogeneral.params))
        ofinal = oracle.run(oalgo, ost)
        ogeneral = dataclasses.replace(ofinal, fit=oracle.model_instance_shape_pose_scale(m, ofinal.params))
        _compare(ggeneral, ogeneral, diag, tol=1e-6)
        assert ggeneral.fit.shape == (600, 3)
    dm.close(); dt.close()


def test_new_reference_argument_errors(ctx, oracle):
    from gingr_b200 import api
    m, target, tt = _problem(oracle, 100, 100, 8, seed=0)
    dm = api.Model(ctx, m.ref, m.mean, m.basis, m.variance, m.tri)
    with pytest.raises(api.GingrError):
        dm.newReference(m.ref[:10], np.array([[0, 1, 99]], dtype=np.int32))       # triangle index out of range
    one = dm.newReference(m.ref[:1])                                               # a single point, no triangles
    assert one.M == 1
    one.close(); dm.close()
