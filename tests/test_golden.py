"""Real-data regression vectors (tests/golden/femur_golden.npz, made by tests/golden/make_golden.py from the
reference's femur example meshes).  CPU: the oracle reproduces them; GPU: the CUDA path reproduces them through the
C ABI.  They are oracle outputs, not outputs of the Scala reference (see the generator's header and DESIGN.md)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "femur_golden.npz"))


def _gpmm(ref):
    import make_golden
    return make_golden.femur_gpmm(ref)


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def test_fixture_is_the_reference_femur(gold):
    assert gold["ref_v"].shape == (1622, 3) and gold["ref_t"].shape == (3240, 3)
    assert gold["tgt_v"].shape == (1622, 3) and gold["tgt_t"].shape == (3240, 3)
    assert gold["lm_ref"].shape == (6, 3) and gold["lm_tgt"].shape == (6, 3)
    assert gold["ref_t"].max() == 1621 and gold["ref_t"].min() == 0


def test_oracle_reproduces_golden_estep_and_icp_correspondence(oracle, gold):
    sub_r, sub_t = gold["ref_v"][::16], gold["tgt_v"][::16]
    P1, Pt1, PX = oracle.P_reductions(oracle.cpd_P(sub_r, sub_t, 1.0, 0.0), sub_t)
    assert _rel(P1, gold["c2_P1"]) < 1e-12 and _rel(PX, gold["c2_PX"]) < 1e-12
    # DemoCPD's sigma2 = 1 on the femur: the dynamic range FP32 cannot hold
    assert gold["c2_P1"].min() < 1e-20
    s1, st1, sx = oracle.cpd_estep(sub_r, sub_t, 1.0, 0.0, fast=True)               # streaming form == literal form
    assert np.max(np.abs(s1 - gold["c2_P1"]) / gold["c2_P1"]) < 1e-9
    cp, w, md, idx = oracle.closest_point_correspondence(oracle.METHOD_TRIANGULAR, gold["ref_v"], gold["ref_t"],
                                                         gold["tgt_v"], gold["tgt_t"])
    assert np.array_equal(idx, gold["c1_idx"]) and np.array_equal(w.astype(np.uint8), gold["c1_w"])
    assert np.array_equal(cp, gold["c1_cp"])


def test_oracle_reproduces_golden_cpd_run(oracle, gold):
    sub_r, sub_t = gold["ref_v"][::16], gold["tgt_v"][::16]
    mean, basis, var = _gpmm(sub_r)
    m = oracle.Gpmm(sub_r, mean, basis, var, None)
    lm_pid, _ = oracle.nearest_vertex(gold["lm_ref"], sub_r)
    lms = oracle.Landmarks(lm_pid.astype(np.int32), gold["lm_tgt"], np.tile(np.eye(3), (6, 1, 1)))
    algo = oracle.CpdAlgorithm(oracle.CpdConfig(max_iterations=100), literal=False)   # streaming mode vs literal golden
    st = algo.initialize(oracle.initial_state(m, sub_t, None, global_transformation=oracle.RIGID_TRANSFORMS, landmarks=lms))
    assert abs(st.sigma2 - float(gold["c2_sigma2_0"])) < 1e-12 * st.sigma2
    diag = np.linalg.norm(sub_r.max(0) - sub_r.min(0))
    for it in range(1, 6):
        st = oracle.propose(algo, st)
        if it in (1, 5):
            assert _rel(st.params.shape, gold[f"c2_alpha_{it}"]) < 1e-7
            assert np.max(np.abs(st.fit - gold[f"c2_fit_{it}"])) < 1e-7 * diag
            assert abs(st.sigma2 - float(gold[f"c2_sigma2_{it}"])) < 1e-7 * st.sigma2


@pytest.mark.gpu
def test_cuda_reproduces_golden_estep(ctx, gold):
    from gingr_b200 import api
    sub_r, sub_t = gold["ref_v"][::16], gold["tgt_v"][::16]
    tgt = api.Target(ctx, sub_t)
    P1, Pt1, PX = api.cpd_estep(ctx, tgt, sub_r, 1.0, 0.0)
    assert np.max(np.abs(P1 - gold["c2_P1"]) / gold["c2_P1"]) < 1e-9        # element-wise relative, values down to 1e-100
    assert np.max(np.abs(Pt1 - gold["c2_Pt1"]) / gold["c2_Pt1"]) < 1e-9
    assert np.max(np.abs(PX - gold["c2_PX"]) / np.maximum(np.abs(gold["c2_PX"]), 1e-300)) < 1e-6
    tgt.close()


@pytest.mark.gpu
def test_cuda_reproduces_golden_cpd_run_with_landmarks(ctx, gold):
    """DemoCPD-like run (CPD, RigidTransforms, the femur's 6 landmarks, sigma2_0 from computeInitialSigma2)."""
    from gingr_b200 import api
    sub_r, sub_t = gold["ref_v"][::16], gold["tgt_v"][::16]
    mean, basis, var = _gpmm(sub_r)
    dm = api.Model(ctx, sub_r, mean, basis, var)
    dt = api.Target(ctx, sub_t)
    reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(maxIterations=100))
    d2 = ((gold["lm_ref"][:, None, :] - sub_r[None, :, :]) ** 2).sum(-1)
    reg.setLandmarks(np.argmin(d2, axis=1).astype(np.int32), gold["lm_tgt"])
    st = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    assert abs(st.sigma2 - float(gold["c2_sigma2_0"])) < 1e-12 * st.sigma2
    diag = np.linalg.norm(sub_r.max(0) - sub_r.min(0))
    for it in range(1, 21):
        st = reg.propose(st)
        if it in (1, 5, 20):
            assert _rel(st.modelParameters.shape, gold[f"c2_alpha_{it}"]) < 1e-6
            assert np.max(np.abs(st.fit - gold[f"c2_fit_{it}"])) < 1e-6 * diag
            assert abs(st.sigma2 - float(gold[f"c2_sigma2_{it}"])) < 1e-6 * st.sigma2
            pose = np.concatenate([[st.modelParameters.scale], st.modelParameters.translation, st.modelParameters.euler])
            assert np.max(np.abs(pose - gold[f"c2_pose_{it}"])) < 1e-6 * diag
    reg.close()


@pytest.mark.gpu
def test_cuda_reproduces_golden_icp_on_the_full_femur(ctx, gold):
    """DemoICP-like: TriangularClosestPoint on the 1622-vertex femur meshes; indices, weights and points exact."""
    from gingr_b200 import api
    rv, rt, tv, tt = gold["ref_v"], gold["ref_t"], gold["tgt_v"], gold["tgt_t"]
    dt = api.Target(ctx, tv, tt)
    idx, cp, w, md = api.icp_closest(ctx, dt, rv, rt, api.TRIANGULAR_CLOSEST_POINT)
    assert np.array_equal(idx, gold["c1_idx"]) and np.array_equal(w, gold["c1_w"]) and np.array_equal(cp, gold["c1_cp"])
    assert abs(md - float(gold["c1_mean_dist"])) < 1e-12 * md
    mean, basis, var = _gpmm(rv)
    dm = api.Model(ctx, rv, mean, basis, var, rt)
    reg = api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration(maxIterations=100, initialSigma=1.0, endSigma=1.0))
    st = reg.initializeState(globalTransformation=api.NO_TRANSFORMS)
    diag = np.linalg.norm(rv.max(0) - rv.min(0))
    for it in range(1, 4):
        st = reg.propose(st)
        assert _rel(st.modelParameters.shape, gold[f"c1_alpha_{it}"]) < 1e-6
        assert np.max(np.abs(st.fit - gold[f"c1_fit_{it}"])) < 1e-6 * diag
    reg.close()
