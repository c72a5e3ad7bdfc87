"""CPU known-answer tests of the probabilistic restatement in oracle/ (no GPU): the SVD-free form of the informed
transition density that libgingr_cuda evaluates (gingr_b200/csrc/mcmc.cuh) equals the literal
GeneratorWrapperStochastic.logTransitionProbability; mixture / evaluator identities; the Philox stream."""
import dataclasses

import numpy as np
import pytest


def _problem(oracle, M=120, N=150, r=12, algo="icp", seed=0):
    from gingr_b200 import synthetic
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, seed)
    om = oracle.Gpmm(ref, mean, basis, var, tri)
    if algo == "icp":
        a = oracle.IcpAlgorithm(oracle.IcpConfig(initial_sigma=2.0, end_sigma=0.5, max_iterations=20))
    else:
        a = oracle.CpdAlgorithm(oracle.CpdConfig(w=0.1), literal=False)
    st = a.initialize(oracle.initial_state(om, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
    return a, st


def _spd_form(oracle, algo, frm, to_mesh):
    """-1/2 u^T Mx u - r/2 log(2 pi) with (S + eps Mx) u = D Phi^T (R^T (toMesh - t) - ref - mean) - S c."""
    posed, c, Minv = oracle.compute_posterior_coefficients(algo, frm)
    Mx = posed._last_Mx
    m = frm.model
    D = np.sqrt(m.variance)
    S = (m.basis * D).T @ (m.basis * D)
    R, t = frm.params.rotation_matrix(), frm.params.translation
    resid = ((to_mesh - t) @ R) - m.ref - m.mean.reshape(-1, 3)
    b = D * (m.basis.T @ resid.reshape(-1)) - S @ c
    u = np.linalg.solve(S + 1e-5 * Mx, b)
    return -0.5 * u @ Mx @ u - 0.5 * len(c) * np.log(2 * np.pi)


@pytest.mark.parametrize("algo", ["icp", "cpd"])
@pytest.mark.parametrize("step_length", [1.0, 0.5])
def test_svd_free_transition_density_equals_literal(oracle, algo, step_length):
    a, st = _problem(oracle, algo=algo)
    st = dataclasses.replace(st, step_length=step_length)
    for _ in range(2):
        st = oracle.propose(a, st)
    nxt = oracle.propose(a, st, True, seed=3)
    lit = oracle.log_transition_informed(a, st, nxt)
    if step_length != 1.0:
        to_mesh = st.model.instance(st.params.shape + (nxt.params.shape - st.params.shape) / step_length)
    else:
        to_mesh = st.fit
    mine = _spd_form(oracle, a, st, to_mesh)
    assert np.isfinite(lit)
    assert abs(mine - lit) <= 1e-7 * abs(lit)


def test_mixture_transition_of_random_moves(oracle):
    a, st = _problem(oracle)
    s = oracle.McmcSettings(random_mixture=0.7)
    for leaf in range(1, 10):
        prop = oracle.random_proposal(s, st, leaf, seed=5, step=leaf)
        assert prop.iteration == st.iteration + 1
        fw = oracle.mixture_log_transition(s, st, prop, -np.inf)
        bw = oracle.mixture_log_transition(s, prop, st, -np.inf)
        assert np.isfinite(fw) and abs(fw - bw) < 1e-12 * max(1.0, abs(fw))     # symmetric Gaussian proposals
        # exactly the leaves of that parameter group contribute
        w = oracle.leaf_weights(0.7)
        p, q = st.params, prop.params
        if leaf <= 3:
            slot = {1: 2, 2: 1, 3: 0}[leaf]
            d = q.euler[slot] - p.euler[slot]
            want = np.log(sum(w[1 + k] * np.exp(oracle.gaussian_logpdf(d if k == leaf - 1 else 0.0, 0.0, s.rot_sdev[k])) for k in range(3)))
        elif leaf <= 6:
            d = q.translation[leaf - 4] - p.translation[leaf - 4]
            want = np.log(sum(w[4 + k] * np.exp(oracle.gaussian_logpdf(d if k == leaf - 4 else 0.0, 0.0, s.trans_sdev[k])) for k in range(3)))
        else:
            ss = np.sum((q.shape - p.shape) ** 2)
            r = len(p.shape)
            want = np.log(sum(w[7 + k] * np.exp(-ss / (2 * s.shape_steps[k] ** 2) - r * np.log(s.shape_steps[k] * np.sqrt(2 * np.pi))) for k in range(3)))
        assert abs(fw - want) < 1e-12 * max(1.0, abs(want))
    # an informed move changes everything: every random leaf is -inf, the mixture is the weighted informed density
    nxt = oracle.propose(a, st, True, seed=1)
    assert abs(oracle.mixture_log_transition(s, st, nxt, -3.0) - (np.log(0.3) - 3.0)) < 1e-12
    assert oracle.mixture_log_transition(s, st, nxt, -np.inf) == -np.inf


def test_evaluators(oracle):
    a, st = _problem(oracle)
    alpha = np.linspace(-1, 1, st.model.rank)
    assert abs(oracle.model_evaluator(alpha) - (-0.5 * alpha @ alpha - 0.5 * len(alpha) * np.log(2 * np.pi))) < 1e-14
    s = oracle.McmcSettings(uncertainty=2.0, mode=oracle.EVAL_SYMMETRIC)
    sym = oracle.distance_evaluator(s, st)
    m2t = oracle.distance_evaluator(dataclasses.replace(s, mode=oracle.EVAL_MODEL_TO_TARGET), st)
    t2m = oracle.distance_evaluator(dataclasses.replace(s, mode=oracle.EVAL_TARGET_TO_MODEL), st)
    assert abs(sym - 0.5 * (m2t + t2m)) < 1e-12 * abs(sym)
    # a point on the surface contributes the peak of the density
    on = dataclasses.replace(st, fit=st.target.copy(), model=dataclasses.replace(st.model, tri=st.target_tri))
    peak = oracle.distance_evaluator(dataclasses.replace(s, mode=oracle.EVAL_MODEL_TO_TARGET), on)
    assert abs(peak - len(st.target) * oracle.gaussian_logpdf(0.0, 0.0, 2.0)) < 1e-6
    ids = np.arange(0, st.model.M, 3)
    sub = oracle.distance_evaluator(dataclasses.replace(s, mode=oracle.EVAL_MODEL_TO_TARGET, model_ids=ids), st)
    assert sub > m2t    # fewer (negative) terms


def test_mcmc_streams_are_counter_based(oracle):
    u = [oracle.mcmc_uniforms(7, k) for k in range(200)]
    uc = np.array([x[0] for x in u])
    ua = np.array([x[1] for x in u])
    assert np.all((uc > 0) & (uc < 1)) and np.all((ua > 0) & (ua < 1))
    assert 0.35 < uc.mean() < 0.65 and 0.35 < ua.mean() < 0.65
    assert oracle.mcmc_uniforms(7, 3) == u[3] and oracle.mcmc_uniforms(8, 3) != u[3]
    z = oracle.mcmc_normals(4001, 11, 2)
    assert abs(z.mean()) < 0.06 and abs(z.std() - 1.0) < 0.05
    assert np.array_equal(oracle.mcmc_normals(10, 11, 2), z[:10])
    # the three purposes (posterior sample, uniforms, perturbations) are distinct streams
    assert not np.allclose(oracle.standard_normals(10, 11, 2), z[:10])


def test_mcmc_steps_run_and_mix(oracle):
    a, st = _problem(oracle, M=80, N=90, r=8)
    s = oracle.McmcSettings(uncertainty=1.5, random_mixture=0.5)
    lp = oracle.log_value(s, st)
    leaves, accepts = set(), 0
    for k in range(12):
        st, lp, info = oracle.mcmc_step(a, s, st, lp, k, seed=4)
        leaves.add(info["leaf"])
        accepts += int(info["accept"])
        assert np.all(np.isfinite(st.fit))
    assert len(leaves) >= 3 and accepts >= 1
