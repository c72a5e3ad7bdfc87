"""Parity against the REAL reference, whenever a maintainer has produced golden files with it.

scala/tools/MakeGolden.scala (run with scala-cli inside a GiNGR checkout; see its header) executes the reference's own
CpdRegistration / IcpRegistration update() + fit refresh on the femur example and writes
tests/golden/reference_<name>.json with the model arrays, the target and the state after every iteration.  This module
replays those inputs through the CPU oracle (CPU run) and through libgingr_cuda (`-m gpu`) and compares every iteration
at the north-star tolerances: coefficients 1e-6 relative to max |alpha|, vertices 1e-6 of the bounding-box diagonal,
sigma2 1e-6 relative, pose 1e-6.  Without such files the comparisons are skipped (parity stays "unpinned") and
only the harness self-test runs: a file in the same schema written from the oracle must read back and compare clean, so
that the first real file meets a checked reader, not an untested one."""
import glob
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "reference_*.json")))
TOL = 1e-6
HOW = "no tests/golden/reference_*.json -- produce them with scala/tools/MakeGolden.scala (scala-cli, inside a GiNGR checkout)"
GT = {"SimilarityTransforms": 0, "RigidTransforms": 1, "NoTransforms": 2}


def load_reference_file(path):
    d = json.load(open(path))
    assert d["schema"].startswith("gingr-b200 reference golden")
    m = d["model"]
    model = dict(ref=np.array(m["reference"]["points"], dtype=np.float64), tri=np.array(m["reference"]["triangles"], dtype=np.int32),
                 mean=np.array(m["mean"], dtype=np.float64), basis=np.array(m["basis"], dtype=np.float64),
                 variance=np.array(m["variance"], dtype=np.float64))
    assert model["basis"].shape == (3 * model["ref"].shape[0], model["variance"].shape[0])
    target = (np.array(d["target"]["points"], dtype=np.float64), np.array(d["target"]["triangles"], dtype=np.int32))
    return d, model, target


def compare_state(got_alpha, got_fit, got_sigma2, got_t, got_euler, got_scale, want, diag, tag):
    a = np.array(want["alpha"])
    assert np.max(np.abs(got_alpha - a)) <= TOL * max(np.max(np.abs(a)), 1e-12), f"{tag}: coefficients"
    assert np.max(np.abs(got_fit - np.array(want["fit"]))) <= TOL * diag, f"{tag}: fitted vertices"
    assert abs(got_sigma2 - want["sigma2"]) <= TOL * abs(want["sigma2"]), f"{tag}: sigma2"
    assert np.max(np.abs(np.asarray(got_t) - np.array(want["translation"]))) <= TOL * diag, f"{tag}: translation"
    assert np.max(np.abs(np.asarray(got_euler) - np.array(want["euler"]))) <= TOL, f"{tag}: Euler angles"
    assert abs(got_scale - want["scale"]) <= TOL, f"{tag}: scale"


def replay_oracle(oracle, d, model, target):
    om = oracle.Gpmm(model["ref"], model["mean"], model["basis"], model["variance"], model["tri"])
    cfg = d["config"]
    if d["algorithm"] == "CPD":
        algo = oracle.CpdAlgorithm(oracle.CpdConfig(w=cfg["w"], lam=cfg["lambda"], initial_sigma=cfg["initialSigma"],
                                                    max_iterations=cfg["maxIterations"]))
    else:
        algo = oracle.IcpAlgorithm(oracle.IcpConfig(initial_sigma=cfg["initialSigma"], end_sigma=cfg["endSigma"],
                                                    max_iterations=cfg["maxIterations"], reverse=cfg.get("reverse", False)))
    st = algo.initialize(oracle.initial_state(om, target[0], target[1], global_transformation=GT[d["globalTransformation"]]))
    diag = float(np.linalg.norm(model["ref"].max(0) - model["ref"].min(0)))
    assert abs(st.sigma2 - d["states"][0]["sigma2"]) <= TOL * d["states"][0]["sigma2"]
    for k, want in enumerate(d["states"][1:], 1):
        st = oracle.propose(algo, st)
        compare_state(st.params.shape, st.fit, st.sigma2, st.params.translation, st.params.euler, st.params.scale, want, diag,
                      f"oracle, iteration {k}")


def replay_device(ctx, d, model, target):
    from gingr_b200 import api
    dm = api.Model(ctx, model["ref"], model["mean"], model["basis"], model["variance"], model["tri"])
    dt = api.Target(ctx, target[0], target[1])
    cfg = d["config"]
    if d["algorithm"] == "CPD":
        reg = api.CpdRegistration(ctx, dm, dt, api.CpdConfiguration(maxIterations=cfg["maxIterations"], w=cfg["w"], initialSigma=cfg["initialSigma"]))
    else:
        reg = api.IcpRegistration(ctx, dm, dt, api.IcpConfiguration(maxIterations=cfg["maxIterations"], initialSigma=cfg["initialSigma"],
                                                                    endSigma=cfg["endSigma"]))
    st = reg.initializeState(globalTransformation=GT[d["globalTransformation"]])
    diag = float(np.linalg.norm(model["ref"].max(0) - model["ref"].min(0)))
    for k, want in enumerate(d["states"][1:], 1):
        st = reg.propose(st)
        p = st.modelParameters
        compare_state(p.shape, st.fit, st.sigma2, p.translation, p.euler, p.scale, want, diag, f"device, iteration {k}")
    reg.close(); dm.close(); dt.close()


def write_in_reference_schema(oracle, path, algorithm):
    """A file in MakeGolden's schema, with ORACLE outputs: exercises reader and comparison only."""
    from gingr_b200 import synthetic
    ref, tri = synthetic.sphere_mesh(90)
    mean, basis, var = synthetic.make_gpmm(ref, 12, 1)
    tv, tt = synthetic.sphere_mesh(110)
    tv = synthetic.make_target(tv, 0)
    om = oracle.Gpmm(ref, mean, basis, var, tri)
    if algorithm == "CPD":
        cfg = {"initialSigma": 1.0, "w": 0.0, "lambda": 1.0, "maxIterations": 100}
        algo = oracle.CpdAlgorithm(oracle.CpdConfig(initial_sigma=1.0))
        gt = "RigidTransforms"
    else:
        cfg = {"initialSigma": 1.0, "endSigma": 1.0, "maxIterations": 100, "reverse": False, "method": "TriangularClosestPoint"}
        algo = oracle.IcpAlgorithm(oracle.IcpConfig(initial_sigma=1.0, end_sigma=1.0))
        gt = "NoTransforms"
    st = algo.initialize(oracle.initial_state(om, tv, tt, global_transformation=GT[gt]))

    def state(s):
        return {"iteration": s.iteration, "status": "None", "sigma2": s.sigma2, "scale": s.params.scale,
                "translation": list(map(float, s.params.translation)), "euler": list(map(float, s.params.euler)),
                "alpha": list(map(float, s.params.shape)), "fit": s.fit.tolist()}
    states = [state(st)]
    for _ in range(3):
        st = oracle.propose(algo, st)
        states.append(state(st))
    json.dump({"schema": "gingr-b200 reference golden v1", "generator": "tests/test_reference_golden.py (ORACLE outputs: harness self-test)",
               "algorithm": algorithm, "config": cfg, "globalTransformation": gt,
               "model": {"reference": {"points": ref.tolist(), "triangles": tri.tolist()}, "mean": mean.tolist(), "variance": var.tolist(),
                         "basis": basis.tolist()},
               "target": {"points": tv.tolist(), "triangles": tt.tolist()}, "states": states}, open(path, "w"))


@pytest.mark.parametrize("algorithm", ["CPD", "ICP"])
def test_harness_reads_and_compares_the_schema(oracle, tmp_path, algorithm):
    path = str(tmp_path / f"reference_selftest_{algorithm}.json")
    write_in_reference_schema(oracle, path, algorithm)
    d, model, target = load_reference_file(path)
    replay_oracle(oracle, d, model, target)
    d["states"][2]["alpha"][0] += 1e-3                       # and a deviation is caught
    with pytest.raises(AssertionError):
        replay_oracle(oracle, d, model, target)


@pytest.mark.skipif(not FILES, reason=HOW)
@pytest.mark.parametrize("path", FILES or ["-"])
def test_oracle_matches_the_reference(oracle, path):
    d, model, target = load_reference_file(path)
    replay_oracle(oracle, d, model, target)


@pytest.mark.gpu
@pytest.mark.skipif(not FILES, reason=HOW)
@pytest.mark.parametrize("path", FILES or ["-"])
def test_device_matches_the_reference(ctx, path):
    d, model, target = load_reference_file(path)
    replay_device(ctx, d, model, target)


@pytest.mark.gpu
@pytest.mark.parametrize("algorithm", ["CPD", "ICP"])
def test_device_harness_on_the_self_test_file(ctx, oracle, tmp_path, algorithm):
    path = str(tmp_path / f"reference_selftest_{algorithm}.json")
    write_in_reference_schema(oracle, path, algorithm)
    d, model, target = load_reference_file(path)
    replay_device(ctx, d, model, target)
