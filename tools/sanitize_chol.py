"""The round-2 kernels under compute-sanitizer: the data-flow Cholesky (single tile, ragged last tile, several block
columns, many extra rows), the back substitution from the block inverses, the posterior covariance (rows of Q riding
through the factorisation), a CPD iteration on top.
    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_chol.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_chol.py small"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GINGR_CUDA_GRAPH"] = "0"
import numpy as np
from gingr_b200 import api, synthetic
what = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = api.Context(0)
rng = np.random.default_rng(0)
for n, nrhs in ((40, 1), (64, 1), (100, 3), (200, 70)) if what == "small" else ((40, 1), (64, 1), (100, 3), (200, 200), (330, 2)):
    Q = rng.normal(size=(n, n + 8))
    A = Q @ Q.T / (n + 8) + np.eye(n) * 1e-2
    B = rng.normal(size=(nrhs, n))
    o = api.spd_solve(ctx, A, B)
    assert np.max(np.abs(o["L"] - np.linalg.cholesky(A))) < 1e-11
    assert np.max(np.abs(A @ o["x"] - B[0])) < 1e-9
print("spd_solve ok")
if what == "all":
    M, r = 150, 70
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    model = api.Model(ctx, ref, mean, basis, var, tri)
    pids = np.arange(0, M, 2, dtype=np.int32)
    cov = api.posterior_covariance(ctx, model, np.eye(3), np.zeros(3), pids, ref[pids] + 0.1, np.full(len(pids), 0.5))
    assert np.all(np.isfinite(cov))
    tv, tt = synthetic.sphere_mesh(180)
    tgt = api.Target(ctx, synthetic.make_target(tv, 0), tt)
    reg = api.CpdRegistration(ctx, model, tgt, api.CpdConfiguration(w=0.1))
    reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    reg.updateChain(2)
    assert np.all(np.isfinite(reg.downloadState().fit))
    print("posterior covariance + CPD iteration ok")
