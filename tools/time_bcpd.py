"""BCPD E-step (gingr_bcpd_estep: BCPD.computeP + reductions, other/algorithms/cpd/BCPD.scala:167-209) at the C4 size through
the host-in / host-out entry point (includes the H2D / D2H of the call); also the CPD E-step entry point beside it.
usage: python tools/time_bcpd.py [M] [N]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gingr_b200 import api, synthetic
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
ctx = api.Context(0)
y = synthetic.fibonacci_sphere(M)
x = synthetic.make_target(synthetic.fibonacci_sphere(N), 0)
tgt = api.Target(ctx, x)
rng = np.random.default_rng(0)
sigma_mm = rng.uniform(0.0, 0.01, M)
alpha = np.full(M, 1.0 / M)
def best(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)
t_b = best(lambda: api.bcpd_estep(ctx, tgt, y, sigma_mm, alpha, 25.0, 1.0, 0.1))
t_c = best(lambda: api.cpd_estep(ctx, tgt, y, 25.0, 0.1))
nu, nup, nhat, xhat = api.bcpd_estep(ctx, tgt, y, sigma_mm, alpha, 25.0, 1.0, 0.1)
assert np.all(np.isfinite(xhat)) and abs(nhat - nup.sum()) < 1e-9 * nhat
print(json.dumps({"M": M, "N": N, "bcpd_estep_call_ms": t_b * 1e3, "bcpd_TFLOPs_71": 71.0 * M * N / t_b / 1e12,
                  "cpd_estep_call_ms": t_c * 1e3, "cpd_TFLOPs_71": 71.0 * M * N / t_c / 1e12, "n_hat": float(nhat)}))
