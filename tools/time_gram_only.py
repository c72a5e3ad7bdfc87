"""K3 Gram at the SURVEY 8(d) 'Gram-only' size: 3n = 600 000 rows (n = M = 200 000 observed points), r = 2000 -- a
9.6 GB basis resident on one B200.  One CPD update() per step with a small target so that the iteration is the
posterior; reports the DMMA Gram kernel (profile slot 2) against the measured FP64 tensor peak and the HBM passes.
usage: python tools/time_gram_only.py [M] [r] [iters]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gingr_b200 import api, synthetic

M = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
r = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
N = 2000
t0 = time.time()
ref = synthetic.fibonacci_sphere(M)
rng = np.random.default_rng(1)
basis = rng.standard_normal(size=(r, 3 * M)).T          # column-major [3M, r] view: no transposed host copy
basis *= 1.0 / np.sqrt(3.0 * M)
var = 100.0 * 0.995 ** np.arange(r)
target = synthetic.make_target(synthetic.fibonacci_sphere(N), 0)
t_gen = time.time() - t0
ctx = api.Context(0)
t0 = time.time()
model = api.Model(ctx, ref, np.zeros(3 * M), basis, var)
t_up = time.time() - t0
tgt = api.Target(ctx, target)
reg = api.CpdRegistration(ctx, model, tgt, api.CpdConfiguration(maxIterations=10 ** 6, w=0.1))
reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
reg.updateChain(2)
ctx.synchronize()
reg.setProfiling(True)
reg.updateChain(iters)
ms, it = reg.getProfile()
st = reg.downloadState()
assert np.all(np.isfinite(st.fit)) and st.status != api.STATUS_MODEL_FLEXIBILITY_ERROR
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "FP64_PEAKS.json")))
dmma = max(peaks["cublas_dgemm_tflops_sustained"], peaks["microbench"]["dmma_tflops_sustained"])
t_g = ms[2] / it
flop = 3.0 * M * r * (r + 1)
out = {"M": M, "rows": 3 * M, "rank": r, "iters": it, "basis_GB": 8.0 * 3 * M * r / 1e9,
       "gram_ms": t_g, "gram_TFLOPs": flop / (t_g * 1e-3) / 1e12, "dmma_peak_TFLOPs": dmma,
       "frac_of_dmma_peak": flop / (t_g * 1e-3) / 1e12 / dmma,
       "gram_HBM_GBps_algorithmic": 8.0 * 3 * M * r / (t_g * 1e-3) / 1e9,
       "estep_ms": (ms[0] + ms[1]) / it, "rest_ms_4_basis_passes_and_small_kernels": (ms[4] - ms[0] - ms[1] - ms[2] - ms[3]) / it,
       "four_basis_passes_at_hbm_peak_ms": 4 * 8.0 * 3 * M * r / 6536e9 * 1e3,
       "cholesky_ms": ms[3] / it, "iteration_ms": ms[4] / it, "host_generate_s": t_gen, "upload_s": t_up}
print(json.dumps(out))
