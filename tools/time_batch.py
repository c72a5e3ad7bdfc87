"""BASELINE config 5 shape: n independent probabilistic ICP chains of the C1 size on one GPU (gingr_update_batch).
usage: python tools/time_batch.py [n_chains] [iters]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gingr_b200 import api, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
M, N, r = 100, 100, 50
ctx = api.Context(0)
ref, tri = synthetic.sphere_mesh(M)
mean, basis, var = synthetic.make_gpmm(ref, r, 1)
tv, tt = synthetic.sphere_mesh(N)
target = synthetic.make_target(tv, 0)
model = api.Model(ctx, ref, mean, basis, var, tri)
tgt = api.Target(ctx, target, tt)
cfg = api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0)
chains = [api.IcpRegistration(ctx, model, tgt, cfg) for _ in range(n)]
for c in chains:
    c.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
api.update_batch(chains, 3, probabilistic=True, seed=0)      # warm-up: captures the graphs
ctx.synchronize()
t0 = time.perf_counter()
api.update_batch(chains, iters, probabilistic=True, seed=0)
ctx.synchronize()
dt = time.perf_counter() - t0
st = chains[-1].downloadState()
assert np.all(np.isfinite(st.fit)) and st.iteration == iters + 3
print(json.dumps({"workload": "probabilistic ICP chains, M=N=100 r=50 (C5 shape)", "chains": n, "iters": iters,
                  "chain_iterations_per_s": n * iters / dt, "ms_per_batch_step": dt / iters * 1e3}))
