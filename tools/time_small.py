"""Device-chain iterations/s of the launch-bound small configurations (C1-C3 sizes), with and without the captured
CUDA graph (GINGR_CUDA_GRAPH=0/1).  usage: python tools/time_small.py [cpd|icp]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gingr_b200 import api, synthetic

algo = sys.argv[1] if len(sys.argv) > 1 else "cpd"
ctx = api.Context(0)
out = {}
for name, (M, N, r) in {"c1": (100, 100, 50), "c3b": (500, 500, 100), "c3c": (1000, 1000, 100)}.items():
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, 0)
    model = api.Model(ctx, ref, mean, basis, var, tri)
    tgt = api.Target(ctx, target, tt)
    if algo == "cpd":
        reg = api.CpdRegistration(ctx, model, tgt, api.CpdConfiguration(maxIterations=10 ** 6, w=0.1))
    else:
        reg = api.IcpRegistration(ctx, model, tgt, api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0))
    reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    reg.updateChain(20)
    ctx.synchronize()
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    iters = 300
    reg.updateChain(iters)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    st = reg.downloadState()
    assert np.all(np.isfinite(st.fit))
    out[name] = {"M": M, "N": N, "r": r, "it_per_s": iters / dt, "us_per_iter": dt / iters * 1e6,
                 "launches_per_iter": (ctx.launch_count - l0) / iters}
    reg.close()
print(json.dumps({"algo": algo, "graph": os.environ.get("GINGR_CUDA_GRAPH", "1"), "results": out}))
