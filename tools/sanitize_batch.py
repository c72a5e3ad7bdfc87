"""Small batched runs for compute-sanitizer (memcheck): 6 Metropolis-Hastings chains x 2 steps and 6 probabilistic ICP
registrations x 2 iterations through the batched kernel sequence (batch.cuh), a rank-50 problem (chol_small_kernel, the
small-rank Gram layout, the pruned surface scan).   usage: compute-sanitizer --tool memcheck python tools/sanitize_batch.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gingr_b200 import api, synthetic

M, N, r = 100, 100, 50
ctx = api.Context(0)
ref, tri = synthetic.sphere_mesh(M)
mean, basis, var = synthetic.make_gpmm(ref, r, 1)
tv, tt = synthetic.sphere_mesh(N)
model = api.Model(ctx, ref, mean, basis, var, tri)
tgt = api.Target(ctx, synthetic.make_target(tv, 0), tt)
cfg = api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0)
chains = []
for _ in range(6):
    c = api.IcpRegistration(ctx, model, tgt, cfg)
    c.configureProbabilistic(api.ProbabilisticSettings(uncertainty=1.0, randomMixture=0.5))
    c.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    chains.append(c)
l0 = ctx.launch_count
api.mcmc_batch(chains, 2, seed=5)
ctx.synchronize()
print("MH batch launches per step:", (ctx.launch_count - l0) / 2, "finite:", bool(np.all(np.isfinite(chains[0].downloadState().fit))))
regs = [api.IcpRegistration(ctx, model, tgt, cfg) for _ in range(6)]
for g in regs:
    g.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
l0 = ctx.launch_count
api.update_batch(regs, 2, probabilistic=True, seed=9)
ctx.synchronize()
print("update batch launches per iteration:", (ctx.launch_count - l0) / 2, "finite:", bool(np.all(np.isfinite(regs[-1].downloadState().fit))))
