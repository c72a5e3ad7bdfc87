"""One Metropolis-Hastings chain of the C1 size, a few steps without CUDA graphs: the launch list of one MH step
(run under ncu --metrics gpu__time_duration.sum).   usage: GINGR_CUDA_GRAPH=0 python tools/mh_step_once.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gingr_b200 import api, synthetic

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
M, N, r = 100, 100, 50
ctx = api.Context(0)
ref, tri = synthetic.sphere_mesh(M)
mean, basis, var = synthetic.make_gpmm(ref, r, 1)
tv, tt = synthetic.sphere_mesh(N)
model = api.Model(ctx, ref, mean, basis, var, tri)
tgt = api.Target(ctx, synthetic.make_target(tv, 0), tt)
c = api.IcpRegistration(ctx, model, tgt, api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0))
c.configureProbabilistic(api.ProbabilisticSettings(uncertainty=1.0, randomMixture=0.5))
c.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
c.mcmcChain(steps, seed=1) if hasattr(c, "mcmcChain") else api.mcmc_batch([c], steps, seed=1)
ctx.synchronize()
print("launches", ctx.launch_count)
