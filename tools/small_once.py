"""A few iterations of a C1-size registration without the CUDA graph -- the command profiled for the launch list of the
launch-bound small configurations.   usage: python tools/small_once.py [cpd|icp|mcmc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GINGR_CUDA_GRAPH"] = "0"
from gingr_b200 import api, synthetic
algo = sys.argv[1] if len(sys.argv) > 1 else "icp"
M, N, r = 100, 100, 50
ref, tri = synthetic.sphere_mesh(M)
mean, basis, var = synthetic.make_gpmm(ref, r, 1)
tv, tt = synthetic.sphere_mesh(N)
target = synthetic.make_target(tv, 0)
ctx = api.Context(0)
model = api.Model(ctx, ref, mean, basis, var, tri)
tgt = api.Target(ctx, target, tt)
if algo == "cpd":
    reg = api.CpdRegistration(ctx, model, tgt, api.CpdConfiguration(maxIterations=10 ** 6, w=0.1))
else:
    reg = api.IcpRegistration(ctx, model, tgt, api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0))
reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
if algo == "mcmc":
    reg.configureProbabilistic(api.ProbabilisticSettings(uncertainty=1.0, randomMixture=0.5))
    reg.mcmcChain(3, 1)
else:
    reg.updateChain(3)
ctx.synchronize()
