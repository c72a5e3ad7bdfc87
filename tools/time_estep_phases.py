"""E-step sweep times (CUDA events of the instrumented iteration) of a CPD update at a given shape, e.g. the per-rank shape of the
8-GPU C4 run (M = 20000, N = 25000).  usage: [GINGR_ESTEP_SPLITS_A=s] [GINGR_ESTEP_SPLITS_B=s] python tools/time_estep_phases.py M N"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gingr_b200 import api, synthetic

M, N = int(sys.argv[1]), int(sys.argv[2])
ctx = api.Context(0)
ref, tri = synthetic.sphere_mesh(M)
mean, basis, var = synthetic.make_gpmm(ref, 64, 1, orthonormal=False)
tv, tt = synthetic.sphere_mesh(N)
model = api.Model(ctx, ref, mean, basis, var, tri)
tgt = api.Target(ctx, synthetic.make_target(tv, 0), tt)
reg = api.CpdRegistration(ctx, model, tgt, api.CpdConfiguration(maxIterations=10 ** 6, w=0.1))
reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
reg.updateChain(3)
reg.setProfiling(True)
reg.updateChain(10)
ctx.synchronize()
ms, it = reg.getProfile()
print(json.dumps({"M": M, "N": N, "A": os.environ.get("GINGR_ESTEP_SPLITS_A"), "B": os.environ.get("GINGR_ESTEP_SPLITS_B"),
                  "sweepA_ms": ms[0] / it, "sweepB_ms": ms[1] / it}))
