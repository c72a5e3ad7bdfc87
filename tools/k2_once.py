"""One ICP update() per correspondence flavour at M = N = 100k with the uniform grids -- the command profiled for
profiles/*_launches_k2_grid.md (ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GINGR_K2_GRID"] = os.environ.get("GINGR_K2_GRID", "1")
os.environ["GINGR_CUDA_GRAPH"] = "0"
from gingr_b200 import api, synthetic
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ref, tri = synthetic.sphere_mesh(M)
mean, basis, var = synthetic.make_gpmm(ref, 16, 1, orthonormal=False)
tv, tt = synthetic.sphere_mesh(N)
target = synthetic.make_target(tv, 0)
ctx = api.Context(0)
model = api.Model(ctx, ref, mean, basis, var, tri)
tgt = api.Target(ctx, target, tt)
for method in sys.argv[2:] or ("POINTCLOUD_CLOSEST_POINT", "TRIANGULAR_CLOSEST_POINT", "ALONG_NORMAL_CLOSEST_POINT"):
    cfg = api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0, correspondenceMethod=getattr(api, method))
    reg = api.IcpRegistration(ctx, model, tgt, cfg)
    reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    reg.updateChain(1)
    ctx.synchronize()
    reg.close()
