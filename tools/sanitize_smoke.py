"""Small run of every kernel family added in this round (uniform grids incl. the warp-per-query surface search, whole ICP
iterations with grids, MH chain, GPMM construction, re-referencing) for compute-sanitizer:
    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_smoke.py grid"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GINGR_K2_GRID"] = "1"
os.environ["GINGR_CUDA_GRAPH"] = "0"
import numpy as np
from gingr_b200 import api, synthetic
what = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = api.Context(0)
tv, tt = synthetic.sphere_mesh(300, radius=97.0)
gv, gt = synthetic.sphere_mesh(400)
gv = synthetic.make_target(gv, 3)
target = api.Target(ctx, gv, gt)
for meth in (api.POINTCLOUD_CLOSEST_POINT, api.TRIANGULAR_CLOSEST_POINT, api.ALONG_NORMAL_CLOSEST_POINT):
    api.icp_closest(ctx, target, tv, tt, meth)
    api.icp_closest_reversal(ctx, target, tv, tt, meth)
far = np.concatenate([tv[:40] * 5.0, tv[:10] * 1e4])          # climbs the levels / falls back to the scan
api.icp_closest(ctx, target, far, None, api.POINTCLOUD_CLOSEST_POINT)
print("grid searches ok")
if what == "all":
    ref, tri = synthetic.sphere_mesh(200)
    mean, basis, var = synthetic.make_gpmm(ref, 12, 1)
    model = api.Model(ctx, ref, mean, basis, var, tri)
    for rev in (False, True):
        reg = api.IcpRegistration(ctx, model, target, api.IcpConfiguration(initialSigma=2.0, endSigma=0.5, reverseCorrespondenceDirection=rev))
        reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
        reg.updateChain(2)
        assert np.all(np.isfinite(reg.downloadState().fit))
        reg.close()
    reg = api.IcpRegistration(ctx, model, target, api.IcpConfiguration(initialSigma=2.0, endSigma=0.5))
    reg.configureProbabilistic(api.ProbabilisticSettings(uncertainty=1.5, randomMixture=0.5, mode=api.EVAL_SYMMETRIC))
    reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    reg.mcmcChain(4, 3)
    assert np.all(np.isfinite(reg.mcmcBest().fit))
    reg.close()
    cpd = api.CpdRegistration(ctx, model, target, api.CpdConfiguration(w=0.1))
    cpd.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    cpd.updateChain(2)
    cpd.close()
    print("iterations / MH chain ok")
    m2 = model.newReference(*synthetic.sphere_mesh(90))
    m3 = api.Model.gaussianMixture(ctx, ref[:80], None, [60.0], [30.0], 0.05)
    assert m3.rank > 0 and np.all(np.isfinite(m3.download()[2]))
    m2.close(); m3.close(); model.close()
    print("re-referencing / GPMM construction ok")
target.close()
ctx.close()
print("SANITIZE_SMOKE_DONE")
