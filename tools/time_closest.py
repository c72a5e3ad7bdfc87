"""K2 at scale: closest-point phase of one ICP update() (profile slot 5: correspondence search incl. the grid rebuild
over the moving fit) with the brute-force scans (GINGR_K2_GRID=0) and the uniform grids (GINGR_K2_GRID=1), CUDA events
on the library stream.   usage: python tools/time_closest.py [M] [N] [iters]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gingr_b200 import api, synthetic

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
r = 16
ref, tri = synthetic.sphere_mesh(M)
mean, basis, var = synthetic.make_gpmm(ref, r, 1, orthonormal=False)
tv, tt = synthetic.sphere_mesh(N)
# K2_NEAR=1: the target is the template sphere scaled by 1.001 (near queries: the steady state of a converged ICP)
target = tv * 1.001 if os.environ.get("K2_NEAR") == "1" else synthetic.make_target(tv, 0)
ctx = api.Context(0)
out = {"M": M, "N": N, "T_template": int(len(tri)), "T_target": int(len(tt)), "rank": r, "iters": iters, "results": {}}
fits = {}
for method in ("POINTCLOUD_CLOSEST_POINT", "TRIANGULAR_CLOSEST_POINT", "ALONG_NORMAL_CLOSEST_POINT"):
    for grid in (0, 1):
        os.environ["GINGR_K2_GRID"] = str(grid)
        model = api.Model(ctx, ref, mean, basis, var, tri)
        tgt = api.Target(ctx, target, tt)
        cfg = api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0, correspondenceMethod=getattr(api, method),
                                   reverseCorrespondenceDirection=os.environ.get("K2_REVERSE") == "1")
        reg = api.IcpRegistration(ctx, model, tgt, cfg)
        reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
        reg.updateChain(1)
        ctx.synchronize()
        reg.setProfiling(True)
        reg.updateChain(iters)
        ms, it = reg.getProfile()
        reg.setProfiling(False)
        st = reg.downloadState()
        fits[(method, grid)] = st.fit
        # algorithmic bytes: queries 24 M + searched vertices 24 N (+ 12 T' triangle indices) read once,
        # idx + cp + w written
        nbytes = 24.0 * M + 24.0 * N + (12.0 * len(tt) if method != "POINTCLOUD_CLOSEST_POINT" else 0.0) + 29.0 * M
        t = ms[5] / max(it, 1)
        out["results"][f"{method}/grid={grid}"] = {
            "closest_ms": t, "iteration_ms": ms[4] / max(it, 1), "algorithmic_bytes": nbytes,
            "achieved_GBps": nbytes / (t * 1e-3) / 1e9 if t > 0 else None,
            "bruteforce_flop_8MN_TFLOPs": 8.0 * M * N / (t * 1e-3) / 1e12 if (grid == 0 and t > 0) else None}
        reg.close(); model.close(); tgt.close()
    out["results"][f"{method}/identical_fit"] = bool(np.array_equal(fits[(method, 0)], fits[(method, 1)]))
print(json.dumps(out))
