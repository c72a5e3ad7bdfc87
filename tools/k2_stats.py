"""Search statistics of the uniform-grid K2 kernels (stats variant of the library: tools/build_variant.sh stats grid.cu
-DGINGR_GRID_STATS; run with GINGR_CUDA_LIB=gingr_b200/lib/variants/libgingr_cuda_stats.so)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GINGR_K2_GRID"] = "1"
os.environ["GINGR_CUDA_GRAPH"] = "0"
from gingr_b200 import api, synthetic, _native
lib = _native.load()
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ref, tri = synthetic.sphere_mesh(M)
mean, basis, var = synthetic.make_gpmm(ref, 16, 1, orthonormal=False)
tv, tt = synthetic.sphere_mesh(N)
target = synthetic.make_target(tv, 0)
ctx = api.Context(0)
model = api.Model(ctx, ref, mean, basis, var, tri)
tgt = api.Target(ctx, target, tt)
names = ["searches", "seed_cells", "ball_cells", "ball_cells_passed", "ball_candidates", "full_triangle_evals", "ball_level_sum", "fallback_scans",
         "w_queries", "w_seed_cells", "w_seed_entries", "w_refine_entries", "w_ball_cells", "w_ball_cells_passed", "w_ball_entries",
         "w_exact_evals", "w_flushes", "w_ball_level_sum", "w_fine_cells_passed", "w11", "w12", "w13", "w14", "w15"]
def stats(reset=True):
    buf = (ctypes.c_ulonglong * 24)()
    lib.gingr_debug_grid_stats(buf, 1 if reset else 0)
    return dict(zip(names, list(buf)))
out = {}
for method in ("TRIANGULAR_CLOSEST_POINT",):
    cfg = api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0, correspondenceMethod=getattr(api, method))
    reg = api.IcpRegistration(ctx, model, tgt, cfg)
    reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    stats()
    for it in range(3):
        reg.updateChain(1)
        st = stats()
        den = max(st["searches"], st["w_queries"], 1)
        out[f"{method}/iter{it}"] = {k: round(v / den, 2) for k, v in st.items() if v} | {"searches_total": den}
    reg.close()
print(json.dumps(out, indent=1))
