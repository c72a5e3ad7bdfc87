"""Critical-path breakdown of the data-flow Cholesky (tuning build with -DDF_TIMING):
  tools/build_variant.sh dft chol_df.cu -DDF_TIMING && GINGR_CUDA_LIB=$PWD/gingr_b200/lib/variants/libgingr_cuda_dft.so python tools/chol_df_timing.py [n]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gingr_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = api.Context(0)
rng = np.random.default_rng(n)
Q = rng.normal(size=(n, n + 8))
A = Q @ Q.T / (n + 8) + np.eye(n) * 1e-3
b = rng.normal(size=(1, n))
api.spd_solve(ctx, A, b, reps=3)
o = api.spd_solve(ctx, A, b, reps=1)
nb = (n + 63) // 64
nbr = (n + 1 + 63) // 64
buf = (ctypes.c_ulonglong * (64 * 64 * 16))()
ctx._lib.gingr_debug_chol_df_timing(buf, 64 * 64 * 16)
st = np.array(buf, dtype=np.uint64).reshape(64 * 64, 16).astype(np.int64)
t0 = st[0 * nb + 0, 0]
rows = []
for j in range(nb):
    d = st[j * nb + j]
    rows.append({"j": j, "start_us": (d[0] - t0) / 1e3, "updates_done_us": (d[1] - t0) / 1e3, "gotZprev_us": (d[2] - t0) / 1e3 if j else None,
                 "Xflag_us": (d[3] - t0) / 1e3 if j else None, "potrf_start_us": (d[4] - t0) / 1e3, "potrf_us": (d[5] - d[4]) / 1e3,
                 "flag_us": (d[6] - t0) / 1e3, "potrf_cycles": int(d[9] - d[8])})
steps = [rows[k + 1]["flag_us"] - rows[k]["flag_us"] for k in range(nb - 1)]
print(json.dumps({"n": n, "ms": o["ms"], "step_us_mean": float(np.mean(steps)) if steps else None,
                  "potrf_us_mean": float(np.mean([r["potrf_us"] for r in rows])),
                  "potrf_cycles_mean": float(np.mean([r["potrf_cycles"] for r in rows]))}))
for r in rows[:4] + rows[-3:]:
    print(json.dumps(r))

if hasattr(ctx._lib, "gingr_debug_potrf_profile"):
    pb = (ctypes.c_longlong * (32 * 8))()
    ctx._lib.gingr_debug_potrf_profile(pb)
    pf = np.array(pb, dtype=np.int64).reshape(32, 8)
    base = pf[0, 0]
    print("potrf profile of tile 0, cycles since the first chain start: owner chain start / factors ready / released | next owner: acquired / pivot pair updated")
    for s in range(0, 32):
        print(s, (pf[s, :5] - base).tolist())
