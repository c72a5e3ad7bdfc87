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
    ph = (d[9:11] - d[8:10]).tolist()
    row = {"j": j, "start_us": (d[0] - t0) / 1e3, "updates_done_us": (d[1] - t0) / 1e3, "potrf_us": (d[3] - d[2]) / 1e3,
           "flag_us": (d[4] - t0) / 1e3, "potrf_cycles": ph}
    if j + 1 < nbr:
        e = st[(j + 1) * nb + j]
        row["below"] = {"start_us": (e[0] - t0) / 1e3, "updates_done_us": (e[1] - t0) / 1e3, "gotZ_us": (e[2] - t0) / 1e3, "flag_us": (e[4] - t0) / 1e3}
    rows.append(row)
steps = [rows[k + 1]["flag_us"] - rows[k]["flag_us"] for k in range(nb - 1)]
print(json.dumps({"n": n, "ms": o["ms"], "step_us_mean": float(np.mean(steps)) if steps else None,
                  "potrf_us_mean": float(np.mean([r["potrf_us"] for r in rows])),
                  "phases_cycles_mean(elimination, unpack)": np.mean([r["potrf_cycles"] for r in rows], axis=0).round(0).tolist()}))
for r in rows[:4] + rows[-3:]:
    print(json.dumps(r))
