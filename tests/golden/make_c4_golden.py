"""Golden values of the seeded C4 chain (BASELINE configs[3]: CPD, M = 20 000, N = 200 000, rank 2000, w = 0.1,
RigidTransforms; inputs = bench.make_inputs("c4")) computed by the CPU oracle in its algorithmic-minimum form
(oracle/fast.py, pinned to the literal restatement by tests/test_oracle_fast.py).  One oracle iteration at this size
takes ~25 s on 8 cores, so the values are cached here for
  * tests/test_c4_golden_gpu.py   (the full-size update() parity test on the GPU box), and
  * bench.py's `parity` field      (every N-GPU run re-plays the first iterations and compares).
    python tests/golden/make_c4_golden.py      # ~3 min, writes tests/golden/c4_chain_oracle.npz
These are ORACLE outputs (parity unpinned by the reference, see oracle/oracle.py), not Scala outputs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import fast, oracle  # noqa: E402

ITERS = 3
oracle.build()
ref, mean, basis, var, target = bench.make_inputs("c4")
om = oracle.Gpmm(ref, mean, basis, var, None)
fm = fast.FastCpdModel(om)
algo = oracle.CpdAlgorithm(oracle.CpdConfig(w=bench.W_OUTLIER), literal=False)
st = algo.initialize(oracle.initial_state(om, target, None, global_transformation=oracle.RIGID_TRANSFORMS))
idx = np.arange(0, om.M, 97)
out = {"sigma2": [st.sigma2], "alpha": [], "fit_idx": idx, "fit_sub": [], "fit_sum": [], "translation": [], "euler": [], "scale": [],
       "input_checksum": np.array([ref.sum(), mean.sum(), float(np.abs(basis).sum()), var.sum(), target.sum()])}
for k in range(ITERS):
    t0 = time.time()
    st = fast.propose(fm, algo, st)
    print(f"iteration {k + 1}: {time.time() - t0:.1f} s  sigma2 {st.sigma2!r} status {st.status}")
    assert st.status == oracle.STATUS_NONE
    out["sigma2"].append(st.sigma2)
    out["alpha"].append(st.params.shape.copy())
    out["fit_sub"].append(st.fit[idx].copy())
    out["fit_sum"].append(st.fit.sum(axis=0))
    out["translation"].append(st.params.translation.copy())
    out["euler"].append(np.array(st.params.euler))
    out["scale"].append(st.params.scale)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c4_chain_oracle.npz"), **{k: np.asarray(v) for k, v in out.items()})
print("written")
