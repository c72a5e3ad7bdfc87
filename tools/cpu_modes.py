"""CPU port (oracle) per-iteration time of CPD update() at the C2 / C3 sizes in two modes:
  reference-faithful   what the JVM path executes per iteration: P rebuilt 4x (strict val of the state case class,
                       SURVEY 3.2), sum(P, Axis._1) recomputed for every point in getUncertainty (CPD.scala:120-128:
                       O(M^2 N)), 1 posterior + 2 full `coefficients` regressions, each with an SVD pseudo-inverse and 3
                       model.transform copies of the basis
  algorithmic-minimum  what libgingr_cuda computes: one streaming E-step, one weighted Gram, one solve, two projections
Both are the C / numpy oracle on this machine's cores -- NOT the JVM; they bound the reference's asymptotics from below.
usage: python tools/cpu_modes.py [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gingr_b200 import synthetic
from oracle import oracle

oracle.build()
out = {"cores": oracle.num_threads(), "sizes": {}}
for name, (M, N, r) in {"C2 (DemoCPD 100x100)": (100, 100, 50), "C3a (500x500)": (500, 500, 100), "C3c (1000x1000)": (1000, 1000, 100)}.items():
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, 0)
    om = oracle.Gpmm(ref, mean, basis, var, tri)
    algo = oracle.CpdAlgorithm(oracle.CpdConfig(w=0.1), literal=True)
    st = algo.initialize(oracle.initial_state(om, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
    st = oracle.propose(algo, st)

    def best(fn, reps=3):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts)
    t_P = best(lambda: oracle.cpd_P(st.fit, st.target, st.sigma2, 0.1))
    P = oracle.cpd_P(st.fit, st.target, st.sigma2, 0.1)
    t_rowsum = best(lambda: P.sum(axis=1))
    t_update_literal = best(lambda: oracle.update(algo, st), reps=2)        # 1 posterior + 2 regressions, SVD pinv, P materialised once
    t_transform = best(lambda: st.model.transform(st.params.rotation_matrix(), st.params.translation))
    faithful = t_update_literal + 3 * t_P + M * t_rowsum + 2 * t_transform   # + 3 more P builds, per-point row sums, extra transforms
    algo_min = oracle.CpdAlgorithm(oracle.CpdConfig(w=0.1), literal=False)
    st2 = algo_min.initialize(oracle.initial_state(om, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
    st2 = oracle.propose(algo_min, st2)
    t_min = best(lambda: oracle.update(algo_min, st2), reps=3)
    out["sizes"][name] = {"M": M, "N": N, "rank": r, "P_build_s": t_P, "row_sum_s": t_rowsum, "literal_update_s": t_update_literal,
                          "reference_faithful_s_per_iteration": faithful, "reference_faithful_it_per_s": 1.0 / faithful,
                          "oracle_streaming_update_s": t_min, "oracle_streaming_it_per_s": 1.0 / t_min}
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
