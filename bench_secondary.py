"""Secondary benchmark entries of bench.py: the other named configurations of BASELINE.json (C1 DemoICP, C3
DemoMultiResolution, C5 batched MCMC chains) and the ICP iteration at the 200k-point scale (K2 closest point with its
roofline).  Each is a short measurement on synthetic seeded inputs of the named shape (the reference's data files do not
travel to the GPU box); the headline number of bench.py stays the C4 line.  Everything goes through the public API of
gingr_b200 (C ABI underneath)."""
from __future__ import annotations

import json
import os
import time

import numpy as np

HBM_FALLBACK_GBS = 6546.9   # MEASURED_PEAKS.json copy bandwidth of this pool's B200 (used when the file is absent)


def _hbm_peak(root):
    p = os.path.join(root, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        for k in ("hbm_copy_gbs", "hbm_gbs", "copy_gbs"):
            if k in d:
                return float(d[k]), "MEASURED_PEAKS.json"
        for v in d.values():
            if isinstance(v, dict):
                for k, x in v.items():
                    if "gb" in k.lower() and isinstance(x, (int, float)):
                        return float(x), "MEASURED_PEAKS.json"
    except Exception:
        pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md / BASELINE.md: 6546.9 GB/s)"


def c1_icp(ctx, iters=99):
    """BASELINE configs[0]: DemoICP shape -- ICP, M = N = 100, sigma2 1 -> 1, NoTransforms, 99 update() calls
    (examples/DemoICP.scala:20-24), deterministic, chained on the device (captured iteration graph)."""
    from gingr_b200 import api, synthetic
    M, N, r = 100, 100, 50
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, 0)
    model = api.Model(ctx, ref, mean, basis, var, tri)
    tgt = api.Target(ctx, target, tt)
    reg = api.IcpRegistration(ctx, model, tgt, api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0))
    reg.initializeState(globalTransformation=api.NO_TRANSFORMS)
    reg.updateChain(20)
    ctx.synchronize()
    best = None
    for _ in range(3):
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        reg.updateChain(iters)
        ctx.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        launches = ctx.launch_count - l0
    st = reg.downloadState()
    ok = bool(np.all(np.isfinite(st.fit)))
    reg.close(); model.close(); tgt.close()
    return {"workload": "ICP M=N=100 rank=50 sigma2 1->1 NoTransforms, 99 updates (DemoICP shape, synthetic spheres)",
            "iterations_per_s": iters / best, "us_per_update": best / iters * 1e6, "registration_ms": best * 1e3,
            "launches_per_update": launches / iters, "finite": ok}


def c3_pipeline(ctx):
    """BASELINE configs[2]: DemoMultiResolution schedule -- CPD 100 -> CPD 500 (50 iterations each, Rigid) -> ICP 1000
    (100 iterations, sigma2 2 -> 0.01) with the hand-over of (pose, shape) between the levels
    (examples/DemoMultiResolution.scala:31-47) on a synthetic 2000 / 2400-vertex pair; the GPMM is built on the device."""
    from gingr_b200 import api, rotation, synthetic
    rv, rt = synthetic.sphere_mesh(2000)
    tv, tt = synthetic.sphere_mesh(2400)
    tv = synthetic.make_target(tv, 0) @ rotation.euler_to_matrix(0.1, 0.1, 0.1).T + np.array([5.0, 5.0, 5.0])
    t0 = time.perf_counter()
    model = api.SimpleTriangleModels3D.create(ctx, rv, rt, api.GaussKernel(scaling=20.0, sigma=60.0))
    target = api.Target(ctx, tv, tt)
    ctx.synchronize()
    t_model = time.perf_counter() - t0
    gi = api.GingrInterface(ctx, model, target)
    stages = {}
    t0 = time.perf_counter()
    coarse = gi.CPD(api.CpdConfiguration(maxIterations=50)).runDecimated(100, 100, globalTransformation=api.RIGID_TRANSFORMS)
    stages["cpd_100_ms"] = (time.perf_counter() - t0) * 1e3
    t1 = time.perf_counter()
    medium = gi.CPD(api.CpdConfiguration(maxIterations=50, initialSigma=coarse.sigma2)).runDecimated(500, 500, generalState=coarse)
    stages["cpd_500_ms"] = (time.perf_counter() - t1) * 1e3
    t1 = time.perf_counter()
    fine = gi.ICP(api.IcpConfiguration(maxIterations=100, initialSigma=2.0, endSigma=0.01)).runDecimated(1000, 1000, generalState=medium)
    stages["icp_1000_ms"] = (time.perf_counter() - t1) * 1e3
    total = time.perf_counter() - t0
    d0 = float(np.sqrt(((rv[:, None, :] - tv[None, ::8, :]) ** 2).sum(-1).min(1)).mean())
    d1 = float(np.sqrt(((fine.fit[:, None, :] - tv[None, ::8, :]) ** 2).sum(-1).min(1)).mean())
    model.close(); target.close()
    return {"workload": "CPD 100 -> CPD 500 -> ICP 1000 (49 + 49 + 99 updates) on a 2000 / 2400-vertex synthetic pair, rank %d; "
                        "includes mesh decimation, model re-referencing and host hand-over (DemoMultiResolution schedule)" % model.rank,
            "pipeline_ms": total * 1e3, "updates": 49 + 49 + 99, "updates_per_s": (49 + 49 + 99) / total, **stages,
            "gpmm_build_ms": t_model * 1e3, "mean_vertex_distance_before": d0, "mean_vertex_distance_after": d1,
            "status": int(fine.status)}


def icp_200k(ctx, root, iters=5, rank=64):
    """ICP update() at M = N = 200 000 (TriangularClosestPoint, the default flavour, and PointcloudClosestPoint): the
    K2 correspondence search at scale with its HBM roofline (algorithmic bytes 24 M + 24 N (+ 12 T') + 29 M)."""
    from gingr_b200 import api, synthetic
    M = N = 200000
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, rank, 1, orthonormal=False)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, 0)
    model = api.Model(ctx, ref, mean, basis, var, tri)
    tgt = api.Target(ctx, target, tt)
    peak, src = _hbm_peak(root)
    out = {"workload": f"ICP update() M=N={M} rank={rank}, uniform-grid search (exact, bit-identical to the scan)", "hbm_peak_gbs": peak,
           "hbm_peak_source": src, "flavours": {}}
    for name, method in (("TriangularClosestPoint", api.TRIANGULAR_CLOSEST_POINT), ("PointcloudClosestPoint", api.POINTCLOUD_CLOSEST_POINT)):
        reg = api.IcpRegistration(ctx, model, tgt, api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=2.0, endSigma=0.5,
                                                                        correspondenceMethod=method))
        reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
        reg.updateChain(2)
        ctx.synchronize()
        t0 = time.perf_counter()
        reg.updateChain(iters)
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / iters
        reg.setProfiling(True)
        reg.updateChain(iters)
        ctx.synchronize()
        ms, it = reg.getProfile()
        reg.setProfiling(False)
        st = reg.downloadState()
        t_cp = ms[5] / max(it, 1)
        T = tt.shape[0]
        alg_bytes = 24.0 * M + 24.0 * N + (12.0 * T if method == api.TRIANGULAR_CLOSEST_POINT else 0.0) + 29.0 * M
        gbs = alg_bytes / (t_cp * 1e-3) / 1e9 if t_cp > 0 else 0.0
        out["flavours"][name] = {"update_ms": dt * 1e3, "iterations_per_s": 1.0 / dt,
                                 "phases_ms": {"closest_point": t_cp, "gram": ms[2] / max(it, 1), "cholesky_backsolve": ms[3] / max(it, 1),
                                               "iteration": ms[4] / max(it, 1)},
                                 "roofline": {"kernel": "K2 grid search (grid_surface_warp_kernel / grid_nn_kernel + predicates)", "bound": "hbm",
                                              "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                                              "algorithmic_bytes": alg_bytes,
                                              "note": "latency / divergence-bound search structure: no exact nearest-neighbour structure reaches the "
                                                      "HBM roofline; the exact scans it replaces are FP64-issue-bound at 8 M N flop "
                                                      "(profiles/r02u_k2_grid.md)"},
                                 "finite": bool(np.all(np.isfinite(st.fit))), "status": int(st.status)}
        reg.close()
    model.close(); tgt.close()
    return out


def c5_mcmc(ctx, world, rank, n_total=1024, iters=50, rho=0.5):
    """BASELINE configs[4]: n_total independent Metropolis-Hastings chains (informed ICP proposal mixed with the random
    pose / shape proposals, point-distance + prior evaluators, accept / reject on the device) of the C1 size, divided among
    the ranks (replicas, no collective).  Returns this rank's (steps, seconds, accepted); bench.py aggregates."""
    from gingr_b200 import api, synthetic
    M, N, r = 100, 100, 50
    first, n = api.chain_range(n_total, world, rank)
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, 0)
    model = api.Model(ctx, ref, mean, basis, var, tri)
    tgt = api.Target(ctx, target, tt)
    cfg = api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0)
    settings = api.ProbabilisticSettings(uncertainty=1.0, randomMixture=rho)
    chains = []
    for _ in range(n):
        c = api.IcpRegistration(ctx, model, tgt, cfg)
        c.configureProbabilistic(settings)
        c.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
        chains.append(c)
    api.mcmc_batch(chains, 3, seed=first)
    ctx.synchronize()
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    api.mcmc_batch(chains, iters, seed=first)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    launches = ctx.launch_count - l0
    acc = 0
    for c in chains:
        v, k = c.mcmcStats()
        acc += int(k[3])
    ok = bool(np.all(np.isfinite(chains[-1].downloadState().fit)))
    for c in chains:
        c.close()
    model.close(); tgt.close()
    return {"chains": n, "steps": iters, "seconds": dt, "accepted": acc, "launches": launches, "finite": ok}


def c5_update_batch(ctx, n=1024, iters=50):
    """The informed proposal alone, batched: n independent probabilistic ICP registrations of the C1 size advancing by
    update(current, probabilistic = true) + propose (gingr_update_batch; one batched kernel sequence per iteration)."""
    from gingr_b200 import api, synthetic
    M, N, r = 100, 100, 50
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    tv, tt = synthetic.sphere_mesh(N)
    model = api.Model(ctx, ref, mean, basis, var, tri)
    tgt = api.Target(ctx, synthetic.make_target(tv, 0), tt)
    cfg = api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=1.0, endSigma=1.0)
    chains = [api.IcpRegistration(ctx, model, tgt, cfg) for _ in range(n)]
    for c in chains:
        c.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    api.update_batch(chains, 3, probabilistic=True, seed=0)
    ctx.synchronize()
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    api.update_batch(chains, iters, probabilistic=True, seed=0)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    launches = ctx.launch_count - l0
    ok = bool(np.all(np.isfinite(chains[-1].downloadState().fit)))
    for c in chains:
        c.close()
    model.close(); tgt.close()
    return {"workload": f"{n} independent probabilistic ICP registrations of the C1 size (informed proposal only), gingr_update_batch",
            "chains": n, "iterations_per_chain": iters, "chain_iterations_per_s": n * iters / dt, "ms_per_iteration_of_all_chains": dt / iters * 1e3,
            "launches_per_iteration_of_all_chains": launches / iters, "finite": ok}
