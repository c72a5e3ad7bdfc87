#!/usr/bin/env python
"""bench.py -- GiNGR update() iterations/s on the BASELINE.json workload (C4: CPD, M = 20 000 moving points,
N = 200 000 target points, rank-2000 GPMM, w = 0.1), synthetic seeded inputs.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c4|c4_small|c3c|c1]

One "step" = one GingrAlgorithm.update + fit refresh (GingrGeneratorWrapper.propose) of ONE registration.
With N > 1 (torchrun, one rank per GPU) the same registration is sharded over the ranks -- target points in the
E-step, basis rows in the Gram / fit passes -- and the partial sums are all-reduced with NCCL inside the
library, so scaling is STRONG (total work fixed).  Prints one JSON line (rank 0).

  value   iterations/s with model/target/state resident in HBM: K iterations chained on the device
          (gingr_update_chain), CUDA events on the library's stream, max over ranks.
  e2e     the same through the reference-facing call gingr_update with HOST state in / out (state + alpha H2D,
          state + alpha + fit D2H every step).
  roofline  dominant kernel, timed live with CUDA events recorded around it on the library's stream, in an
            event-instrumented pass of the same chain right after the timed one (the timed chain replays the captured
            iteration graph, which cannot carry per-iteration events).
  cpu_baseline  the CPU oracle (port of the reference path, NOT the JVM) on this box's cores on a bounded sample.

--impl reference times the CPU port as the main line (the JVM reference cannot run here: no JVM, scalismo /
Breeze jars absent).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (M, N, r, description)
    "c4": (20000, 200000, 2000, "CPD M=20000 N=200000 rank=2000 w=0.1 RigidTransforms (BASELINE configs[3], 200k-point CPD)"),
    "c4_small": (4000, 20000, 500, "CPD M=4000 N=20000 rank=500 w=0.1 (reduced C4, smoke only)"),
    "c3c": (1000, 1000, 100, "CPD M=N=1000 rank=100 (DemoMultiResolution size)"),
    "c1": (100, 100, 50, "CPD M=N=100 rank=50 (DemoCPD size)"),
}
W_OUTLIER = 0.1
FLOP_PER_PAIR = 71.0   # two sweeps, exp counted as 22 flop


_INPUT_CACHE = {}


def make_inputs(name: str, seed: int = 0):
    if (name, seed) not in _INPUT_CACHE:
        _INPUT_CACHE[(name, seed)] = _make_inputs(name, seed)
    return _INPUT_CACHE[(name, seed)]


def _make_inputs(name: str, seed: int = 0):
    from gingr_b200 import synthetic
    M, N, r, _ = WORKLOADS[name]
    ref = synthetic.fibonacci_sphere(M)
    target = synthetic.make_target(synthetic.fibonacci_sphere(N), seed)
    mean, basis, var = synthetic.make_gpmm(ref, r, seed + 1, orthonormal=(M <= 20000 and r <= 600))
    return ref, mean, basis, var, target


def load_peaks():
    peaks = {}
    for fn in ("MEASURED_PEAKS.json", "FP64_PEAKS.json"):
        p = os.path.join(ROOT, fn)
        if os.path.exists(p):
            with open(p) as f:
                peaks[fn] = json.load(f)
    return peaks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            hi = [s for s in sm if s >= 0.5 * max(sm)]   # samples under load
            out["sm_mhz"] = statistics.median(hi)
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------------
# CPU port (oracle) timing -- only the cpu_baseline / --impl reference legs touch oracle/
# ---------------------------------------------------------------------------------------------------
def cpu_iteration_seconds(name: str, budget_s: float = 20.0):
    """Seconds per update() of the CPU port in algorithmic-minimum mode (one streaming E-step, one weighted
    Gram, SVD pseudo-inverse, two projections), measured on bounded samples and scaled by the exact work ratio:
    E-step on a column subsample (cost is linear in N), Gram on a row subsample (linear in 3M), the r x r algebra
    at full size."""
    from oracle import oracle
    oracle.build()
    M, N, r, _ = WORKLOADS[name]
    ref, mean, basis, var, target = make_inputs(name)
    cores = oracle.num_threads()
    sample = {}
    # E-step: choose N_s so that the sample takes a few seconds (~ 2 * M * N_s exp at ~1e8/s/core)
    n_s = int(min(N, max(256, 3.0 * 1e8 * cores / (2.0 * M))))
    fit = ref + mean.reshape(-1, 3)
    t0 = time.perf_counter()
    P1, Pt1, PX = oracle.cpd_estep(fit, target[:n_s], 25.0, W_OUTLIER, fast=True)
    t_e = (time.perf_counter() - t0) * (N / n_s)
    sample["estep"] = f"{M}x{n_s} of {M}x{N} pairs"
    # Gram: rows subsample
    m_s = int(min(M, max(64, 2.0e10 * 1.0 / (2.0 * 3 * r * r))))
    rows = 3 * m_s
    Q = basis[:rows] * np.sqrt(var)[None, :]
    wts = np.repeat(np.random.default_rng(0).uniform(0.5, 2.0, m_s), 3)
    t0 = time.perf_counter()
    Mx = (Q * wts[:, None]).T @ Q + np.eye(r)
    t_g = (time.perf_counter() - t0) * (M / m_s)
    sample["gram"] = f"{rows}x{r} of {3 * M}x{r} rows"
    # r x r algebra at full size: pinv (SVD) as Breeze does, plus the matrix-vector products
    t0 = time.perf_counter()
    Minv = oracle.breeze_pinv(Mx)
    c = Minv @ np.ones(r)
    t_s = time.perf_counter() - t0
    # HBM-like passes: 3 basis passes (rhs, instances, projection) -- measured on the row subsample
    t0 = time.perf_counter()
    v = Q @ c
    q = Q.T @ v
    v2 = Q @ q
    t_p = (time.perf_counter() - t0) * (M / m_s)
    total = t_e + t_g + t_s + t_p
    return total, cores, ("update() of the CPU port (C oracle + numpy/OpenBLAS), algorithmic-minimum mode; "
                          f"E-step sampled on {sample['estep']}, Gram and basis passes on {sample['gram']}, each scaled by "
                          f"the exact work ratio; r x r pseudo-inverse at full size; parts: estep {t_e:.2f}s gram {t_g:.2f}s "
                          f"pinv {t_s:.2f}s passes {t_p:.2f}s")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    M, N, r, desc = WORKLOADS[name]
    steps = max(1, args.steps)
    vals = []
    note = ""
    cores = 1
    for k in range(args.warmup + steps):
        # each step is one bounded-sample measurement; keep the whole run within a few minutes
        if k >= 1 and time.time() - run_reference.t0 > 150:
            break
        sec, cores, note = cpu_iteration_seconds(name)
        if k >= min(args.warmup, 1):
            vals.append(1.0 / sec)
    v = statistics.median(vals)
    line = {
        "impl": "reference", "metric": "GiNGR update() iterations/s", "value": v, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": len(vals), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 / v,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "M": M, "N": N, "rank": r, "w": W_OUTLIER, "note": "CPU port of the reference path (no JVM in this image)"},
        "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": note},
        "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


run_reference.t0 = time.time()


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from gingr_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    name = args.workload
    M, N, r, desc = WORKLOADS[name]
    ref, mean, basis, var, target = make_inputs(name)

    ctx = api.Context(local_rank)
    if world > 1:
        uid = [api.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
    model = api.Model(ctx, ref, mean, basis, var)
    tgt = api.Target(ctx, target)
    reg = api.CpdRegistration(ctx, model, tgt, api.CpdConfiguration(maxIterations=10 ** 6, w=W_OUTLIER))
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident chain: `value` ---------------------------------------------------------------
    # The timed region replays the captured iteration graph (the product path).  The per-kernel times of the
    # roofline come from a second, event-instrumented pass of the same chain right after it (events cannot live
    # inside the replayed graph); its total is reported as phases_ms.iteration next to ms_per_step.
    state0 = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    reg.updateChain(max(args.warmup, 3))
    barrier()
    l0 = ctx.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    reg.updateChain(args.steps)
    e1.record(stream)
    barrier()
    launches = ctx.launch_count - l0
    ms_chain = e0.elapsed_time(e1)
    reg.setProfiling(True)
    barrier()
    reg.updateChain(args.steps)
    barrier()
    clocks = sampler.stop()
    prof_ms, prof_it = reg.getProfile()
    reg.setProfiling(False)
    final = reg.downloadState()
    if not np.all(np.isfinite(final.fit)) or final.status == api.STATUS_MODEL_FLEXIBILITY_ERROR:
        raise SystemExit("bench.py: registration diverged / ModelFlexibilityError inside the timed region")

    # ---- host in / host out through gingr_update: `e2e` -------------------------------------------------
    st = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
    for _ in range(max(args.warmup, 3)):
        st = reg.propose(st)
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(args.steps):
        st = reg.propose(st)
    e3.record(stream)
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), 0.0)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(ms_e2e, wall_e2e if world == 1 else ms_e2e)

    if world > 1:
        t = torch.tensor([ms_chain, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_chain, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = load_peaks()
        fp = peaks.get("FP64_PEAKS.json", {})
        mb = fp.get("microbench", {})
        dfma_peak = mb.get("dfma_tflops_sustained") or mb.get("dfma_tflops") or 33.7
        dmma_peak = fp.get("cublas_dgemm_tflops_sustained") and max(fp.get("cublas_dgemm_tflops_sustained"), mb.get("dmma_tflops_sustained", 0)) or 37.0
        it = max(prof_it, 1)
        n_local = api.shard_range(N, world, 0)[1]
        m_local = api.shard_range(M, world, 0)[1]
        t_a, t_b, t_g, t_c, t_it = (prof_ms[k] / it for k in (0, 1, 2, 3, 4))
        estep_flop = FLOP_PER_PAIR * M * n_local
        gram_flop = 3.0 * m_local * r * (r + 1)
        estep_tf = estep_flop / ((t_a + t_b) * 1e-3) / 1e12 if (t_a + t_b) > 0 else 0.0
        gram_tf = gram_flop / (t_g * 1e-3) / 1e12 if t_g > 0 else 0.0
        # DRAM bytes per launch from the committed ncu --set full capture of this workload (profiles/ncu_traffic.json);
        # only valid for the configuration it was captured on (C4, 1 GPU)
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if name == "c4" and world == 1 and os.path.exists(tpath):
            with open(tpath) as f:
                traffic = {k: v["dram_bytes"] for k, v in json.load(f)["kernels"].items()}
        tr_estep = (traffic.get("estep_colsum_kernel", 0) + traffic.get("estep_rowsum_kernel", 0)) or None
        rl_estep = {"kernel": "estep_colsum_kernel+estep_rowsum_kernel (K1)", "bound": "fp64", "achieved": estep_tf,
                    "peak": dfma_peak, "unit": "TFLOP/s", "frac": estep_tf / dfma_peak, "traffic": tr_estep,
                    "ms": t_a + t_b, "algorithmic_flop": estep_flop,
                    "algorithmic_bytes": 24.0 * (M + n_local) + 8.0 * (4 * M + n_local),
                    "peak_source": "FP64_PEAKS.json dfma_tflops_sustained (measured, tools/fp64_peaks.cu)",
                    "note": "achieved = SURVEY 8(d) ALGORITHMIC count (71 flop per pair, exp = 22) / time; the kernels issue 14 + 17 "
                            "FP64 instructions per pair (expanded-distance path, table-driven 2^x), ncu FP64 pipe active 78-79 % "
                            "(profiles/r02_ncu_top_kernels.md)"}
        rl_gram = {"kernel": "gram_ws_kernel (K3, DMMA.8x8x4, warp-specialised, cp.async.bulk)", "bound": "tensor",
                   "achieved": gram_tf, "peak": dmma_peak, "unit": "TFLOP/s", "frac": gram_tf / dmma_peak,
                   "traffic": traffic.get("gram_ws_kernel"), "ms": t_g, "algorithmic_flop": gram_flop,
                   "algorithmic_bytes": 8.0 * 3 * m_local * r,
                   "peak_source": "FP64_PEAKS.json max(cuBLAS DGEMM sustained, DMMA issue peak) (measured)"}
        dominant = rl_estep if (t_a + t_b) >= t_g else rl_gram
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                sec, cores, note = cpu_iteration_seconds(name)
                cpu = {"value": 1.0 / sec, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": note}
            except Exception as ex:  # the oracle is test infrastructure: its absence must not fail the bench
                cpu = {"value": None, "unit": "iterations/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
        h2d = 112 + 8 * r
        d2h = 112 + 8 * r + 24 * M
        line = {
            "metric": "GiNGR update() iterations/s", "value": args.steps / (ms_chain * 1e-3), "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_chain / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "M": M, "N": N, "rank": r, "w": W_OUTLIER, "algorithm": "CPD",
                       "sharding": f"targets/{world} (E-step), basis rows/{world} (Gram, fit); NCCL all-reduce" if world > 1 else "single GPU",
                       "timed_path": "captured CUDA graph of the iteration, replayed K times on the device",
                       "l2": "inputs larger than L2 (basis 8*3M*r bytes = %.0f MB re-read every step)" % (8.0 * 3 * M * r / 1e6)},
            "clocks": clocks,
            "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": "iterations/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": dominant,
            "rooflines": [rl_estep, rl_gram],
            "phases_ms": {"estep_sweepA": t_a, "estep_sweepB": t_b, "gram": t_g, "cholesky_backsolve": t_c,
                          "iteration": t_it},
            "cpu_baseline": cpu,
            "final_sigma2": final.sigma2,
        }
        print(json.dumps(line))
    reg.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
