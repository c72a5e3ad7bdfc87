"""(Under torchrun: the chains are divided among the ranks -- replicas, no collective -- and rank 0 prints the aggregate.)
BASELINE config 5 as a real MCMC benchmark: n independent Metropolis-Hastings chains (informed ICP proposal mixed
with the random pose / shape proposals, point-distance + prior evaluators, accept / reject on the device) of the C1
size on one GPU (gingr_mcmc_batch).   usage: python tools/time_mcmc.py [n_chains] [iters] [random_mixture]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gingr_b200 import api, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rho = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
M, N, r = 100, 100, 50
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = n
first_chain, n = api.chain_range(n_total, world, rank)
ctx = api.Context(local)
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
api.mcmc_batch(chains, 3, seed=first_chain)      # warm-up: primes the chains and captures the step graphs
ctx.synchronize()
if world > 1:
    dist.barrier()
l0 = ctx.launch_count
t0 = time.perf_counter()
api.mcmc_batch(chains, iters, seed=first_chain)
ctx.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    tt_ = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
    dt = float(tt_[0])
launches = ctx.launch_count - l0
n_local = n
acc, leaves, lp0, lpb = 0, np.zeros(10, dtype=np.int64), [], []
for c in chains:
    v, k = c.mcmcStats()
    assert k[0] == iters + 3
    acc += int(k[3])
    leaves += k[8:18]
    lpb.append(v[8])
    lp0.append(v[0] + v[1])
st = chains[-1].downloadState()
assert np.all(np.isfinite(st.fit))
if world > 1:
    agg = torch.tensor([float(acc)] + [float(x) for x in leaves] + [float(np.sum(lp0)), float(np.sum(lpb))], dtype=torch.float64, device="cuda")
    dist.all_reduce(agg)
    acc, leaves = int(agg[0]), agg[1:11].cpu().numpy().astype(np.int64)
    lp0, lpb = [float(agg[11]) / n_total], [float(agg[12]) / n_total]
    dist.destroy_process_group()
    if rank != 0:
        sys.exit(0)
    n = n_total
print(json.dumps({"workload": "Metropolis-Hastings ICP chains, M=N=100 r=50 (C5 shape)", "gpus": world, "chains": n, "mh_steps": iters,
                  "random_mixture": rho, "chain_steps_per_s": n * iters / dt, "ms_per_batch_step": dt / iters * 1e3,
                  "launches_per_chain_step": launches / (n_local * iters), "acceptance_rate": acc / (n * (iters + 3)),
                  "proposals_per_leaf": leaves.tolist(), "mean_log_value_current": float(np.mean(lp0)),
                  "mean_log_value_best": float(np.mean(lpb))}))
