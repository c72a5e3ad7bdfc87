"""ICP update() at M = N = 200 000 on 1..8 GPUs (torchrun): the K2 correspondence search splits its queries across the
ranks (each rank searches the vertices of its own basis shard).  Prints one JSON line on rank 0.
usage: [torchrun ...] python tools/time_icp_mgpu.py [M] [rank] [method]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gingr_b200 import api, synthetic
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
r = int(sys.argv[2]) if len(sys.argv) > 2 else 64
method = sys.argv[3] if len(sys.argv) > 3 else "TRIANGULAR_CLOSEST_POINT"
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
ctx = api.Context(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [api.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
ref, tri = synthetic.sphere_mesh(M)
mean, basis, var = synthetic.make_gpmm(ref, r, 1, orthonormal=False)
tv, tt = synthetic.sphere_mesh(N)
target = synthetic.make_target(tv, 0)
model = api.Model(ctx, ref, mean, basis, var, tri)
tgt = api.Target(ctx, target, tt)
reg = api.IcpRegistration(ctx, model, tgt, api.IcpConfiguration(maxIterations=10 ** 6, initialSigma=2.0, endSigma=0.5,
                                                                correspondenceMethod=getattr(api, method)))
reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
reg.updateChain(2)
ctx.synchronize()
if world > 1:
    dist.barrier()
iters = 10
t0 = time.perf_counter()
reg.updateChain(iters)
ctx.synchronize()
dt = (time.perf_counter() - t0) / iters
reg.setProfiling(True)
reg.updateChain(iters)
ctx.synchronize()
ms, it = reg.getProfile()
st = reg.downloadState()
if world > 1:
    t = torch.tensor([dt, ms[5] / max(it, 1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt, cp_ms = float(t[0]), float(t[1])
    dist.destroy_process_group()
else:
    cp_ms = ms[5] / max(it, 1)
if rank == 0:
    print(json.dumps({"workload": f"ICP update() M=N={M} rank={r} {method}", "gpus": world, "update_ms": dt * 1e3, "iterations_per_s": 1.0 / dt,
                      "closest_point_ms": cp_ms, "alpha_checksum": float(np.sum(st.modelParameters.shape)), "sigma2": st.sigma2,
                      "finite": bool(np.all(np.isfinite(st.fit)))}))
