"""Device-chain iterations/s at C4 under torchrun without the profiling events (so the CUDA graph path is used when
GINGR_CUDA_GRAPH allows it).  usage: torchrun ... tools/time_chain_mgpu.py [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from gingr_b200 import api

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ref, mean, basis, var, target = bench.make_inputs("c4")
ctx = api.Context(local)
if world > 1:
    uid = [api.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
model = api.Model(ctx, ref, mean, basis, var)
tgt = api.Target(ctx, target)
reg = api.CpdRegistration(ctx, model, tgt, api.CpdConfiguration(maxIterations=10 ** 6, w=0.1))
reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
reg.updateChain(3)
ctx.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
reg.updateChain(steps)
ctx.synchronize()
dt = time.perf_counter() - t0
st = reg.downloadState()
if rank == 0:
    print(json.dumps({"world": world, "graph": os.environ.get("GINGR_CUDA_GRAPH", "1"), "it_per_s": steps / dt,
                      "ms_per_iter": dt / steps * 1e3, "sigma2": st.sigma2}))
reg.close()
if world > 1:
    dist.destroy_process_group()
