"""Multi-GPU parity worker, launched by tests/test_multi_gpu.py (and by hand) under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_worker.py

Every rank joins the library's NCCL communicator, uploads the same model / target (the library keeps only its
shard: target columns for the E-step, basis rows for Gram / fit), and runs CPD and ICP updates.  Rank 0 checks the
result against the CPU oracle with the single-GPU tolerances; all ranks must agree bit-for-bit with rank 0 (the
all-reduced quantities are identical on every rank by construction)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from gingr_b200 import api, synthetic

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    uid = [api.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])

    M, N, r = 301, 403, 61  # odd sizes: uneven shards
    ref, tri = synthetic.sphere_mesh(M)
    mean, basis, var = synthetic.make_gpmm(ref, r, 1)
    tv, tt = synthetic.sphere_mesh(N)
    target = synthetic.make_target(tv, 0)
    model = api.Model(ctx, ref, mean, basis, var, tri)
    tgt = api.Target(ctx, target, tt)
    diag = float(np.linalg.norm(ref.max(0) - ref.min(0)))
    ok = True
    results = {}
    ICP_METHODS = {"icp": api.TRIANGULAR_CLOSEST_POINT, "icp_pointcloud": api.POINTCLOUD_CLOSEST_POINT,
                   "icp_normal": api.ALONG_NORMAL_CLOSEST_POINT, "icp_reversed": api.TRIANGULAR_CLOSEST_POINT}
    NAMES = ("cpd",) + tuple(ICP_METHODS)   # the ICP flavours split their QUERIES across the ranks (reversed: replicated)
    for name in NAMES:
        if name == "cpd":
            reg = api.CpdRegistration(ctx, model, tgt, api.CpdConfiguration(w=0.1))
            reg.setLandmarks([3, 250], target[[5, 300]], None)
        else:
            reg = api.IcpRegistration(ctx, model, tgt, api.IcpConfiguration(initialSigma=2.0, endSigma=0.5,
                                                                            correspondenceMethod=ICP_METHODS[name],
                                                                            reverseCorrespondenceDirection=(name == "icp_reversed")))
        st = reg.initializeState(globalTransformation=api.RIGID_TRANSFORMS)
        for _ in range(3):
            st = reg.propose(st)
        reg.updateChain(2)
        fin = reg.downloadState()
        results[name] = (st, fin)
        # every rank holds the same state
        v = torch.tensor(np.concatenate([fin.fit.ravel(), fin.modelParameters.shape, [fin.sigma2]]), device="cuda")
        ref_v = v.clone()
        dist.broadcast(ref_v, src=0)
        if not torch.equal(v, ref_v):
            print(f"rank {rank}: {name} state differs from rank 0: {float((v - ref_v).abs().max())}")
            ok = False
        reg.close()
    if rank == 0:
        from oracle import oracle
        oracle.build()
        om = oracle.Gpmm(ref, mean, basis, var, tri)
        OMETHOD = {"icp": oracle.METHOD_TRIANGULAR, "icp_pointcloud": oracle.METHOD_POINTCLOUD, "icp_normal": oracle.METHOD_ALONG_NORMAL,
                   "icp_reversed": oracle.METHOD_TRIANGULAR}
        for name in NAMES:
            if name == "cpd":
                algo = oracle.CpdAlgorithm(oracle.CpdConfig(w=0.1))
                lm = oracle.Landmarks(np.array([3, 250], dtype=np.int32), target[[5, 300]], np.tile(np.eye(3), (2, 1, 1)))
                ost = algo.initialize(oracle.initial_state(om, target, tt, global_transformation=oracle.RIGID_TRANSFORMS,
                                                           landmarks=lm))
            else:
                algo = oracle.IcpAlgorithm(oracle.IcpConfig(initial_sigma=2.0, end_sigma=0.5, method=OMETHOD[name],
                                                            reverse=(name == "icp_reversed")))
                ost = algo.initialize(oracle.initial_state(om, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
            for _ in range(3):
                ost = oracle.propose(algo, ost)
            st, fin = results[name]
            for label, g, o in (("after 3 host updates", st, ost),):
                e_fit = float(np.max(np.abs(g.fit - o.fit))) / diag
                e_a = float(np.max(np.abs(g.modelParameters.shape - o.params.shape)) / max(np.max(np.abs(o.params.shape)), 1e-12))
                e_s = abs(g.sigma2 - o.sigma2) / abs(o.sigma2)
                print(f"mgpu world={world} {name} {label}: fit {e_fit:.2e} alpha {e_a:.2e} sigma2 {e_s:.2e} status {g.status}/{o.status}")
                ok = ok and e_fit < 1e-6 and e_a < 1e-6 and e_s < 1e-6 and g.status == o.status
            for _ in range(2):
                ost = oracle.propose(algo, ost)
            e_fit = float(np.max(np.abs(fin.fit - ost.fit))) / diag
            print(f"mgpu world={world} {name} after 2 chained updates: fit {e_fit:.2e} iteration {fin.iteration}/{ost.iteration}")
            # gingr_update leaves `iteration += 1` to the host-side propose (include/gingr_cuda.h): the bump of the
            # last host-driven propose never reached the device counter the chain continues from
            ok = ok and e_fit < 1e-6 and fin.iteration == ost.iteration - 1
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.destroy_process_group()
    if int(flag) != 1:
        sys.exit(1)
    if rank == 0:
        print("MGPU_PARITY_OK")


if __name__ == "__main__":
    main()
