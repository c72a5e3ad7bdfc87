"""N > 1 coverage.

CPU (`-m "not gpu"`): world_size-2 gloo processes prove the sharding algebra the library relies on:
the E-step shards TARGET columns (exact denominators per column; P1 / PX / sum Pt1|x|^2 are partial sums combined by
an all-reduce), the posterior shards basis ROWS (partial Gram + rhs, all-reduce), the fit is all-gathered.  The
per-rank arithmetic is the CPU oracle's; the collective is torch.distributed gloo.

GPU (`-m gpu`, needs >= 2 devices, else skipped): tests/mgpu_worker.py under torchrun with NCCL inside the library.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard(n, world, rank):
    base, rem = divmod(n, world)
    b = rank * base + min(rank, rem)
    return b, base + (1 if rank < rem else 0)


def test_shard_range_matches_library_partition():
    from gingr_b200 import api
    for n in (1, 2, 7, 100, 200000):
        for world in (1, 2, 3, 8):
            got = [api.shard_range(n, world, k) for k in range(world)]
            assert got == [_shard(n, world, k) for k in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (b0, c0), (b1, _) in zip(got, got[1:]):
                assert b0 + c0 == b1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gingr_b200 import synthetic
    from oracle import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M, N, r = 90, 131, 17
        ref, _ = synthetic.sphere_mesh(M)
        mean, basis, var = synthetic.make_gpmm(ref, r, 3)
        tv, _ = synthetic.sphere_mesh(N)
        target = synthetic.make_target(tv, 1)
        fit = ref + mean.reshape(-1, 3)
        sigma2, w = 30.0, 0.1
        # ---- E-step sharded by target columns -----------------------------------------------------------
        n0, nl = _shard(N, world, rank)
        # outlier constant uses the GLOBAL M/N (CPD.scala:69-70): evaluate the shard with the literal formula
        d2 = ((target[n0:n0 + nl][None, :, :] - fit[:, None, :]) ** 2).sum(-1)
        K = np.exp(-d2 / (2 * sigma2))
        c = w / (1 - w) * (2 * np.pi * sigma2) ** 1.5 * M / N
        P = K / (K.sum(0) + c)[None, :]
        part = np.concatenate([P.sum(1), (P @ target[n0:n0 + nl]).ravel(),
                               [float((P.sum(0) * (target[n0:n0 + nl] ** 2).sum(1)).sum())]])
        t = torch.from_numpy(part.copy())
        dist.all_reduce(t)
        P1, Pt1, PX = oracle.cpd_estep(fit, target, sigma2, w)
        full = np.concatenate([P1, PX.ravel(), [float((Pt1 * (target ** 2).sum(1)).sum())]])
        e_estep = float(np.max(np.abs(t.numpy() - full) / np.maximum(np.abs(full), 1e-300)))
        # ---- Gram + rhs sharded by basis rows -----------------------------------------------------------
        m0, ml = _shard(M, world, rank)
        Q = basis * np.sqrt(var)[None, :]
        wrow = np.repeat(P1 / (sigma2 * 1.0), 3)
        u = np.random.default_rng(0).normal(size=3 * M)
        rows = slice(3 * m0, 3 * (m0 + ml))
        G = torch.from_numpy(np.concatenate([((Q[rows] * wrow[rows, None]).T @ Q[rows]).ravel(),
                                             Q[rows].T @ (wrow[rows] * u[rows])]))
        dist.all_reduce(G)
        Gf = np.concatenate([((Q * wrow[:, None]).T @ Q).ravel(), Q.T @ (wrow * u)])
        e_gram = float(np.max(np.abs(G.numpy() - Gf)) / np.max(np.abs(Gf)))
        # ---- fit all-gather with padded shards ----------------------------------------------------------
        mmax = -(-M // world)
        loc = np.zeros(3 * mmax)
        loc[:3 * ml] = fit[m0:m0 + ml].ravel()
        outs = [torch.zeros(3 * mmax, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(outs, torch.from_numpy(loc))
        gathered = np.concatenate([outs[k].numpy()[:3 * _shard(M, world, k)[1]] for k in range(world)])
        e_gather = float(np.max(np.abs(gathered - fit.ravel())))
        q.put((rank, e_estep, e_gram, e_gather))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_partials_allreduce_to_full_result_gloo(world):
    import torch.multiprocessing as mp
    from oracle import oracle
    oracle.build()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctxm.Process(target=_gloo_worker, args=(k, world, port, q)) for k in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_estep, e_gram, e_gather in res:
        assert e_estep < 1e-12, (rank, e_estep)
        assert e_gram < 1e-12, (rank, e_gram)
        assert e_gather == 0.0, (rank, e_gather)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_multi_gpu_update_matches_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    # exact scans and the uniform grids of the K2 search (the grids are otherwise only used from 8192 vertices up)
    for grid in ("0", "1"):
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, GINGR_K2_GRID=grid))
        sys.stdout.write(p.stdout[-4000:])
        sys.stderr.write(p.stderr[-4000:])
        assert p.returncode == 0 and "MGPU_PARITY_OK" in p.stdout, f"GINGR_K2_GRID={grid}"


def _gloo_chain_worker(rank, world, port, n_chains, steps, q):
    """Replica-only division of independent MH chains (BASELINE config 5): every rank advances ITS chains with the
    oracle's chain step (the device consumes the same Philox stream), results are gathered with gloo."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gingr_b200 import api, synthetic
        from oracle import oracle
        first, count = api.chain_range(n_chains, world, rank)
        ref, tri = synthetic.sphere_mesh(40)
        mean, basis, var = synthetic.make_gpmm(ref, 6, 1)
        tv, tt = synthetic.sphere_mesh(45)
        target = synthetic.make_target(tv, 0)
        om = oracle.Gpmm(ref, mean, basis, var, tri)
        settings = oracle.McmcSettings(uncertainty=1.5, random_mixture=0.5)
        out = torch.zeros(n_chains, 3, dtype=torch.float64)
        for k in range(first, first + count):
            algo = oracle.IcpAlgorithm(oracle.IcpConfig(initial_sigma=2.0, end_sigma=0.5, max_iterations=20))
            st = algo.initialize(oracle.initial_state(om, target, tt, global_transformation=oracle.RIGID_TRANSFORMS))
            lp = oracle.log_value(settings, st)
            acc = 0
            for s in range(steps):
                st, lp, info = oracle.mcmc_step(algo, settings, st, lp, s, seed=100 + k)
                acc += int(info["accept"])
            out[k] = torch.tensor([float(acc), lp[0] + lp[1], float(np.sum(st.params.shape))])
        dist.all_reduce(out)          # every chain was written by exactly one rank
        q.put((rank, first, count, out.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_chain_range_partitions_every_chain_once():
    from gingr_b200 import api
    for n in (1, 7, 128, 1024, 1027):
        for world in (1, 2, 3, 8):
            seen = []
            for rank in range(world):
                first, count = api.chain_range(n, world, rank)
                seen.extend(range(first, first + count))
            assert seen == list(range(n))
            counts = [api.chain_range(n, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


@pytest.mark.parametrize("world", [2, 3])
def test_divided_chains_equal_undivided_chains_gloo(world):
    """The same chains, whichever rank runs them: world-size independence of config 5."""
    import torch.multiprocessing as mp
    from oracle import oracle
    oracle.build()
    n_chains, steps = 5, 3
    ctxm = mp.get_context("spawn")
    results = {}
    for w in (1, world):
        q = ctxm.Queue()
        port = 29700 + 10 * w + (os.getpid() % 100)
        procs = [ctxm.Process(target=_gloo_chain_worker, args=(k, w, port, n_chains, steps, q)) for k in range(w)]
        for p in procs:
            p.start()
        res = [q.get(timeout=240) for _ in range(w)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        results[w] = res[0][3]
        for r in res:
            assert np.array_equal(r[3], res[0][3])
    assert np.array_equal(results[1], results[world])
    assert len(set(results[1][:, 2].tolist())) > 1          # different seeds, different chains
