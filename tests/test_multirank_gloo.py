"""world_size-2 gloo test (CPU): partitions shard across ranks with no data-path collective, and ONE allreduce of the
packed buffer reproduces the single-process combine (dlsa/dlsa.py:30-59).  Per-rank partition fits come from the
oracle here (the CUDA kernels are covered by the -m gpu tests); the code under test is dlsa_b200/parallel.py, the
exact helpers engine.combine_partitions uses on the GPUs with NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import games_data
from oracle import numpy_dlsa as onp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, K, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dlsa_b200 import parallel
    X, y, _, g = games_data()
    pid = np.arange(X.shape[0]) % K
    lo, hi = parallel.shard_partitions(K, world, rank)
    coefs, sigs = [], []
    for k in range(lo, hi):
        c, S, _ = onp.logistic_fit(X[pid == k], y[pid == k])
        coefs.append(c)
        sigs.append(S)
    coefs = torch.as_tensor(np.stack(coefs))
    sigs = torch.as_tensor(np.stack(sigs))
    smc = torch.einsum("kij,kj->ki", sigs, coefs)
    ok = torch.ones(hi - lo, dtype=torch.bool)
    if rank == 1:
        ok[-1] = False  # a skipped partition contributes nothing but still counts in K_total
    packed = parallel.pack_host(sigs, smc, coefs, ok)
    parallel.allreduce_packed(packed)
    S, v, c, K_eff, K_total = parallel.unpack(packed, coefs.shape[1])
    if rank == 0:
        ret["S"], ret["v"], ret["c"], ret["K_eff"], ret["K_total"] = S.numpy().copy(), v.numpy().copy(), c.numpy().copy(), K_eff, K_total
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_combine_equals_single_process():
    K, world = 5, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), K, ret), nprocs=world, join=True)
    X, y, _, g = games_data()
    pid = np.arange(X.shape[0]) % K
    # single-process answer over the same effective partitions (partition K-1 skipped)
    coefs, sigs = zip(*[onp.logistic_fit(X[pid == k], y[pid == k])[:2] for k in range(K - 1)])
    theta, _, S = onp.combine(np.stack(coefs), np.stack(sigs))
    assert ret["K_eff"] == K - 1 and ret["K_total"] == K
    assert np.linalg.norm(ret["S"] - S) / np.linalg.norm(S) < 1e-14
    th = np.linalg.solve(ret["S"], ret["v"])
    assert np.max(np.abs(th - theta)) < 1e-10
    assert np.allclose(ret["c"] / K, np.stack(coefs).sum(0) / K)


def test_shard_partitions_covers_everything_once():
    from dlsa_b200 import parallel
    for K in (1, 5, 16, 120, 1024, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = parallel.shard_partitions(K, world, r)
                seen += list(range(lo, hi))
            assert seen == list(range(K))
