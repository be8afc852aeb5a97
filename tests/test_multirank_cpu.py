"""world_size-2 gloo tests (CPU) of the sharding / gather logic used on the multi-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gperks_b200.parallel import run_jobs_round_robin, scatter_rows, shard_bounds, sharded_predict


def test_shard_bounds_cover_and_balance():
    for M in (0, 1, 7, 8, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(M, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_predict(X):
    # deterministic per-row function, so sharded == unsharded bit for bit
    return np.sin(X).sum(1), np.sqrt(1.0 + (X * X).sum(1))


def _worker(rank, world, port, M, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        X = rng.random((M, 3))
        ref_m, ref_s = _fake_predict(X)
        # the product contract: X* lives on the source rank only, rows are scattered, results gathered back to it
        mean, std = sharded_predict(_fake_predict, X if rank == 0 else None)
        if rank == 0:
            ok = np.array_equal(mean, ref_m) and np.array_equal(std, ref_s)
        else:
            ok = mean is None and std is None
        # another source rank, result broadcast to everyone
        mean, std = sharded_predict(_fake_predict, X if rank == 1 else None, src=1, all_ranks=True)
        ok = ok and np.array_equal(mean, ref_m) and np.array_equal(std, ref_s)
        # every rank already holds X*: no scatter
        mean, std = sharded_predict(_fake_predict, X, replicated=True)
        ok = ok and ((np.array_equal(mean, ref_m) and np.array_equal(std, ref_s)) if rank == 0 else mean is None)
        # the row dealer on its own
        rows, lo, hi, m_all = scatter_rows(X if rank == 0 else None)
        ok = ok and m_all == M and (lo, hi) == shard_bounds(M, world, rank) and np.array_equal(rows.numpy(), X[lo:hi])
        jobs = list(range(5))
        res = run_jobs_round_robin(jobs, lambda j: {"job": j, "rank": dist.get_rank(), "score": j * j})
        ok = ok and [r["job"] for r in res] == jobs and [r["rank"] for r in res] == [j % world for j in jobs]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("M", [1001, 2])
def test_sharded_predict_and_job_dealing_world2(M):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, M, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def test_single_process_paths():
    X = np.random.default_rng(1).random((10, 2))
    m, s = sharded_predict(_fake_predict, X)
    assert np.array_equal(m, _fake_predict(X)[0])
    assert run_jobs_round_robin([1, 2, 3], lambda j: j + 1) == [2, 3, 4]
