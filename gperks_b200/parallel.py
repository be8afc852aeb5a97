"""One-process-per-GPU helpers for the parts of the path that shard.

* predict / history matching / GSA evaluation: test points are independent given the replicated factors, so
  rows of X* are split contiguously over ranks and only the 16 B/point results are exchanged at the end;
* K-fold x restart fits: independent jobs dealt round-robin to ranks, results gathered as Python objects.

Works with any initialised ``torch.distributed`` backend (NCCL on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n_rows`` owned by ``rank``; sizes differ by at most one row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def sharded_predict(predict_fn: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray]], X_new: np.ndarray,
                    group=None, device: Optional[torch.device] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Every rank holds the same ``X_new`` and the same (replicated) emulator; rank r evaluates rows
    ``shard_bounds(M, world, r)`` with ``predict_fn`` (e.g. ``emulator.predict``) and all ranks receive the full
    (mean, std).  The result is bitwise identical to a single-rank ``predict_fn(X_new)`` because no reduction
    crosses a shard boundary."""
    world, rank = _world(group)
    X_new = np.asarray(X_new)
    M = X_new.shape[0]
    if world == 1:
        return predict_fn(X_new)
    lo, hi = shard_bounds(M, world, rank)
    mean, std = predict_fn(X_new[lo:hi])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    sizes = [shard_bounds(M, world, r) for r in range(world)]
    width = max(h - l for l, h in sizes)
    local = torch.zeros(2, width, dtype=torch.float64, device=device)
    local[0, : hi - lo] = torch.as_tensor(np.asarray(mean, dtype=np.float64), device=device)
    local[1, : hi - lo] = torch.as_tensor(np.asarray(std, dtype=np.float64), device=device)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local, group=group)
    out_mean = np.empty(M, dtype=np.float64)
    out_std = np.empty(M, dtype=np.float64)
    for (l, h), g in zip(sizes, gathered):
        g = g.cpu().numpy()
        out_mean[l:h] = g[0, : h - l]
        out_std[l:h] = g[1, : h - l]
    return out_mean, out_std


def run_jobs_round_robin(jobs: Sequence, run: Callable, group=None) -> List:
    """Deal ``jobs`` round-robin over ranks (job i -> rank i % world, as GPErks deals folds over its device
    list, GPErks/perks/cross_validation.py:101-103), run the local ones and return every job's result, in job
    order, on every rank (``all_gather_object``; payloads are state dicts / scores, a few KB)."""
    world, rank = _world(group)
    local = [(i, run(job)) for i, job in enumerate(jobs) if i % world == rank]
    if world == 1:
        return [r for _, r in local]
    bucket: List = [None] * world
    dist.all_gather_object(bucket, local, group=group)
    out: List = [None] * len(jobs)
    for part in bucket:
        for i, r in part:
            out[i] = r
    return out
