"""One-process-per-GPU helpers for the parts of the path that shard.

* predict / history matching / GSA evaluation: test points are independent given the replicated factors, so
  rows of X* are dealt contiguously over the ranks (``scatter`` from the rank that holds them) and only the
  16 B/point results come back (``gather`` to that rank) -- no rank but the source ever holds all of X*;
* K-fold x restart fits: independent jobs dealt round-robin to ranks, results gathered as Python objects.

Works with any initialised ``torch.distributed`` backend (NCCL on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.distributed as dist

F64 = torch.float64


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n_rows`` owned by ``rank``; sizes differ by at most one row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _global_rank(group, rank: int) -> int:
    return dist.get_global_rank(group, rank) if group is not None else rank


def _comm_device(group, device) -> torch.device:
    if device is not None:
        return torch.device(device)
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def scatter_rows(X_src: Optional[np.ndarray], group=None, device=None, src: int = 0) -> Tuple[torch.Tensor, int, int, int]:
    """Deal the rows of ``X_src`` (a host array on group rank ``src``; ignored elsewhere) contiguously over the ranks.
    Returns (my rows as a [hi-lo, D] tensor on the communication device, lo, hi, M).  One ``scatter`` of equal-sized
    (zero-padded) blocks; the shape travels first as a 2-int broadcast."""
    world, rank = _world(group)
    device = _comm_device(group, device)
    gsrc = _global_rank(group, src)
    shape = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        X_src = np.asarray(X_src, dtype=np.float64)
        if X_src.ndim == 1:
            X_src = X_src.reshape(-1, 1)
        shape = torch.tensor(X_src.shape, dtype=torch.int64, device=device)
    dist.broadcast(shape, src=gsrc, group=group)
    M, D = int(shape[0]), int(shape[1])
    bounds = [shard_bounds(M, world, r) for r in range(world)]
    width = max(max(h - l for l, h in bounds), 1)
    recv = torch.empty(width, D, dtype=F64, device=device)
    blocks = None
    if rank == src:
        padded = torch.zeros(world * width, D, dtype=F64, device=device)
        host = torch.from_numpy(np.ascontiguousarray(X_src))
        for r, (l, h) in enumerate(bounds):
            if h > l:
                padded[r * width: r * width + (h - l)].copy_(host[l:h], non_blocking=True)
        blocks = list(padded.view(world, width, D).unbind(0))
    dist.scatter(recv, blocks, src=gsrc, group=group)
    lo, hi = bounds[rank]
    return recv[: hi - lo], lo, hi, M


def gather_columns(local: torch.Tensor, lo: int, hi: int, M: int, group=None, dst: int = 0) -> Optional[np.ndarray]:
    """Inverse of ``scatter_rows`` for per-point results: ``local`` is [C, hi-lo] on the communication device; group rank
    ``dst`` receives the [C, M] host array, every other rank None.  One ``gather`` of equal-sized blocks."""
    world, rank = _world(group)
    gdst = _global_rank(group, dst)
    bounds = [shard_bounds(M, world, r) for r in range(world)]
    width = max(max(h - l for l, h in bounds), 1)
    C = local.shape[0]
    block = torch.zeros(C, width, dtype=local.dtype, device=local.device)
    block[:, : hi - lo] = local
    parts = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    dist.gather(block, parts, dst=gdst, group=group)
    if rank != dst:
        return None
    out = np.empty((C, M), dtype=np.float64)
    for (l, h), part in zip(bounds, parts):
        if h > l:
            out[:, l:h] = part[:, : h - l].cpu().numpy()
    return out


Target = Union[Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray]], object]


def sharded_predict(target: Target, X_new: Optional[np.ndarray] = None, group=None, device: Optional[torch.device] = None,
                    src: int = 0, replicated: bool = False, all_ranks: bool = False, **predict_kwargs):
    """``predict`` of one emulator over all ranks of an initialised ``torch.distributed`` group (one process per GPU, the same
    emulator -- same training data and hyper-parameters -- on every rank, so every rank builds bitwise equal factors).

    ``target``: a ``GPEmulator`` (rows stay on the GPU: NCCL scatter -> ``predict_device`` -> NCCL gather) or any callable
    ``f(X) -> (mean, std)`` on host arrays (used by the consumers and the gloo tests).
    ``X_new`` is read on group rank ``src`` only (pass None elsewhere) unless ``replicated=True`` (every rank already holds it:
    no scatter).  Returns ``(mean, std)`` on ``src`` and ``(None, None)`` elsewhere, or on every rank with ``all_ranks=True``.
    Bitwise identical to a single-rank ``predict`` of the same rows: no reduction crosses a shard boundary.
    Contract of the predict itself: GPErks/train/emulator.py:348-363."""
    world, rank = _world(group)
    is_emulator = hasattr(target, "predict_device")
    if world == 1:
        return target.predict(X_new, **predict_kwargs) if is_emulator else target(np.asarray(X_new))
    device = _comm_device(group, device)
    if replicated:
        X_new = np.asarray(X_new, dtype=np.float64)
        if X_new.ndim == 1:
            X_new = X_new.reshape(-1, 1)
        M = X_new.shape[0]
        lo, hi = shard_bounds(M, world, rank)
        x_local = torch.from_numpy(np.ascontiguousarray(X_new[lo:hi])).to(device)
    else:
        x_local, lo, hi, M = scatter_rows(X_new if rank == src else None, group=group, device=device, src=src)
    if is_emulator and device.type == "cuda":
        target.model.eval()
        mean_d, std_d, linear = target.predict_device(x_local.contiguous(), **predict_kwargs)
        if not linear:            # non-linear output scaler: finish on the host like the reference does
            m_h, s_h = target.scaled_data.scy.inverse_transform(mean_d.cpu().numpy(), ystd_=std_d.cpu().numpy())
            mean_d, std_d = torch.as_tensor(m_h, device=device), torch.as_tensor(s_h, device=device)
        local = torch.stack([mean_d, std_d])
    else:
        fn = (lambda x: target.predict(x, **predict_kwargs)) if is_emulator else target
        m_h, s_h = fn(x_local.cpu().numpy()) if hi > lo else (np.empty(0), np.empty(0))
        local = torch.stack([torch.as_tensor(np.asarray(m_h, dtype=np.float64)),
                             torch.as_tensor(np.asarray(s_h, dtype=np.float64))]).to(device)
    out = gather_columns(local, lo, hi, M, group=group, dst=src)
    if all_ranks:
        buf = torch.from_numpy(out).to(device) if rank == src else torch.empty(2, M, dtype=F64, device=device)
        dist.broadcast(buf, src=_global_rank(group, src), group=group)
        out = buf.cpu().numpy()
    if out is None:
        return None, None
    return out[0], out[1]


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
