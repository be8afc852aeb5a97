"""Sobol' global sensitivity analysis with emulator uncertainty (API of GPErks/perks/gsa.py:21-148).

``estimate_Sobol_indices_with_emulator`` follows the reference: a Saltelli design of n(d+2) points from scipy's
Sobol' engine, posterior-predictive draws of the emulator at all of them, scipy's Saltelli-2010 estimators per
draw.  Two additions for B200-sized designs:
  * ``chunk``: joint draws within consecutive chunks of points (the reference's single joint draw needs the
    dense M x M covariance, impossible beyond M ~ 3e4; the reference itself already approximates that root with
    a rank-100 Lanczos decomposition above M = 800);
  * ``estimate_Sobol_indices_with_emulator_mean``: indices of the posterior mean over an arbitrarily large design
    (n = 2^19 ... 2^20), everything -- the n(d+2) predictions and the Saltelli sums -- staying on the GPU."""
from typing import Callable, Optional

import numpy as np
import pandas as pd
import torch
from scipy.stats import _sensitivity_analysis, sobol_indices, uniform

from ..constants import (
    DEFAULT_GSA_CONF_LEVEL,
    DEFAULT_GSA_N,
    DEFAULT_GSA_N_BOOTSTRAP,
    DEFAULT_GSA_N_DRAWS,
    DEFAULT_GSA_THRESHOLD,
)
from ..gp.data.dataset import Dataset
from ..utils.array import get_minmax


class SobolGSA:
    def __init__(self, dataset: Dataset, n: int = DEFAULT_GSA_N, seed: Optional[int] = None):
        self.n = n
        self.seed = seed
        self.d = dataset.input_size
        self.xlabels = dataset.x_labels
        self.ylabel = dataset.y_label
        if dataset.l_bounds is None and dataset.u_bounds is None:
            self.minmax = get_minmax(dataset.X_train)
        else:
            self.minmax = np.hstack((np.array(dataset.l_bounds).reshape(-1, 1), np.array(dataset.u_bounds).reshape(-1, 1)))
        self.ST = None
        self.S1 = None
        self.df = None

    def _dists(self):
        return [uniform(loc=lo, scale=hi - lo) for lo, hi in self.minmax]

    def saltelli_design(self):
        """(A, B, X) with X the stacked n(d+2) x d matrix  [A | B | AB pages interleaved]  (gsa.py:71-81)."""
        A, B = _sensitivity_analysis.sample_A_B(n=self.n, dists=self._dists(), rng=self.seed)
        AB = _sensitivity_analysis.sample_AB(A=A, B=B)
        d, _, n = AB.shape
        AB = np.moveaxis(AB, 0, -1).reshape(d, n * d)
        return A, B, np.hstack((A, B, AB)).T

    def estimate_Sobol_indices_with_simulator(self, f: Callable[[np.ndarray], np.ndarray]):
        indices = sobol_indices(func=f, n=self.n, dists=self._dists(), rng=self.seed)
        self.boot = indices.bootstrap(confidence_level=DEFAULT_GSA_CONF_LEVEL, n_resamples=DEFAULT_GSA_N_BOOTSTRAP)
        self.ST = indices.total_order.reshape(1, -1)
        self.S1 = indices.first_order.reshape(1, -1)

    def estimate_Sobol_indices_with_emulator(self, emulator, n_draws: int = DEFAULT_GSA_N_DRAWS,
                                             chunk: Optional[int] = None):
        _, _, X = self.saltelli_design()
        n, d = self.n, self.d
        f_all = emulator.sample(X, n_draws) if chunk is None else emulator.sample(X, n_draws, chunk=chunk)
        f_A, f_B, f_AB = f_all[:, :n], f_all[:, n:2 * n], f_all[:, 2 * n:]
        f_AB = np.moveaxis(f_AB.reshape((-1, n, d)), -1, 0)
        indices = sobol_indices(func={"f_A": f_A, "f_B": f_B, "f_AB": f_AB}, n=n, rng=self.seed)
        self.ST = np.atleast_2d(indices.total_order)
        self.S1 = np.atleast_2d(indices.first_order)

    # ---- device pipeline ---------------------------------------------------------------------
    def _base_design(self) -> np.ndarray:
        """The 2d-column scrambled Sobol' block scipy's ``sample_A_B`` draws (``qmc.Sobol(d=2d, seed, bits=64).random(n)``), as a
        C-contiguous [n, 2d] host array, cached per (n, seed).  A = U[:, :d] * scale + loc and B = U[:, d:] * scale + loc are
        exactly scipy's ``dist.ppf`` values for the uniform marginals this class uses -- formed on the device -- and the AB rows
        are assembled there too, so the n (d + 2) x d matrix of ``saltelli_design`` never exists on the host."""
        key = (self.n, self.seed)
        if getattr(self, "_design_key", None) != key:
            from scipy.stats import qmc
            self._design = np.ascontiguousarray(qmc.Sobol(d=2 * self.d, seed=self.seed, bits=64).random(self.n))
            self._design_key = key
        return self._design

    def estimate_Sobol_indices_with_emulator_device(self, emulator, n_draws: int = 0, draw_seed: Optional[int] = None,
                                                    block_rows: int = 65536, group=None):
        """Saltelli-2010 first-order / total indices with everything after the base Sobol' block on the GPU: A, B and the AB rows
        are formed on the device, all n (d + 2) design points are predicted there in blocks of ``block_rows`` design rows, and
        the sums scipy's estimators need are reduced by ``gpx_saltelli_partial`` / ``gpx_saltelli_reduce``; only (n_draws, d)
        indices come back (``self.S1``, ``self.ST``, like GPErks/perks/gsa.py:66-88).

        ``n_draws = 0``: indices of the posterior MEAN (one row); only the mean is evaluated (``gpx_predict_mean``).
        ``n_draws > 0``: per-draw indices from MARGINAL posterior-predictive draws, f_k(x) = mean(x) + std(x) z_k(x) with z
        independent per point (Philox keyed by ``draw_seed``).  The reference draws jointly over all points, which needs the
        dense n(d+2)-square covariance and is impossible at these sizes; marginal draws keep every point's predictive law but
        drop the correlation between points, so the spread of the indices over draws is a lower bound of the joint one.
        With an initialised ``torch.distributed`` group (one process per GPU, same emulator everywhere) rank 0's base block is
        broadcast, the chunks of design rows are dealt contiguously over the ranks, the per-chunk partial sums are gathered on
        rank 0 and added in chunk order: bitwise identical for any number of GPUs; every rank receives the indices."""
        import torch.distributed as dist
        from .. import _native as nv
        from ..parallel import shard_bounds
        lib = nv.lib()
        n, d = self.n, self.d
        dev = emulator._compute_device()
        sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        world, rank = (dist.get_world_size(group), dist.get_rank(group)) if sharded else (1, 0)
        F64 = torch.float64
        JC = int(lib.gpx_saltelli_chunk_rows())
        n_chunks = (n + JC - 1) // JC
        c_lo, c_hi = shard_bounds(n_chunks, world, rank)
        j_lo, j_hi = c_lo * JC, min(n, c_hi * JC)
        NS, kd = 4 + 3 * d, max(int(n_draws), 1)
        seed = int(self.seed if draw_seed is None and self.seed is not None else (draw_seed or 0)) & (2 ** 64 - 1)
        with torch.no_grad(), torch.cuda.device(dev):
            emulator.model.eval()
            emulator.model.likelihood.eval()
            if sharded and rank != 0:
                U = torch.empty(n, 2 * d, dtype=F64, device=dev)
            else:
                U = torch.from_numpy(self._base_design()).to(dev)
            if sharded:
                src = dist.get_global_rank(group, 0) if group is not None else 0
                if dist.get_backend(group) == "nccl":
                    dist.broadcast(U, src=src, group=group)
                else:
                    Uc = U.cpu()
                    dist.broadcast(Uc, src=src, group=group)
                    U = Uc.to(dev)
            loc = torch.as_tensor(np.asarray(self.minmax[:, 0], dtype=np.float64), device=dev)
            scale = torch.as_tensor(np.asarray(self.minmax[:, 1] - self.minmax[:, 0], dtype=np.float64), device=dev)
            m_loc = j_hi - j_lo
            fm = torch.empty(max(m_loc, 1) * (d + 2), dtype=F64, device=dev)            # [A | B | AB] of MY rows
            fs = torch.empty(max(m_loc, 1) * (d + 2), dtype=F64, device=dev) if n_draws > 0 else None
            for lo in range(j_lo, j_hi, block_rows):
                hi = min(j_hi, lo + block_rows)
                m = hi - lo
                A = U[lo:hi, :d] * scale + loc                       # uniform.ppf(u) = u * scale + loc, scipy's operation order
                B = U[lo:hi, d:] * scale + loc
                AB = A.repeat_interleave(d, dim=0).view(m, d, d)      # row (j, p): A_j with coordinate p from B_j
                AB.diagonal(dim1=1, dim2=2).copy_(B)
                X = torch.cat([A, B, AB.view(m * d, d)], 0).contiguous()
                o = lo - j_lo
                if n_draws > 0:
                    mu, sd, linear = emulator.predict_device(X)
                else:
                    (mu, linear), sd = emulator.predict_mean_device(X), None
                if not linear:
                    raise NotImplementedError("device-side GSA needs the (linear) StandardScaler on the outputs")
                for dst, src_t in ((fm, mu), (fs, sd)):
                    if dst is None:
                        continue
                    dst[o:o + m] = src_t[:m]
                    dst[m_loc + o: m_loc + o + m] = src_t[m:2 * m]
                    dst[2 * m_loc + o * d: 2 * m_loc + (o + m) * d] = src_t[2 * m:]
            my_chunks = c_hi - c_lo
            partial = torch.zeros(max(my_chunks, 1) * kd * NS, dtype=F64, device=dev)
            if my_chunks > 0:
                p = nv.ptr
                mA, mB, mAB = fm[:m_loc], fm[m_loc:2 * m_loc], fm[2 * m_loc:]
                sA = sB = sAB = None
                if fs is not None:
                    sA, sB, sAB = fs[:m_loc], fs[m_loc:2 * m_loc], fs[2 * m_loc:]
                # the kernel indexes rows globally (the Philox counter is the design row): pass pointers shifted back by j_lo
                sh = lambda t, w=1: (p(t) - 8 * j_lo * w) if t is not None else None  # noqa: E731
                nv.check(lib.gpx_saltelli_partial(sh(mA), sh(mB), sh(mAB, d), sh(sA), sh(sB), sh(sAB, d), n, d, j_lo, j_hi,
                                                  int(n_draws), seed, p(partial), nv.stream_ptr()), "gpx_saltelli_partial")
            if sharded:
                counts = [shard_bounds(n_chunks, world, r) for r in range(world)]
                width = max(max(h - l for l, h in counts), 1) * kd * NS
                block = torch.zeros(width, dtype=F64, device=dev)
                block[: my_chunks * kd * NS] = partial[: my_chunks * kd * NS]
                gloo = dist.get_backend(group) != "nccl"
                if gloo:
                    block = block.cpu()
                parts = [torch.empty_like(block) for _ in range(world)] if rank == 0 else None
                gdst = dist.get_global_rank(group, 0) if group is not None else 0
                dist.gather(block, parts, dst=gdst, group=group)
                if rank == 0:
                    partial = torch.cat([pt[: (h - l) * kd * NS] for (l, h), pt in zip(counts, parts)]).to(dev)
            out = torch.empty(2, kd, d, dtype=F64, device=dev)
            if rank == 0:
                sums = torch.empty(kd * NS, dtype=F64, device=dev)
                nv.check(lib.gpx_saltelli_reduce(nv.ptr(partial), n_chunks, d, int(n_draws), nv.ptr(sums), nv.stream_ptr()),
                         "gpx_saltelli_reduce")
                sums = sums.view(kd, NS)
                mean = (sums[:, 0] + sums[:, 1]) / (2 * n)
                var = (sums[:, 2] + sums[:, 3]) / (2 * n) - mean * mean
                per_p = sums[:, 4:].reshape(kd, d, 3)
                out[0] = (per_p[:, :, 0] - mean.unsqueeze(1) * per_p[:, :, 1]) / n / var.unsqueeze(1)      # first order
                out[1] = 0.5 * per_p[:, :, 2] / n / var.unsqueeze(1)                                        # total order
            if sharded:
                if dist.get_backend(group) == "nccl":
                    dist.broadcast(out, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
                else:
                    oc = out.cpu()
                    dist.broadcast(oc, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
                    out = oc
            res = out.cpu().numpy()
        self.S1, self.ST = res[0], res[1]

    def estimate_Sobol_indices_with_emulator_mean(self, emulator, block_points: int = 4_000_000, group=None):
        """Saltelli-2010 indices of the emulator's posterior mean over a design of any size (n = 2^19 ... 2^20): one row of
        indices; design rows, predictions and sums all stay on the device (``estimate_Sobol_indices_with_emulator_device`` with
        ``n_draws = 0``)."""
        rows = max(1024, int(block_points) // (self.d + 2))
        self.estimate_Sobol_indices_with_emulator_device(emulator, n_draws=0, block_rows=rows, group=group)

    def correct_Sobol_indices(self, threshold: float = DEFAULT_GSA_THRESHOLD):
        for S in (self.ST, self.S1):
            low = np.where(np.percentile(S, q=25, axis=0) < threshold)[0]
            S[:, low] = 0.0

    def assemble_dataframe(self):
        value = np.concatenate((self.ST.T.ravel(), self.S1.T.ravel()))
        param = np.repeat(self.xlabels, self.S1.shape[0])
        index = np.repeat(["ST", "S1"], len(param))
        self.df = pd.DataFrame.from_dict(dict(Parameter=np.concatenate((param, param)), Index=index, Value=value))

    def summary(self):
        self.assemble_dataframe()
        piv = pd.pivot_table(self.df, values="Value", index="Parameter", columns="Index", aggfunc="median")
        print(pd.DataFrame(np.round(piv["ST"], 6).values.reshape(-1, 1), index=piv.index, columns=["STi"]))
        print(pd.DataFrame(np.round(piv["S1"], 6).values.reshape(-1, 1), index=piv.index, columns=["Si"]))

    def plot(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the B200 hot path (matplotlib not required)")
