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

    def estimate_Sobol_indices_with_emulator_mean(self, emulator, block_points: int = 4_000_000, group=None):
        """Saltelli-2010 indices of the emulator's posterior mean; predictions and sums stay on the device.
        With an initialised ``torch.distributed`` process group (one process per GPU, the same emulator and seed on every
        rank) the design rows are split contiguously over the ranks and only the 8 B/point means are all-gathered; the
        sums are then formed redundantly on every rank, so the result does not depend on the number of GPUs."""
        import torch.distributed as dist
        from ..parallel import shard_bounds
        _, _, X = self.saltelli_design()
        n, d = self.n, self.d
        dev = emulator._compute_device()
        sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        world, rank = (dist.get_world_size(group), dist.get_rank(group)) if sharded else (1, 0)
        row_lo, row_hi = shard_bounds(X.shape[0], world, rank)
        with torch.no_grad(), torch.cuda.device(dev):
            f = torch.empty(X.shape[0], dtype=torch.float64, device=dev)
            for lo in range(row_lo, row_hi, block_points):
                hi = min(row_hi, lo + block_points)
                x_dev = torch.from_numpy(np.ascontiguousarray(X[lo:hi])).pin_memory().to(dev, non_blocking=True)
                emulator.model.eval()
                mu, _, linear = emulator.predict_device(x_dev)
                if not linear:
                    raise NotImplementedError("device-side GSA needs the (linear) StandardScaler on the outputs")
                f[lo:hi] = mu
            if sharded:
                bounds = [shard_bounds(X.shape[0], world, r) for r in range(world)]
                width = max(h - l for l, h in bounds)
                local = torch.zeros(width, dtype=torch.float64, device=dev)
                local[: row_hi - row_lo] = f[row_lo:row_hi]
                if dist.get_backend(group) != "nccl":
                    local = local.cpu()
                parts = [torch.empty_like(local) for _ in range(world)]
                dist.all_gather(parts, local, group=group)
                for (l, h), part in zip(bounds, parts):
                    f[l:h] = part[: h - l].to(dev)
            f_A, f_B = f[:n], f[n:2 * n]
            f_AB = f[2 * n:].reshape(n, d)                       # row j, column p: A_j with coordinate p from B_j
            mean = torch.cat([f_A, f_B]).mean()
            f_A, f_B, f_AB = f_A - mean, f_B - mean, f_AB - mean
            var = torch.cat([f_A, f_B]).var(unbiased=False)
            S1 = (f_B.unsqueeze(1) * (f_AB - f_A.unsqueeze(1))).mean(0) / var
            ST = 0.5 * ((f_A.unsqueeze(1) - f_AB) ** 2).mean(0) / var
            self.S1 = S1.cpu().numpy().reshape(1, -1)
            self.ST = ST.cpu().numpy().reshape(1, -1)

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
