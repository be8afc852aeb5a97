"""Bayesian history matching waves (API of GPErks/perks/history_matching.py).

``Wave.compute_impl`` keeps the test points on the GPU: X is uploaded once, every emulator predicts from the same
device buffer, and the per-point implausibility (maxno-th largest over the emulators) is evaluated by a CUDA
epilogue (csrc/gpx_hm.cu) -- the reference loops over points in Python with two ``np.sort`` calls each
(history_matching.py:57-62) and copies every emulator's predictions to the host first."""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from .. import _native as nv
from ..log.logger import get_logger
from ..utils.array import get_minmax
from ..utils.indices import diff, part_and_select, whereq_whernot
from ..utils.jsonfiles import load_json, save_json

log = get_logger()
F64 = torch.float64


class Wave:
    def __init__(self, emulator=None, Itrain=None, cutoff=None, maxno=None, mean=None, var=None):
        self.emulator = emulator
        self.Itrain = Itrain
        self.cutoff = cutoff
        self.maxno = maxno
        self.mean = mean
        self.var = var
        self.I = None
        self.PV = None
        self.NIMP = None
        self.nimp_idx = None
        self.IMP = None
        self.imp_idx = None

    # ------------------------------------------------------------------------------------------
    def compute_impl(self, X, block_points: int = 4_000_000):
        """(I, PV) as numpy float64, one entry per row of X.  ``block_points`` bounds the device working set
        (E x block x 16 B for the per-emulator predictions)."""
        nv.require_cuda()
        lib = nv.lib()
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        n_samples, E = X.shape[0], len(self.emulator)
        if not (1 <= int(self.maxno) <= min(E, 8)):
            raise ValueError("maxno must be in [1, min(len(emulator), 8)]")
        dev = self.emulator[0]._compute_device()
        I = np.empty(n_samples, dtype=np.float64)
        PV = np.empty(n_samples, dtype=np.float64)
        with torch.no_grad(), torch.cuda.device(dev):
            z = torch.as_tensor(np.asarray(self.mean, dtype=np.float64).reshape(-1), device=dev)
            v = torch.as_tensor(np.asarray(self.var, dtype=np.float64).reshape(-1), device=dev)
            for lo in range(0, max(n_samples, 1), block_points):
                hi = min(n_samples, lo + block_points)
                m = hi - lo
                if m <= 0:
                    break
                x_dev = torch.from_numpy(X[lo:hi]).pin_memory().to(dev, non_blocking=True)
                means = torch.empty(E, m, dtype=F64, device=dev)
                stds = torch.empty(E, m, dtype=F64, device=dev)
                for j, emul in enumerate(self.emulator):
                    emul.model.eval()
                    mu, sd, linear = emul.predict_device(x_dev)
                    if not linear:       # non-linear output scaler: finish on the host like the reference
                        mu_h, sd_h = emul.scaled_data.scy.inverse_transform(mu.cpu().numpy(), ystd_=sd.cpu().numpy())
                        mu, sd = torch.as_tensor(mu_h, device=dev), torch.as_tensor(sd_h, device=dev)
                    means[j].copy_(mu)
                    stds[j].copy_(sd)
                Id = torch.empty(m, dtype=F64, device=dev)
                PVd = torch.empty(m, dtype=F64, device=dev)
                nv.check(lib.gpx_implausibility(nv.ptr(means), nv.ptr(stds), m, E, nv.ptr(z), nv.ptr(v), int(self.maxno),
                                                nv.ptr(Id), nv.ptr(PVd), nv.stream_ptr()), "gpx_implausibility")
                I[lo:hi] = Id.cpu().numpy()
                PV[lo:hi] = PVd.cpu().numpy()
        return I, PV

    def compute_impl_sharded(self, X, group=None, block_points: int = 4_000_000, src: int = 0, all_ranks: bool = False):
        """``compute_impl`` over all ranks of an initialised ``torch.distributed`` group (one process per GPU, every rank
        holding the same emulators): the rows of ``X`` -- read on group rank ``src`` only, pass None elsewhere -- are
        scattered contiguously over the ranks, rank r scores its slice with all emulators resident on its own GPU, so the
        per-point maximum over emulators stays local, and only (I, PV) -- 16 B/point -- are gathered back to ``src``
        (``all_ranks=True``: broadcast to everyone).  Bitwise identical to the single-GPU result.  """
        from ..parallel import sharded_predict
        return sharded_predict(lambda x: self.compute_impl(x, block_points), X, group=group, src=src, all_ranks=all_ranks)

    def find_regions(self, X):
        n_samples = X.shape[0]
        I, PV = self.compute_impl(X)
        keep = list(np.where(I < self.cutoff)[0])
        drop = diff(range(n_samples), keep)
        self.I, self.PV = I, PV
        self.nimp_idx, self.NIMP = keep, X[keep]
        self.imp_idx, self.IMP = drop, X[drop]

    def print_stats(self):
        nimp, imp = len(self.nimp_idx), len(self.imp_idx)
        tests = nimp + imp
        stats = pd.DataFrame(index=["TESTS", "IMP", "NIMP", "PERC"], columns=["#POINTS"],
                             data=[tests, imp, nimp, f"{100 * nimp / tests:.4f} %"])
        print(stats)

    def reconstruct_tests(self):
        n = self.NIMP.shape[0] + self.IMP.shape[0]
        X = np.zeros((n, self.NIMP.shape[1]), dtype=float)
        X[self.nimp_idx] = self.NIMP
        X[self.imp_idx] = self.IMP
        return X

    def save(self, filename):
        save_json({k: v for k, v in vars(self).items() if k != "emulator"}, filename)

    def load(self, filename):
        for k, v in load_json(filename).items():
            setattr(self, k, v)

    def get_nimps(self, n_points):
        if n_points >= len(self.nimp_idx) - 1:
            raise ValueError("Not enough NIMP points to choose from! n_points must be strictly less than "
                             "W.NIMP.shape[0] - 1.")
        _, X = part_and_select(self.NIMP, n_points)
        _, rest = whereq_whernot(self.NIMP, X)
        return X, self.NIMP[rest]

    def augment_nimp(self, n_total_points, scaling=0.1, n_max=2000):
        """'Cloud' resampling around the current NIMP points until ``n_total_points`` are available
        (GPErks/perks/history_matching.py:127-198); every candidate batch is screened on the GPU."""
        X = np.copy(self.NIMP)
        lb, ub = self.Itrain[:, 0].reshape(1, -1), self.Itrain[:, 1].reshape(1, -1)
        it = 0
        while X.shape[0] < n_total_points:
            it += 1
            bounds = get_minmax(X)
            scale = scaling * (bounds[:, 1] - bounds[:, 0])
            cand = np.random.normal(loc=X, scale=scale)
            for attempt in range(n_max + 1):
                outside = np.logical_or((cand <= lb).any(axis=1), (cand >= ub).any(axis=1))
                if not outside.any():
                    break
                if attempt == n_max:
                    cand = cand[~outside]
                    break
                cand[outside] = np.random.normal(loc=X[outside], scale=scale)
            I, _ = self.compute_impl(cand)
            X = np.vstack((X, cand[I < self.cutoff]))
            log.info(f"[Iteration: {it}] Found: {min(X.shape[0], n_total_points)} / {n_total_points}")
        nimp = len(self.nimp_idx)
        _, extra = part_and_select(X[nimp:], n_total_points - nimp)
        I, PV = self.compute_impl(extra)
        self.NIMP = np.vstack((X[:nimp], extra))
        self.I = np.concatenate((self.I, I))
        self.PV = np.concatenate((self.PV, PV))
        imp = len(self.imp_idx)
        self.nimp_idx = list(self.nimp_idx) + list(range(nimp + imp, imp + n_total_points))

    def get_trains(self, X, n_points):
        if n_points > X.shape[0]:
            raise ValueError("Cannot return more points than totally available points! "
                             "Choose n_points <= X_train.shape[0].")
        if n_points == X.shape[0]:
            return X
        I, _ = self.compute_impl(X)
        return X[np.argsort(I)[:n_points]]

    def copy(self):
        W = type(self)(emulator=self.emulator, Itrain=self.Itrain, cutoff=self.cutoff, maxno=self.maxno, mean=self.mean,
                 var=self.var)
        W.I, W.PV = np.copy(self.I), np.copy(self.PV)
        W.NIMP, W.IMP = np.copy(self.NIMP), np.copy(self.IMP)
        W.nimp_idx, W.imp_idx = list(self.nimp_idx), list(self.imp_idx)
        return W

    def plot_wave(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the B200 hot path (matplotlib not required)")
