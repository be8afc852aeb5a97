"""Bayesian history matching waves (API of GPErks/perks/history_matching.py).

``Wave.compute_impl`` keeps the test points on the GPU: X is uploaded once, every emulator predicts from the same
device buffer, and the per-point implausibility (maxno-th largest over the emulators) is evaluated by a CUDA
epilogue (csrc/gpx_hm.cu) -- the reference loops over points in Python with two ``np.sort`` calls each
(history_matching.py:57-62) and copies every emulator's predictions to the host first."""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import torch

from .. import _native as nv
from ..log.logger import get_logger
from ..utils.array import get_minmax
from ..utils.indices import part_and_select, whereq_whernot
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
    def _staging(self, dev, blk: int, D: int):
        """Two pinned host slots + device buffers for blocks of ``blk`` points, allocated once per Wave and reused by every
        ``compute_impl`` call (pinning memory costs more than the copies it serves)."""
        st = getattr(self, "_stage", None)
        if st is None or st["device"] != dev or st["blk"] < blk or st["D"] != D:
            st = {
                "device": dev, "blk": blk, "D": D, "copy": torch.cuda.Stream(device=dev), "mma": torch.cuda.Stream(device=dev, priority=-1),
                "xin": [torch.empty(blk * D, dtype=F64).pin_memory() for _ in range(2)],
                "xdev": [torch.empty(blk * D, dtype=F64, device=dev) for _ in range(2)],
                "odev": [torch.empty(2, blk, dtype=F64, device=dev) for _ in range(2)],           # I, PV
                "out": [torch.empty(2, blk, dtype=F64).pin_memory() for _ in range(2)],
                "idx": [torch.empty(blk, dtype=torch.int64, device=dev) for _ in range(2)],        # kept indices (block-local)
                "cnt": [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(2)],
                "cnt_host": [torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(2)],
                "ev_h2d": [torch.cuda.Event() for _ in range(2)], "ev_done": [torch.cuda.Event() for _ in range(2)],
                "ev_d2h": [torch.cuda.Event() for _ in range(2)],
                "compact_ws": torch.empty(int(nv.lib().gpx_compact_below_workspace_bytes(blk)), dtype=torch.uint8, device=dev),
                "ws": None,
            }
            self._stage = st
        return st

    def _wave_descriptors(self, dev):
        """gpx_i8_emulator array for the one-launch-sequence path (every emulator on a sliced-integer mode, on ``dev``, with a
        polynomial prior mean and the linear output scaler), else None -> the per-emulator loop."""
        descs, keep = [], []
        for emul in self.emulator:
            if emul._compute_device() != dev or emul.scaled_data.input_size != self.emulator[0].scaled_data.input_size:
                return None, None
            d = emul.i8_descriptor()
            if d is None:
                return None, None
            descs.append(d[0])
            keep.append(d[1])
        return (nv.I8EmulatorDesc * len(descs))(*descs), keep

    def compute_impl(self, X, block_points: int = 4_000_000):
        """(I, PV) as numpy float64, one entry per row of X  (GPErks/perks/history_matching.py:43-64)."""
        I, PV, _ = self._compute(X, block_points, want_keep=False)
        return I, PV

    def _compute(self, X, block_points: int, want_keep: bool):
        """Blocks of ``block_points`` rows stream through two pinned staging slots (H2D of block b+1 and D2H of block b-1 on a
        copy stream while block b computes).  Per block all emulators predict from the same device-resident rows: with every
        emulator on a sliced-integer mode through ONE call of ``gpx_wave_i8`` (one two-stream chunk pipeline over (chunk,
        emulator) items + the implausibility epilogue; per-emulator predictions never leave the chip), otherwise emulator by
        emulator + ``gpx_implausibility``.  ``want_keep``: also the ascending indices with I < cutoff, compacted on the device
        (``gpx_compact_below``) so that only the kept indices cross PCIe."""
        nv.require_cuda()
        lib = nv.lib()
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        n_samples, D = X.shape
        E = len(self.emulator)
        if not (1 <= int(self.maxno) <= min(E, 8)):
            raise ValueError("maxno must be in [1, min(len(emulator), 8)]")
        dev = self.emulator[0]._compute_device()
        I = np.empty(n_samples, dtype=np.float64)
        PV = np.empty(n_samples, dtype=np.float64)
        keep_parts = []
        if n_samples == 0:
            return I, PV, np.empty(0, dtype=np.int64)
        with torch.no_grad(), torch.cuda.device(dev):
            z = torch.as_tensor(np.asarray(self.mean, dtype=np.float64).reshape(-1), device=dev)
            v = torch.as_tensor(np.asarray(self.var, dtype=np.float64).reshape(-1), device=dev)
            descs, keepalive = self._wave_descriptors(dev)
            blk = max(1, min(n_samples, int(block_points)))
            st = self._staging(dev, blk, D)
            comp, copy = torch.cuda.current_stream(), st["copy"]
            Mc = 32768
            if descs is not None:
                need = int(lib.gpx_wave_i8_workspace_bytes(descs, E, Mc))
                if st["ws"] is None or st["ws"].numel() < need:
                    st["ws"] = None
                    st["ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
            pending = {}

            def flush(slot):
                lo_, hi_ = pending.pop(slot)
                st["ev_d2h"][slot].synchronize()
                m_ = hi_ - lo_
                I[lo_:hi_] = st["out"][slot][0, :m_].numpy()
                PV[lo_:hi_] = st["out"][slot][1, :m_].numpy()
                if want_keep:
                    k = int(st["cnt_host"][slot].item())
                    keep_parts.append(st["idx"][slot][:k].cpu().numpy() + lo_)

            for b, lo in enumerate(range(0, n_samples, blk)):
                i = b % 2
                hi = min(n_samples, lo + blk)
                m = hi - lo
                if i in pending:
                    flush(i)
                st["xin"][i][: m * D].copy_(torch.from_numpy(X[lo:hi]).reshape(-1))
                with torch.cuda.stream(copy):
                    st["xdev"][i][: m * D].copy_(st["xin"][i][: m * D], non_blocking=True)
                    st["ev_h2d"][i].record(copy)
                comp.wait_event(st["ev_h2d"][i])
                x_dev = st["xdev"][i][: m * D].view(m, D)
                Id, PVd = st["odev"][i][0, :m], st["odev"][i][1, :m]
                if descs is not None:
                    nv.check(lib.gpx_wave_i8(descs, E, D, nv.ptr(x_dev), m, nv.ptr(z), nv.ptr(v), int(self.maxno), nv.ptr(Id), nv.ptr(PVd),
                                             None, None, nv.ptr(st["ws"]), st["ws"].numel(), Mc, nv.stream_ptr(),
                                             st["mma"].cuda_stream if os.environ.get("GPX_I8_OVERLAP", "1") != "0" else None), "gpx_wave_i8")
                else:
                    means = torch.empty(E, m, dtype=F64, device=dev)
                    stds = torch.empty(E, m, dtype=F64, device=dev)
                    for j, emul in enumerate(self.emulator):
                        emul.model.eval()
                        mu, sd, linear = emul.predict_device(x_dev, out_mean=means[j], out_std=stds[j])
                        if not linear:       # non-linear output scaler: finish on the host like the reference
                            mu_h, sd_h = emul.scaled_data.scy.inverse_transform(mu.cpu().numpy(), ystd_=sd.cpu().numpy())
                            means[j].copy_(torch.as_tensor(mu_h, device=dev))
                            stds[j].copy_(torch.as_tensor(sd_h, device=dev))
                    nv.check(lib.gpx_implausibility(nv.ptr(means), nv.ptr(stds), m, E, nv.ptr(z), nv.ptr(v), int(self.maxno),
                                                    nv.ptr(Id), nv.ptr(PVd), nv.stream_ptr()), "gpx_implausibility")
                if want_keep:
                    nv.check(lib.gpx_compact_below(nv.ptr(Id), m, float(self.cutoff), nv.ptr(st["idx"][i]), nv.ptr(st["cnt"][i]),
                                                   nv.ptr(st["compact_ws"]), st["compact_ws"].numel(), nv.stream_ptr()), "gpx_compact_below")
                st["ev_done"][i].record(comp)
                with torch.cuda.stream(copy):
                    copy.wait_event(st["ev_done"][i])
                    st["out"][i][:, :m].copy_(st["odev"][i][:, :m], non_blocking=True)
                    if want_keep:
                        st["cnt_host"][i].copy_(st["cnt"][i], non_blocking=True)
                    st["ev_d2h"][i].record(copy)
                pending[i] = (lo, hi)
            for slot in sorted(pending, key=lambda k: pending[k][0]):
                flush(slot)
            del keepalive
        keep = np.concatenate(keep_parts) if keep_parts else np.empty(0, dtype=np.int64)
        return I, PV, keep

    def compute_impl_sharded(self, X, group=None, block_points: int = 4_000_000, src: int = 0, all_ranks: bool = False):
        """``compute_impl`` over all ranks of an initialised ``torch.distributed`` group (one process per GPU, every rank
        holding the same emulators): the rows of ``X`` -- read on group rank ``src`` only, pass None elsewhere -- are
        scattered contiguously over the ranks, rank r scores its slice with all emulators resident on its own GPU, so the
        per-point maximum over emulators stays local, and only (I, PV) -- 16 B/point -- are gathered back to ``src``
        (``all_ranks=True``: broadcast to everyone).  Bitwise identical to the single-GPU result.  """
        from ..parallel import sharded_predict
        return sharded_predict(lambda x: self.compute_impl(x, block_points), X, group=group, src=src, all_ranks=all_ranks)

    def find_regions(self, X, block_points: int = 4_000_000):
        """NIMP / IMP split of X (GPErks/perks/history_matching.py:66-77); the kept indices come from the device compaction."""
        n_samples = X.shape[0]
        if type(self).compute_impl is Wave.compute_impl:
            I, PV, keep = self._compute(X, block_points, want_keep=True)
        else:                                   # a subclass supplies its own implausibility: the reference's host-side selection
            I, PV = self.compute_impl(X)
            keep = np.where(I < self.cutoff)[0]
        mask = np.ones(n_samples, dtype=bool)
        mask[keep] = False
        keep, drop = keep.tolist(), np.nonzero(mask)[0].tolist()
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
