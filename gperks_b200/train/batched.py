"""All restarts of one fit trained at the same time on one GPU ("batched multi-restart fits
... as a batch of Cholesky factorisations on one GPU for small N").

For small training sets a restart is bound by the host, not the GPU: an Adam epoch at N=200 is ~15 kernel launches
plus a few dozen tiny torch ops, 2 ms of Python for <0.3 ms of kernel time.  Here every restart gets its own replica
of the model and its own ``ExactGPEngine``; one evaluation of a replica -- K_hat build, Cholesky, L^-1, alpha, -mll,
K_hat^-1, gradient reduction and the train/validation-set predictions the per-epoch metrics need -- is captured once
into a CUDA graph, and an epoch replays the R graphs on R streams (the tiny kernels of different restarts overlap on
the GPU), steps ONE optimizer that holds the parameters of all replicas, and reads everything the bookkeeping needs
(losses, `info`, predictions) back with a single device-to-host copy.

Semantics relative to the sequential loop (GPErks/train/emulator.py:96-112, 218-321):
* every restart starts from the raw hyper-parameter draw the sequential loop would give it (drawn up front, in order);
* each restart has its own TrainStats, early-stopping and snapshotting criterion (deep copies), stops on its own, and
  writes the same files; the best restart is selected by the same rule afterwards;
* the optimizer state is per parameter, so the restarts do not interact -- like a fresh optimizer per restart (the
  sequential loop carries its moment estimates from one restart into the next);
* -mll, its gradient and the predictions come from the same kernels, so one restart's trajectory is the one
  ``train(devices=...)`` produces for it.
"""
from __future__ import annotations

from copy import deepcopy
from typing import List, Optional

import torch

from .. import _native as nv
from ..engine import ExactGPEngine, NotPSDError
from ..log.logger import get_logger
from ..serialization.path import posix_path
from ..utils.metrics import get_metric_name
from .train_stats import TrainStats

log = get_logger()
F64 = torch.float64
MAX_BATCHED_N = 2048          # beyond this a restart keeps the GPU busy on its own
_FORCE_GENERIC = False        # tests: evaluate replica by replica even when the stacked fast path applies


class _BatchedNegMLL(torch.autograd.Function):
    """loss_r(theta_r, resid_r) = -mll of replica r, for the rows listed in ``rows``; one graph replay per row."""

    @staticmethod
    def forward(ctx, runner, rows, theta, resid):
        runner.evaluate(rows, theta.detach(), resid.detach())
        idx = torch.as_tensor(rows, device=theta.device)
        D = runner.D
        g = runner.grad_all.index_select(0, idx)                     # wrt (lengthscale, outputscale, noise)
        gt = g.clone()
        inv_ls = theta.detach()[:, :D]
        gt[:, :D] = -g[:, :D] / (inv_ls * inv_ls)                    # chain to 1/lengthscale
        ctx.save_for_backward(gt, runner.alpha_all.index_select(0, idx) / runner.N)
        return runner.loss_all.index_select(0, idx).clone()

    @staticmethod
    def backward(ctx, grad_out):
        gt, gr = ctx.saved_tensors
        return None, None, grad_out.unsqueeze(1) * gt, grad_out.unsqueeze(1) * gr


class _Runner:
    """R engines over the same training inputs with static input/output buffers and one CUDA graph each."""

    def __init__(self, X: torch.Tensor, kind: int, R: int, y: torch.Tensor, X_val: Optional[torch.Tensor], device):
        self.device = device
        self.R = R
        self.engines: List[ExactGPEngine] = [ExactGPEngine(X, kind, device) for _ in range(R)]
        e0 = self.engines[0]
        self.N, self.D = e0.N, e0.D
        f64 = dict(dtype=F64, device=device)
        self.y = y.detach().to(device, F64).reshape(-1).contiguous()
        self.X_val = None if X_val is None else X_val.detach().to(device, F64).contiguous()
        Mv = 0 if self.X_val is None else self.X_val.shape[0]
        self.theta_in = torch.zeros(R, self.D + 2, **f64)
        self.resid_in = torch.zeros(R, self.N, **f64)
        self.val_prior = torch.zeros(R, max(Mv, 1), **f64)          # mean_module_r(X_val), refreshed by the caller
        self.loss_all = torch.zeros(R, **f64)
        self.grad_all = torch.zeros(R, self.D + 2, **f64)
        self.alpha_all = torch.zeros(R, self.N, **f64)
        self.info_all = torch.zeros(R, dtype=torch.int32, device=device)
        self.pred_train = torch.zeros(R, 2, self.N, **f64)          # [r][0] mean, [r][1] noisy std on X_train
        self.pred_val = torch.zeros(R, 2, max(Mv, 1), **f64)
        self.streams = [torch.cuda.Stream(device=device) for _ in range(R)]
        self.ev_in = torch.cuda.Event()
        self.ev_done = [torch.cuda.Event() for _ in range(R)]
        self.graphs: List[Optional[torch.cuda.CUDAGraph]] = [None] * R
        self.use_graphs = True
        self._last = None                                            # (rows, theta, resid) of the last evaluation
        with torch.cuda.device(device):
            for r in range(R):                                      # warm-up: allocations, helper streams, attributes
                self.theta_in[r, : self.D] = 1.0
                self.theta_in[r, self.D:] = 1.0
                self._sequence(r)
            torch.cuda.synchronize(device)
            try:
                for r in range(R):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._sequence(r)
                    self.graphs[r] = g
            except Exception as exc:                                # pragma: no cover - depends on the driver
                log.warning(f"CUDA graph capture failed ({exc!r}); replaying the launches eagerly")
                self.use_graphs = False
                self.graphs = [None] * R
                torch.cuda.synchronize(device)

    def _sequence(self, r: int) -> None:
        """One replica's evaluation at (theta_in[r], resid_in[r]); only enqueues (no host read)."""
        eng = self.engines[r]
        eng.factorize(self.theta_in[r], self.resid_in[r], check=False)
        eng.gradients()
        self.loss_all[r].copy_(eng.out[0])
        self.grad_all[r].copy_(eng.grad)
        self.alpha_all[r].copy_(eng.alpha[: self.N])
        self.info_all[r].copy_(eng.info[0])
        prior = self.y - self.resid_in[r]                            # mean_module(X_train) = y - resid
        eng.predict(eng.X, prior_mean=prior, out_mean=self.pred_train[r, 0], out_std=self.pred_train[r, 1])
        if self.X_val is not None:
            eng.predict(self.X_val, prior_mean=self.val_prior[r], out_mean=self.pred_val[r, 0], out_std=self.pred_val[r, 1])

    def evaluate(self, rows: List[int], theta: torch.Tensor, resid: torch.Tensor) -> None:
        """Evaluate replicas ``rows`` at theta [len(rows), D+2], resid [len(rows), N] (device tensors)."""
        last = self._last
        if last is not None and last[0] == rows and torch.equal(last[1], theta) and torch.equal(last[2], resid):
            return
        with torch.cuda.device(self.device):
            idx = torch.as_tensor(rows, device=self.device)
            self.theta_in.index_copy_(0, idx, theta)
            self.resid_in.index_copy_(0, idx, resid)
            cur = torch.cuda.current_stream(self.device)
            self.ev_in.record(cur)
            for r in rows:
                s = self.streams[r]
                s.wait_event(self.ev_in)
                with torch.cuda.stream(s):
                    if self.graphs[r] is not None:
                        self.graphs[r].replay()
                    else:
                        self._sequence(r)
                    self.ev_done[r].record(s)
            for r in rows:
                cur.wait_event(self.ev_done[r])
            info = self.info_all.index_select(0, idx).cpu()          # the one host sync of an evaluation
            for k, r in enumerate(rows):
                eng = self.engines[r]
                if int(info[k]) != 0:                                # jitter schedule of psd_safe_cholesky, eagerly
                    eng.factorize(theta[k], resid[k])                # raises NotPSDError when it cannot be rescued
                    eng.gradients()
                    self.loss_all[r].copy_(eng.out[0])
                    self.grad_all[r].copy_(eng.grad)
                    self.alpha_all[r].copy_(eng.alpha[: self.N])
                    eng.predict(eng.X, prior_mean=self.y - resid[k], out_mean=self.pred_train[r, 0], out_std=self.pred_train[r, 1])
                    if self.X_val is not None:
                        eng.predict(self.X_val, prior_mean=self.val_prior[r], out_mean=self.pred_val[r, 0],
                                    out_std=self.pred_val[r, 1])
                eng._mark_factorized(has_kinv=True)                  # the graph bypassed the engine's Python-side bookkeeping
        self._last = (list(rows), theta.clone(), resid.clone())


class _StackedParams:
    """The parameters of all replicas as ONE leaf tensor per parameter name ([R, ...]), with every replica's own
    ``nn.Parameter`` re-pointed at its row (shared storage), so that (theta, resid) of all replicas come from a handful
    of vectorised torch ops and the optimizer steps five tensors instead of 5 R.  Understands the standard GPErks
    model: [ScaleKernel(] RBF/Matern kernel [)], Gaussian likelihood, and the polynomial means of this package;
    anything else makes ``build`` return None and the trainer evaluates replica by replica."""

    @staticmethod
    def build(replicas, X_train, y_train, X_val, device):
        from ..gp.mean import LinearMean as PolyLinearMean
        from ..means import ConstantMean, LinearMean, ZeroMean
        m0 = replicas[0]
        mm = m0.mean_module
        if not isinstance(mm, (PolyLinearMean, LinearMean, ConstantMean, ZeroMean)):
            return None
        names = [n for n, _ in m0.named_parameters()]
        known = {"mean_module.weights", "mean_module.bias", "mean_module.raw_constant"}
        ls_name = [n for n in names if n.endswith("raw_lengthscale")]
        os_name = [n for n in names if n.endswith("raw_outputscale")]
        nz_name = [n for n in names if n.endswith("raw_noise")]
        if len(ls_name) != 1 or len(nz_name) != 1 or len(os_name) > 1:
            return None
        if any(n not in known and n not in ls_name + os_name + nz_name for n in names):
            return None
        cov = m0.covar_module
        base = cov.base_kernel if hasattr(cov, "base_kernel") else cov
        if bool(os_name) != hasattr(cov, "base_kernel"):
            return None
        self = _StackedParams()
        self.R, self.device = len(replicas), device
        self.leaves = {}
        for n in names:
            ps = [dict(m.named_parameters())[n] for m in replicas]
            st = torch.stack([p.detach() for p in ps]).clone().requires_grad_(ps[0].requires_grad)
            for r, p_ in enumerate(ps):
                p_.data = st.data[r]                   # the replica's parameter now IS row r of the stacked tensor
            self.leaves[n] = st
        self.ls_name, self.nz_name = ls_name[0], nz_name[0]
        self.os_name = os_name[0] if os_name else None
        self.ls_c = base._modules.get("raw_lengthscale_constraint")
        self.os_c = cov._modules.get("raw_outputscale_constraint") if os_name else None
        self.nz_c = m0.likelihood.noise_covar._modules.get("raw_noise_constraint")
        self.D = X_train.shape[1] if X_train.dim() > 1 else 1
        X2 = X_train if X_train.dim() > 1 else X_train.unsqueeze(-1)
        self.y = y_train.reshape(-1).to(F64)
        self.mean_kind = ("poly" if isinstance(mm, PolyLinearMean) else "linear" if isinstance(mm, LinearMean)
                          else "const" if isinstance(mm, ConstantMean) else "zero")
        feats = (lambda x: mm._poly.transform(x)) if self.mean_kind == "poly" else (lambda x: x)
        self.phi_train = feats(X2).detach() if self.mean_kind in ("poly", "linear") else None
        self.n_train = X2.shape[0]
        self.phi_val = self.n_val = None
        if X_val is not None:
            Xv = X_val if X_val.dim() > 1 else X_val.unsqueeze(-1)
            self.n_val = Xv.shape[0]
            self.phi_val = feats(Xv).detach() if self.mean_kind in ("poly", "linear") else None
        self.has_bias = "mean_module.bias" in self.leaves
        return self

    def parameters(self):
        return list(self.leaves.values())

    @staticmethod
    def _tr(c, raw):
        return raw if c is None else c.transform(raw)

    def theta(self, idx: torch.Tensor) -> torch.Tensor:
        k = idx.numel()
        ls = self._tr(self.ls_c, self.leaves[self.ls_name].index_select(0, idx)).reshape(k, -1)
        if ls.shape[1] == 1 and self.D > 1:
            ls = ls.expand(k, self.D)
        nz = self._tr(self.nz_c, self.leaves[self.nz_name].index_select(0, idx)).reshape(k, 1)
        if self.os_name is not None:
            os_ = self._tr(self.os_c, self.leaves[self.os_name].index_select(0, idx)).reshape(k, 1)
        else:
            os_ = torch.ones(k, 1, dtype=ls.dtype, device=ls.device)
        return torch.cat([1.0 / ls, os_, nz], dim=1).to(self.device, F64)

    def _mean(self, idx: torch.Tensor, phi, n: int) -> torch.Tensor:
        k = idx.numel()
        if self.mean_kind in ("poly", "linear"):
            W = self.leaves["mean_module.weights"].index_select(0, idx)              # [k, P, 1]
            res = phi.matmul(W.to(phi.device, phi.dtype)).squeeze(-1)                  # [k, n]
            if self.has_bias:
                res = res + self.leaves["mean_module.bias"].index_select(0, idx).to(phi.device, phi.dtype)
            return res
        if self.mean_kind == "const":
            return self.leaves["mean_module.raw_constant"].index_select(0, idx).reshape(k, 1).expand(k, n)
        return torch.zeros(k, n, dtype=F64)

    def resid(self, idx: torch.Tensor) -> torch.Tensor:
        return (self.y.unsqueeze(0) - self._mean(idx, self.phi_train, self.n_train).to(F64)).to(self.device)

    def val_prior(self, idx: torch.Tensor) -> torch.Tensor:
        return self._mean(idx, self.phi_val, self.n_val).to(F64).to(self.device)


def unsupported_reason(emulator) -> Optional[str]:
    sd = emulator.scaled_data
    if sd.X_train.shape[0] > MAX_BATCHED_N:
        return f"more than {MAX_BATCHED_N} training points"
    return None


def train_restarts_batched(emulator, optimizer, early_stopping_criterion, snapshotting_criterion):
    """Returns (all_stats, best_epochs) like the sequential restart loop of ``GPEmulator.train``."""
    sd = emulator.scaled_data
    R = emulator.experiment.n_restarts
    dev = emulator._compute_device()
    X_train, y_train, X_val, y_val = sd.X_train, sd.y_train, sd.X_val, sd.y_val
    with_val = bool(sd.with_val)
    metric_names = list(map(get_metric_name, emulator.metrics))

    # ---- replicas: the sequential loop's initial state + its r-th raw hyper-parameter draw
    replicas, stats, escs, snpcs = [], [], [], []
    for r in range(R):
        raw_ls, raw_os, raw_noise = emulator._draw_restart_init()
        m = deepcopy(emulator.model)
        m.load_state_dict(emulator.init_state)
        m.initialize(**{"covar_module.base_kernel.raw_lengthscale": raw_ls, "covar_module.raw_outputscale": raw_os})
        if raw_noise is not None:
            m.likelihood.initialize(raw_noise=raw_noise)
        m.train()
        st = TrainStats(metric_names)
        esc, snpc = deepcopy(early_stopping_criterion), deepcopy(snapshotting_criterion)
        esc.enable(m, st)
        snpc.enable(m, st)
        replicas.append(m)
        stats.append(st)
        escs.append(esc)
        snpcs.append(snpc)
    stacked = None if _FORCE_GENERIC else _StackedParams.build(replicas, X_train, y_train, X_val if with_val else None, dev)
    params = stacked.parameters() if stacked is not None else [p for m in replicas for p in m.parameters()]
    opt = optimizer.__class__(params, **optimizer.defaults)
    frozen = {}            # row -> {name: value} of replicas that stopped (stacked mode: their rows still see the optimizer)

    with torch.cuda.device(dev):
        runner = _Runner(X_train, replicas[0].covar_module.kind, R, y_train, X_val if with_val else None, dev)
        for r, m in enumerate(replicas):
            m._engine = runner.engines[r]
            # the validation loss goes through the replica's own posterior: it must use the factors of the batched
            # evaluation as they are, not re-derive (theta, resid) and compare them bit by bit
            m._factorized_engine = (lambda e=runner.engines[r]: e)
        y_cpu = y_train.detach().cpu()
        yv_cpu = y_val.detach().cpu() if with_val else None
        active = list(range(R))
        best_epochs: List[Optional[int]] = [None] * R
        max_epochs = early_stopping_criterion.max_epochs
        width = len(str(max_epochs))

        def evaluate(rows):
            """-mll of the replicas in ``rows`` at their current parameters, differentiable; refreshes the predictions."""
            if stacked is not None:
                idx = torch.as_tensor(rows)
                thetas, resids = stacked.theta(idx), stacked.resid(idx)
                if with_val:
                    with torch.no_grad():
                        runner.val_prior.index_copy_(0, idx.to(dev), stacked.val_prior(idx))
                return _BatchedNegMLL.apply(runner, list(rows), thetas, resids)
            thetas = torch.stack([replicas[r]._theta(dev) for r in rows])
            resids = torch.stack([replicas[r]._resid(dev) for r in rows])
            if with_val:
                with torch.no_grad():
                    for r in rows:
                        runner.val_prior[r].copy_(replicas[r].mean_module(X_val).reshape(-1).to(dev, F64))
            return _BatchedNegMLL.apply(runner, list(rows), thetas, resids)

        loss = evaluate(active)
        while active:
            rows = list(active)
            train_losses = loss.detach().cpu().tolist()
            opt.zero_grad(set_to_none=True)
            loss.sum().backward()
            opt.step()
            if frozen:          # a stopped replica keeps the state its criterion left it in
                with torch.no_grad():
                    for r_, saved in frozen.items():
                        for n_, v_ in saved.items():
                            stacked.leaves[n_].data[r_].copy_(v_)
            msgs = {}
            for k, r in enumerate(rows):
                st = stats[r]
                st.current_epoch += 1
                st.train_loss.append(train_losses[k])
                msgs[r] = f"[restart {r + 1}] [{st.current_epoch:>{width}}/{max_epochs:>{width}}] Training Loss: {train_losses[k]:.4f}"
                snpcs[r].maybe_save(r + 1, st.current_epoch)
            # the metrics of an epoch describe the model AFTER its step: evaluate there once, differentiably, and reuse
            # that evaluation as the next epoch's training step
            loss = evaluate(rows)
            pred_tr = runner.pred_train.cpu()
            pred_va = runner.pred_val.cpu() if with_val else None
            still = []
            for k, r in enumerate(rows):
                m, st = replicas[r], stats[r]
                for metric in emulator.metrics:
                    name = get_metric_name(metric)
                    score = _score(metric, name, pred_tr[r], y_cpu)
                    msgs[r] += f" - {name}: {score:.4f}"
                    st.train_metrics_score[name].append(float(score))
                if with_val:
                    m.eval()
                    with torch.no_grad():
                        prev, emulator.model, emulator.criterion = (emulator.model, emulator.criterion), m, _criterion(m)
                        try:
                            val_loss = emulator._val_step(X_val, y_val)
                        finally:
                            emulator.model, emulator.criterion = prev
                    m.train()
                    st.val_loss.append(val_loss)
                    msgs[r] += f" | Validation Loss: {val_loss:.4f}"
                    for metric in emulator.metrics:
                        name = get_metric_name(metric)
                        score = _score(metric, name, pred_va[r], yv_cpu)
                        msgs[r] += f" - {name}: {score:.4f}"
                        st.val_metrics_score[name].append(float(score))
                log.info(msgs[r])
                best_epoch, best_model = escs[r].evaluate()
                if escs[r].is_verified:
                    snpcs[r].model = best_model
                    snpcs[r].save(r + 1, best_epoch)
                    st.best_epoch = best_epoch
                    best_epochs[r] = best_epoch
                    st.save_to_file(posix_path(snpcs[r].snapshot_dir.format(restart=r + 1), "train_stats.json"))
                    snpcs[r].keep_snapshots_until(r + 1, best_epoch)
                    if stacked is not None:
                        frozen[r] = {n_: v_.data[r].clone() for n_, v_ in stacked.leaves.items()}
                else:
                    still.append(r)
            if still and len(still) != len(rows):
                # a stopped replica may be rewound to its best state by its criterion: rebuild the autograd graph (and
                # the cached evaluation) over the remaining replicas only
                loss = evaluate(still)
            active = still
    return stats, best_epochs


def _criterion(model):
    from ..mlls import ExactMarginalLogLikelihood
    return ExactMarginalLogLikelihood(model.likelihood, model)


def _score(metric, name, pred, y):
    if name == "IndependentStandardError":
        return metric(pred[0], pred[1], y)
    return metric(pred[0], y)
