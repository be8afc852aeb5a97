"""``ExactGP`` base model with gpytorch's calling convention (gpytorch.models.ExactGP), backed by the
sm_100a engine.  Train mode returns the lazy prior over the training inputs; eval mode returns the exact
posterior predictive distribution.  Call sites: GPErks/gp/model.py:7-21, GPErks/train/emulator.py:326,
336, 356, 392."""
from __future__ import annotations

import warnings
import weakref
from typing import Optional

import torch

from .distributions import MultivariateNormal, PosteriorDistribution
from .engine import ExactGPEngine, make_theta
from .module import Module

F64 = torch.float64


class GPInputWarning(UserWarning):
    """Mirrors gpytorch.utils.warnings.GPInputWarning."""


class ExactGP(Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        if train_inputs is not None and torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        if train_inputs is not None:
            train_inputs = tuple(t.unsqueeze(-1) if t.dim() == 1 else t for t in train_inputs)
        self.train_inputs = train_inputs
        self.train_targets = train_targets
        self.likelihood = likelihood
        self._engine: Optional[ExactGPEngine] = None
        self._factor_theta: Optional[torch.Tensor] = None

    # -- pickling / device moves must not drag device workspaces along ---------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_factor_theta"] = None
        state["_pred_cache"] = None
        return state

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        if self.train_inputs is not None:
            self.train_inputs = tuple(fn(t) for t in self.train_inputs)
            self.train_targets = fn(self.train_targets)
        return out

    def set_train_data(self, inputs=None, targets=None, strict: bool = True):
        if inputs is not None:
            if torch.is_tensor(inputs):
                inputs = (inputs,)
            self.train_inputs = tuple(t.unsqueeze(-1) if t.dim() == 1 else t for t in inputs)
            self._engine = None
        if targets is not None:
            self.train_targets = targets

    def train(self, mode: bool = True):
        return super().train(mode)

    # -- engine plumbing ------------------------------------------------------------------------
    def _kernel_parts(self):
        return self.covar_module.scale_and_base()

    def _get_engine(self) -> ExactGPEngine:
        kind = self.covar_module.kind
        X = self.train_inputs[0]
        eng = self._engine
        if eng is None or eng.kind != kind or eng.N != X.shape[0] or eng.D != X.shape[1] \
                or eng.device.index != torch.cuda.current_device():
            eng = ExactGPEngine(X, kind)
            self._engine = eng
        return eng

    def _theta(self, device) -> torch.Tensor:
        s, base = self._kernel_parts()
        D = self.train_inputs[0].shape[1]
        return make_theta(base.lengthscale_vector(D), s, self.likelihood.noise, device)

    def _resid(self, device) -> torch.Tensor:
        # evaluated where the (tiny) mean parameters live, then moved once: avoids a host->device hop per parameter
        X = self.train_inputs[0]
        y = self.train_targets
        return (y.reshape(-1).to(F64) - self.mean_module(X).reshape(-1).to(F64)).to(device)

    def _factorized_engine(self) -> ExactGPEngine:
        """Engine whose cached factors match the current hyper-parameters (refactorises if stale)."""
        eng = self._get_engine()
        with torch.no_grad():
            theta = self._theta(eng.device)
            resid = self._resid(eng.device)
        if not eng.matches(theta, resid):
            eng.factorize(theta, resid)
            self._pred_cache = None
            self._pred_cache_mean = None
        return eng

    def _cached_prediction(self, x: torch.Tensor, noisy: bool, compute):
        """Per-epoch metrics call ``likelihood(model(X))`` once per metric on the same inputs: keep the last
        (mean, variance) keyed on the input tensor, its version and the factorisation they were computed from."""
        eng = self._factorized_engine()
        key = (x._version, tuple(x.shape), noisy, eng.n_factorizations)
        # ``noisy == "mean"``: the mean-only prediction (its own slot: a later request for the variance must not evict it mid-epoch)
        slot = "_pred_cache_mean" if noisy == "mean" else "_pred_cache"
        cache = getattr(self, slot, None)
        # the weak reference guards against a new tensor re-using the id of a collected one
        if cache is not None and cache[0] == key and cache[1]() is x:
            return cache[2]
        value = compute(eng)
        setattr(self, slot, (key, weakref.ref(x), value))
        return value

    # -- gpytorch calling convention ------------------------------------------------------------
    def __call__(self, *args, **kwargs):
        inputs = [a.unsqueeze(-1) if a.dim() == 1 else a for a in args]
        x = inputs[0]
        if self.training:
            if self.train_inputs is None:
                raise RuntimeError("train_inputs, train_targets cannot be None in training mode. "
                                   "Call .eval() for prior predictions, or call .set_train_data() to add training data.")
            tx = self.train_inputs[0]
            if x is not tx and not (x.shape == tx.shape and torch.equal(x.to(tx.device, tx.dtype), tx)):
                raise RuntimeError("You must train on the training inputs!")
            out = self.forward(*inputs, **kwargs)
            out._gp_model = self
            return out
        tx = self.train_inputs[0]
        if x.shape == tx.shape and torch.equal(x.to(tx.device, tx.dtype), tx):
            warnings.warn("The input matches the stored training data. Did you forget to call model.train()?",
                          GPInputWarning)
        return PosteriorDistribution(self, x, noisy=False)
