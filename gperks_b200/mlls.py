"""ExactMarginalLogLikelihood (gpytorch.mlls.ExactMarginalLogLikelihood): log p(y | X) / N with analytic
gradients from the CUDA kernels, exposed to torch autograd so user-built optimizers work unchanged.
Call sites: GPErks/train/emulator.py:248-250, 326, 332 and 202-206."""
from __future__ import annotations

import torch

from .distributions import MultivariateNormal, PosteriorDistribution
from .kernels import LazyKernelTensor
from .module import Module

F64 = torch.float64


class _NegMLL(torch.autograd.Function):
    """loss(theta, resid) = -mll.  theta = (1/lengthscale_d, outputscale, noise) on the engine's device."""

    @staticmethod
    def forward(ctx, engine, theta, resid):
        if not engine.matches(theta, resid):
            engine.factorize(theta, resid)
        need = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        ctx.need = need
        if need:
            g = engine.gradients()                      # wrt (lengthscale, outputscale, noise)
            D = engine.D
            gt = g.clone()
            inv_ls = theta.detach()[:D]
            gt[:D] = -g[:D] / (inv_ls * inv_ls)          # chain to 1/lengthscale
            ctx.save_for_backward(gt, engine.alpha[: engine.N] / engine.N)
        return engine.out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        gt, gr = ctx.saved_tensors
        return None, grad_out * gt, grad_out * gr


class MarginalLogLikelihood(Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model


class ExactMarginalLogLikelihood(MarginalLogLikelihood):
    def forward(self, function_dist, target, *params, **kwargs):
        model = self.model
        if isinstance(function_dist, PosteriorDistribution):
            # eval mode: posterior-predictive log density of held-out targets (GPEmulator._val_step)
            noisy = function_dist.with_noise(self.likelihood)
            M = function_dist.x.shape[0]
            return noisy.log_prob(target) / M
        if not isinstance(function_dist, MultivariateNormal):
            raise RuntimeError("ExactMarginalLogLikelihood can only operate on Gaussian random variables")
        covar = function_dist.lazy_covariance_matrix
        tx = model.train_inputs[0]
        if not (isinstance(covar, LazyKernelTensor) and covar.symmetric and covar.x1.shape == tx.shape):
            raise NotImplementedError("the B200 exact-GP path evaluates the MLL on the model's training inputs")
        eng = model._get_engine()
        theta = model._theta(eng.device)
        resid = target.to(eng.device, F64).reshape(-1) - function_dist.mean.to(eng.device, F64).reshape(-1)
        loss = _NegMLL.apply(eng, theta, resid)
        res = -loss
        dt = target.dtype if target.dtype.is_floating_point else F64
        return res.to(target.device, dt)
