"""Stationary ARD kernels with gpytorch's constructor signatures, attribute names and state_dict keys
(``covar_module.raw_outputscale``, ``covar_module.base_kernel.raw_lengthscale`` -- SURVEY App. A.1).

The classes only hold parameters and describe the kernel (``kind``); the arithmetic is done by the
sm_100a kernels (csrc/gpx_kmat.cu).  Calling a kernel returns a lazy tensor whose ``to_dense()`` launches
the fused builder.  Replaces gpytorch.kernels.{ScaleKernel,RBFKernel,MaternKernel} at the call sites
``examples/example_1.py:44-56`` and ``GPErks/gp/experiment.py:118,150``."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _native as nv
from .constraints import Positive
from .module import PARAM_DTYPE, Module


class LazyKernelTensor:
    """k(x1, x2) not yet evaluated (stands in for linear_operator's LazyEvaluatedKernelTensor)."""

    def __init__(self, kernel: "Kernel", x1: torch.Tensor, x2: Optional[torch.Tensor] = None):
        self.kernel = kernel
        self.x1 = x1 if x1.dim() > 1 else x1.unsqueeze(-1)
        self.x2 = self.x1 if x2 is None else (x2 if x2.dim() > 1 else x2.unsqueeze(-1))
        self.symmetric = x2 is None

    @property
    def shape(self):
        return torch.Size([self.x1.shape[0], self.x2.shape[0]])

    def to_dense(self) -> torch.Tensor:
        return self.kernel.dense(self.x1, self.x2)

    evaluate = to_dense

    def diagonal(self, *args, **kwargs) -> torch.Tensor:
        s = self.kernel.scale_and_base()[0]
        return s.reshape(()).to(self.x1.device, self.x1.dtype).expand(self.x1.shape[0])

    diag = diagonal


class Kernel(Module):
    has_lengthscale = False
    kind: Optional[int] = None

    def __init__(self, ard_num_dims: Optional[int] = None, batch_shape=torch.Size([]), active_dims=None,
                 lengthscale_prior=None, lengthscale_constraint=None, eps: float = 1e-6, **kwargs):
        super().__init__()
        if len(batch_shape) != 0:
            raise NotImplementedError("batched kernels are outside the GPErks exact-GP path")
        if active_dims is not None:
            raise NotImplementedError("active_dims is outside the GPErks exact-GP path")
        self.ard_num_dims = ard_num_dims
        self.batch_shape = torch.Size(batch_shape)
        self.eps = eps
        if self.has_lengthscale:
            n = 1 if ard_num_dims is None else int(ard_num_dims)
            self.register_parameter("raw_lengthscale", nn.Parameter(torch.zeros(1, n, dtype=PARAM_DTYPE)))
            self.register_constraint("raw_lengthscale", lengthscale_constraint or Positive())

    @property
    def lengthscale(self) -> Optional[torch.Tensor]:
        return self._constrained("raw_lengthscale") if self.has_lengthscale else None

    @lengthscale.setter
    def lengthscale(self, value):
        self._set_constrained("raw_lengthscale", value)

    # -- description used by the engine ---------------------------------------------------------
    def scale_and_base(self):
        """(outputscale tensor, base kernel) -- a bare base kernel has outputscale 1."""
        ref = self.raw_lengthscale
        return torch.ones((), dtype=ref.dtype, device=ref.device), self

    def lengthscale_vector(self, D: int) -> torch.Tensor:
        ls = self.lengthscale.reshape(-1)
        if ls.numel() == 1 and D > 1:
            ls = ls.expand(D)
        if ls.numel() != D:
            raise RuntimeError(f"kernel has {ls.numel()} lengthscales but inputs have {D} columns")
        return ls

    # -- evaluation -----------------------------------------------------------------------------
    def forward(self, x1, x2=None, diag: bool = False, **params):
        lazy = LazyKernelTensor(self, x1, x2)
        return lazy.diagonal() if diag else lazy

    def dense(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        """Dense k(x1, x2) through gpx_kmat_cross (no noise); result on x1's device/dtype."""
        nv.require_cuda()
        lib = nv.lib()
        s, base = self.scale_and_base()
        dev = torch.device("cuda", torch.cuda.current_device())
        a = x1.detach().to(dev, torch.float64).contiguous()
        b = x2.detach().to(dev, torch.float64).contiguous()
        M, D = a.shape
        N = b.shape[0]
        ls = base.lengthscale_vector(D).detach().to(dev, torch.float64)
        theta = torch.cat([1.0 / ls, s.detach().reshape(1).to(dev, torch.float64),
                           torch.zeros(1, dtype=torch.float64, device=dev)])
        Np, Mc = nv.padded(N), nv.padded(M)
        Ks = torch.empty(Mc, Np, dtype=torch.float64, device=dev)
        mp = torch.empty(Np // nv.TILE, Mc, dtype=torch.float64, device=dev)
        zeros = torch.zeros(Np, dtype=torch.float64, device=dev)
        nv.check(lib.gpx_kmat_cross(nv.ptr(a), M, nv.ptr(b), N, D, nv.ptr(theta), None, None, base.kind,
                                    nv.ptr(zeros), nv.ptr(Ks), Np, nv.ptr(mp), Mc, nv.stream_ptr()), "gpx_kmat_cross")
        return Ks[:M, :N].to(x1.device, x1.dtype if x1.dtype.is_floating_point else torch.float64)


class RBFKernel(Kernel):
    has_lengthscale = True
    kind = nv.KIND_RBF


class MaternKernel(Kernel):
    has_lengthscale = True

    def __init__(self, nu: float = 2.5, **kwargs):
        if nu not in {0.5, 1.5, 2.5}:
            raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
        super().__init__(**kwargs)
        self.nu = nu

    @property
    def kind(self):
        return {0.5: nv.KIND_MATERN12, 1.5: nv.KIND_MATERN32, 2.5: nv.KIND_MATERN52}[self.nu]


class ScaleKernel(Kernel):
    """``outputscale * base_kernel`` (gpytorch.kernels.ScaleKernel)."""

    def __init__(self, base_kernel: Kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        super().__init__(**kwargs)
        if getattr(base_kernel, "kind", None) is None:
            raise NotImplementedError("ScaleKernel supports RBFKernel / MaternKernel base kernels")
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", nn.Parameter(torch.zeros((), dtype=PARAM_DTYPE)))
        self.register_constraint("raw_outputscale", outputscale_constraint or Positive())

    @property
    def outputscale(self) -> torch.Tensor:
        return self._constrained("raw_outputscale")

    @outputscale.setter
    def outputscale(self, value):
        self._set_constrained("raw_outputscale", value)

    @property
    def kind(self):
        return self.base_kernel.kind

    def scale_and_base(self):
        return self.outputscale, self.base_kernel
