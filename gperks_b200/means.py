"""Mean functions with gpytorch's names and parameter shapes (gpytorch.means.{ZeroMean,ConstantMean,
LinearMean}); GPErks' own polynomial ``LinearMean(degree, input_size)`` lives in gperks_b200/gp/mean.py.

Every mean here is a polynomial in the inputs, so ``poly_spec()`` lets the predict epilogue evaluate it
inside the CUDA kernel; arbitrary user means fall back to evaluating ``forward`` with torch on the device."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from .module import PARAM_DTYPE, Module


class Mean(Module):
    def forward(self, x: torch.Tensor) -> torch.Tensor:  # pragma: no cover
        raise NotImplementedError

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 1:
            x = x.unsqueeze(-1)
        return super().__call__(x)

    def poly_spec(self, input_size: int) -> Optional[Tuple[List[List[int]], Optional[torch.Tensor], Optional[torch.Tensor]]]:
        """(feature index lists, weights [P] or None, bias [1] or None) if this mean is a polynomial."""
        return None


class ZeroMean(Mean):
    def forward(self, x):
        return torch.zeros(x.shape[:-1], dtype=x.dtype, device=x.device)

    def poly_spec(self, input_size):
        return [], None, None


class ConstantMean(Mean):
    def __init__(self, constant_prior=None, constant_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.register_parameter("raw_constant", nn.Parameter(torch.zeros(batch_shape, dtype=PARAM_DTYPE)))

    @property
    def constant(self):
        return self.raw_constant

    @constant.setter
    def constant(self, value):
        self.initialize(raw_constant=value)

    def forward(self, x):
        return self.raw_constant.to(x.device, x.dtype).expand(x.shape[:-1])

    def poly_spec(self, input_size):
        return [], None, self.raw_constant.reshape(1)


class LinearMean(Mean):
    """gpytorch.means.LinearMean: m(x) = x @ weights + bias, weights [D,1], bias [1] (zeros-initialised...
    gpytorch uses randn for weights/bias; we follow that so seeded runs line up)."""

    def __init__(self, input_size: int, batch_shape=torch.Size(), bias: bool = True):
        super().__init__()
        self.input_size = input_size
        self.register_parameter("weights", nn.Parameter(torch.randn(*batch_shape, input_size, 1).to(PARAM_DTYPE)))
        if bias:
            self.register_parameter("bias", nn.Parameter(torch.randn(*batch_shape, 1).to(PARAM_DTYPE)))
        else:
            self.bias = None

    def forward(self, x):
        res = x.matmul(self.weights.to(x.device, x.dtype)).squeeze(-1)
        if self.bias is not None:
            res = res + self.bias.to(x.device, x.dtype)
        return res

    def poly_spec(self, input_size):
        return [[d] for d in range(input_size)], self.weights.reshape(-1), self.bias
