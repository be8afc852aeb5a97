"""Parameter constraints with gpytorch's semantics (gpytorch.constraints.Interval / GreaterThan /
Positive / LessThan): ``value = transform(raw)`` with a softplus for one-sided bounds and a sigmoid for
two-sided ones.  Bounds are registered buffers so they appear in ``state_dict`` under
``<param>_constraint.lower_bound`` / ``.upper_bound`` exactly as in the reference's snapshots
(notebooks/example_1.ipynb:556-575)."""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn.functional import softplus


def _inv_softplus(x: torch.Tensor) -> torch.Tensor:
    return x + torch.log(-torch.expm1(-x))


class Interval(nn.Module):
    enforced = True

    def __init__(self, lower_bound, upper_bound, transform=None, inv_transform=None, initial_value=None):
        super().__init__()
        lb = torch.as_tensor(lower_bound, dtype=torch.float64)
        ub = torch.as_tensor(upper_bound, dtype=torch.float64)
        if bool((lb >= ub).any()):
            raise ValueError("Got parameter bounds with empty intervals.")
        self.register_buffer("lower_bound", lb)
        self.register_buffer("upper_bound", ub)
        self.initial_value = initial_value

    def _bounds(self, like: torch.Tensor):
        return self.lower_bound.to(like.device, like.dtype), self.upper_bound.to(like.device, like.dtype)

    def transform(self, raw: torch.Tensor) -> torch.Tensor:
        lb, ub = self._bounds(raw)
        return torch.sigmoid(raw) * (ub - lb) + lb

    def inverse_transform(self, value: torch.Tensor) -> torch.Tensor:
        lb, ub = self._bounds(value)
        u = (value - lb) / (ub - lb)
        return torch.log(u) - torch.log1p(-u)

    def check(self, value: torch.Tensor) -> bool:
        lb, ub = self._bounds(value)
        return bool(torch.all(value <= ub) and torch.all(value >= lb))

    def __repr__(self):
        return f"{type(self).__name__}({float(self.lower_bound):.3E}, {float(self.upper_bound):.3E})"


class GreaterThan(Interval):
    def __init__(self, lower_bound, transform=None, inv_transform=None, initial_value=None):
        super().__init__(lower_bound, math.inf, initial_value=initial_value)

    def transform(self, raw):
        lb, _ = self._bounds(raw)
        return softplus(raw) + lb

    def inverse_transform(self, value):
        lb, _ = self._bounds(value)
        return _inv_softplus(value - lb)

    def __repr__(self):
        return f"{type(self).__name__}({float(self.lower_bound):.3E})"


class Positive(GreaterThan):
    def __init__(self, transform=None, inv_transform=None, initial_value=None):
        super().__init__(0.0, initial_value=initial_value)

    def __repr__(self):
        return "Positive()"


class LessThan(Interval):
    def __init__(self, upper_bound, transform=None, inv_transform=None, initial_value=None):
        super().__init__(-math.inf, upper_bound, initial_value=initial_value)

    def transform(self, raw):
        _, ub = self._bounds(raw)
        return ub - softplus(-raw)

    def inverse_transform(self, value):
        _, ub = self._bounds(value)
        return -_inv_softplus(ub - value)
