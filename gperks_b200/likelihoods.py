"""GaussianLikelihood with gpytorch's structure: ``likelihood.noise_covar.raw_noise`` [1] under a
``GreaterThan(1e-4)`` constraint (state_dict keys per SURVEY App. A.1); ``likelihood(dist)`` returns the
noisy predictive distribution.  Call sites: GPErks/train/emulator.py:54-58, 243-244, 336, 356, 392."""
from __future__ import annotations

import torch
from torch import nn

from .constraints import GreaterThan
from .module import PARAM_DTYPE, Module


class HomoskedasticNoise(Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), num_tasks: int = 1):
        super().__init__()
        self.register_parameter("raw_noise", nn.Parameter(torch.zeros(*batch_shape, num_tasks, dtype=PARAM_DTYPE)))
        self.register_constraint("raw_noise", noise_constraint or GreaterThan(1e-4))

    @property
    def noise(self) -> torch.Tensor:
        return self._constrained("raw_noise")

    @noise.setter
    def noise(self, value):
        self._set_constrained("raw_noise", value)


class Likelihood(Module):
    pass


class GaussianLikelihood(Likelihood):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.noise_covar = HomoskedasticNoise(noise_prior=noise_prior, noise_constraint=noise_constraint,
                                              batch_shape=batch_shape)

    @property
    def noise(self) -> torch.Tensor:
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def raw_noise(self) -> torch.Tensor:
        return self.noise_covar.raw_noise

    @raw_noise.setter
    def raw_noise(self, value):
        self.noise_covar.initialize(raw_noise=value)

    def forward(self, function_dist, *args, **kwargs):
        return function_dist.with_noise(self)

    def __call__(self, function_dist, *args, **kwargs):
        return self.forward(function_dist, *args, **kwargs)
