import torch

from ..distributions import MultivariateNormal
from ..models import ExactGP


class ExactGPModel(ExactGP):
    """GPErks/gp/model.py:7-21 -- an exact GP made of a mean module and a covariance module."""

    def __init__(self, X_train: torch.Tensor, y_train: torch.Tensor, likelihood, mean_module, covar_module):
        super().__init__(X_train, y_train, likelihood)
        self.mean_module = mean_module
        self.covar_module = covar_module

    def forward(self, x):
        return MultivariateNormal(self.mean_module(x), self.covar_module(x))
