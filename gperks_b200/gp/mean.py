"""GPErks' polynomial mean ``LinearMean(degree, input_size, bias=True)``: m(x) = phi_degree(x) @ weights + bias
with ``weights`` [P,1] and ``bias`` [1] drawn from N(0,1) under the experiment seed
(GPErks/gp/mean.py:9-36)."""
import torch
from torch import nn

from ..means import Mean
from ..module import PARAM_DTYPE
from ..utils.polynomialfeatures import PolynomialFeatures


class LinearMean(Mean):
    def __init__(self, degree, input_size, batch_shape=torch.Size(), bias=True):
        super().__init__()
        self.degree = degree
        self.input_size = input_size
        self._poly = PolynomialFeatures(degree).fit(input_size)
        n_weights = len(self._poly.indices)
        # drawn in the default dtype so that seeded initial values coincide with the reference's
        self.register_parameter("weights", nn.Parameter(torch.randn(*batch_shape, n_weights, 1).to(PARAM_DTYPE)))
        if bias:
            self.register_parameter("bias", nn.Parameter(torch.randn(*batch_shape, 1).to(PARAM_DTYPE)))
        else:
            self.bias = None

    def forward(self, x):
        res = self._poly.transform(x).matmul(self.weights.to(x.device, x.dtype)).squeeze(-1)
        if self.bias is not None:
            res = res + self.bias.to(x.device, x.dtype)
        return res

    def poly_spec(self, input_size):
        return self._poly.indices, self.weights.reshape(-1), self.bias
