"""Monomial features of degree 1..p: index multisets ordered by degree, then lexicographically
(same ordering as GPErks/utils/polynomialfeatures.py:12-42; D=2, p=2 -> x0, x1, x0^2, x0 x1, x1^2)."""
import itertools

import torch


class PolynomialFeatures:
    def __init__(self, degree: int):
        self.degree = degree
        self.indices = None
        self._gather = None          # [P, degree] index table into [x, 1]; shorter monomials are padded with the ones column
        self._input_size = None

    def fit(self, input_size: int):
        self.indices = [list(c) for k in range(1, self.degree + 1)
                        for c in itertools.combinations_with_replacement(range(input_size), k)]
        self._input_size = input_size
        pad = input_size
        self._gather = torch.tensor([idx + [pad] * (self.degree - len(idx)) for idx in self.indices], dtype=torch.long)
        return self

    def transform(self, x: torch.Tensor) -> torch.Tensor:
        if not self.indices:
            raise ValueError("self.indices is set to None. Did you forget to call 'fit'?")
        if self.degree == 1:
            return x                                   # the features of a degree-1 model are the inputs themselves
        # one gather + one product instead of P separate products (the reference stacks P `prod` results)
        ones = torch.ones(x.shape[:-1] + (1,), dtype=x.dtype, device=x.device)
        xe = torch.cat([x, ones], dim=-1)
        g = self._gather.to(x.device)
        return xe[..., g].prod(dim=-1)
