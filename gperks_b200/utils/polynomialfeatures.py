"""Monomial features of degree 1..p: index multisets ordered by degree, then lexicographically
(same ordering as GPErks/utils/polynomialfeatures.py:12-42; D=2, p=2 -> x0, x1, x0^2, x0 x1, x1^2)."""
import itertools

import torch


class PolynomialFeatures:
    def __init__(self, degree: int):
        self.degree = degree
        self.indices = None

    def fit(self, input_size: int):
        self.indices = [list(c) for k in range(1, self.degree + 1)
                        for c in itertools.combinations_with_replacement(range(input_size), k)]
        return self

    def transform(self, x: torch.Tensor) -> torch.Tensor:
        if not self.indices:
            raise ValueError("self.indices is set to None. Did you forget to call 'fit'?")
        return torch.stack([x[..., idx].prod(dim=-1) for idx in self.indices], dim=-1)
