from typing import Optional

from scipy.stats import qmc

from .random import RandomEngine


class Sampler:
    """Space-filling designs: 'srs', 'lhs' or scrambled 'sobol' (GPErks/utils/sampling.py:8-32)."""

    _ENGINES = {
        "srs": lambda d, seed: RandomEngine(d=d, seed=seed),
        "lhs": lambda d, seed: qmc.LatinHypercube(d=d, seed=seed),
        "sobol": lambda d, seed: qmc.Sobol(d=d, scramble=True, seed=seed),
    }

    def __init__(self, design: str = "lhs", dim: Optional[int] = None, seed: Optional[int] = None):
        if design not in self._ENGINES:
            raise ValueError("Not a valid sampling design! Choose between 'srs', 'lhs', 'sobol'")
        self.dim = dim
        self.design = design
        self.engine = self._ENGINES[design](dim, seed)

    def sample(self, n_samples, l_bounds=None, u_bounds=None):
        x = self.engine.random(n=n_samples)
        if l_bounds is not None and u_bounds is not None:
            x = qmc.scale(x, l_bounds, u_bounds)
        return x
