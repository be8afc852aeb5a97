import os
import random
from typing import Optional

import numpy
import torch
from scipy.stats import qmc


def set_seed(seed: Optional[int] = None):
    """Seed python / numpy / torch (GPErks/utils/random.py:10-20)."""
    if seed is None:
        return
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    numpy.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class RandomEngine(qmc.QMCEngine):
    """Simple random sampling behind scipy's QMCEngine interface (GPErks/utils/random.py:23-36)."""

    def __init__(self, d, seed=None):
        super().__init__(d=d, seed=seed)

    def _random(self, n=1, *, workers=1):
        return self.rng.random((n, self.d))

    def reset(self):
        super().__init__(d=self.d, seed=self.rng_seed)
        return self

    def fast_forward(self, n):
        self.random(n)
        return self
