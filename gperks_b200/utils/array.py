import numpy as np
import torch


def get_minmax(X):
    """Per-column [min, max] pairs, shape (D, 2)  (GPErks/utils/array.py:5-9)."""
    X = np.asarray(X)
    return np.stack([X.min(axis=0), X.max(axis=0)], axis=1)


def tensorize(X, dtype=torch.float64):
    """numpy -> torch.  The reference casts to float32 (GPErks/utils/array.py:12-13); the B200 path
    keeps float64 end to end (pass ``dtype=torch.float32`` to reproduce the reference's rounding)."""
    return torch.from_numpy(np.ascontiguousarray(X)).to(dtype)
