"""Context managers with gpytorch.settings' names so that reference code such as
``with torch.no_grad(), gpytorch.settings.fast_pred_var():`` (GPErks/train/emulator.py:281,354,391) ports unchanged.

They are deliberate no-ops: this implementation always evaluates the exact Cholesky expressions those settings
approximate (fast_pred_var's LOVE/Lanczos cache, max_cholesky_size's CG switch, ...), at every problem size."""
from contextlib import ContextDecorator


class _Flag(ContextDecorator):
    def __init__(self, state=True, *args, **kwargs):
        self.state = state

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    @classmethod
    def on(cls):
        return True

    @classmethod
    def off(cls):
        return False


class fast_pred_var(_Flag):
    """Exact predictive variances are already faster than LOVE here."""


class fast_pred_samples(_Flag):
    pass


class fast_computations(_Flag):
    def __init__(self, covar_root_decomposition=True, log_prob=True, solves=True):
        super().__init__(True)


class max_cholesky_size(_Flag):
    """gpytorch leaves the Cholesky path above this size (default 800); the CUDA path never does."""

    def __init__(self, value=800):
        super().__init__(value)


class cholesky_jitter(_Flag):
    def __init__(self, float_value=None, double_value=None, half_value=None):
        super().__init__(double_value)


class debug(_Flag):
    pass


class min_variance(_Flag):
    @staticmethod
    def value(dtype=None):
        return 1e-10
