"""Metric objects.  torchmetrics is not a dependency here: ``MeanSquaredError`` / ``R2Score`` are minimal
callables with torchmetrics' class names (GPErks dispatches on the class-name string:
GPErks/train/emulator.py:340, GPErks/perks/cross_validation.py:135)."""
import torch
from torch import Tensor


def get_metric_name(metric) -> str:
    return metric.__class__.__name__


class MeanSquaredError:
    def __call__(self, preds: Tensor, target: Tensor) -> Tensor:
        d = preds.double() - target.double()
        return (d * d).mean()


class MeanAbsoluteError:
    def __call__(self, preds: Tensor, target: Tensor) -> Tensor:
        return (preds.double() - target.double()).abs().mean()


class R2Score:
    def __call__(self, preds: Tensor, target: Tensor) -> Tensor:
        preds, target = preds.double(), target.double()
        ss_res = ((target - preds) ** 2).sum()
        ss_tot = ((target - target.mean()) ** 2).sum()
        return 1.0 - ss_res / ss_tot


class IndependentStandardError:
    """Fraction of points whose standardised error is below ``ci`` (GPErks/utils/metrics.py:11-22)."""

    def __init__(self, ci: float = 2.0):
        self.ci = ci
        self.total = 0.0

    def __call__(self, y_pred_mean: Tensor, y_pred_std: Tensor, y_true: Tensor) -> Tensor:
        ise = (y_true - y_pred_mean).abs() / y_pred_std
        return torch.as_tensor(float((ise < self.ci).sum()) / ise.numel(), dtype=torch.float64)
