"""Scores of a trained emulator on the dataset's test split (API of GPErks/perks/inference.py:13-57)."""
import numpy as np
import pandas as pd

from ..utils.array import tensorize
from ..utils.metrics import get_metric_name


class Inference:
    def __init__(self, emulator):
        self.emulator = emulator
        self.X_test = emulator.experiment.dataset.X_test
        self.y_test = emulator.experiment.dataset.y_test
        self.metrics = emulator.experiment.metrics
        self.y_pred_mean, self.y_pred_std = emulator.predict(self.X_test)
        self.scores_dct = {}

    def summary(self, printtoconsole=True):
        self.scores_dct = {}
        mean, std, true = tensorize(self.y_pred_mean), tensorize(self.y_pred_std), tensorize(np.asarray(self.y_test))
        for metric in self.metrics:
            name = get_metric_name(metric)
            score = metric(mean, std, true) if name == "IndependentStandardError" else metric(mean, true)
            self.scores_dct[name] = score.cpu().numpy()
        if printtoconsole:
            print(pd.DataFrame(data=np.around(np.array(list(self.scores_dct.values())), decimals=4).reshape(-1, 1),
                               index=list(self.scores_dct.keys()), columns=["Score"]))

    def interpolate_2Dgrid(self, f=None, grid_dim: int = 50):
        """Predict on a regular grid over the 2-D training box; returns (X_grid, mean, std[, f(X_grid)])."""
        if self.X_test.shape[1] != 2:
            raise ValueError("Not a 2D input!")
        X_train = self.emulator.experiment.dataset.X_train
        g1 = np.linspace(X_train[:, 0].min(), X_train[:, 0].max(), grid_dim)
        g2 = np.linspace(X_train[:, 1].min(), X_train[:, 1].max(), grid_dim)
        x1, x2 = np.meshgrid(g1, g2)
        X_grid = np.hstack((x1.reshape(-1, 1), x2.reshape(-1, 1)))
        mean, std = self.emulator.predict(X_grid)
        out = (X_grid, mean.reshape(grid_dim, grid_dim), std.reshape(grid_dim, grid_dim))
        if f is not None:
            out += (np.asarray(f(X_grid.T)).reshape(grid_dim, grid_dim),)
        return out

    def plot(self):
        raise NotImplementedError("plotting is outside the B200 hot path (matplotlib not required)")
