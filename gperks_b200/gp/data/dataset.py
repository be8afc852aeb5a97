"""Dataset container (train / optional validation / optional test) -- behaviour of
GPErks/gp/data/dataset.py without the plotting helpers (matplotlib is optional here)."""
import json
from functools import partial
from typing import Callable, List, Optional

import numpy as np
from scipy.stats import qmc

from ...utils.sampling import Sampler


class Dataset:
    def __init__(self, X_train: np.ndarray, y_train: np.ndarray, X_val: Optional[np.ndarray] = None,
                 y_val: Optional[np.ndarray] = None, X_test: Optional[np.ndarray] = None,
                 y_test: Optional[np.ndarray] = None, x_labels: Optional[List[str]] = None,
                 y_label: Optional[str] = None, l_bounds: Optional[List[float]] = None,
                 u_bounds: Optional[List[float]] = None, name: Optional[str] = None, descr: Optional[str] = None):
        self.X_train, self.y_train = X_train, y_train
        self.X_val, self.y_val = X_val, y_val
        self.X_test, self.y_test = X_test, y_test
        self.with_val = X_val is not None and y_val is not None
        self.with_test = X_test is not None and y_test is not None
        self.sample_size = X_train.shape[0]
        self.input_size = X_train.shape[1] if X_train.ndim > 1 else 1
        self.x_labels = list(x_labels) if x_labels else [f"X{i + 1}" for i in range(self.input_size)]
        self.y_label = y_label if y_label else "y"
        self.l_bounds = l_bounds if l_bounds else np.min(X_train, axis=0)
        self.u_bounds = u_bounds if u_bounds else np.max(X_train, axis=0)
        self.name = name if name else "TestExperiment"
        self.descr = descr if descr else "An example dataset to test GPErks' power!"

    @property
    def discrepancy(self):
        unit = qmc.scale(self.X_train, self.l_bounds, self.u_bounds, reverse=True)
        return qmc.discrepancy(unit)

    def summary(self):
        val = f"Yes (size = {self.X_val.shape[0]})" if self.with_val else "No"
        test = f"Yes (size = {self.X_test.shape[0]})" if self.with_test else "No"
        lines = [f"\n{self.name} dataset loaded."]
        if self.descr:
            lines.append(f'Notes from the author:\n"{self.descr}"')
        lines += ["Dataset properties:", f"-Input size: {self.input_size}", f"-Input parameters: {self.x_labels}",
                  "-Output size: 1", f"-Output feature: {[self.y_label]}", f"-Sample size: {self.sample_size}",
                  f"-Discrepancy: {self.discrepancy:.4f}", f"-Validation data available: {val}",
                  f"-Testing data available: {test}"]
        print("\n".join(lines))

    def plot(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the B200 hot path (matplotlib not required)")

    plot_train = plot_val = plot_test = plot_pairwise = plot

    @classmethod
    def build_from_file(cls, posix_path):
        """JSON layout of the reference: X_train / Y_train (columns = outputs) [+ _val, _test], optional
        x_labels, y_labels, l_bounds, u_bounds, info.  Returns {y_label: Dataset}."""
        with open(posix_path, "r") as f:
            d = json.load(f)
        X_train = np.array(d["X_train"])
        Y_train = np.array(d["Y_train"])
        if Y_train.ndim == 1:
            Y_train = Y_train[:, None]
        n_out = Y_train.shape[1]

        def opt(key):
            return np.array(d[key]) if key in d else None

        X_val, Y_val, X_test, Y_test = opt("X_val"), opt("Y_val"), opt("X_test"), opt("Y_test")
        x_labels = d.get("x_labels")
        y_labels = d.get("y_labels") or [f"y{i + 1}" for i in range(n_out)]
        name = str(posix_path).rsplit("/", 1)[-1].rsplit(".", 1)[0]

        def col(Y, i):
            if Y is None:
                return None
            return Y[:, i] if Y.ndim > 1 else Y

        return {
            y_labels[i]: cls(X_train, Y_train[:, i], X_val=X_val, y_val=col(Y_val, i), X_test=X_test,
                             y_test=col(Y_test, i), x_labels=x_labels, y_label=y_labels[i],
                             l_bounds=d.get("l_bounds"), u_bounds=d.get("u_bounds"),
                             name=f"{name}_{y_labels[i]}", descr=d.get("info"))
            for i in range(n_out)
        }

    @classmethod
    def build_from_function(cls, f: Callable[[np.ndarray], np.ndarray], d: int, n_train_samples: int,
                            n_val_samples: Optional[int] = None, n_test_samples: Optional[int] = None,
                            design: str = "lhs", seed: Optional[int] = None, x_labels=None, y_label=None,
                            l_bounds=None, u_bounds=None):
        """Train, validation and test sets are drawn one after the other from ONE sampling engine
        (GPErks/gp/data/dataset.py:234-249), so a seed reproduces all three."""
        name = f.func.__name__ if isinstance(f, partial) else f.__name__
        sampler = Sampler(design, d, seed)

        def draw(n):
            if n is None:
                return None, None
            X = sampler.sample(n, l_bounds, u_bounds)
            return X, np.squeeze(f(X.T))

        X_train, y_train = draw(n_train_samples)
        X_val, y_val = draw(n_val_samples)
        X_test, y_test = draw(n_test_samples)
        return cls(X_train, y_train, X_val=X_val, y_val=y_val, X_test=X_test, y_test=y_test, x_labels=x_labels,
                   y_label=y_label, l_bounds=l_bounds, u_bounds=u_bounds, name=name,
                   descr=f"This dataset was generated by using {name} function evaluations.")
