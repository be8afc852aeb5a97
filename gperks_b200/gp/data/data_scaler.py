"""Input / output scalers (behaviour of GPErks/gp/data/data_scaler.py:34-96)."""
from abc import ABCMeta, abstractmethod

import numpy as np


class InputDataScaler(metaclass=ABCMeta):
    @abstractmethod
    def fit(self, X): ...

    @abstractmethod
    def transform(self, X): ...

    @abstractmethod
    def inverse_transform(self, X_): ...


class OutputDataScaler(metaclass=ABCMeta):
    @abstractmethod
    def fit(self, y): ...

    @abstractmethod
    def transform(self, y): ...

    @abstractmethod
    def inverse_transform(self, y_, ystd_=None): ...


class UnitCubeScaler(InputDataScaler):
    """Column-wise affine map of the training box onto [0, 1]^D."""

    def __init__(self):
        self.a = None
        self.b = None

    def fit(self, X):
        self.a, self.b = np.min(X, axis=0), np.max(X, axis=0)

    def transform(self, X):
        return (X - self.a) / (self.b - self.a)

    def inverse_transform(self, X_):
        return (self.b - self.a) * X_ + self.a


class StandardScaler(OutputDataScaler):
    """(y - mean) / population std; the inverse also rescales a predictive std."""

    def __init__(self):
        self.mu = None
        self.sigma = None

    def fit(self, y):
        self.mu, self.sigma = np.mean(y), np.std(y)

    def transform(self, y):
        return (y - self.mu) / self.sigma

    def inverse_transform(self, y_, ystd_=None):
        return self.sigma * y_ + self.mu, (None if ystd_ is None else self.sigma * ystd_)


class StandardLogScaler(OutputDataScaler):
    """Standardises log(y); the inverse returns log-normal mean and variance."""

    def __init__(self):
        self.mu = None
        self.sigma = None

    def fit(self, y):
        ly = np.log(y)
        self.mu, self.sigma = np.mean(ly), np.std(ly)

    def transform(self, y):
        return (np.log(y) - self.mu) / self.sigma

    def inverse_transform(self, y_, ystd_=None):
        if ystd_ is None:
            raise ValueError("StandardLogScaler cannot back-transform data without ystd_")
        m = self.sigma * y_ + self.mu
        s2 = (self.sigma * ystd_) ** 2
        return np.exp(m + 0.5 * s2), (np.exp(s2) - 1.0) * np.exp(2.0 * m + s2)
