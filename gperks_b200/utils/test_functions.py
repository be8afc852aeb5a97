"""Analytic benchmark functions; inputs are arrays of shape (d, n) (one row per coordinate)."""
import numpy as np


def forrester(x):
    """1-D, x in [0, 1]: (6x - 2)^2 sin(12x - 4)."""
    return (6.0 * x - 2.0) ** 2 * np.sin(12.0 * x - 4.0)


def currin_exp(x):
    """2-D, x in [0, 1]^2 (Currin et al. exponential function)."""
    x0, x1 = x[0], x[1]
    num = 2300.0 * x0 ** 3 + 1900.0 * x0 ** 2 + 2092.0 * x0 + 60.0
    den = 100.0 * x0 ** 3 + 500.0 * x0 ** 2 + 4.0 * x0 + 20.0
    return (1.0 - np.exp(np.power(-2.0 * x1, -1))) * num / den


def lim_poly(x):
    """2-D, x in [0, 1]^2 (Lim et al. polynomial)."""
    x0, x1 = x[0], x[1]
    return (9.0 + 2.5 * x0 - 17.5 * x1 + 2.5 * x0 * x1 + 19.0 * x1 ** 2 - 7.5 * x0 ** 3
            - 2.5 * x0 * x1 ** 2 - 5.5 * x1 ** 4 + x0 ** 3 * x1 ** 2)


def branin_rescaled(x):
    """2-D, x in [0, 1]^2 (rescaled Branin)."""
    u, v = 15.0 * x[0] - 5.0, 15.0 * x[1]
    core = (v - 1.275 * u ** 2 / np.pi ** 2 + 5.0 * u / np.pi - 6.0) ** 2
    return (core + 10.0 * (1.0 - 0.125 / np.pi) * np.cos(u) - 44.81) / 51.95
