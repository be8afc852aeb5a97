"""Sensitivity-analysis benchmarks with closed-form Sobol' indices (Ishigami, Sobol'-Saltelli G*)."""
from itertools import combinations

import numpy as np
import pandas as pd


def Ishigami(x, a=7.0, b=0.1):
    """sin(X1) + a sin^2(X2) + b X3^4 sin(X1), Xi ~ U[-pi, pi]."""
    return np.sin(x[0]) + a * np.sin(x[1]) ** 2 + b * x[2] ** 4 * np.sin(x[0])


def _frames(STi, Si, Sij, D):
    names = [f"X{i + 1}" for i in range(D)]
    pairs = [f"({p}, {q})" for p, q in combinations(names, 2)]
    mk = lambda v, idx, col: pd.DataFrame(np.round(v, 6).reshape(-1, 1), index=idx, columns=[col])  # noqa: E731
    return mk(np.array(STi), names, "STi"), mk(np.array(Si), names, "Si"), mk(np.array(Sij), pairs, "Sij")


def Ishigami_theoretical_Si(a=7.0, b=0.1):
    pi4, pi8 = np.pi ** 4, np.pi ** 8
    V = 0.5 + a ** 2 / 8 + b * pi4 / 5 + b ** 2 * pi8 / 18
    V1 = 0.5 * (1 + b * pi4 / 5) ** 2
    V2 = a ** 2 / 8
    V13 = (1 / 18 - 1 / 50) * b ** 2 * pi8
    VT3 = 8 * b ** 2 * pi8 / 225
    return _frames([(V1 + VT3) / V, V2 / V, VT3 / V], [V1 / V, V2 / V, 0.0], [0.0, V13 / V, 0.0], 3)


def SobolGstar(x, a, delta, alpha):
    """prod_i ((1+alpha_i)|2(X_i + delta_i - floor(X_i + delta_i)) - 1|^alpha_i + a_i)/(1 + a_i)."""
    G = 1.0
    for i in range(len(a)):
        frac = x[i] + delta[i] - np.floor(x[i] + delta[i])
        G = G * (((1 + alpha[i]) * np.abs(2 * frac - 1) ** alpha[i] + a[i]) / (1 + a[i]))
    return G


def SobolGstar_theoretical_Si(a, delta, alpha):
    D = len(a)
    Vi = np.array([alpha[i] ** 2 / ((1 + 2 * alpha[i]) * (1 + a[i]) ** 2) for i in range(D)])
    V = np.prod(1 + Vi) - 1
    VTi = [Vi[i] * np.prod(np.delete(1 + Vi, i)) for i in range(D)]
    Vij = [Vi[p] * Vi[q] for p, q in combinations(range(D), 2)]
    return _frames(np.array(VTi) / V, Vi / V, np.array(Vij) / V, D)
