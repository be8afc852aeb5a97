"""Bastos & O'Hagan (2009) emulator diagnostics on the test split (API of GPErks/perks/diagnostics.py):
chi-squared, Mahalanobis distance, credible interval, individual / Cholesky / pivoted-Cholesky errors."""
import numpy as np
import pandas as pd
from scipy import linalg, stats


class Diagnostics:
    def __init__(self, emulator):
        self.emulator = emulator
        ds = emulator.experiment.dataset
        self.X_test, self.y_test = ds.X_test, ds.y_test
        self.y_pred_mean, self.y_pred_std, self.y_pred_covar = emulator.predict(self.X_test, with_covar=True)
        self.n, self.q = ds.sample_size, ds.input_size
        self.m = self.X_test.shape[0] if self.X_test.ndim > 1 else len(self.X_test)

    def chi_squared(self):
        return DChi2(self.y_test, self.y_pred_mean, self.y_pred_std), self.m, np.sqrt(2 * self.m)

    def mahalanobis_distance(self):
        sd = np.sqrt(2 * self.m * (self.m + self.n - self.q - 2) / (self.n - self.q - 4))
        return DMD(self.y_test, self.y_pred_mean, self.y_pred_covar), self.m, sd

    def credible_interval(self):
        draws = self.emulator.sample(self.X_test)
        ci = [DCI(self.y_test, y, self.y_pred_std) for y in draws]
        return DCI(self.y_test, self.y_pred_mean, self.y_pred_std), np.mean(ci), np.std(ci)

    def summary(self):
        rows = [list(self.chi_squared()), list(self.mahalanobis_distance()), list(self.credible_interval())]
        print(pd.DataFrame(data=np.around(np.array(rows), decimals=4),
                           index=["Chi-Squared", "Mahalanobis Distance", "Credible Interval"],
                           columns=["Observed", "Expected", "Std"]))

    def errors(self, errors_type: str = "pivoted"):
        if errors_type == "correlated":
            return DI(self.y_test, self.y_pred_mean, self.y_pred_std)
        if errors_type == "uncorrelated":
            return DG(self.y_test, self.y_pred_mean, self.y_pred_covar)
        if errors_type == "pivoted":
            return DPC(self.y_test, self.y_pred_mean, self.y_pred_covar)
        raise ValueError("Not a valid errors type! Choose among: 'correlated', 'uncorrelated', 'pivoted'.")

    def plot(self, errors_type: str = "pivoted"):
        raise NotImplementedError("plotting is outside the B200 hot path (matplotlib not required)")


def pivoted_Cholesky_factorization(A):
    """LAPACK pstrf; returns (L, P) with P^T A P = L L^T."""
    pstrf = linalg.get_lapack_funcs("pstrf", dtype=float)
    n = A.shape[0]
    U, piv, _rank, _info = pstrf(A)
    P = np.zeros((n, n), dtype=float)
    P[piv - 1, np.arange(n)] = 1
    return np.triu(U).T, P


def DI(y_true, y_pred_mean, y_pred_std):
    return (y_true - y_pred_mean) / y_pred_std


def DChi2(y_true, y_pred_mean, y_pred_std):
    return np.sum(DI(y_true, y_pred_mean, y_pred_std) ** 2)


def DMD(y_true, y_pred_mean, y_pred_covar):
    x = (y_true - y_pred_mean).reshape(-1, 1)
    return np.squeeze(x.T @ linalg.cho_solve(linalg.cho_factor(y_pred_covar), x))


def DG(y_true, y_pred_mean, y_pred_covar):
    x = (y_true - y_pred_mean).reshape(-1, 1)
    return np.squeeze(linalg.solve(linalg.cholesky(y_pred_covar, lower=True), x))


def DPC(y_true, y_pred_mean, y_pred_covar):
    # NOTE: the reference unpacks (P, L) from a function returning (L, P) (diagnostics.py:189); kept as is.
    x = (y_true - y_pred_mean).reshape(-1, 1)
    P, L = pivoted_Cholesky_factorization(y_pred_covar)
    return np.squeeze(linalg.solve(L, P.T @ x))


def DCI(y_true, y_pred_mean, y_pred_std, marginal_credible_interval: float = 95.0):
    z = stats.norm.ppf(0.5 + 0.01 * marginal_credible_interval / 2)
    x = DI(y_true, y_pred_mean, y_pred_std)
    return np.count_nonzero(np.abs(x) < z) / len(x)
