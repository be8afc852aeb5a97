"""CPU float64 oracle for the exact-GP hot path of GPErks' ``GPEmulator``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gperks_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and there only as the checker / the baseline being timed.

What it restates
----------------
The reference (``/root/reference``, GPErks v0.1.2) is pure Python and delegates every FLOP of
this path to gpytorch / linear_operator (gpytorch==1.11, linear_operator==0.5.1 as pinned by
botorch==0.9.5, ``requirements/requirements.txt:1-2``).  Those packages are NOT in the image and
cannot be installed (no network), so this file restates their published exact-Cholesky
algorithm in plain torch (CPU, float64) -- the same ATen/LAPACK entry points gpytorch's
N <= ``max_cholesky_size`` path bottoms out in -- and follows the reference's own call sites:

* ``GPErks/train/emulator.py:348-383``  predict(): scale X, posterior mean, *noisy* variance,
  un-standardise (``likelihood(model(X))`` => +sigma^2 on the diagonal).
* ``GPErks/train/emulator.py:323-333``  loss = -mll(model(X), y)  (gpytorch divides by N).
* ``GPErks/train/emulator.py:385-401``  sample(): joint draw from the noisy predictive MVN.
* ``GPErks/gp/mean.py:29-36`` + ``GPErks/utils/polynomialfeatures.py:12-42``  polynomial mean.
* ``GPErks/gp/data/data_scaler.py:34-68``  UnitCubeScaler / StandardScaler (population std).
* ``GPErks/train/emulator.py:53-58``  fixed-noise mode (GreaterThan(1e-6), noise=1e-4).

Parity pin
----------
Pinned against the only result goldens the reference carries -- the executed cell outputs of
``notebooks/example_1.ipynb``, ``example_2.ipynb`` and ``example_9.ipynb`` (see
``tests/golden/notebook_goldens.json`` and ``tests/test_oracle_golden.py``).  These pin: softplus
constraints + 1e-4 noise floor, RBF and Matern-5/2 ARD forms, outputscale, polynomial-mean
ordering, noisy predictive variance, full predictive covariance, scaler conventions, MLL/N.
NOT pinned by any reference artefact ("parity unpinned" for these): Matern nu=1/2 and 3/2, MLL
gradients, fixed-noise mode, sample() values, N > 800 (where stock gpytorch switches to CG/Lanczos
approximations of the same quantities).  Those are covered by autograd / finite-difference
cross-checks inside this oracle only.
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

DTYPE = torch.float64

KIND_RBF = 0
KIND_MATERN12 = 1
KIND_MATERN32 = 2
KIND_MATERN52 = 3

_KIND_BY_NU = {0.5: KIND_MATERN12, 1.5: KIND_MATERN32, 2.5: KIND_MATERN52}


def kind_from_name(name: str, nu: Optional[float] = None) -> int:
    name = name.lower()
    if name in ("rbf", "rbfkernel"):
        return KIND_RBF
    if name in ("matern", "maternkernel"):
        return _KIND_BY_NU[float(nu if nu is not None else 2.5)]
    raise ValueError(f"unknown kernel {name!r}")


# ----------------------------------------------------------------------------------------------
# constraints  (gpytorch.constraints.Positive / GreaterThan: softplus transform)
# ----------------------------------------------------------------------------------------------
def softplus(raw: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.softplus(raw)


def inv_softplus(value: torch.Tensor) -> torch.Tensor:
    return value + torch.log(-torch.expm1(-value))


def constrained(raw, lower_bound: float = 0.0) -> torch.Tensor:
    """theta = softplus(raw) + lower_bound   (SURVEY App. A.2)."""
    raw = torch.as_tensor(raw, dtype=DTYPE)
    return softplus(raw) + lower_bound


# ----------------------------------------------------------------------------------------------
# polynomial mean  (GPErks/utils/polynomialfeatures.py:12-42, GPErks/gp/mean.py:29-36)
# ----------------------------------------------------------------------------------------------
def poly_indices(input_size: int, degree: int) -> List[List[int]]:
    """Sorted index multisets of size 1..degree, by degree then lexicographic."""
    out: List[List[int]] = []
    for k in range(1, degree + 1):
        out += [list(c) for c in itertools.combinations_with_replacement(range(input_size), k)]
    return out


def poly_features(X: torch.Tensor, degree: int) -> torch.Tensor:
    idx = poly_indices(X.shape[-1], degree)
    return torch.stack([torch.prod(X[..., i], dim=-1) for i in idx], dim=-1)


def mean_fn(X: torch.Tensor, weights, bias, degree: int = 1) -> torch.Tensor:
    """m(x) = phi_degree(x) @ w + b ; weights shape [P] or [P,1], bias scalar/[1]/None."""
    w = torch.as_tensor(weights, dtype=DTYPE).reshape(-1)
    phi = poly_features(X, degree)
    res = phi @ w
    if bias is not None:
        res = res + torch.as_tensor(bias, dtype=DTYPE).reshape(())
    return res


# ----------------------------------------------------------------------------------------------
# kernels  (gpytorch ScaleKernel(RBFKernel | MaternKernel(nu)), ARD)   SURVEY App. A.3
# ----------------------------------------------------------------------------------------------
def sq_dist(X1: torch.Tensor, X2: torch.Tensor, lengthscale: torch.Tensor, formulation: str = "direct") -> torch.Tensor:
    """r^2_ij = sum_d ((x1_id - x2_jd)/l_d)^2.

    ``direct``   : explicit differences (what the CUDA kernels do; most accurate).
    ``gpytorch`` : gpytorch's op sequence -- divide by l, centre on x1's mean, expand
                   |a|^2 - 2ab + |b|^2 with one matmul, clamp at 0.
    """
    ls = torch.as_tensor(lengthscale, dtype=DTYPE).reshape(1, -1)
    a = X1 / ls
    b = X2 / ls
    if formulation == "direct":
        diff = a.unsqueeze(1) - b.unsqueeze(0)
        return (diff * diff).sum(-1)
    if formulation == "gpytorch":
        adj = a.mean(0, keepdim=True)
        a = a - adj
        b = b - adj
        a_n = (a * a).sum(-1, keepdim=True)
        b_n = (b * b).sum(-1, keepdim=True)
        a_ = torch.cat([-2.0 * a, a_n, torch.ones_like(a_n)], dim=-1)
        b_ = torch.cat([b, torch.ones_like(b_n), b_n], dim=-1)
        return (a_ @ b_.T).clamp_min(0.0)
    raise ValueError(formulation)


def kernel_from_r2(r2: torch.Tensor, kind: int, outputscale) -> torch.Tensor:
    s = torch.as_tensor(outputscale, dtype=DTYPE)
    if kind == KIND_RBF:
        return s * torch.exp(-0.5 * r2)
    r = torch.sqrt(r2.clamp_min(1e-30))
    if kind == KIND_MATERN12:
        return s * torch.exp(-r)
    if kind == KIND_MATERN32:
        c = math.sqrt(3.0)
        return s * (1.0 + c * r) * torch.exp(-c * r)
    if kind == KIND_MATERN52:
        c = math.sqrt(5.0)
        return s * (1.0 + c * r + (5.0 / 3.0) * r2) * torch.exp(-c * r)
    raise ValueError(kind)


def kernel_matrix(X1, X2, kind: int, lengthscale, outputscale, formulation: str = "direct") -> torch.Tensor:
    return kernel_from_r2(sq_dist(X1, X2, lengthscale, formulation), kind, outputscale)


# ----------------------------------------------------------------------------------------------
# exact marginal log likelihood  (gpytorch ExactMarginalLogLikelihood; loss = -mll)  App. A.4
# ----------------------------------------------------------------------------------------------
def factorize(X, resid, kind, lengthscale, outputscale, noise):
    """K_hat = K + noise*I = L L^T ; alpha = K_hat^{-1} resid."""
    N = X.shape[0]
    K = kernel_matrix(X, X, kind, lengthscale, outputscale)
    K = K + torch.as_tensor(noise, dtype=DTYPE) * torch.eye(N, dtype=DTYPE)
    L = torch.linalg.cholesky(K)
    alpha = torch.cholesky_solve(resid.reshape(-1, 1), L).reshape(-1)
    return K, L, alpha


def neg_mll(X, y, kind, lengthscale, outputscale, noise, mean_x) -> torch.Tensor:
    """loss = -mll = [ 1/2 r^T K_hat^-1 r + sum log L_ii + N/2 log 2pi ] / N  (differentiable)."""
    N = X.shape[0]
    resid = y - mean_x
    _, L, alpha = factorize(X, resid, kind, lengthscale, outputscale, noise)
    quad = resid @ alpha
    logdet_half = torch.log(torch.diagonal(L)).sum()
    return (0.5 * quad + logdet_half + 0.5 * N * math.log(2.0 * math.pi)) / N


def neg_mll_analytic_grads(X, y, kind, lengthscale, outputscale, noise, mean_x) -> Dict[str, torch.Tensor]:
    """Closed-form gradients of ``neg_mll`` w.r.t. the CONSTRAINED hyper-parameters and the residual.

    d(loss)/d(theta) = (1/2N) sum_ij (K_hat^-1 - alpha alpha^T)_ij dK_hat_ij/dtheta   (App. A.4);
    d(loss)/d(resid) = alpha / N.  The CUDA path implements exactly these reductions.
    """
    N, D = X.shape
    ls = torch.as_tensor(lengthscale, dtype=DTYPE).reshape(-1)
    s = torch.as_tensor(outputscale, dtype=DTYPE)
    resid = y - mean_x
    _, L, alpha = factorize(X, resid, kind, ls, s, noise)
    Kinv = torch.cholesky_inverse(L)
    W = Kinv - torch.outer(alpha, alpha)
    r2 = sq_dist(X, X, ls)
    Kmat = kernel_from_r2(r2, kind, s)
    r = torch.sqrt(r2.clamp_min(1e-30))
    if kind == KIND_RBF:
        g = Kmat
    elif kind == KIND_MATERN52:
        c = math.sqrt(5.0)
        g = s * (5.0 / 3.0) * (1.0 + c * r) * torch.exp(-c * r)
    elif kind == KIND_MATERN32:
        c = math.sqrt(3.0)
        g = 3.0 * s * torch.exp(-c * r)
    else:  # MATERN12: dk/dl_d = s e^{-r}/r * delta_d^2/l_d^3, 0 at r = 0
        g = torch.where(r2 > 0, s * torch.exp(-r) / r, torch.zeros_like(r))
    g_ls = torch.empty(D, dtype=DTYPE)
    for d in range(D):
        delta = X[:, d].unsqueeze(1) - X[:, d].unsqueeze(0)
        g_ls[d] = (W * g * delta * delta).sum() / ls[d] ** 3 / (2.0 * N)
    return {
        "lengthscale": g_ls,
        "outputscale": (W * Kmat).sum() / s / (2.0 * N),
        "noise": torch.diagonal(W).sum() / (2.0 * N),
        "resid": alpha / N,
    }


# ----------------------------------------------------------------------------------------------
# prediction  (gpytorch DefaultPredictionStrategy, exact path + GaussianLikelihood)  App. A.5
# ----------------------------------------------------------------------------------------------
def predict(X, y, kind, lengthscale, outputscale, noise, mean_x, Xs, mean_xs,
            with_covar: bool = False, min_variance: float = 1e-10):
    """mu* = m(X*) + K* alpha ;  var* = k** - ||L^-1 K*^T||^2 + noise (noisy predictive variance)."""
    resid = y - mean_x
    _, L, alpha = factorize(X, resid, kind, lengthscale, outputscale, noise)
    Ks = kernel_matrix(Xs, X, kind, lengthscale, outputscale)          # [M, N]
    mean = mean_xs + Ks @ alpha
    V = torch.linalg.solve_triangular(L, Ks.T, upper=False)            # [N, M]
    s = torch.as_tensor(outputscale, dtype=DTYPE)
    nz = torch.as_tensor(noise, dtype=DTYPE)
    if with_covar:
        Kss = kernel_matrix(Xs, Xs, kind, lengthscale, outputscale)
        covar = Kss - V.T @ V + nz * torch.eye(Xs.shape[0], dtype=DTYPE)
        var = torch.diagonal(covar).clamp_min(min_variance)
        return mean, var, covar
    var = (s - (V * V).sum(0) + nz).clamp_min(min_variance)
    return mean, var


def sample(mean, covar, n_draws: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Joint MVN draws mu + R z with R the Cholesky root of the (noisy) predictive covariance."""
    M = mean.shape[0]
    R = torch.linalg.cholesky(covar)
    z = torch.randn(M, n_draws, dtype=DTYPE, generator=generator)
    return (mean.unsqueeze(1) + R @ z).T


# ----------------------------------------------------------------------------------------------
# scalers  (GPErks/gp/data/data_scaler.py:34-68)
# ----------------------------------------------------------------------------------------------
class UnitCube:
    def __init__(self, X: np.ndarray):
        self.a = np.min(X, axis=0)
        self.b = np.max(X, axis=0)

    def transform(self, X):
        return (X - self.a) / (self.b - self.a)


class Standard:
    def __init__(self, y: np.ndarray):
        self.mu = float(np.mean(y))
        self.sigma = float(np.std(y))

    def transform(self, y):
        return (y - self.mu) / self.sigma

    def inverse(self, y_, ystd_=None):
        return self.sigma * y_ + self.mu, (self.sigma * ystd_ if ystd_ is not None else None)


# ----------------------------------------------------------------------------------------------
# end-to-end emulator restatement (GPExperiment + GPEmulator.predict / loss)
# ----------------------------------------------------------------------------------------------
class OracleEmulator:
    """Restates ``GPExperiment.__init__`` (gp/experiment.py:79-104) + ``GPEmulator.predict`` /
    ``_train_step`` loss / ``sample`` for one fixed set of hyper-parameters, in float64."""

    def __init__(self, X_train: np.ndarray, y_train: np.ndarray, *, kind: int,
                 lengthscale, outputscale, noise, weights, bias, degree: int = 1):
        X_train = np.asarray(X_train, dtype=np.float64)
        if X_train.ndim == 1:
            X_train = X_train[:, None]
        self.scx = UnitCube(X_train)
        self.scy = Standard(np.asarray(y_train, dtype=np.float64))
        self.X = torch.from_numpy(self.scx.transform(X_train)).to(DTYPE)
        self.y = torch.from_numpy(self.scy.transform(np.asarray(y_train, dtype=np.float64))).to(DTYPE)
        self.kind = kind
        self.ls = torch.as_tensor(lengthscale, dtype=DTYPE).reshape(-1)
        self.s = torch.as_tensor(outputscale, dtype=DTYPE).reshape(())
        self.noise = torch.as_tensor(noise, dtype=DTYPE).reshape(())
        self.w = torch.as_tensor(weights, dtype=DTYPE).reshape(-1)
        self.b = None if bias is None else torch.as_tensor(bias, dtype=DTYPE).reshape(())
        self.degree = degree

    def _mean(self, X):
        return mean_fn(X, self.w, self.b, self.degree)

    def loss(self) -> float:
        return float(neg_mll(self.X, self.y, self.kind, self.ls, self.s, self.noise, self._mean(self.X)))

    def grads(self) -> Dict[str, torch.Tensor]:
        return neg_mll_analytic_grads(self.X, self.y, self.kind, self.ls, self.s, self.noise, self._mean(self.X))

    def predict_scaled(self, Xs_scaled: torch.Tensor, with_covar=False):
        return predict(self.X, self.y, self.kind, self.ls, self.s, self.noise, self._mean(self.X),
                       Xs_scaled, self._mean(Xs_scaled), with_covar=with_covar)

    def predict(self, X_new: np.ndarray, with_covar: bool = False, chunk: int = 4096):
        X_new = np.asarray(X_new, dtype=np.float64)
        if X_new.ndim == 1:
            X_new = X_new[:, None]
        Xs = torch.from_numpy(self.scx.transform(X_new)).to(DTYPE)
        if with_covar:
            mean, var, covar = self.predict_scaled(Xs, with_covar=True)
            y_mean, y_std = self.scy.inverse(mean.numpy(), np.sqrt(var.numpy()))
            return y_mean, y_std, (self.scy.sigma ** 2) * covar.numpy()
        # factor once, then stream X* in chunks (the reference materialises K* in one call)
        resid = self.y - self._mean(self.X)
        _, L, alpha = factorize(self.X, resid, self.kind, self.ls, self.s, self.noise)
        means, variances = [], []
        for lo in range(0, Xs.shape[0], chunk):
            xs = Xs[lo:lo + chunk]
            Ks = kernel_matrix(xs, self.X, self.kind, self.ls, self.s)
            means.append(self._mean(xs) + Ks @ alpha)
            V = torch.linalg.solve_triangular(L, Ks.T, upper=False)
            variances.append((self.s - (V * V).sum(0) + self.noise).clamp_min(1e-10))
        mean = torch.cat(means).numpy()
        std = np.sqrt(torch.cat(variances).numpy())
        return self.scy.inverse(mean, std)

    def sample(self, X_new: np.ndarray, n_draws: int, seed: int = 0) -> np.ndarray:
        X_new = np.asarray(X_new, dtype=np.float64)
        Xs = torch.from_numpy(self.scx.transform(X_new)).to(DTYPE)
        mean, _, covar = self.predict_scaled(Xs, with_covar=True)
        g = torch.Generator().manual_seed(seed)
        draws = sample(mean, covar, n_draws, g).numpy()
        return self.scy.sigma * draws + self.scy.mu


# ----------------------------------------------------------------------------------------------
# metrics used by the notebook goldens (torchmetrics MeanSquaredError / R2Score; ISE; chi^2; MD)
# ----------------------------------------------------------------------------------------------
def mse(pred, true) -> float:
    pred, true = np.asarray(pred), np.asarray(true)
    return float(np.mean((pred - true) ** 2))


def r2score(pred, true) -> float:
    pred, true = np.asarray(pred), np.asarray(true)
    return float(1.0 - np.sum((true - pred) ** 2) / np.sum((true - true.mean()) ** 2))


def ise(mean, std, true, ci: float = 2.0) -> float:
    """GPErks/utils/metrics.py:17-22."""
    e = np.abs(np.asarray(true) - np.asarray(mean)) / np.asarray(std)
    return float(np.sum(e < ci) / len(e))


def chi_squared(mean, std, true) -> float:
    """GPErks/perks/diagnostics.py: sum of squared standardised errors."""
    return float(np.sum(((np.asarray(true) - np.asarray(mean)) / np.asarray(std)) ** 2))


def mahalanobis(mean, covar, true) -> float:
    e = np.asarray(true) - np.asarray(mean)
    return float(e @ np.linalg.solve(covar, e))
