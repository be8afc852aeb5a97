"""Independent pin of the oracle where the reference carries no golden (SURVEY.md section 8c "parity unpinned": Matern nu=1/2
and 3/2, the MLL gradients, the latent-vs-noisy variance): scikit-learn's GaussianProcessRegressor is a separate
implementation of the same published exact-GP algorithm (Rasmussen & Williams, Alg. 2.1) with the same kernel forms as
gpytorch's RBFKernel / MaternKernel (ARD length scales, ConstantKernel = outputscale, WhiteKernel = likelihood noise).
It is not the reference, so DESIGN.md keeps the "unpinned against the reference" wording for these items; this file only
rules out an error shared by the oracle's own cross-checks."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

sk_gp = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel  # noqa: E402

KINDS = [(O.KIND_RBF, None), (O.KIND_MATERN12, 0.5), (O.KIND_MATERN32, 1.5), (O.KIND_MATERN52, 2.5)]


def _problem(seed, N=70, D=5, M=23):
    rng = np.random.default_rng(seed)
    X = rng.random((N, D))
    y = np.sin(4 * X[:, 0]) + X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(N)
    Xs = rng.random((M, D))
    ls = 0.3 + rng.random(D)
    return X, y, Xs, ls, 1.7, 0.03


def _sk_kernel(nu, ls, s, nz):
    base = RBF(length_scale=ls) if nu is None else Matern(length_scale=ls, nu=nu)
    return ConstantKernel(s) * base + WhiteKernel(nz)


@pytest.mark.parametrize("kind,nu", KINDS)
def test_kernel_matrix_matches_sklearn(kind, nu):
    X, _, Xs, ls, s, nz = _problem(3)
    k = _sk_kernel(nu, ls, s, nz)
    ours = O.kernel_matrix(torch.as_tensor(Xs), torch.as_tensor(X), kind, torch.as_tensor(ls), s).numpy()
    np.testing.assert_allclose(ours, k(Xs, X), rtol=1e-12, atol=1e-14)      # WhiteKernel is zero off the training set


@pytest.mark.parametrize("kind,nu", KINDS)
def test_mll_and_gradients_match_sklearn(kind, nu):
    X, y, _, ls, s, nz = _problem(4)
    N, D = X.shape
    gpr = sk_gp.GaussianProcessRegressor(kernel=_sk_kernel(nu, ls, s, nz), alpha=0.0, optimizer=None, normalize_y=False).fit(X, y)
    lml, dlml = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    Xt, yt, lst = torch.as_tensor(X), torch.as_tensor(y), torch.as_tensor(ls)
    zero = torch.zeros(N, dtype=O.DTYPE)
    loss = float(O.neg_mll(Xt, yt, kind, lst, torch.tensor(s, dtype=O.DTYPE), torch.tensor(nz, dtype=O.DTYPE), zero))
    np.testing.assert_allclose(loss, -lml / N, rtol=1e-11)
    g = O.neg_mll_analytic_grads(Xt, yt, kind, lst, torch.tensor(s, dtype=O.DTYPE), torch.tensor(nz, dtype=O.DTYPE), zero)
    # sklearn differentiates with respect to log-parameters, ordered (constant, length scales..., noise level)
    ours = np.concatenate([[float(g["outputscale"]) * s], g["lengthscale"].numpy() * ls, [float(g["noise"]) * nz]])
    np.testing.assert_allclose(ours, -dlml / N, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("kind,nu", KINDS)
def test_predictive_mean_and_noisy_std_match_sklearn(kind, nu):
    X, y, Xs, ls, s, nz = _problem(5)
    N, M = X.shape[0], Xs.shape[0]
    gpr = sk_gp.GaussianProcessRegressor(kernel=_sk_kernel(nu, ls, s, nz), alpha=0.0, optimizer=None, normalize_y=False).fit(X, y)
    mu_sk, cov_sk = gpr.predict(Xs, return_cov=True)      # K** includes the WhiteKernel diagonal => noisy covariance
    mu, var, C = O.predict(torch.as_tensor(X), torch.as_tensor(y), kind, torch.as_tensor(ls), s, nz,
                           torch.zeros(N, dtype=O.DTYPE), torch.as_tensor(Xs), torch.zeros(M, dtype=O.DTYPE), with_covar=True)
    np.testing.assert_allclose(mu.numpy(), mu_sk, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(C.numpy(), cov_sk, rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(var.numpy(), np.diag(cov_sk), rtol=1e-8, atol=1e-11)
