"""Distribution objects returned by ``model(x)`` / ``likelihood(model(x))``.

* ``MultivariateNormal``   -- train-mode prior N(m(X), k(X,X)) with a lazy covariance: nothing is
  evaluated until the marginal log-likelihood consumes it on the GPU.
* ``PosteriorDistribution`` -- eval-mode predictive distribution; ``.mean`` / ``.variance`` run the fused
  predict kernels, ``.covariance_matrix`` / ``.sample`` / ``.log_prob`` form the dense M x M covariance
  (test-set sized M) with the TRMM-store / SYRK kernels and factor it with the blocked CUDA Cholesky
  (``engine.DenseCholesky``); the ``L z`` of a draw is the same TRMM-store kernel (``_lower_times``) -- no library GEMM on the path.

They stand in for gpytorch.distributions.MultivariateNormal as used at GPErks/gp/model.py:21 and
GPErks/train/emulator.py:336-346, 356-360, 392-394."""
from __future__ import annotations

import warnings
from typing import Optional

import torch

F64 = torch.float64


def _psd_safe_cholesky(A: torch.Tensor) -> torch.Tensor:
    """Lower Cholesky factor of a dense covariance through the blocked CUDA factorisation (jitter retries as
    gpytorch's psd_safe_cholesky)."""
    from . import _native as nv
    from .engine import DenseCholesky
    nv.require_cuda()
    dev = A.device if A.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return DenseCholesky(A.detach().to(dev, F64).contiguous()).L.to(A.device)


def _lower_times(L: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """(L z)^T -> [n, M] for a lower-triangular L [M, M] and z [M, n], on the FP64 DMMA tile engine: gpx_trmm_store computes
    out[r][i] = sum_{k <= i} Zt[r][k] L[i][k] -- the kernel behind V^T = K* L^-T, handed L itself and z^T as its two operands."""
    from . import _native as nv
    nv.require_cuda()
    dev = L.device if L.is_cuda else torch.device("cuda", torch.cuda.current_device())
    M, n = z.shape
    Mp, npad = nv.padded(M), nv.padded(n)
    Lp = torch.zeros(Mp, Mp, dtype=F64, device=dev)
    Lp[:M, :M] = torch.tril(L.detach().to(dev, F64))
    Zt = torch.zeros(npad, Mp, dtype=F64, device=dev)
    Zt[:n, :M] = z.detach().to(dev, F64).T
    out = torch.empty(npad, Mp, dtype=F64, device=dev)
    with torch.cuda.device(dev):
        nv.check(nv.lib().gpx_trmm_store(nv.ptr(Lp), Mp, nv.ptr(Zt), npad, nv.ptr(out), nv.stream_ptr()), "gpx_trmm_store")
    return out[:n, :M].to(L.device)


class MultivariateNormal:
    def __init__(self, mean: torch.Tensor, covariance_matrix, validate_args: bool = False):
        self._mean = mean
        self._covar = covariance_matrix
        self.noise_module = None       # set by likelihood(dist): adds noise * I

    # gpytorch API ---------------------------------------------------------------------------
    @property
    def mean(self) -> torch.Tensor:
        return self._mean

    loc = mean

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self) -> torch.Tensor:
        K = self._covar.to_dense() if hasattr(self._covar, "to_dense") else self._covar
        if self.noise_module is not None:
            nz = self.noise_module.noise.detach().reshape(()).to(K.device, K.dtype)
            K = K + nz * torch.eye(K.shape[-1], dtype=K.dtype, device=K.device)
        return K

    @property
    def variance(self) -> torch.Tensor:
        if hasattr(self._covar, "diagonal") and not torch.is_tensor(self._covar):
            v = self._covar.diagonal()
        else:
            v = torch.diagonal(self._covar)
        if self.noise_module is not None:
            v = v + self.noise_module.noise.detach().reshape(()).to(v.device, v.dtype)
        return v

    @property
    def stddev(self) -> torch.Tensor:
        return self.variance.sqrt()

    def with_noise(self, likelihood) -> "MultivariateNormal":
        out = MultivariateNormal(self._mean, self._covar)
        out.noise_module = likelihood
        return out

    def rsample(self, sample_shape=torch.Size()) -> torch.Tensor:
        K = self.covariance_matrix
        L = _psd_safe_cholesky(K.to(F64))
        n = int(torch.Size(sample_shape).numel()) if len(sample_shape) else 1
        z = torch.randn(K.shape[-1], n, dtype=F64, device=K.device)
        draws = self._mean.to(K.device, F64).unsqueeze(0) + _lower_times(L, z)
        return draws.reshape(*sample_shape, K.shape[-1]).to(self._mean.dtype)

    def sample(self, sample_shape=torch.Size()) -> torch.Tensor:
        with torch.no_grad():
            return self.rsample(sample_shape)

    def confidence_region(self):
        s2 = self.stddev * 2.0
        return self.mean - s2, self.mean + s2


class PosteriorDistribution:
    """Exact-GP predictive distribution at ``x`` (scaled inputs), evaluated lazily on the GPU."""

    def __init__(self, model, x: torch.Tensor, noisy: bool = False):
        self.model = model
        self._x_orig = x                 # identity of the caller's tensor keys the per-epoch prediction cache
        self.x = x if x.dim() > 1 else x.unsqueeze(-1)
        self.noisy = noisy
        self._mv = None
        self._m = None
        self._cov = None
        self._cov_mean = None

    # -- helpers --------------------------------------------------------------------------------
    def _out(self, t: torch.Tensor) -> torch.Tensor:
        dt = self.x.dtype if self.x.dtype.is_floating_point else F64
        return t.to(self.x.device, dt)

    def _mean_device(self) -> torch.Tensor:
        """Posterior mean alone (``gpx_predict_mean``: bit-identical to the mean of the full predict at ~1 % of its cost) -- what the
        per-epoch metrics of a training loop ask of ``likelihood(model(X))``; the variance contraction runs only when a moment
        that needs it is read."""
        if self._mv is not None:
            return self._mv[0]
        if self._m is None:
            with torch.no_grad():
                def compute(eng):
                    xd = self.x.detach().to(eng.device, F64).contiguous()
                    prior = self.model.mean_module(self.x.detach()).detach().to(eng.device, F64)
                    return eng.predict_mean(xd, prior_mean=prior)
                self._m = self.model._cached_prediction(self._x_orig, "mean", compute)
        return self._m

    def _mean_var_device(self):
        if self._mv is None:
            with torch.no_grad():
                def compute(eng):
                    xd = self.x.detach().to(eng.device, F64).contiguous()
                    prior = self.model.mean_module(self.x.detach()).detach().to(eng.device, F64)
                    # large emulators: the variance contraction on the integer tensor cores where that keeps 1e-6 (engine.auto_precision)
                    mean, std = eng.predict(xd, prior_mean=prior, noisy=self.noisy, precision=eng.auto_precision(self.noisy))
                    return (mean, std * std)
                self._mv = self.model._cached_prediction(self._x_orig, self.noisy, compute)
        return self._mv

    def _cov_device(self) -> torch.Tensor:
        if self._cov is None:
            with torch.no_grad():
                eng = self.model._factorized_engine()
                prior = self.model.mean_module(self.x.detach()).detach()
                mean, cov = eng.predict_covar(self.x, prior, noisy=self.noisy)
                self._cov = cov
                self._cov_mean = mean
        return self._cov

    # -- gpytorch API ---------------------------------------------------------------------------
    @property
    def mean(self) -> torch.Tensor:
        return self._out(self._mean_device())

    loc = mean

    @property
    def variance(self) -> torch.Tensor:
        return self._out(self._mean_var_device()[1])

    @property
    def stddev(self) -> torch.Tensor:
        return self.variance.sqrt()

    @property
    def covariance_matrix(self) -> torch.Tensor:
        return self._out(self._cov_device())

    @property
    def lazy_covariance_matrix(self):
        return self.covariance_matrix

    def with_noise(self, likelihood) -> "PosteriorDistribution":
        return PosteriorDistribution(self.model, self._x_orig, noisy=True)

    def confidence_region(self):
        s2 = self.stddev * 2.0
        return self.mean - s2, self.mean + s2

    def rsample(self, sample_shape=torch.Size()) -> torch.Tensor:
        mean = self._mean_var_device()[0]
        L = _psd_safe_cholesky(self._cov_device())
        shape = torch.Size(sample_shape)
        n = int(shape.numel()) if len(shape) else 1
        z = torch.randn(mean.shape[0], n, dtype=F64, device=mean.device)
        draws = mean.unsqueeze(0) + _lower_times(L, z)
        return self._out(draws.reshape(*shape, mean.shape[0]))

    def sample(self, sample_shape=torch.Size()) -> torch.Tensor:
        with torch.no_grad():
            return self.rsample(sample_shape)

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        from .engine import DenseCholesky
        cov = self._cov_device()
        mean = self._cov_mean
        diff = value.detach().to(mean.device, F64).reshape(-1) - mean
        chol = DenseCholesky(cov, need_inverse=True)
        z = chol.solve_lower(diff)
        M = mean.shape[0]
        lp = -0.5 * (z @ z) - chol.logdet_half - 0.5 * M * 1.8378770664093453
        return self._out(lp)
