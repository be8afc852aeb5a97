"""Device-side exact-GP engine: owns the HBM-resident factors of one training set and drives the
hand-written sm_100a kernels through the C ABI (include/gpx.h).

HBM layout for N training points (Np = N rounded up to 128, nb = Np/128), all FP64 row-major:
    K   [Np,Np]  K_hat lower tiles -> overwritten by L
    S   [Np,Np]  L^-1 (lower) + L^-T mirrored in the upper off-diagonal blocks
    W   [Np,Np]  scratch of the recursive inverse, then K_hat^-1 for the gradient
    DT  [nb,128,128] transposed diagonal blocks of L^-1
    r, u, alpha [Np]; theta [D+2]; logdet_blk [nb]; out [3]; grad [D+2]; partial [nb*nb*(D+2)]
3*Np^2*8 B dominates: 1.6 GB at N=8192, 16.5 GB at N=26214 -- far below the B200's 180 GB.

This module replaces the arithmetic gpytorch performs behind GPErks' call sites
``GPErks/train/emulator.py:326-327`` (MLL forward/backward) and ``:356`` (prediction).
"""
from __future__ import annotations

import os
import warnings
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _native as nv
from .constants import (
    AUTO_I8_MAX_N,
    AUTO_I8_MAX_SCALE_TO_NOISE,
    AUTO_I8_MIN_N,
    AUTO_I8_PROBE_POINTS,
    AUTO_I8_PROBE_POOL,
    AUTO_I8_PROBE_TOL,
    AUTO_I8_SLICES,
    AUTO_I8_W5_MAX_SCALE_TO_NOISE,
    AUTO_I8_W6_MAX_SCALE_TO_NOISE,
    AUTO_I8_WIDE_MAX_N,
)

F64 = torch.float64

# precision name -> (slices, bits per slice) of the sliced-integer variance path (gpx_i8.cu)
I8_MODES = {"int8x5": (5, 7), "int8x6": (6, 7), "int8x5w": (5, 8), "int8x6w": (6, 8)}
PRECISIONS = ("fp64", "tf32x3", "tf32x3_fast") + tuple(I8_MODES)


class NotPSDError(RuntimeError):
    """Mirrors linear_operator.utils.errors.NotPSDError (raised after the jitter retries fail)."""


class NumericalWarning(RuntimeWarning):
    """Mirrors gpytorch.utils.warnings.NumericalWarning."""


@dataclass
class PolyMeanSpec:
    """Polynomial prior mean evaluated inside the predict epilogue (GPErks/gp/mean.py:29-36)."""
    feat_idx: torch.Tensor        # int32 [P, degree] on device, -1 padded
    degree: int
    weights: torch.Tensor         # [P] device f64
    bias: Optional[torch.Tensor]  # [1] device f64 or None


class ExactGPEngine:
    """Factorisation cache + kernel launcher for one (X_train, kernel kind)."""

    def __init__(self, X: torch.Tensor, kind: int, device: Optional[torch.device] = None):
        nv.require_cuda()
        self.lib = nv.lib()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("ExactGPEngine needs a CUDA device")
        X = X.detach()
        if X.dim() == 1:
            X = X.unsqueeze(-1)
        self.X = X.to(self.device, F64).contiguous()
        self.N, self.D = self.X.shape
        if self.D > self.lib.gpx_max_dim():
            raise ValueError(f"input dimension {self.D} exceeds the kernel limit {self.lib.gpx_max_dim()}")
        self.kind = int(kind)
        self.Np = nv.padded(self.N)
        self.nb = self.Np // nv.TILE
        dev = self.device
        Np, nb, D = self.Np, self.nb, self.D
        self.K = torch.empty(Np, Np, dtype=F64, device=dev)
        self.S = torch.empty(Np, Np, dtype=F64, device=dev)
        self.W = torch.empty(Np, Np, dtype=F64, device=dev)
        self.DT = torch.empty(nb, nv.TILE, nv.TILE, dtype=F64, device=dev)
        self.logdet_blk = torch.zeros(nb, dtype=F64, device=dev)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        self.r = torch.zeros(Np, dtype=F64, device=dev)
        self.u = torch.zeros(Np, dtype=F64, device=dev)
        self.alpha = torch.zeros(Np, dtype=F64, device=dev)
        self.out = torch.zeros(3, dtype=F64, device=dev)
        self.grad = torch.zeros(D + 2, dtype=F64, device=dev)
        self.partial = torch.zeros(nb * nb * (D + 2), dtype=F64, device=dev)
        self.theta = torch.zeros(D + 2, dtype=F64, device=dev)
        self.theta_in = torch.zeros(D + 2, dtype=F64, device=dev)      # theta as passed in (before jitter)
        self.theta_latent = torch.zeros(D + 2, dtype=F64, device=dev)  # theta with noise = 0 (latent variance)
        self.factored = False
        self.factor_key = None
        self.has_kinv = False
        self.jitter_used = 0.0
        self._ws: Optional[torch.Tensor] = None
        self.n_factorizations = 0
        self.has_split = False
        self.S_hi: Optional[torch.Tensor] = None      # TF32 mode: (hi, lo) float split of L^-1
        self.S_lo: Optional[torch.Tensor] = None
        self.i8_mode = None                           # sliced-integer mode: (slices, bits) of the int8 planes of L^-1 held (None = stale)
        self._i8_probe = {}                           # precision -> (n_factorizations, tol, ok) of int8_probe_ok()
        self._i8_bound = {}                           # precision -> (key, ok, bound) of i8_sums_exact()
        self._probe_set = None                        # (n_factorizations, points, their FP64 std)
        self._mma_stream: Optional[torch.cuda.Stream] = None   # high-priority stream of the tensor-core kernel (chunk pipeline)
        # K_hat^-1 = L^-T L^-1 for the gradient: six-slice integer tensor-core product where its tiles pay and its int32 sums
        # cannot overflow (entries within ~1e-10 of their row/column scale of the DMMA result), else the FP64 DMMA lauum;
        # GPX_LAUUM=fp64 forces the latter
        self.lauum_slices = (AUTO_I8_SLICES if (AUTO_I8_MIN_N <= self.N <= AUTO_I8_MAX_N
                                                and os.environ.get("GPX_LAUUM", "") != "fp64") else 0)
        # The merge levels of the recursive inverse with segments of >= 1024 rows can run on the integer tensor cores as well
        # (gpx_trtri_i8: six slices of 8 bits for Np <= 16384, of 7 bits above).  L^-1 then carries ~1e-10 (8 bit) / ~1e-8 (7 bit) of
        # its norm instead of ~1e-13 -- enough for the loss, its gradients and the per-epoch metrics, NOT for predictions at small
        # noise (amplification ~ outputscale / noise).  So it is used only while ``approx_inverse_ok`` is set -- GPEmulator's
        # training loops set it around their epochs -- and the first prediction afterwards completes the factors with the FP64
        # inverse (``ensure_exact_inverse``).  GPX_TRTRI=i8 (or engine.trtri_i8 = True) forces the integer inverse everywhere
        # (experiments); GPX_TRTRI=fp64 disables it.
        self.trtri_i8 = bool(self.lauum_slices and self.Np >= 2048 and os.environ.get("GPX_TRTRI", "") == "i8")
        self._trtri_i8_capable = bool(self.lauum_slices and self.Np >= 2048 and os.environ.get("GPX_TRTRI", "") != "fp64")
        # Inside the same loops the Cholesky factorisation itself sends its bulk trailing updates through the integer tensor
        # cores (gpx_potrf_i8: six 8-bit digits per operand, L to ~1e-12 of its scale); GPX_POTRF=fp64 keeps them on FP64 DMMA.
        # Needs the two-level schedule (more than 24 block columns with integer updates) to have any bulk update at all.
        self._potrf_i8_capable = bool(self._trtri_i8_capable and self.Np > 24 * 128 and os.environ.get("GPX_POTRF", "") != "fp64")
        self.approx_inverse_ok = False       # True inside a training loop: factorize() may use the integer inverse / Cholesky
        self.inverse_exact = True            # False while S / alpha / out come from the integer inverse
        self.chol_exact = True               # False while L itself comes from gpx_potrf_i8
        self._lauum_ws: Optional[torch.Tensor] = None
        if self.lauum_slices:       # one workspace for both, allocated up front: gradients() may run inside a CUDA-graph capture
            nbytes = int(self.lib.gpx_lauum_i8_workspace_bytes(self.Np, self.lauum_slices))
            if self._trtri_i8_capable:
                nbytes = max(nbytes, int(self.lib.gpx_trtri_i8_workspace_bytes(self.Np)))
            if self._potrf_i8_capable:
                nbytes = max(nbytes, int(self.lib.gpx_potrf_i8_workspace_bytes(self.Np)))
            self._lauum_ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.S_i8: Optional[torch.Tensor] = None      # [slices, Np, Np] int8 planes of L^-1
        self.S_i8_inv_scale: Optional[torch.Tensor] = None   # [Np] inverse power-of-two row scales

    # ------------------------------------------------------------------------------------------
    def _launch(self, name: str, *args) -> None:
        nv.check(getattr(self.lib, name)(*args), name)

    def factorize(self, theta: torch.Tensor, resid: torch.Tensor, key=None, check: bool = True) -> None:
        """K_hat(theta) = L L^T, S = L^-1, alpha = K_hat^-1 resid, out = (-mll, quad, logdet/2).

        ``theta``: device [D+2] = (1/lengthscale, outputscale, noise).  On a failed factorisation the
        gpytorch ``psd_safe_cholesky`` schedule is followed: jitter 1e-8 * 10^i (i<3) added to the
        diagonal with a NumericalWarning, then NotPSDError."""
        with torch.cuda.device(self.device):
            st = nv.stream_ptr()
            self.factored = False
            self.theta.copy_(theta.detach().reshape(-1))
            self.theta_in.copy_(self.theta)
            self.theta_latent.copy_(self.theta)
            self.theta_latent[self.D + 1:].zero_()      # device-side (a Python scalar assignment is a host copy: not capturable)
            self.r[: self.N].copy_(resid.detach().reshape(-1))
            p = nv.ptr
            jitter = 0.0
            potrf_i8 = self._potrf_i8_capable and self.approx_inverse_ok and not self.trtri_i8
            if potrf_i8:
                need = int(self.lib.gpx_potrf_i8_workspace_bytes(self.Np))
                if self._lauum_ws is None or self._lauum_ws.numel() < need:
                    self._lauum_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            for attempt in range(4):
                if attempt > 0:
                    new_jitter = 1e-8 * 10 ** (attempt - 1)
                    warnings.warn(f"A not p.d., added jitter of {new_jitter:.1e} to the diagonal", NumericalWarning)
                    self.theta[self.D + 1] += new_jitter - jitter
                    jitter = new_jitter
                self._launch("gpx_kmat_sym", p(self.X), self.N, self.D, p(self.theta), self.kind, p(self.K),
                             self.Np, st)
                if potrf_i8:
                    self._launch("gpx_potrf_i8", p(self.K), self.Np, self.N, p(self.S), p(self.DT), p(self.logdet_blk),
                                 p(self.info), p(self._lauum_ws), self._lauum_ws.numel(), st)
                else:
                    self._launch("gpx_potrf", p(self.K), self.Np, self.N, p(self.S), p(self.DT), p(self.logdet_blk),
                                 p(self.info), st)
                if not check:
                    break
                info = int(self.info.item())       # the one host sync of a factorisation
                if info == 0:
                    break
                if attempt == 3:
                    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to "
                                      f"{jitter:.1e} (first non-positive pivot at {info}).")
            self.jitter_used = jitter
            use_i8 = self._trtri_i8_capable and (self.trtri_i8 or self.approx_inverse_ok)
            if use_i8:
                need = int(self.lib.gpx_trtri_i8_workspace_bytes(self.Np))
                if self._lauum_ws is None or self._lauum_ws.numel() < need:      # switched on after construction
                    self._lauum_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                self._launch("gpx_trtri_i8", p(self.K), p(self.S), p(self.DT), p(self.W), self.Np,
                             8 if self.Np <= int(self.lib.gpx_i8_safe_n(8)) else 7, p(self._lauum_ws), self._lauum_ws.numel(), st)
            else:
                self._launch("gpx_trtri", p(self.K), p(self.S), p(self.DT), p(self.W), self.Np, st)
            self._launch("gpx_solve_alpha", p(self.S), p(self.DT), self.Np, p(self.r), p(self.u), p(self.alpha), st)
            self._launch("gpx_mll_value", p(self.r), p(self.alpha), self.Np, self.N, p(self.logdet_blk),
                         p(self.out), st)
        self.factor_key = key
        self._mark_factorized()
        self.inverse_exact = not (use_i8 and not self.trtri_i8)      # a forced integer inverse (GPX_TRTRI=i8) is never completed
        self.chol_exact = not potrf_i8

    def ensure_exact_inverse(self) -> None:
        """Completes factors whose inverse came from the integer tensor cores (training loop) with the FP64 inverse: L is still in
        ``K``, so only trtri, alpha and the loss are redone (~40 % of a factorisation); every derived cache is dropped."""
        if self.inverse_exact or not self.factored:
            return
        if not self.chol_exact:
            # L itself came from the integer tensor cores: the whole factorisation is redone in FP64 for the same inputs
            ok, self.approx_inverse_ok = self.approx_inverse_ok, False
            try:
                nf = self.n_factorizations
                self.factorize(self.theta_in.clone(), self.r[: self.N].clone(), key=self.factor_key)
                self.n_factorizations = nf       # same factorisation, new version of its derived quantities
            finally:
                self.approx_inverse_ok = ok
            self.n_inverse_refreshes = getattr(self, "n_inverse_refreshes", 0) + 1
            return
        with torch.cuda.device(self.device):
            st, p = nv.stream_ptr(), nv.ptr
            # the diagonal blocks of S / DT are potrf's (exact, untouched by either inverse): every merge level is simply redone
            self._launch("gpx_trtri", p(self.K), p(self.S), p(self.DT), p(self.W), self.Np, st)
            self._launch("gpx_solve_alpha", p(self.S), p(self.DT), self.Np, p(self.r), p(self.u), p(self.alpha), st)
            self._launch("gpx_mll_value", p(self.r), p(self.alpha), self.Np, self.N, p(self.logdet_blk), p(self.out), st)
        self._mark_factorized()
        self.n_factorizations -= 1           # same factorisation, new version of its derived quantities
        self.n_inverse_refreshes = getattr(self, "n_inverse_refreshes", 0) + 1
        self.inverse_exact = True

    def _mark_factorized(self, has_kinv: bool = False) -> None:
        """Python-side state after K, S, DT, alpha, out have been (re)computed on the device -- by ``factorize`` or by a replayed
        CUDA graph (train/batched.py): every cache derived from the previous factors is stale."""
        self.has_kinv = has_kinv
        self.has_split = False
        self.i8_mode = None
        self.factored = True
        self.n_factorizations += 1

    def matches(self, theta: torch.Tensor, resid: torch.Tensor) -> bool:
        """True when the cached factors were computed for exactly these hyper-parameters / residuals
        (bitwise; one tiny device comparison + host read)."""
        if not self.factored:
            return False
        t = theta.detach().reshape(-1)
        r = resid.detach().reshape(-1)
        if t.device != self.device or r.device != self.device:
            t, r = t.to(self.device, F64), r.to(self.device, F64)
        return bool(torch.equal(t, self.theta_in)) and bool(torch.equal(r, self.r[: self.N]))

    def gradients(self) -> torch.Tensor:
        """d(-mll)/d(lengthscale_d, outputscale, noise) for the cached factors -> device [D+2]."""
        with torch.cuda.device(self.device):
            st = nv.stream_ptr()
            p = nv.ptr
            if not self.has_kinv:
                if self.lauum_slices:
                    # K_hat^-1 on the integer tensor cores (lower 128-tiles only: what gpx_mll_grad reads)
                    nbytes = int(self.lib.gpx_lauum_i8_workspace_bytes(self.Np, self.lauum_slices))
                    if self._lauum_ws is None or self._lauum_ws.numel() < nbytes:
                        self._lauum_ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                    self._launch("gpx_lauum_i8", p(self.S), p(self.DT), p(self.W), self.Np, self.lauum_slices,
                                 p(self._lauum_ws), self._lauum_ws.numel(), st)
                else:
                    self._launch("gpx_lauum", p(self.S), p(self.DT), p(self.W), self.Np, st)
                self.has_kinv = True
            self._launch("gpx_mll_grad", p(self.X), self.N, self.D, p(self.theta), self.kind, p(self.W), self.Np,
                         p(self.alpha), p(self.partial), p(self.grad), st)
        return self.grad

    @property
    def loss(self) -> torch.Tensor:
        return self.out[0]

    # ------------------------------------------------------------------------------------------
    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() * 8 < nbytes:
            self._ws = None
            self._ws = torch.empty((nbytes + 7) // 8, dtype=F64, device=self.device)
        return self._ws

    def pick_chunk(self, M: int, budget_bytes: int = 2 << 30) -> int:
        """Chunk of test points whose K* tile (Mc x Np doubles) stays within ``budget_bytes``."""
        mc = max(nv.TILE, (budget_bytes // (self.Np * 8)) // nv.TILE * nv.TILE)
        mc = min(mc, 32768)
        return int(min(mc, nv.padded(M)))

    def predict(self, Xs: torch.Tensor, *, xa: Optional[torch.Tensor] = None, xr: Optional[torch.Tensor] = None,
                prior_mean: Optional[torch.Tensor] = None, poly: Optional[PolyMeanSpec] = None,
                y_mu: float = 0.0, y_sigma: float = 1.0, min_var: float = 1e-10, noisy: bool = True,
                precision: str = "fp64", tf32_kchunk: int = 32,
                out_mean: Optional[torch.Tensor] = None, out_std: Optional[torch.Tensor] = None,
                chunk: Optional[int] = None):
        """Posterior mean and noisy std for device-resident points Xs [M, D] using the cached factors.

        ``xa``/``xr``: unit-cube scaler (min, range) applied in-kernel to raw inputs; omit for inputs that
        are already scaled.  Returns (mean, std) device tensors, un-standardised with (y_mu, y_sigma)."""
        if not self.factored:
            raise RuntimeError("predict() before factorize()")
        if not self.approx_inverse_ok:
            self.ensure_exact_inverse()
        if Xs.dim() == 1:
            Xs = Xs.unsqueeze(-1)
        Xs = Xs.to(self.device, F64).contiguous()
        M = Xs.shape[0]
        if Xs.shape[1] != self.D:
            raise ValueError(f"expected inputs with {self.D} columns, got {Xs.shape[1]}")
        if out_mean is None:
            out_mean = torch.empty(M, dtype=F64, device=self.device)
        if out_std is None:
            out_std = torch.empty(M, dtype=F64, device=self.device)
        if M == 0:
            return out_mean, out_std
        Mc = int(chunk) if chunk else self.pick_chunk(M)
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        # int8x5 / int8x6 (7-bit slices), int8x5w / int8x6w (8-bit slices): the FP64 contraction from 5 / 6 int8 slices per
        # operand on the integer tensor cores (exact int32 accumulation), mean bit-identical to fp64 (gpx_i8.cu)
        i8, i8_bits = I8_MODES.get(precision, (0, 0))
        if i8 and self.Np > int(self.lib.gpx_i8_max_n(i8_bits)):
            raise ValueError(f"the sliced-integer mode holds its sums in int32: N <= {int(self.lib.gpx_i8_max_n(i8_bits))} "
                             f"for {i8_bits}-bit slices")
        if i8 and not self.i8_sums_exact(precision):
            raise ValueError(f"precision={precision!r}: the digits of this L^-1 do not keep the int32 sums of {i8_bits}-bit slices exact at "
                             f"N_pad = {self.Np} (above {int(self.lib.gpx_i8_safe_n(i8_bits))} that depends on the data); use 'int8x6'")
        # tf32x3: K* evaluated in FP64 then split (posterior mean bit-identical to fp64; variance on tcgen05);
        # tf32x3_fast: K* evaluated in FP32 as well (mean moves by ~1e-4 of its scale for noise-like targets)
        tf32 = precision.startswith("tf32")
        kchunk = abs(int(tf32_kchunk)) or 32
        if precision == "tf32x3":
            kchunk = -kchunk
        if tf32:
            self._ensure_split()
            Mc = (Mc + 255) // 256 * 256         # the tcgen05 kernel works on pairs of 128-point tiles
            ws_bytes = int(self.lib.gpx_predict_tf32_workspace_bytes(self.Np, Mc))
        elif i8:
            self._ensure_i8(i8, i8_bits)
            ws_bytes = int(self.lib.gpx_predict_i8_workspace_bytes(self.Np, Mc, i8))
        else:
            ws_bytes = int(self.lib.gpx_predict_workspace_bytes(self.Np, Mc))
        ws = self._workspace(ws_bytes)
        p = nv.ptr
        feat = wts = bias = None
        n_feat = degree = 0
        if prior_mean is not None:
            prior_mean = prior_mean.detach().to(self.device, F64).contiguous()
        elif poly is not None:
            feat, wts, bias = poly.feat_idx, poly.weights, poly.bias
            n_feat, degree = int(feat.shape[0]), int(poly.degree)
        with torch.cuda.device(self.device):
            theta = self.theta if noisy else self.theta_latent
            if tf32:
                self._launch("gpx_predict_meanvar_tf32", p(Xs), M, p(self.X), self.N, self.D, p(theta), p(xa), p(xr),
                             self.kind, p(self.S_hi), p(self.S_lo), self.Np, p(self.alpha), p(prior_mean), p(feat),
                             n_feat, degree, p(wts), p(bias), float(y_mu), float(y_sigma), float(min_var),
                             p(out_mean), p(out_std), p(ws), ws_bytes, Mc, kchunk, nv.stream_ptr())
            elif i8:
                # chunk pipeline: the tensor-core kernel of chunk c on a high-priority stream of this engine while the K*
                # builder of chunk c+1 runs on the current stream (GPX_I8_OVERLAP=0: everything on the current stream)
                mma = None
                if M > Mc and os.environ.get("GPX_I8_OVERLAP", "1") != "0" and not torch.cuda.is_current_stream_capturing():
                    if self._mma_stream is None:
                        self._mma_stream = torch.cuda.Stream(device=self.device, priority=-1)
                    mma = self._mma_stream.cuda_stream
                self._launch("gpx_predict_meanvar_i8", p(Xs), M, p(self.X), self.N, self.D, p(theta), p(xa), p(xr),
                             self.kind, p(self.S_i8), p(self.S_i8_inv_scale), i8, i8_bits, self.Np, p(self.alpha),
                             p(prior_mean), p(feat), n_feat, degree, p(wts), p(bias), float(y_mu), float(y_sigma),
                             float(min_var), p(out_mean), p(out_std), p(ws), ws_bytes, Mc, nv.stream_ptr(), mma)
            else:
                self._launch("gpx_predict_meanvar", p(Xs), M, p(self.X), self.N, self.D, p(theta), p(xa), p(xr),
                             self.kind, p(self.S), self.Np, p(self.alpha), p(prior_mean), p(feat), n_feat, degree,
                             p(wts), p(bias), float(y_mu), float(y_sigma), float(min_var), p(out_mean), p(out_std),
                             p(ws), ws_bytes, Mc, nv.stream_ptr())
        return out_mean, out_std

    def predict_mean(self, Xs: torch.Tensor, *, xa: Optional[torch.Tensor] = None, xr: Optional[torch.Tensor] = None,
                     prior_mean: Optional[torch.Tensor] = None, poly: Optional[PolyMeanSpec] = None, y_mu: float = 0.0,
                     y_sigma: float = 1.0, out_mean: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Posterior mean only for device-resident points Xs [M, D] (``gpx_predict_mean``: K* is never materialised and the
        variance contraction -- N^2 FLOP per point -- is skipped); bit-identical to the mean ``predict`` returns."""
        if not self.factored:
            raise RuntimeError("predict_mean() before factorize()")
        if not self.approx_inverse_ok:
            self.ensure_exact_inverse()
        if Xs.dim() == 1:
            Xs = Xs.unsqueeze(-1)
        Xs = Xs.to(self.device, F64).contiguous()
        M = Xs.shape[0]
        if Xs.shape[1] != self.D:
            raise ValueError(f"expected inputs with {self.D} columns, got {Xs.shape[1]}")
        if out_mean is None:
            out_mean = torch.empty(M, dtype=F64, device=self.device)
        if M == 0:
            return out_mean
        p = nv.ptr
        feat = wts = bias = None
        n_feat = degree = 0
        if prior_mean is not None:
            prior_mean = prior_mean.detach().to(self.device, F64).contiguous()
        elif poly is not None:
            feat, wts, bias = poly.feat_idx, poly.weights, poly.bias
            n_feat, degree = int(feat.shape[0]), int(poly.degree)
        with torch.cuda.device(self.device):
            self._launch("gpx_predict_mean", p(Xs), M, p(self.X), self.N, self.D, p(self.theta), p(xa), p(xr), self.kind,
                         p(self.alpha), self.Np, p(prior_mean), p(feat), n_feat, degree, p(wts), p(bias), float(y_mu),
                         float(y_sigma), p(out_mean), nv.stream_ptr())
        return out_mean

    def _ensure_split(self) -> None:
        """(hi, lo) TF32 split of L^-1 for the tensor-core variance path; refreshed after every factorisation."""
        if self.has_split:
            return
        if self.S_hi is None:
            self.S_hi = torch.empty(self.Np, self.Np, dtype=torch.float32, device=self.device)
            self.S_lo = torch.empty(self.Np, self.Np, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._launch("gpx_split_tf32", nv.ptr(self.S), self.Np * self.Np, nv.ptr(self.S_hi), nv.ptr(self.S_lo),
                         self.Np, nv.stream_ptr())
        self.has_split = True

    @staticmethod
    def auto_candidates(n: int, outputscale: float, noise: float):
        """Candidates of ``precision="auto"`` for a train size and an outputscale / noise ratio, fastest first, always ending in
        ``"fp64"`` (``constants.AUTO_I8_*``; the limits come from measured error constants)."""
        if not (AUTO_I8_MIN_N <= n <= AUTO_I8_MAX_N) or not (noise > 0.0 and outputscale > 0.0):
            return ["fp64"]
        ratio = outputscale / noise
        cands = []
        if nv.padded(n) <= AUTO_I8_WIDE_MAX_N:
            if ratio <= AUTO_I8_W5_MAX_SCALE_TO_NOISE:
                cands.append("int8x5w")
            if ratio <= AUTO_I8_W6_MAX_SCALE_TO_NOISE:
                cands.append("int8x6w")
        else:
            # 8-bit digits above 16384 rows: exact only if the digits of this L^-1 allow it (i8_sums_exact, checked where the
            # candidate is picked); five of them are 1.4 x cheaper than the six 7-bit digits that are always safe
            if ratio <= AUTO_I8_W5_MAX_SCALE_TO_NOISE:
                cands.append("int8x5w")
            if ratio <= AUTO_I8_MAX_SCALE_TO_NOISE:
                cands.append("int8x6")
        return cands + ["fp64"]

    def auto_precision(self, noisy: bool = True) -> str:
        """What ``model(x).variance`` / ``likelihood(model(x))`` use for the CURRENT factors: the first ``auto`` candidate -- checked
        against the FP64 path on the probe set outside training loops, by its ratio limit alone inside them (per-epoch metrics).
        The latent variance (``noisy=False``) has no noise floor under its cancellation and stays on the FP64 path."""
        if not noisy or not self.factored:
            return "fp64"
        key = self.n_factorizations
        if getattr(self, "_auto_key", None) != key:
            th = self.theta[self.D:].tolist()              # one small host read per factorisation
            self._auto_cands = self.auto_candidates(self.N, float(th[0]), float(th[1]))
            self._auto_key = key
        for c in self._auto_cands:
            if c == "fp64":
                return c
            if not self.i8_sums_exact(c):            # 8-bit digits above 16384 rows whose int32 sums this L^-1 does not bound
                continue
            if self.approx_inverse_ok or self.int8_probe_ok(c, AUTO_I8_PROBE_TOL):
                return c
        return "fp64"

    def _probe_points(self):
        """Probe set of the ``precision="auto"`` guard for the CURRENT factors: out of up to AUTO_I8_PROBE_POOL evenly strided
        training inputs, the AUTO_I8_PROBE_POINTS of smallest FP64 (DMMA) predictive std -- where s + noise - q cancels the most,
        i.e. where an absolute error of q costs the most relative accuracy -- and the same points moved by 1e-4 (the training
        inputs live in the unit cube) along a fixed pseudo-random direction, so near-duplicates of training inputs are covered
        too.  Returns (points, their FP64 std)."""
        if self._probe_set is not None and self._probe_set[0] == self.n_factorizations:
            return self._probe_set[1], self._probe_set[2]
        stride = max(1, -(-self.N // AUTO_I8_PROBE_POOL))
        pool = self.X[::stride]
        zero = torch.zeros(pool.shape[0], dtype=F64, device=self.device)
        _, s_pool = self.predict(pool, prior_mean=zero, precision="fp64")
        k = min(AUTO_I8_PROBE_POINTS, pool.shape[0])
        idx = torch.topk(s_pool, k, largest=False).indices
        base = pool.index_select(0, idx)
        g = torch.Generator(device="cpu").manual_seed(8)
        shift = (torch.rand(k, self.D, generator=g, dtype=F64) - 0.5).to(self.device) * 2e-4
        pts = torch.cat([base, base + shift], 0).contiguous()
        _, s64 = self.predict(pts, prior_mean=torch.zeros(2 * k, dtype=F64, device=self.device), precision="fp64")
        self._probe_set = (self.n_factorizations, pts, s64.clone())
        return pts, self._probe_set[2]

    def int8_probe_ok(self, precision: str, tol: float) -> bool:
        """Guard of ``precision="auto"``: True when the sliced-integer std of mode ``precision`` agrees with the FP64 (DMMA) std to
        ``tol`` on the probe set of ``_probe_points`` for the CURRENT factors.  Evaluated once per factorisation and mode (one
        FP64 predict of the pool, two small predicts and one host read), then cached."""
        hit = self._i8_probe.get(precision)
        if hit is not None and hit[0] == self.n_factorizations and hit[1] == float(tol):
            return hit[2]
        if not self.i8_sums_exact(precision):        # the mode's int32 sums are not bounded for these factors: not a candidate
            self._i8_probe[precision] = (self.n_factorizations, float(tol), False)
            return False
        pts, s64 = self._probe_points()
        _, s8 = self.predict(pts, prior_mean=torch.zeros(pts.shape[0], dtype=F64, device=self.device), precision=precision)
        ok = bool(((s8 - s64).abs() <= tol * s64).all())
        self._i8_probe[precision] = (self.n_factorizations, float(tol), ok)
        return ok

    def i8_sums_exact(self, precision: str) -> bool:
        """True when the int32 accumulators of sliced-integer mode ``precision`` cannot overflow for the CURRENT factors: always up to
        ``gpx_i8_safe_n(bits)`` (any digits); above it (8-bit digits, 16384 < N_pad <= 32768) when 128 x the largest row sum of
        |digits| of the sliced L^-1 stays below 2^31 (``gpx_i8_digit_bound``: one pass over the planes and one host read per
        factorisation and mode)."""
        slices, bits = I8_MODES[precision]
        if self.Np <= int(self.lib.gpx_i8_safe_n(bits)):
            return True
        if self.Np > int(self.lib.gpx_i8_max_n(bits)):
            return False
        key = (precision, self.n_factorizations, self.inverse_exact)
        hit = self._i8_bound.get(precision)
        if hit is not None and hit[0] == key:
            return hit[1]
        self._ensure_i8(slices, bits)
        out = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._launch("gpx_i8_digit_bound", nv.ptr(self.S_i8), slices, self.Np, nv.ptr(out), nv.stream_ptr())
        bound = int(out.item())
        ok = 128 * bound < 2 ** 31
        self._i8_bound[precision] = (key, ok, bound)
        return ok

    def _ensure_i8(self, slices: int, bits: int) -> None:
        """int8 slices and row scales of L^-1 for the sliced-integer variance path; refreshed after every factorisation."""
        if self.i8_mode == (slices, bits):
            return
        if self.S_i8 is None or self.S_i8.shape[0] != slices:
            self.S_i8 = None
            self.S_i8 = torch.empty(slices, self.Np, self.Np, dtype=torch.int8, device=self.device)
            self.S_i8_inv_scale = torch.empty(self.Np, dtype=F64, device=self.device)
        with torch.cuda.device(self.device):
            self._launch("gpx_slice_i8", nv.ptr(self.S), self.Np, self.Np, self.Np, nv.ptr(self.S_i8), self.Np, slices, bits, 1,
                         nv.ptr(self.S_i8_inv_scale), None, nv.stream_ptr())
        self.i8_mode = (slices, bits)

    # ------------------------------------------------------------------------------------------
    def cross_kernel(self, Xs: torch.Tensor, xa=None, xr=None) -> torch.Tensor:
        """Dense K*[M, N] (debug / with_covar / sample paths; test-set sized M)."""
        if Xs.dim() == 1:
            Xs = Xs.unsqueeze(-1)
        Xs = Xs.to(self.device, F64).contiguous()
        M = Xs.shape[0]
        Mc = nv.padded(M)
        Ks = torch.empty(Mc, self.Np, dtype=F64, device=self.device)
        mp = torch.empty(self.nb, Mc, dtype=F64, device=self.device)
        with torch.cuda.device(self.device):
            self._launch("gpx_kmat_cross", nv.ptr(Xs), M, nv.ptr(self.X), self.N, self.D, nv.ptr(self.theta),
                         nv.ptr(xa), nv.ptr(xr), self.kind, nv.ptr(self.alpha), nv.ptr(Ks), self.Np, nv.ptr(mp), Mc,
                         nv.stream_ptr())
        return Ks[:M, : self.N]

    def predict_covar(self, Xs: torch.Tensor, prior_mean: torch.Tensor, noisy: bool = True):
        """Posterior mean and the dense M x M predictive covariance for scaled device inputs Xs (test-set sized M):
        K** - (L^-1 K*^T)^T (L^-1 K*^T) [+ noise I], every product on the DMMA/TMA tile engine."""
        if not self.factored:
            raise RuntimeError("predict_covar() before factorize()")
        if not self.approx_inverse_ok:
            self.ensure_exact_inverse()
        if Xs.dim() == 1:
            Xs = Xs.unsqueeze(-1)
        Xs = Xs.detach().to(self.device, F64).contiguous()
        M = Xs.shape[0]
        Mp, Np, nb = nv.padded(M), self.Np, self.nb
        dev, p = self.device, nv.ptr
        Ks = torch.empty(Mp, Np, dtype=F64, device=dev)
        mpart = torch.empty(nb, Mp, dtype=F64, device=dev)
        Vt = torch.empty(Mp, Np, dtype=F64, device=dev)
        Kss = torch.empty(Mp, Mp, dtype=F64, device=dev)
        mpart2 = torch.empty(Mp // nv.TILE, Mp, dtype=F64, device=dev)
        zeros = torch.zeros(Mp, dtype=F64, device=dev)
        with torch.cuda.device(dev):
            st = nv.stream_ptr()
            self._launch("gpx_kmat_cross", p(Xs), M, p(self.X), self.N, self.D, p(self.theta), None, None, self.kind,
                         p(self.alpha), p(Ks), Np, p(mpart), Mp, st)
            # inside a training loop (the validation loss of every epoch) both products run on the integer tensor cores
            i8 = bool(self.approx_inverse_ok and self._trtri_i8_capable and Mp <= Np and Mp >= 1024 and self._lauum_ws is not None
                      and self._lauum_ws.numel() >= int(self.lib.gpx_trtri_i8_workspace_bytes(Np)))
            if not i8:
                self._launch("gpx_trmm_store", p(self.S), Np, p(Ks), Mp, p(Vt), st)
            self._launch("gpx_kmat_cross", p(Xs), M, p(Xs), M, self.D, p(self.theta_latent), None, None, self.kind,
                         p(zeros), p(Kss), Mp, p(mpart2), Mp, st)
            if i8:
                self._launch("gpx_predict_covar_i8", p(self.S), Np, p(Ks), Mp, p(Vt), p(Kss), p(self.W),
                             8 if Np <= int(self.lib.gpx_i8_safe_n(8)) else 7, p(self._lauum_ws), self._lauum_ws.numel(), st)
                self.has_kinv = False            # W served as scratch
            else:
                self._launch("gpx_syrk_sub", p(Vt), Mp, Np, p(Kss), st)
        mean = prior_mean.detach().to(dev, F64).reshape(-1) + mpart[:, :M].sum(0)
        cov = Kss[:M, :M]
        if i8:                                   # only the tiles on or below the diagonal were formed
            low = torch.tril(cov)
            cov = low + torch.tril(cov, -1).T
        else:
            cov = 0.5 * (cov + cov.T)
        if noisy:
            cov = cov + self.theta[self.D + 1] * torch.eye(M, dtype=F64, device=dev)
        return mean, cov

    def linv(self) -> torch.Tensor:
        """Dense lower-triangular L^-1 [N, N] (copy; for the small-M covariance paths and tests)."""
        if not self.approx_inverse_ok:
            self.ensure_exact_inverse()
        return torch.tril(self.S[: self.N, : self.N])

    def chol(self) -> torch.Tensor:
        return torch.tril(self.K[: self.N, : self.N])


class DenseCholesky:
    """Cholesky factor / inverse / log-determinant of a small dense SPD matrix (predictive covariances) through the
    same blocked kernels as the training factorisation, with gpytorch's jitter retry schedule."""

    def __init__(self, A: torch.Tensor, need_inverse: bool = False):
        nv.require_cuda()
        lib = nv.lib()
        dev = A.device
        M = A.shape[0]
        Mp = nv.padded(M)
        nb = Mp // nv.TILE
        self.M, self.Mp = M, Mp
        p = nv.ptr
        S = torch.empty(Mp, Mp, dtype=F64, device=dev)
        DT = torch.empty(nb, nv.TILE, nv.TILE, dtype=F64, device=dev)
        logdet_blk = torch.zeros(nb, dtype=F64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        jitter = 0.0
        with torch.cuda.device(dev):
            for attempt in range(4):
                K = torch.eye(Mp, dtype=F64, device=dev)
                K[:M, :M] = A
                if attempt > 0:
                    jitter = 1e-8 * 10 ** (attempt - 1)
                    warnings.warn(f"A not p.d., added jitter of {jitter:.1e} to the diagonal", NumericalWarning)
                    K[:M, :M].diagonal().add_(jitter)
                nv.check(lib.gpx_potrf(p(K), Mp, M, p(S), p(DT), p(logdet_blk), p(info), nv.stream_ptr()), "gpx_potrf")
                if int(info.item()) == 0:
                    break
                if attempt == 3:
                    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {jitter:.1e}.")
            if need_inverse:
                W = torch.empty(Mp, Mp, dtype=F64, device=dev)
                nv.check(lib.gpx_trtri(p(K), p(S), p(DT), p(W), Mp, nv.stream_ptr()), "gpx_trtri")
        self.K, self.S, self.DT = K, S, DT
        self.logdet_half = logdet_blk.sum()          # sum log L_ii
        self.has_inverse = need_inverse

    @property
    def L(self) -> torch.Tensor:
        return torch.tril(self.K[: self.M, : self.M])

    def solve_lower(self, b: torch.Tensor) -> torch.Tensor:
        """L^-1 b through the explicit triangular inverse (one mat-vec kernel)."""
        assert self.has_inverse
        lib = nv.lib()
        dev = self.K.device
        r = torch.zeros(self.Mp, dtype=F64, device=dev)
        r[: self.M] = b.reshape(-1)
        u = torch.empty(self.Mp, dtype=F64, device=dev)
        a = torch.empty(self.Mp, dtype=F64, device=dev)
        with torch.cuda.device(dev):
            nv.check(lib.gpx_solve_alpha(nv.ptr(self.S), nv.ptr(self.DT), self.Mp, nv.ptr(r), nv.ptr(u), nv.ptr(a),
                                         nv.stream_ptr()), "gpx_solve_alpha")
        return u[: self.M]


def make_theta(lengthscale: torch.Tensor, outputscale: torch.Tensor, noise: torch.Tensor,
               device: torch.device) -> torch.Tensor:
    """(1/lengthscale_d, outputscale, noise) as one device vector -- differentiable torch plumbing."""
    ls = lengthscale.reshape(-1).to(device, F64)
    return torch.cat([1.0 / ls, outputscale.reshape(1).to(device, F64), noise.reshape(1).to(device, F64)])
