"""The DEFAULT predict path (``precision="auto"`` -> sliced-integer tensor-core variance) at BASELINE.json's own shapes against
an INDEPENDENT float64 checker: the oracle's formulas (oracle/gp_oracle.py: kernel forms, noisy predictive variance through a
triangular solve) evaluated with torch on the device, i.e. cuSOLVER potrf + cuBLAS trsm/gemm in FP64 -- none of this
package's kernels.  Bar: north_star's rel 1e-6 on mean and std (written in every assert), on points that sit ON training
inputs, 1e-4 beside them and far from them, down to the likelihood's noise floor of 1e-4.

Reference behaviour being matched: GPEmulator.predict, GPErks/train/emulator.py:348-363."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, assume, example, given, settings
from hypothesis import strategies as st

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import gp_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
F64 = torch.float64
KIND_OF = {"rbf": O.KIND_RBF, "m12": O.KIND_MATERN12, "m32": O.KIND_MATERN32, "m52": O.KIND_MATERN52}
RTOL = 1e-6          # north_star: "rel 1e-6 in FP64"


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _emulator(X, y, kind, ls, s, nz, w, b, learn_noise=True):
    from gperks_b200.constraints import GreaterThan
    from gperks_b200.gp.data.dataset import Dataset
    from gperks_b200.gp.experiment import GPExperiment
    from gperks_b200.gp.mean import LinearMean
    from gperks_b200.kernels import MaternKernel, RBFKernel, ScaleKernel
    from gperks_b200.likelihoods import GaussianLikelihood
    from gperks_b200.train.emulator import GPEmulator
    D = X.shape[1]
    base = {"rbf": lambda: RBFKernel(ard_num_dims=D), "m12": lambda: MaternKernel(nu=0.5, ard_num_dims=D),
            "m32": lambda: MaternKernel(nu=1.5, ard_num_dims=D), "m52": lambda: MaternKernel(nu=2.5, ard_num_dims=D)}[kind]()
    exp = GPExperiment(Dataset(X, y), GaussianLikelihood(), LinearMean(1, D), ScaleKernel(base), n_restarts=1, seed=8)
    if nz < 1e-4:
        exp.model.likelihood.noise_covar.register_constraint("raw_noise", GreaterThan(1e-9))
    em = GPEmulator(exp, "cuda:0")
    t = lambda v: torch.as_tensor(v, dtype=F64)  # noqa: E731
    em.model.likelihood.noise = float(nz)
    em.model.covar_module.outputscale = float(s)
    em.model.covar_module.base_kernel.lengthscale = t(ls)
    em.model.initialize(**{"mean_module.weights": t(w), "mean_module.bias": t([b])})
    return em


class DeviceChecker:
    """Exact-Cholesky GP in float64 with stock torch ops on the device (cuSOLVER / cuBLAS); scaled inputs, standardised targets."""

    def __init__(self, em, kind, dev):
        sd = em.scaled_data
        self.em, self.kind, self.dev = em, KIND_OF[kind], dev
        self.X = sd.X_train.to(dev, F64)
        m = em.model
        self.ls = m.covar_module.base_kernel.lengthscale.detach().reshape(-1).to(dev, F64)
        self.s = m.covar_module.outputscale.detach().to(dev, F64)
        self.nz = m.likelihood.noise.detach().reshape(()).to(dev, F64)
        self.w = m.mean_module.weights.detach().reshape(-1).to(dev, F64)
        self.b = m.mean_module.bias.detach().reshape(()).to(dev, F64)
        N = self.X.shape[0]
        K = torch.empty(N, N, dtype=F64, device=dev)
        for lo in range(0, N, 2048):
            K[lo:lo + 2048] = O.kernel_from_r2(O.sq_dist(self.X[lo:lo + 2048], self.X, self.ls), self.kind, self.s)
        K.diagonal().add_(self.nz)
        self.L = torch.linalg.cholesky(K)
        del K
        resid = sd.y_train.to(dev, F64) - (self.X @ self.w + self.b)
        self.alpha = torch.cholesky_solve(resid.reshape(-1, 1), self.L).reshape(-1)

    def predict(self, Xq_raw: np.ndarray, chunk: int = 4096):
        sd = self.em.scaled_data
        xs_all = torch.from_numpy(sd.scx.transform(Xq_raw)).to(self.dev, F64)
        mean = torch.empty(xs_all.shape[0], dtype=F64, device=self.dev)
        var = torch.empty_like(mean)
        for lo in range(0, xs_all.shape[0], chunk):
            xs = xs_all[lo:lo + chunk]
            Ks = O.kernel_from_r2(O.sq_dist(xs, self.X, self.ls), self.kind, self.s)
            mean[lo:lo + chunk] = xs @ self.w + self.b + Ks @ self.alpha
            V = torch.linalg.solve_triangular(self.L, Ks.T, upper=False)
            var[lo:lo + chunk] = (self.s + self.nz - (V * V).sum(0)).clamp_min(1e-10)
        sig, mu = float(sd.scy.sigma), float(sd.scy.mu)
        return (sig * mean + mu).cpu().numpy(), (sig * var.sqrt()).cpu().numpy()


def _query_points(X, M, rng):
    """M raw points: a quarter ON training inputs, a quarter 1e-4 beside them, the rest spread over (and slightly outside) the box."""
    D = X.shape[1]
    q = M // 4
    idx = rng.integers(0, X.shape[0], size=2 * q)
    Xq = rng.random((M, D)) * 1.1 - 0.05
    Xq[:q] = X[idx[:q]]
    Xq[q:2 * q] = X[idx[q:]] + 1e-4 * (rng.random((q, D)) - 0.5)
    return Xq


def _assert_matches(em, chk, Xq, expect_precision=None):
    picked = None
    if expect_precision is not None:
        from gperks_b200.constants import AUTO_I8_PROBE_TOL
        eng = em.model._factorized_engine()
        picked = next(c for c in em.auto_precisions() if c == "fp64" or eng.int8_probe_ok(c, AUTO_I8_PROBE_TOL))
        assert picked == expect_precision, (picked, expect_precision)
    mean, std = em.predict(Xq)                       # the default: precision="auto"
    m_ref, s_ref = chk.predict(Xq)
    np.testing.assert_allclose(std, s_ref, rtol=RTOL)
    np.testing.assert_allclose(mean, m_ref, rtol=RTOL, atol=RTOL * float(np.abs(m_ref).max()))
    return float(np.max(np.abs(std / s_ref - 1.0))), picked


# ---- BASELINE configs[1]: N=8192, D=8, Matern-5/2 (the headline), and the same at the likelihood's noise floor --------------
@pytest.mark.parametrize("noise,want", [(1e-2, "int8x5w"), (1e-4, "int8x6w")])
def test_default_predict_at_config2_shape(dev, noise, want):
    from bench import synthetic_problem
    X, y, _, hyper = synthetic_problem(n_train=8192, dim=8, m=0)
    em = _emulator(X, y, "m52", hyper["lengthscale"], hyper["outputscale"], noise, hyper["weights"], hyper["bias"])
    chk = DeviceChecker(em, "m52", dev)
    Xq = _query_points(X, 100_000, np.random.default_rng(1))
    worst, _ = _assert_matches(em, chk, Xq, expect_precision=want)
    print(f"config2 noise={noise}: {want} worst rel std error {worst:.2e} over 1e5 points")


# ---- BASELINE configs[2]: N=4096, D=12 (the GSA emulator) ---------------------------------------------------------------
@pytest.mark.parametrize("kind,noise,want", [("m52", 1e-2, "int8x5w"), ("rbf", 1e-3, "int8x6w")])
def test_default_predict_at_config3_shape(dev, kind, noise, want):
    from bench import synthetic_problem
    X, y, _, hyper = synthetic_problem(n_train=4096, dim=12, m=0)
    em = _emulator(X, y, kind, hyper["lengthscale"], hyper["outputscale"], noise, hyper["weights"], hyper["bias"])
    chk = DeviceChecker(em, kind, dev)
    Xq = _query_points(X, 100_000, np.random.default_rng(2))
    worst, _ = _assert_matches(em, chk, Xq, expect_precision=want)
    print(f"config3 {kind} noise={noise}: {want} worst rel std error {worst:.2e}")


# ---- BASELINE configs[3]: 8 output emulators, N=2048, D=6, through Wave.compute_impl ---------------------------------
def test_history_matching_wave_at_config4_shape(dev):
    from bench import synthetic_problem
    from gperks_b200.perks.history_matching import Wave
    rng = np.random.default_rng(3)
    E, N, D, M = 8, 2048, 6, 60_000
    ems, chks, X0 = [], [], None
    for e in range(E):
        X, y, _, hyper = synthetic_problem(n_train=N, dim=D, m=0, seed=20 + e)
        X0 = X if X0 is None else X0
        kind = ("m52", "rbf", "m32", "m52")[e % 4]
        em = _emulator(X, y + 0.1 * e, kind, [l * (1 + 0.1 * e) for l in hyper["lengthscale"]], 0.5 + 0.25 * e, 1e-2 / (1 + e),
                       hyper["weights"], hyper["bias"])
        ems.append(em)
        chks.append(DeviceChecker(em, kind, dev))
    Xq = _query_points(X0, M, rng)
    z = np.array([0.3 + 0.05 * e for e in range(E)])
    v = np.array([0.01 * (1 + e) for e in range(E)])
    W = Wave(emulator=ems, Itrain=np.array([[0.0, 1.0]] * D), cutoff=3.0, maxno=2, mean=z, var=v)
    I, PV = W.compute_impl(Xq)
    mean = np.empty((M, E))
    var = np.empty((M, E))
    for e, chk in enumerate(chks):
        m_, s_ = chk.predict(Xq)
        mean[:, e], var[:, e] = m_, s_ ** 2
    In = np.sqrt((mean - z) ** 2 / (var + v))              # GPErks/perks/history_matching.py:57-62
    I_ref = np.sort(In, axis=1)[:, -W.maxno]
    PV_ref = np.sort(var / v, axis=1)[:, -W.maxno]
    np.testing.assert_allclose(I, I_ref, rtol=2 * RTOL, atol=2 * RTOL * float(np.abs(mean - z).max()))
    np.testing.assert_allclose(PV, PV_ref, rtol=2 * RTOL)           # variance = std^2: twice the relative error of std
    keep = np.where(I < W.cutoff)[0]
    W.find_regions(Xq)
    assert np.array_equal(np.asarray(W.nimp_idx), keep) and W.NIMP.shape[0] + W.IMP.shape[0] == M


# ---- BASELINE configs[4]: one fold of N=32768, D=16 (26214 training points, RBF) ----------------------------------------
def test_default_predict_at_config5_fold_shape(dev):
    from bench import synthetic_problem
    X, y, _, hyper = synthetic_problem(n_train=26214, dim=16, m=0)
    em = _emulator(X, y, "rbf", hyper["lengthscale"], hyper["outputscale"], 1e-2, hyper["weights"], hyper["bias"])
    chk = DeviceChecker(em, "rbf", dev)
    Xq = _query_points(X, 20_000, np.random.default_rng(4))
    # padded size 26240 > 16384: five 8-bit digits are used because the digits of THIS L^-1 bound every int32 accumulator
    # (ExactGPEngine.i8_sums_exact: 128 x the largest row sum of |digits| < 2^31); otherwise it would be six 7-bit ones
    eng0 = em.model._factorized_engine()
    assert eng0.Np > int(eng0.lib.gpx_i8_safe_n(8)) and eng0.i8_sums_exact("int8x5w") and eng0.i8_sums_exact("int8x6")
    bound = eng0._i8_bound["int8x5w"][2]
    planes = eng0.S_i8
    rows = torch.randint(0, eng0.N, (64,), device=dev)
    live = torch.arange(eng0.Np, device=dev).view(1, -1) < ((rows // 128 + 1) * 128).view(-1, 1)   # the k range read per row
    ref = (planes[:, rows, :].to(torch.int32).abs() * live.unsqueeze(0)).sum(dim=(0, 2)).max()
    assert 0 < int(ref) <= bound < 2 ** 24                                  # the kernel's maximum covers the sampled rows
    worst, picked = _assert_matches(em, chk, Xq, expect_precision="int8x5w")
    print(f"config5 fold: {picked}, digit bound {bound} (limit {2 ** 24}), worst rel std error {worst:.2e}")
    _, s6 = em.predict(Xq[:4096], precision="int8x6")                       # the always-safe 7-bit mode still agrees
    _, s5 = em.predict(Xq[:4096], precision="int8x5w")
    np.testing.assert_allclose(s5, s6, rtol=2e-7)
    # the training objective at this size against the same independent factorisation
    eng = em.model._factorized_engine()
    N = 26214
    resid = em.scaled_data.y_train.to(dev, F64) - (chk.X @ chk.w + chk.b)
    loss = (0.5 * resid @ chk.alpha + torch.log(torch.diagonal(chk.L)).sum() + 0.5 * N * np.log(2 * np.pi)) / N
    assert float(eng.loss) == pytest.approx(float(loss), rel=1e-9)


# ---- property sweep over hyper-parameters at N >= 1024 (where "auto" leaves the FP64 path) -------------------------------
_CACHE = {}


def _sweep_data(N, D):
    key = (N, D)
    if key not in _CACHE:
        rng = np.random.default_rng(N * 31 + D)
        X = rng.random((N, D))
        y = np.sin(3 * X[:, 0]) + (X[:, 1 % D] - 0.5) ** 2 + 0.1 * rng.standard_normal(N)
        _CACHE[key] = (X, y, _query_points(X, 3000, rng))
    return _CACHE[key]


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(kind=st.sampled_from(["rbf", "m12", "m32", "m52"]), N=st.sampled_from([1024, 1100, 1536]), D=st.integers(2, 8),
       log_lmin=st.floats(-1.3, 0.3), log_ratio=st.floats(0.0, 3.0), log_s=st.floats(-1.5, 1.3), log_nz=st.floats(-4.0, -1.0),
       seed=st.integers(0, 1000))
# the reference's own trained hyper-parameters of notebooks/example_9.ipynb: lengthscales 0.056 vs 101.7, outputscale 0.047, noise 1e-4
@example(kind="m52", N=1100, D=2, log_lmin=np.log10(0.056), log_ratio=np.log10(101.7 / 0.056), log_s=np.log10(0.047), log_nz=-4.0, seed=9)
@example(kind="rbf", N=1536, D=8, log_lmin=-1.3, log_ratio=3.0, log_s=1.0, log_nz=-4.0, seed=1)
@example(kind="m12", N=1024, D=3, log_lmin=0.3, log_ratio=0.0, log_s=-1.5, log_nz=-1.0, seed=2)
def test_auto_precision_property_sweep(dev, kind, N, D, log_lmin, log_ratio, log_s, log_nz, seed):
    """ARD length-scale ratios up to 1e3, outputscale 0.03 .. 20, noise from the floor (1e-4) up, all four kernel families: whatever
    ``auto`` picks (int8x5w / int8x6w / fp64, guarded by the probe) must stay within rel 1e-6 of the independent FP64 checker."""
    assume(log_s - log_nz <= 5.0)        # beyond outputscale/noise = 1e5 two independent FP64 evaluations differ by ~1e-6 themselves
    X, y, Xq = _sweep_data(N, D)
    rng = np.random.default_rng(seed)
    ls = 10.0 ** (log_lmin + log_ratio * rng.random(D))
    ls[0], ls[-1] = 10.0 ** log_lmin, 10.0 ** (log_lmin + log_ratio)
    em = _emulator(X, y, kind, ls, 10.0 ** log_s, 10.0 ** log_nz, np.linspace(-0.5, 0.5, D), 0.05)
    chk = DeviceChecker(em, kind, dev)
    _assert_matches(em, chk, Xq)


def test_auto_precision_after_train_auto(dev):
    """Hyper-parameters that come out of the package's own L-BFGS-B fit (train_auto) at a size where auto uses the integer path."""
    import tempfile
    X, y, Xq = _sweep_data(1100, 4)
    em = _emulator(X, y, "m52", [0.5] * 4, 1.0, 1e-2, [0.0] * 4, 0.0)
    em.train_auto(tempfile.mkdtemp(prefix="gpx_auto_"), maxiter=60)
    chk = DeviceChecker(em, "m52", dev)
    worst, _ = _assert_matches(em, chk, Xq)
    nz = float(em.model.likelihood.noise)
    s = float(em.model.covar_module.outputscale)
    print(f"train_auto: outputscale/noise = {s / nz:.3g}, candidates {em.auto_precisions()}, worst rel std error {worst:.2e}")


def test_probe_set_is_where_the_variance_cancels_most(dev):
    """The guard of precision='auto' probes the training inputs of smallest FP64 predictive std (+ 1e-4 neighbours), not the
    first rows of X."""
    X, y, _ = _sweep_data(1536, 5)
    em = _emulator(X, y, "m52", [0.3, 0.5, 0.8, 1.2, 2.0], 1.0, 1e-3, [0.0] * 5, 0.0)
    eng = em.model._factorized_engine()
    pts, s64 = eng._probe_points()
    zero = torch.zeros(eng.N, dtype=F64, device=dev)
    _, s_all = eng.predict(eng.X, prior_mean=zero, precision="fp64")
    k = pts.shape[0] // 2
    assert k == 128
    kth = torch.kthvalue(s_all, k).values
    assert float(s64[:k].max()) <= float(kth) * (1 + 1e-12)            # exactly the k smallest of the pool (pool = all rows here)
    assert float((pts[k:] - pts[:k]).abs().max()) <= 1e-4 and float((pts[k:] - pts[:k]).abs().max()) > 0
    assert not torch.equal(pts[:k], eng.X[:k])
