"""CPU-only checks of the host side: API surface, constraints, scalers, dataset reproducibility, the
early-stopping protocol, the C-ABI export list -- and that the product path refuses to run without CUDA."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from gperks_b200 import _native
from gperks_b200.constraints import GreaterThan, Interval, Positive
from gperks_b200.gp.data.data_scaler import StandardLogScaler, StandardScaler, UnitCubeScaler
from gperks_b200.gp.data.dataset import Dataset
from gperks_b200.gp.experiment import GPExperiment, load_experiment_from_config_file
from gperks_b200.gp.mean import LinearMean
from gperks_b200.kernels import MaternKernel, RBFKernel, ScaleKernel
from gperks_b200.likelihoods import GaussianLikelihood
from gperks_b200.means import ConstantMean, LinearMean as GpLinearMean, ZeroMean
from gperks_b200.train.early_stop import (GLEarlyStoppingCriterion, NoEarlyStoppingCriterion,
                                          PkEarlyStoppingCriterion, PQEarlyStoppingCriterion,
                                          SimpleEarlyStoppingCriterion, UPEarlyStoppingCriterion)
from gperks_b200.train.emulator import GPEmulator
from gperks_b200.train.snapshot import EveryNEpochsSnapshottingCriterion, NeverSaveSnapshottingCriterion
from gperks_b200.train.train_stats import TrainStats, load_train_stats_from_file
from gperks_b200.utils.metrics import IndependentStandardError, MeanSquaredError, R2Score, get_metric_name
from gperks_b200.utils.polynomialfeatures import PolynomialFeatures
from gperks_b200.utils.test_functions import currin_exp, forrester
from oracle import gp_oracle as O

ROOT = Path(__file__).resolve().parent.parent


def _experiment(**kw):
    ds = Dataset.build_from_function(currin_exp, 2, 20, None, 25, "lhs", 8)
    return GPExperiment(ds, GaussianLikelihood(), LinearMean(2, 2), ScaleKernel(MaternKernel(ard_num_dims=2)),
                        n_restarts=2, seed=8, metrics=[MeanSquaredError(), R2Score()], **kw)


def test_cabi_exports_every_declared_symbol():
    header = (ROOT / "include" / "gpx.h").read_text()
    declared = set(re.findall(r"\b(gpx_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _native.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in gpx.h but not exported"
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    assert lib.gpx_version() >= 100
    assert lib.gpx_padded(1) == 128 and lib.gpx_padded(128) == 128 and lib.gpx_padded(129) == 256
    # argument validation happens before any launch, so it is callable without a GPU
    assert lib.gpx_kmat_sym(None, 10, 2, None, 0, None, 128, None) == -1
    assert lib.gpx_predict_workspace_bytes(256, 128) == 128 * 256 * 8 + 2 * 2 * 128 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    em = GPEmulator(_experiment(), "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        em.predict(np.random.rand(5, 2))
    opt = torch.optim.Adam(em.model.parameters(), lr=0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        em.train(opt)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        em.model.covar_module(torch.rand(3, 2)).to_dense()


def test_auto_precision_is_resolved_from_host_values():
    """precision="auto" (GPEmulator.predict's default) -> the sliced-integer path only where its int32 sums cannot overflow,
    the tiles pay and outputscale/noise leaves the slicing error inside 1e-6; host-side values only, no CUDA call."""
    from gperks_b200 import constants as K
    rng = np.random.default_rng(0)

    def emulator(n):
        X = rng.random((n, 3))
        exp = GPExperiment(Dataset(X, X.sum(1)), GaussianLikelihood(), LinearMean(1, 3), ScaleKernel(MaternKernel(ard_num_dims=3)),
                           n_restarts=1, seed=8)
        return GPEmulator(exp, "cpu")
    small, big = emulator(K.AUTO_I8_MIN_N - 1), emulator(K.AUTO_I8_MIN_N)
    assert small.resolve_precision("auto") == "fp64"
    big.model.covar_module.outputscale = 1.0
    big.model.likelihood.noise = 1e-2
    assert big.auto_precisions() == ["int8x5w", "int8x6w", "fp64"] and big.resolve_precision("auto") == "int8x5w"
    big.model.likelihood.noise = 1e-3
    big.model.covar_module.outputscale = 0.99e-3 * K.AUTO_I8_W5_MAX_SCALE_TO_NOISE
    assert big.resolve_precision("auto") == "int8x5w"
    big.model.covar_module.outputscale = 1.01e-3 * K.AUTO_I8_W5_MAX_SCALE_TO_NOISE       # five 8-bit slices no longer hold 1e-6
    assert big.auto_precisions() == ["int8x6w", "fp64"]
    big.model.covar_module.outputscale = 1.01e-3 * K.AUTO_I8_W6_MAX_SCALE_TO_NOISE       # outputscale / noise above every limit
    assert big.auto_precisions() == ["fp64"] and big.resolve_precision("auto") == "fp64"
    # 8-bit slices keep their int32 sums exact for ANY digits up to a padded train size of 16384; beyond (to 32768) five 8-bit
    # slices stay a candidate that the engine admits only after bounding the sums with the actual digits of L^-1
    # (ExactGPEngine.i8_sums_exact, a device pass: tests/test_gpu_baseline_shapes.py), followed by the always-safe six 7-bit slices
    assert K.AUTO_I8_WIDE_MAX_N == 16384 and K.AUTO_I8_MAX_N == 32768
    from gperks_b200.engine import ExactGPEngine
    assert ExactGPEngine.auto_candidates(26214, 1.0, 1e-2) == ["int8x5w", "int8x6", "fp64"]
    assert ExactGPEngine.auto_candidates(26214, 1.0, 1e-3) == ["int8x6", "fp64"]            # ratio 1000 > 300: five slices no longer hold 1e-6
    assert ExactGPEngine.auto_candidates(26214, 1.0, 1e-4) == ["fp64"]
    assert ExactGPEngine.auto_candidates(16384, 1.0, 1e-2) == ["int8x5w", "int8x6w", "fp64"]
    assert ExactGPEngine.auto_candidates(32769, 1.0, 1e-2) == ["fp64"]
    for p in ("fp64", "int8x5", "int8x6", "int8x5w", "int8x6w", "tf32x3", "tf32x3_fast"):
        assert big.resolve_precision(p) == p


def test_product_never_imports_oracle():
    for path in (ROOT / "gperks_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in text.replace("# oracle", ""), f"{path} mentions the oracle"


def test_state_dict_keys_match_reference_snapshots():
    exp = _experiment()
    keys = set(exp.model.state_dict().keys())
    assert keys == {
        "likelihood.noise_covar.raw_noise",
        "likelihood.noise_covar.raw_noise_constraint.lower_bound",
        "likelihood.noise_covar.raw_noise_constraint.upper_bound",
        "mean_module.weights", "mean_module.bias",
        "covar_module.raw_outputscale",
        "covar_module.raw_outputscale_constraint.lower_bound",
        "covar_module.raw_outputscale_constraint.upper_bound",
        "covar_module.base_kernel.raw_lengthscale",
        "covar_module.base_kernel.raw_lengthscale_constraint.lower_bound",
        "covar_module.base_kernel.raw_lengthscale_constraint.upper_bound",
    }
    sd = exp.model.state_dict()
    assert sd["covar_module.base_kernel.raw_lengthscale"].shape == (1, 2)
    assert sd["covar_module.raw_outputscale"].shape == ()
    assert sd["likelihood.noise_covar.raw_noise"].shape == (1,)
    assert sd["mean_module.weights"].shape == (5, 1) and sd["mean_module.bias"].shape == (1,)
    # a float32 state dict written by the reference loads into the float64 parameters
    ref_style = {k: v.float() for k, v in sd.items()}
    exp.model.load_state_dict(ref_style)


def test_constraints_and_default_values():
    lk = GaussianLikelihood()
    assert float(lk.noise.detach()) == pytest.approx(np.log(2.0) + 1e-4, rel=1e-12)
    k = ScaleKernel(RBFKernel(ard_num_dims=3))
    assert k.base_kernel.lengthscale.shape == (1, 3)
    assert float(k.outputscale.detach()) == pytest.approx(np.log(2.0))
    k.base_kernel.lengthscale = torch.tensor([0.1, 1.0, 10.0], dtype=torch.float64)
    np.testing.assert_allclose(k.base_kernel.lengthscale.detach().numpy().ravel(), [0.1, 1.0, 10.0], rtol=1e-12)
    lk.noise = 1e-3
    assert float(lk.noise.detach()) == pytest.approx(1e-3, rel=1e-10)
    # oracle agrees on the transform
    raw = torch.tensor(-3.6342, dtype=torch.float64)
    assert float(GreaterThan(1e-4).transform(raw)) == pytest.approx(float(O.constrained(raw, 1e-4)), rel=1e-14)
    c = Interval(0.5, 2.0)
    v = torch.tensor([0.7, 1.9], dtype=torch.float64)
    np.testing.assert_allclose(c.transform(c.inverse_transform(v)).numpy(), v.numpy(), rtol=1e-12)
    with pytest.raises(RuntimeError):
        MaternKernel(nu=2.0)


def test_fixed_noise_mode_and_initialize():
    exp = _experiment(learn_noise=False)
    em = GPEmulator(exp, "cpu")
    lk = em.model.likelihood
    assert float(lk.noise.detach()) == pytest.approx(1e-4, rel=1e-9)
    assert not lk.noise_covar.raw_noise.requires_grad
    assert float(lk.noise_covar.raw_noise_constraint.lower_bound) == pytest.approx(1e-6)
    em.model.initialize(**{"covar_module.base_kernel.raw_lengthscale": torch.tensor([0.3, -0.2]),
                           "covar_module.raw_outputscale": torch.tensor([1.5])})
    np.testing.assert_allclose(em.model.covar_module.base_kernel.raw_lengthscale.detach().numpy(), [[0.3, -0.2]])
    assert float(em.model.covar_module.raw_outputscale.detach()) == 1.5
    torch.manual_seed(0)
    em.restart_idx = 1
    em._reinit_hyperparameters()
    raw = em.model.covar_module.base_kernel.raw_lengthscale.detach()
    assert bool(((raw >= np.log(0.1)) & (raw <= np.log(10.0))).all())
    assert float(lk.noise.detach()) == pytest.approx(1e-4, rel=1e-9)   # untouched when the noise is fixed


def test_seeded_mean_init_matches_reference_float32_stream():
    torch.manual_seed(8)
    w32 = torch.randn(5, 1)
    b32 = torch.randn(1)
    torch.manual_seed(8)
    m = LinearMean(2, 2)
    assert torch.equal(m.weights.detach(), w32.double()) and torch.equal(m.bias.detach(), b32.double())


def test_polynomial_features_order_and_mean_forward():
    assert PolynomialFeatures(2).fit(2).indices == [[0], [1], [0, 0], [0, 1], [1, 1]]
    assert PolynomialFeatures(3).fit(2).indices[-4:] == [[0, 0, 0], [0, 0, 1], [0, 1, 1], [1, 1, 1]]
    m = LinearMean(2, 3)
    x = torch.rand(7, 3, dtype=torch.float64)
    ref = O.mean_fn(x, m.weights.detach(), m.bias.detach(), degree=2)
    np.testing.assert_allclose(m(x).detach().numpy(), ref.numpy(), rtol=1e-13)
    g = GpLinearMean(3)
    ref = O.mean_fn(x, g.weights.detach(), g.bias.detach(), degree=1)
    np.testing.assert_allclose(g(x).detach().numpy(), ref.numpy(), rtol=1e-13)
    assert ZeroMean()(x).abs().sum() == 0
    assert ConstantMean()(x).shape == (7,)
    idx, w, b = m.poly_spec(3)
    assert len(idx) == 9 and w.shape == (9,) and b.shape == (1,)


def test_scalers_follow_reference_conventions():
    X = np.array([[1.0, 10.0], [3.0, 30.0], [2.0, 50.0]])
    sc = UnitCubeScaler()
    sc.fit(X)
    np.testing.assert_allclose(sc.transform(X), [[0, 0], [1, 0.5], [0.5, 1]])
    np.testing.assert_allclose(sc.inverse_transform(sc.transform(X)), X)
    y = np.array([1.0, 2.0, 3.0, 6.0])
    s = StandardScaler()
    s.fit(y)
    assert s.sigma == pytest.approx(np.std(y))           # population std
    m, sd = s.inverse_transform(s.transform(y), ystd_=np.ones(4))
    np.testing.assert_allclose(m, y)
    np.testing.assert_allclose(sd, np.std(y))
    with pytest.raises(ValueError):
        ls = StandardLogScaler()
        ls.fit(y)
        ls.inverse_transform(ls.transform(y))


def test_dataset_reproduces_reference_sampling(goldens):
    g = goldens["example_2"]["data"]
    ds = Dataset.build_from_function(currin_exp, 2, 20, None, 25, "lhs", 8)
    np.testing.assert_array_equal(ds.X_train, np.array(g["X_train"]))
    np.testing.assert_array_equal(ds.X_test, np.array(g["X_test"]))
    np.testing.assert_allclose(ds.y_test, np.array(g["y_test"]), rtol=1e-15)
    g1 = goldens["example_1"]["data"]
    ds1 = Dataset.build_from_function(forrester, 1, 10, None, 10, "srs", 8)
    np.testing.assert_array_equal(ds1.X_train, np.array(g1["X_train"]))
    np.testing.assert_allclose(ds1.y_train, np.array(g1["y_train"]), rtol=1e-15)
    assert ds.input_size == 2 and ds.sample_size == 20 and not ds.with_val and ds.with_test
    assert ds.x_labels == ["X1", "X2"]


def test_experiment_config_round_trip(tmp_path):
    exp = _experiment()
    exp.metrics.append(IndependentStandardError(ci=3.0))
    f = tmp_path / "exp.ini"
    exp.save_to_config_file(f)
    exp2 = load_experiment_from_config_file(f, exp.dataset)
    assert exp2.n_restarts == 2 and exp2.seed == 8 and exp2.learn_noise
    assert [get_metric_name(m) for m in exp2.metrics] == ["MeanSquaredError", "R2Score", "IndependentStandardError"]
    assert exp2.metrics[2].ci == 3.0
    assert isinstance(exp2.model.covar_module.base_kernel, MaternKernel) and exp2.model.covar_module.base_kernel.nu == 2.5
    assert exp2.model.mean_module.degree == 2
    # configs written by the reference name gpytorch classes
    txt = f.read_text().replace("gperks_b200.kernels.MaternKernel", "gpytorch.kernels.matern_kernel.MaternKernel")
    f.write_text(txt)
    exp3 = load_experiment_from_config_file(f, exp.dataset)
    assert isinstance(exp3.model.covar_module.base_kernel, MaternKernel)


class _Model(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))


def _drive(crit, train_losses, val_losses=None):
    model, stats = _Model(), TrainStats([])
    crit.enable(model, stats)
    best = None
    for e, tl in enumerate(train_losses, 1):
        stats.current_epoch = e
        stats.train_loss.append(tl)
        if val_losses is not None:
            stats.val_loss.append(val_losses[e - 1])
        with torch.no_grad():
            model.p.fill_(float(e))
        best, _ = crit.evaluate()
        if crit.is_verified:
            break
    return best, stats, model


def test_early_stopping_protocol():
    best, stats, _ = _drive(NoEarlyStoppingCriterion(5), [5, 4, 3, 2, 1])
    assert best == 5 and stats.current_epoch == 5 and len(stats.early_stopping_criterion_evaluations) == 5
    # Simple: stop after `patience` epochs without validation improvement; best = epoch of the minimum
    best, stats, model = _drive(SimpleEarlyStoppingCriterion(100, patience=3), [1] * 10, [5, 4, 3, 3.5, 3.6, 3.7, 9, 9, 9, 9])
    assert stats.current_epoch == 6 and best == 3 and float(model.p.detach()) == 3.0
    # GL: generalisation loss above alpha for `patience` epochs
    best, stats, model = _drive(GLEarlyStoppingCriterion(100, alpha=10.0, patience=2), [1] * 8, [1.0, 0.9, 1.2, 1.3, 1.4, 1, 1, 1])
    assert stats.current_epoch == 4 and best == 2 and float(model.p.detach()) == 2.0
    # Pk: training progress stalls
    losses = [10, 5, 2, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]
    best, stats, _ = _drive(PkEarlyStoppingCriterion(100, alpha=0.01, patience=2, strip_length=3), losses)
    assert crit_stopped(stats, 7) and best == 5
    # UP: validation loss increases over successive strips
    best, stats, _ = _drive(UPEarlyStoppingCriterion(100, strip_length=2, successive_strips=2), [1] * 10, [5, 4, 3, 4, 5, 6, 7, 8, 9, 10])
    assert stats.current_epoch == 5 and best == 5 - (2 + 2 - 1)
    # PQ runs and records one evaluation per epoch
    best, stats, _ = _drive(PQEarlyStoppingCriterion(6, alpha=0.5, patience=3, strip_length=2), [3, 2.5, 2.2, 2.1, 2.05, 2.0], [3, 2.6, 2.7, 2.8, 2.9, 3.0])
    assert len(stats.early_stopping_criterion_evaluations) == stats.current_epoch


def crit_stopped(stats, epoch):
    return stats.current_epoch == epoch


def test_snapshot_paths_and_train_stats_io(tmp_path):
    crit = EveryNEpochsSnapshottingCriterion((tmp_path / "restart_{restart}").as_posix(), "epoch_{epoch}.pth", 2)
    model, stats = _Model(), TrainStats(["MeanSquaredError"])
    crit.enable(model, stats)
    for e in range(1, 6):
        stats.current_epoch = e
        crit.maybe_save(1, e)
    assert sorted(p.name for p in (tmp_path / "restart_1").iterdir()) == ["epoch_2.pth", "epoch_4.pth"]
    crit.keep_snapshots_until(1, 2)
    assert sorted(p.name for p in (tmp_path / "restart_1").iterdir()) == ["epoch_2.pth"]
    never = NeverSaveSnapshottingCriterion((tmp_path / "r_{restart}").as_posix(), "e_{epoch}.pth")
    never.enable(model, stats)
    never.maybe_save(1, 1)
    assert not (tmp_path / "r_1").exists()
    never.save(1, 3)                       # the training loop always saves the best model
    assert (tmp_path / "r_1" / "e_3.pth").exists()
    stats.train_loss = [1.0, 0.5]
    stats.train_metrics_score["MeanSquaredError"] = [0.3, 0.2]
    stats.save_to_file(tmp_path / "s.json")
    back = load_train_stats_from_file(tmp_path / "s.json")
    assert back.train_loss == [1.0, 0.5] and back.train_metrics_score["MeanSquaredError"] == [0.3, 0.2]


def test_metrics_shims():
    p = torch.tensor([1.0, 2.0, 3.0])
    t = torch.tensor([1.0, 2.5, 2.0])
    assert float(MeanSquaredError()(p, t)) == pytest.approx(O.mse(p.numpy(), t.numpy()))
    assert float(R2Score()(p, t)) == pytest.approx(O.r2score(p.numpy(), t.numpy()))
    assert float(IndependentStandardError()(p, torch.tensor([0.1, 1.0, 0.4]), t)) == pytest.approx(2 / 3)
    assert get_metric_name(R2Score()) == "R2Score"


@pytest.mark.parametrize("mean", ["poly2", "gp_linear", "constant", "zero"])
@pytest.mark.parametrize("scaled", [True, False])
def test_stacked_restart_parameters_match_per_replica_evaluation(mean, scaled):
    """The batched trainer's vectorised (theta, resid) over all replicas -- values and gradients -- against each replica's
    own ``_theta`` / ``_resid``; and that a replica's parameters are views of its row of the stacked tensors."""
    from copy import deepcopy
    from gperks_b200.models import ExactGP
    from gperks_b200.train.batched import _StackedParams
    torch.manual_seed(3)
    D, N, R = 3, 17, 4
    X, y = torch.rand(N, D, dtype=torch.float64), torch.randn(N, dtype=torch.float64)
    Xv = torch.rand(5, D, dtype=torch.float64)
    mm = {"poly2": lambda: LinearMean(2, D), "gp_linear": lambda: GpLinearMean(D), "constant": ConstantMean, "zero": ZeroMean}[mean]()
    base = MaternKernel(nu=1.5, ard_num_dims=D)

    class Model(ExactGP):
        def __init__(self):
            super().__init__(X, y, GaussianLikelihood())
            self.mean_module = mm
            self.covar_module = ScaleKernel(base) if scaled else base

        def forward(self, x):
            from gperks_b200.distributions import MultivariateNormal
            return MultivariateNormal(self.mean_module(x), self.covar_module(x))

    m0 = Model()
    replicas = []
    for r in range(R):
        m = deepcopy(m0)
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.3 * torch.randn_like(p))
        replicas.append(m)
    dev = torch.device("cpu")
    ref_theta = torch.stack([m._theta(dev) for m in replicas])
    ref_resid = torch.stack([m._resid(dev) for m in replicas])
    wt, wr = torch.randn_like(ref_theta), torch.randn_like(ref_resid)
    ((ref_theta * wt).sum() + (ref_resid * wr).sum()).backward()
    ref_grads = [{n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None} for m in replicas]
    ref_vprior = torch.stack([m.mean_module(Xv).reshape(-1) for m in replicas]).detach()

    st = _StackedParams.build(replicas, X, y, Xv, dev)
    assert st is not None
    idx = torch.arange(R)
    theta, resid = st.theta(idx), st.resid(idx)
    np.testing.assert_allclose(theta.detach().numpy(), ref_theta.detach().numpy(), rtol=1e-15, atol=0)
    np.testing.assert_allclose(resid.detach().numpy(), ref_resid.detach().numpy(), rtol=0, atol=1e-14)
    np.testing.assert_allclose(st.val_prior(idx).detach().numpy(), ref_vprior.numpy(), rtol=0, atol=1e-14)
    ((theta * wt).sum() + (resid * wr).sum()).backward()
    for r in range(R):
        for n, g in ref_grads[r].items():
            np.testing.assert_allclose(st.leaves[n].grad[r].numpy(), g.numpy(), rtol=1e-12, atol=1e-14)
    # subset of rows, and shared storage between a replica's parameter and its row
    sub = torch.tensor([2, 0])
    np.testing.assert_allclose(st.theta(sub).detach().numpy(), ref_theta.detach().numpy()[[2, 0]], rtol=1e-15)
    with torch.no_grad():
        for leaf in st.parameters():
            leaf.add_(1.0)
    for r, m in enumerate(replicas):
        for n, p in m.named_parameters():
            assert torch.equal(p.detach(), st.leaves[n].detach()[r])


def test_settings_shim_accepts_gpytorch_idioms():
    from gperks_b200 import settings
    with torch.no_grad(), settings.fast_pred_var(), settings.max_cholesky_size(4000), settings.cholesky_jitter(1e-6):
        pass
    with settings.fast_computations(covar_root_decomposition=False, log_prob=False, solves=False):
        assert settings.fast_pred_var.on() and not settings.fast_pred_var.off()


def test_worker_initializer_binds_a_device_and_shares_out_the_cores():
    """Pool workers of train(devices=...) / KFoldCrossValidation(devices=...): one device per worker for its whole life, and the
    host's cores split between the workers (eight workers at torch's default thread count contended for the same 16 threads and
    a config-5 fit took 2.2 s instead of 0.9 s)."""
    import multiprocessing
    import queue
    from gperks_b200.utils import concurrency as cc
    q = queue.Queue()
    q.put("cuda:3")
    n0 = torch.get_num_threads()
    try:
        cc._bind_worker(q, 4)
        assert cc._WORKER_DEVICE == "cuda:3"
        assert torch.get_num_threads() == max(1, multiprocessing.cpu_count() // 4)
        assert cc._call_on_worker_device(lambda a, b: (a, b), 1, ("x", "cuda:0")) == ("x", "cuda:3")
    finally:
        torch.set_num_threads(n0)
        cc._WORKER_DEVICE = None
