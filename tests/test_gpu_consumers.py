"""GPU tests of the consumers that sit on the hot path: history matching, Sobol' GSA,
K-fold cross validation, Inference / Diagnostics -- against the reference's formulas restated in numpy, the
notebook goldens and analytic Sobol' indices."""
import json

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _fit(f, d, n_train, kernel="matern", seed=8, n_test=None, l_bounds=None, u_bounds=None, auto=True, degree=1):
    from gperks_b200.gp.data.dataset import Dataset
    from gperks_b200.gp.experiment import GPExperiment
    from gperks_b200.gp.mean import LinearMean
    from gperks_b200.kernels import MaternKernel, RBFKernel, ScaleKernel
    from gperks_b200.likelihoods import GaussianLikelihood
    from gperks_b200.train.emulator import GPEmulator
    from gperks_b200.utils.metrics import IndependentStandardError, MeanSquaredError, R2Score
    ds = Dataset.build_from_function(f, d, n_train, None, n_test, "lhs", seed, l_bounds=l_bounds, u_bounds=u_bounds)
    base = MaternKernel(ard_num_dims=d) if kernel == "matern" else RBFKernel(ard_num_dims=d)
    exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(degree, d), ScaleKernel(base), n_restarts=1, seed=seed,
                       metrics=[MeanSquaredError(), R2Score(), IndependentStandardError()])
    em = GPEmulator(exp, "cuda:0")
    if auto:
        em.train_auto(snap_dir="/tmp/gperks_b200_test_snap")
    return em, ds


def test_history_matching_matches_reference_loop(dev):
    from gperks_b200.perks.history_matching import Wave
    from gperks_b200.utils.test_functions import branin_rescaled, currin_exp, lim_poly
    emulators = [_fit(f, 2, 40, seed=s)[0] for f, s in ((currin_exp, 8), (lim_poly, 9), (branin_rescaled, 10))]
    rng = np.random.default_rng(0)
    X = rng.random((5003, 2))
    z, v = [5.0, 8.0, -0.2], [0.5, 1.0, 0.01]
    for maxno in (1, 2, 3):
        W = Wave(emulator=emulators, Itrain=np.array([[0, 1], [0, 1]]), cutoff=3.0, maxno=maxno, mean=z, var=v)
        I, PV = W.compute_impl(X, block_points=2048)
        # the reference's arithmetic (history_matching.py:46-62), on the same predictions
        M = np.zeros((X.shape[0], 3))
        V = np.zeros((X.shape[0], 3))
        for j, em in enumerate(emulators):
            mean, std = em.predict(X)
            M[:, j], V[:, j] = mean, np.power(std, 2)
        Iref = np.array([np.sort(np.sqrt(np.power(M[i] - np.array(z), 2) / (V[i] + np.array(v))))[-maxno] for i in range(X.shape[0])])
        PVref = np.array([np.sort(V[i] / np.array(v))[-maxno] for i in range(X.shape[0])])
        assert np.array_equal(I, Iref) and np.array_equal(PV, PVref)      # byte-identical
    W = Wave(emulator=emulators, Itrain=np.array([[0.0, 1.0], [0.0, 1.0]]), cutoff=3.0, maxno=1, mean=z, var=v)
    W.find_regions(X)
    assert len(W.nimp_idx) + len(W.imp_idx) == X.shape[0]
    assert np.array_equal(W.reconstruct_tests(), X)
    if 10 < len(W.nimp_idx) < 4000:
        n0 = len(W.nimp_idx)
        np.random.seed(0)
        W.augment_nimp(n0 + 50)
        assert W.NIMP.shape[0] == n0 + 50 and len(W.I) == X.shape[0] + 50
        assert (W.I[-50:] < W.cutoff).all()


def test_multi_emulator_wave_equals_per_emulator_loop(dev):
    """gpx_wave_i8 (all emulators of a wave through ONE two-stream chunk pipeline + the implausibility epilogue, per-emulator
    predictions kept on chip) against the per-emulator predict_device loop + gpx_implausibility: byte-identical I / PV, with
    emulators of different train sizes, kernel families and slice modes; kept indices from the device compaction ==
    np.where(I < cutoff).  Replaces GPErks/perks/history_matching.py:43-77."""
    from gperks_b200.gp.data.dataset import Dataset
    from gperks_b200.gp.experiment import GPExperiment
    from gperks_b200.gp.mean import LinearMean
    from gperks_b200.kernels import MaternKernel, RBFKernel, ScaleKernel
    from gperks_b200.likelihoods import GaussianLikelihood
    from gperks_b200.perks.history_matching import Wave
    from gperks_b200.train.emulator import GPEmulator
    rng = np.random.default_rng(4)
    D = 3
    ems = []
    for e, (n, noise) in enumerate(((1100, 1e-2), (1300, 1e-4), (2048, 3e-3), (1024, 1e-2))):
        X = rng.random((n, D))
        y = np.sin(3 * X[:, 0] + e) + X[:, 1] * X[:, 2] + 0.05 * rng.standard_normal(n)
        base = MaternKernel(nu=(2.5, 1.5, 0.5, 2.5)[e], ard_num_dims=D) if e != 2 else RBFKernel(ard_num_dims=D)
        exp = GPExperiment(Dataset(X, y), GaussianLikelihood(), LinearMean(1 + (e % 2), D), ScaleKernel(base), n_restarts=1, seed=8)
        em = GPEmulator(exp, "cuda:0")
        em.model.covar_module.base_kernel.lengthscale = torch.tensor([0.4, 0.7, 1.1], dtype=torch.float64)
        em.model.covar_module.outputscale = 0.8 + 0.3 * e
        em.model.likelihood.noise = noise
        ems.append(em)
    assert [em.auto_precisions()[0] for em in ems] == ["int8x5w", "int8x6w", "int8x6w", "int8x5w"]
    Xq = rng.random((70_001, D)) * 1.1 - 0.05
    z, v = [0.2, 0.5, -0.1, 0.4], [0.05, 0.02, 0.1, 0.03]
    for maxno in (1, 3):
        W = Wave(emulator=ems, Itrain=np.array([[0.0, 1.0]] * D), cutoff=2.5, maxno=maxno, mean=z, var=v)
        assert W._wave_descriptors(dev)[0] is not None                  # the fused path is the one that runs
        I, PV = W.compute_impl(Xq, block_points=40_000)                 # two blocks: the pinned ring and the chunk tail
        W2 = Wave(emulator=ems, Itrain=W.Itrain, cutoff=2.5, maxno=maxno, mean=z, var=v)
        W2._wave_descriptors = lambda dev_: (None, None)                # force the per-emulator loop
        I2, PV2 = W2.compute_impl(Xq, block_points=40_000)
        assert np.array_equal(I, I2) and np.array_equal(PV, PV2)
        W.find_regions(Xq, block_points=40_000)
        assert np.array_equal(np.asarray(W.nimp_idx), np.where(W.I < W.cutoff)[0])
        assert np.array_equal(W.I, I) and len(W.nimp_idx) + len(W.imp_idx) == Xq.shape[0]
        assert np.array_equal(W.reconstruct_tests(), Xq)


@pytest.mark.parametrize("M", [0, 1, 255, 4095, 4096, 4097, 1_000_003])
def test_compact_below_c_abi(dev, M):
    from gperks_b200 import _native as nv
    lib = nv.lib()
    g = torch.Generator(device="cpu").manual_seed(M)
    vals = torch.rand(max(M, 1), generator=g, dtype=torch.float64).to(dev)
    for cutoff in (-1.0, 0.3, 2.0):
        idx = torch.full((max(M, 1),), -7, dtype=torch.int64, device=dev)
        cnt = torch.full((1,), -1, dtype=torch.int64, device=dev)
        ws = torch.empty(int(lib.gpx_compact_below_workspace_bytes(M)), dtype=torch.uint8, device=dev)
        nv.check(lib.gpx_compact_below(nv.ptr(vals), M, cutoff, nv.ptr(idx), nv.ptr(cnt), nv.ptr(ws), ws.numel(), nv.stream_ptr()), "compact")
        ref = torch.nonzero(vals[:M] < cutoff).reshape(-1)
        assert int(cnt) == ref.numel()
        assert torch.equal(idx[: ref.numel()], ref)


@pytest.mark.parametrize("E,maxno", [(1, 1), (3, 2), (4, 1), (5, 5), (8, 1), (8, 3), (8, 8), (11, 2), (13, 7)])
def test_implausibility_c_abi_is_byte_identical_to_numpy(dev, E, maxno):
    """gpx_implausibility for every code path of the kernel (blocks of four emulators + remainder, every list length),
    against the reference's arithmetic (GPErks/perks/history_matching.py:46-62) written with numpy."""
    from gperks_b200 import _native as nv
    rng = np.random.default_rng(100 * E + maxno)
    M = 3001
    means = rng.standard_normal((E, M))
    stds = rng.random((E, M)) + 0.05
    stds[:, ::97] = 0.0                                # exact zeros in the variance
    means[0, ::101] = means[E - 1, ::101]              # ties
    z = rng.standard_normal(E)
    v = rng.random(E) + 0.01
    d_m, d_s = torch.as_tensor(means, device=dev), torch.as_tensor(stds, device=dev)
    d_z, d_v = torch.as_tensor(z, device=dev), torch.as_tensor(v, device=dev)
    Id = torch.empty(M, dtype=torch.float64, device=dev)
    PVd = torch.empty(M, dtype=torch.float64, device=dev)
    nv.check(nv.lib().gpx_implausibility(nv.ptr(d_m), nv.ptr(d_s), M, E, nv.ptr(d_z), nv.ptr(d_v), maxno, nv.ptr(Id), nv.ptr(PVd),
                                         nv.stream_ptr()), "gpx_implausibility")
    V = np.power(stds, 2)
    Iref = np.sort(np.sqrt(np.power(means - z[:, None], 2) / (V + v[:, None])), axis=0)[-maxno]
    PVref = np.sort(V / v[:, None], axis=0)[-maxno]
    assert np.array_equal(Id.cpu().numpy(), Iref) and np.array_equal(PVd.cpu().numpy(), PVref)
    assert nv.lib().gpx_implausibility(nv.ptr(d_m), nv.ptr(d_s), M, E, nv.ptr(d_z), nv.ptr(d_v), E + 1, nv.ptr(Id), nv.ptr(PVd),
                                       nv.stream_ptr()) < 0


def test_sobol_gsa_against_analytic_indices(dev):
    from functools import partial
    from gperks_b200.perks.gsa import SobolGSA
    from gperks_b200.utils.test_functions_gsa import SobolGstar, SobolGstar_theoretical_Si
    d = 4
    a, delta, alpha = [0.0, 1.0, 4.5, 99.0], [0.1, 0.3, 0.5, 0.7], [2.0] * d
    f = partial(SobolGstar, a=a, delta=delta, alpha=alpha)
    em, ds = _fit(f, d, 300, l_bounds=[0.0] * d, u_bounds=[1.0] * d)
    ST_th, S1_th, _ = SobolGstar_theoretical_Si(a, delta, alpha)
    gsa = SobolGSA(ds, n=2 ** 15, seed=8)
    gsa.estimate_Sobol_indices_with_emulator_mean(em)          # 196608 predictions + sums on the device
    assert gsa.S1.shape == (1, d) and gsa.ST.shape == (1, d)
    np.testing.assert_allclose(gsa.S1[0], S1_th["Si"].values, atol=0.08)
    np.testing.assert_allclose(gsa.ST[0], ST_th["STi"].values, atol=0.08)
    # the device pipeline equals scipy's estimator applied to the same predictions
    _, _, X = gsa.saltelli_design()
    mean, _ = em.predict(X)
    from scipy.stats import sobol_indices
    n = gsa.n
    f_AB = np.moveaxis(mean[2 * n:].reshape((-1, n, d)), -1, 0)
    ref = sobol_indices(func={"f_A": mean[:n].reshape(1, -1), "f_B": mean[n:2 * n].reshape(1, -1), "f_AB": f_AB}, n=n)
    np.testing.assert_allclose(gsa.S1[0], np.ravel(ref.first_order), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gsa.ST[0], np.ravel(ref.total_order), rtol=1e-9, atol=1e-12)
    # reference-style path: joint posterior draws at n(d+2) points
    small = SobolGSA(ds, n=64, seed=8)
    small.estimate_Sobol_indices_with_emulator(em, n_draws=40)
    assert small.S1.shape == (40, d) and small.ST.shape == (40, d)
    assert np.argmax(np.median(small.ST, axis=0)) == 0            # X1 dominates (a_1 = 0)
    small.estimate_Sobol_indices_with_emulator(em, n_draws=10, chunk=100)
    assert small.ST.shape == (10, d)
    small.correct_Sobol_indices()
    small.assemble_dataframe()
    assert set(small.df.columns) == {"Parameter", "Index", "Value"}


def test_sobol_gsa_marginal_draws_on_device(dev):
    """Per-draw indices (n_draws, d) from marginal posterior-predictive draws, reduced on the device (gpx_saltelli_partial):
    deterministic for a seed, different across seeds, centred on the indices of the posterior mean, every row a valid set of
    indices; and the mean-only predict kernel equals the mean of the full predict bit for bit."""
    from functools import partial
    from gperks_b200.perks.gsa import SobolGSA
    from gperks_b200.utils.test_functions_gsa import SobolGstar
    d = 3
    f = partial(SobolGstar, a=[0.0, 1.0, 9.0], delta=[0.2, 0.4, 0.6], alpha=[1.0] * d)
    em, ds = _fit(f, d, 200, l_bounds=[0.0] * d, u_bounds=[1.0] * d)
    g0 = SobolGSA(ds, n=2 ** 14, seed=8)
    g0.estimate_Sobol_indices_with_emulator_device(em, n_draws=0)
    g = SobolGSA(ds, n=2 ** 14, seed=8)
    g.estimate_Sobol_indices_with_emulator_device(em, n_draws=200, draw_seed=5, block_rows=5000)    # ragged blocks, 4 chunks of 4096 rows
    assert g.S1.shape == (200, d) and g.ST.shape == (200, d)
    S1a, STa = g.S1.copy(), g.ST.copy()
    g.estimate_Sobol_indices_with_emulator_device(em, n_draws=200, draw_seed=5, block_rows=16384)
    assert np.array_equal(S1a, g.S1) and np.array_equal(STa, g.ST)              # blocks never change a bit, draws are keyed by (seed, row, draw)
    g.estimate_Sobol_indices_with_emulator_device(em, n_draws=200, draw_seed=6)
    assert not np.array_equal(S1a, g.S1)
    assert np.all(np.isfinite(S1a)) and np.all(STa > -1e-3) and STa.std(axis=0).max() > 0
    # marginal noise adds variance to f: indices shrink towards zero by var_f / (var_f + mean noise^2), slightly; centre stays close
    np.testing.assert_allclose(np.median(STa, axis=0), g0.ST[0], atol=0.03)
    np.testing.assert_allclose(np.median(S1a, axis=0), g0.S1[0], atol=0.03)
    # mean-only kernel == mean of the full predict
    x = torch.rand(5000, d, dtype=torch.float64, device=dev)
    m_full, _, _ = em.predict_device(x)
    m_only, _ = em.predict_mean_device(x)
    assert torch.equal(m_full, m_only)


def test_saltelli_sums_c_abi(dev):
    """gpx_saltelli_partial / gpx_saltelli_reduce against numpy on known inputs: the noise-free sums exactly (up to summation
    order), and the Philox normals through their first two moments and their independence across points."""
    from gperks_b200 import _native as nv
    lib = nv.lib()
    n, d = 10_000, 5
    JC = lib.gpx_saltelli_chunk_rows()
    rng = np.random.default_rng(0)
    fA, fB, fAB = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal((n, d))
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)  # noqa: E731
    dA, dB, dAB = t(fA), t(fB), t(fAB)
    n_chunks = (n + JC - 1) // JC
    NS = 4 + 3 * d
    part = torch.zeros(n_chunks * NS, dtype=torch.float64, device=dev)
    sums = torch.zeros(NS, dtype=torch.float64, device=dev)
    nv.check(lib.gpx_saltelli_partial(nv.ptr(dA), nv.ptr(dB), nv.ptr(dAB), None, None, None, n, d, 0, n, 0, 1, nv.ptr(part), nv.stream_ptr()), "partial")
    nv.check(lib.gpx_saltelli_reduce(nv.ptr(part), n_chunks, d, 0, nv.ptr(sums), nv.stream_ptr()), "reduce")
    D = fAB - fA[:, None]
    ref = np.concatenate([[fA.sum(), fB.sum(), (fA ** 2).sum(), (fB ** 2).sum()],
                          np.stack([(fB[:, None] * D).sum(0), D.sum(0), (D ** 2).sum(0)], 1).ravel()])
    np.testing.assert_allclose(sums.cpu().numpy(), ref, rtol=1e-12, atol=1e-10)
    # unit normals: mean 0, std 1 everywhere -> sum f_A^2 / n -> 1, sum f_A / n -> 0, sum f_B D_p / n -> 0 (independence), sum D^2 / n -> 2
    K = 64
    z0, o1 = torch.zeros(n, dtype=torch.float64, device=dev), torch.ones(n, dtype=torch.float64, device=dev)
    z0d, o1d = torch.zeros(n, d, dtype=torch.float64, device=dev), torch.ones(n, d, dtype=torch.float64, device=dev)
    part = torch.zeros(n_chunks * K * NS, dtype=torch.float64, device=dev)
    sums = torch.zeros(K * NS, dtype=torch.float64, device=dev)
    nv.check(lib.gpx_saltelli_partial(nv.ptr(z0), nv.ptr(z0), nv.ptr(z0d), nv.ptr(o1), nv.ptr(o1), nv.ptr(o1d), n, d, 0, n, K, 12345,
                                      nv.ptr(part), nv.stream_ptr()), "partial")
    nv.check(lib.gpx_saltelli_reduce(nv.ptr(part), n_chunks, d, K, nv.ptr(sums), nv.stream_ptr()), "reduce")
    S = sums.view(K, NS).cpu().numpy() / n
    assert abs(S[:, 0].mean()) < 4 / np.sqrt(n * K) and abs(S[:, 1].mean()) < 4 / np.sqrt(n * K)
    assert abs(S[:, 2].mean() - 1) < 6 * np.sqrt(2 / (n * K)) and abs(S[:, 3].mean() - 1) < 6 * np.sqrt(2 / (n * K))
    per_p = S[:, 4:].reshape(K, d, 3)
    assert np.abs(per_p[:, :, 0].mean(0)).max() < 6 * np.sqrt(2 / (n * K))          # f_B independent of f_AB - f_A
    assert np.abs(per_p[:, :, 2].mean(0) - 2).max() < 6 * np.sqrt(8 / (n * K))      # var(z_AB - z_A) = 2: the two are independent
    assert np.unique(np.round(S[:, 2], 12)).size == K                               # every draw is a different stream
    assert lib.gpx_saltelli_partial(nv.ptr(dA), nv.ptr(dB), nv.ptr(dAB), None, None, None, n, d, 100, n, 0, 1, nv.ptr(part), nv.stream_ptr()) < 0


def test_inference_and_diagnostics_reproduce_notebook(dev, goldens):
    from test_gpu_parity import _emulator_from_golden
    from gperks_b200.perks.diagnostics import Diagnostics
    from gperks_b200.perks.inference import Inference
    g = goldens["example_2"]
    em, ds = _emulator_from_golden(g, "example_2")
    inf = Inference(em)
    inf.summary(printtoconsole=False)
    assert float(inf.scores_dct["MeanSquaredError"]) == pytest.approx(g["printed"]["MSE"], abs=5e-4)
    assert float(inf.scores_dct["R2Score"]) == pytest.approx(g["printed"]["R2"], abs=5e-4)
    diag = Diagnostics(em)
    assert diag.chi_squared()[0] == pytest.approx(g["printed"]["chi2"], rel=2e-4)
    md, md_mean, _ = diag.mahalanobis_distance()
    assert float(md) == pytest.approx(g["printed"]["mahalanobis"], rel=2e-4) and md_mean == 25
    torch.manual_seed(0)
    ci, ci_mean, ci_std = diag.credible_interval()
    assert ci == pytest.approx(g["printed"]["CI"], abs=1e-12) and 0.5 < ci_mean <= 1.0
    for kind in ("correlated", "uncorrelated", "pivoted"):
        assert diag.errors(kind).shape == (25,)


def test_kfold_cross_validation(dev, tmp_path):
    from gperks_b200.perks.cross_validation import KFoldCrossValidation
    from gperks_b200.train.early_stop import NoEarlyStoppingCriterion
    from gperks_b200.train.snapshot import NeverSaveSnapshottingCriterion
    from gperks_b200.utils.test_functions import currin_exp
    em, ds = _fit(currin_exp, 2, 45, auto=False)
    exp = em.experiment
    opt = torch.optim.Adam(exp.model.parameters(), lr=0.1)
    snap = NeverSaveSnapshottingCriterion((tmp_path / "split_{split}" / "restart_{restart}").as_posix(), "epoch_{epoch}.pth")
    cv = KFoldCrossValidation(exp, ["cuda:0"], n_splits=3, max_workers=1)
    models, stats = cv.train(opt, NoEarlyStoppingCriterion(30), snap, leftout_is_val=True)
    assert set(models) == {0, 1, 2} and all(len(stats[s].train_loss) == 30 for s in stats)
    assert len(cv.best_test_scores_structured_dct["R2Score"]) == 3
    assert cv.best_split == int(np.argmax(cv.best_test_scores_structured_dct["R2Score"]))
    mean, std = cv.emulator.predict(ds.X_train[:5])
    assert mean.shape == (5,) and (std > 0).all()
    for s in range(3):
        assert (tmp_path / f"split_{s}" / "best_model.pth").exists()


def _restart_experiment(n_restarts, seed=8):
    from gperks_b200.gp.data.dataset import Dataset
    from gperks_b200.gp.experiment import GPExperiment
    from gperks_b200.gp.mean import LinearMean
    from gperks_b200.kernels import MaternKernel, ScaleKernel
    from gperks_b200.likelihoods import GaussianLikelihood
    from gperks_b200.utils.metrics import MeanSquaredError, R2Score
    from gperks_b200.utils.test_functions import currin_exp
    ds = Dataset.build_from_function(currin_exp, 2, 40, None, 15, "lhs", seed)
    torch.manual_seed(seed)     # LinearMean draws its initial weights when it is constructed, before GPExperiment seeds
    exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, 2), ScaleKernel(MaternKernel(ard_num_dims=2)),
                       n_restarts=n_restarts, seed=seed, metrics=[MeanSquaredError(), R2Score()])
    return exp, ds


def test_restarts_as_device_jobs_replay_the_sequential_draws(dev, tmp_path):
    """`train(devices=[...])` runs every restart as a self-contained job.  Restart r must start from the raw
    hyper-parameters the sequential loop draws for it; restart 1 (fresh optimizer in both modes) must reproduce the
    sequential trajectory exactly, and the selection / snapshot links must work from the jobs' files."""
    from gperks_b200.train.early_stop import NoEarlyStoppingCriterion
    from gperks_b200.train.emulator import GPEmulator
    from gperks_b200.train.snapshot import NeverSaveSnapshottingCriterion
    runs = {}
    for mode in ("sequential", "jobs"):
        exp, ds = _restart_experiment(3)
        em = GPEmulator(exp, "cuda:0")
        opt = torch.optim.Adam(exp.model.parameters(), lr=0.1)
        root = tmp_path / mode
        snap = NeverSaveSnapshottingCriterion((root / "restart_{restart}").as_posix(), "epoch_{epoch}.pth")
        kw = {"devices": ["cuda:0"], "max_workers": 1} if mode == "jobs" else {}
        best_model, best_stats = em.train(opt, NoEarlyStoppingCriterion(25), snap, **kw)
        stats = [json.loads((root / f"restart_{r}" / "train_stats.json").read_text()) for r in (1, 2, 3)]
        runs[mode] = (stats, best_stats, em.predict(ds.X_test), torch.rand(1).item())
        assert (root / "best_model.pth").exists() and (root / "best_train_stats.json").exists()
    seq, job = runs["sequential"], runs["jobs"]
    np.testing.assert_allclose(seq[0][0]["train_loss"], job[0][0]["train_loss"], rtol=1e-10)      # restart 1: identical
    for r in (1, 2):            # later restarts: same starting point, optimizer moments not carried over
        assert seq[0][r]["train_loss"][0] == pytest.approx(job[0][r]["train_loss"][0], rel=1e-10)
    assert seq[3] == job[3], "the caller's random stream must end where the sequential loop leaves it"
    assert len(job[1].train_loss) == 25 and np.isfinite(job[2][0]).all()


def test_kfold_with_fold_restart_jobs(dev, tmp_path):
    from gperks_b200.perks.cross_validation import KFoldCrossValidation
    from gperks_b200.train.early_stop import NoEarlyStoppingCriterion
    from gperks_b200.train.snapshot import NeverSaveSnapshottingCriterion
    exp, ds = _restart_experiment(2)
    opt = torch.optim.Adam(exp.model.parameters(), lr=0.1)
    snap = NeverSaveSnapshottingCriterion((tmp_path / "split_{split}" / "restart_{restart}").as_posix(), "epoch_{epoch}.pth")
    cv = KFoldCrossValidation(exp, ["cuda:0"], n_splits=3, max_workers=1)
    models, stats = cv.train(opt, NoEarlyStoppingCriterion(20), snap, parallel_restarts=True)
    assert set(models) == {0, 1, 2} and all(len(stats[s].train_loss) == 20 for s in stats)
    for s in range(3):
        for r in (1, 2):
            assert (tmp_path / f"split_{s}" / f"restart_{r}" / "train_stats.json").exists()
        assert (tmp_path / f"split_{s}" / "best_model.pth").exists()
    assert min(cv.best_test_scores_structured_dct["R2Score"]) > 0.8
    mean, std = cv.emulator.predict(ds.X_train[:4])
    assert mean.shape == (4,) and (std > 0).all()


@pytest.mark.parametrize("with_val", [False, True])
def test_batched_restarts_match_restart_jobs(dev, tmp_path, with_val):
    """All restarts trained at once (CUDA graph per restart, one optimizer over every replica) must follow, restart by
    restart, the trajectory of the same restart run on its own (`devices` mode: same draw, fresh optimizer state)."""
    from gperks_b200.train.early_stop import GLEarlyStoppingCriterion, NoEarlyStoppingCriterion
    from gperks_b200.train.emulator import GPEmulator
    from gperks_b200.train.snapshot import EveryEpochSnapshottingCriterion
    import gperks_b200.train.batched as batched_mod
    runs = {}
    for mode in ("jobs", "batched", "batched_generic"):
        batched_mod._FORCE_GENERIC = mode == "batched_generic"
        exp, ds = _restart_experiment(4)
        if with_val:
            from gperks_b200.gp.data.dataset import Dataset
            from gperks_b200.gp.experiment import GPExperiment
            d2 = Dataset(ds.X_train, ds.y_train, X_val=ds.X_test, y_val=ds.y_test, X_test=ds.X_test, y_test=ds.y_test)
            torch.manual_seed(8)
            exp = GPExperiment(d2, exp.likelihood, exp.mean_module, exp.covar_module, n_restarts=4, seed=8, metrics=exp.metrics)
        em = GPEmulator(exp, "cuda:0")
        opt = torch.optim.Adam(exp.model.parameters(), lr=0.1)
        root = tmp_path / mode
        snap = EveryEpochSnapshottingCriterion((root / "restart_{restart}").as_posix(), "epoch_{epoch}.pth")
        esc = GLEarlyStoppingCriterion(40, alpha=0.5, patience=3) if with_val else NoEarlyStoppingCriterion(30)
        kw = {"devices": ["cuda:0"], "max_workers": 1} if mode == "jobs" else {"batched_restarts": True}
        best_model, best_stats = em.train(opt, esc, snap, **kw)
        stats = [json.loads((root / f"restart_{r}" / "train_stats.json").read_text()) for r in (1, 2, 3, 4)]
        runs[mode] = (stats, best_stats, em.predict(ds.X_test), {k: v.clone() for k, v in best_model.items()})
        assert (root / "best_model.pth").exists()
    batched_mod._FORCE_GENERIC = False
    for other in ("batched", "batched_generic"):
        for a, b in zip(runs["jobs"][0], runs[other][0]):
            assert len(a["train_loss"]) == len(b["train_loss"]) and a["best_epoch"] == b["best_epoch"]
            np.testing.assert_allclose(a["train_loss"], b["train_loss"], rtol=1e-8, atol=1e-10)
            for name in a["train_metrics_score"]:
                np.testing.assert_allclose(a["train_metrics_score"][name], b["train_metrics_score"][name], rtol=1e-7, atol=1e-9)
            if with_val:
                np.testing.assert_allclose(a["val_loss"], b["val_loss"], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(runs["jobs"][2][0], runs[other][2][0], rtol=1e-7)
        for k, v in runs["jobs"][3].items():
            np.testing.assert_allclose(v.numpy(), runs[other][3][k].numpy(), rtol=1e-7, atol=1e-10)
