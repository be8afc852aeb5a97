"""Consumer-level measurements on one B200 (run under gpurun): config-1 style Adam training, history matching
(config 4 shape, bounded sample), Saltelli GSA on the device (config 3 shape)."""
import json
import sys
import time
from functools import partial
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from gperks_b200.gp.data.dataset import Dataset  # noqa: E402
from gperks_b200.gp.experiment import GPExperiment  # noqa: E402
from gperks_b200.gp.mean import LinearMean  # noqa: E402
from gperks_b200.kernels import MaternKernel, RBFKernel, ScaleKernel  # noqa: E402
from gperks_b200.likelihoods import GaussianLikelihood  # noqa: E402
from gperks_b200.perks.gsa import SobolGSA  # noqa: E402
from gperks_b200.perks.history_matching import Wave  # noqa: E402
from gperks_b200.train.early_stop import NoEarlyStoppingCriterion  # noqa: E402
from gperks_b200.train.emulator import GPEmulator  # noqa: E402
from gperks_b200.train.snapshot import NeverSaveSnapshottingCriterion  # noqa: E402
from gperks_b200.utils.metrics import MeanSquaredError, R2Score  # noqa: E402
from gperks_b200.utils.test_functions_gsa import SobolGstar  # noqa: E402
import logging
logging.getLogger("gperks_b200").setLevel(logging.WARNING)
for h in logging.getLogger("gperks_b200").handlers:
    h.setLevel(logging.WARNING)

out = {}


def gstar(d):
    a = ([0, 1, 4.5, 9, 99, 99, 99, 99] + [99] * 8)[:d]
    return partial(SobolGstar, a=a, delta=np.linspace(0.1, 0.9, d), alpha=[1.0] * d)


def emulator(d, n, kernel, seed=8, fixed=True):
    ds = Dataset.build_from_function(gstar(d), d, n, None, 64, "sobol", seed, l_bounds=[0.0] * d, u_bounds=[1.0] * d)
    base = RBFKernel(ard_num_dims=d) if kernel == "rbf" else MaternKernel(ard_num_dims=d)
    exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, d), ScaleKernel(base), n_restarts=1, seed=seed,
                       metrics=[MeanSquaredError(), R2Score()])
    em = GPEmulator(exp, "cuda:0")
    if fixed:
        m = em.model
        m.likelihood.noise = 1e-2
        m.covar_module.outputscale = 1.0
        m.covar_module.base_kernel.lengthscale = torch.tensor([0.5 * (1 + i / d) for i in range(d)], dtype=torch.float64)
    return em, ds


# ---- config 1: Adam on ExactMLL, N=200, D=8, RBF + linear mean, then predict 1e4 points
em, ds = emulator(8, 200, "rbf", fixed=False)
opt = torch.optim.Adam(em.model.parameters(), lr=0.1)
snap = NeverSaveSnapshottingCriterion("/tmp/gpx_c1/restart_{restart}", "epoch_{epoch}.pth")
em.train(opt, NoEarlyStoppingCriterion(20), snap)            # warm-up (allocations, first launches)
torch.cuda.synchronize()
t0 = time.perf_counter()
em.train(opt, NoEarlyStoppingCriterion(100), snap)
torch.cuda.synchronize()
t_train = time.perf_counter() - t0
X1 = np.random.default_rng(0).random((10_000, 8))
em.predict(X1)
t0 = time.perf_counter()
em.predict(X1)
t_pred = time.perf_counter() - t0
out["config1"] = {"N": 200, "D": 8, "epochs": 100, "ms_per_epoch": 1e3 * t_train / 100, "predict_1e4_ms": 1e3 * t_pred,
                  "note": "epoch = train step + 2 train-set metrics, one restart; reference notebooks log ~3.3 ms/epoch at N=10-20 on CPU"}
print(json.dumps(out["config1"]), flush=True)

# ---- config 4 shape: 8 emulators N=2048 D=6, implausibility over M points (bounded sample of the 5e7 grid)
ems = [emulator(6, 2048, "matern", seed=8 + i)[0] for i in range(8)]
import os
M = int(os.environ.get("GPX_C4_M", 4_000_000))      # BASELINE config 4 is 5e7; default is a bounded sample
X4 = np.random.default_rng(1).random((M, 6))
W = Wave(emulator=ems, Itrain=np.array([[0.0, 1.0]] * 6), cutoff=3.0, maxno=1, mean=[1.0] * 8, var=[0.05] * 8)
W.compute_impl(X4[:200_000])
torch.cuda.synchronize()
t0 = time.perf_counter()
I, PV = W.compute_impl(X4)
t_hm = time.perf_counter() - t0
t0 = time.perf_counter()
W.find_regions(X4)
t_fr = time.perf_counter() - t0
out["config4"] = {"emulators": 8, "N": 2048, "D": 6, "M": M, "seconds": t_hm, "points_per_s": M / t_hm,
                  "emulator_evals_per_s": 8 * M / t_hm, "nimp_fraction": float((I < 3.0).mean()),
                  "find_regions_seconds": t_fr, "fused_wave_path": W._wave_descriptors(torch.device("cuda:0"))[0] is not None,
                  "precisions": [e.auto_precisions()[0] for e in ems], "overlap": os.environ.get("GPX_I8_OVERLAP", "1"),
                  "reference_python_loop_us_per_point": 8.4}
print(json.dumps(out["config4"]), flush=True)

# ---- config 3 shape: Saltelli design N_train=4096, D=12: n=2^20 -> 1.47e7 emulator evaluations (posterior-mean indices), and
# n=2^19 -> 7.34e6 evaluations of mean+std with 1000 marginal draws reduced on the device
em3, ds3 = emulator(12, 4096, "matern")
gsa_small = SobolGSA(ds3, n=2 ** 12, seed=8)
gsa_small.estimate_Sobol_indices_with_emulator_device(em3, n_draws=0)
gsa_small.estimate_Sobol_indices_with_emulator_device(em3, n_draws=8)
torch.cuda.synchronize()
gsa = SobolGSA(ds3, n=2 ** 20, seed=8)
t0 = time.perf_counter()
gsa._base_design()
t_design = time.perf_counter() - t0
t0 = time.perf_counter()
gsa.estimate_Sobol_indices_with_emulator_device(em3, n_draws=0)
t_gsa = time.perf_counter() - t0
out["config3_mean"] = {"N": 4096, "D": 12, "n": 2 ** 20, "evaluations": int(2 ** 20 * 14), "seconds_total_wall": t_gsa + t_design,
                       "seconds_sobol_base_block_host": t_design, "seconds_device_pipeline": t_gsa,
                       "evals_per_s": 2 ** 20 * 14 / (t_gsa + t_design), "ST": np.round(gsa.ST[0], 4).tolist()}
print(json.dumps(out["config3_mean"]), flush=True)
gsa2 = SobolGSA(ds3, n=2 ** 19, seed=8)
t0 = time.perf_counter()
gsa2.estimate_Sobol_indices_with_emulator_device(em3, n_draws=1000, draw_seed=1)
t_draws = time.perf_counter() - t0
out["config3_draws"] = {"N": 4096, "D": 12, "n": 2 ** 19, "evaluations": int(2 ** 19 * 14), "n_draws": 1000, "seconds_total_wall": t_draws,
                        "evals_per_s": 2 ** 19 * 14 / t_draws, "ST_median": np.round(np.median(gsa2.ST, 0), 4).tolist(),
                        "ST_std_over_draws": np.round(gsa2.ST.std(0), 5).tolist(), "shape": list(gsa2.ST.shape)}
print(json.dumps(out["config3_draws"]), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"{os.environ.get('GPX_TAG', 'r02')}_consumer_bench.json").write_text(json.dumps(out, indent=1))
