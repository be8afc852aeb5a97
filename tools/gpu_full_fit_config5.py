"""One full fit of a config-5 fold as SURVEY section 8(d) asks: N=32768 -> 26214 training points, D=16, RBF + linear mean, 100 Adam
epochs (every epoch: MLL + gradient, optimizer step, two train-set metrics, validation loss on the 6554 left-out points), through
train_restart_job on one GPU; wall time and the loss trajectory's ends.  usage: gpu_full_fit_config5.py [epochs] [N]"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import synthetic_problem  # noqa: E402
from sklearn.model_selection import KFold  # noqa: E402
from gperks_b200.gp.data.dataset import Dataset  # noqa: E402
from gperks_b200.gp.experiment import GPExperiment  # noqa: E402
from gperks_b200.gp.mean import LinearMean  # noqa: E402
from gperks_b200.kernels import RBFKernel, ScaleKernel  # noqa: E402
from gperks_b200.likelihoods import GaussianLikelihood  # noqa: E402
from gperks_b200.train.early_stop import NoEarlyStoppingCriterion  # noqa: E402
from gperks_b200.train.emulator import train_restart_job  # noqa: E402
from gperks_b200.train.snapshot import NeverSaveSnapshottingCriterion  # noqa: E402
from gperks_b200.utils.metrics import MeanSquaredError, R2Score  # noqa: E402
import logging  # noqa: E402

logging.disable(logging.INFO)
epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
D = 16
X, y, _, _ = synthetic_problem(n_train=N, dim=D, m=0)
tr, lo = list(KFold(n_splits=5).split(X))[0]


def fit(n_epochs):
    ds = Dataset(X[tr], y[tr], X_val=X[lo], y_val=y[lo], X_test=X[lo], y_test=y[lo])
    torch.manual_seed(8)
    exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, D), ScaleKernel(RBFKernel(ard_num_dims=D)), n_restarts=8, seed=8,
                       metrics=[MeanSquaredError(), R2Score()])
    snpc = NeverSaveSnapshottingCriterion("/tmp/gpx_full_fit/restart_{restart}", "epoch_{epoch}.pth")
    return train_restart_job(exp, None, "cuda:0", 1, torch.get_rng_state(), (torch.optim.Adam, {"lr": 0.1}),
                             NoEarlyStoppingCriterion(n_epochs), snpc)


fit(2)          # warm-up (max_epochs = 1 fails in the reference too: the criterion repeats an evaluation it never made): library load, first launches, the optimizer's imports
torch.cuda.synchronize()
t0 = time.perf_counter()
stats, best_epoch = fit(epochs)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
out = {"workload": f"one config-5 fit: {len(tr)} training points, D={D}, RBF + linear mean, {epochs} Adam epochs (lr 0.1), 2 train-set metrics + "
                   f"validation loss on {len(lo)} points per epoch", "seconds": wall, "seconds_per_epoch": wall / epochs,
       "train_loss_first_last": [float(stats.train_loss[0]), float(stats.train_loss[-1])],
       "val_loss_first_last": [float(stats.val_loss[0]), float(stats.val_loss[-1])],
       "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(out))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r02_full_fit_config5.json").write_text(json.dumps(out, indent=1))
