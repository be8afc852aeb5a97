"""Where one config-5 fit spends its time: one (fold, restart) job of bench.py --workload kfold (N=32768 -> 26214 training points,
D=16, RBF, 3 Adam epochs) under cProfile, CUDA launches made blocking so that host time = device time.  usage: gpu_fit_profile.py [N] [epochs]"""
import cProfile
import os
import pstats
import sys
import time
from pathlib import Path

os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
import torch  # noqa: E402

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

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
D = 16
X, y, _, _ = synthetic_problem(n_train=N, dim=D, m=0)
tr, lo = list(KFold(n_splits=5).split(X))[0]


def fit():
    ds = Dataset(X[tr], y[tr], X_val=X[lo], y_val=y[lo], X_test=X[lo], y_test=y[lo])
    torch.manual_seed(8)
    exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, D), ScaleKernel(RBFKernel(ard_num_dims=D)), n_restarts=8, seed=8,
                       metrics=[MeanSquaredError(), R2Score()])
    snpc = NeverSaveSnapshottingCriterion("/tmp/gpx_fit_profile/restart_{restart}", "epoch_{epoch}.pth")
    return train_restart_job(exp, None, "cuda:0", 1, torch.get_rng_state(), (torch.optim.Adam, {"lr": 0.1}),
                             NoEarlyStoppingCriterion(epochs), snpc)


fit() if N <= 4096 else None
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
fit()
torch.cuda.synchronize()
pr.disable()
print("fit seconds", time.perf_counter() - t0)
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
