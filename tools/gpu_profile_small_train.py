"""cProfile of the Adam training loop at N=200 (config 1 shape): where does the per-epoch host time go?"""
import cProfile
import io
import logging
import pstats
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
from gperks_b200.kernels import RBFKernel, ScaleKernel  # noqa: E402
from gperks_b200.likelihoods import GaussianLikelihood  # noqa: E402
from gperks_b200.train.early_stop import NoEarlyStoppingCriterion  # noqa: E402
from gperks_b200.train.emulator import GPEmulator  # noqa: E402
from gperks_b200.train.snapshot import NeverSaveSnapshottingCriterion  # noqa: E402
from gperks_b200.utils.metrics import MeanSquaredError, R2Score  # noqa: E402
from gperks_b200.utils.test_functions_gsa import SobolGstar  # noqa: E402

lg = logging.getLogger("gperks_b200")
lg.setLevel(logging.WARNING)
for h in lg.handlers:
    h.setLevel(logging.WARNING)
d, n = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 200
f = partial(SobolGstar, a=[0, 1, 4.5, 9, 99, 99, 99, 99], delta=np.linspace(0.1, 0.9, d), alpha=[1.0] * d)
ds = Dataset.build_from_function(f, d, n, None, 64, "lhs", 8)
exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, d), ScaleKernel(RBFKernel(ard_num_dims=d)), n_restarts=1, seed=8,
                   metrics=[MeanSquaredError(), R2Score()])
em = GPEmulator(exp, "cuda:0")
opt = torch.optim.Adam(em.model.parameters(), lr=0.1)
snap = NeverSaveSnapshottingCriterion("/tmp/gpx_prof/restart_{restart}", "epoch_{epoch}.pth")
em.train(opt, NoEarlyStoppingCriterion(10), snap)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
em.train(opt, NoEarlyStoppingCriterion(100), snap)
pr.disable()
print("ms/epoch", 1e3 * (time.perf_counter() - t0) / 100)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:75]))
