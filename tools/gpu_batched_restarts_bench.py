"""Config-1 shaped training (N=200, D=8, RBF + linear mean, Adam, 100 epochs): restarts one after the other vs all at
once (`train(batched_restarts=True)`), for 3 and 8 restarts.  Run under gpurun."""
import json
import logging
import sys
import tempfile
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

logging.disable(logging.CRITICAL)
d = 8
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
f = partial(SobolGstar, a=[0, 1, 4.5, 9, 99, 99, 99, 99], delta=np.linspace(0.1, 0.9, d), alpha=[1.0] * d)
ds = Dataset.build_from_function(f, d, n, None, 64, "lhs", 8)
out = {"N": n, "D": d, "epochs": 100}
for R in (1, 3, 8):
    for mode in ("sequential", "batched"):
        torch.manual_seed(8)
        exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, d), ScaleKernel(RBFKernel(ard_num_dims=d)), n_restarts=R, seed=8,
                           metrics=[MeanSquaredError(), R2Score()])
        em = GPEmulator(exp, "cuda:0")
        opt = torch.optim.Adam(exp.model.parameters(), lr=0.1)
        root = Path(tempfile.mkdtemp())
        snap = NeverSaveSnapshottingCriterion((root / "restart_{restart}").as_posix(), "epoch_{epoch}.pth")
        kw = {"batched_restarts": True} if mode == "batched" else {}
        em.train(opt, NoEarlyStoppingCriterion(5), snap, **kw)            # warm-up: first launches, graph capture
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, best = em.train(opt, NoEarlyStoppingCriterion(100), snap, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[f"R{R}_{mode}_s"] = round(dt, 4)
        out[f"R{R}_{mode}_final_loss"] = round(best.train_loss[-1], 6)
    out[f"R{R}_speedup"] = round(out[f"R{R}_sequential_s"] / out[f"R{R}_batched_s"], 2)
print(json.dumps(out))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "batched_restarts.json").write_text(json.dumps(out, indent=1))
