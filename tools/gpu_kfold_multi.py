"""K-fold cross validation with folds dealt over several GPUs in spawned workers (run under gpurun --gpus 2)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from gperks_b200.gp.data.dataset import Dataset
    from gperks_b200.gp.experiment import GPExperiment
    from gperks_b200.gp.mean import LinearMean
    from gperks_b200.kernels import MaternKernel, ScaleKernel
    from gperks_b200.likelihoods import GaussianLikelihood
    from gperks_b200.perks.cross_validation import KFoldCrossValidation
    from gperks_b200.train.early_stop import NoEarlyStoppingCriterion
    from gperks_b200.train.snapshot import NeverSaveSnapshottingCriterion
    from gperks_b200.utils.metrics import MeanSquaredError, R2Score
    from gperks_b200.utils.test_functions import currin_exp
    n_gpu = torch.cuda.device_count()
    devices = [f"cuda:{i}" for i in range(n_gpu)]
    ds = Dataset.build_from_function(currin_exp, 2, 400, None, None, "lhs", 8)
    exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, 2), ScaleKernel(MaternKernel(ard_num_dims=2)), n_restarts=1,
                       seed=8, metrics=[MeanSquaredError(), R2Score()])
    opt = torch.optim.Adam(exp.model.parameters(), lr=0.1)
    snap = NeverSaveSnapshottingCriterion("/tmp/gpx_cv/split_{split}/restart_{restart}", "epoch_{epoch}.pth")
    cv = KFoldCrossValidation(exp, devices, n_splits=4, max_workers=n_gpu)
    t0 = time.perf_counter()
    models, stats = cv.train(opt, NoEarlyStoppingCriterion(40), snap)
    dt = time.perf_counter() - t0
    print("devices", devices, "splits", sorted(models), "seconds", round(dt, 2))
    # the same with (fold, restart) jobs: 4 folds x 2 restarts dealt over the GPUs
    exp2 = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, 2), ScaleKernel(MaternKernel(ard_num_dims=2)), n_restarts=2,
                        seed=8, metrics=[MeanSquaredError(), R2Score()])
    snap2 = NeverSaveSnapshottingCriterion("/tmp/gpx_cv2/split_{split}/restart_{restart}", "epoch_{epoch}.pth")
    cv2 = KFoldCrossValidation(exp2, devices, n_splits=4, max_workers=n_gpu)
    t0 = time.perf_counter()
    models2, _ = cv2.train(torch.optim.Adam(exp2.model.parameters(), lr=0.1), NoEarlyStoppingCriterion(40), snap2,
                           parallel_restarts=True)
    print("fold x restart jobs:", len(models2), "splits, R2", np.round(cv2.best_test_scores_structured_dct["R2Score"], 4),
          "seconds", round(time.perf_counter() - t0, 2))
    assert min(cv2.best_test_scores_structured_dct["R2Score"]) > 0.9
    print("R2 per split", np.round(cv.best_test_scores_structured_dct["R2Score"], 4), "best", cv.best_split)
    mean, std = cv.emulator.predict(ds.X_train[:3])
    print("best-split emulator predicts", mean, std)
    assert len(models) == 4 and min(cv.best_test_scores_structured_dct["R2Score"]) > 0.9
    print("KFOLD-OK")


if __name__ == "__main__":
    main()
