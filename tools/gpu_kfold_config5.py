"""BASELINE configs[4] through the product API: KFoldCrossValidation(devices=[cuda:0..G-1]).train(..., parallel_restarts=True)
on N=32768, D=16 (RBF + linear mean), 5 folds x 8 restarts = 40 fits of 26214/26215 points dealt over the GPUs of the box in
spawned workers (GPErks/perks/cross_validation.py:99-143).  Epochs per restart are scaled to the GPU-minute budget (argument 1,
default 3; the reference's default is 100).  Wall time, per-split R2 and the per-GPU utilisation sampled with nvidia-smi go to
gpurun_out/<tag>_kfold_config5.json.   usage: gpu_kfold_config5.py [epochs] [N] [tag]"""
import json
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class UtilSampler:
    def __init__(self):
        self.samples, self._stop = [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=utilization.gpu,memory.used", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
                self.samples.append([[float(v) for v in line.split(",")] for line in out])
            except Exception:
                pass
            self._stop.wait(0.5)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {}
        a = np.array([s for s in self.samples if len(s) == len(self.samples[0])])
        return {"samples": int(a.shape[0]), "busy_fraction_per_gpu": (a[:, :, 0] > 5).mean(0).round(3).tolist(),
                "mean_utilization_per_gpu": a[:, :, 0].mean(0).round(1).tolist(), "max_memory_mib_per_gpu": a[:, :, 1].max(0).tolist()}


def main():
    from bench import synthetic_problem
    from gperks_b200.gp.data.dataset import Dataset
    from gperks_b200.gp.experiment import GPExperiment
    from gperks_b200.gp.mean import LinearMean
    from gperks_b200.kernels import RBFKernel, ScaleKernel
    from gperks_b200.likelihoods import GaussianLikelihood
    from gperks_b200.perks.cross_validation import KFoldCrossValidation
    from gperks_b200.train.early_stop import NoEarlyStoppingCriterion
    from gperks_b200.train.snapshot import NeverSaveSnapshottingCriterion
    from gperks_b200.utils.metrics import MeanSquaredError, R2Score
    epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    tag = sys.argv[3] if len(sys.argv) > 3 else "r02"
    D, n_splits, n_restarts = 16, 5, 8
    n_gpu = torch.cuda.device_count()
    devices = [f"cuda:{i}" for i in range(n_gpu)]
    X, y, _, _ = synthetic_problem(n_train=N, dim=D, m=0)
    exp = GPExperiment(Dataset(X, y), GaussianLikelihood(), LinearMean(1, D), ScaleKernel(RBFKernel(ard_num_dims=D)), n_restarts=n_restarts,
                       seed=8, metrics=[MeanSquaredError(), R2Score()])
    opt = torch.optim.Adam(exp.model.parameters(), lr=0.1)
    import tempfile
    root = tempfile.mkdtemp(prefix="gpx_c5_")
    snap = NeverSaveSnapshottingCriterion(root + "/split_{split}/restart_{restart}", "epoch_{epoch}.pth")
    cv = KFoldCrossValidation(exp, devices, n_splits=n_splits, max_workers=n_gpu)
    with UtilSampler() as util:
        t0 = time.perf_counter()
        models, stats = cv.train(opt, NoEarlyStoppingCriterion(epochs), snap, leftout_is_val=True, parallel_restarts=True)
        wall = time.perf_counter() - t0
    out = {"workload": "configs[4]: 5-fold x 8 restarts, N=%d D=%d RBF, KFoldCrossValidation(devices=%d GPUs).train(parallel_restarts=True)" % (N, D, n_gpu),
           "fits": n_splits * n_restarts, "fold_train_size": int(N - N // n_splits), "epochs_per_restart": epochs, "wall_seconds": wall,
           "fits_per_hour": n_splits * n_restarts / wall * 3600, "epochs_per_second_all_gpus": n_splits * n_restarts * epochs / wall,
           "R2_per_split": np.round(cv.best_test_scores_structured_dct["R2Score"], 4).tolist(), "best_split": int(cv.best_split),
           "gpus": util.summary(), "note": "wall clock includes spawning one worker per GPU (CUDA context + library load in each) and the parent's "
                                           "best-restart selection + Inference on every fold"}
    print(json.dumps(out), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"{tag}_kfold_config5.json").write_text(json.dumps(out, indent=1))
    print("KFOLD5-OK")


if __name__ == "__main__":
    main()
