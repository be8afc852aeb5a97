"""K-fold cross validation of an experiment (API of GPErks/perks/cross_validation.py:40-274).

Folds are independent fits: they are dealt round-robin onto ``devices`` (one B200 each) and run in spawned
worker processes; only state dicts (numpy) and scores travel back.  Under ``torchrun`` the same folds can be
dealt over ranks with ``gperks_b200.parallel.run_jobs_round_robin``."""
from copy import deepcopy
from itertools import cycle
from typing import List, Optional

import numpy as np
import pandas as pd
import torch
from sklearn.model_selection import KFold

from ..constants import (
    DEFAULT_CROSS_VALIDATION_MAX_WORKERS,
    DEFAULT_CROSS_VALIDATION_N_SPLITS,
    DEFAULT_TRAIN_MAX_EPOCH,
    DEFAULT_TRAIN_SNAPSHOT_DIR,
    DEFAULT_TRAIN_SNAPSHOT_EPOCH_TEMPLATE,
    DEFAULT_TRAIN_SNAPSHOT_RESTART_TEMPLATE,
    DEFAULT_TRAIN_SNAPSHOT_SPLIT_TEMPLATE,
)
from ..gp.data.dataset import Dataset
from ..gp.experiment import GPExperiment
from ..log.logger import get_logger
from ..serialization.path import posix_path
from ..train.early_stop import EarlyStoppingCriterion, NoEarlyStoppingCriterion
from ..train.emulator import GPEmulator
from ..train.snapshot import NeverSaveSnapshottingCriterion, SnapshottingCriterion
from ..train.trainable import Trainable
from ..utils.concurrency import execute_task_in_parallel
from ..utils.metrics import get_metric_name
from .inference import Inference

log = get_logger()


class KFoldCrossValidation(Trainable):
    def __init__(self, experiment: GPExperiment, devices: List[str],
                 n_splits: int = DEFAULT_CROSS_VALIDATION_N_SPLITS,
                 max_workers: int = DEFAULT_CROSS_VALIDATION_MAX_WORKERS, *, shuffle: bool = False,
                 random_state: Optional[int] = None, leftout_is_val: Optional[bool] = None):
        self.experiment = experiment
        self.devices = devices
        self.n_splits = n_splits
        self.max_workers = max_workers
        self.split_generator = KFold(n_splits=n_splits, shuffle=shuffle, random_state=random_state)
        self.leftout_is_val = leftout_is_val
        self.best_split = None
        self.best_split_idx = [None, None]
        self.best_test_scores_structured_dct = dict()
        self.emulator: Optional[GPEmulator] = None

    def train(self, optimizer, early_stopping_criterion: Optional[EarlyStoppingCriterion] = None,
              snapshotting_criterion: Optional[SnapshottingCriterion] = None, leftout_is_val: bool = False):
        if early_stopping_criterion is None:
            early_stopping_criterion = NoEarlyStoppingCriterion(DEFAULT_TRAIN_MAX_EPOCH)
        if snapshotting_criterion is None:
            snapshotting_criterion = NeverSaveSnapshottingCriterion(
                posix_path(DEFAULT_TRAIN_SNAPSHOT_DIR, DEFAULT_TRAIN_SNAPSHOT_SPLIT_TEMPLATE,
                           DEFAULT_TRAIN_SNAPSHOT_RESTART_TEMPLATE), DEFAULT_TRAIN_SNAPSHOT_EPOCH_TEMPLATE)
        self.leftout_is_val = leftout_is_val
        X, y = self.experiment.dataset.X_train, self.experiment.dataset.y_train
        opt_spec = (optimizer.__class__, dict(optimizer.defaults))
        tasks = {
            i: (opt_spec, early_stopping_criterion, snapshotting_criterion, i, device, X[tr], y[tr], X[lo], y[lo], tr, lo)
            for i, (device, (tr, lo)) in enumerate(zip(cycle(self.devices), self.split_generator.split(X)))
        }
        models, stats, scores, idx = {}, {}, {}, {}
        for split, (best_model, best_stats, test_scores, split_idx) in execute_task_in_parallel(
                self._train_split, tasks, self.max_workers).items():
            models[split] = {k: torch.as_tensor(v) for k, v in best_model.items()}
            stats[split], scores[split], idx[split] = best_stats, test_scores, split_idx

        names = [get_metric_name(m) for m in self.experiment.metrics]
        self.best_test_scores_structured_dct = {n: [float(np.asarray(scores[s][n]).item()) for s in range(self.n_splits)]
                                                for n in names}
        self.best_split = int(np.argmax(self.best_test_scores_structured_dct["R2Score"])) if "R2Score" in names else 0
        self.best_split_idx = idx[self.best_split]
        self.emulator = self._get_emulator(models[self.best_split])
        return models, stats

    def summary(self):
        data = np.array(list(self.best_test_scores_structured_dct.values()))
        data = np.hstack((data, data.mean(axis=1, keepdims=True), data.std(axis=1, keepdims=True)))
        print(pd.DataFrame(data=np.around(data, decimals=4), index=list(self.best_test_scores_structured_dct.keys()),
                           columns=[f"Split {i}" for i in range(self.n_splits)] + ["Mean", "Std"]))

    def _fold_experiment(self, dataset: Dataset) -> GPExperiment:
        e = self.experiment
        return GPExperiment(dataset, deepcopy(e.likelihood), deepcopy(e.mean_module), deepcopy(e.covar_module),
                            n_restarts=e.n_restarts, seed=e.seed, metrics=e.metrics, learn_noise=e.learn_noise)

    def _get_emulator(self, best_state=None) -> GPEmulator:
        """Emulator of the best split, rebuilt from that split's data and its best state dict."""
        ds = self.experiment.dataset
        tr, lo = self.best_split_idx
        dataset = Dataset(ds.X_train[tr], ds.y_train[tr], X_test=ds.X_train[lo], y_test=ds.y_train[lo],
                          x_labels=ds.x_labels, y_label=ds.y_label, name=ds.name, descr=ds.descr)
        emulator = GPEmulator(self._fold_experiment(dataset), "cpu")
        if best_state is None:
            best_state = torch.load(posix_path(DEFAULT_TRAIN_SNAPSHOT_DIR,
                                               DEFAULT_TRAIN_SNAPSHOT_SPLIT_TEMPLATE.format(split=self.best_split),
                                               "best_model.pth"), map_location=torch.device("cpu"))
        emulator.model.load_state_dict(best_state)
        return emulator

    def _train_split(self, opt_spec, early_stopping_criterion, snapshotting_criterion, i, device, X_train, y_train,
                     X_leftout, y_leftout, idx_train, idx_leftout):
        log.info(f"Running K-fold split {i}...")
        ds = self.experiment.dataset
        if self.leftout_is_val:
            X_val, y_val = X_leftout, y_leftout
        else:
            X_val, y_val = ds.X_val, ds.y_val
        dataset = Dataset(X_train, y_train, X_val=X_val, y_val=y_val, X_test=X_leftout, y_test=y_leftout,
                          x_labels=ds.x_labels, y_label=ds.y_label)
        experiment = self._fold_experiment(dataset)
        opt_cls, opt_defaults = opt_spec
        optimizer = opt_cls(experiment.model.parameters(), **opt_defaults)
        esc, snpc = deepcopy(early_stopping_criterion), deepcopy(snapshotting_criterion)
        snpc.snapshot_dir = snpc.snapshot_dir.replace("{restart}", "{{restart}}").format(split=i)
        emulator = GPEmulator(experiment, device)
        best_model, best_stats = emulator.train(optimizer, early_stopping_criterion=esc, snapshotting_criterion=snpc)
        log.info(f"Run K-fold split {i}.")
        inference = Inference(emulator)
        inference.summary(printtoconsole=False)
        return ({k: v.cpu().detach().numpy() for k, v in best_model.items()}, best_stats, inference.scores_dct,
                [idx_train, idx_leftout])
