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
              snapshotting_criterion: Optional[SnapshottingCriterion] = None, leftout_is_val: bool = False, *,
              parallel_restarts: bool = False):
        """``parallel_restarts`` (extension): deal the n_splits x n_restarts independent fits, not just
        the folds, over ``devices`` -- e.g. 5 folds x 8 restarts = 40 jobs on 8 GPUs in 5 waves instead of 5 GPUs each
        running 8 restarts in a row.  Every (fold, restart) job starts from the raw hyper-parameter draw it would get
        in the sequential run and uses a fresh optimizer (see ``GPEmulator.train``)."""
        if early_stopping_criterion is None:
            early_stopping_criterion = NoEarlyStoppingCriterion(DEFAULT_TRAIN_MAX_EPOCH)
        if snapshotting_criterion is None:
            snapshotting_criterion = NeverSaveSnapshottingCriterion(
                posix_path(DEFAULT_TRAIN_SNAPSHOT_DIR, DEFAULT_TRAIN_SNAPSHOT_SPLIT_TEMPLATE,
                           DEFAULT_TRAIN_SNAPSHOT_RESTART_TEMPLATE), DEFAULT_TRAIN_SNAPSHOT_EPOCH_TEMPLATE)
        self.leftout_is_val = leftout_is_val
        X, y = self.experiment.dataset.X_train, self.experiment.dataset.y_train
        opt_spec = (optimizer.__class__, dict(optimizer.defaults))
        if parallel_restarts:
            return self._train_fold_restart_jobs(opt_spec, early_stopping_criterion, snapshotting_criterion, X, y)
        tasks = {
            i: (opt_spec, early_stopping_criterion, snapshotting_criterion, i, device, X[tr], y[tr], X[lo], y[lo], tr, lo)
            for i, (device, (tr, lo)) in enumerate(zip(cycle(self.devices), self.split_generator.split(X)))
        }
        models, stats, scores, idx = {}, {}, {}, {}
        for split, (best_model, best_stats, test_scores, split_idx) in execute_task_in_parallel(
                self._train_split, tasks, self.max_workers).items():
            models[split] = {k: torch.as_tensor(v) for k, v in best_model.items()}
            stats[split], scores[split], idx[split] = best_stats, test_scores, split_idx
        return self._finish(models, stats, scores, idx)

    def _finish(self, models, stats, scores, idx):
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

    def _train_fold_restart_jobs(self, opt_spec, early_stopping_criterion, snapshotting_criterion, X, y):
        splits = list(self.split_generator.split(X))
        n_restarts = self.experiment.n_restarts
        devices = cycle(self.devices)
        tasks = {(i, r): (opt_spec, early_stopping_criterion, snapshotting_criterion, i, r, next(devices),
                          X[tr], y[tr], X[lo], y[lo])
                 for i, (tr, lo) in enumerate(splits) for r in range(1, n_restarts + 1)}
        # one worker per device, bound to it: (fold, restart) jobs are pulled by whichever GPU becomes free (argument 5 = device)
        results = execute_task_in_parallel(self._train_split_restart, tasks, self.max_workers, devices=self.devices, device_arg=5)
        models, stats, scores, idx = {}, {}, {}, {}

        def finish_fold(i, tr, lo):
            # best restart of the fold, its snapshot and its test scores: same selection as GPEmulator.train
            emulator = GPEmulator(self._fold_experiment(self._fold_dataset(X[tr], y[tr], X[lo], y[lo])),
                                  self.devices[i % len(self.devices)])
            snpc = self._split_snapshotting(snapshotting_criterion, i)
            best_model, best_stats = emulator._select_best_restart(
                [results[(i, r)][0] for r in range(1, n_restarts + 1)],
                [results[(i, r)][1] for r in range(1, n_restarts + 1)], snpc)
            inference = Inference(emulator)
            inference.summary(printtoconsole=False)
            return {k: torch.as_tensor(v) for k, v in best_model.items()}, best_stats, inference.scores_dct

        # every fold's selection + test-set scores on its own GPU, side by side (the device work -- an FP64 factorisation of the
        # fold and a predict -- releases the GIL); one fold after the other when there is a single device
        if len(self.devices) > 1 and len(splits) > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(len(self.devices), len(splits))) as pool:
                futs = {i: pool.submit(finish_fold, i, tr, lo) for i, (tr, lo) in enumerate(splits)}
                done = {i: f.result() for i, f in futs.items()}
        else:
            done = {i: finish_fold(i, tr, lo) for i, (tr, lo) in enumerate(splits)}
        for i, (tr, lo) in enumerate(splits):
            models[i], stats[i], scores[i] = done[i]
            idx[i] = [tr, lo]
        return self._finish(models, stats, scores, idx)

    def _train_split_restart(self, opt_spec, early_stopping_criterion, snapshotting_criterion, i, restart, device,
                             X_train, y_train, X_leftout, y_leftout):
        from ..train.emulator import train_restart_job
        log.info(f"Running K-fold split {i}, restart {restart}...")
        experiment = self._fold_experiment(self._fold_dataset(X_train, y_train, X_leftout, y_leftout))
        # the random stream right after the fold's experiment is built is where the sequential run starts drawing
        return train_restart_job(experiment, None, device, restart, torch.get_rng_state(), opt_spec,
                                 early_stopping_criterion, self._split_snapshotting(snapshotting_criterion, i))

    def _fold_dataset(self, X_train, y_train, X_leftout, y_leftout) -> Dataset:
        ds = self.experiment.dataset
        if self.leftout_is_val:
            X_val, y_val = X_leftout, y_leftout
        else:
            X_val, y_val = ds.X_val, ds.y_val
        return Dataset(X_train, y_train, X_val=X_val, y_val=y_val, X_test=X_leftout, y_test=y_leftout,
                       x_labels=ds.x_labels, y_label=ds.y_label)

    @staticmethod
    def _split_snapshotting(snapshotting_criterion, i):
        snpc = deepcopy(snapshotting_criterion)
        snpc.snapshot_dir = snpc.snapshot_dir.replace("{restart}", "{{restart}}").format(split=i)
        return snpc

    def _train_split(self, opt_spec, early_stopping_criterion, snapshotting_criterion, i, device, X_train, y_train,
                     X_leftout, y_leftout, idx_train, idx_leftout):
        log.info(f"Running K-fold split {i}...")
        experiment = self._fold_experiment(self._fold_dataset(X_train, y_train, X_leftout, y_leftout))
        opt_cls, opt_defaults = opt_spec
        optimizer = opt_cls(experiment.model.parameters(), **opt_defaults)
        esc, snpc = deepcopy(early_stopping_criterion), self._split_snapshotting(snapshotting_criterion, i)
        emulator = GPEmulator(experiment, device)
        best_model, best_stats = emulator.train(optimizer, early_stopping_criterion=esc, snapshotting_criterion=snpc)
        log.info(f"Run K-fold split {i}.")
        inference = Inference(emulator)
        inference.summary(printtoconsole=False)
        return ({k: v.cpu().detach().numpy() for k, v in best_model.items()}, best_stats, inference.scores_dct,
                [idx_train, idx_leftout])
