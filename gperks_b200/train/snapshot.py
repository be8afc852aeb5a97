"""Snapshotting criteria: ``torch.save(model.state_dict())`` to
``<snapshot_dir.format(restart=r)>/<snapshot_file.format(epoch=e)>`` (GPErks/train/snapshot.py)."""
import os
from abc import ABCMeta, abstractmethod
from pathlib import Path

import torch

from ..log.logger import get_logger

log = get_logger()


class SnapshottingCriterion(metaclass=ABCMeta):
    def __init__(self, snapshot_dir: str, snapshot_file: str):
        self.model = None
        self.train_stats = None
        self.snapshot_dir = snapshot_dir
        self.snapshot_file = snapshot_file
        self._reached_epoch = -1

    def enable(self, model, train_stats):
        self.model, self.train_stats = model, train_stats
        self._reached_epoch = -1

    def get_snapshot_file_path(self, restart, epoch):
        folder = Path(self.snapshot_dir.format(restart=restart))
        folder.mkdir(parents=True, exist_ok=True)
        return (folder / self.snapshot_file.format(epoch=epoch)).as_posix()

    def save(self, restart, epoch):
        self._save_snapshot(self.get_snapshot_file_path(restart, epoch))
        self._reached_epoch = epoch

    def maybe_save(self, restart, epoch):
        if self._should_save():
            self.save(restart, epoch)

    def keep_snapshots_until(self, restart, epoch):
        """Drop the snapshots taken after the best epoch."""
        for later in range(epoch + 1, self._reached_epoch + 1):
            try:
                os.remove(self.get_snapshot_file_path(restart, later))
            except FileNotFoundError:
                pass

    def _save_snapshot(self, path):
        state = {k: v.detach().cpu() for k, v in self.model.state_dict().items()}
        torch.save(state, path)

    @abstractmethod
    def _should_save(self) -> bool: ...


class EveryNEpochsSnapshottingCriterion(SnapshottingCriterion):
    def __init__(self, snapshot_dir: str, snapshot_file: str, n: int):
        super().__init__(snapshot_dir, snapshot_file)
        self.n = n

    def _should_save(self) -> bool:
        return self.train_stats.current_epoch % self.n == 0


class EveryEpochSnapshottingCriterion(EveryNEpochsSnapshottingCriterion):
    def __init__(self, snapshot_dir: str, snapshot_file: str):
        super().__init__(snapshot_dir, snapshot_file, 1)


class NeverSaveSnapshottingCriterion(SnapshottingCriterion):
    """Never saves on its own; GPEmulator still saves the best model of each restart through ``save``."""

    def _should_save(self) -> bool:
        return False
