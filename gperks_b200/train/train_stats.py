import json
from typing import Dict, List


class TrainStats:
    """Per-restart training record, JSON-serialisable (fields as GPErks/train/train_stats.py:10-26)."""

    def __init__(self, metrics_names):
        self.current_epoch: int = 0
        self.best_epoch: int = 0
        self.train_loss: List[float] = []
        self.val_loss: List[float] = []
        self.train_metrics_score: Dict[str, List[float]] = {m: [] for m in metrics_names}
        self.val_metrics_score: Dict[str, List[float]] = {m: [] for m in metrics_names}
        self.early_stopping_criterion_evaluations: List[float] = []

    def save_to_file(self, output_file):
        with open(output_file, "w") as f:
            json.dump(self.__dict__, f)

    def plot(self, with_early_stopping_criterion: bool = False):
        raise NotImplementedError("plotting is outside the B200 hot path (matplotlib not required)")


def load_train_stats_from_file(filename) -> TrainStats:
    stats = TrainStats([])
    with open(filename, "r") as f:
        stats.__dict__.update(json.load(f))
    return stats
