"""Early-stopping criteria after Prechelt, "Early Stopping -- But When?" (1997), with the protocol GPErks'
training loop drives once per epoch: ``enable`` -> (``evaluate`` until ``is_verified``)
(GPErks/train/early_stop.py).  Behaviourally equivalent restatement:

  No      never stops early; best epoch = last epoch
  Pk      training progress  P_k = |mean(strip)/min(strip) - 1| below alpha for `patience` epochs
  Simple  validation loss has not improved for `patience` epochs
  GL      generalisation loss GL = 100 |E_va/E_opt - 1| above alpha for `patience` epochs
  UP      validation loss went up over `successive_strips` consecutive strips of `strip_length`
  PQ      GL / (1000 * P_k) above alpha for `patience` epochs
"""
from abc import ABCMeta, abstractmethod
from copy import deepcopy
from typing import List, Optional, Tuple

import numpy

from ..log.logger import get_logger

log = get_logger()


class EarlyStoppingCriterion(metaclass=ABCMeta):
    def __init__(self, max_epochs: int):
        self.max_epochs = max_epochs
        self.model = None
        self.train_stats = None
        self.save_path = None
        self.is_verified = False

    def enable(self, model, train_stats):
        self.model, self.train_stats = model, train_stats
        self.is_verified = False

    def evaluate(self) -> Tuple[Optional[int], Optional[object]]:
        stats = self.train_stats
        evals = stats.early_stopping_criterion_evaluations
        if stats.current_epoch == self.max_epochs:
            # budget exhausted: repeat the last evaluation so the series has one entry per epoch
            evals.append(evals[-1])
            stop = True
        else:
            stop, value = self._should_stop()
            evals.append(value)
        if not stop:
            self._on_continue()
            return None, None
        self.is_verified = True
        outcome = self._on_stop()
        self._reset()
        return outcome

    def _on_continue(self):
        pass

    @abstractmethod
    def _reset(self): ...

    @abstractmethod
    def _should_stop(self) -> Tuple[bool, float]: ...

    @abstractmethod
    def _on_stop(self) -> Tuple[int, object]: ...


class NoEarlyStoppingCriterion(EarlyStoppingCriterion):
    def _reset(self):
        pass

    def _should_stop(self):
        return False, 0

    def _on_stop(self):
        return self.train_stats.current_epoch, self.model


class _PatienceCriterion(EarlyStoppingCriterion):
    """Shared machinery: a counter of consecutive 'bad' epochs and the state dict of the last good one."""

    def __init__(self, max_epochs: int, patience: int):
        super().__init__(max_epochs)
        self.patience = patience
        self.counter = 0
        self.current_best_state_dict = {}

    def _remember(self):
        self.current_best_state_dict = deepcopy(self.model.state_dict())

    def _lookback(self) -> int:
        return self.patience

    def _on_stop(self):
        self.model.load_state_dict(self.current_best_state_dict)
        return self.train_stats.current_epoch - self._lookback(), self.model

    def _tick(self, bad: bool) -> bool:
        if bad:
            self.counter += 1
        else:
            self.counter = 0
            self._remember()
        return self.counter == self.patience


def _progress(train_loss: List[float], k: int) -> float:
    strip = train_loss[-k:]
    return float(numpy.sum(strip) / (k * numpy.min(strip)) - 1)


class PkEarlyStoppingCriterion(_PatienceCriterion):
    """Uses training data only."""

    def __init__(self, max_epochs: int, alpha: float, patience: int, strip_length: int):
        super().__init__(max_epochs, patience)
        self.alpha = alpha
        self.strip_length = strip_length
        self._reset()

    def _reset(self):
        self.counter = 0
        self.strip_counter = 0
        self.computation_enabled = False
        self.Pk: List[float] = []

    def _should_stop(self):
        self.strip_counter += 1
        if self.strip_counter % self.strip_length == 0:
            self.computation_enabled = True
        self.Pk.append(abs(_progress(self.train_stats.train_loss, self.strip_length))
                       if self.computation_enabled else 0.0)
        return self._tick(self.computation_enabled and self.Pk[-1] < self.alpha), self.Pk[-1]


class SimpleEarlyStoppingCriterion(_PatienceCriterion):
    """Uses validation data."""

    def __init__(self, max_epochs: int, patience: int):
        super().__init__(max_epochs, patience)
        self.Eva_opt = numpy.inf

    def _reset(self):
        self.counter = 0
        self.Eva_opt = numpy.inf

    def _should_stop(self):
        current = self.train_stats.val_loss[-1]
        if current < self.Eva_opt:
            self.Eva_opt = current
            self._remember()
            self.counter = 0
        else:
            self.counter += 1
        return self.counter == self.patience, current


class GLEarlyStoppingCriterion(_PatienceCriterion):
    """Uses validation data."""

    def __init__(self, max_epochs: int, alpha: float, patience: int):
        super().__init__(max_epochs, patience)
        self.alpha = alpha
        self._reset()

    def _reset(self):
        self.counter = 0
        self.Eva_opt = numpy.inf
        self.GL: List[float] = []

    def _generalisation_loss(self) -> float:
        current = self.train_stats.val_loss[-1]
        self.Eva_opt = min(self.Eva_opt, current)
        self.GL.append(float(numpy.abs(100 * (current / self.Eva_opt - 1))))
        return self.GL[-1]

    def _should_stop(self):
        gl = self._generalisation_loss()
        return self._tick(gl > self.alpha), gl


class UPEarlyStoppingCriterion(_PatienceCriterion):
    """Uses validation data."""

    def __init__(self, max_epochs: int, strip_length: int, successive_strips: int):
        super().__init__(max_epochs, successive_strips)
        self.strip_length = strip_length
        self.successive_strips = successive_strips
        self._reset()

    def _reset(self):
        self.strip_counter = 0
        self.computation_enabled = False
        self.sstrips_counter = 0
        self.UP: List[float] = []

    def _lookback(self) -> int:
        return self.strip_length + self.successive_strips - 1

    def _should_stop(self):
        self.strip_counter += 1
        if self.strip_counter % self.strip_length == 0:
            self.computation_enabled = True
        if not self.computation_enabled:
            self.UP.append(0.0)
        else:
            val = self.train_stats.val_loss
            self.UP.append(float(numpy.sign(val[-1] - val[-self.strip_length])))
            if self.UP[-1] == 1:
                self.sstrips_counter += 1
            else:
                self.sstrips_counter = 0
                self._remember()
        return self.sstrips_counter == self.successive_strips, self.UP[-1]


class PQEarlyStoppingCriterion(GLEarlyStoppingCriterion):
    """Uses validation data (quotient of generalisation loss and training progress)."""

    def __init__(self, max_epochs: int, alpha: float, patience: int, strip_length: int):
        self.strip_length = strip_length
        super().__init__(max_epochs, alpha, patience)

    def _reset(self):
        super()._reset()
        self.strip_counter = 0
        self.computation_enabled = False
        self.Pk: List[float] = []
        self.PQ: List[float] = []

    def _should_stop(self):
        gl = self._generalisation_loss()
        self.strip_counter += 1
        if self.strip_counter % self.strip_length == 0:
            self.computation_enabled = True
        if self.computation_enabled:
            self.Pk.append(abs(1000 * _progress(self.train_stats.train_loss, self.strip_length)))
            self.PQ.append(gl / self.Pk[-1])
        else:
            self.Pk.append(0.0)
            self.PQ.append(0.0)
        return self._tick(self.PQ[-1] > self.alpha), self.PQ[-1]
