from abc import ABCMeta, abstractmethod


class Trainable(metaclass=ABCMeta):
    @abstractmethod
    def train(self, optimizer, early_stopping_criterion, snapshotting_criterion): ...
