from torch import Tensor

from ...utils.array import tensorize
from .data_scaler import InputDataScaler, OutputDataScaler
from .dataset import Dataset


class ScaledData:
    """Fits the scalers on the training split only and holds the scaled tensors
    (GPErks/gp/data/scaled_data.py:8-30).  Tensors are float64 (the reference keeps float32)."""

    def __init__(self, dataset: Dataset, scx: InputDataScaler, scy: OutputDataScaler):
        self.scx, self.scy = scx, scy
        scx.fit(dataset.X_train)
        scy.fit(dataset.y_train)
        self.X_train: Tensor = tensorize(scx.transform(dataset.X_train))
        self.y_train: Tensor = tensorize(scy.transform(dataset.y_train))
        self.sample_size = dataset.sample_size
        self.input_size = dataset.input_size
        self.with_val = dataset.with_val
        self.X_val, self.y_val = dataset.X_val, dataset.y_val
        if self.with_val:
            self.X_val = tensorize(scx.transform(dataset.X_val))
            self.y_val = tensorize(scy.transform(dataset.y_val))
