"""GPExperiment: wiring of dataset -> scalers -> ExactGPModel (GPErks/gp/experiment.py:67-104) plus the
``.ini`` round trip (``save_to_config_file`` / ``load_experiment_from_config_file``, :106-160)."""
import importlib
from ast import literal_eval
from configparser import ConfigParser
from typing import List, Optional

from ..constants import DEFAULT_EXP_N_RESTARTS
from ..utils.random import set_seed
from .data.data_scaler import StandardScaler, UnitCubeScaler
from .data.dataset import Dataset
from .data.scaled_data import ScaledData
from .model import ExactGPModel


class GPExperiment:
    def __init__(self, dataset: Dataset, likelihood, mean_module, covar_module, *,
                 n_restarts: int = DEFAULT_EXP_N_RESTARTS, seed: Optional[int] = None,
                 metrics: Optional[List] = None, learn_noise: bool = True):
        set_seed(seed)   # first, so everything below is reproducible
        self.seed = seed
        self.dataset = dataset
        self.scaled_data = ScaledData(dataset, UnitCubeScaler(), StandardScaler())
        self.likelihood = likelihood
        self.mean_module = mean_module
        self.covar_module = covar_module
        self.n_restarts = n_restarts
        self.metrics = metrics or []
        self.learn_noise = learn_noise
        self.model = ExactGPModel(self.scaled_data.X_train, self.scaled_data.y_train, likelihood, mean_module,
                                  covar_module)

    # -- configuration files ---------------------------------------------------------------------
    def save_to_config_file(self, experiment_file_path):
        cfg = ConfigParser()
        cfg["GPExperiment"] = {"n_restarts": str(self.n_restarts), "seed": str(self.seed),
                               "learn_noise": str(self.learn_noise)}
        for i, metric in enumerate(self.metrics):
            cfg[f"Metric_{i}"] = _dump(metric)
        cfg["Likelihood"] = _dump(self.model.likelihood)
        cfg["Mean"] = _dump(self.model.mean_module)
        cfg["Kernel"] = _dump(self.model.covar_module.base_kernel)
        with open(experiment_file_path, "w") as f:
            cfg.write(f)


def _dump(obj) -> dict:
    cls = type(obj)
    out = {"class": f"{cls.__module__}.{cls.__qualname__}"}
    name = cls.__name__
    if name == "LinearMean" and hasattr(obj, "degree"):
        out.update(degree=str(obj.degree), input_size=str(obj.input_size), bias=str(obj.bias is not None))
    elif name == "LinearMean":
        out.update(input_size=str(obj.input_size), bias=str(obj.bias is not None))
    elif name == "RBFKernel":
        out.update(ard_num_dims=str(obj.ard_num_dims))
    elif name == "MaternKernel":
        out.update(nu=str(obj.nu), ard_num_dims=str(obj.ard_num_dims))
    elif name == "IndependentStandardError":
        out.update(ci=str(obj.ci))
    return out


# reference configs name gpytorch / torchmetrics / GPErks classes; map them onto this package
_ALIASES = {
    "gpytorch.kernels.rbf_kernel.RBFKernel": "gperks_b200.kernels.RBFKernel",
    "gpytorch.kernels.matern_kernel.MaternKernel": "gperks_b200.kernels.MaternKernel",
    "gpytorch.likelihoods.gaussian_likelihood.GaussianLikelihood": "gperks_b200.likelihoods.GaussianLikelihood",
    "gpytorch.means.linear_mean.LinearMean": "gperks_b200.means.LinearMean",
    "gpytorch.means.constant_mean.ConstantMean": "gperks_b200.means.ConstantMean",
    "gpytorch.means.zero_mean.ZeroMean": "gperks_b200.means.ZeroMean",
    "GPErks.gp.mean.LinearMean": "gperks_b200.gp.mean.LinearMean",
    "GPErks.utils.metrics.IndependentStandardError": "gperks_b200.utils.metrics.IndependentStandardError",
    "torchmetrics.regression.mse.MeanSquaredError": "gperks_b200.utils.metrics.MeanSquaredError",
    "torchmetrics.regression.r2.R2Score": "gperks_b200.utils.metrics.R2Score",
}


def _build(section) -> object:
    kw = dict(section.items())
    path = _ALIASES.get(kw["class"], kw["class"])
    del kw["class"]
    mod, _, cls_name = path.rpartition(".")
    cls = getattr(importlib.import_module(mod), cls_name)
    args = {}
    for k, v in kw.items():
        try:
            args[k] = literal_eval(v)
        except (ValueError, SyntaxError):
            args[k] = v
    return cls(**args)


def load_experiment_from_config_file(experiment_file_path, dataset: Dataset) -> GPExperiment:
    from ..kernels import ScaleKernel

    cfg = ConfigParser()
    if not cfg.read(experiment_file_path):
        raise FileNotFoundError(experiment_file_path)
    exp = cfg["GPExperiment"]
    seed = None if exp.get("seed") == "None" else int(exp.get("seed"))
    set_seed(seed)
    metrics = [_build(cfg[s]) for s in sorted((s for s in cfg.sections() if s.startswith("Metric_")),
                                              key=lambda s: int(s.split("_")[1]))]
    likelihood = _build(cfg["Likelihood"])
    mean = _build(cfg["Mean"])
    kernel = ScaleKernel(_build(cfg["Kernel"]))
    return GPExperiment(dataset, likelihood, mean, kernel, n_restarts=exp.getint("n_restarts"), seed=seed,
                        metrics=metrics, learn_noise=exp.getboolean("learn_noise"))
