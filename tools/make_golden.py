"""Regenerates tests/golden/notebook_goldens.json.

Run in the build container only (needs /root/reference):  ``python tools/make_golden.py``.

The datasets are re-created with the reference's OWN importable code --
``GPErks.utils.sampling.Sampler`` (GPErks/utils/sampling.py:8-32) and the analytic test functions
(GPErks/utils/test_functions.py, test_functions_gsa.py) -- following
``Dataset.build_from_function`` (GPErks/gp/data/dataset.py:214-264): train, then val, then test are
drawn sequentially from one engine.  The hyper-parameters and the printed scores are transcribed
from the executed notebook cells (file:line given per entry).  ``import GPErks`` itself fails in this
image (pkg_resources lookup + missing gpytorch), so the package is stubbed and only leaf modules
are imported.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)
pkg = types.ModuleType("GPErks")
pkg.__path__ = [f"{REF}/GPErks"]
sys.modules["GPErks"] = pkg

from GPErks.utils.sampling import Sampler  # noqa: E402
from GPErks.utils.test_functions import currin_exp, forrester  # noqa: E402
from GPErks.utils.test_functions_gsa import SobolGstar, SobolGstar_theoretical_Si  # noqa: E402,F401


def build(f, d, n_train, n_val, n_test, design, seed):
    sampler = Sampler(design, d, seed)
    out = {}
    X = sampler.sample(n_train)
    out["X_train"], out["y_train"] = X, np.squeeze(f(X.T))
    if n_val is not None:
        X = sampler.sample(n_val)
        out["X_val"], out["y_val"] = X, np.squeeze(f(X.T))
    if n_test is not None:
        X = sampler.sample(n_test)
        out["X_test"], out["y_test"] = X, np.squeeze(f(X.T))
    return {k: np.asarray(v).tolist() for k, v in out.items()}


def main():
    goldens = {}

    # notebooks/example_2.ipynb: currin_exp, LHS, N=20, test 25, D=2, seed 8;
    # gpytorch LinearMean + ScaleKernel(RBFKernel(ard_num_dims=2)); state_dict :534-553
    goldens["example_2"] = {
        "source": "notebooks/example_2.ipynb:534-553 (state), :580-581 (MSE/R2), :698-700 (chi2/MD/CI), :420 (loss)",
        "data": build(currin_exp, 2, 20, None, 25, "lhs", 8),
        "kernel": "rbf",
        "mean_degree": 1,
        "raw": {
            "raw_noise": -3.6342,
            "noise_lower_bound": 1e-4,
            "weights": [-0.0557, -2.5851],
            "bias": 1.2667,
            "raw_outputscale": -2.0184,
            "raw_lengthscale": [-2.0959, 2.1659],
        },
        "printed": {
            "MSE": 0.7254, "R2": 0.9013, "chi2": 46.6326, "mahalanobis": 44.0926, "CI": 0.92,
            "train_loss": 0.0870, "train_MSE": 0.0153, "train_R2": 0.9847,
        },
    }

    # notebooks/example_9.ipynb: same data; GPErks LinearMean(degree=2); Matern-5/2; train_auto
    goldens["example_9"] = {
        "source": "notebooks/example_9.ipynb:114-118 (hyperparameters, post-softplus), :146-148 (scores)",
        "data": goldens["example_2"]["data"],
        "kernel": "matern52",
        "mean_degree": 2,
        "actual": {
            "noise": 0.0001,
            "weights": [1.2054, -5.6002, -2.2104, 1.0741, 2.5829],
            "bias": 1.7794,
            "outputscale": 0.0472,
            "lengthscale": [0.0564, 101.7413],
        },
        "printed": {"MSE": 0.6714, "R2": 0.9086, "ISE": 0.2400},
    }

    # notebooks/example_1.ipynb: forrester 1-D, srs, N=10, test 10, seed 8; RBF
    goldens["example_1"] = {
        "source": "notebooks/example_1.ipynb:556-575 (state), :599-600 (MSE/R2)",
        "data": build(forrester, 1, 10, None, 10, "srs", 8),
        "kernel": "rbf",
        "mean_degree": 1,
        "raw": {
            "raw_noise": -10.7605,
            "noise_lower_bound": 1e-4,
            "weights": [2.9474],
            "bias": -1.0693,
            "raw_outputscale": 1.7124,
            "raw_lengthscale": [-1.6003],
        },
        "printed": {"MSE": 0.2552940845489502, "R2": 0.8388684249013869},
    }

    out = Path(__file__).resolve().parent.parent / "tests" / "golden" / "notebook_goldens.json"
    out.write_text(json.dumps(goldens, indent=1))
    print("wrote", out)


if __name__ == "__main__":
    main()
