"""gperks_b200 -- B200-native (sm_100a) implementation of GPErks' exact-GP emulator path.

Drop-in surface: ``gperks_b200.gp.experiment.GPExperiment``, ``gperks_b200.train.emulator.GPEmulator``,
gpytorch-compatible ``kernels`` / ``means`` / ``likelihoods`` / ``mlls`` / ``constraints`` modules, and the
``perks`` consumers.  All GP arithmetic runs in hand-written CUDA kernels behind the C ABI in
``include/gpx.h``; there is no CPU fallback."""
from . import constraints, kernels, likelihoods, means, mlls, models, settings  # noqa: F401

__version__ = "0.1.0"
