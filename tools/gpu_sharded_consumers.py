"""History matching and Saltelli GSA sharded over the GPUs of one node (run under `gpurun --gpus 2` with torchrun):
every rank must reproduce the single-GPU result bit for bit."""
import os
import sys
from functools import partial
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import logging  # noqa: E402
logging.disable(logging.CRITICAL)
from gperks_b200.gp.data.dataset import Dataset  # noqa: E402
from gperks_b200.gp.experiment import GPExperiment  # noqa: E402
from gperks_b200.gp.mean import LinearMean  # noqa: E402
from gperks_b200.kernels import MaternKernel, ScaleKernel  # noqa: E402
from gperks_b200.likelihoods import GaussianLikelihood  # noqa: E402
from gperks_b200.perks.gsa import SobolGSA  # noqa: E402
from gperks_b200.perks.history_matching import Wave  # noqa: E402
from gperks_b200.train.emulator import GPEmulator  # noqa: E402
from gperks_b200.utils.test_functions_gsa import SobolGstar  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")


def emulator(d, n, seed):
    f = partial(SobolGstar, a=([0, 1, 4.5, 9, 99, 99][:d]), delta=np.linspace(0.1, 0.9, d), alpha=[1.0] * d)
    ds = Dataset.build_from_function(f, d, n, None, 32, "sobol", seed, l_bounds=[0.0] * d, u_bounds=[1.0] * d)
    torch.manual_seed(seed)
    exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, d), ScaleKernel(MaternKernel(ard_num_dims=d)), n_restarts=1, seed=seed)
    em = GPEmulator(exp, "cpu")          # parameters on the host: the math runs on this rank's current GPU
    em.model.likelihood.noise = 1e-2
    em.model.covar_module.base_kernel.lengthscale = torch.tensor([0.5 + 0.1 * i for i in range(d)], dtype=torch.float64)
    return em, ds


ems = [emulator(4, 512, 8 + i)[0] for i in range(3)]
X = np.random.default_rng(1).random((300_001, 4))
W = Wave(emulator=ems, Itrain=np.array([[0.0, 1.0]] * 4), cutoff=3.0, maxno=2, mean=[1.0, 0.9, 1.1], var=[0.05] * 3)
I1, PV1 = W.compute_impl(X)
I2, PV2 = W.compute_impl_sharded(X if rank == 0 else None, all_ranks=True)
assert np.array_equal(I1, I2) and np.array_equal(PV1, PV2)

em, ds = emulator(4, 512, 8)
g1 = SobolGSA(ds, n=2 ** 14, seed=8)
g1.estimate_Sobol_indices_with_emulator_mean(em, group=None if dist.get_world_size() == 1 else dist.group.WORLD)
g0 = SobolGSA(ds, n=2 ** 14, seed=8)
dist.barrier()
# single-GPU reference of the same design: temporarily pretend there is no group
_real = dist.is_initialized
dist.is_initialized = lambda: False
g0.estimate_Sobol_indices_with_emulator_mean(em)
dist.is_initialized = _real
assert np.array_equal(g0.ST, g1.ST) and np.array_equal(g0.S1, g1.S1), (g0.ST, g1.ST)
dist.barrier()
if rank == 0:
    print("world", dist.get_world_size(), "HM + GSA sharded == single GPU, bitwise; ST", np.round(g1.ST[0], 4).tolist())
    print("SHARDED-OK")
dist.destroy_process_group()
