"""BASELINE configs[3] (history matching: 8 emulators, N=2048, D=6, 5e7 points) through the product's multi-GPU API,
Wave.compute_impl_sharded, under torchrun: X lives on rank 0 (host numpy), its rows are scattered over the ranks, every rank scores its
slice with all 8 emulators resident, (I, PV) are gathered back to rank 0.  Wall clock on rank 0 between barriers; a slice is compared
bit for bit with the single-GPU Wave.compute_impl.  usage (under torchrun): gpu_wave_config4_sharded.py [points]"""
import json
import os
import sys
import time
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
from gperks_b200.perks.history_matching import Wave  # noqa: E402
from gperks_b200.train.emulator import GPEmulator  # noqa: E402
from gperks_b200.utils.random import set_seed  # noqa: E402
from gperks_b200.utils.test_functions_gsa import SobolGstar  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
d, n = 6, 2048


def emulator(seed):
    set_seed(seed)      # before the modules are built, as GPErks' examples do: LinearMean draws its weights at construction, and a fresh
                        # process starts from a random torch seed -- without this the first emulator differs from rank to rank
    f = partial(SobolGstar, a=[0, 1, 4.5, 9, 99, 99], delta=np.linspace(0.1, 0.9, d), alpha=[1.0] * d)
    ds = Dataset.build_from_function(f, d, n, None, 64, "sobol", seed, l_bounds=[0.0] * d, u_bounds=[1.0] * d)
    exp = GPExperiment(ds, GaussianLikelihood(), LinearMean(1, d), ScaleKernel(MaternKernel(ard_num_dims=d)), n_restarts=1, seed=seed)
    em = GPEmulator(exp, f"cuda:{local}")
    m = em.model
    m.likelihood.noise = 1e-2
    m.covar_module.outputscale = 1.0
    m.covar_module.base_kernel.lengthscale = torch.tensor([0.5 * (1 + i / d) for i in range(d)], dtype=torch.float64)
    return em


ems = [emulator(8 + i) for i in range(8)]
W = Wave(emulator=ems, Itrain=np.array([[0.0, 1.0]] * d), cutoff=3.0, maxno=1, mean=[1.0] * 8, var=[0.05] * 8)
X = np.random.default_rng(1).random((M, d)) if rank == 0 else None
warm = np.random.default_rng(2).random((400_000, d))
W.compute_impl(warm)                                    # every rank: factorise, slice, first launches
W.compute_impl_sharded(warm if rank == 0 else None)     # communicator warm-up
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
out = W.compute_impl_sharded(X)
torch.cuda.synchronize()
dist.barrier()
wall = time.perf_counter() - t0
if rank == 0:
    I, PV = out
    I1, PV1 = W.compute_impl(X[:300_000])
    lo = M // world - 100_000                           # around the first shard boundary
    I2, PV2 = W.compute_impl(X[lo:lo + 200_000])
    same = bool(np.array_equal(I[:300_000], I1) and np.array_equal(PV[:300_000], PV1)
                and np.array_equal(I[lo:lo + 200_000], I2) and np.array_equal(PV[lo:lo + 200_000], PV2))
    if not same:
        for name, a, b in (("I head", I[:300_000], I1), ("PV head", PV[:300_000], PV1), ("I boundary", I[lo:lo + 200_000], I2),
                           ("PV boundary", PV[lo:lo + 200_000], PV2)):
            bad = np.nonzero(a != b)[0]
            print(name, "mismatches", bad.size, "first", bad[:5].tolist(), "max rel", float(np.max(np.abs(a - b) / np.abs(b))) if bad.size else 0.0,
                  flush=True)
    res = {"workload": f"configs[3]: 8 emulators N={n} D={d}, {M} points, Wave.compute_impl_sharded over {world} GPUs (X on rank 0's host)",
           "seconds": wall, "points_per_s": M / wall, "emulator_evals_per_s": 8 * M / wall, "n_gpus": world,
           "nimp_fraction": float((I < 3.0).mean()), "bitwise_equal_to_single_gpu_slices": same,
           "timing": "wall clock on rank 0 between barriers: host X -> device, scatter, 8-emulator wave on every rank, gather, device -> host"}
    print(json.dumps(res), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"r02_wave_config4_{world}gpu.json").write_text(json.dumps(res, indent=1))
    assert same
    print("WAVE-SHARDED-OK")
dist.destroy_process_group()
