"""Short program for ncu: the K* builder + integer tensor-core kernel at a config-4 emulator's shape (N=2048, D=6, Matern-5/2),
two chunks of 32768 points through GPEmulator.predict.  usage: ncu_builder_target.py [N] [D] [points]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import build_emulator, synthetic_problem  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = int(sys.argv[2]) if len(sys.argv) > 2 else 6
M = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
X, y, Xs, hyper = synthetic_problem(n_train=N, dim=D, m=M)
em = build_emulator(X, y, hyper, "cuda:0")
mean, std = em.predict(Xs)
torch.cuda.synchronize()
print("NCU-TARGET-OK", float(mean[0]), float(std[0]), em.resolve_precision("auto"))
