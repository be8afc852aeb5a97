"""Short program for ncu: the headline predict (BASELINE configs[1], default precision) over two chunks of 32768 points, i.e. two
launches each of the K* builder, the persistent integer tensor-core kernel and the finalize kernel, plus one MLL + gradient
evaluation in training mode.  usage: ncu_target.py [points]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import build_emulator, synthetic_problem  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
X, y, Xs, hyper = synthetic_problem(m=M)
em = build_emulator(X, y, hyper, "cuda:0")
mean, std = em.predict(Xs)
eng = em.model._get_engine()
eng.approx_inverse_ok = True
eng.factorize(eng.theta_in.clone(), eng.r[: eng.N].clone())
eng.gradients()
eng.approx_inverse_ok = False
torch.cuda.synchronize()
print("NCU-TARGET-OK", float(mean[0]), float(std[0]), em.resolve_precision("auto"))
