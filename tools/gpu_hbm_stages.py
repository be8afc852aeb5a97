"""One MLL + gradient evaluation, one 32768-point predict and one history-matching epilogue, for an ncu capture of the
elementwise / reduction stages (kmat_sym, mll_grad_tile/final, trmv, mll_value, predict_finalize, implausibility)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv
from gperks_b200.engine import ExactGPEngine
dev = torch.device("cuda:0")
F64 = torch.float64
N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 8
g = torch.Generator().manual_seed(0)
X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
y = torch.randn(N, dtype=F64, generator=g).to(dev)
ls = torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=F64, device=dev)
theta = torch.cat([1.0 / ls, torch.tensor([1.0, 1e-2], dtype=F64, device=dev)])
eng = ExactGPEngine(X, 3)
for _ in range(2):
    eng.factorize(theta, y)
    gr = eng.gradients()
Xs = torch.rand(32768, D, dtype=F64, generator=g).to(dev)
mu, sd = eng.predict(Xs)
E, M = 8, 4_000_000
means = torch.randn(E, M, dtype=F64, device=dev)
stds = torch.rand(E, M, dtype=F64, device=dev) + 0.1
z = torch.zeros(E, dtype=F64, device=dev)
v = torch.full((E,), 0.01, dtype=F64, device=dev)
Id = torch.empty(M, dtype=F64, device=dev)
PVd = torch.empty(M, dtype=F64, device=dev)
for _ in range(2):
    nv.check(nv.lib().gpx_implausibility(nv.ptr(means), nv.ptr(stds), M, E, nv.ptr(z), nv.ptr(v), 1, nv.ptr(Id), nv.ptr(PVd),
                                         nv.stream_ptr()), "gpx_implausibility")
torch.cuda.synchronize()
print("done", float(eng.loss), float(mu[0]), float(sd[0]), float(Id[0]))
