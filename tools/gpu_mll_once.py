"""Two MLL + gradient evaluations at N (default 8192), for ncu captures of the Cholesky / inverse kernels.
usage: gpu_mll_once.py [N] [D] [train]     (train: training mode -- integer-tensor-core Cholesky updates and inverse)"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200.engine import ExactGPEngine
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 8
g = torch.Generator().manual_seed(0)
X = torch.rand(N, D, dtype=torch.float64, generator=g).to(dev)
y = torch.randn(N, dtype=torch.float64, generator=g).to(dev)
ls = torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=torch.float64, device=dev)
theta = torch.cat([1.0 / ls, torch.tensor([1.0, 1e-2], dtype=torch.float64, device=dev)])
eng = ExactGPEngine(X, 3)
eng.approx_inverse_ok = len(sys.argv) > 3 and sys.argv[3] == "train"
for _ in range(2):
    eng.factorize(theta, y)
    gr = eng.gradients()
torch.cuda.synchronize()
print("done", float(eng.loss), gr[:3].tolist())
