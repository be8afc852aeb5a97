"""One-off sweep over sizes at the tile-shape / schedule selection boundaries (block rows 18|19, 37|38, 74|75, switch point 32|33,
OB 95|96): potrf, L^-1, K^-1, alpha and the loss against cuSOLVER / cuBLAS on the same device."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200.engine import ExactGPEngine
dev = torch.device("cuda:0")
F64 = torch.float64
sizes = [int(a) for a in sys.argv[1:]] or [2304, 2305, 2433, 4096, 4097, 4736, 4737, 4865, 9472, 9473, 9601, 12160, 12161, 12289]
worst = 0.0
for N in sizes:
    D = 5
    g = torch.Generator().manual_seed(N)
    X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
    y = torch.randn(N, dtype=F64, generator=g).to(dev)
    ls = torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=F64, device=dev)
    theta = torch.cat([1.0 / ls, torch.tensor([1.2, 1e-2], dtype=F64, device=dev)])
    eng = ExactGPEngine(X, 3)
    eng.lauum_slices = 0          # this tool checks the FP64 lauum (full symmetric K^-1)
    eng.factorize(theta, y)
    eng.gradients()
    K = torch.empty(N, N, dtype=F64, device=dev)
    a = X / ls
    for lo in range(0, N, 2048):
        r = torch.cdist(a[lo:lo + 2048], a).clamp_min(0)
        K[lo:lo + 2048] = 1.2 * (1 + np.sqrt(5) * r + 5.0 / 3.0 * r * r) * torch.exp(-np.sqrt(5) * r)
    K.diagonal().add_(1e-2)
    L = torch.linalg.cholesky(K)
    del K
    e_l = float((eng.chol() - L).abs().max() / L.abs().max())
    alpha = torch.cholesky_solve(y.reshape(-1, 1), L).reshape(-1)
    e_a = float((eng.alpha[:N] - alpha).abs().max() / alpha.abs().max())
    Kinv = torch.cholesky_inverse(L)
    e_k = float((eng.W[:N, :N] - Kinv).abs().max() / Kinv.abs().max())
    del Kinv
    Linv = torch.linalg.solve_triangular(L, torch.eye(N, dtype=F64, device=dev), upper=False)
    e_i = float((eng.linv() - Linv).abs().max() / Linv.abs().max())
    del Linv, L
    # cdist-based K differs from the direct-difference kernel by rounding of r: compare at 1e-7, not 1e-11
    worst = max(worst, e_l, e_a, e_k, e_i)
    print(f"N={N:6d} nb={eng.nb:3d}  chol {e_l:.1e}  alpha {e_a:.1e}  Kinv {e_k:.1e}  Linv {e_i:.1e}", flush=True)
    del eng
    torch.cuda.empty_cache()
print("WORST", worst)
assert worst < 1e-6
print("BOUNDARY-OK")
