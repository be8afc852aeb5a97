"""Config-5 sized fit (N_train = 26214, D = 16, RBF): one MLL + gradient evaluation and a predict, validated against
cuSOLVER/cuBLAS (checker only); run under gpurun."""
import json
import math
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200.engine import ExactGPEngine  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
F64 = torch.float64
N, D = (int(sys.argv[1]) if len(sys.argv) > 1 else 26214), 16
g = torch.Generator().manual_seed(0)
X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
y = torch.randn(N, dtype=F64, generator=g).to(dev)
ls = torch.tensor([0.8 * (1 + d / D) for d in range(D)], dtype=F64, device=dev)
theta = torch.cat([1.0 / ls, torch.tensor([1.0, 1e-2], dtype=F64, device=dev)])
eng = ExactGPEngine(X, O.KIND_RBF)
res = {"N": N, "D": D}
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.factorize(theta, y)
    gr = eng.gradients().clone()
    b.record()
    torch.cuda.synchronize()
    res["mll_grad_ms"] = a.elapsed_time(b)
res["loss"] = float(eng.loss)
res["k_inverse"] = f"int8x{eng.lauum_slices}" if eng.lauum_slices else "fp64 DMMA"
res["mll_grad_tflops"] = N ** 3 / res["mll_grad_ms"] / 1e9
if eng.lauum_slices:        # the all-FP64 edition beside it
    keep, eng.lauum_slices = eng.lauum_slices, 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.factorize(theta, y)
    gr64 = eng.gradients().clone()
    b.record()
    torch.cuda.synchronize()
    res["mll_grad_ms_all_fp64"] = a.elapsed_time(b)
    res["grad_rel_diff_int8_vs_fp64_lauum"] = float((gr - gr64).abs().max() / gr64.abs().max())
    eng.lauum_slices = keep
# checker: cuSOLVER Cholesky of the same matrix, built in row blocks to bound memory
K = torch.empty(N, N, dtype=F64, device=dev)
for lo in range(0, N, 2048):
    K[lo:lo + 2048] = O.kernel_from_r2(O.sq_dist(X[lo:lo + 2048], X, ls), O.KIND_RBF, 1.0)
K.diagonal().add_(1e-2)
t0 = time.perf_counter()
L = torch.linalg.cholesky(K)
torch.cuda.synchronize()
res["cusolver_potrf_ms"] = 1e3 * (time.perf_counter() - t0)
del K
alpha = torch.cholesky_solve(y.reshape(-1, 1), L).reshape(-1)
loss_ref = (0.5 * y @ alpha + torch.log(torch.diagonal(L)).sum() + 0.5 * N * math.log(2 * math.pi)) / N
res["loss_rel_err"] = abs(res["loss"] - float(loss_ref)) / abs(float(loss_ref))
res["alpha_rel_err"] = float((eng.alpha[:N] - alpha).abs().max() / alpha.abs().max())
res["chol_rel_err"] = float((eng.chol() - L).abs().max() / L.abs().max())
M = 65536
Xs = torch.rand(M, D, dtype=F64, device=dev)
pm = torch.zeros(M, dtype=F64, device=dev)
for prec in ("fp64", "int8x6", "int8x5w", "tf32x3"):
    if prec == "int8x5w":
        res["int8x5w_sums_exact"] = bool(eng.i8_sums_exact(prec))
        res["int8x5w_digit_bound"] = eng._i8_bound.get(prec, (None, None, None))[2]
    eng.predict(Xs[:4096], prior_mean=pm[:4096], precision=prec)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    mean, std = eng.predict(Xs, prior_mean=pm, precision=prec)
    b.record()
    torch.cuda.synchronize()
    res[f"predict_{prec}_pts_per_s"] = M / a.elapsed_time(b) * 1e3
    if prec == "fp64":
        m64, s64 = mean.clone(), std.clone()
        Ks = O.kernel_from_r2(O.sq_dist(Xs[:2048], X, ls), O.KIND_RBF, 1.0)
        V = torch.linalg.solve_triangular(L, Ks.T, upper=False)
        vref = (1.0 + 1e-2 - (V * V).sum(0)).clamp_min(1e-10)
        res["pred_std_rel_err"] = float(((std[:2048] - vref.sqrt()).abs() / vref.sqrt()).max())
        res["pred_mean_rel_err"] = float((mean[:2048] - Ks @ alpha).abs().max() / (Ks @ alpha).abs().max())
    else:
        res[f"{prec}_std_rel_err"] = float(((std - s64).abs() / s64).max())
res["gpu_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(res))
