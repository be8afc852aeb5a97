"""Stage-by-stage numerical self-check + timings of libgpx on a B200 (run under gpurun).

Prints, per problem size, the max relative error of every C-ABI stage against torch float64 on the same
GPU (cuSOLVER/cuBLAS used only as the checker here), plus CUDA-event timings and the FP64 peak probes.
Writes gpurun_out/selfcheck.json.
"""
import json
import math
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from gperks_b200 import _native as nv  # noqa: E402
from gperks_b200.engine import ExactGPEngine  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402

F64 = torch.float64
dev = torch.device("cuda:0")


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def peaks():
    lib = nv.lib()
    sink = torch.zeros(8, dtype=F64, device=dev)
    out = {}
    for name in ("gpx_peak_dmma", "gpx_peak_dfma"):
        fn = getattr(lib, name)
        blocks, iters = 148 * 8, 20000
        flops = fn(blocks, iters, nv.ptr(sink), nv.stream_ptr())
        ms = ev_time(lambda: fn(blocks, iters, nv.ptr(sink), nv.stream_ptr()), reps=3)
        out[name] = flops / ms / 1e9
    n = 8192
    a = torch.randn(n, n, dtype=F64, device=dev)
    b = torch.randn(n, n, dtype=F64, device=dev)
    ms = ev_time(lambda: torch.matmul(a, b), reps=3)
    out["cublas_dgemm_8192_tflops"] = 2 * n ** 3 / ms / 1e9
    a32, b32 = a.float(), b.float()
    torch.backends.cuda.matmul.allow_tf32 = True
    ms = ev_time(lambda: torch.matmul(a32, b32), reps=3)
    out["cublas_tf32_8192_tflops"] = 2 * n ** 3 / ms / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    ms = ev_time(lambda: torch.matmul(a32, b32), reps=3)
    out["cublas_sgemm_8192_tflops"] = 2 * n ** 3 / ms / 1e9
    x = torch.empty(1 << 28, dtype=F64, device=dev)
    y = torch.empty_like(x)
    ms = ev_time(lambda: y.copy_(x), reps=5)
    out["copy_gbs"] = 2 * x.numel() * 8 / ms / 1e6
    return out


def check(N, D, kind, M, noise=1e-2, seed=0, timing=True):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(N, D, dtype=F64, generator=g)
    y = torch.randn(N, dtype=F64, generator=g)
    Xs = torch.rand(M, D, dtype=F64, generator=g) * 1.2 - 0.1
    ls = torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=F64)
    s = torch.tensor(1.3, dtype=F64)
    nz = torch.tensor(noise, dtype=F64)
    w = torch.linspace(-1, 1, D, dtype=F64)
    b = torch.tensor(0.1, dtype=F64)
    res = {"N": N, "D": D, "kind": kind, "M": M, "noise": noise}

    Xd, yd, Xsd = X.to(dev), y.to(dev), Xs.to(dev)
    lsd, sd, nzd = ls.to(dev), s.to(dev), nz.to(dev)
    mean_x = Xd @ w.to(dev) + b.to(dev)
    resid = yd - mean_x
    eng = ExactGPEngine(Xd, kind)
    eng.lauum_slices = 0          # this tool checks the FP64 lauum (full symmetric K^-1)
    theta = torch.cat([1.0 / lsd, sd.reshape(1), nzd.reshape(1)])
    t0 = time.time()
    eng.factorize(theta, resid)
    torch.cuda.synchronize()
    res["factorize_first_s"] = time.time() - t0

    # torch reference on the GPU (checker only)
    Kref = O.kernel_from_r2(O.sq_dist(Xd, Xd, lsd), kind, sd) + nzd * torch.eye(N, dtype=F64, device=dev)
    Lref = torch.linalg.cholesky(Kref)
    # stage: kmat_sym (rebuild K because potrf overwrote it)
    lib = nv.lib()
    Ktmp = torch.empty_like(eng.K)
    nv.check(lib.gpx_kmat_sym(nv.ptr(eng.X), N, D, nv.ptr(eng.theta), kind, nv.ptr(Ktmp), eng.Np, nv.stream_ptr()), "kmat")
    res["kmat_sym"] = relerr(torch.tril(Ktmp[:N, :N]), torch.tril(Kref))
    pad_ok = True
    if eng.Np > N:
        padblk = torch.tril(Ktmp)[N:, :]
        eye = torch.zeros_like(padblk)
        eye[:, N:] = torch.eye(eng.Np - N, dtype=F64, device=dev)
        pad_ok = bool((padblk == eye).all())
    res["pad_identity"] = pad_ok
    res["potrf"] = relerr(eng.chol(), Lref)
    res["info"] = int(eng.info.item())
    Linv_ref = torch.linalg.solve_triangular(Lref, torch.eye(N, dtype=F64, device=dev), upper=False)
    res["trtri"] = relerr(eng.linv(), Linv_ref)
    # mirror check: upper off-diagonal blocks of S hold L^-T
    Sfull = eng.S[:N, :N]
    nbv = (N + 127) // 128
    mirr = 0.0
    for bi in range(nbv):
        for bj in range(bi + 1, nbv):
            blk = Sfull[bi * 128:(bi + 1) * 128, bj * 128:(bj + 1) * 128]
            ref = Linv_ref[bj * 128:(bj + 1) * 128, bi * 128:(bi + 1) * 128].T
            mirr = max(mirr, float((blk - ref).abs().max()))
        if bi > 3:
            break
    res["trtri_mirror_abs"] = mirr
    alpha_ref = torch.cholesky_solve(resid.reshape(-1, 1), Lref).reshape(-1)
    res["alpha"] = relerr(eng.alpha[:N], alpha_ref)
    loss_ref = (0.5 * resid @ alpha_ref + torch.log(torch.diagonal(Lref)).sum() + 0.5 * N * math.log(2 * math.pi)) / N
    res["loss"] = abs(float(eng.loss) - float(loss_ref)) / abs(float(loss_ref))
    res["loss_value"] = float(eng.loss)
    # gradients
    gr = eng.gradients().clone()
    Kinv_ref = torch.cholesky_inverse(Lref)
    res["lauum"] = relerr(eng.W[:N, :N], Kinv_ref)
    if N <= 2048:
        gref = O.neg_mll_analytic_grads(X, y, kind, ls, s, nz, (X @ w + b))
        gvec = torch.cat([gref["lengthscale"], gref["outputscale"].reshape(1), gref["noise"].reshape(1)])
        res["grad"] = relerr(gr.cpu(), gvec)
    # predict
    mean, std = eng.predict(Xsd, poly=None, prior_mean=Xsd @ w.to(dev) + b.to(dev))
    Ks = O.kernel_from_r2(O.sq_dist(Xsd, Xd, lsd), kind, sd)
    mref = Xsd @ w.to(dev) + b.to(dev) + Ks @ alpha_ref
    V = torch.linalg.solve_triangular(Lref, Ks.T, upper=False)
    vref = (sd - (V * V).sum(0) + nzd).clamp_min(1e-10)
    res["pred_mean"] = relerr(mean, mref)
    res["pred_std_rel"] = float(((std - vref.sqrt()).abs() / vref.sqrt()).max())
    res["cross"] = relerr(eng.cross_kernel(Xsd), Ks)

    if timing:
        p = nv.ptr
        st = nv.stream_ptr()
        res["t_kmat_sym_ms"] = ev_time(lambda: lib.gpx_kmat_sym(p(eng.X), N, D, p(eng.theta), kind, p(Ktmp), eng.Np, st))

        def potrf_only():
            eng.K.copy_(Ktmp)
            lib.gpx_potrf(p(eng.K), eng.Np, N, p(eng.S), p(eng.DT), p(eng.logdet_blk), p(eng.info), st)
        t_copy = ev_time(lambda: eng.K.copy_(Ktmp))
        res["t_potrf_ms"] = ev_time(potrf_only) - t_copy
        res["t_trtri_ms"] = ev_time(lambda: lib.gpx_trtri(p(eng.K), p(eng.S), p(eng.DT), p(eng.W), eng.Np, st))
        res["t_lauum_ms"] = ev_time(lambda: lib.gpx_lauum(p(eng.S), p(eng.DT), p(eng.W), eng.Np, st))
        eng.has_kinv = True
        res["t_grad_ms"] = ev_time(lambda: eng.gradients())
        res["t_alpha_ms"] = ev_time(lambda: lib.gpx_solve_alpha(p(eng.S), p(eng.DT), eng.Np, p(eng.r), p(eng.u), p(eng.alpha), st))
        res["t_torch_cholesky_ms"] = ev_time(lambda: torch.linalg.cholesky(Kref))
        pm = Xsd @ w.to(dev) + b.to(dev)
        Mt = M if N < 8192 else 131072
        Xt = torch.rand(Mt, D, dtype=F64, device=dev)
        pmt = Xt @ w.to(dev) + b.to(dev)
        res["t_predict_ms"] = ev_time(lambda: eng.predict(Xt, prior_mean=pmt), reps=2)
        res["predict_M"] = Mt
        res["predict_pts_per_s"] = Mt / res["t_predict_ms"] * 1e3
        flops_pt = eng.Np ** 2
        res["predict_tflops"] = flops_pt * Mt / res["t_predict_ms"] / 1e9
        res["potrf_tflops"] = N ** 3 / 3 / res["t_potrf_ms"] / 1e9
        res["trtri_tflops"] = N ** 3 / 3 / res["t_trtri_ms"] / 1e9
        res["lauum_tflops"] = N ** 3 / 3 / res["t_lauum_ms"] / 1e9
    return res


def main():
    out = {"device": torch.cuda.get_device_name(0)}
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    results = []
    cases = [
        (20, 2, O.KIND_RBF, 25, 1e-2),
        (200, 8, O.KIND_RBF, 1000, 1e-2),
        (300, 3, O.KIND_MATERN12, 500, 1e-2),
        (1000, 5, O.KIND_MATERN32, 3000, 1e-3),
        (2048, 6, O.KIND_MATERN52, 20000, 1e-2),
        (8192, 8, O.KIND_MATERN52, 16384, 1e-2),
    ]
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        cases = cases[:4]
    for (N, D, kind, M, nz) in cases:
        try:
            r = check(N, D, kind, M, noise=nz)
        except Exception as e:  # keep going: we want every stage's status from one GPU call
            r = {"N": N, "D": D, "kind": kind, "error": repr(e)}
        print(json.dumps(r), flush=True)
        results.append(r)
    out["cases"] = results
    try:
        out["peaks"] = peaks()
        print(json.dumps(out["peaks"]), flush=True)
    except Exception as e:
        out["peaks_error"] = repr(e)
    (ROOT / "gpurun_out" / "selfcheck.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
