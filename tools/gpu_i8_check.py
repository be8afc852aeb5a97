"""Sliced-integer (tcgen05 kind::i8) variance path: slice reconstruction, the integer contraction against the same slices
multiplied in FP64, predict against the FP64 mode at several sizes, and kernel timing.  usage: gpu_i8_check.py [N ...]"""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv
from gperks_b200.engine import ExactGPEngine
dev = torch.device("cuda:0")
F64 = torch.float64
lib = nv.lib()
sizes = [int(a) for a in sys.argv[1:]] or [200, 1000, 8192]
out = {"cases": []}


def engine(N, D, noise, kind=3):
    g = torch.Generator().manual_seed(N)
    X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
    y = torch.randn(N, dtype=F64, generator=g).to(dev)
    ls = torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=F64, device=dev)
    theta = torch.cat([1.0 / ls, torch.tensor([1.3, noise], dtype=F64, device=dev)])
    eng = ExactGPEngine(X, kind)
    eng.factorize(theta, y)
    return eng, X, theta, g


for N in sizes:
    D = 8
    for noise in (1e-2, 1e-4):
        eng, X, theta, g = engine(N, D, noise)
        Np = eng.Np
        M = 4096 if N >= 4096 else 700
        Xs = torch.rand(M, D, dtype=F64, generator=g).to(dev)
        Xs[:64] = X[:64] + 1e-3 * torch.rand(64, D, dtype=F64, generator=g).to(dev)
        mu0, sd0 = eng.predict(Xs)
        case = {"N": N, "noise": noise, "M": M}
        for S in (5, 6):
            if noise == 1e-2 and S == 5 and N <= 1000:
                # --- unit checks on the slices and the integer contraction ---
                eng._ensure_i8(S)
                planes = eng.S_i8.to(F64)
                w = torch.tensor([2.0 ** (-7 * (j + 1)) for j in range(S)], dtype=F64, device=dev).view(S, 1, 1)
                rec = (planes * w).sum(0) * eng.S_i8_inv_scale.view(-1, 1)
                Sl = torch.tril(eng.S)
                mask = torch.tril(torch.ones(Np, Np, dtype=torch.bool, device=dev))
                case["slice_rec_err"] = float(((rec - Sl).abs() * mask).max() / Sl.abs().max())
                case["slice_absmax"] = int(eng.S_i8.abs().max())
                Mc = 256
                Ks = eng.cross_kernel(Xs[:Mc])                       # [Mc, N]
                Kp_src = torch.zeros(Mc, Np, dtype=F64, device=dev)
                Kp_src[:, :N] = Ks
                Kp = torch.empty(S, Mc, Np, dtype=torch.int8, device=dev)
                amax = theta[D:D + 1].contiguous()
                nv.check(lib.gpx_slice_i8(nv.ptr(Kp_src), Mc, Np, Np, nv.ptr(Kp), Mc, S, 0, None, nv.ptr(amax), nv.stream_ptr()), "slice")
                nt = lib.gpx_i8_row_tiles(Np)
                qpart = torch.full((nt, Mc), -1.0, dtype=F64, device=dev)
                nv.check(lib.gpx_trmm_sumsq_i8(nv.ptr(eng.S_i8), nv.ptr(eng.S_i8_inv_scale), Np, nv.ptr(Kp), Mc, S, nv.ptr(amax),
                                               nv.ptr(qpart), nv.stream_ptr()), "trmm_i8")
                torch.cuda.synchronize()
                # the same slices multiplied in FP64 (exact: integers < 2^53)
                Kpf = Kp.to(F64)
                import math
                ek = math.frexp(float(amax))[1]
                V = torch.zeros(Np, Mc, dtype=F64, device=dev)
                for a in range(S):
                    for b in range(S - a):
                        V += (torch.tril(planes[b]) @ Kpf[a].T) * 2.0 ** (-7 * (a + b + 2))
                V = V * eng.S_i8_inv_scale.view(-1, 1) * 2.0 ** (ek + 1)
                qref = (V * V).view(nt, 64, Mc).sum(1)
                case["qpart_vs_sliced_fp64"] = float((qpart - qref).abs().max() / qref.abs().max())
                Vt = torch.tril(eng.S) @ Kp_src.T
                qtrue = (Vt * Vt).view(nt, 64, Mc).sum(1)
                case["qpart_vs_fp64"] = float((qpart - qtrue).abs().max() / qtrue.abs().max())
            mu, sd = eng.predict(Xs, precision=f"int8x{S}")
            torch.cuda.synchronize()
            case[f"int8x{S}_std_rel"] = float(((sd - sd0).abs() / sd0).max())
            case[f"int8x{S}_mean_equal"] = bool(torch.equal(mu, mu0))
        print(json.dumps(case), flush=True)
        out["cases"].append(case)
        if N >= 1024 and noise == 1e-2:
            Mt = 32768
            Xt = torch.rand(Mt, D, dtype=F64, generator=g).to(dev)
            for prec in ("fp64", "int8x5", "int8x6", "tf32x3"):
                for _ in range(2):
                    eng.predict(Xt, precision=prec)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    eng.predict(Xt, precision=prec)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                t = {"N": N, "precision": prec, "ms_per_32768": ms, "points_per_s": Mt / ms * 1e3, "fp64_equiv_tflops": N * N * Mt / ms / 1e9}
                print(json.dumps(t), flush=True)
                out["cases"].append(t)
        del eng
        torch.cuda.empty_cache()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
import os
(ROOT / "gpurun_out" / f"i8_check_bkb{os.environ.get('GPX_I8_BKB', '64')}.json").write_text(json.dumps(out, indent=1))
print("I8-CHECK-DONE")
