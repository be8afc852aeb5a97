"""potrf only: time (CUDA events, K restore subtracted) and max error against torch.linalg.cholesky, per N.
Usage: python tools/gpu_potrf_time.py [N ...]   (env switches of gpx_chol.cu apply)"""
import json
import os
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv
from gperks_b200.engine import ExactGPEngine

dev = torch.device("cuda:0")
lib = nv.lib()


def ev_time(fn, reps=5):
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


for N in [int(a) for a in sys.argv[1:]] or [8192]:
    D = 8
    g = torch.Generator().manual_seed(0)
    X = torch.rand(N, D, dtype=torch.float64, generator=g).to(dev)
    ls = torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=torch.float64, device=dev)
    theta = torch.cat([1.0 / ls, torch.tensor([1.0, 1e-2], dtype=torch.float64, device=dev)])
    eng = ExactGPEngine(X, 3)
    p, st = nv.ptr, nv.stream_ptr()
    eng.theta.copy_(theta)
    lib.gpx_kmat_sym(p(eng.X), N, D, p(eng.theta), 3, p(eng.K), eng.Np, st)
    K0 = eng.K.clone()

    def potrf_only():
        eng.K.copy_(K0)
        lib.gpx_potrf(p(eng.K), eng.Np, N, p(eng.S), p(eng.DT), p(eng.logdet_blk), p(eng.info), st)
    t_copy = ev_time(lambda: eng.K.copy_(K0))
    t = ev_time(potrf_only) - t_copy
    Kfull = torch.tril(K0[:N, :N]) + torch.tril(K0[:N, :N], -1).T
    Lref = torch.linalg.cholesky(Kfull)
    t_ref = ev_time(lambda: torch.linalg.cholesky(Kfull), reps=3)
    L = torch.tril(eng.K[:N, :N])
    err = float((L - Lref).abs().max() / Lref.abs().max())
    print(json.dumps({"N": N, "t_potrf_ms": round(t, 3), "tflops": round(N ** 3 / 3 / t / 1e9, 2), "t_cusolver_ms": round(t_ref, 3),
                      "max_rel_err": err, "info": int(eng.info.item()),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("GPX_")}}), flush=True)
    del eng, K0, Kfull, Lref, L
