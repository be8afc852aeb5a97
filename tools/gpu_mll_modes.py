"""MLL + gradient evaluation (ExactGPEngine.factorize + gradients) timed in training mode (integer-tensor-core Cholesky bulk updates,
inverse merge levels and K_hat^-1) and all-FP64, with the stage breakdown and the deviation of loss / gradients / L between them.
usage: gpu_mll_modes.py [--sizes N:D ...] [--reps 10] [--tag name]      env: GPX_POTRF_OB_I8=1..8, GPX_POTRF=fp64"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv  # noqa: E402
from gperks_b200.engine import ExactGPEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", nargs="*", default=["8192:8"])
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--tag", default="r02")
args = ap.parse_args()
dev = torch.device("cuda:0")
F64 = torch.float64
out = []


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for spec in args.sizes:
    N, D = (int(v) for v in spec.split(":"))
    g = torch.Generator().manual_seed(N)
    X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
    y = torch.randn(N, dtype=F64, generator=g).to(dev)
    ls = torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=F64, device=dev)
    theta = torch.cat([1.0 / ls, torch.tensor([1.0, 1e-2], dtype=F64, device=dev)])
    eng = ExactGPEngine(X, 3)
    lib, p, st = eng.lib, nv.ptr, nv.stream_ptr()
    row = {"N": N, "D": D, "ob_i8": os.environ.get("GPX_POTRF_OB_I8", "default")}

    def full():
        eng.factorize(theta, y, check=False)
        eng.gradients()

    eng.approx_inverse_ok = False
    row["ms_all_fp64_inverse_i8_lauum"] = timed(full, args.reps)
    L64, loss64, g64 = eng.chol().clone(), float(eng.loss), eng.gradients().clone()
    eng.approx_inverse_ok = True
    row["ms_training_mode"] = timed(full, args.reps)
    row["evals_per_s_training_mode"] = 1e3 / row["ms_training_mode"]
    row["potrf_i8"] = not eng.chol_exact
    row["L_rel_diff"] = float((eng.chol() - L64).abs().max() / L64.abs().max())
    row["loss_rel_diff"] = abs(float(eng.loss) - loss64) / abs(loss64)
    row["grad_max_rel_diff"] = float(((eng.gradients() - g64).abs() / g64.abs().clamp_min(1e-300)).max())
    # stages
    def kmat():
        nv.check(lib.gpx_kmat_sym(p(eng.X), N, D, p(eng.theta), 3, p(eng.K), eng.Np, st), "kmat")

    def potrf64():
        kmat()
        nv.check(lib.gpx_potrf(p(eng.K), eng.Np, N, p(eng.S), p(eng.DT), p(eng.logdet_blk), p(eng.info), st), "potrf")

    def potrf8():
        kmat()
        nv.check(lib.gpx_potrf_i8(p(eng.K), eng.Np, N, p(eng.S), p(eng.DT), p(eng.logdet_blk), p(eng.info), p(eng._lauum_ws),
                                  eng._lauum_ws.numel(), st), "potrf_i8")

    row["ms_kmat_sym"] = timed(kmat, args.reps)
    row["ms_potrf_fp64"] = timed(potrf64, args.reps) - row["ms_kmat_sym"]
    if eng._potrf_i8_capable:
        row["ms_potrf_i8"] = timed(potrf8, args.reps) - row["ms_kmat_sym"]
    print(json.dumps(row), flush=True)
    out.append(row)
    del eng
    torch.cuda.empty_cache()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"{args.tag}_mll_modes.json").write_text(json.dumps(out, indent=1))
