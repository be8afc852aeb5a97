"""TF32-mode (tcgen05 3xTF32) predict vs the FP64 path on the same inputs; run under gpurun."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv  # noqa: E402
from gperks_b200.engine import ExactGPEngine  # noqa: E402

F64 = torch.float64
dev = torch.device("cuda:0")


def run(N, D, M, kind=3, noise=1e-2):
    g = torch.Generator().manual_seed(N)
    X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
    y = torch.randn(N, dtype=F64, generator=g).to(dev)
    Xs = torch.rand(M, D, dtype=F64, generator=g).to(dev)
    ls = torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=F64, device=dev)
    theta = torch.cat([1.0 / ls, torch.tensor([1.0, noise], dtype=F64, device=dev)])
    eng = ExactGPEngine(X, kind)
    eng.factorize(theta, y)
    pm = torch.zeros(M, dtype=F64, device=dev)
    m64, s64 = eng.predict(Xs, prior_mean=pm)
    m32, s32 = eng.predict(Xs, prior_mean=pm, precision="tf32x3")
    torch.cuda.synchronize()
    rel = ((s32 - s64).abs() / s64)
    out = {"N": N, "D": D, "M": M, "noise": noise, "mean_rel_max": float((m32 - m64).abs().max() / m64.abs().max()),
           "std_rel_max": float(rel.max()), "std_rel_median": float(rel.median())}
    for prec in ("fp64", "tf32x3"):
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.predict(Xs, prior_mean=pm, precision=prec)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out[f"{prec}_pts_per_s"] = M / min(ts) * 1e3
        out[f"{prec}_useful_tflops"] = eng.Np ** 2 * M / min(ts) / 1e9
    return out


if __name__ == "__main__":
    cases = [(512, 4, 1024), (2048, 6, 16384), (8192, 8, 131072), (8192, 8, 131072, 3, 1e-4), (4096, 12, 131072, 0)]
    for c in cases:
        try:
            print(json.dumps(run(*c)), flush=True)
        except Exception as e:
            print("ERROR", c, repr(e), flush=True)
            break
