"""Small-shape tour of every hand-written kernel family for compute-sanitizer (memcheck / racecheck / synccheck; one tool per
run):  two-stream Cholesky + inverse + gradient (FP64 DMMA and integer tensor-core lauum), predict on the FP64, TF32 and both
sliced-integer paths (persistent clustered tcgen05 kernel, two-stream chunk pipeline), the multi-emulator wave, the device
compaction and the Saltelli sums.  Prints SANITIZE-TARGET-OK when every result is finite and the paths agree."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv  # noqa: E402
from gperks_b200.engine import ExactGPEngine  # noqa: E402

dev = torch.device("cuda:0")
F64 = torch.float64
lib = nv.lib()
g = torch.Generator().manual_seed(0)
for N, D, kind in ((300, 3, 3), (1100, 4, 0)):
    X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
    y = torch.randn(N, dtype=F64, generator=g).to(dev)
    theta = torch.cat([torch.full((D,), 2.0, dtype=F64), torch.tensor([1.0, 1e-2], dtype=F64)]).to(dev)
    eng = ExactGPEngine(X, kind)
    eng.factorize(theta, y)
    grad = eng.gradients().clone()
    assert torch.isfinite(grad).all() and int(eng.info) == 0
    Xs = torch.rand(700, D, dtype=F64, generator=g).to(dev)
    m0, s0 = eng.predict(Xs, precision="fp64")
    for prec in ("int8x5w", "int8x6w", "int8x6", "tf32x3"):
        m1, s1 = eng.predict(Xs, precision=prec, chunk=256)          # three chunks: the two-stream pipeline runs
        tol = 1e-3 if prec == "tf32x3" else 1e-6
        assert torch.equal(m0, m1) and float(((s1 - s0).abs() / s0).max()) < tol, prec
    mo = eng.predict_mean(Xs)
    assert torch.isfinite(mo).all()
    torch.cuda.synchronize()
    print("engine ok", N, D, kind, flush=True)

# compaction + Saltelli sums
vals = torch.rand(10_000, dtype=F64, device=dev)
idx = torch.empty(10_000, dtype=torch.int64, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
ws = torch.empty(int(lib.gpx_compact_below_workspace_bytes(10_000)), dtype=torch.uint8, device=dev)
nv.check(lib.gpx_compact_below(nv.ptr(vals), 10_000, 0.3, nv.ptr(idx), nv.ptr(cnt), nv.ptr(ws), ws.numel(), nv.stream_ptr()), "compact")
assert int(cnt) == int((vals < 0.3).sum())
n, d, K = 5000, 3, 16
JC = lib.gpx_saltelli_chunk_rows()
nc = (n + JC - 1) // JC
f = [torch.randn(n, dtype=F64, device=dev), torch.randn(n, dtype=F64, device=dev), torch.randn(n, d, dtype=F64, device=dev)]
sd = [torch.rand(n, dtype=F64, device=dev), torch.rand(n, dtype=F64, device=dev), torch.rand(n, d, dtype=F64, device=dev)]
part = torch.zeros(nc * K * (4 + 3 * d), dtype=F64, device=dev)
sums = torch.zeros(K * (4 + 3 * d), dtype=F64, device=dev)
nv.check(lib.gpx_saltelli_partial(*[nv.ptr(t) for t in f + sd], n, d, 0, n, K, 7, nv.ptr(part), nv.stream_ptr()), "saltelli")
nv.check(lib.gpx_saltelli_reduce(nv.ptr(part), nc, d, K, nv.ptr(sums), nv.stream_ptr()), "reduce")
assert torch.isfinite(sums).all()
torch.cuda.synchronize()
print("SANITIZE-TARGET-OK", flush=True)
