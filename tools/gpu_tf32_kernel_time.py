"""Kernel-only timing of the tcgen05 TF32 contraction vs accumulation-chunk length (run under gpurun)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv
lib = nv.lib()
dev = torch.device("cuda:0")
Np, Mc = 8192, 32768
Khi = torch.randn(Mc, Np, device=dev); Klo = torch.randn(Mc, Np, device=dev) * 1e-4
Shi = torch.tril(torch.randn(Np, Np, device=dev)); Slo = torch.tril(torch.randn(Np, Np, device=dev)) * 1e-4
nt = lib.gpx_tf32_row_tiles(Np)
q = torch.empty(nt, Mc, dtype=torch.float64, device=dev)
for kc in (32, 64, 128, 256, 1024, 8192):
    ts = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.gpx_trmm_sumsq_tf32(nv.ptr(Shi), nv.ptr(Slo), Np, nv.ptr(Khi), nv.ptr(Klo), Mc, nv.ptr(q), kc, nv.stream_ptr())
        b.record(); torch.cuda.synchronize()
        assert rc == 0
        if i: ts.append(a.elapsed_time(b))
    ms = min(ts)
    # useful FLOPs: triangular, N^2 per point
    print(f"kchunk={kc:5d}  {ms:7.3f} ms  raw TF32 {3*Np*Np*Mc/ms/1e9:7.1f} TF/s  useful {Np*Np*Mc/ms/1e9:6.1f} TF/s", flush=True)
