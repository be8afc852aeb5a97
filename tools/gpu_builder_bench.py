"""K* builder (gpx_kmat_cross_i8) timed alone: `reps` back-to-back launches per case after a warm-up, SM clock sampled while they run.
usage: gpu_builder_bench.py [--sizes N:D ...] [--kinds 0 1 2 3] [--reps 200] [--tag name]"""
import argparse
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", nargs="*", default=["2048:6", "4096:12", "8192:8", "8192:16"])
ap.add_argument("--kinds", nargs="*", type=int, default=[3])
ap.add_argument("--reps", type=int, default=200)
ap.add_argument("--tag", default="r02")
args = ap.parse_args()
dev = torch.device("cuda:0")
F64 = torch.float64
lib = nv.lib()
Mc, S, bits = 32768, 5, 8


def smi_clock():
    try:
        return float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"],
                                    capture_output=True, text=True, timeout=5).stdout.strip())
    except Exception:
        return None


out = []
for spec in args.sizes:
    N, D = (int(v) for v in spec.split(":"))
    Np = (N + 127) // 128 * 128
    g = torch.Generator().manual_seed(N)
    X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
    Xt = torch.rand(Mc, D, dtype=F64, generator=g).to(dev)
    theta = torch.cat([1.0 / torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=F64), torch.tensor([1.0, 1e-2], dtype=F64)]).to(dev)
    alpha = torch.randn(Np, dtype=F64, generator=g).to(dev)
    planes = torch.empty(S * Mc * Np, dtype=torch.int8, device=dev)
    mp = torch.empty((Np // 128) * Mc, dtype=F64, device=dev)
    st = nv.stream_ptr()
    for kind in args.kinds:
        def launch():
            nv.check(lib.gpx_kmat_cross_i8(nv.ptr(Xt), Mc, nv.ptr(X), N, D, nv.ptr(theta), None, None, kind, nv.ptr(alpha), nv.ptr(planes),
                                           Mc, S, bits, Np, nv.ptr(mp), Mc, st), "builder")
        for _ in range(20):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            launch()
        e1.record()
        mhz = smi_clock()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        ent = float(N) * Mc
        row = {"N": N, "D": D, "kind": kind, "ms": ms, "entries_per_s": ent / ms * 1e3, "sm_mhz": mhz,
               "cycles_per_warp_entry_per_smsp": (ms * 1e-3 * (mhz or 1965.0) * 1e6) / (ent / 32 / (148 * 4)),
               "plane_write_GBps": S * ent / ms / 1e6}
        print(json.dumps(row), flush=True)
        out.append(row)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"{args.tag}_builder_bench.json").write_text(json.dumps(out, indent=1))
