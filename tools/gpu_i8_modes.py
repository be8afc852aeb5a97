"""Sliced-integer variance modes (tcgen05 kind::i8): accuracy of every mode against the FP64 (DMMA) path and timing of the
whole predict (chunk pipeline on / off), of the tensor-core kernel alone and of the K* builder alone.
usage: gpu_i8_modes.py [--sizes N:D ...] [--noises ...] [--points M] [--tag name]"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv  # noqa: E402
from gperks_b200.engine import I8_MODES, ExactGPEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", nargs="*", default=["2048:6", "4096:12", "8192:8"])
ap.add_argument("--noises", nargs="*", type=float, default=[1e-2, 1e-4])
ap.add_argument("--points", type=int, default=262144)
ap.add_argument("--acc-points", type=int, default=16384)
ap.add_argument("--tag", default="r02")
ap.add_argument("--modes", nargs="*", default=list(I8_MODES))
args = ap.parse_args()
dev = torch.device("cuda:0")
F64 = torch.float64
lib = nv.lib()
out = {"impl": os.environ.get("GPX_I8_IMPL", "persistent"), "cases": []}


def smi_clock():
    import subprocess
    try:
        return float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"],
                                    capture_output=True, text=True, timeout=5).stdout.strip())
    except Exception:
        return None


LAST_CLOCK = [None]


def timed(fn, reps=3, clock=False):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    if clock:                      # sampled while the launches above are still running
        LAST_CLOCK[0] = smi_clock()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for spec in args.sizes:
    N, D = (int(v) for v in spec.split(":"))
    for noise in args.noises:
        g = torch.Generator().manual_seed(N)
        X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
        y = torch.randn(N, dtype=F64, generator=g).to(dev)
        ls = torch.tensor([0.5 * (1 + d / D) for d in range(D)], dtype=F64, device=dev)
        theta = torch.cat([1.0 / ls, torch.tensor([1.0, noise], dtype=F64, device=dev)])
        eng = ExactGPEngine(X, 3)
        eng.factorize(theta, y)
        Ma = args.acc_points
        Xa = torch.rand(Ma, D, dtype=F64, generator=g).to(dev)
        k = min(N, Ma // 4)
        Xa[:k] = X[:k] + 1e-4 * (torch.rand(k, D, dtype=F64, generator=g).to(dev) - 0.5)     # next to training inputs
        Xa[k:2 * k] = X[:k]                                                                   # and on them
        mu0, sd0 = eng.predict(Xa)
        case = {"N": N, "D": D, "noise": noise, "acc_points": Ma}
        for mode in args.modes:
            S, bits = I8_MODES[mode]
            if eng.Np > lib.gpx_i8_max_n(bits):
                continue
            mu, sd = eng.predict(Xa, precision=mode)
            rel = (sd - sd0).abs() / sd0
            case[mode] = {"std_rel_max": float(rel.max()), "std_rel_median": float(rel.median()), "mean_equal": bool(torch.equal(mu, mu0)),
                          "digit_absmax": int(eng.S_i8.to(torch.int32).abs().max())}
        print(json.dumps(case), flush=True)
        out["cases"].append(case)
        if noise != args.noises[0]:
            continue
        M = args.points
        Xt = torch.rand(M, D, dtype=F64, generator=g).to(dev)
        om = torch.empty(M, dtype=F64, device=dev)
        os_ = torch.empty(M, dtype=F64, device=dev)
        for mode in ["fp64"] + args.modes:
            if mode != "fp64" and eng.Np > lib.gpx_i8_max_n(I8_MODES[mode][1]):
                continue
            t = {"N": N, "D": D, "precision": mode, "points": M}
            for ov in (("1", "0") if mode != "fp64" else ("1",)):
                os.environ["GPX_I8_OVERLAP"] = ov
                ms = timed(lambda: eng.predict(Xt, precision=mode, out_mean=om, out_std=os_), reps=2 if mode == "fp64" else 4)
                key = "ms" if mode == "fp64" else ("ms_overlap" if ov == "1" else "ms_serial")
                t[key] = ms
                t[key.replace("ms", "points_per_s")] = M / ms * 1e3
            os.environ["GPX_I8_OVERLAP"] = "1"
            if mode != "fp64":
                S, bits = I8_MODES[mode]
                Mc = eng.pick_chunk(M)
                ws = eng._workspace(int(lib.gpx_predict_i8_workspace_bytes(eng.Np, Mc, S)))
                Kp = ws.data_ptr() + int(lib.gpx_predict_i8_planes_offset(eng.Np, Mc, S, 0))
                nt = int(lib.gpx_i8_row_tiles(eng.Np))
                qp = torch.empty(nt * Mc, dtype=F64, device=dev)
                sched = torch.zeros(2, dtype=torch.int32, device=dev)
                mp = torch.empty(eng.nb * Mc, dtype=F64, device=dev)
                st = nv.stream_ptr()
                t["chunk"] = Mc
                t["mma_kernel_ms"] = timed(lambda: nv.check(lib.gpx_trmm_sumsq_i8(
                    nv.ptr(eng.S_i8), nv.ptr(eng.S_i8_inv_scale), eng.Np, Kp, Mc, S, bits, nv.ptr(eng.theta) + 8 * D, nv.ptr(qp), nv.ptr(sched), st), "mma"),
                    reps=max(3, int(300 / max(t["ms_serial"] * Mc / M, 0.1))), clock=True)
                t["builder_kernel_ms"] = timed(lambda: nv.check(lib.gpx_kmat_cross_i8(
                    nv.ptr(Xt), Mc, nv.ptr(eng.X), N, D, nv.ptr(eng.theta), None, None, 3, nv.ptr(eng.alpha), Kp, Mc, S, bits, eng.Np,
                    nv.ptr(mp), Mc, st), "builder"))
                # co-residency probe: the tensor-core kernel on a high-priority stream and the K* builder (writing the OTHER plane
                # buffer) on the current stream, launched together; ~max(mma, builder) when both are resident on every SM
                Kp1 = ws.data_ptr() + int(lib.gpx_predict_i8_planes_offset(eng.Np, Mc, S, 1))
                hp = torch.cuda.Stream(priority=-1)
                cur = torch.cuda.current_stream()

                def both(order):
                    ev = torch.cuda.Event()
                    ev.record(cur)
                    hp.wait_event(ev)
                    launch_m = lambda: nv.check(lib.gpx_trmm_sumsq_i8(nv.ptr(eng.S_i8), nv.ptr(eng.S_i8_inv_scale), eng.Np, Kp, Mc, S, bits,  # noqa: E731
                                                                      nv.ptr(eng.theta) + 8 * D, nv.ptr(qp), nv.ptr(sched), hp.cuda_stream), "mma")
                    launch_b = lambda: nv.check(lib.gpx_kmat_cross_i8(nv.ptr(Xt), Mc, nv.ptr(eng.X), N, D, nv.ptr(eng.theta), None, None, 3,  # noqa: E731
                                                                      nv.ptr(eng.alpha), Kp1, Mc, S, bits, eng.Np, nv.ptr(mp), Mc, st), "builder")
                    for f in ((launch_m, launch_b) if order == "mb" else (launch_b, launch_m)):
                        f()
                    ev2 = torch.cuda.Event()
                    ev2.record(hp)
                    cur.wait_event(ev2)
                t["concurrent_mma_then_builder_ms"] = timed(lambda: both("mb"), reps=10)
                t["concurrent_builder_then_mma_ms"] = timed(lambda: both("bm"), reps=10)
                t["clusters"] = int(lib.gpx_i8_clusters())
                t["sm_mhz_during_mma_kernel"] = LAST_CLOCK[0]
                t["mma_raw_int8_tops"] = S * (S + 1) / 2 * N * N * Mc / t["mma_kernel_ms"] / 1e9
                t["mma_fp64_equiv_tflops"] = N * N * Mc / t["mma_kernel_ms"] / 1e9
            print(json.dumps(t), flush=True)
            out["cases"].append(t)
        del eng
        torch.cuda.empty_cache()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"{args.tag}_i8_modes_{out['impl']}.json").write_text(json.dumps(out, indent=1))
print("I8-MODES-DONE")
