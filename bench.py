#!/usr/bin/env python
"""Headline benchmark: GPEmulator.predict (mean + noisy std) throughput, N_train=8192, D=8, Matern-5/2 ARD
(BASELINE.json configs[1]), on 1..8 B200s of one node.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (float64 oracle, all host threads)

A step = one pass of the predict hot path over M points per GPU (weak scaling: each rank owns its own M-point
shard; factors L^-1 and alpha are replicated; the only collective is the final gather of the 16 B/point
results).  `value` = points all ranks processed / max-over-ranks device time with X* resident in HBM;
`e2e` = the same through GPEmulator.predict with host numpy buffers (pinned H2D + D2H inside the timed
region).  Both run at the precision GPEmulator.predict uses by default for this emulator (`--precision auto`: the
FP64 variance contraction from six 7-bit integer slices on tcgen05 kind::i8, std within 1e-8 of the FP64 DMMA path,
reported under `accuracy`); the FP64 DMMA path itself is timed beside it (`fp64_dmma_mode`, or `--precision fp64` to
make it the headline), as are the TF32 mode and the five-slice edition.  One JSON line on stdout (rank 0)."""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore", message="The balance properties of Sobol")

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_TRAIN, DIM = 8192, 8
NCU_TRAFFIC_I8_BYTES = {"int8x6": 11.877e9}      # dram read+write per launch of trmm_sumsq_i8_kernel<6>, profiles/r01_ncu_i8.txt (N=8192, chunk 32768)
NCU_TRAFFIC_BYTES = 13.546e9   # dram read+write per launch, profiles/r01_ncu_trmm_tma.txt (N=8192, chunk 32768)
KIND_NAME = "matern52"
NOISE = 1e-2


def synthetic_problem(n_train=N_TRAIN, dim=DIM, m=1_000_000, seed=8, rank=0):
    """Seeded synthetic data of BASELINE's shape: X ~ Sobol(seed 8) in [0,1]^D,
    y = Sobol G* standardised, fixed hyper-parameters l_d = 0.5(1+d/D), s = 1, noise = 1e-2,
    linear mean w = linspace(-1,1,D), b = 0.1; X* ~ Sobol(seed 9 + rank)."""
    from scipy.stats import qmc
    X = qmc.Sobol(d=dim, scramble=True, seed=seed).random(n_train)
    a = np.array([0, 1, 4.5, 9, 99, 99, 99, 99][:dim] + [99] * max(0, dim - 8), dtype=float)
    delta = np.random.default_rng(seed).random(dim)
    frac = X + delta - np.floor(X + delta)
    y = np.prod((2.0 * np.abs(2 * frac - 1) + a) / (1 + a), axis=1)
    Xs = qmc.Sobol(d=dim, scramble=True, seed=seed + 1 + rank).random(m)
    hyper = dict(lengthscale=[0.5 * (1 + d / dim) for d in range(dim)], outputscale=1.0, noise=NOISE,
                 weights=np.linspace(-1, 1, dim).tolist(), bias=0.1)
    return X, y, Xs, hyper


def build_emulator(X, y, hyper, device):
    from gperks_b200.gp.data.dataset import Dataset
    from gperks_b200.gp.experiment import GPExperiment
    from gperks_b200.gp.mean import LinearMean
    from gperks_b200.kernels import MaternKernel, ScaleKernel
    from gperks_b200.likelihoods import GaussianLikelihood
    from gperks_b200.train.emulator import GPEmulator
    D = X.shape[1]
    exp = GPExperiment(Dataset(X, y), GaussianLikelihood(), LinearMean(1, D), ScaleKernel(MaternKernel(nu=2.5, ard_num_dims=D)),
                       n_restarts=1, seed=8)
    em = GPEmulator(exp, device)
    t = lambda v: torch.as_tensor(v, dtype=torch.float64)  # noqa: E731
    m = em.model
    m.likelihood.noise = hyper["noise"]
    m.covar_module.outputscale = hyper["outputscale"]
    m.covar_module.base_kernel.lengthscale = t(hyper["lengthscale"])
    m.initialize(**{"mean_module.weights": t(hyper["weights"]), "mean_module.bias": t([hyper["bias"]])})
    return em


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
_ORACLE_CACHE = {}


def cpu_oracle_predict_rate(X, y, Xs, hyper, sample_pts, chunk=4096):
    """The reference's CPU path restated (oracle/gp_oracle.py, torch float64, all host threads): factorise
    once, then predict `sample_pts` points in chunks; returns (pts/s over the predict part, seconds, threads)."""
    from oracle import gp_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    if "em" not in _ORACLE_CACHE:
        em = O.OracleEmulator(X, y, kind=O.KIND_MATERN52, lengthscale=hyper["lengthscale"],
                              outputscale=hyper["outputscale"], noise=hyper["noise"], weights=hyper["weights"],
                              bias=hyper["bias"], degree=1)
        t0 = time.perf_counter()
        resid = em.y - em._mean(em.X)
        _, L, alpha = O.factorize(em.X, resid, em.kind, em.ls, em.s, em.noise)
        _ORACLE_CACHE.update(em=em, L=L, alpha=alpha, t_fac=time.perf_counter() - t0)
    em, L, alpha, t_fac = (_ORACLE_CACHE[k] for k in ("em", "L", "alpha", "t_fac"))
    Xq = torch.from_numpy(em.scx.transform(Xs[:sample_pts]))
    t0 = time.perf_counter()
    for lo in range(0, sample_pts, chunk):
        xs = Xq[lo:lo + chunk]
        Ks = O.kernel_matrix(xs, em.X, em.kind, em.ls, em.s)
        mean = em._mean(xs) + Ks @ alpha
        V = torch.linalg.solve_triangular(L, Ks.T, upper=False)
        var = (em.s - (V * V).sum(0) + em.noise).clamp_min(1e-10)
        _ = em.scy.sigma * mean + em.scy.mu, em.scy.sigma * var.sqrt()
    t_pred = time.perf_counter() - t0
    return sample_pts / t_pred, t_pred, t_fac, torch.get_num_threads()


def torch_gpu_predict_rate(X, y, Xs, hyper, dev, sample_pts, chunk=4096):
    """"Baseline B" of SURVEY.md section 8d: the same exact-GP math written with plain torch ops on the same B200
    (cuSOLVER potrf, cuBLAS trsm/gemm in FP64, unfused pointwise kernels) -- what a GPErks user gets from
    device="cuda" with stock libraries.  Uses the oracle's formulas on device tensors; reported, never shipped."""
    from oracle import gp_oracle as O
    em = O.OracleEmulator(X, y, kind=O.KIND_MATERN52, lengthscale=hyper["lengthscale"], outputscale=hyper["outputscale"],
                          noise=hyper["noise"], weights=hyper["weights"], bias=hyper["bias"], degree=1)
    Xd, yd = em.X.to(dev), em.y.to(dev)
    ls, s, nz, w, b = em.ls.to(dev), em.s.to(dev), em.noise.to(dev), em.w.to(dev), em.b.to(dev)
    K = torch.empty(Xd.shape[0], Xd.shape[0], dtype=torch.float64, device=dev)
    for lo in range(0, Xd.shape[0], 2048):
        K[lo:lo + 2048] = O.kernel_from_r2(O.sq_dist(Xd[lo:lo + 2048], Xd, ls, "gpytorch"), em.kind, s)
    K.diagonal().add_(nz)
    L = torch.linalg.cholesky(K)
    del K
    alpha = torch.cholesky_solve((yd - (Xd @ w + b)).reshape(-1, 1), L).reshape(-1)
    Xq = torch.from_numpy(em.scx.transform(Xs[:sample_pts])).to(dev)

    def run():
        for lo in range(0, sample_pts, chunk):
            xs = Xq[lo:lo + chunk]
            Ks = O.kernel_from_r2(O.sq_dist(xs, Xd, ls, "gpytorch"), em.kind, s)
            mean = xs @ w + b + Ks @ alpha
            V = torch.linalg.solve_triangular(L, Ks.T, upper=False)
            var = (s - (V * V).sum(0) + nz).clamp_min(1e-10)
            _ = em.scy.sigma * mean + em.scy.mu, em.scy.sigma * var.sqrt()
    run()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e)
    return {"value": sample_pts / ms * 1e3, "unit": "points/s", "what": "plain torch float64 on the same B200 "
            "(cuSOLVER/cuBLAS trsm + unfused pointwise kernels), device-resident", "sample": f"{sample_pts} points, chunks of {chunk}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, y, Xs, hyper = synthetic_problem(m=max(args.ref_points, 4096))
    rates, secs = [], []
    for i in range(args.warmup + args.steps):
        r, t_pred, t_fac, threads = cpu_oracle_predict_rate(X, y, Xs, hyper, args.ref_points)
        if i >= args.warmup:
            rates.append(r)
            secs.append(t_pred)
    value = args.ref_points * len(secs) / sum(secs)
    line = {
        "impl": "reference", "metric": "predict points/sec (mean+var) N=8192 D=8", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: Matern-5/2 ARD emulator, N_train=8192 D=8, predict mean+noisy std",
                   "points_per_step": args.ref_points, "note": "bounded sample of the same workload; cost is linear in M"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"{args.ref_points} of the 1e6 X* points per step, chunks of 4096, factorisation ({t_fac:.1f} s) excluded"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000, help="predict points per GPU per step")
    ap.add_argument("--ref-points", type=int, default=65536, help="points per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=98304, help="points of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mll-evals", type=int, default=3)
    ap.add_argument("--no-tf32", action="store_true", help="skip the TF32-mode (tcgen05) arm")
    ap.add_argument("--no-int8", action="store_true", help="skip the sliced-integer (tcgen05 kind::i8) arm")
    ap.add_argument("--precision", default="auto",
                    help="precision of the headline arms: auto (GPEmulator.predict's default: int8x6 for this workload), fp64, int8x5, int8x6")
    ap.add_argument("--ncu-traffic", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from an ncu --set full capture (profiles/)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from gperks_b200 import _native as nv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    M = args.points
    X, y, Xs, hyper = synthetic_problem(m=M, rank=rank)
    em = build_emulator(X, y, hyper, f"cuda:{local_rank}")
    lib = nv.lib()

    # ---- factorise once (replicated on every rank), time MLL+grad evaluations as the secondary metric
    eng = em.model._factorized_engine()
    torch.cuda.synchronize()
    theta, resid = eng.theta_in.clone(), eng.r[:eng.N].clone()
    def time_mll():
        ms = []
        for i in range(args.mll_evals + 1 if args.mll_evals > 0 else 0):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.factorize(theta, resid)
            eng.gradients()
            b.record()
            torch.cuda.synchronize()
            if i > 0:
                ms.append(a.elapsed_time(b))
        return ms
    mll_ms = time_mll()                      # the engine's default: K_hat^-1 on the integer tensor cores at this size
    lauum_default = eng.lauum_slices
    eng.lauum_slices = 0                     # all-FP64 edition (DMMA lauum) beside it
    mll_ms_fp64 = time_mll() if lauum_default else mll_ms
    eng.lauum_slices = lauum_default
    mll_ms_i8_trtri = None
    if getattr(eng, "_trtri_i8_capable", False):      # experimental, opt-in: the top merge levels of the inverse on the integer tensor cores too
        eng.trtri_i8 = True
        mll_ms_i8_trtri = time_mll()
        eng.trtri_i8 = False
    eng.factorize(theta, resid)

    # ---- device-resident arm (`value`)
    scx, scy = em.scaled_data.scx, em.scaled_data.scy
    xa = torch.as_tensor(np.asarray(scx.a, dtype=np.float64), device=dev)
    xr = torch.as_tensor(np.asarray(scx.b - scx.a, dtype=np.float64), device=dev)
    poly = em._poly_spec(dev)
    x_dev = torch.from_numpy(Xs).to(dev)
    out_mean = torch.empty(M, dtype=torch.float64, device=dev)
    out_std = torch.empty(M, dtype=torch.float64, device=dev)
    Mc = eng.pick_chunk(M)
    n_chunks = math.ceil(M / Mc)
    PREC = em.resolve_precision(args.precision)        # what GPEmulator.predict(X) runs by default for this emulator
    gathered = [torch.empty(2, M, dtype=torch.float64, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    local_out = torch.empty(2, M, dtype=torch.float64, device=dev)

    def step():
        eng.predict(x_dev, xa=xa, xr=xr, poly=poly, y_mu=float(scy.mu), y_sigma=float(scy.sigma),
                    out_mean=out_mean, out_std=out_std, chunk=Mc, precision=PREC)
        if world > 1:       # the path's only exchange: gather the 16 B/point results on rank 0
            local_out[0].copy_(out_mean)
            local_out[1].copy_(out_std)
            dist.gather(local_out, gathered, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    with ClockSampler(local_rank) as clocks:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            step()
        b.record()
        barrier()
        elapsed_ms = a.elapsed_time(b)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * M * args.steps / (elapsed_ms / 1e3)

    # ---- the FP64 DMMA arm (precision="fp64"): the reference point every other mode's error is measured against
    f_mean = torch.empty(M, dtype=torch.float64, device=dev)
    f_std = torch.empty(M, dtype=torch.float64, device=dev)

    def f64_step():
        eng.predict(x_dev, xa=xa, xr=xr, poly=poly, y_mu=float(scy.mu), y_sigma=float(scy.sigma),
                    out_mean=f_mean, out_std=f_std, chunk=Mc, precision="fp64")
    f64_step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        f64_step()
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    f64_ms = float(t.item())
    f64_value = world * M * args.steps / (f64_ms / 1e3)
    headline_err_std = float(((out_std - f_std).abs() / f_std).max())
    headline_err_mean = float((out_mean - f_mean).abs().max() / f_mean.abs().max())

    # ---- the FP64 variance TRMM alone on one full chunk, CUDA events on its stream
    ws = eng._workspace(int(lib.gpx_predict_workspace_bytes(eng.Np, Mc)))
    Ks_ptr = ws.data_ptr()
    qpart = torch.empty(eng.nb * Mc, dtype=torch.float64, device=dev)
    st = nv.stream_ptr()
    k_ms = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nv.check(lib.gpx_trmm_sumsq(nv.ptr(eng.S), eng.Np, Ks_ptr, Mc, nv.ptr(qpart), st), "trmm")
        b.record()
        torch.cuda.synchronize()
        if i > 0:
            k_ms.append(a.elapsed_time(b))
    k_ms = sum(k_ms) / len(k_ms)
    flops_per_launch = float(eng.N) ** 2 * Mc            # N^2 FLOP per point for the variance
    achieved = flops_per_launch / (k_ms / 1e3) / 1e12
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    peak_ms = []
    for i in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        peak_flops = lib.gpx_peak_dmma(148 * 8, 20000, nv.ptr(sink), st)
        b.record()
        torch.cuda.synchronize()
        peak_ms.append(a.elapsed_time(b))
    peak = peak_flops / (min(peak_ms) / 1e3) / 1e12
    kernel_share = k_ms * (M / Mc) / (f64_ms / args.steps)      # TRMM work is linear in the points

    # ---- end-to-end arm: GPEmulator.predict on host numpy buffers (pinned H2D + D2H inside)
    e2e_steps = max(1, min(args.steps, 2))
    em.predict(Xs[: min(M, 200_000)], precision=args.precision)       # warm the pinned allocator
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        y_mean, y_std = em.predict(Xs, precision=args.precision)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * M * e2e_steps / float(t.item())
    parity = float(np.max(np.abs(y_mean - out_mean.cpu().numpy())))     # the two arms must agree bit for bit

    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (one chunk of Mc points) from the
    # committed `ncu --set full` capture (profiles/r01_ncu_trmm_tma.txt); only valid for the default chunk size
    TRAFFIC = args.ncu_traffic if args.ncu_traffic is not None else (NCU_TRAFFIC_BYTES if Mc == 32768 and eng.N == 8192 else None)

    # ---- TF32 mode (tcgen05 3xTF32 variance contraction), reported beside the FP64 headline
    tf32 = None
    if not args.no_tf32:
        s64 = f_std
        tmean = torch.empty_like(out_mean)
        tstd = torch.empty_like(out_std)

        def tf32_arm(precision):
            def tf32_step():
                eng.predict(x_dev, xa=xa, xr=xr, poly=poly, y_mu=float(scy.mu), y_sigma=float(scy.sigma),
                            out_mean=tmean, out_std=tstd, chunk=Mc, precision=precision)
            for _ in range(3):
                tf32_step()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                tf32_step()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            rel_s = ((tstd - s64).abs() / s64)
            rel_m = (tmean - f_mean).abs().max() / f_mean.abs().max()
            return {"value": world * M * args.steps / (ms / 1e3), "unit": "points/s", "ms_per_step": ms / args.steps,
                    "max_rel_err_std_vs_fp64": float(rel_s.max()), "median_rel_err_std_vs_fp64": float(rel_s.median()),
                    "max_err_mean_vs_fp64_rel_to_scale": float(rel_m)}
        fast = tf32_arm("tf32x3_fast")
        exact = tf32_arm("tf32x3")
        tf32_ms = exact["ms_per_step"] * args.steps
        rel = ((tstd - s64).abs() / s64)
        # the tensor-core kernel alone, and the cuBLAS TF32 GEMM rate measured here as its denominator
        Khi = ws.view(torch.float32)[: Mc * eng.Np]
        Klo = ws.view(torch.float32)[Mc * eng.Np: 2 * Mc * eng.Np]
        nt = int(lib.gpx_tf32_row_tiles(eng.Np))
        qp = torch.empty(nt * Mc, dtype=torch.float64, device=dev)
        tk = []
        for i in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nv.check(lib.gpx_trmm_sumsq_tf32(nv.ptr(eng.S_hi), nv.ptr(eng.S_lo), eng.Np, Khi.data_ptr(), Klo.data_ptr(), Mc,
                                             nv.ptr(qp), 32, st), "tf32 trmm")
            b.record()
            torch.cuda.synchronize()
            if i > 0:
                tk.append(a.elapsed_time(b))
        tk = sum(tk) / len(tk)
        a32 = torch.randn(8192, 8192, device=dev)
        b32 = torch.randn(8192, 8192, device=dev)
        torch.backends.cuda.matmul.allow_tf32 = True
        tg = []
        for i in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            torch.matmul(a32, b32)
            b.record()
            torch.cuda.synchronize()
            if i > 0:
                tg.append(a.elapsed_time(b))
        torch.backends.cuda.matmul.allow_tf32 = False
        tf32_peak = 2 * 8192 ** 3 / (min(tg) / 1e3) / 1e12
        raw = 3.0 * float(eng.N) ** 2 * Mc / (tk / 1e3) / 1e12
        tf32 = {"tf32x3": exact, "tf32x3_fast": fast, "tolerance": 1e-3, "kchunk": 32,
                "note": "tf32x3: K* evaluated in FP64 (mean bit-identical to fp64); tf32x3_fast: K* evaluated in FP32",
                "roofline": {"kernel": "trmm_sumsq_tf32_kernel (tcgen05.mma kind::tf32, 3 MMAs per product, TMEM accumulators)",
                             "bound": "tensor", "achieved": raw, "peak": tf32_peak, "unit": "TFLOP/s (TF32, raw = 3 x useful)",
                             "frac": raw / tf32_peak, "useful_tflops": raw / 3.0, "kernel_ms": tk,
                             "peak_source": "cuBLAS TF32 GEMM 8192^3 (torch.matmul, allow_tf32) measured in this run"}}

    # ---- sliced-integer mode (tcgen05 kind::i8, exact int32 accumulation; std inside the FP64 bar of 1e-6), beside the headline
    int8 = None
    if not args.no_int8 or PREC.startswith("int8"):
        s64 = f_std
        imean = torch.empty_like(out_mean)
        istd = torch.empty_like(out_std)
        f_std_host = f_std.cpu().numpy()
        int8 = {"tolerance": 1e-6, "note": "FP64 contraction from S signed 7-bit slices per operand on the integer tensor cores; "
                                           "K* and the posterior mean stay on the FP64 kernels (mean bit-identical)"}
        try:
            ai = torch.randint(-64, 64, (8192, 8192), device=dev, dtype=torch.int8)
            bi = torch.randint(-64, 64, (8192, 8192), device=dev, dtype=torch.int8)
            tg = []
            for i in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                torch._int_mm(ai, bi)
                b.record()
                torch.cuda.synchronize()
                if i > 0:
                    tg.append(a.elapsed_time(b))
            i8_peak = 2 * 8192 ** 3 / (min(tg) / 1e3) / 1e12
            del ai, bi
        except Exception as exc:          # noqa: BLE001 -- the probe is optional; the mode itself never falls back
            i8_peak = None
            int8["peak_probe_error"] = repr(exc)[:200]
        # the tensor core's own limit: kind::i8 MMAs (M=128, N=256, K=32) issued back to back from resident shared memory
        # burst: best of three 5 ms launches after an idle gap; sustained: 1.5 s of back-to-back launches, the second half timed
        # (what a kernel inside a long step can get once the power cap has pulled the clock down)
        torch.cuda.synchronize()
        time.sleep(0.5)
        pk = []
        for i in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pk_ops = lib.gpx_peak_i8mma(148, 40000, st)
            b.record()
            torch.cuda.synchronize()
            pk.append(a.elapsed_time(b))
        i8_issue_peak_burst = pk_ops / (min(pk) / 1e3) / 1e12
        n_sus = 280
        for _ in range(n_sus // 2):
            lib.gpx_peak_i8mma(148, 40000, st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_sus // 2):
            lib.gpx_peak_i8mma(148, 40000, st)
        b.record()
        torch.cuda.synchronize()
        i8_issue_peak = pk_ops * (n_sus // 2) / (a.elapsed_time(b) / 1e3) / 1e12
        int8["int8_mma_issue_peak_tops"] = i8_issue_peak
        int8["int8_mma_issue_peak_burst_tops"] = i8_issue_peak_burst
        int8["cublaslt_int8_gemm_tops"] = i8_peak
        for S_ in (5, 6):
            prec = f"int8x{S_}"

            def i8_step():
                eng.predict(x_dev, xa=xa, xr=xr, poly=poly, y_mu=float(scy.mu), y_sigma=float(scy.sigma),
                            out_mean=imean, out_std=istd, chunk=Mc, precision=prec)
            for _ in range(3):
                i8_step()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                i8_step()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            rel_s = ((istd - s64).abs() / s64)
            entry = {"value": world * M * args.steps / (ms / 1e3), "unit": "points/s", "ms_per_step": ms / args.steps,
                     "max_rel_err_std_vs_fp64": float(rel_s.max()), "median_rel_err_std_vs_fp64": float(rel_s.median()),
                     "mean_bit_identical_to_fp64": bool(torch.equal(imean, f_mean))}
            # end to end through GPEmulator.predict(numpy)
            em.predict(Xs[: min(M, 200_000)], precision=prec)
            barrier()
            t0 = time.perf_counter()
            ym8, ys8 = em.predict(Xs, precision=prec)
            barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            entry["e2e"] = {"value": world * M / float(t.item()), "unit": "points/s", "api": f"GPEmulator.predict(numpy, precision='{prec}')",
                            "max_rel_err_std_vs_fp64_e2e": float(np.max(np.abs(ys8 - f_std_host) / f_std_host))}
            # the integer tensor-core kernel alone on the slices left in the workspace by the last chunk
            ws8 = eng._workspace(int(lib.gpx_predict_i8_workspace_bytes(eng.Np, Mc, S_)))
            Kp_ptr = ws8.data_ptr() + Mc * eng.Np * 8
            nt8 = int(lib.gpx_i8_row_tiles(eng.Np))
            qp8 = torch.empty(nt8 * Mc, dtype=torch.float64, device=dev)
            tk = []
            for i in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                nv.check(lib.gpx_trmm_sumsq_i8(nv.ptr(eng.S_i8), nv.ptr(eng.S_i8_inv_scale), eng.Np, Kp_ptr, Mc, S_,
                                               nv.ptr(eng.theta) + 8 * DIM, nv.ptr(qp8), st), "i8 trmm")
                b.record()
                torch.cuda.synchronize()
                if i > 0:
                    tk.append(a.elapsed_time(b))
            tk = sum(tk) / len(tk)
            raw = (S_ * (S_ + 1) / 2) * float(eng.N) ** 2 * Mc / (tk / 1e3) / 1e12
            entry["roofline"] = {"kernel": f"trmm_sumsq_i8_kernel<{S_}> (tcgen05.mma kind::i8, {S_ * (S_ + 1) // 2} slice products, int32 TMEM accumulators)",
                                 "bound": "tensor", "achieved": raw, "peak": i8_issue_peak, "unit": "TOP/s (int8, raw)",
                                 "frac": raw / i8_issue_peak, "peak_burst": i8_issue_peak_burst, "frac_of_burst_peak": raw / i8_issue_peak_burst,
                                 "frac_of_cublaslt_int8_gemm": (raw / i8_peak) if i8_peak else None,
                                 "fp64_equivalent_tflops": float(eng.N) ** 2 * Mc / (tk / 1e3) / 1e12,
                                 "kernel_ms": tk, "kernel_share_of_step": tk * (M / Mc) / (ms / args.steps),
                                 "peak_source": "gpx_peak_i8mma probe measured in this run (kind::i8 MMA issue rate from resident shared memory, "
                                                "one CTA per SM; MEASURED_PEAKS.json has no int8 figure): `peak` = SUSTAINED (1.5 s of back-to-back "
                                                "probe launches, second half timed -- this kernel is timed inside a long, power-capped run), "
                                                "`peak_burst` = best 5 ms launch after an idle gap; the cuBLASLt int8 GEMM 8192^3 (torch._int_mm) "
                                                "rate of the same run is recorded beside them"}
            int8[prec] = entry

    fp64_roofline = {"kernel": "trmm_sumsq_tma_kernel (variance contraction L^-1 K*^T: TMA + mbarrier ring, FP64 DMMA)", "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": TRAFFIC,
                     "peak_source": "gpx_peak_dmma probe measured in this run (FP64 DMMA.8x8x4 issue peak; MEASURED_PEAKS.json "
                                    "has no FP64 figure; cuBLAS DGEMM 8192^3 measured 35.5 TF/s on this pool, see profiles/)",
                     "algorithmic_flops_per_launch": flops_per_launch, "kernel_ms": k_ms, "kernel_share_of_step": kernel_share}
    if PREC.startswith("int8") and int8 is not None and PREC in int8:
        # headline = the sliced-integer path: its dominant kernel is the integer tensor-core contraction.  `achieved` counts the
        # int8 operations the kernel actually issues (S(S+1)/2 slice products of the algorithmic N^2/2 multiply-adds per
        # point); the FP64-equivalent rate (algorithmic N^2 FLOP per point / time) is given beside it.
        roofline = dict(int8[PREC]["roofline"])
        roofline["traffic"] = args.ncu_traffic if args.ncu_traffic is not None else (
            NCU_TRAFFIC_I8_BYTES.get(PREC) if Mc == 32768 and eng.N == 8192 else None)
        roofline["algorithmic_flops_per_launch"] = flops_per_launch
        roofline["nominal_int8_peak_tops"] = 4500.0
        roofline["kernel_share_of_step"] = roofline["kernel_ms"] * (M / Mc) / (elapsed_ms / args.steps)
        dtype = f"i8 slices ({PREC[-1]} x 7 bit, exact i32 accumulation on tcgen05, FP64 combination): f64-equivalent, see accuracy"
        launches_per_chunk = 4       # kmat_cross, slice_i8, trmm_sumsq_i8, predict_finalize
    else:
        roofline, dtype, launches_per_chunk = fp64_roofline, "f64", 3
    line = {
        "metric": "predict points/sec (mean+var) N=8192 D=8", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": "configs[1]: Matern-5/2 ARD emulator, N_train=8192 D=8, predict mean+noisy std",
                   "points_per_gpu_per_step": M, "chunk_points": Mc, "parallelism": f"x*-shard x{world}, L^-1/alpha replicated",
                   "precision": PREC, "precision_requested": args.precision,
                   "l2": "working set per chunk (K* 2 GiB + its slices + L^-1) exceeds the 126 MB L2; no flush needed"},
        "accuracy": {"max_rel_err_std_vs_fp64_dmma_path": headline_err_std, "max_err_mean_vs_fp64_dmma_path_rel_to_scale": headline_err_mean,
                     "tolerance": 1e-6, "points": int(M),
                     "note": "every point of the timed step compared with the precision='fp64' (DMMA) arm of the same run; "
                             "the FP64 arm itself is within 3e-12 / 5e-13 of cuSOLVER/cuBLAS FP64 math (tests)"},
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": int(M * DIM * 8), "d2h_bytes_per_step": int(M * 16),
                "steps": e2e_steps, "api": "GPEmulator.predict(numpy)  [default precision='auto']", "max_abs_diff_vs_device_arm": parity},
        "gpu_launches": int(args.steps * n_chunks * launches_per_chunk),
        "roofline": roofline,
        "fp64_dmma_mode": {"value": f64_value, "unit": "points/s", "ms_per_step": f64_ms / args.steps, "roofline": fp64_roofline,
                           "note": "precision='fp64': the same predict with the variance contraction on the FP64 tensor path (DMMA)"},
        "mll_grad": ({"evals_per_s": 1e3 / (sum(mll_ms) / len(mll_ms)), "ms_per_eval": sum(mll_ms) / len(mll_ms),
                      "n_train": eng.N, "what": "K_hat build + Cholesky + L^-1 + alpha + -mll + K_hat^-1 + gradient reduce",
                      "k_inverse": (f"int8x{lauum_default} slices on tcgen05 kind::i8 (gpx_lauum_i8)" if lauum_default else "FP64 DMMA lauum"),
                      "ms_per_eval_all_fp64": sum(mll_ms_fp64) / len(mll_ms_fp64),
                      "ms_per_eval_experimental_i8_trtri": (sum(mll_ms_i8_trtri) / len(mll_ms_i8_trtri)) if mll_ms_i8_trtri else None,
                      "experimental_i8_trtri_note": "opt-in (GPX_TRTRI=i8): L^-1 accurate to ~1e-8 of its norm -- enough for loss and "
                                                    "gradients, not for the predictive variance at small noise; not used by any other number here"}
                     if mll_ms else None),
        "tf32_mode": tf32,
        "int8_mode": int8,
        "clocks": clocks.summary(),
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        Xc, yc, Xsc, hyp = X, y, Xs, hyper
        r, t_pred, t_fac, threads = cpu_oracle_predict_rate(Xc, yc, Xsc, hyp, min(args.cpu_sample, M))
        line["cpu_baseline"] = {"value": r, "unit": "points/s", "cores": threads, "kind": "port",
                                "sample": f"first {min(args.cpu_sample, M)} X* points, chunks of 4096, {t_pred:.1f} s "
                                          f"(factorisation {t_fac:.1f} s excluded)"}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["torch_gpu_baseline"] = torch_gpu_predict_rate(X, y, Xs, hyper, dev, min(32768, M))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
