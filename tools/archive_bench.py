"""Archive gpurun_out/{bench_final.json, bench_ref.json, launches_final.csv} into profiles/ (round tag from argv[1], default r01):
pretty-printed bench lines and the per-kernel aggregate of the ncu launch list with the TRMM share of the FP64 predict step."""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
sfx = sys.argv[2] if len(sys.argv) > 2 else "final"      # "int8": the run with the sliced-integer default as headline
out, prof = ROOT / "gpurun_out", ROOT / "profiles"
d = json.loads((out / "bench_final.json").read_text().strip().splitlines()[-1])
(prof / f"{tag}_bench_1gpu_{sfx}.json").write_text(json.dumps(d, indent=1))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "frac", round(d["roofline"]["frac"], 4),
      "mll", d["mll_grad"], "cpu", d.get("cpu_baseline", {}).get("value"), "torch_gpu", d.get("torch_gpu_baseline", {}).get("value"))
if d.get("tf32_mode"):
    print({k: v["value"] for k, v in d["tf32_mode"].items() if isinstance(v, dict) and "value" in v})
ref = out / "bench_ref.json"
if ref.exists():
    r = json.loads(ref.read_text().strip().splitlines()[-1])
    (prof / (f"{tag}_bench_reference_arm.json" if sfx == "final" else f"{tag}_bench_reference_arm_{sfx}.json")).write_text(json.dumps(r, indent=1))
    print("reference arm", r["value"], r.get("cpu_baseline", {}).get("sample"))
rows = list(csv.reader(l for l in open(out / "launches_final.csv") if not l.startswith("==")))
hdr, agg = None, collections.OrderedDict()
unit = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
for r_ in rows:
    if r_ and r_[0] == "ID":
        hdr = r_
        continue
    if hdr and len(r_) == len(hdr):
        dd = dict(zip(hdr, r_))
        if dd.get("Metric Name") == "gpu__time_duration.sum":
            a = agg.setdefault(dd["Kernel Name"], [0, 0.0])
            a[0] += 1
            a[1] += float(dd["Metric Value"].replace(",", "")) * unit[dd["Metric Unit"]]
tot = sum(t for _, t in agg.values())
lines = ["# ncu launch list (gpu__time_duration.sum, --clock-control none), SHIPPED kernels (final)",
         "# command: python bench.py --steps 1 --warmup 3 --points 65536 --no-cpu-baseline --mll-evals 1",
         "# covers: factorisations (+gradients), the headline predict arm (3 warm-up + 1 timed step), the FP64 DMMA arm + roofline probes, e2e predict, TF32- and int8-mode arms",
         "# per-launch times are cold-cache and serialised under ncu: compare SHARES",
         f"# total {tot / 1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches", "kernel,launches,total_ms,share,mean_us"]
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"{k.split('(')[0][:110]},{n},{t / 1e3:.3f},{t / tot:.4f},{t / n:.1f}")


def mean(name):
    v = [t for k, t in agg.items() if name in k]
    return v[0][1] / v[0][0] if v else float("nan")


trmm, kc, fin = mean("trmm_sumsq_tma_kernel"), mean("kmat_cross_kernel<3>"), mean("predict_finalize_kernel")
lines.append(f"# FP64 predict step (one 32768-point chunk) under ncu: trmm {trmm / 1e3:.2f} ms, kmat_cross {kc / 1e3:.2f} ms, "
             f"finalize {fin / 1e3:.3f} ms -> TRMM share {trmm / (trmm + kc + fin):.3f} (bench.py live: {d.get('fp64_dmma_mode', d)['roofline']['kernel_share_of_step']:.3f})")
ti, ki = mean("trmm_sumsq_i8_kernel<6, 64, 1>"), mean("kmat_cross_i8_kernel<3, 6>")
if ti == ti:
    lines.append(f"# int8x6 predict step (one 32768-point chunk) under ncu: trmm_i8 {ti / 1e3:.2f} ms, kmat_cross_i8 {ki / 1e3:.2f} ms, "
                 f"finalize {fin / 1e3:.3f} ms -> integer-kernel share {ti / (ti + ki + fin):.3f}"
                 + (f" (bench.py live: {d['roofline']['kernel_share_of_step']:.3f})" if "i8" in d["roofline"]["kernel"] else ""))
(prof / (f"{tag}_launches_shipped.csv" if sfx == "final" else f"{tag}_launches_{sfx}.csv")).write_text("\n".join(lines) + "\n")
print("\n".join(lines[5:26]))
print(lines[-1])
