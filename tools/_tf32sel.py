import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(l.strip()); continue
    print(d["N"], d["noise"], "std max", f'{d["std_rel_max"]:.2e}', "med", f'{d["std_rel_median"]:.2e}', "mean", f'{d["mean_rel_max"]:.1e}', "tf32 Mpts/s", round(d["tf32x3_pts_per_s"]/1e6, 3), "useful TF", round(d["tf32x3_useful_tflops"], 1))
