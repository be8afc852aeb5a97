"""Print the per-stage errors and timings of gpurun_out/selfcheck.json (written by tools/gpu_selfcheck.py), one line per case."""
import json
import sys
from pathlib import Path

path = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent.parent / "gpurun_out" / "selfcheck.json"
d = json.loads(path.read_text())
for c in d["cases"]:
    if "error" in c:
        print(c)
        continue
    errs = {k: f"{v:.1e}" for k, v in c.items() if k in ("kmat_sym", "potrf", "trtri", "alpha", "loss", "grad", "pred_mean", "pred_std_rel")}
    times = {k[2:-3]: round(v, 3) for k, v in c.items() if k.startswith("t_") and k.endswith("_ms")}
    print(f"N={c['N']:<6} D={c['D']:<3} errors {errs}  ms {times}")
if "peaks" in d:
    print({k: round(v, 1) for k, v in d["peaks"].items()})
