import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if "N" in d: print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k.startswith("t_") or k in ("N", "potrf", "trtri", "loss", "error")})
