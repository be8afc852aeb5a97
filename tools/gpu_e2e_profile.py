"""Where the end-to-end predict (host numpy in/out) loses time against the device-resident arm: wall time of
GPEmulator.predict for 1, 2, 4 staging blocks against eng.predict on resident inputs, plus a cProfile of one call."""
import cProfile
import pstats
import sys
import time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench as B
M = 1_000_000
X, y, Xs, hyper = B.synthetic_problem(m=M)
em = B.build_emulator(X, y, hyper, "cuda:0")
em.predict(Xs[:300000])
for m in (262144, 524288, 1_000_000):
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        em.predict(Xs[:m])
        t1 = time.perf_counter()
    x_dev = torch.from_numpy(Xs[:m]).cuda()
    em.predict_device(x_dev)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    em.predict_device(x_dev)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"M={m}: e2e {1e3 * (t1 - t0):.1f} ms, device {1e3 * (t3 - t2):.1f} ms, gap {1e3 * ((t1 - t0) - (t3 - t2)):.1f} ms", flush=True)
pr = cProfile.Profile()
pr.enable()
em.predict(Xs)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
