"""Diagnostics (GPU): lanes per query in the FIRST match of a C3 registration (far queries,
divergent paths) — host wall of iteration 0 incl. its general reject/solve kernel.  Not product."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair
import torch

X_fix, X_mov, _ = make_pair(1_000_000, 0)
K = 100_000
with _capi.Engine() as e:
    e.set_clouds(X_fix, X_mov)
    e.set_selected(sb.pointcloud.subsample_indices(len(X_fix), K).astype(np.int64))
    e.estimate_normals(10, download=False)
    lsq = e.lsq_params(np.zeros(6), np.zeros(6), np.zeros(6), 1.0)
    p = e.run_params(0.3, 1.0, 100, lsq)
    for mg in (0, 2, 4, 8, 16, 4):
        e.set_option("match_group", mg)
        ts = []
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e.iterate(p, x_in=np.zeros(6), want_record=True)
            ts.append((time.perf_counter() - t0) * 1e6)
        t1 = []
        for _ in range(3):
            e.iterate(p, x_in=np.zeros(6), want_record=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e.iterate(p, want_record=True)
            t1.append((time.perf_counter() - t0) * 1e6)
        print(f"match_group={mg}: iteration 0 {min(ts):.0f} us, iteration 1 {min(t1):.0f} us (host wall incl. sync)")
