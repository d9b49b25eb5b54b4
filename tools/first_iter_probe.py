"""Diagnostics (GPU): cost of the FIRST iterations of a C3 registration for each lanes-per-query
setting of the cooperative match kernel."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair

n, K = 1_000_000, 100_000
X_fix, X_mov, _ = make_pair(n, 0)
idx = sb.pointcloud.subsample_indices(n, K).astype(np.int64)
with _capi.Engine() as e:
    e.set_clouds(X_fix, X_mov)
    e.set_selected(idx)
    e.estimate_normals(10)
    p = e.run_params(0.3, 1.0, 100, e.lsq_params(np.zeros(6), np.zeros(6), np.zeros(6), 1.0))
    for mg in (0, 4, 8, 16):
        e.set_option("match_group", mg)
        e.iterate(p, x_in=np.zeros(6), want_record=True)
        rows = []
        for rep in range(3):
            ts = []
            for it in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                e.iterate(p, x_in=np.zeros(6) if it == 0 else None, want_record=True)
                ts.append((time.perf_counter() - t0) * 1e6)
            rows.append(ts)
        m = np.median(np.array(rows), axis=0)
        print(f"match_group={mg:2d}: iteration 1 {m[0]:6.0f} us, 2 {m[1]:6.0f} us, 3 {m[2]:6.0f} us")
