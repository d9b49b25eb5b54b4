"""Diagnostics (GPU): steady-state kernel times of the C3 pair (optionally for the library named
by SICP_B200_LIB, see tools/build_variants.py).  One line per run."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
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
    e.iterate(p, x_in=np.zeros(6), want_record=True)
    for _ in range(6):
        rec = e.iterate(p, want_record=True)
    out = []
    for rep in range(5):
        sw = e.time_stages(p, 30, False)
        sc = e.time_stages(p, 30, True)
        out.append((sw["match_grid"] * 1e3, sc["match_grid"] * 1e3, sw["reject_solve"] * 1e3, sc["reject_solve"] * 1e3, sw["iteration"] * 1e3, sc["iteration"] * 1e3))
    o = np.median(np.array(out), axis=0)
    tag = os.path.basename(os.environ.get("SICP_B200_LIB", "default"))
    print(f"{tag:24s} match warm {o[0]:6.1f} cold {o[1]:6.1f} | rs warm {o[2]:6.1f} cold {o[3]:6.1f} | iter warm {o[4]:6.1f} cold {o[5]:6.1f} | kept {rec.n_kept}")
