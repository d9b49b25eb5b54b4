"""Workloads for the ncu captures of round 2 (GPU; not part of the product).

    python tools/profile_driver.py c3       # one C3 registration cut at 4 iterations (grid build, k-NN, PCA,
                                            # first-iteration match, general + barrier-free reject/solve, final
                                            # residuals, transform)
    python tools/profile_driver.py bunny    # partial overlap: overlap filter, bounded search
    python tools/profile_driver.py brute    # the TMA brute-force engine (4096 queries x 1M points)
    python tools/profile_driver.py batch    # the batched engine on 16 small pairs
    python tools/profile_driver.py steady   # C3 steady state: 14 iterations (capture the last ones)
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_c5_pairs, make_pair

mode = sys.argv[1] if len(sys.argv) > 1 else "c3"
n, K = 1_000_000, 100_000
if mode in ("c3", "steady", "brute"):
    X_fix, X_mov, _ = make_pair(n, 0)
if mode == "c3":
    with _capi.Engine() as e:
        sb.register(X_fix, X_mov, correspondences=K, engine=e, want_normals=False, max_iterations=4)
elif mode == "bunny":
    from conftest import load_pair

    Xf, Xm = load_pair("bunny")
    with _capi.Engine() as e:
        sb.register(Xf, Xm, engine=e, want_normals=False, max_overlap_distance=1.0, max_iterations=3)
elif mode == "brute":
    with _capi.Engine() as e:
        e.set_option("nn_engine", _capi.NN_BRUTE)
        sb.register(X_fix, X_mov, correspondences=4096, engine=e, want_normals=False, max_iterations=2)
elif mode == "batch":
    with _capi.Engine() as e:
        lsq = e.lsq_params(np.zeros(6), np.zeros(6), np.zeros(6), 1.0)
        e.register_batch(make_c5_pairs(range(16), 100_000), 1000, 10, e.run_params(0.3, 1.0, 6, lsq))
else:
    with _capi.Engine() as e:
        e.set_clouds(X_fix, X_mov)
        e.set_selected(sb.pointcloud.subsample_indices(n, K).astype(np.int64))
        e.estimate_normals(10, download=False)
        lsq = e.lsq_params(np.zeros(6), np.zeros(6), np.zeros(6), 1.0)
        p = e.run_params(0.3, 1.0, 100, lsq)
        e.iterate(p, x_in=np.zeros(6), want_record=True)
        for _ in range(13):
            e.iterate(p, want_record=True)
print("done", mode)
