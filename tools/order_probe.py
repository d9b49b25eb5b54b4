"""Diagnostics (GPU): how much of the k-NN and 1-NN search time is the ORDER of the queries?
The C3 synthetic cloud is in random order (consecutive queries are far apart: every lane of a warp
walks a different part of the grid); lidar files come in scan order.  Same pair, fixed cloud
pre-sorted along a Morton curve on the host, so that the reference's index-equidistant selection
yields spatially coherent warps.  Not part of the product."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair


def morton_order(X, bits=10):
    lo, hi = X.min(axis=0), X.max(axis=0)
    q = np.minimum(((X - lo) / (hi - lo).max() * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    code = np.zeros(len(X), dtype=np.uint64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return np.argsort(code, kind="stable")


def run(tag, X_fix, X_mov, K):
    n = len(X_fix)
    with _capi.Engine() as e:
        e.set_clouds(X_fix, X_mov)
        e.set_selected(sb.pointcloud.subsample_indices(n, K).astype(np.int64))
        for mode in (0, 1, 2):
            e.set_option("knn_coop", mode)
            e.estimate_normals(10, download=False)
            e.estimate_normals(10, download=False)
            print(f"{tag}: normals knn_coop={mode}: {e.timings()['normals_ms'] * 1e3:.1f} us")
        e.set_option("knn_coop", -1)
        lsq = e.lsq_params(np.zeros(6), np.zeros(6), np.zeros(6), 1.0)
        p = e.run_params(0.3, 1.0, 100, lsq)
        e.iterate(p, x_in=np.zeros(6), want_record=True)
        for _ in range(12):
            e.iterate(p, want_record=True)
        sc, sw = e.time_stages(p, 20, True), e.time_stages(p, 20, False)
        print(f"{tag}: match cold {sc['match_grid']*1e3:.1f} warm {sw['match_grid']*1e3:.1f} | rs cold {sc['reject_solve']*1e3:.1f} "
              f"warm {sw['reject_solve']*1e3:.1f} | iteration cold {sc['iteration']*1e3:.1f} warm {sw['iteration']*1e3:.1f} us")


if __name__ == "__main__":
    n, K = 1_000_000, 100_000
    X_fix, X_mov, _ = make_pair(n, 0)
    run("C3 random order", X_fix, X_mov, K)
    o = morton_order(X_fix)
    run("C3 fixed cloud in Morton order", np.ascontiguousarray(X_fix[o]), X_mov, K)
    om = morton_order(X_mov)
    run("C3 both clouds in Morton order", np.ascontiguousarray(X_fix[o]), np.ascontiguousarray(X_mov[om]), K)
