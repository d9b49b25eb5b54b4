"""Experiment (GPU): how much would spatially ordered QUERIES buy?  The fixed cloud of the C3 pair
is in random spatial order; here it is additionally given (a) sorted by a coarse Morton key and
the same K-subsample is taken, so that consecutive queries are spatial neighbours.  Prints match
and normals kernel times for both orders.  Not part of the product."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair


def morton_order(X, bits=10):
    lo, hi = X.min(0), X.max(0)
    q = ((X - lo) / (hi - lo + 1e-12) * ((1 << bits) - 1)).astype(np.uint64)
    key = np.zeros(len(X), dtype=np.uint64)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return np.argsort(key, kind="stable")


def run(tag, X_fix, X_mov, idx, K):
    with _capi.Engine() as e:
        e.set_clouds(X_fix, X_mov)
        e.set_selected(idx)
        e.estimate_normals(10)
        p = e.run_params(0.3, 1.0, 100, e.lsq_params(np.zeros(6), np.zeros(6), np.zeros(6), 1.0))
        recs = [e.iterate(p, x_in=np.zeros(6), want_record=True)]
        for _ in range(5):
            recs.append(e.iterate(p, want_record=True))
        sw = e.time_stages(p, 20, False)
        sc = e.time_stages(p, 20, True)
        tm = e.timings()
        print(f"{tag:28s} match warm {sw['match_grid']*1e3:6.1f} cold {sc['match_grid']*1e3:6.1f} us | rs warm {sw['reject_solve']*1e3:6.1f} cold {sc['reject_solve']*1e3:6.1f} us | "
              f"iter warm {sw['iteration']*1e3:6.1f} cold {sc['iteration']*1e3:6.1f} | normals {tm['normals_ms']:.3f} ms kept {recs[-1].n_kept}")


def main():
    n, K = 1_000_000, 100_000
    X_fix, X_mov, _ = make_pair(n, 0)
    idx = sb.pointcloud.subsample_indices(n, K).astype(np.int64)
    run("random order (as bench)", X_fix, X_mov, idx, K)
    # fixed cloud stored in Morton order: the subsample (ascending indices) then visits space coherently
    Xf2 = X_fix[morton_order(X_fix)]
    run("fixed cloud Morton-sorted", Xf2, X_mov, idx, K)
    Xm2 = X_mov[morton_order(X_mov)]
    run("both clouds Morton-sorted", Xf2, Xm2, idx, K)


if __name__ == "__main__":
    main()
