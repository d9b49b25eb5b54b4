"""Diagnostics (GPU): programmatic dependent launch between the two kernels of an iteration
(option "pdl") — steady-state iteration of C3 (events around whole iterations only) and whole
registrations of the small configurations, with and without.  Not part of the product."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair
from conftest import load_pair


def steady(tag, X_fix, X_mov, K, e):
    e.set_clouds(X_fix, X_mov)
    e.set_selected(sb.pointcloud.subsample_indices(len(X_fix), K).astype(np.int64) if K < len(X_fix) else None)
    e.estimate_normals(10, download=False)
    lsq = e.lsq_params(np.zeros(6), np.zeros(6), np.zeros(6), 1.0)
    p = e.run_params(0.3, 1.0, 100, lsq)
    for pdl in (0, 1, 0, 1):
        e.set_option("pdl", pdl)
        e.iterate(p, x_in=np.zeros(6), want_record=True)
        for _ in range(12):
            e.iterate(p, want_record=True)
        sc = e.time_stages(p, 40, True, True)
        sw = e.time_stages(p, 40, False, True)
        print(f"{tag} pdl={pdl}: iteration cold {sc['iteration']*1e3:.1f} us, warm {sw['iteration']*1e3:.1f} us")


def whole(tag, X_fix, X_mov, e, **kw):
    for pdl in (0, 1, 0, 1):
        e.set_option("pdl", pdl)
        sb.register(X_fix, X_mov, engine=e, want_normals=False, **kw)
        t0 = time.perf_counter()
        for _ in range(8):
            r = sb.register(X_fix, X_mov, engine=e, want_normals=False, **kw)
        dt = (time.perf_counter() - t0) / 8
        print(f"{tag} pdl={pdl}: {dt*1e3:.3f} ms per call, {r.iterations} iterations, loop {r.loop_ms:.3f} ms ({r.loop_ms/r.iterations*1e3:.1f} us/it)")


if __name__ == "__main__":
    X_fix, X_mov, _ = make_pair(1_000_000, 0)
    with _capi.Engine() as e:
        steady("C3", X_fix, X_mov, 100_000, e)
        whole("C3", X_fix, X_mov, e, correspondences=100_000)
        for name, kw in (("dragon", {}), ("bunny", {"max_overlap_distance": 1.0})):
            Xf, Xm = load_pair(name)
            steady(name, Xf, Xm, 1000, e)
            whole(name, Xf, Xm, e, **kw)
