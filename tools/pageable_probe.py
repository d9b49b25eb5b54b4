"""Diagnostics (GPU): why bench.py's e2e_simpleicp_pageable is above tools/upload_probe.py's number.
Times simpleicp() on NumPy arrays with the library-owned engine and with an explicit one, call by
call, before and after the pinned end-to-end registrations bench.py runs first.  Result: 5.0-5.7 ms
for the second to fourth call after an engine's creation, 4.7-5.0 ms from then on (the first call
of a new engine costs 120-200 ms of allocations) -- bench.py now warms up twice and reports the
median of five calls.  Not part of the product."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import simpleicp_b200 as sb
from bench import make_pair
from simpleicp_b200 import _capi

n, K = 1_000_000, 100_000
X_fix, X_mov, _ = make_pair(n, 0)


def run(tag, **kw):
    ts = []
    for _ in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sb.simpleicp(X_fix, X_mov, correspondences=K, **kw)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{tag}: " + " ".join(f"{t:.2f}" for t in ts), flush=True)


run("default engine        ")
eng = _capi.Engine()
run("explicit engine       ", engine=eng)
run("default engine again  ")
# what bench.py did before its measurement: pinned copies + an explicit engine registering them
Xf_p = _capi.pinned_empty(X_fix.shape, np.float64); Xf_p[:] = X_fix
Xm_p = _capi.pinned_empty(X_mov.shape, np.float64); Xm_p[:] = X_mov
out = _capi.pinned_empty(X_mov.shape, np.float64)
for _ in range(3):
    sb.register(Xf_p, Xm_p, correspondences=K, engine=eng, transform_out=out, want_normals=False)
run("after pinned e2e      ")
