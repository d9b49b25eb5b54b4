"""Diagnostics (GPU): why bench.py's e2e_simpleicp_pageable is above tools/upload_probe.py's number.
Times simpleicp() on NumPy arrays with the library-owned engine and with an explicit one, before
and after the process has run the CPU oracle (thread pools of SciPy / BLAS).  Not part of the product."""
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
from oracle import simpleicp_oracle as O

t0 = time.perf_counter()
O.simpleicp(X_fix[:200000], X_mov[:200000], correspondences=2000, max_iterations=2)
print(f"oracle call {time.perf_counter() - t0:.2f} s")
run("after an oracle call  ")
run("explicit, after oracle", engine=eng)
