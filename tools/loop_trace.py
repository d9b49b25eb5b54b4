"""Diagnostics (GPU): host-side timeline of one C3 registration (SICP_TRACE_RUN=1: a stderr line
at every launch / synchronisation of sicp_register and its loop).  Not part of the product."""
import os
import sys
from pathlib import Path

os.environ["SICP_TRACE_RUN"] = "1"
import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair

X_fix, X_mov, _ = make_pair(1_000_000, 0)
import torch
Xf = torch.from_numpy(X_fix).pin_memory().numpy()
Xm = torch.from_numpy(X_mov).pin_memory().numpy()
with _capi.Engine() as e:
    for i in range(4):
        sys.stderr.write(f"==== registration {i}\n")
        r = sb.register(Xf, Xm, correspondences=100_000, engine=e, want_normals=False)
    print(r.iterations, r.loop_ms, e.timings())
