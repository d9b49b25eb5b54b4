"""Diagnostics (GPU): where the host time of one end-to-end registration goes."""
import cProfile, pstats, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair

n, K = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
X_fix, X_mov, _ = make_pair(n, 0)
Xf = torch.from_numpy(X_fix).pin_memory().numpy(); Xm = torch.from_numpy(X_mov).pin_memory().numpy()
out = torch.empty((n, 3), dtype=torch.float64).pin_memory().numpy()
import os
if os.environ.get('SICP_OWN_STREAM'):
    _st = torch.cuda.Stream()
    eng = _capi.Engine(0, stream=int(_st.cuda_stream))
    print('engine on its own non-default stream')
else:
    eng = _capi.Engine()
for _ in range(2):
    sb.register(Xf, Xm, correspondences=K, engine=eng, transform_out=out, want_normals=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    r = sb.register(Xf, Xm, correspondences=K, engine=eng, transform_out=out, want_normals=False)
print("ms per registration:", (time.perf_counter() - t0) / 5 * 1e3, "iterations", r.iterations, "loop_ms", r.loop_ms)
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    sb.register(Xf, Xm, correspondences=K, engine=eng, transform_out=out, want_normals=False)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
