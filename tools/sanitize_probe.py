"""Run under compute-sanitizer (GPU): one small registration through every engine path —
fused register, stage-by-stage, linearised variant, brute-force engine, K > 4096 cooperative
path — so memcheck sees all kernels.  python tools/sanitize_probe.py"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair

X_fix, X_mov, H_true = make_pair(60_000, 0)
with _capi.Engine() as e:
    a = sb.register(X_fix, X_mov, correspondences=2000, engine=e, want_normals=False)       # fused, single block
    b = sb.register(X_fix, X_mov, correspondences=2000, engine=e, want_normals=True)        # staged
    assert np.array_equal(a.H, b.H)
    c = sb.register(X_fix, X_mov, correspondences=6000, engine=e, want_normals=False)       # cooperative path
    d = sb.register(X_fix, X_mov, correspondences=2000, engine=e, max_overlap_distance=2.0, want_normals=False)
    e.set_option("nn_engine", _capi.NN_BRUTE)
    f = sb.register(X_fix, X_mov, correspondences=500, engine=e, want_normals=True)
    e.set_option("nn_engine", _capi.NN_AUTO)
    e.set_option("grid_max_rings", 1)                                                       # forces the TMA fallback
    g = sb.register(X_fix, X_mov, correspondences=2000, engine=e, want_normals=True)
    e.set_option("grid_max_rings", 8)
    assert np.linalg.norm(g.H - b.H) < 1e-9
    l = sb.simpleicp_linearized(X_fix, X_mov, correspondences=6000, engine=e)
    print("ok", a.iterations, c.iterations, d.iterations, f.iterations, g.iterations, l.iterations,
          float(np.linalg.norm(c.H - H_true)), float(np.linalg.norm(l.T - H_true)))
