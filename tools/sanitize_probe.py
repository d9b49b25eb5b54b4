"""Run under compute-sanitizer (GPU): one small registration through every engine path —
fused register, stage-by-stage, linearised variant, brute-force engine, K > 4096 (cooperative
first iteration, barrier-free kernel with several blocks afterwards), the three k-NN kernels,
2 / 4 / 16 lanes per query, movable-side attributes, the batched engine — so memcheck sees all
kernels.  python tools/sanitize_probe.py"""
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
    # round 2: k-NN kernel variants, lanes per query, cooperative-only loop, movable attributes
    for mode in (0, 1, 2):
        e.set_option("knn_coop", mode)
        h = sb.register(X_fix, X_mov, correspondences=3000, engine=e, want_normals=True)
        assert np.linalg.norm(h.H - H_true) < 5e-2
    e.set_option("knn_coop", -1)
    for mg in (2, 4, 16):
        e.set_option("match_group", mg)
        h = sb.register(X_fix, X_mov, correspondences=6000, engine=e, want_normals=False)
        assert np.linalg.norm(h.H - c.H) < 1e-9
    e.set_option("match_group", 0)
    e.set_option("fused", 0)
    h = sb.register(X_fix, X_mov, correspondences=6000, engine=e, want_normals=False)
    assert np.linalg.norm(h.H - c.H) < 1e-9
    e.set_option("fused", 1)
    pc = sb.PointCloud(X_mov[:20000], columns=["x", "y", "z"])
    pc.estimate_normals(10)
    mov = tuple(pc[col].to_numpy().astype(np.float32) for col in ("nx", "ny", "nz", "planarity"))
    m = sb.register(X_fix, X_mov[:20000], correspondences=6000, engine=e, mov_normals=mov, max_angle_between_normals=20.0)
    assert np.linalg.norm(m.H - H_true) < 5e-2
pairs = [sb.synthetic.c5_pair(i, 20_000) for i in range(4)]
tab = sb.simpleicp_batch([(p[0], p[1]) for p in pairs], engine="batched", correspondences=500)
assert np.isfinite(tab).all()
print("ok", a.iterations, m.iterations, int(tab[:, 16].sum()), c.iterations, d.iterations, f.iterations, g.iterations, l.iterations,
      float(np.linalg.norm(c.H - H_true)), float(np.linalg.norm(l.T - H_true)))
