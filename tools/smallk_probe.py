"""Diagnostics (GPU): per-iteration records of the small reference configurations (dragon, bunny):
Gauss-Newton steps, brute-force counts, and the device time of each kernel group."""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from conftest import load_pair

for name, kw in (("dragon", {}), ("bunny", {"max_overlap_distance": 1.0})):
    X_fix, X_mov = load_pair(name)
    with _capi.Engine() as e:
        r = sb.register(X_fix, X_mov, engine=e, want_normals=False, **kw)
        print(name, "iterations", r.iterations, "loop_ms", round(r.loop_ms, 3))
        print("  lm steps :", [rec["lm_iterations"] for rec in r.records])
        print("  kept     :", [rec["n_kept"] for rec in r.records])
        p = e.run_params(0.3, 1.0, 100, e.lsq_params(np.zeros(6), np.zeros(6), np.zeros(6), 1.0))
        e.iterate(p, x_in=np.array(r.records[-1]["x"]), want_record=True)
        for _ in range(3):
            e.iterate(p, want_record=True)
        st = e.time_stages(p, 30, False)
        t = e.phase_times()
        print("  steady state us:", {k: round(v * 1e3, 1) for k, v in st.items()},
              "| rs phases: median %.1f mad %.1f accum %.1f solve %.1f exit %.1f" % (t[1], t[2], t[3], t[5], t[9]))
