"""Diagnostics (GPU): throughput of the TMA-staged brute-force engine on the full C3 problem."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_pair

n, K = 1_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
X_fix, X_mov, H_true = make_pair(n, 0)
with _capi.Engine() as e:
    e.set_clouds(X_fix, X_mov)
    idx = sb.pointcloud.subsample_indices(n, K).astype(np.int64)
    e.set_selected(idx)
    e.estimate_normals(10)
    idx_g, d_g = e.match(H_true)
    e.set_option("nn_engine", _capi.NN_BRUTE)
    idx_b, d_b = e.match(H_true)  # warm-up + correctness
    print("brute == grid:", np.array_equal(idx_b, idx_g), "max |dd|", np.abs(d_b - d_g).max())
    ts = []
    for _ in range(3):
        e.match(H_true)
        ts.append(e.timings()["match_ms"])
    ms = min(ts)
    pairs = float(n) * K
    print(f"brute force K={K}: {ms:.2f} ms, {pairs / ms / 1e6:.1f} G pairs/s, {8 * pairs / ms / 1e9:.1f} TFLOP/s (8 flop/pair), "
          f"{7 * pairs / ms / 1e9 / 32:.2f} T warp-instr/s")
