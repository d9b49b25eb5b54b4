"""Diagnostics (GPU): BASELINE configs[4] on one GPU — pairs/s of the batched engine against the
per-pair engine pool.  Not part of the product."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from bench import make_c5_pairs


def main(n_pairs=64, n_pts=100_000):
    pairs = make_c5_pairs(range(n_pairs), n_pts, pinned=True)
    for label, kw in (("batched", dict(engine="batched", batch_size=n_pairs)),
                      ("batched/32", dict(engine="batched", batch_size=32)),
                      ("pool x8", dict(engine="pool", concurrency=8))):
        sb.simpleicp_batch(pairs, on_error="nan", **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tab = sb.simpleicp_batch(pairs, on_error="nan", **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{label:12s}: {n_pairs / dt:8.0f} pairs/s ({dt * 1e3:.1f} ms for {n_pairs} pairs), mean iterations {tab[:, 16].mean():.1f}, "
              f"failed {(tab[:, 16] < 0).sum()}")
    eng = sb.batch._BATCH_ENGINES[0]
    print({k: round(v, 3) for k, v in eng.timings().items()})


if __name__ == "__main__":
    main()
