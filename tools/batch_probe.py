"""C5 probe (BASELINE configs[4]): independent 100k-point scan pairs, K = 1000, through
simpleicp_batch on ONE GPU at several engine concurrencies.  Prints pairs/s and the stage times
of a single registration.  Run on the GPU box: python tools/batch_probe.py [n_pairs]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import simpleicp_b200 as sb  # noqa: E402
from simpleicp_b200 import _capi  # noqa: E402


def surface(n, seed, extent):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, extent, n)
    y = rng.uniform(0, extent, n)
    s = 100.0 / extent  # same undulation per point spacing as the C3 surface, scaled to the patch
    z = (0.05 * x * s + 0.03 * y * s + 2 * np.sin(2 * np.pi * x * s / 25) * np.cos(2 * np.pi * y * s / 40)) / s
    z = z + rng.normal(0, 0.01 / s, n)
    return np.column_stack([x, y, z])


def c5_pair(i, n=100_000):
    rng = np.random.default_rng(99 + 1000 * i)
    a = np.radians(rng.uniform(-1, 1, 3))
    t = rng.uniform(-0.2, 0.2, 3)
    from simpleicp_b200 import mathutils

    H = mathutils.create_homogeneous_transformation_matrix(mathutils.euler_angles_to_rotation_matrix(*a), t)
    Xf = surface(n, 10_000 + 2 * i, 30.0)
    Xm = surface(n, 10_001 + 2 * i, 30.0)
    Hi = np.linalg.inv(H)
    return Xf, Xm @ Hi[:3, :3].T + Hi[:3, 3], H


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    pairs = [c5_pair(i) for i in range(n_pairs)]
    P = [(a, b) for a, b, _ in pairs]
    eng = _capi.Engine()
    r = sb.register(*P[0], engine=eng, want_normals=False)
    t0 = time.perf_counter()
    for k in range(8):
        r = sb.register(*P[k % n_pairs], engine=eng, want_normals=False)
    dt = (time.perf_counter() - t0) / 8
    print(f"single engine: {1e3 * dt:.2f} ms/pair, iterations {r.iterations}, |H-H_true| {np.linalg.norm(r.H - pairs[(7) % n_pairs][2]):.2e}")
    print("  stage ms:", {k: round(v, 3) for k, v in r.timings.items() if k.endswith("_ms")}, "loop_ms", round(r.loop_ms, 3))
    eng.close()
    if len(sys.argv) > 2 and sys.argv[2] == "pinned":
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
        P = [(pin(a), pin(b)) for a, b in P]
        print("inputs in pinned host memory")
    for conc in (1, 2, 4, 8, 16):
        sb.simpleicp_batch(P[: min(conc, n_pairs)], concurrency=conc, want_normals=False)  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tab = sb.simpleicp_batch(P, concurrency=conc, want_normals=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        err = max(np.linalg.norm(tab[i, :16].reshape(4, 4) - pairs[i][2]) for i in range(n_pairs))
        its = tab[:, 16].astype(int)
        print(f"concurrency {conc:2d}: {n_pairs / dt:8.1f} pairs/s  ({1e3 * dt / n_pairs:.2f} ms/pair)  max |H-H_true| {err:.2e}"
              f"  iterations min/median/mean/max {its.min()}/{int(np.median(its))}/{its.mean():.1f}/{its.max()} ({int((its >= 100).sum())} hit the limit)")


if __name__ == "__main__":
    main()
