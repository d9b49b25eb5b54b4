"""Diagnostics (GPU): host->device transfer of PAGEABLE clouds (NumPy arrays) through the worker
threads of csrc/upload.cu vs the driver's own staging (option "upload_threads" = 0), on the C3
pair: simpleicp() (one sicp_register call) and SimpleICP().run() (staged calls on pandas
containers).  Not part of the product."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import simpleicp_b200 as sb
from bench import make_pair

n, K = 1_000_000, 100_000
X_fix, X_mov, _ = make_pair(n, 0)
from simpleicp_b200 import _capi
import importlib

drv = importlib.import_module("simpleicp_b200.simpleicp")  # the module (the package attribute of that name is the function)

eng = _capi.Engine()
drv.default_engine = lambda device=None: eng  # calls without engine= use this one, options kept
ref = None
for T, ck in ((0, 2048), (2, 2048), (3, 2048), (4, 2048), (2, 1024), (3, 1024), (4, 1024), (2, 4096), (3, 512), (2, 2048), (0, 2048),
              (3, 2048), (2, 1024)):
    eng.set_option("upload_threads", T)
    eng.set_option("upload_chunk_kb", ck)
    ts, up = [], []
    for _ in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        H, X_t, rbp, res = sb.simpleicp(X_fix, X_mov, correspondences=K, engine=eng)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        up.append(eng.timings()["upload_ms"])
    if ref is None:
        ref = (H.copy(), np.asarray(X_t).copy())
    same = np.array_equal(H, ref[0]) and np.array_equal(np.asarray(X_t), ref[1])
    print(f"upload_threads={T} chunk={ck} KB: simpleicp() pageable min {min(ts[1:]):.2f} ms median {np.median(ts[1:]):.2f} ms "
          f"(first-cloud upload median {np.median(up[1:]):.2f} ms) identical={same}")
eng.set_option("upload_chunk_kb", 2048)
for T in (0, 2):
    eng.set_option("upload_threads", T)
    ts = []
    for _ in range(4):
        pc_fix = sb.PointCloud(X_fix, columns=["x", "y", "z"])
        pc_mov = sb.PointCloud(X_mov, columns=["x", "y", "z"], copy=True)
        icp = sb.SimpleICP(verbose=False)
        icp.add_point_clouds(pc_fix, pc_mov)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Hc, Xc, rbpc, rc = icp.run(correspondences=K)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"upload_threads={T}: SimpleICP().run() min {min(ts[1:]):.2f} ms  H==functional {np.array_equal(Hc, ref[0])} "
          f"X==pc_mov.X {np.array_equal(Xc, pc_mov.X)}")
eng.set_option("upload_threads", 2)
