"""BASELINE configs[0] and [1] end to end (GPU): dragon (100k<->100k, K = 1000) and bunny
(max_overlap_distance = 1) through simpleicp(), wall time per registration with a reused engine,
and the CLI's own "Finished in" line on .xyz files."""
import re, subprocess, sys, tempfile, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch
import simpleicp_b200 as sb
from simpleicp_b200 import _capi
from conftest import load_pair, load_golden

for name, kw in (("dragon", {}), ("bunny", {"max_overlap_distance": 1.0})):
    X_fix, X_mov = load_pair(name)
    g = load_golden(name)
    eng = _capi.Engine()
    for _ in range(3):
        r = sb.register(X_fix, X_mov, engine=eng, want_normals=False, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        r = sb.register(X_fix, X_mov, engine=eng, want_normals=False, **kw)
    dt = (time.perf_counter() - t0) / 20
    print(f"{name}: {1e3*dt:.2f} ms per registration (engine reused, pageable inputs), {r.iterations} iterations, "
          f"loop {r.loop_ms:.2f} ms, |H - H_reference|_F = {np.linalg.norm(r.H - g['H']):.1e}; stages {({k: round(v, 3) for k, v in r.timings.items() if k.endswith('_ms')})}")
    print("   brute-force queries per iteration:", [rec["n_bruteforce"] for rec in r.records])
    eng.close()
    with tempfile.TemporaryDirectory() as d:
        f1, f2 = Path(d) / "a.xyz", Path(d) / "b.xyz"
        sb.write_xyz(f1, X_fix, decimals=6, header=False); sb.write_xyz(f2, X_mov, decimals=6, header=False)
        args = [str(REPO / "simpleicp_b200" / "sicp_cli"), "-f", str(f1), "-m", str(f2)] + (["-o", "1"] if kw else [])
        outs = []
        for _ in range(3):
            t0 = time.perf_counter(); out = subprocess.run(args, capture_output=True, text=True).stdout; wall = time.perf_counter() - t0
            outs.append((float(re.search(r"Finished in ([\d.]+) seconds", out).group(1)), wall))
        print(f"{name}: sicp_cli 'Finished in' {min(o[0] for o in outs):.3f} s (process wall incl. CUDA context and file parsing {min(o[1] for o in outs):.2f} s)")
