#!/usr/bin/env python
"""bench.py — correspondences/sec per ICP iteration on the BASELINE.json C3 workload.

    python bench.py --gpus 1 --steps 20 --warmup 3                 # B200 arm
    python bench.py --impl reference --steps 20 --warmup 3          # CPU arm (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W       # weak scaling

Workload (BASELINE.json configs[2]): synthetic 1M <-> 1M point pair (tilted,
undulating plane with 1 cm noise), correspondences = 100 000, all other reference defaults.
One STEP = one ICP iteration of the hot path: nearest-neighbour match of the 100 000 selected
fixed points into the 1M-point movable cloud + point-to-plane distances, planarity / median-MAD
rejection, 6-DoF least-squares solve, residual statistics and stop-rule evaluation.

value      = K * steps / (sum of the per-step device times); every step starts with a COLD L2
             (a 256 MiB buffer is overwritten between steps, outside the timed intervals),
             inputs resident in HBM.  Per-step times are CUDA-event intervals on the launching
             stream; with N ranks the slowest rank's total is used.
e2e        = the same metric through the public API call  simpleicp(X_fix, X_mov,
             correspondences=K)  on pinned HOST arrays: upload, grid builds, normals, the
             whole iteration loop, final transform and download are all inside the timed region.
N > 1      = independent pairs, one per GPU (weak scaling); the only collective is the NCCL
             all-gather of the per-pair result records, inside the e2e region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "correspondences/sec per ICP iter (1M<->1M pts)"
UNIT = "corr/s"


def make_pair(n: int, rank: int):
    """C3 generator; rank r uses seeds shifted by 10 r (independent pairs)."""
    from oracle.simpleicp_oracle import rbp_to_H, surface, transform_by_H  # input generator only

    H_true = rbp_to_H([np.deg2rad(0.3), np.deg2rad(-0.2), np.deg2rad(0.5), 0.15, -0.10, 0.05])
    X_fix = surface(n, 1234 + 10 * rank)
    X_mov = transform_by_H(surface(n, 5678 + 10 * rank), np.linalg.inv(H_true))
    return np.ascontiguousarray(X_fix), np.ascontiguousarray(X_mov), H_true


def algorithmic_bytes(n_mov: int, K: int, n_kept: int):
    """B_iter = 16 N_mov + 60 K + 48 K_kept, split by kernel."""
    match = 16 * n_mov + 44 * K          # cloud stream + query + normal + (idx, dist) out
    reject_solve = 16 * K + 48 * n_kept  # select/compaction + fp64 gathers for the solve
    return match, reject_solve


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples while the benchmark runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.marks = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            p = [v.strip() for v in line.split(",")]
            if len(p) >= 7:
                try:
                    self.samples.append((time.time(), float(p[0]), float(p[1]), p[3:7]))
                except ValueError:
                    pass

    def summary(self, t0: float, t1: float):
        if self.proc is not None:
            self.proc.terminate()
        sel = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({self.NAMES[i] for s in sel for i, v in enumerate(s[3]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median([s[1] for s in sel])), "sm_max_mhz": sel[0][2],
                "reasons": reasons, "samples": len(sel)}


def run_reference(args):
    """CPU arm: the oracle port of the reference's algorithm (NumPy + SciPy cKDTree + TRF), the
    same SciPy calls the reference makes, on this host's cores.  One step = one ICP iteration."""
    from oracle import simpleicp_oracle as O

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X_fix, X_mov, _ = make_pair(args.points, 0)
    tr = O.Trace(light=True)
    n_it = args.warmup + args.steps
    t0 = time.perf_counter()
    O.simpleicp(X_fix, X_mov, correspondences=args.correspondences, min_change=0.0,
                max_iterations=n_it, trace=tr)
    total = time.perf_counter() - t0
    it_s = tr.iter_seconds[args.warmup:]
    loop = float(np.sum(it_s))
    value = args.correspondences * len(it_s) / loop
    e2e = args.correspondences * n_it / total
    cores = os.cpu_count()
    sample = (f"{n_it} ICP iterations (first {args.warmup} untimed) of the oracle port on the full C3 pair, "
              f"K={args.correspondences}; cKDTree queries use all {cores} cores (workers=-1), tree build, "
              "transforms, eig loop and TRF solve are single-threaded as in the reference")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(it_s), "warmup": args.warmup, "ms_per_step": 1e3 * loop / len(it_s),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"C3 synthetic surface pair {args.points}<->{args.points}, correspondences={args.correspondences}, neighbors=10",
                   "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup_s": {"normals": tr.timings.get("normals"), "total": total},
    }))


def run_b200(args):
    import torch
    import torch.distributed as dist

    import simpleicp_b200 as sb
    from simpleicp_b200 import _capi, batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, n = args.correspondences, args.points

    X_fix, X_mov, H_true = make_pair(n, rank)
    Xf_pin = torch.from_numpy(X_fix).pin_memory()
    Xm_pin = torch.from_numpy(X_mov).pin_memory()
    out_pin = torch.empty((n, 3), dtype=torch.float64).pin_memory()

    sampler = ClockSampler(local) if rank == 0 else None
    eng = _capi.Engine(local)

    # ---- setup (untimed for `value`): resident inputs, grid, selection, normals
    eng.set_clouds(Xf_pin.numpy(), Xm_pin.numpy())
    idx = sb.pointcloud.subsample_indices(n, K).astype(np.int64)
    eng.set_selected(idx)
    nrm = eng.estimate_normals(10)
    lsq = eng.lsq_params(np.zeros(6), np.zeros(6), np.zeros(6), 1.0)
    params = eng.run_params(0.3, 1.0, 100, lsq)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up then the timed region (cold L2 before every step)
    rec = eng.iterate(params, x_in=np.zeros(6), want_record=True)
    for _ in range(max(args.warmup - 1, 0)):
        rec = eng.iterate(params, want_record=True)
    launches0 = eng.timings()["kernel_launches"]
    barrier()
    t_region0 = time.time()
    st = eng.time_stages(params, args.steps, True)
    barrier()
    launches = eng.timings()["kernel_launches"] - launches0
    total_ms = st["iteration"] * args.steps
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * K * args.steps / (total_ms * 1e-3)
    rec = eng.iterate(params, want_record=True)
    n_kept = int(rec.n_kept)

    # ---- same loop, warm L2, iterations queued back to back with no host sync (what sicp_run does)
    eng.iterate(params, x_in=np.zeros(6))
    for _ in range(args.warmup):
        eng.iterate(params)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        eng.iterate(params)
    e1.record()
    torch.cuda.synchronize()
    warm_ms = e0.elapsed_time(e1) / args.steps
    eng.iterate(params, want_record=True)  # lets the engine see that no query needs the fallback
    st_warm = eng.time_stages(params, args.steps, False)

    # ---- sustained load for the clock record (~1.5 s of the same step)
    t_load0 = time.time()
    while time.time() - t_load0 < 1.5:
        for _ in range(200):
            eng.iterate(params)
        torch.cuda.synchronize()
    t_load1 = time.time()

    # ---- end to end through the public API on pinned host arrays
    def one_e2e():
        res = sb.register(Xf_pin.numpy(), Xm_pin.numpy(), correspondences=K, engine=eng,
                          transform_out=out_pin.numpy(), want_normals=False)  # = what simpleicp() does
        tab = None
        if world > 1:
            last = res.records[res.iterations - 1]
            local_rec = batch.pack_record(res.H, res.iterations, last["n_kept"], last["mean_res"], last["std_res"])[None]
            tab = batch.gather_records(local_rec, world, world, rank, dist, torch.device("cuda", local))
        return res, tab

    one_e2e()  # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    its = 0
    for _ in range(args.e2e_steps):
        res, tab = one_e2e()
        its += res.iterations
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s, float(its)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        e2e_s, its_total = float(tmax[0].item()), float(tsum[1].item())
    else:
        its_total = float(its)
    e2e_value = K * its_total / e2e_s
    dH_true = float(np.linalg.norm(res.H - H_true))
    tm = eng.timings()

    # ---- the linearised variant on the same pair: same two kernels
    # per iteration, one linear solve instead of the Gauss-Newton loop; reported beside the headline
    variants = None
    if world == 1:
        eng.set_option("variant", 1)
        eng.iterate(params, x_in=np.zeros(6), want_record=True)
        for _ in range(max(args.warmup, 3)):
            eng.iterate(params, want_record=True)
        st_lin = eng.time_stages(params, args.steps, True)
        t0 = time.perf_counter()
        rl = sb.simpleicp_linearized(Xf_pin.numpy(), Xm_pin.numpy(), correspondences=K, engine=eng,
                                     transform_out=out_pin.numpy())
        torch.cuda.synchronize()
        lin_s = time.perf_counter() - t0
        eng.set_option("variant", 0)
        variants = {"linearized": {
            "ms_per_step": st_lin["iteration"], "value": K / (st_lin["iteration"] * 1e-3), "unit": UNIT,
            "kernels_ms_cold_l2": st_lin, "e2e_ms_per_registration": 1e3 * lin_s, "e2e_iterations": rl.iterations,
            "H_frobenius_vs_H_true": float(np.linalg.norm(rl.T - H_true)),
            "api": "simpleicp_b200.simpleicp_linearized(X_fix, X_mov, correspondences=K, engine=<reused>)"}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    clocks = sampler.summary(t_load0, t_load1)
    clocks["window"] = "1.5 s of back-to-back steps right after the timed region"

    # ---- roofline of the dominant kernel (cold-L2 per-launch time measured above)
    b_match, b_rs = algorithmic_bytes(n, K, n_kept)
    peaks = {}
    try:
        peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text())
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    if st["match_grid"] >= st["reject_solve"]:
        dom, dom_ms, dom_bytes = "k_match_grid_coop", st["match_grid"], b_match
    else:
        dom, dom_ms, dom_bytes = "k_reject_solve", st["reject_solve"], b_rs
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    prof = REPO / "profiles" / "ncu_traffic.json"
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get(dom)
        except ValueError:
            pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                "kernels_ms_cold_l2": st, "kernels_ms_warm_l2": st_warm,
                "iteration_bytes": b_match + b_rs,
                "note": "the search structure is L2-resident in the real loop; see DESIGN.md for the L2/latency view"}

    # ---- CPU baseline beside it: oracle port on a bounded sample of the same workload (N = 1)
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import simpleicp_oracle as O

        tr = O.Trace()
        n_it = 3
        t0 = time.perf_counter()
        H_o, _, x_o, _, _ = O.simpleicp(X_fix, X_mov, correspondences=K, min_change=0.0, max_iterations=n_it, trace=tr)
        t_cpu = time.perf_counter() - t0
        loop = float(np.sum(tr.iter_seconds))
        cpu = {"value": K * n_it / loop, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"{n_it} ICP iterations of the oracle port on the full C3 pair (K={K}); loop only "
                         f"({loop:.1f} s; normals {tr.timings['normals']:.1f} s and final transform excluded; "
                         f"whole run {t_cpu:.1f} s); cKDTree queries on all cores, the rest single-threaded"}
        # live parity check on the same inputs: 3 lock-step iterations, oracle normals injected
        full = [np.full(n, np.nan, dtype=np.float32) for _ in range(4)]
        for a in range(3):
            full[a][tr.idx_sel] = tr.normals[:, a]
        full[3][tr.idx_sel] = tr.planarity
        r3 = sb.register(X_fix, X_mov, correspondences=K, min_change=0.0, max_iterations=n_it,
                         normals=tuple(full), engine=eng)
        parity = {"H_frobenius_vs_oracle_after_3_iterations": float(np.linalg.norm(r3.H - H_o)),
                  "kept_gpu": r3.records[-1]["n_kept"], "kept_oracle": int(tr.iterations[-1].keep.sum())}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C3 synthetic surface pair {n}<->{n} per GPU, correspondences={K}, neighbors=10, "
                               "min_planarity=0.3 (BASELINE.json configs[2])",
                   "l2": "flushed: 256 MiB buffer overwritten before every timed step",
                   "nn_engine": "grid (float64) + TMA brute-force fallback", "kept_per_iteration": n_kept},
        "value_l2_warm_queued": world * K / (warm_ms * 1e-3),
        "ms_per_step_l2_warm_queued": warm_ms,
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 24,
                "d2h_bytes_per_step": n * 24 + 8 * n_kept,
                "ms_per_registration": 1e3 * e2e_s / args.e2e_steps, "iterations": its_total / (args.e2e_steps * world),
                "registrations": args.e2e_steps,
                "api": "simpleicp_b200.register(X_fix, X_mov, correspondences=K, engine=<reused>, transform_out=<pinned>)",
                "stage_ms": {k: v for k, v in tm.items() if k.endswith("_ms")},
                "H_frobenius_vs_H_true": dH_true},
        "gpu_launches": int(launches), "clocks": clocks, "parity": parity, "variants": variants,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--correspondences", type=int, default=100_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "b200" else 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
