"""CPU: the C-ABI library loads and exports every declared symbol; host-side logic."""
import os
import re
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from conftest import REPO, load_pair

import simpleicp_b200 as sb
import importlib

from simpleicp_b200 import _capi, batch, mathutils, pointcloud

drv = importlib.import_module("simpleicp_b200.simpleicp")
from oracle import simpleicp_oracle as O


def test_header_symbols_exported():
    hdr = (REPO / "include" / "sicp_b200.h").read_text()
    declared = set(re.findall(r"\b(sicp_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"sicp_ctx"}
    assert declared == set(_capi.SYMBOLS)
    lib = _capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sicp_abi_version() == _capi.ABI_VERSION


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every struct, as gcc sees include/sicp_b200.h, equal the ctypes mirror."""
    import ctypes as C

    src = tmp_path / "layout.c"
    src.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "sicp_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(sicp_lsq_params), sizeof(sicp_run_params),
         sizeof(sicp_iter_record), sizeof(sicp_run_result), sizeof(sicp_timings));
  printf("%zu %zu %zu %zu %zu\n", offsetof(sicp_run_params, lsq), offsetof(sicp_iter_record, x),
         offsetof(sicp_iter_record, lm_iterations), offsetof(sicp_run_result, H),
         offsetof(sicp_run_result, loop_ms));
  printf("%zu %zu %zu\n", sizeof(sicp_register_params), offsetof(sicp_register_params, max_overlap_distance),
         offsetof(sicp_register_params, run));
  return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(REPO / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(v) for v in out]
    want = [C.sizeof(_capi.LsqParams), C.sizeof(_capi.RunParams), C.sizeof(_capi.IterRecord),
            C.sizeof(_capi.RunResult), C.sizeof(_capi.Timings), _capi.RunParams.lsq.offset,
            _capi.IterRecord.x.offset, _capi.IterRecord.lm_iterations.offset,
            _capi.RunResult.H.offset, _capi.RunResult.loop_ms.offset,
            C.sizeof(_capi.RegisterParams), _capi.RegisterParams.max_overlap_distance.offset,
            _capi.RegisterParams.run.offset]
    assert got == want


def test_no_cpu_fallback_without_device():
    from conftest import has_gpu

    if has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.SicpError):
        _capi.Engine()
    X_fix, X_mov = load_pair("bunny")
    with pytest.raises(_capi.SicpError):
        sb.simpleicp(X_fix, X_mov)


def test_cli_and_linearized_fail_loudly_without_device(tmp_path):
    """No CPU path anywhere: the command line and the linearised front end error out too."""
    from conftest import has_gpu

    if has_gpu():
        pytest.skip("a GPU is present")
    X_fix, X_mov = load_pair("bunny")
    with pytest.raises(_capi.SicpError):
        sb.simpleicp_linearized(X_fix, X_mov)
    f = tmp_path / "a.xyz"
    sb.write_xyz(f, X_fix[:100], decimals=4, header=False)
    cli = REPO / "simpleicp_b200" / "sicp_cli"
    r = subprocess.run([str(cli), "-f", str(f), "-m", str(f)], capture_output=True, text=True)
    assert r.returncode == 1 and "Caught exception: no CUDA device available" in r.stderr
    h = subprocess.run([str(cli), "--help"], capture_output=True, text=True)
    assert h.returncode == 0 and "--variant" in h.stdout and "--max_overlap_distance" in h.stdout


def test_product_does_not_import_oracle():
    for f in (REPO / "simpleicp_b200").glob("*.py"):
        assert "oracle" not in f.read_text().replace("oracle/make_golden", ""), f


def test_mathutils_against_oracle():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = rng.uniform(-3, 3, 3)
        R = mathutils.euler_angles_to_rotation_matrix(*a)
        assert np.array_equal(R, O.euler_angles_to_rotation_matrix(*a))
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-14)
    a = np.array([0.3, -0.2, 0.5])
    R = mathutils.euler_angles_to_rotation_matrix(*a)
    np.testing.assert_allclose(mathutils.rotation_matrix_to_euler_angles(R), a, atol=1e-14)
    H = mathutils.create_homogeneous_transformation_matrix(R, [1, 2, 3])
    assert np.array_equal(H, O.create_homogeneous_transformation_matrix(R, [1, 2, 3]))
    X = rng.normal(size=(5, 3))
    Xh = mathutils.euler_coord_to_homogeneous_coord(X)
    np.testing.assert_allclose(mathutils.homogeneous_coord_to_euler_coord(Xh @ H.T), O.transform_by_H(X, H), atol=1e-15)


def test_subsample_matches_reference_rule():
    for m, n in [(100000, 1000), (7943, 1000), (1001, 1000), (10, 3), (5, 5)]:
        ref = O.select_n_points(np.arange(m), n)
        if m > n:
            got = np.unique(pointcloud.subsample_indices(m, n))
        else:
            got = np.arange(m)
        assert np.array_equal(ref, got)


def test_pointcloud_container():
    X = np.arange(30, dtype=float).reshape(10, 3)
    pc = sb.PointCloud(X, columns=["x", "y", "z"])
    assert pc.num_points == 10 and pc.num_selected_points == 10
    assert np.array_equal(pc.X, X) and np.array_equal(pc.x, X[:, 0])
    pc.select_n_points(4)
    assert list(pc.idx_selected) == [0, 3, 6, 9]
    assert np.array_equal(pc.X_selected, X[[0, 3, 6, 9]])
    pc.select_by_indices([3, 4, 9])
    assert list(pc.idx_selected) == [3, 9]
    pc.unselect_all_points()
    assert pc.num_selected_points == 0
    pc.select_all_points()
    assert pc.num_selected_points == 10
    pc.set_normals(np.array([1, 2]), [0.0, 1.0], [0.0, 0.0], [1.0, 0.0], [0.5, 0.9])
    assert str(pc["planarity"].dtype) == "Sparse[float32, nan]"  # stored as the reference does
    assert np.isnan(pc["nx"].to_numpy()[0]) and pc["planarity"].to_numpy()[2] == np.float32(0.9)
    with pytest.raises(sb.PointCloudException, match='"z" is missing'):
        sb.PointCloud(X[:, :2], columns=["x", "y"])


def test_sparse_normal_columns_equal_the_reference_construction():
    """PointCloud.set_normals assembles the sparse float32 columns from (index, value) pairs; the
    result must be the array the reference builds (pointcloud.py:200-203: SparseArray of a dense
    NaN-filled column): same values, dtype, sparse index, and NaN values not stored."""
    import pandas as pd

    from simpleicp_b200.pointcloud import _sparse_f32_column

    rng = np.random.default_rng(3)
    n = 5000
    for idx in (np.sort(rng.choice(n, 700, replace=False)), np.array([0]), np.array([n - 1]), np.arange(n),
                rng.permutation(n)[:50], np.array([], dtype=np.int64)):  # unsorted / empty: dense route
        vals = rng.standard_normal(idx.size).astype(np.float32)
        if idx.size > 3:
            vals[2] = np.nan  # a degenerate neighbourhood: NaN normal
        col = np.full(n, np.nan, dtype=np.float32)
        col[idx] = vals
        ref = pd.arrays.SparseArray(col)
        got = _sparse_f32_column(n, np.asarray(idx, dtype=np.int64), vals)
        assert got.dtype == ref.dtype and str(got.dtype) == "Sparse[float32, nan]"
        assert np.array_equal(np.asarray(got), np.asarray(ref), equal_nan=True)
        assert got.sp_index.equals(ref.sp_index) and np.array_equal(got.sp_values, ref.sp_values)
    pc = sb.PointCloud(rng.random((n, 3)), columns=["x", "y", "z"])
    idx = np.arange(0, n, 7)
    v = rng.random(idx.size).astype(np.float32)
    pc.set_normals(idx, v, v, v, v)
    assert pc["nx"].to_numpy()[7] == v[1] and np.isnan(pc["nz"].to_numpy()[8])
    assert pc["planarity"].sparse.density == pytest.approx(idx.size / n)


def test_argument_checks_match_reference_messages():
    X = np.zeros((10, 3))
    icp = sb.SimpleICP(verbose=False)
    icp.add_point_clouds(sb.PointCloud(X, columns=["x", "y", "z"]), sb.PointCloud(X, columns=["x", "y", "z"]))
    with pytest.raises(sb.SimpleICPException, match="distance_weights must be > 0"):
        icp.run(distance_weights=0)
    with pytest.raises(sb.SimpleICPException, match="rbp_observed_values must have exactly 6"):
        icp.run(rbp_observed_values=(0, 0))
    with pytest.raises(sb.SimpleICPException, match="rbp_observation_weights must have exactly 6"):
        icp.run(rbp_observation_weights=(0, 0))
    with pytest.raises(sb.SimpleICPException, match="must be >= 0"):
        icp.run(rbp_observation_weights=(-1, 0, 0, 0, 0, 0))
    with pytest.raises(sb.SimpleICPException, match="must be finite"):
        icp.run(rbp_observation_weights=(np.inf,) * 6)


def test_observed_values_conversion_keeps_reference_quirk():
    obs = drv._observed_in_radians((0.0, 0.0, -60.0, -0.05, -0.09, 0.0))
    assert obs[2] == -60.0 * np.pi / 180
    assert drv._observed_in_radians((0, 0, 60, 0, 0, 0)).dtype.kind == "i"  # as the reference


def test_rigid_body_parameters():
    rbp = sb.RigidBodyParameters()
    rbp.set_parameter_attributes_from_list("estimated_value", [0.1, -0.2, 0.3, 1, 2, 3])
    assert np.array_equal(rbp.H, O.rbp_to_H([0.1, -0.2, 0.3, 1, 2, 3]))
    assert rbp.alpha1.estimated_value_scaled == pytest.approx(0.1 * 180 / np.pi)
    assert rbp.get_parameter_attributes_as_list("estimated_value")[3] == 1
    assert np.isnan(rbp.tx.estimated_uncertainty)


def test_shard_pairs_partition():
    for n, w in [(512, 8), (7, 3), (2, 4), (0, 2)]:
        seen = sorted(i for r in range(w) for i in batch.shard_pairs(n, w, r))
        assert seen == list(range(n))
    assert batch.shard_pairs(10, 4, 1) == [1, 5, 9]


def test_tile_slabs():
    rng = np.random.default_rng(0)
    Xf = rng.uniform(0, 100, size=(5000, 3))
    Xm = rng.uniform(0, 100, size=(4000, 3))
    slabs = batch.tile_slabs(Xf, Xm, 4, overlap=2.5)
    assert len(slabs) == 4
    cuts = np.quantile(Xf[:, 0], [0.25, 0.5, 0.75])
    # every fixed point is in at least one slab, interior points near a cut are in two
    total = sum(len(a) for a, _ in slabs)
    assert total > len(Xf)
    for s, (a, b) in enumerate(slabs):
        lo = -np.inf if s == 0 else cuts[s - 1] - 2.5
        hi = np.inf if s == 3 else cuts[s] + 2.5
        assert ((a[:, 0] >= lo) & (a[:, 0] < hi)).all() and ((b[:, 0] >= lo) & (b[:, 0] < hi)).all()
        assert abs(len(a) - (len(Xf) / 4 + (0.05 if 0 < s < 3 else 0.025) * len(Xf))) < 0.03 * len(Xf)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from simpleicp_b200 import batch

class R:  # stand-in for a registration result (the collective and sharding are what is tested)
    def __init__(self, i):
        self.H = np.eye(4) * (i + 1); self.iterations = i + 2
        self.records = [dict(n_kept=0, mean_res=0.0, std_res=0.0)] * (i + 1) + [dict(n_kept=100 + i, mean_res=0.5 * i, std_res=0.25 * i)]

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 7
calls = []
def fake_register(Xf, Xm, **kw):
    calls.append(int(Xf[0, 0])); return R(int(Xf[0, 0]))
table = batch.simpleicp_batch(lambda i: (np.full((1, 3), float(i)), np.zeros((1, 3))), n, rank=rank,
                              world_size=world, dist=dist, register_fn=fake_register)
assert calls == batch.shard_pairs(n, world, rank), calls
# a pair that cannot be registered fails ALONE: its rank still joins the collective, every rank gets
# the full table and the same BatchError
class TooFew(Exception):
    code = 3
def flaky(Xf, Xm, **kw):
    if int(Xf[0, 0]) == 4:
        raise TooFew("Too few correspondences! ...")
    return R(int(Xf[0, 0]))
try:
    batch.simpleicp_batch(lambda i: (np.full((1, 3), float(i)), np.zeros((1, 3))), n, rank=rank,
                          world_size=world, dist=dist, register_fn=flaky)
    raise SystemExit("BatchError expected")
except batch.BatchError as e:
    assert e.failed.tolist() == [4] and e.table.shape == (n, 20) and np.isnan(e.table[4, :16]).all()
    assert e.table[4, 16] == -3 and np.array_equal(e.table[5, :16].reshape(4, 4), np.eye(4) * 6)
t2 = batch.simpleicp_batch(lambda i: (np.full((1, 3), float(i)), np.zeros((1, 3))), n, rank=rank,
                           world_size=world, dist=dist, register_fn=flaky, on_error="nan")
assert t2[4, 16] == -3 and t2[3, 16] == 5
for i in range(n):
    assert np.array_equal(table[i, :16].reshape(4, 4), np.eye(4) * (i + 1))
    assert table[i, 16] == i + 2 and table[i, 17] == 100 + i and table[i, 18] == 0.5 * i and table[i, 19] == 0.25 * i
dist.destroy_process_group()
print("ok", rank)
'''


def test_batch_gloo_world2(tmp_path):
    """N>1 path on CPU: world_size-2 gloo run of the sharding + all-gather of the H records."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), str(REPO)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_synthetic_generators_equal_the_oracles():
    """bench.py's product arm takes its inputs from simpleicp_b200.synthetic (no oracle import);
    the arrays are the oracle's bit for bit, so both arms see the same workload."""
    from simpleicp_b200 import synthetic

    a, b = synthetic.c3_pair(3000), O.c3_pair(3000)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    Xf, Xm, H = synthetic.c5_pair(3, 2000)
    assert np.array_equal(Xf, O.surface(2000, 10_006, extent=30.0))
    assert np.array_equal(Xm, O.transform_by_H(O.surface(2000, 10_007, extent=30.0), np.linalg.inv(H)))
    src = (REPO / "bench.py").read_text()
    b200_arm = src[src.index("def run_b200"):src.index("# ---- CPU baseline beside it")]
    assert "oracle" not in b200_arm  # only the cpu_baseline / parity legs of bench.py touch oracle/


def test_angle_rejection_needs_movable_normals():
    """Argument check of the facade happens before any device work."""
    import simpleicp_b200 as sb
    X = np.random.default_rng(0).random((50, 3))
    with pytest.raises(sb.SimpleICPException, match="needs the normals of the movable"):
        sb.simpleicp(X, X, max_angle_between_normals=10.0)
    with pytest.raises(sb.SimpleICPException, match=r"within \[0, 90\]"):
        sb.simpleicp(X, X, mov_normals=(X[:, 0],) * 4, max_angle_between_normals=120.0)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside the product arm): ONE JSON
    line on stdout with the contract's keys, the workload string of the product arm, the e2e and
    cpu_baseline objects of the tier's reference arm.  Small sizes: the line, not the number."""
    import json
    import subprocess

    import bench

    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--points", "20000", "--correspondences", "2000"], capture_output=True, text=True,
                       timeout=600, cwd=str(REPO))
    assert r.returncode == 0, r.stderr[-800:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # stdout is the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["metric"].startswith("correspondences/sec") and d["unit"] == "corr/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["config"]["workload"] == bench.workload_string(20000, 2000)   # same string as the product arm
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and e["value"] > 0 and e["unit"] == "corr/s"
