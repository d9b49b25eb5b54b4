"""Generate tests/golden/*.npz by running the UNMODIFIED reference package.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # all six reference test configs + fixtures

The reference (python/simpleicp, imported from /root/reference/python, never copied) is run
with oracle/lmfit_standin on sys.path for its one absent dependency, and instrumented from the
outside (method wrappers) to record every stage of the hot path:

    idx_sel, normals/planarity (float32), per iteration: pc2_idx, point-to-plane distances,
    kept pc1_idx, estimated x, residual mean/std/len, and the final H, sigma, residuals.

The configs replayed are exactly those of python/simpleicp/tests/test_simpleicp.py:35-104.
Input clouds small enough to travel are stored losslessly as scaled int32 (the .xyz files
carry 2-4 decimals): tests/golden/data_<name>.npz.
"""
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(REPO / "oracle" / "lmfit_standin"))
sys.path.insert(0, str(REF / "python"))

from simpleicp import PointCloud, SimpleICP  # noqa: E402  (the reference package)
from simpleicp import corrpts, optimization, pointcloud  # noqa: E402

GOLD = REPO / "tests" / "golden"

CONFIGS = {
    "dragon": ("dragon1.xyz", "dragon2.xyz", {}),
    "bunny": ("bunny_part1.xyz", "bunny_part2.xyz", {"max_overlap_distance": 1}),
    "multisensor": (
        "multisensor_lidar.xyz",
        "multisensor_radar.xyz",
        {
            "max_overlap_distance": 1,
            "rbp_observed_values": (-0.5, 0.0, 0.0, 0.0, 0.0, 0.0),
            "rbp_observation_weights": (np.inf, np.inf, 0.0, 0.0, 0.0, 0.0),
        },
    ),
    "webots": (
        "webots1.xyz",
        "webots2.xyz",
        {
            "neighbors": 40,
            "max_overlap_distance": 0.5,
            "rbp_observed_values": (0.0, 0.0, -60.0, -0.05, -0.09, 0.0),
            "rbp_observation_weights": (0.0, 0.0, 0.0, 0.0, 0.0, 0.0),
        },
    ),
    "airborne": ("airborne_lidar1.xyz", "airborne_lidar2.xyz", {}),
    "terrestrial": ("terrestrial_lidar1.xyz", "terrestrial_lidar2.xyz", {}),
    # extra configs exercising branches the reference tests do not reach
    "dragon_observed": (
        "dragon1.xyz",
        "dragon2.xyz",
        {
            "distance_weights": None,
            "rbp_observed_values": (-1.0, -2.0, -3.0, -0.2, -0.4, -0.6),
            "rbp_observation_weights": (50.0, 50.0, 0.0, 20.0, np.inf, 0.0),
        },
    ),
    # pc_mov carries normal columns (pc_mov.estimate_normals before run): the second branch of
    # CorrPts.reject_wrt_planarity (corrpts.py:157-162), which none of the reference tests reaches
    "dragon_movnormals": ("dragon1.xyz", "dragon2.xyz", {"_mov_normals": 10}),
}
TRAVEL = ("dragon", "bunny", "multisensor", "webots")  # inputs committed as fixtures
TRAVEL_LARGE = ("airborne", "terrestrial")  # > 1M points each: delta + byte-plane + LZMA (see below)
DEBUG_DUMPS = ("bunny",)  # configs whose debug_dirpath output is captured (debug_<name>.npz)


def save_cloud_fixture(name, X_fix, X_mov):
    for dec in range(0, 7):
        s = 10 ** dec
        if np.array_equal(np.round(X_fix * s) / s, X_fix) and np.array_equal(
            np.round(X_mov * s) / s, X_mov
        ):
            break
    else:
        raise RuntimeError("cloud is not a fixed-decimal text file")
    fi = np.round(X_fix * s).astype(np.int32)
    mi = np.round(X_mov * s).astype(np.int32)
    assert np.array_equal(fi / s, X_fix) and np.array_equal(mi / s, X_mov)
    np.savez_compressed(GOLD / f"data_{name}.npz", fix=fi, mov=mi, scale=np.int64(s))


def encode_cloud(X):
    """Lossless compact form of a fixed-decimal scan-ordered cloud: scaled int32, first
    differences along the scan, zig-zag, the four byte planes of every column stored one after
    the other (the high planes are almost all zero), LZMA.  5.7 MB for the 1.34 M-point airborne
    cloud (32 MB of text).  tests/conftest.py::decode_cloud is the inverse."""
    import lzma

    for dec in range(0, 7):
        s = 10 ** dec
        if np.array_equal(np.round(X * s) / s, X):
            break
    else:
        raise RuntimeError("cloud is not a fixed-decimal text file")
    I = np.round(X * s).astype(np.int64)
    assert np.abs(I).max() < 2 ** 30 and np.array_equal(I / s, X)
    D = np.diff(I, axis=0, prepend=0).astype(np.int32)
    Z = ((D << 1) ^ (D >> 31)).astype(np.uint32)
    planes = np.ascontiguousarray(Z.T).view(np.uint8).reshape(3, -1, 4).transpose(0, 2, 1)
    blob = lzma.compress(np.ascontiguousarray(planes).tobytes(), preset=9)
    return np.frombuffer(blob, dtype=np.uint8), s


def save_large_cloud_fixture(name, X_fix, X_mov):
    fb, fs = encode_cloud(X_fix)
    mb, ms = encode_cloud(X_mov)
    np.savez(GOLD / f"data_{name}.npz", codec=np.array("delta-zigzag-byteplane-lzma"),
             fix_blob=fb, mov_blob=mb, n_fix=np.int64(len(X_fix)), n_mov=np.int64(len(X_mov)),
             fix_scale=np.int64(fs), mov_scale=np.int64(ms))


def read_xyz(path):
    import pandas as pd

    X = pd.read_csv(path, sep=r"\s+", header=None).to_numpy(dtype=np.float64)
    G = np.genfromtxt(path, max_rows=2000)  # the reference's reader on a sample: same doubles
    assert np.array_equal(G, X[: len(G)])
    return X


def capture_debug(name, X_fix, X_mov, kwargs):
    """Run the unmodified reference with debug_dirpath and keep a compact description of every
    file it writes (simpleicp.py:141-143, 189-221, 317-320; corrpts.py:213-237;
    pointcloud.py:219-226): name, header line, row count, first and last text lines, column sums,
    and the full numeric content of the correspondence dumps."""
    import tempfile

    import logging

    out = {}
    lines = []

    class Grab(logging.Handler):
        def emit(self, record):
            lines.append(record.getMessage())

    grab = Grab()
    ref_log = logging.getLogger("simpleicp")
    ref_log.addHandler(grab)
    ref_log.setLevel(logging.INFO)
    with tempfile.TemporaryDirectory() as td:
        pc_fix = PointCloud(X_fix, columns=["x", "y", "z"])
        pc_mov = PointCloud(X_mov, columns=["x", "y", "z"], copy=True)
        icp = SimpleICP(verbose=False)
        icp.add_point_clouds(pc_fix, pc_mov)
        try:
            icp.run(debug_dirpath=td, **kwargs)
        finally:
            ref_log.removeHandler(grab)
        out["log_lines"] = np.array([ln.replace(td, "<debug_dir>") for ln in lines])
        names = sorted(p.name for p in Path(td).iterdir())
        out["names"] = np.array(names)
        for nm in names:
            lines = (Path(td) / nm).read_text().splitlines()
            key = nm.replace(".", "_")
            out[f"{key}__header"] = np.array(lines[0])
            out[f"{key}__rows"] = np.int64(len(lines) - 1)
            out[f"{key}__head"] = np.array(lines[1:4])
            out[f"{key}__tail"] = np.array(lines[-1])
            data = np.loadtxt(Path(td) / nm, comments="//")
            out[f"{key}__colsum"] = data.sum(axis=0)
            if "correspondences" in nm:
                out[f"{key}__data"] = data
    np.savez_compressed(GOLD / f"debug_{name}.npz", **out)
    print(f"{name}: debug dump captured, {len(names)} files")


def capture(name, file1, file2, kwargs):
    if name in TRAVEL_LARGE:
        X_fix, X_mov = read_xyz(REF / "data" / file1), read_xyz(REF / "data" / file2)
        save_large_cloud_fixture(name, X_fix, X_mov)
    else:
        X_fix = np.genfromtxt(REF / "data" / file1)
        X_mov = np.genfromtxt(REF / "data" / file2)
    if name in TRAVEL:
        save_cloud_fixture(name, X_fix, X_mov)
    if name in DEBUG_DUMPS:
        capture_debug(name, X_fix, X_mov, kwargs)
    if "--fixtures-only" in sys.argv:
        return

    rec = {"it_pc2_idx": [], "it_dist": [], "it_kept_pc1": [], "it_x": [], "it_res_stats": [],
           "it_w": []}

    orig_match = corrpts.CorrPts.match
    orig_rej = corrpts.CorrPts.reject_wrt_point_to_plane_distances
    orig_est = optimization.SimpleICPOptimization.estimate_parameters
    orig_nrm = pointcloud.PointCloud.estimate_normals
    orig_inrange = pointcloud.PointCloud.select_in_range

    def match(self):
        orig_match(self)
        rec["it_pc2_idx"].append(self._df["pc2_idx"].to_numpy().astype(np.int32))
        rec["it_dist"].append(self._df["point_to_plane_distances"].to_numpy().copy())
        if "idx_sel" not in rec:
            rec["idx_sel"] = self._df["pc1_idx"].to_numpy().astype(np.int32)

    def rej(self):
        orig_rej(self)
        rec["it_kept_pc1"].append(self._df["pc1_idx"].to_numpy().astype(np.int32))

    def est(self):
        r = orig_est(self)
        rec["it_x"].append(np.array(self.rbp.get_parameter_attributes_as_list("estimated_value")))
        rec["it_res_stats"].append(np.array([len(r), np.mean(r), np.std(r)]))
        rec["it_w"].append(float(self._distance_weights))
        return r

    def nrm(self, neighbors):
        orig_nrm(self, neighbors)
        idx = self.idx_selected
        rec["normals"] = np.column_stack(
            [self[c].to_numpy()[idx] for c in ("nx", "ny", "nz")]
        ).astype(np.float32)
        rec["planarity"] = self["planarity"].to_numpy()[idx].astype(np.float32)

    def inrange(self, X, max_range):
        orig_inrange(self, X, max_range)
        rec["idx_overlap"] = self.idx_selected.astype(np.int32)

    kwargs = dict(kwargs)
    mov_neighbors = kwargs.pop("_mov_normals", None)
    pc_mov_pre = None
    if mov_neighbors:
        # all movable points are selected by default: normals for every one of them, in pc_mov's own frame
        pc_mov_pre = PointCloud(X_mov, columns=["x", "y", "z"], copy=True)
        pc_mov_pre.estimate_normals(mov_neighbors)
        rec["mov_normals"] = np.column_stack(
            [pc_mov_pre[c].to_numpy() for c in ("nx", "ny", "nz")]).astype(np.float32)
        rec["mov_planarity"] = pc_mov_pre["planarity"].to_numpy().astype(np.float32)

    corrpts.CorrPts.match = match
    corrpts.CorrPts.reject_wrt_point_to_plane_distances = rej
    optimization.SimpleICPOptimization.estimate_parameters = est
    pointcloud.PointCloud.estimate_normals = nrm
    pointcloud.PointCloud.select_in_range = inrange
    try:
        pc_fix = PointCloud(X_fix, columns=["x", "y", "z"])
        pc_mov = pc_mov_pre if pc_mov_pre is not None else PointCloud(X_mov, columns=["x", "y", "z"], copy=True)
        icp = SimpleICP(verbose=False)
        icp.add_point_clouds(pc_fix, pc_mov)
        t = time.time()
        H, X_t, rbp, res = icp.run(**kwargs)
        wall = time.time() - t
    finally:
        corrpts.CorrPts.match = orig_match
        corrpts.CorrPts.reject_wrt_point_to_plane_distances = orig_rej
        optimization.SimpleICPOptimization.estimate_parameters = orig_est
        pointcloud.PointCloud.estimate_normals = orig_nrm
        pointcloud.PointCloud.select_in_range = orig_inrange

    n_it = len(rec["it_x"])
    K = len(rec["idx_sel"])
    kept = np.zeros((n_it, K), dtype=bool)
    pos = {int(v): i for i, v in enumerate(rec["idx_sel"])}
    for i, kp in enumerate(rec["it_kept_pc1"]):
        kept[i, [pos[int(v)] for v in kp]] = True
    out = dict(
        H=H,
        x=np.array(rbp.get_parameter_attributes_as_list("estimated_value")),
        sigma=np.array(rbp.get_parameter_attributes_as_list("estimated_uncertainty")),
        residuals=res,
        idx_sel=rec["idx_sel"],
        normals=rec["normals"],
        planarity=rec["planarity"],
        it_pc2_idx=np.stack(rec["it_pc2_idx"]),
        it_dist=np.stack(rec["it_dist"]),
        it_keep=kept,
        it_x=np.stack(rec["it_x"]),
        it_res_stats=np.stack(rec["it_res_stats"]),
        it_w=np.array(rec["it_w"]),
        X_mov_t_head=X_t[:64].copy(),
        X_mov_t_sum=X_t.sum(axis=0),
        wall_s=np.float64(wall),
        n_fix=np.int64(len(X_fix)),
        n_mov=np.int64(len(X_mov)),
    )
    if "idx_overlap" in rec:
        out["idx_overlap"] = rec["idx_overlap"]
    if "mov_normals" in rec:
        out["mov_normals"], out["mov_planarity"] = rec["mov_normals"], rec["mov_planarity"]
    kw = {k: v for k, v in kwargs.items()}
    out["kwargs_repr"] = np.array(repr(kw))
    np.savez_compressed(GOLD / f"ref_{name}.npz", **out)
    print(f"{name}: {n_it} iterations, kept {[int(k.sum()) for k in kept]}, wall {wall:.2f}s")
    print(np.array2string(H, precision=13))


if __name__ == "__main__":
    GOLD.mkdir(parents=True, exist_ok=True)
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or list(CONFIGS)
    for nm in names:
        capture(nm, *CONFIGS[nm])
