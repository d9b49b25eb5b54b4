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
}
TRAVEL = ("dragon", "bunny", "multisensor", "webots")  # inputs committed as fixtures


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


def capture(name, file1, file2, kwargs):
    X_fix = np.genfromtxt(REF / "data" / file1)
    X_mov = np.genfromtxt(REF / "data" / file2)
    if name in TRAVEL:
        save_cloud_fixture(name, X_fix, X_mov)

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

    corrpts.CorrPts.match = match
    corrpts.CorrPts.reject_wrt_point_to_plane_distances = rej
    optimization.SimpleICPOptimization.estimate_parameters = est
    pointcloud.PointCloud.estimate_normals = nrm
    pointcloud.PointCloud.select_in_range = inrange
    try:
        pc_fix = PointCloud(X_fix, columns=["x", "y", "z"])
        pc_mov = PointCloud(X_mov, columns=["x", "y", "z"], copy=True)
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
    kw = {k: v for k, v in kwargs.items()}
    out["kwargs_repr"] = np.array(repr(kw))
    np.savez_compressed(GOLD / f"ref_{name}.npz", **out)
    print(f"{name}: {n_it} iterations, kept {[int(k.sum()) for k in kept]}, wall {wall:.2f}s")
    print(np.array2string(H, precision=13))


if __name__ == "__main__":
    GOLD.mkdir(parents=True, exist_ok=True)
    names = sys.argv[1:] or list(CONFIGS)
    for nm in names:
        capture(nm, *CONFIGS[nm])
