"""CPU oracle: NumPy/SciPy restatement of the Python simpleICP hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package
``simpleicp_b200`` never does (it fails loudly when its CUDA library is missing).

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference package from
``/root/reference/python`` (with ``oracle/lmfit_standin`` supplying the one absent, unvendored
dependency) and ``tests/test_oracle_golden.py`` checks this restatement against those captured
outputs (tests/golden/*.npz) stage by stage.

Every function cites the reference lines it follows (paths relative to /root/reference/python/
simpleicp/).  The restatement keeps the reference's arithmetic *order* (no fused ops, same
SciPy/NumPy calls: cKDTree, np.cov, np.linalg.eig, np.median, scipy least_squares TRF) and drops
only the pandas container.

Third-party arithmetic that is not under /root/reference:
  * lmfit (unpinned, setup.py:24) -> scipy.optimize.least_squares(method='trf', jac='2-point',
    ftol=xtol=gtol=1e-8, max_nfev=2*2000*(nvar+1)); restated in ``estimate_parameters``.
  * scipy.spatial.cKDTree, scipy.stats.median_abs_deviation (scale=1.0), numpy.linalg.eig.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy import spatial
from scipy.optimize import least_squares

PARAM_NAMES = ("alpha1", "alpha2", "alpha3", "tx", "ty", "tz")

# threads of the cKDTree queries (the reference passes workers=-1: all cores).  bench.py lowers it
# when it runs several reference processes side by side (one per GPU of the arm it is compared with).
WORKERS = -1


# --------------------------------------------------------------------------- mathutils.py
def euler_angles_to_rotation_matrix(a1: float, a2: float, a3: float) -> np.ndarray:
    """mathutils.py:39-68 — R = Rx(a1) Ry(a2) Rz(a3) written out element by element."""
    c1, s1 = np.cos(a1), np.sin(a1)
    c2, s2 = np.cos(a2), np.sin(a2)
    c3, s3 = np.cos(a3), np.sin(a3)
    return np.array(
        [
            [c2 * c3, -c2 * s3, s2],
            [c1 * s3 + s1 * s2 * c3, c1 * c3 - s1 * s2 * s3, -s1 * c2],
            [s1 * s3 - c1 * s2 * c3, s1 * c3 + c1 * s2 * s3, c1 * c2],
        ]
    )


def create_homogeneous_transformation_matrix(R: np.ndarray, t: Sequence[float]) -> np.ndarray:
    """mathutils.py:81-93."""
    H = np.eye(4)
    H[0:3, 0:3] = R
    H[0:3, 3] = np.asarray(t, dtype=float)
    return H


def rbp_to_H(x: Sequence[float]) -> np.ndarray:
    """optimization.py:335-350 (RigidBodyParameters.H)."""
    return create_homogeneous_transformation_matrix(
        euler_angles_to_rotation_matrix(x[0], x[1], x[2]), x[3:6]
    )


def transform_by_H(X: np.ndarray, H: np.ndarray) -> np.ndarray:
    """pointcloud.py:205-217 with mathutils.py:10-26 (homogeneous multiply, then divide by w)."""
    Xh = np.column_stack((X, np.ones(X.shape[0])))
    Xh = np.transpose(H @ Xh.T)
    return np.column_stack((Xh[:, 0] / Xh[:, 3], Xh[:, 1] / Xh[:, 3], Xh[:, 2] / Xh[:, 3]))


# --------------------------------------------------------------------------- pointcloud.py
def select_in_range(
    X_fix: np.ndarray, idx_sel: np.ndarray, X_other: np.ndarray, max_range: float
) -> np.ndarray:
    """pointcloud.py:149-171 — keep selected points whose NN in X_other is *strictly* closer
    than max_range (cKDTree distance_upper_bound semantics)."""
    kdtree = spatial.cKDTree(X_other)
    distances, _ = kdtree.query(
        X_fix[idx_sel], k=1, p=2, distance_upper_bound=max_range, workers=WORKERS
    )
    return idx_sel[np.isfinite(distances)]


def select_n_points(idx_sel: np.ndarray, n: int) -> np.ndarray:
    """pointcloud.py:132-147 — round-half-even(linspace) subsample of the selected indices."""
    if idx_sel.size > n:
        sub = np.round(np.linspace(0, idx_sel.size - 1, n)).astype(int)
        # the reference re-marks a boolean column, which also de-duplicates and sorts
        return np.unique(idx_sel[sub])
    return idx_sel


def estimate_normals(
    X_fix: np.ndarray, idx_sel: np.ndarray, neighbors: int
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """pointcloud.py:173-203 — k-NN (self included) -> np.cov (ddof=1) -> np.linalg.eig ->
    normal = eigenvector of the smallest eigenvalue, planarity = (l_mid - l_min) / l_max,
    both stored as float32.  Returns (normals f32 [K,3], planarity f32 [K], idxNN [K,k])."""
    kdtree = spatial.cKDTree(X_fix)
    _, idxNN_all = kdtree.query(X_fix[idx_sel], k=neighbors, p=2, workers=WORKERS)
    if idxNN_all.ndim == 1:
        idxNN_all = idxNN_all[:, None]
    normals = np.full((idx_sel.size, 3), np.nan, dtype=np.float32)
    planarity = np.full((idx_sel.size,), np.nan, dtype=np.float32)
    for i, idxNN in enumerate(idxNN_all):
        pts = X_fix[idxNN, :]
        C = np.cov(pts.T, bias=False)
        eig_vals, eig_vecs = np.linalg.eig(C)
        order = eig_vals.argsort()[::-1]
        eig_vals = eig_vals[order]
        eig_vecs = eig_vecs[:, order]
        normals[i, 0] = eig_vecs[0, 2]
        normals[i, 1] = eig_vecs[1, 2]
        normals[i, 2] = eig_vecs[2, 2]
        planarity[i] = (eig_vals[1] - eig_vals[2]) / eig_vals[0]
    return normals, planarity, idxNN_all


# --------------------------------------------------------------------------- corrpts.py
def match(
    X_fix: np.ndarray, idx_sel: np.ndarray, normals_f32: np.ndarray, X_mov_transformed: np.ndarray
) -> Tuple[np.ndarray, np.ndarray]:
    """corrpts.py:124-137 + 195-211 — kd-tree on the transformed movable cloud, 1-NN for each
    selected fixed point, signed point-to-plane distance (dx*nx + dy*ny) + dz*nz with the
    float32 normal promoted to float64."""
    kdtree = spatial.cKDTree(X_mov_transformed)
    _, idx_nn = kdtree.query(X_fix[idx_sel], k=1, p=2, workers=WORKERS)
    p1 = X_fix[idx_sel]
    p2 = X_mov_transformed[idx_nn]
    n = normals_f32.astype(np.float64)
    dx = p2[:, 0] - p1[:, 0]
    dy = p2[:, 1] - p1[:, 1]
    dz = p2[:, 2] - p1[:, 2]
    d = dx * n[:, 0] + dy * n[:, 1] + dz * n[:, 2]
    return idx_nn, d


def reject(
    d: np.ndarray, planarity_f32: np.ndarray, min_planarity: float,
    mov_planarity_f32: Optional[np.ndarray] = None,
    angle_ok: Optional[np.ndarray] = None,
) -> Tuple[np.ndarray, float, float]:
    """corrpts.py:139-163 then 165-188 — planarity mask first (float32 values compared in float64, NaN -> drop),
    then |d - median| <= 3 * MAD on the survivors with MAD *unscaled* (SciPy default scale=1.0;
    corrpts.py:186).  Returns (keep mask over K, median, mad).

    ``mov_planarity_f32`` (K values: the planarity of each correspondence's movable point) is
    the second branch of reject_wrt_planarity (corrpts.py:157-162), taken when pc_mov carries
    a planarity column.  ``angle_ok`` (K booleans) is NOT reference behaviour — the reference's
    reject_wrt_to_angle_between_normals raises NotImplementedError (corrpts.py:190-193); it is
    applied where the reference would call it: after the distance rejection (simpleicp.py:207)."""
    # corrpts.py:152-155 compares `Sparse[float32].to_numpy() >= min_planarity`; pandas hands the
    # sparse float32 column back as float64, so the comparison is a float64 one.
    keep1 = planarity_f32.astype(np.float64) >= min_planarity
    if mov_planarity_f32 is not None:  # corrpts.py:157-162
        with np.errstate(invalid="ignore"):
            keep1 = keep1 & (mov_planarity_f32.astype(np.float64) >= min_planarity)
    ds = d[keep1]
    if ds.size == 0:
        return np.zeros_like(keep1), np.nan, np.nan
    median = np.median(ds)
    mad = np.median(np.abs(ds - median))
    keep2 = np.abs(ds - median) <= 3 * mad
    keep = np.zeros_like(keep1)
    keep[np.flatnonzero(keep1)[keep2]] = True
    if angle_ok is not None:
        keep &= angle_ok
    return keep, float(median), float(mad)


def angle_between_normals_ok(normals_f32: np.ndarray, mov_normals_f32: np.ndarray, H: np.ndarray,
                             max_angle_deg: float) -> np.ndarray:
    """Extension (no reference behaviour, see ``reject``): |n_fix . (R n_mov)| >= cos(max angle),
    R the rotation of the transform the movable cloud is matched under; normals are axes, so the
    angle is taken modulo their sign; NaN normals fail."""
    rn = mov_normals_f32.astype(np.float64) @ H[:3, :3].T
    with np.errstate(invalid="ignore"):
        c = np.abs(np.sum(rn * normals_f32.astype(np.float64), axis=1))
        return c >= (0.0 if max_angle_deg >= 90.0 else np.cos(np.deg2rad(max_angle_deg)))


# --------------------------------------------------------------------------- optimization.py
def _residual_vector(
    x_full: np.ndarray,
    p1: np.ndarray,
    n1: np.ndarray,
    p2: np.ndarray,
    w: float,
    obs: np.ndarray,
    w_obs: np.ndarray,
) -> np.ndarray:
    """optimization.py:172-288 — weighted point-to-plane residuals of the *untransformed*
    matched movable points under the cumulative parameters, then one row per observed param."""
    H = rbp_to_H(x_full)
    p2t = transform_by_H(p2, H)
    dx = p2t[:, 0] - p1[:, 0]
    dy = p2t[:, 1] - p1[:, 1]
    dz = p2t[:, 2] - p1[:, 2]
    r = w * (dx * n1[:, 0] + dy * n1[:, 1] + dz * n1[:, 2])
    extra = [
        w_obs[j] * (x_full[j] - obs[j]) for j in range(6) if w_obs[j] > 0 and np.isfinite(w_obs[j])
    ]
    return np.concatenate((r, np.array(extra, dtype=float)))


def estimate_parameters(
    p1: np.ndarray,
    n1_f32: np.ndarray,
    p2: np.ndarray,
    w: float,
    x0: Sequence[float],
    obs: Sequence[float],
    w_obs: Sequence[float],
) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """optimization.py:65-124 via lmfit -> scipy TRF (see module docstring).
    Returns (x[6], unweighted distance residuals, weighted residual vector, jacobian)."""
    n1 = n1_f32.astype(np.float64)
    obs = np.asarray(obs, dtype=float)
    w_obs = np.asarray(w_obs, dtype=float)
    x_full = np.array(x0, dtype=float)
    vary = np.isfinite(w_obs)
    names = np.flatnonzero(vary)

    def fun(xv):
        x_full[names] = xv
        return _residual_vector(x_full, p1, n1, p2, w, obs, w_obs)

    if names.size:
        ret = least_squares(
            fun, x_full[names].copy(), bounds=(-np.inf, np.inf), max_nfev=2 * 2000 * (names.size + 1)
        )
        x_full[names] = ret.x
        jac = ret.jac
    else:
        jac = np.zeros((p1.shape[0], 0))
    res_w = _residual_vector(x_full, p1, n1, p2, w, obs, w_obs)
    return x_full.copy(), res_w[: p1.shape[0]] / w, res_w, jac


def estimate_parameter_uncertainties(
    res_w: np.ndarray, jac: np.ndarray, n_corr: int, w: float, w_obs: Sequence[float]
) -> np.ndarray:
    """optimization.py:126-170 — Cxx = s0^2 (A^T P A)^-1 with the reference's (unusual)
    P = diag(weights), A = J / weights.  The dense K x K np.diag of :155 is evaluated in its
    mathematically identical O(K) form (A.T * weights) @ A.  NaN for fixed parameters."""
    w_obs = np.asarray(w_obs, dtype=float)
    weights = np.full((n_corr,), float(w))
    for j in range(6):
        if w_obs[j] > 0 and np.isfinite(w_obs[j]):
            weights = np.append(weights, w_obs[j])
    A = jac / weights[:, None]
    r = res_w / weights
    N = (A.T * weights) @ A
    Qxx = np.linalg.inv(N)
    vPv = np.sum(weights * r ** 2)
    num_obs, num_prm = jac.shape
    s0 = np.sqrt(vPv / (num_obs - num_prm))
    Cxx = s0 ** 2 * Qxx
    sig = np.full(6, np.nan)
    k = 0
    for j in range(6):
        if np.isfinite(w_obs[j]):
            sig[j] = np.sqrt(Cxx[k, k])
            k += 1
    return sig


# --------------------------------------------------------------------------- simpleicp.py
def check_convergence_criteria(new: np.ndarray, old: np.ndarray, min_change: float) -> bool:
    """simpleicp.py:355-379 — relative change (in %) of mean and *population* std both below
    min_change."""

    def change(a, b):
        if b == 0:
            return 0.0 if a == 0 else np.inf
        return np.abs((a - b) / b * 100)

    return bool(
        change(np.mean(new), np.mean(old)) < min_change
        and change(np.std(new), np.std(old)) < min_change
    )


class OracleICPError(Exception):
    """Mirrors SimpleICPException (simpleicp.py:382)."""


@dataclass
class IterationTrace:
    H_in: np.ndarray
    pc2_idx: np.ndarray
    distances: np.ndarray
    keep: np.ndarray
    median: float
    mad: float
    x: np.ndarray
    H: np.ndarray
    residuals: np.ndarray
    w: float


@dataclass
class Trace:
    idx_overlap: Optional[np.ndarray] = None
    idx_sel: Optional[np.ndarray] = None
    normals: Optional[np.ndarray] = None
    planarity: Optional[np.ndarray] = None
    idxNN: Optional[np.ndarray] = None
    iterations: List[IterationTrace] = field(default_factory=list)
    sigma: Optional[np.ndarray] = None
    timings: Dict[str, float] = field(default_factory=dict)
    iter_seconds: List[float] = field(default_factory=list)  # wall time of each iteration
    light: bool = False  # True: keep only timings (no per-iteration arrays)


def simpleicp(
    X_fix: np.ndarray,
    X_mov: np.ndarray,
    correspondences: int = 1000,
    neighbors: int = 10,
    min_planarity: float = 0.3,
    max_overlap_distance: float = np.inf,
    min_change: float = 1.0,
    max_iterations: int = 100,
    distance_weights: Optional[float] = 1,
    rbp_observed_values: Sequence[float] = (0.0,) * 6,
    rbp_observation_weights: Sequence[float] = (0.0,) * 6,
    normals: Optional[Tuple[np.ndarray, np.ndarray]] = None,
    trace: Optional[Trace] = None,
    static_tree: bool = False,
    mov_normals: Optional[Tuple[np.ndarray, np.ndarray]] = None,
    max_angle_between_normals: Optional[float] = None,
):
    """SimpleICP.run (simpleicp.py:75-324), restated without pandas.

    ``normals=(normals_f32[K,3], planarity_f32[K])`` plays the role of the reference's
    "columns nx, ny, nz, planarity already present" hook (simpleicp.py:176-178).
    ``static_tree=True`` is NOT the reference algorithm: it builds the kd-tree once and moves
    the queries by inv(H) (the product's strategy); it exists so tests can show the two are
    equivalent.  ``mov_normals=(normals_f32[n_mov,3], planarity_f32[n_mov])`` stands for a pc_mov
    that carries normal columns (corrpts.py:157-162); ``max_angle_between_normals`` is the
    extension described in ``reject``.  Returns (H, X_mov_transformed, x[6], sigma[6], residuals).
    """
    X_fix = np.ascontiguousarray(X_fix, dtype=float)
    X2 = np.array(X_mov, dtype=float, copy=True)  # the reference mutates pc2 in place
    t0 = time.perf_counter()

    # simpleicp.py:326-353
    if distance_weights is not None and distance_weights <= 0:
        raise OracleICPError("distance_weights must be > 0.")
    if len(rbp_observed_values) != 6 or len(rbp_observation_weights) != 6:
        raise OracleICPError("rbp tuples must have exactly 6 elements.")
    if not all(w >= 0 for w in rbp_observation_weights):
        raise OracleICPError("All elements of rbp_observation_weights must be >= 0.")
    if not any(np.isfinite(rbp_observation_weights)):
        raise OracleICPError("At least one element in rbp_observation_weights must be finite.")

    # simpleicp.py:146-156 (np.array keeps an integer dtype if the caller passed ints — as the
    # reference does)
    obs = np.array(rbp_observed_values)
    for i in range(3):
        obs[i] = obs[i] * np.pi / 180
    w_obs = np.asarray(rbp_observation_weights, dtype=float)
    H = rbp_to_H(obs)

    idx_sel = np.arange(X_fix.shape[0])
    # simpleicp.py:158-170
    if np.isfinite(max_overlap_distance):
        X2 = transform_by_H(X2, H)
        idx_sel = select_in_range(X_fix, idx_sel, X2, max_overlap_distance)
        X2 = transform_by_H(X2, np.linalg.inv(H))
        if idx_sel.size == 0:
            raise OracleICPError(
                "Point clouds do not overlap within max_overlap_distance = "
                f"{max_overlap_distance:.5f}! Consider increasing the value of "
                "max_overlap_distance."
            )
        if trace is not None:
            trace.idx_overlap = idx_sel.copy()

    idx_sel = select_n_points(idx_sel, correspondences)  # simpleicp.py:172-174
    t1 = time.perf_counter()
    if normals is None:  # simpleicp.py:176-178
        nrm, plan, idxNN = estimate_normals(X_fix, idx_sel, neighbors)
    else:
        nrm, plan = normals
        idxNN = None
    t2 = time.perf_counter()
    if trace is not None:
        trace.idx_sel, trace.normals, trace.planarity, trace.idxNN = idx_sel, nrm, plan, idxNN

    tree_static = spatial.cKDTree(X2) if static_tree else None
    residuals_all: List[np.ndarray] = []
    x_est = None
    w = distance_weights
    res_w = jac = None
    n_corr = 0
    it = -1
    for it in range(max_iterations):  # simpleicp.py:184
        H_in = H
        t_it = time.perf_counter()
        if static_tree:
            q = transform_by_H(X_fix[idx_sel], np.linalg.inv(H))
            _, idx_nn = tree_static.query(q, k=1, p=2, workers=WORKERS)
            p2t = transform_by_H(X2[idx_nn], H)
            p1 = X_fix[idx_sel]
            n64 = nrm.astype(np.float64)
            dd = p2t - p1
            d = dd[:, 0] * n64[:, 0] + dd[:, 1] * n64[:, 1] + dd[:, 2] * n64[:, 2]
        else:
            X2 = transform_by_H(X2, H)  # :188
            idx_nn, d = match(X_fix, idx_sel, nrm, X2)  # :201
            X2 = transform_by_H(X2, np.linalg.inv(H))  # :202
        mov_plan = angle_ok = None
        if mov_normals is not None:
            mov_plan = mov_normals[1][idx_nn]
            if max_angle_between_normals is not None:
                angle_ok = angle_between_normals_ok(nrm, mov_normals[0][idx_nn], H, max_angle_between_normals)
        keep, med, mad = reject(d, plan, min_planarity, mov_plan, angle_ok)  # :205-206
        n_corr = int(keep.sum())
        if n_corr < 6:  # :209-214
            raise OracleICPError(
                "Too few correspondences! At least 6 correspondences are needed to estimate "
                "the 6 rigid body transformation parameters. The current number of "
                f"correspondences is {n_corr}."
            )
        x0 = obs if it == 0 else x_est  # :223-227
        if w is None:  # :233-234
            w = 1 / (np.std(d[keep]) ** 2)
        p1 = X_fix[idx_sel[keep]]
        p2 = X2[idx_nn[keep]]
        x_est, res, res_w, jac = estimate_parameters(p1, nrm[keep], p2, w, x0, obs, w_obs)
        H = rbp_to_H(x_est)
        residuals_all.append(res)
        if trace is not None:
            if not trace.light:
                trace.iterations.append(
                    IterationTrace(H_in, idx_nn, d, keep, med, mad, x_est.copy(), H.copy(), res, float(w))
                )
        stop = it > 0 and check_convergence_criteria(residuals_all[it], residuals_all[it - 1], min_change)
        if trace is not None:
            trace.iter_seconds.append(time.perf_counter() - t_it)
        if stop:
            break
    t3 = time.perf_counter()
    sigma = estimate_parameter_uncertainties(res_w, jac, n_corr, w, w_obs)
    X2 = transform_by_H(X2, H)  # :316
    t4 = time.perf_counter()
    if trace is not None:
        trace.sigma = sigma
        trace.timings = {
            "select": t1 - t0,
            "normals": t2 - t1,
            "loop": t3 - t2,
            "final": t4 - t3,
            "total": t4 - t0,
            "iterations": it + 1,
        }
    return H, X2, x_est, sigma, residuals_all[it]


# --------------------------------------------------------------------------- synthetic inputs
def surface(n: int, seed: int, extent: float = 100.0) -> np.ndarray:
    """SURVEY.md §8(d) C3 generator: tilted, gently undulating plane with 1 cm noise."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, extent, n)
    y = rng.uniform(0, extent, n)
    z = 0.05 * x + 0.03 * y + 2 * np.sin(2 * np.pi * x / 25) * np.cos(2 * np.pi * y / 40)
    z = z + rng.normal(0, 0.01, n)
    return np.column_stack((x, y, z))


def c3_pair(n: int = 1_000_000) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """C3: X_fix = surface(n, 1234); X_mov = H_true^-1 * surface(n, 5678)."""
    H_true = rbp_to_H(
        [np.deg2rad(0.3), np.deg2rad(-0.2), np.deg2rad(0.5), 0.15, -0.10, 0.05]
    )
    X_fix = surface(n, 1234)
    X_mov = transform_by_H(surface(n, 5678), np.linalg.inv(H_true))
    return X_fix, X_mov, H_true
