"""Linearised simpleICP variant on the GPU: the algorithm of the C++ driver (and, up to two
second-order details listed in include/sicp_b200.h next to ``sicp_variant``, of its Rust / Julia /
MATLAB siblings, whose composition order ``dH * H`` is the default here).

Mirror of the C++ entry point ``SimpleICP(X_fix, X_mov, correspondences, neighbors, min_planarity,
max_overlap_distance, min_change, max_iterations)`` (/root/reference/c++/src/simpleicp.h,
simpleicp.cpp:8-129): same arguments, same defaults as its command line
(c++/src/simpleicp-cli.cpp:15-35: ``max_overlap_distance = -1`` means "no overlap filter"), same
screen output, returns the 4 x 4 matrix H.  What the variant changes relative to the Python
package's algorithm is listed next to ``sicp_variant`` in include/sicp_b200.h.

The whole loop runs in the same two kernels per iteration as the default variant
(``sicp_set_option(ctx, "variant", ...)``); nothing here computes on the CPU beyond the index
arithmetic of the subsample.

Known, documented deviations from the C++ sources (DESIGN.md "linearised variant"):
  * normals are stored as float32 (the default variant's storage); the C++ keeps float64.  The
    effect on H is ~1e-8 (measured in tests/test_gpu_linearized.py against a float64 CPU restatement);
  * the sign of a normal is this library's convention (LAPACK dgeev emulation), not Eigen's; the
    estimated transform does not depend on it, the sign of individual residuals (and therefore the
    printed residual mean) does;
  * the overlap filter keeps ``dist < max_overlap_distance`` (the Python rule); the C++ keeps
    ``<=`` (pointcloud.cpp:68-74).  They differ only for a distance exactly equal to the bound.
"""
from __future__ import annotations

import logging
import time
from typing import Optional, Tuple

import numpy as np

from . import _capi

_log = logging.getLogger(__name__)

VARIANT_LINEARIZED = 1       # C++ arithmetic, reports dH * H (composition order of rust/src/icp.rs:164, matlab/simpleicp.m:55)
VARIANT_LINEARIZED_CPP = 2   # reports H * dH  (C++ simpleicp.cpp:66)


class LinearizedResult:
    H: np.ndarray                 # reported matrix (composition rule of the chosen driver)
    T: np.ndarray                 # transform actually applied to the movable cloud
    X_mov_transformed: np.ndarray
    records: list
    iterations: int
    converged: bool
    idx_selected: np.ndarray
    loop_ms: float
    table: str
    normals: Optional[tuple]      # (nx, ny, nz, planarity) float32 over the selected points, on request


def subsample_indices_cpp(m: int, n: int) -> np.ndarray:
    """c++/src/pointcloud.cpp:78-98 -- LinSpaced(n, 0, m-1) rounded with C round() (half away
    from zero; the Python package uses rint).  Strictly increasing for n < m."""
    lin = np.linspace(0.0, float(m - 1), int(n))
    return np.floor(lin + 0.5).astype(np.int64)


def format_table(records, iterations: int, converged: bool) -> str:
    """The rows of c++/src/simpleicp.cpp:82-101.  The C++ driver breaks BEFORE printing the row of
    the iteration that met the stop rule."""
    rows = ["%9s | %15s | %15s | %15s" % ("Iteration", "correspondences", "mean(residuals)", "std(residuals)")]
    if records:
        r0 = records[0]
        rows.append("%9s | %15d | %15.4f | %15.4f" % ("orig:0", r0["n_kept"], r0["mean_dist"], r0["std_dist"]))
    shown = iterations - 1 if converged else iterations
    for k in range(shown):
        r = records[k]
        rows.append("%9d | %15d | %15.4f | %15.4f" % (k + 1, r["n_kept"], r["mean_res"], r["std_res"]))
    if converged:
        rows.append("Convergence criteria fulfilled -> stop iteration!")
    return "\n".join(rows)


def format_matrix(H: np.ndarray) -> str:
    """c++/src/simpleicp.cpp:103-123."""
    return "\n".join("[%12.6f %12.6f %12.6f %12.6f]" % tuple(H[i]) for i in range(4))


def simpleicp_linearized(
    X_fix,
    X_mov,
    correspondences: int = 1000,
    neighbors: int = 10,
    min_planarity: float = 0.3,
    max_overlap_distance: float = -1.0,
    min_change: float = 1.0,
    max_iterations: int = 100,
    *,
    compose: str = "dH*H",
    normals: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]] = None,
    engine: Optional[_capi.Engine] = None,
    transform_out=None,
    verbose: bool = False,
    want_normals: bool = False,
) -> LinearizedResult:
    """One registration with the linearised variant.  ``compose``: "dH*H" (Rust/MATLAB/Julia) or
    "H*dH" (the C++ driver's reported matrix).  ``normals = (nx, ny, nz, planarity)`` are arrays
    over the SELECTED fixed points, supplied instead of being estimated (tests)."""
    if compose not in ("dH*H", "H*dH"):
        raise ValueError('compose must be "dH*H" or "H*dH"')
    say = print if verbose else _log.info
    t0 = time.time()
    own = engine is None
    eng = engine or _capi.Engine()
    try:
        say("Create point cloud objects ...")
        eng.set_option("variant", VARIANT_LINEARIZED_CPP if compose == "H*dH" else VARIANT_LINEARIZED)
        zero = [0.0] * 6
        params = eng.run_params(min_planarity, min_change, max_iterations,
                                eng.lsq_params(zero, zero, zero, 1.0))
        has_filter = max_overlap_distance > 0 and np.isfinite(max_overlap_distance)
        X_t = None
        if normals is None and not want_normals:
            # the whole driver as one library call (sicp_register; the variant option selects the
            # native rounding of the subsample and the linearised loop)
            for msg in ("Consider partial overlap of point clouds ..." if has_filter else None,
                        "Select points for correspondences in fixed point cloud ...",
                        "Estimate normals of selected points ...", "Start iterations ..."):
                if msg:
                    say(msg)
            try:
                out, log, _, X_t = eng.register_fused(X_fix, X_mov, correspondences, neighbors,
                                                      max_overlap_distance if has_filter else -1.0, params,
                                                      transform_out)
            except _capi.SicpError as e:
                if e.code == _capi.SICP_ERR_NO_OVERLAP:
                    raise RuntimeError(
                        "Point clouds do not overlap within max_overlap_distance = %.5f. "
                        "Consider increasing the value of max_overlap_distance." % max_overlap_distance
                    ) from None
                raise
            nrm = None
            idx = eng.select_n_points(eng.K)  # downloads the selection (no-op on the device)
        else:
            eng.set_clouds(X_fix, X_mov)
            n_fix = eng.n_fix
            idx = None
            if has_filter:
                say("Consider partial overlap of point clouds ...")
                eng.set_selected(None)
                try:
                    keep = eng.select_in_range(np.eye(4), float(max_overlap_distance))
                except _capi.SicpError as e:
                    if e.code == _capi.SICP_ERR_NO_OVERLAP:
                        raise RuntimeError(
                            "Point clouds do not overlap within max_overlap_distance = %.5f. "
                            "Consider increasing the value of max_overlap_distance." % max_overlap_distance
                        ) from None
                    raise
                idx = np.arange(n_fix, dtype=np.int64)[keep]
            say("Select points for correspondences in fixed point cloud ...")
            m = n_fix if idx is None else idx.size
            if correspondences < m:
                pick = subsample_indices_cpp(m, correspondences)
                idx = (pick if idx is None else idx[pick]).astype(np.int64)
            eng.set_selected(idx)
            if idx is None:
                idx = np.arange(n_fix, dtype=np.int64)
            if normals is None:
                say("Estimate normals of selected points ...")
                nrm = eng.estimate_normals(neighbors, download=want_normals)
            else:
                nrm = tuple(np.asarray(a, dtype=np.float32) for a in normals)
                eng.set_normals(*nrm)
            say("Start iterations ...")
            out, log, _ = eng.run(params)
        records = [
            dict(n_kept=int(r.n_kept), median=r.median, mad=r.mad, mean_dist=r.mean_dist,
                 std_dist=r.std_dist, x=np.array(r.x), mean_res=r.mean_res, std_res=r.std_res,
                 n_bruteforce=int(r.n_bruteforce))
            for r in log
        ]
        res = LinearizedResult()
        res.H = np.array(out.H).reshape(4, 4)
        res.T = eng.get_transform()
        res.records, res.iterations, res.converged = records, int(out.iterations), bool(out.converged)
        res.idx_selected, res.loop_ms = idx, float(out.loop_ms)
        res.normals = nrm if (want_normals or normals is not None) else None
        res.table = format_table(records, res.iterations, res.converged)
        say(res.table)
        say("Estimated transformation matrix H:")
        say(format_matrix(res.H))
        res.X_mov_transformed = X_t if X_t is not None else eng.transform(res.T, out=transform_out)
        say("Finished in %.3f seconds!" % (time.time() - t0))
        return res
    finally:
        eng.set_option("variant", 0)
        if own:
            eng.close()
