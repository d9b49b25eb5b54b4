"""CPU: the oracle restatement against the golden vectors captured from the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import load_golden, load_pair
from oracle import simpleicp_oracle as O


@pytest.mark.parametrize("name", ["dragon", "bunny", "multisensor", "webots", "dragon_observed"])
def test_oracle_reproduces_reference(name):
    g = load_golden(name)
    X_fix, X_mov = load_pair(name)
    tr = O.Trace()
    H, X_t, x, sigma, res = O.simpleicp(X_fix, X_mov, trace=tr, **g["kwargs"])
    n = len(tr.iterations)
    assert n == g["it_x"].shape[0]
    assert np.array_equal(tr.idx_sel, g["idx_sel"])
    if "idx_overlap" in g:
        assert np.array_equal(tr.idx_overlap, g["idx_overlap"])
    assert np.array_equal(tr.normals, g["normals"])
    assert np.array_equal(tr.planarity, g["planarity"], equal_nan=True)
    for i in range(n):
        assert np.array_equal(tr.iterations[i].pc2_idx, g["it_pc2_idx"][i])
        assert np.array_equal(tr.iterations[i].keep, g["it_keep"][i])
        np.testing.assert_allclose(tr.iterations[i].distances, g["it_dist"][i], rtol=0, atol=1e-15)
        np.testing.assert_allclose(tr.iterations[i].x, g["it_x"][i], rtol=0, atol=1e-14)
    np.testing.assert_allclose(H, g["H"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(res, g["residuals"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(sigma, g["sigma"], rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(X_t[:64], g["X_mov_t_head"], rtol=0, atol=1e-12)


def test_oracle_movable_side_planarity():
    """pc_mov with normal columns: the second branch of CorrPts.reject_wrt_planarity
    (corrpts.py:157-162), captured from the unmodified reference (make_golden.py config
    dragon_movnormals).  Also pins estimate_normals on all 100 000 points of a cloud."""
    g = load_golden("dragon_movnormals")
    X_fix, X_mov = load_pair("dragon_movnormals")
    tr = O.Trace()
    mov = (g["mov_normals"], g["mov_planarity"])
    H, X_t, x, sigma, res = O.simpleicp(X_fix, X_mov, trace=tr, mov_normals=mov, **g["kwargs"])
    n = len(tr.iterations)
    assert n == g["it_x"].shape[0]
    for i in range(n):
        assert np.array_equal(tr.iterations[i].pc2_idx, g["it_pc2_idx"][i])
        assert np.array_equal(tr.iterations[i].keep, g["it_keep"][i])
        np.testing.assert_allclose(tr.iterations[i].x, g["it_x"][i], rtol=0, atol=1e-14)
    np.testing.assert_allclose(H, g["H"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(res, g["residuals"], rtol=0, atol=1e-14)
    # the filter did something: fewer correspondences than the plain dragon run keeps
    assert g["it_keep"][0].sum() < load_golden("dragon")["it_keep"][0].sum()
    # the restated estimate_normals reproduces the movable-side columns bit for bit
    sub = np.arange(0, len(X_mov), 37)
    nrm, plan, _ = O.estimate_normals(X_mov, sub, 10)
    assert np.array_equal(nrm, g["mov_normals"][sub])
    assert np.array_equal(plan, g["mov_planarity"][sub], equal_nan=True)


def test_oracle_angle_rejection_extension():
    """The angle test is an extension (the reference raises NotImplementedError): check its
    definition on hand-made data, and that 90 degrees accepts everything."""
    n1 = np.array([[0, 0, 1], [0, 0, 1], [0, 0, 1], [0, 0, 1]], dtype=np.float32)
    n2 = np.array([[0, 0, -1], [0, np.sin(0.2), np.cos(0.2)], [1, 0, 0], [np.nan, 0, 0]], dtype=np.float32)
    ok = O.angle_between_normals_ok(n1, n2, np.eye(4), 15.0)
    assert ok.tolist() == [True, True, False, False]
    Rz = np.eye(4)
    Rz[:3, :3] = O.euler_angles_to_rotation_matrix(0.3, 0.0, 0.0)  # tilts n2 by 0.3 rad about x
    assert O.angle_between_normals_ok(n1[:1], n2[:1], Rz, 15.0).tolist() == [False]
    assert O.angle_between_normals_ok(n1[:3], n2[:3], np.eye(4), 90.0).all()
    d = np.array([0.0, 0.1, -0.1, 0.05, 5.0])
    pl = np.full(5, 0.9, dtype=np.float32)
    keep, med, mad = O.reject(d, pl, 0.3, None, np.array([True, False, True, True, True]))
    assert keep.tolist() == [True, False, True, True, False] and med == 0.05  # median/MAD still see entry 1


def test_known_answer_readme_bunny():
    """python/README.md:62-65 prints the bunny H to 6 decimals (older release)."""
    g = load_golden("bunny")
    H_readme = np.array(
        [[0.984798, -0.173702, -0.000053, 0.000676],
         [0.173702, 0.984798, 0.000084, -0.001150],
         [0.000038, -0.000092, 1.000000, 0.000113],
         [0.0, 0.0, 0.0, 1.0]])
    assert np.abs(g["H"] - H_readme).max() < 1e-6


def test_static_tree_equivalence():
    """Moving the K queries by inv(H) into a static tree (the product's strategy) equals the
    reference's transform-the-cloud-and-rebuild strategy."""
    X_fix, X_mov = load_pair("dragon")
    H1, *_ = O.simpleicp(X_fix, X_mov)
    H2, *_ = O.simpleicp(X_fix, X_mov, static_tree=True)
    assert np.linalg.norm(H1 - H2) < 1e-9


def test_oracle_errors():
    X_fix, X_mov = load_pair("bunny")
    with pytest.raises(O.OracleICPError, match="do not overlap"):
        O.simpleicp(X_fix, X_mov + 1000.0, max_overlap_distance=0.5)
    with pytest.raises(O.OracleICPError, match="distance_weights"):
        O.simpleicp(X_fix, X_mov, distance_weights=0)
    with pytest.raises(O.OracleICPError, match="finite"):
        O.simpleicp(X_fix, X_mov, rbp_observation_weights=(np.inf,) * 6)


@pytest.mark.parametrize("name", ["airborne", "terrestrial"])
def test_oracle_large_lidar(name):
    """The two 1.3 M-point lidar pairs of the reference's test file
    (python/simpleicp/tests/test_simpleicp.py:44-63), inputs from the compressed fixtures:
    the restatement reproduces every stage of the unmodified reference bit for bit."""
    g = load_golden(name)
    X_fix, X_mov = load_pair(name)
    assert X_fix.shape[0] == int(g["n_fix"]) and X_mov.shape[0] == int(g["n_mov"])
    tr = O.Trace()
    H, X_t, x, sigma, res = O.simpleicp(X_fix, X_mov, trace=tr, **g["kwargs"])
    assert len(tr.iterations) == g["it_x"].shape[0]
    assert np.array_equal(tr.idx_sel, g["idx_sel"])
    assert np.array_equal(tr.normals, g["normals"])
    for i, it in enumerate(tr.iterations):
        assert np.array_equal(it.pc2_idx, g["it_pc2_idx"][i])
        assert np.array_equal(it.keep, g["it_keep"][i])
    np.testing.assert_allclose(H, g["H"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(res, g["residuals"], rtol=0, atol=1e-13)


def test_reference_result_depends_on_the_eigenvector_sign_convention():
    """Why the stand-alone GPU tolerances differ per data set (tests/test_gpu_parity.py::
    test_full_run_standalone): the reference takes each normal's SIGN from np.linalg.eig, which
    leaves it to LAPACK.  Flipping ALL normals changes nothing (d -> -d, median -> -median), but
    flipping a FEW -- an equally valid eigen-decomposition, and what any other LAPACK build or
    k-NN tie order may produce -- moves the median the rejection is centred on.  On dragon
    (noise-free, 1000 correspondences) that moves H by < 1e-6; on multisensor (a 316-point radar cloud)
    1 % of the signs moves the reference's OWN H by 1e-3 .. 1e-2.  A parity bar tighter than that
    on those inputs would test the sign convention, not the algorithm; with the reference's signs
    handed over every data set reproduces to 3e-10 (test_full_run_with_reference_normals)."""
    moved = {}
    for name, frac in (("dragon", 0.01), ("multisensor", 0.01), ("webots", 0.05)):
        g = load_golden(name)
        X_fix, X_mov = load_pair(name)
        nrm, pl = g["normals"].astype(np.float32), g["planarity"].astype(np.float32)
        rng = np.random.default_rng(0)
        out = []
        for _ in range(3):
            n2 = nrm.copy()
            n2[rng.random(len(n2)) < frac] *= -1
            out.append(np.linalg.norm(O.simpleicp(X_fix, X_mov, normals=(n2, pl), **g["kwargs"])[0] - g["H"]))
        moved[name] = out
        flipped = O.simpleicp(X_fix, X_mov, normals=(-nrm, pl), **g["kwargs"])[0]
        assert np.linalg.norm(flipped - g["H"]) < 1e-12  # global sign: no effect at all
    print({k: ["%.1e" % v for v in vs] for k, vs in moved.items()})
    assert max(moved["dragon"]) < 1e-6   # measured 4e-7: below the graded bar of 1e-5
    assert max(moved["multisensor"]) > 1e-3
    assert max(moved["webots"]) > 1e-4
