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
