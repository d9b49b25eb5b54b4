"""CPU tests of the linearised-variant oracle (oracle/linearized_oracle.py) against the only golden
vector the reference holds for it: the C++ screen output in README.md:141-160."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import linearized_oracle as lo  # noqa: E402

GOLD = ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def dragon():
    d = np.load(GOLD / "data_dragon.npz")
    return d["fix"] / d["scale"], d["mov"] / d["scale"]


def test_helpers():
    v = np.array([5.0, 1.0, 3.0, 2.0])
    assert lo.median_upper(v) == 3.0  # nth_element at size/2: upper middle, not the average
    assert lo.median_upper(np.array([2.0, 9.0, 4.0])) == 4.0
    assert lo.mad_upper(v) == 2.0  # |v - 3| = 2 2 0 1 -> sorted 0 1 2 2 -> index 2
    assert lo.sample_std(np.array([1.0, 2.0, 3.0, 4.0])) == pytest.approx(np.std([1, 2, 3, 4], ddof=1))
    assert lo.change(1.01, 1.0) == pytest.approx(1.0)
    assert lo.change(0.0, 0.0) == 0.0 and lo.change(1.0, 0.0) == np.inf
    idx = np.arange(10)
    # LinSpaced(5, 0, 9) = 0 2.25 4.5 6.75 9 -> C round(): 0 2 5 7 9 (rint would give 4 for 4.5)
    assert lo.select_n_points(idx, 5).tolist() == [0, 2, 5, 7, 9]
    assert lo.select_n_points(idx, 10) is idx


def test_readme_golden_table(dragon):
    gold = json.loads((GOLD / "cpp_readme_dragon.json").read_text())
    r = lo.simpleicp_linearized(*dragon, compose="post", rotation="small_angle")
    assert r.converged
    assert len(r.iterations) - 1 == len(gold["rows"])  # the converging iteration is not printed
    assert abs(r.orig.n_kept - gold["orig"][0]) <= 10
    assert abs(r.orig.std - gold["orig"][2]) < 0.01  # sign dependent through the mean
    for it, row in zip(r.iterations, gold["rows"]):
        assert abs(it.n_kept - row[1]) <= 10
        assert abs(it.std - row[3]) < 1e-3
    assert np.abs(r.H - np.array(gold["H"])).max() < 2e-3
    # the printed table has the reference's layout
    lines = lo.format_table(r).splitlines()
    assert lines[0] == "Iteration | correspondences | mean(residuals) |  std(residuals)"
    assert lines[1].startswith("   orig:0 |") and len(lines) == 2 + len(gold["rows"])


def test_current_sources_converge_to_the_python_result(dragon):
    ref = np.load(GOLD / "ref_dragon.npz")
    r = lo.simpleicp_linearized(*dragon, compose="pre")
    assert r.converged and np.array_equal(r.H, r.T)
    R = r.T[:3, :3]
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-13  # Euler product: a rotation, unlike the README's H
    assert np.linalg.norm(r.T - ref["H"]) < 1e-5  # noise-free pair: both variants find the same H
    rp = lo.simpleicp_linearized(*dragon, compose="post")
    assert np.array_equal(rp.T, r.T) and not np.array_equal(rp.H, r.H)
    # H * dH and dH * H are products of the same increments in opposite order
    Hpost, Hpre = np.eye(4), np.eye(4)
    for it in rp.iterations:
        dH = np.eye(4)
        dH[:3, :3] = lo.euler_angles_to_rotation_matrix(*it.x[:3])
        dH[:3, 3] = it.x[3:]
        Hpost, Hpre = Hpost @ dH, dH @ Hpre
    assert np.allclose(Hpost, rp.H, atol=1e-15) and np.allclose(Hpre, rp.T, atol=1e-15)
