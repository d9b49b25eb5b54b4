"""CPU: .xyz reader / writer of the library (host C++) against NumPy's text parser."""
import numpy as np
import pytest

import simpleicp_b200 as sb
from conftest import load_pair


def test_read_matches_genfromtxt(tmp_path):
    X, _ = load_pair("bunny")
    f = tmp_path / "bunny.xyz"
    np.savetxt(f, X, fmt="%.4f")
    A = sb.read_xyz(f)
    B = np.genfromtxt(f)
    assert A.dtype == np.float64 and A.shape == X.shape
    assert np.array_equal(A, B) and np.array_equal(A, X)  # correctly rounded parse


def test_comments_blank_lines_formats(tmp_path):
    f = tmp_path / "mixed.xyz"
    f.write_text("//X Y Z\n# comment\n\n1 2 3\n  -1.5e-3\t+2.25   3e2  extra columns ignored\r\n7.0,8.0,9.0\n1e-320 0 0")
    A = sb.read_xyz(f)
    np.testing.assert_array_equal(A, [[1, 2, 3], [-1.5e-3, 2.25, 300.0], [7, 8, 9], [1e-320, 0, 0]])


def test_large_file_is_parsed_in_parallel_chunks(tmp_path):
    rng = np.random.default_rng(0)
    X = rng.normal(size=(300_000, 3)) * 1000
    f = tmp_path / "big.xyz"
    np.savetxt(f, X, fmt="%.17g")
    A = sb.read_xyz(f)
    assert np.array_equal(A, X)  # round-trip precision, order preserved across chunk borders


def test_write_like_reference(tmp_path):
    X = np.array([[1.23456, -2.0, 3.0005], [0.0004, 5.5, -6.25]])
    f = tmp_path / "out.xyz"
    sb.write_xyz(f, X)  # the reference's PointCloud.write_xyz: "//X Y Z" header, %.3f
    lines = f.read_text().splitlines()
    assert lines[0] == "//X Y Z" and lines[1] == "1.235 -2.000 3.001" and lines[2] == "0.000 5.500 -6.250"
    sb.write_xyz(f, X, decimals=-1, header=False)
    assert np.array_equal(sb.read_xyz(f), X)


def test_errors(tmp_path):
    with pytest.raises(OSError, match="cannot open"):
        sb.read_xyz(tmp_path / "missing.xyz")
    f = tmp_path / "bad.xyz"
    f.write_text("1 2 3\n4 5\n")
    with pytest.raises(OSError, match="malformed"):
        sb.read_xyz(f)
    with pytest.raises(ValueError):
        sb.write_xyz(tmp_path / "x.xyz", np.zeros((3, 2)))
