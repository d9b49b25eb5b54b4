"""The stand-in headers of oracle/cpp_standin (Eigen, nanoflann subsets the C++ reference is
compiled against here) checked on their own against NumPy / SciPy: symmetric eigen-decomposition,
SVD least squares (full rank and rank deficient), LinSpaced, exact k-NN with ties by index.
Needs g++ (skipped otherwise); independent of /root/reference."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest
from scipy import spatial

ROOT = Path(__file__).resolve().parent.parent
STANDIN = ROOT / "oracle" / "cpp_standin"

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def selftest(tmp_path_factory):
    exe = tmp_path_factory.mktemp("standin") / "selftest"
    subprocess.run(["g++", "-std=c++14", "-O2", "-ffp-contract=off", f"-I{STANDIN}", str(STANDIN / "selftest.cpp"),
                    "-o", str(exe)], check=True)

    def run(text):
        out = subprocess.run([str(exe)], input=text, capture_output=True, text=True, check=True).stdout
        return [np.array(line.split(), dtype=float) for line in out.strip().splitlines()]

    return run


def fmt(a):
    return " ".join(repr(float(v)) for v in np.asarray(a, dtype=float).ravel())


def test_self_adjoint_eigen_solver(selftest):
    rng = np.random.default_rng(0)
    cases = []
    for n in (3, 3, 3, 5):
        B = rng.standard_normal((n, n))
        cases.append(B @ B.T)
    P = rng.standard_normal((10, 3)) * [1.0, 1.0, 1e-4]      # a flat neighbourhood: tiny smallest eigenvalue
    cases.append(np.cov(P.T))
    cases.append(np.diag([2.0, 2.0, 5.0]))                    # repeated eigenvalue
    outs = selftest("\n".join(f"eig {len(A)} {fmt(A)}" for A in cases))
    for A, o in zip(cases, outs):
        n = len(A)
        w, V = o[:n], o[n:].reshape(n, n).T                   # printed column by column
        w_ref = np.linalg.eigvalsh(A)
        assert np.all(np.diff(w) >= 0)
        assert np.abs(w - w_ref).max() <= 1e-13 * max(1.0, np.abs(w_ref).max())
        assert np.abs(V.T @ V - np.eye(n)).max() < 1e-13      # orthonormal
        assert np.abs(A @ V - V * w).max() <= 1e-12 * max(1.0, np.abs(w_ref).max())


def test_svd_least_squares(selftest):
    rng = np.random.default_rng(1)
    A1, b1 = rng.standard_normal((200, 6)), rng.standard_normal(200)
    A2 = rng.standard_normal((50, 6)) * [1, 1e-3, 1e3, 1, 1e-6, 1]   # badly scaled columns
    b2 = rng.standard_normal(50)
    A3 = rng.standard_normal((40, 4))
    A3[:, 3] = A3[:, 0] - 2 * A3[:, 1]                               # rank 3: minimum-norm solution
    b3 = rng.standard_normal(40)
    outs = selftest("\n".join(f"lsq {A.shape[0]} {A.shape[1]} {fmt(A)} {fmt(b)}" for A, b in ((A1, b1), (A2, b2), (A3, b3))))
    for (A, b), x in zip(((A1, b1), (A2, b2), (A3, b3)), outs):
        x_ref = np.linalg.lstsq(A, b, rcond=None)[0]
        assert np.abs(x - x_ref).max() <= 1e-9 * max(1.0, np.abs(x_ref).max())
    assert np.abs(outs[0] - np.linalg.lstsq(A1, b1, rcond=None)[0]).max() < 1e-13


def test_linspaced_is_numpy_linspace(selftest):
    cases = [(5, 0.0, 9.0), (1000, 0.0, 99999.0), (7, -3.0, 2.0), (4, 5.0, -1.0), (1, 2.0, 2.0), (100000, 0.0, 1342905.0)]
    outs = selftest("\n".join(f"lin {n} {lo!r} {hi!r}" for n, lo, hi in cases))
    for (n, lo, hi), v in zip(cases, outs):
        ref = np.linspace(lo, hi, n)
        if abs(hi) >= abs(lo):  # low + i * step, the reference's case (low = 0): bit-identical to NumPy
            assert np.array_equal(v, ref)
        else:                   # mirrored evaluation (Eigen's rule when |high| < |low|)
            assert np.abs(v - ref).max() < 1e-14


def test_knn_is_exact_with_ties_by_index(selftest):
    rng = np.random.default_rng(2)
    X = np.round(rng.random((500, 3)) * 10) / 10          # a 0.1 lattice: many exactly equidistant points
    qs = [X[17], np.array([0.55, 0.55, 0.55]), np.array([20.0, -3.0, 0.5])]
    k = 12
    outs = selftest("\n".join(f"knn {len(X)} {k} {fmt(X)} {fmt(q)}" for q in qs))
    for q, o in zip(qs, outs):
        idx, d2 = o[:k].astype(int), o[k:]
        d2_all = ((X - q) ** 2).sum(axis=1)
        order = np.lexsort((np.arange(len(X)), d2_all))   # by distance, then by index
        assert np.array_equal(idx, order[:k])
        assert np.allclose(d2, d2_all[idx], rtol=0, atol=1e-15)
        dk, _ = spatial.cKDTree(X).query(q, k=k)
        assert np.allclose(np.sqrt(d2), dk, atol=1e-12)
