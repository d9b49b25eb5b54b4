import ast
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
GOLD = REPO / "tests" / "golden"
REFERENCE = Path("/root/reference")

TRAVEL = ("dragon", "bunny", "multisensor", "webots")
ALIAS = {"dragon_observed": "dragon"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pair(name):
    """Input clouds of a reference test config, stored losslessly as scaled int32."""
    z = np.load(GOLD / f"data_{ALIAS.get(name, name)}.npz")
    s = float(z["scale"])
    return z["fix"] / s, z["mov"] / s


def load_golden(name):
    g = dict(np.load(GOLD / f"ref_{name}.npz"))
    inf = np.inf  # noqa: F841  (used by eval below)
    g["kwargs"] = eval(str(g["kwargs_repr"]), {"inf": np.inf, "nan": np.nan})
    return g


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not has_gpu():
        pytest.skip("no CUDA device")
    return True
