import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_ops():
    return dict(np.load(os.path.join(GOLDEN, "operators.npz")))


@pytest.fixture(scope="session")
def golden_traj():
    return dict(np.load(os.path.join(GOLDEN, "trajectories.npz")))


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def first_fork(D, Dref):
    """Index of the first differing discrete line-search outcome (L_k, gain G_k, gamma_k), or the common length."""
    D, Dref = np.asarray(D), np.asarray(Dref)
    n = min(len(D), len(Dref))
    idx = np.nonzero(D[:n] != Dref[:n])[0]
    return int(idx[0]) if len(idx) else n


def assert_trajectory(F, Fref, D, Dref, tol=1e-9, min_prefix=None, floor=1e-3):
    """Fork-aware trajectory parity.

    The drivers take discrete decisions (line-search comparisons) on differences of nearly equal FP64
    numbers.  A 1-ulp change in any operator output can flip one of them, after which two *correct*
    implementations follow different, equally valid trajectories; tests/test_noise_floor.py shows the
    reference itself does this under 1-ulp perturbations.  So: F must agree to `tol` on the whole prefix
    on which every discrete decision agrees (F[k] is evaluated before decision k is taken), and that
    prefix must be at least `min_prefix` long (default: the whole run).
    """
    F, Fref = np.asarray(F), np.asarray(Fref)
    k = first_fork(D, Dref)
    full = min(len(F), len(Fref))
    if min_prefix is None:
        min_prefix = full
    assert k >= min(min_prefix, full), f"discrete decisions fork at iteration {k} (< {min_prefix})"
    upto = min(k + 1, full)
    err = float(np.max(np.abs(F[:upto] - Fref[:upto]) / np.maximum(np.abs(Fref[:upto]), floor)))
    assert err <= tol, f"F differs by {err:.3e} on the agreed prefix [0, {upto})"
    return k
