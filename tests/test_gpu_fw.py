"""GPU parity: D_opt_FW / D_opt_FW_away (device-resident loop) against the oracle and the golden runs.
F_k, the slacks SP/SN and the selected vertex indices; the FW vertex index must be bit-exact."""
import numpy as np
import torch
import pytest

from conftest import relerr
from oracle import accbpg_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acc():
    import accbpg_and_fw_b200 as a
    return a


def close(a, b, tol, floor=1e-3):
    assert a.shape == b.shape, (a.shape, b.shape)
    err = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
    assert err <= tol, err


def test_fw_golden(acc, golden_traj):
    t = golden_traj
    f, h, L, x0 = acc.D_opt_design(80, 200, randseed=10)
    x, F, SP, SN, T = acc.D_opt_FW(f.H, x0, 1e-8, 2000, verbose=False)
    close(F, t["dfw_F"], 1e-9); close(SP, t["dfw_SP"], 1e-7); close(SN, t["dfw_SN"], 1e-7)
    close(x, t["dfw_x"], 1e-6, floor=1e-6)
    assert T.shape == F.shape and np.all(np.diff(T) >= 0)


def test_fw_away_golden(acc, golden_traj):
    t = golden_traj
    f, h, L, x0 = acc.D_opt_design(80, 200, randseed=10)
    x, F, SP, SN, T = acc.D_opt_FW_away(f.H, x0, 1e-8, 2000, verbose=False)
    close(F, t["dfwa_F"], 1e-9); close(SP, t["dfwa_SP"], 1e-7); close(SN, t["dfwa_SN"], 1e-7)
    close(x, t["dfwa_x"], 1e-6, floor=1e-6)
    x, F, SP, SN, T = acc.D_opt_FW_away(f.H, t["ky_x0"], 1e-8, 1000, verbose=False)
    close(F, t["dfwa_ky_F"], 1e-9); close(SP, t["dfwa_ky_SP"], 1e-7)


@pytest.mark.parametrize("away", [0, 1])
@pytest.mark.parametrize("m,n,seed,its", [(80, 200, 10, 400), (30, 1000, 3, 500), (13, 506, 0, 300)])
def test_fw_vertex_indices_bit_exact(acc, golden_ops, away, m, n, seed, its):
    if seed == 0:
        V = golden_ops["housing_H"]
    else:
        np.random.seed(seed)
        V = np.random.randn(m, n)
    x0 = np.ones(n) / n
    log_o, log_g = [], []
    fo = orc.D_opt_FW_away if away else orc.D_opt_FW
    fg = acc.D_opt_FW_away if away else acc.D_opt_FW
    xo, Fo, SPo, SNo, To = fo(V, x0, 1e-9, its, index_log=log_o)
    xg, Fg, SPg, SNg, Tg = fg(V, x0, 1e-9, its, verbose=False, index_log=log_g)
    assert len(log_g) == len(log_o)
    assert [p[0] for p in log_g] == [p[0] for p in log_o]          # argmax vertex i
    assert [p[1] for p in log_g] == [p[1] for p in log_o]          # (masked) argmin vertex j
    if away:
        assert [p[2] for p in log_g[:-1]] == [int(p[2]) for p in log_o[:-1]]   # toward / away decisions
    close(Fg, Fo, 1e-9)
    close(xg, xo, 1e-6, floor=1e-6)


def test_fw_stops_like_reference(acc):
    np.random.seed(4)
    V = np.random.randn(10, 40)
    x0 = np.ones(40) / 40
    a = acc.D_opt_FW_away(V, x0, 1e-3, 5000, verbose=False, batch=16)
    b = orc.D_opt_FW_away(V, x0, 1e-3, 5000)
    assert a[1].shape == b[1].shape and a[1].shape[0] < 5000
    close(a[1], b[1], 1e-9)
    a = acc.D_opt_FW(V, x0, 1e-3, 5000, verbose=False, batch=7)
    b = orc.D_opt_FW(V, x0, 1e-3, 5000)
    assert a[1].shape == b[1].shape
    close(a[1], b[1], 1e-9)


def test_fw_midsize_vs_oracle(acc):
    np.random.seed(2)
    V = np.random.randn(200, 5000)
    x0 = np.ones(5000) / 5000
    a = acc.D_opt_FW_away(V, x0, 1e-8, 300, verbose=False)
    b = orc.D_opt_FW_away(V, x0, 1e-8, 300)
    close(a[1], b[1], 1e-9); close(a[2], b[2], 1e-7); close(a[3], b[3], 1e-7)


def test_ky_init_golden(acc, golden_traj):
    """D_opt_KYinit (applications.py:59-95) with the q^T V passes and the arg-extrema on the device: same support as
    the reference (x0 is a function of the chosen indices only), from a host matrix and from a CUDA tensor."""
    np.random.seed(10)
    H = np.random.randn(80, 200)
    np.random.seed(77)
    assert np.array_equal(acc.D_opt_KYinit(H), golden_traj["ky_x0"])
    np.random.seed(77)
    assert np.array_equal(acc.D_opt_KYinit(torch.tensor(H, device="cuda")), golden_traj["ky_x0"])
    assert np.array_equal(acc.D_opt_KYinit(H[:, :150]), np.ones(150) / 150)       # n <= 2m: uniform
    xa, Fa, SPa, SNa, Ta = acc.D_opt_FW_away(H, golden_traj["ky_x0"], 1e-8, 1000, verbose=False)
    n = min(len(Fa), len(golden_traj["dfwa_ky_F"]))
    assert np.max(np.abs(Fa[:n] - golden_traj["dfwa_ky_F"][:n]) / np.abs(golden_traj["dfwa_ky_F"][:n])) <= 1e-9
