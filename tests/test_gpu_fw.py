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
@pytest.mark.parametrize("m,n,seed,its", [(80, 200, 10, 400), (30, 1000, 3, 500), (13, 506, 0, 300), (24, 1001, 6, 300)])
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


@pytest.mark.parametrize("away,world", [(1, 2), (0, 3), (1, 4)])
def test_sharded_fw_building_blocks_on_one_gpu(acc, away, world):
    """The column-sharded Frank-Wolfe loop emulated on one GPU: `world` slabs processed one after the other, records
    and the chosen column exchanged by hand exactly as the collectives would.  F, slacks and the vertex sequence must
    be those of the single-slab device loop (and of the oracle): the decision does not depend on the sharding."""
    from accbpg_and_fw_b200 import _native as nat
    from accbpg_and_fw_b200.dopt_fw import NCTRL, C_IMAX, C_JMIN, C_STOP, C_NITER
    lib = nat.lib
    rt = acc.Runtime.get()
    m, n, its = 12, 203, 60
    rng = np.random.RandomState(5)
    V = rng.randn(m, n)
    x0 = np.ones(n) / n
    offs = acc.ColumnShard.partition(n, world)
    dev = "cuda"
    Vs = [torch.tensor(np.ascontiguousarray(V[:, offs[r]:offs[r + 1]]), device=dev) for r in range(world)]
    xs = [torch.tensor(x0[offs[r]:offs[r + 1]].copy(), device=dev) for r in range(world)]
    nl = [offs[r + 1] - offs[r] for r in range(world)]
    wss = [torch.zeros(lib.accbpg_fw_workspace_bytes(m, nl[r]), dtype=torch.uint8, device=dev) for r in range(world)]
    M = torch.zeros(m, m, dtype=torch.float64, device=dev)
    Mr = torch.empty(m, m, dtype=torch.float64, device=dev)
    for r in range(world):
        nat.check(lib.accbpg_dopt_gram(rt.ctx, rt.stream, Vs[r].data_ptr(), m, nl[r], nl[r], xs[r].data_ptr(),
                                       wss[r].data_ptr(), Mr.data_ptr()))
        M += Mr
    Hinv = [torch.empty(m, m, dtype=torch.float64, device=dev) for _ in range(world)]
    ws_ = [torch.empty(nl[r], dtype=torch.float64, device=dev) for r in range(world)]
    ctrl = [torch.zeros(NCTRL, dtype=torch.float64, device=dev) for _ in range(world)]
    hist = [torch.zeros(4, its, dtype=torch.float64, device=dev) for _ in range(world)]
    rd = lib.accbpg_fw_record_bytes() // 8
    recs = torch.zeros(world * rd, dtype=torch.float64, device=dev)
    vcols = [torch.zeros(m, dtype=torch.float64, device=dev) for _ in range(world)]
    for r in range(world):
        nat.check(lib.accbpg_fw_setup_from_gram(rt.ctx, rt.stream, Vs[r].data_ptr(), m, nl[r], nl[r], M.data_ptr(),
                                                wss[r].data_ptr(), Hinv[r].data_ptr(), ws_[r].data_ptr(), ctrl[r].data_ptr()))
        nat.check(lib.accbpg_fw_select_local(rt.ctx, rt.stream, nl[r], offs[r], away, xs[r].data_ptr(), ws_[r].data_ptr(),
                                             wss[r].data_ptr(), m, recs[r * rd:].data_ptr()))
    log = []
    for k in range(its):
        for r in range(world):
            h = hist[r]
            nat.check(lib.accbpg_fw_decide(rt.ctx, rt.stream, Vs[r].data_ptr(), m, nl[r], nl[r], offs[r], away, 1e-9, k,
                                           recs.data_ptr(), world, wss[r].data_ptr(), ctrl[r].data_ptr(), h[0].data_ptr(),
                                           h[1].data_ptr(), h[2].data_ptr(), h[3].data_ptr(), vcols[r].data_ptr()))
        vsum = sum(vcols)                                   # the all-reduce
        c0 = ctrl[0].cpu().numpy()
        if c0[C_STOP] != 0:
            break
        log.append((int(c0[C_IMAX]), int(c0[C_JMIN])))
        new = torch.zeros_like(recs)
        for r in range(world):
            vr = vsum.clone()
            nat.check(lib.accbpg_fw_step(rt.ctx, rt.stream, Vs[r].data_ptr(), m, nl[r], nl[r], offs[r], away, k,
                                         wss[r].data_ptr(), Hinv[r].data_ptr(), vr.data_ptr(), xs[r].data_ptr(),
                                         ws_[r].data_ptr(), ctrl[r].data_ptr(), new[r * rd:].data_ptr()))
        torch.cuda.synchronize()
        recs = new                                          # the all-gather
    done = int(ctrl[0].cpu().numpy()[C_NITER])
    F = hist[0][0, :done].cpu().numpy()
    fn = orc.D_opt_FW_away if away else orc.D_opt_FW
    olog = []
    xo, Fo, SPo, SNo, To = fn(V, x0, 1e-9, its, index_log=olog)
    nn = min(len(F), len(Fo))
    assert nn >= its - 1
    assert np.max(np.abs(F[:nn] - Fo[:nn]) / np.maximum(np.abs(Fo[:nn]), 1e-3)) <= 1e-9
    assert [(a, b) for a, b in log[:nn - 1]] == [(a, b) for a, b, *_ in olog[:nn - 1]]
    xfull = np.concatenate([t.cpu().numpy() for t in xs])
    assert np.max(np.abs(xfull - xo)) <= 1e-9
    for r in range(1, world):                               # replicated state stays identical
        assert torch.equal(ctrl[r][:15], ctrl[0][:15]) and torch.equal(Hinv[r], Hinv[0])


_FW_PATHS_SCRIPT = r"""
import hashlib, sys, numpy as np
sys.path.insert(0, {root!r})
import accbpg_and_fw_b200 as acc
out = []
for (m, n, seed, eps, its, batch) in [(200, 5000, 2, 1e-8, 200, 64), (64, 3000, 5, 1e-2, 4000, 32), (10, 40, 4, 1e-3, 5000, 16)]:
    np.random.seed(seed)
    V = np.random.randn(m, n)
    x0 = np.ones(n) / n
    for fn in (acc.D_opt_FW_away, acc.D_opt_FW):
        x, F, SP, SN, T = fn(V, x0, eps, its, verbose=False, batch=batch)
        h = hashlib.sha256()
        for a in (x, F, SP, SN):
            h.update(np.ascontiguousarray(a).tobytes())
        out.append(f"{{len(F)}}:{{h.hexdigest()[:16]}}")
print("FWHASH " + " ".join(out))
"""


def test_fw_paths_bit_identical():
    """The three forms of the loop - ring-fed persistent launch (default on one GPU), launch chain with the register-fed
    pass, launch chain with the ring-fed pass - give the same bits (iterates, histories, stopping iteration).  The
    switches are read once per process, hence the subprocesses."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for name, env in (("persistent ring", {}), ("chain", {"ACCBPG_FW_PERSIST": "0"}),
                      ("chain + ring pass", {"ACCBPG_FW_PERSIST": "0", "ACCBPG_FW_RING": "1"}),
                      ("chain, nothing before the wait", {"ACCBPG_FW_PERSIST": "0", "ACCBPG_FW_EARLY": "0"})):
        e = dict(os.environ)
        for k in ("ACCBPG_FW_PERSIST", "ACCBPG_FW_RING", "ACCBPG_FW_EARLY"):
            e.pop(k, None)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", _FW_PATHS_SCRIPT.format(root=root)], env=e, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, (name, r.stderr[-2000:])
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("FWHASH ")]
        assert line, (name, r.stdout[-500:])
        res[name] = line[-1]
    assert len(set(res.values())) == 1, res
