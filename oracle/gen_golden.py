"""
Generate tests/golden/*.npz by running the REAL reference (/root/reference/accbpg).

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no
/root/reference):      python oracle/gen_golden.py

The reference imports cvxpy, jax and matplotlib at module scope although the hot
path never touches them; they are absent here, so permissive stub modules are
registered first (recipe from SURVEY.md section 8c).  The reference tree is never
modified and none of its source is copied: only inputs (or the seeds that
regenerate them) and the reference's numerical outputs are stored.
"""
import os
import sys
import types
import contextlib
import io

import numpy as np

REF_ROOT = os.environ.get("ACCBPG_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference(ref_root=REF_ROOT):
    """Import the reference package with stubbed optional dependencies (oracle/ref_loader.py)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ref_loader
    return ref_loader.import_reference(ref_root)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.RandomState(12345)

    # ---------------- operator-level vectors -----------------------------
    ops = {}
    # D-opt on three shapes (H regenerated from the legacy seed by the tests)
    for tag, (m, n, seed) in {"dopt_80x200": (80, 200, 10), "dopt_30x1000": (30, 1000, 3),
                              "dopt_129x517": (129, 517, 7)}.items():
        f, h, L, x0 = ref.D_opt_design(m, n, randseed=seed)
        x = rng.rand(n)
        x /= x.sum()
        fx, g = f.func_grad(x)
        ops[tag + "_x"] = x
        ops[tag + "_f"] = fx
        ops[tag + "_g"] = g
        ops[tag + "_f0"] = f(x0)
        ops[tag + "_Hsum"] = f.H.sum()
    # housing LIBSVM instance
    f, h, L, x0 = ref.D_opt_libsvm(os.path.join(REF_ROOT, "parameters_free_fw/data/housing.txt"))
    ops["housing_H"] = f.H
    ops["housing_f0"] = f(x0)
    ops["housing_g0"] = f.gradient(x0)

    # Poisson / KL oracles
    f, h, L, x0 = ref.Poisson_regrL1(200, 100, noise=1e-4, lamda=0, randseed=1)
    ops["poisson_A"], ops["poisson_b"], ops["poisson_L"] = f.A, f.b, L
    x = 0.05 + rng.rand(100)
    ops["poisson_x"] = x
    ops["poisson_f"], ops["poisson_g"] = f.func_grad(x)
    ops["poisson_f0"] = f(x0)
    f, h, L, x0 = ref.KL_nonneg_regr(300, 120, noise=0.01, lamdaL1=0.001, randseed=1)
    ops["kl_A"], ops["kl_b"], ops["kl_L"] = f.A, f.b, L
    x = 0.05 + rng.rand(120)
    ops["kl_x"] = x
    ops["kl_f"], ops["kl_g"] = f.func_grad(x)
    f1000, h1000, L1000, x01000 = ref.KL_nonneg_regr(1000, 100, noise=0.01, lamdaL1=0.001, randseed=1)
    ops["kl1000_F0"] = f1000(x01000) + h1000.extra_Psi(x01000)

    # Bregman kernels on a shared set of vectors (n = 1000)
    n = 1000
    xv = 0.01 + rng.rand(n)
    yv = 0.01 + rng.rand(n)
    gv = rng.randn(n)
    gpos = 0.1 + rng.rand(n)
    ops["vec_x"], ops["vec_y"], ops["vec_g"], ops["vec_gpos"] = xv, yv, gv, gpos
    Lc = 0.7
    burg = ref.BurgEntropy()
    ops["burg_h"] = burg(xv)
    ops["burg_grad"] = burg.gradient(xv)
    ops["burg_div"] = burg.divergence(xv, yv)
    ops["burg_prox"] = burg.prox_map(gpos, Lc)
    ops["burg_divprox"] = burg.div_prox_map(yv, gpos, Lc)
    b1 = ref.BurgEntropyL1(lamda=0.3)
    ops["burgl1_prox"] = b1.prox_map(gpos, Lc)
    ops["burgl1_divprox"] = b1.div_prox_map(yv, gv * 0.1, Lc)
    ops["burgl1_psi"] = b1.extra_Psi(xv)
    b2 = ref.BurgEntropyL2(lamda=0.3)
    ops["burgl2_prox"] = b2.prox_map(gv, Lc)
    ops["burgl2_divprox"] = b2.div_prox_map(yv, gv, Lc)
    ops["burgl2_psi"] = b2.extra_Psi(xv)
    bs = ref.BurgEntropySimplex()
    ops["burgs_prox"] = bs.prox_map(gv, Lc)
    ys = yv / yv.sum()
    ops["vec_ysimplex"] = ys
    ops["burgs_divprox"] = bs.div_prox_map(ys, gv, Lc)
    sh = ref.ShannonEntropy()
    ops["sh_h"] = sh(xv)
    ops["sh_grad"] = sh.gradient(xv)
    ops["sh_div"] = sh.divergence(xv, yv)
    ops["sh_prox"] = sh.prox_map(gv, Lc)
    ops["sh_divprox"] = sh.div_prox_map(yv, gv, Lc)
    s1 = ref.ShannonEntropyL1(lamda=0.3)
    ops["shl1_prox"] = s1.prox_map(gv, Lc)
    ops["shl1_divprox"] = s1.div_prox_map(yv, gv, Lc)
    ss = ref.ShannonEntropySimplex()
    ops["shs_prox"] = ss.prox_map(gv, Lc)
    ops["shs_divprox"] = ss.div_prox_map(ys, gv, Lc)
    # LMOs
    gt = np.round(gv, 1)
    gt[[640, 137, 903]] = gt.min() - 0.5      # exact three-way tie at the minimum: index 137 must win
    ops["lmo_gties"] = gt
    ops["lmo_simplex"] = ref.lmo_simplex(2.0)(gt)
    ops["lmo_l2"] = ref.lmo_l2_ball(1.5)(gv)
    ops["lmo_l2_center"] = ref.lmo_l2_ball(1.5, center=xv)(gv)
    ops["lmo_l2pos"] = ref.lmo_l2_ball_positive_orthant(1.5, center=xv, epsilon=1e-3)(gv)
    ops["lmo_linf"] = ref.lmo_linf_ball(0.5, center=xv)(gv)
    Gm = gv.reshape(25, 40)
    ops["lmo_msimplex"] = ref.lmo_matrix_simplex(3.0)(Gm)
    ops["lmo_mbox"] = ref.lmo_matrix_box(-np.ones((25, 40)), 2 * np.ones((25, 40)))(Gm)
    np.savez_compressed(os.path.join(OUT, "operators.npz"), **ops)

    # ---------------- driver trajectories -------------------------------
    tr = {}
    f, h, L, x0 = ref.D_opt_design(80, 200, randseed=10)
    V = f.H
    out = quiet(ref.BPG, f, h, L, x0, maxitrs=1000, linesearch=True, ls_ratio=1.2, verbskip=100)
    tr["bpg_ls_x"], tr["bpg_ls_F"], tr["bpg_ls_Ls"] = out[0], out[1], out[2]
    out = quiet(ref.BPG, f, h, L, x0, maxitrs=300, linesearch=False, verbskip=100)
    tr["bpg_x"], tr["bpg_F"] = out[0], out[1]
    out = quiet(ref.ABPG, f, h, L, x0, gamma=2, maxitrs=1000, theta_eq=False, verbskip=100)
    tr["abpg_x"], tr["abpg_F"], tr["abpg_G"] = out[0], out[1], out[2]
    out = quiet(ref.ABPG, f, h, L, x0, gamma=2, maxitrs=1000, theta_eq=True, verbskip=100)
    tr["abpg_eq_F"], tr["abpg_eq_G"] = out[1], out[2]
    out = quiet(ref.ABPG, f, h, L, x0, gamma=2, maxitrs=400, theta_eq=True, restart=True, verbskip=100)
    tr["abpg_rs_F"], tr["abpg_rs_G"] = out[1], out[2]
    out = quiet(ref.ABPG_expo, f, h, L, x0, gamma0=3, maxitrs=600, theta_eq=True, verbskip=100)
    tr["expo_x"], tr["expo_F"], tr["expo_Gamma"], tr["expo_G"] = out[0], out[1], out[2], out[3]
    out = quiet(ref.ABPG_gain, f, h, L, x0, gamma=2, maxitrs=1000, G0=0.1, theta_eq=True, verbskip=100)
    tr["gain_x"], tr["gain_F"], tr["gain_Gain"], tr["gain_Gdiv"], tr["gain_Gavg"] = out[:5]
    out = quiet(ref.ABPG_gain, f, h, L, x0, gamma=2, maxitrs=400, G0=1, theta_eq=False, restart=True,
                verbskip=100)
    tr["gain_rs_F"], tr["gain_rs_Gain"] = out[1], out[2]
    out = quiet(ref.ABDA, f, h, L, x0, gamma=2, maxitrs=600, theta_eq=True, verbskip=100)
    tr["abda_x"], tr["abda_F"], tr["abda_G"] = out[0], out[1], out[2]
    out = quiet(ref.FW_alg_div_step, f, h, L, x0, maxitrs=300, gamma=2.0, lmo=ref.lmo_simplex(),
                ls_ratio=2, verbskip=100)
    tr["fwdiv_x"], tr["fwdiv_F"], tr["fwdiv_Ls"] = out[0], out[1], out[2]
    out = quiet(ref.FW_alg_descent_step, f, h, x0, maxitrs=300, lmo=ref.lmo_simplex(), verbskip=100)
    tr["fwdesc_x"], tr["fwdesc_F"] = out[0], out[1]
    out = quiet(ref.FW_alg_L0_L1_shortest_step, f, h, 1.0, 1.0, x0, 200, 2.0, ref.lmo_simplex(), ls_ratio=2,
                verbskip=1000)
    tr["fwl0l1s_x"], tr["fwl0l1s_F"], tr["fwl0l1s_Ls"] = out[0], out[1], out[2]
    out = quiet(ref.FW_l0l1_log_and_linear_step, f, h, 1.0, 1.0, x0, 200, ref.lmo_simplex(), 2, verbskip=1000)
    tr["fwl0l1ll_x"], tr["fwl0l1ll_F"], tr["fwl0l1ll_Ls"], tr["fwl0l1ll_LOG"] = out[0], out[1], out[2], out[3]
    out = quiet(ref.FW_l0l1_log_only, f, h, 1.0, 1.0, x0, 200, ref.lmo_simplex(), 2, verbskip=1000)
    tr["fwl0l1lo_x"], tr["fwl0l1lo_F"], tr["fwl0l1lo_Ls"], tr["fwl0l1lo_LOG"] = out[0], out[1], out[2], out[3]
    out = quiet(ref.D_opt_FW, V, x0, 1e-8, 2000, verbskip=1000)
    tr["dfw_x"], tr["dfw_F"], tr["dfw_SP"], tr["dfw_SN"] = out[:4]
    out = quiet(ref.D_opt_FW_away, V, x0, 1e-8, 2000, verbskip=1000)
    tr["dfwa_x"], tr["dfwa_F"], tr["dfwa_SP"], tr["dfwa_SN"] = out[:4]
    np.random.seed(77)
    xky = ref.D_opt_KYinit(V)
    tr["ky_x0"] = xky
    out = quiet(ref.D_opt_FW_away, V, xky, 1e-8, 1000, verbskip=1000)
    tr["dfwa_ky_F"], tr["dfwa_ky_SP"], tr["dfwa_ky_SN"] = out[1], out[2], out[3]

    # housing: BPG and BPG-LS rows at k=1000 of ipynb/ex_Dopt_LIBSVM.ipynb (stored stdout of cell 6)
    f, h, L, x0 = ref.D_opt_libsvm(os.path.join(REF_ROOT, "parameters_free_fw/data/housing.txt"))
    out = quiet(ref.BPG, f, h, L, x0, maxitrs=1001, linesearch=True, ls_ratio=1.2, verbskip=1000)
    tr["housing_bpg_ls_F"], tr["housing_bpg_ls_Ls"] = out[1], out[2]
    out = quiet(ref.BPG, f, h, L, x0, maxitrs=1001, linesearch=False, verbskip=1000)
    tr["housing_bpg_F"] = out[1]

    # KL + Shannon-L1, Poisson + Burg-L1 / Burg-L2
    f, h, L, x0 = ref.KL_nonneg_regr(300, 120, noise=0.01, lamdaL1=0.001, randseed=1)
    out = quiet(ref.BPG, f, h, L, x0, maxitrs=300, linesearch=True, verbskip=100)
    tr["kl_bpg_F"], tr["kl_bpg_Ls"] = out[1], out[2]
    out = quiet(ref.ABPG_gain, f, h, L, x0, gamma=2.0, maxitrs=300, verbskip=100)
    tr["kl_gain_F"], tr["kl_gain_Gain"] = out[1], out[2]
    f, h, L, x0 = ref.Poisson_regrL1(200, 100, noise=1e-4, lamda=0, randseed=1)
    out = quiet(ref.BPG, f, h, L, x0, maxitrs=300, linesearch=True, verbskip=100)
    tr["poi_bpg_F"], tr["poi_bpg_Ls"] = out[1], out[2]
    f, h, L, x0 = ref.Poisson_regrL2(200, 100, noise=1e-3, lamda=1e-3, randseed=1)
    out = quiet(ref.ABPG_gain, f, h, L, x0, gamma=2.0, maxitrs=300, verbskip=100)
    tr["poi2_gain_F"], tr["poi2_gain_Gain"] = out[1], out[2]
    # KL with Shannon simplex kernel (config C3 shape family), x* on the simplex
    np.random.seed(5)
    A = np.random.rand(150, 400)
    A = A / A.sum(axis=0)
    xs = np.random.rand(400)
    xs /= xs.sum()
    b = np.dot(A, xs) * (1 + 0.01 * (np.random.rand(150) - 0.5))
    tr["kls_A"], tr["kls_b"] = A, b
    f = ref.KLdivRegression(A, b)
    h = ref.ShannonEntropySimplex()
    x0 = np.ones(400) / 400
    out = quiet(ref.ABPG_gain, f, h, 1.0, x0, gamma=2.0, maxitrs=300, verbskip=100)
    tr["kls_gain_x"], tr["kls_gain_F"], tr["kls_gain_Gain"] = out[0], out[1], out[2]
    out = quiet(ref.FW_alg_div_step, f, h, 1.0, x0, maxitrs=100, gamma=2.0, lmo=ref.lmo_simplex(),
                verbskip=100)
    tr["kls_fw_F"], tr["kls_fw_Ls"] = out[1], out[2]
    np.savez_compressed(os.path.join(OUT, "trajectories.npz"), **tr)
    for name in ("operators.npz", "trajectories.npz"):
        print(name, os.path.getsize(os.path.join(OUT, name)), "bytes")


if __name__ == "__main__":
    main()
