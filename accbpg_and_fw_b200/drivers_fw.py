"""Generic Frank-Wolfe drivers over GPU-resident iterates.

FW_alg_div_step (Bregman-divergence step size with backtracking on L), FW_alg_descent_step (2/(k+2) step) and the
three (L0,L1)-smooth variants FW_alg_L0_L1_shortest_step, FW_l0l1_log_and_linear_step, FW_l0l1_log_only, with the
signatures, defaults, error behaviour and return tuples of accbpg/algorithms_fw.py:6-75, :78-207, :210-247, :250-349
and :352-453.  `lmo` is a callable g -> s; the simplex LMO of this
package runs on the device, any other callable is used through its public (NumPy/tensor) form.
"""
import math
import time

import numpy as np

from . import _native as nat
from .drivers import _Loop

lib = nat.lib


def _call_lmo(lp, lmo, g):
    """Vertex s for gradient g as a device vector."""
    rt = lp.rt
    if hasattr(lmo, "_enq"):
        s = rt.empty(lp.n)
        lmo._enq(g, s, rt.S_AUX0)
        return s
    return rt.to_device(lmo(g))


def _vertex_image(lp, f, lmo, s):
    """Objective image of the LMO's answer: for the simplex LMO over a D-optimal objective from H H^T and the chosen
    column (DOptimalObj._img_vertex; the LMO left the index in S_AUX0 + 1), otherwise one pass (None when images are off)."""
    if lp.lin and hasattr(lmo, "_enq") and hasattr(f, "_img_vertex"):
        return f._img_vertex(lp.rt.S_AUX0 + 1, 1e-15, lmo.radius)
    return lp.img(s)


def _step(lp, x, s, alpha):
    """x + alpha*(s - x)   (algorithms_fw.py:34,54 / :227-231)."""
    rt = lp.rt
    out = rt.empty(lp.n)
    nat.check(lib.accbpg_vec_step_toward(rt.ctx, rt.stream, lp.n, x.data_ptr(), s.data_ptr(), float(alpha),
                                         out.data_ptr()))
    return out


def FW_alg_div_step(f, h, L, x0, maxitrs, gamma, lmo,
                    epsilon=1e-14, linesearch=True, ls_ratio=2,
                    verbose=True, verbskip=1):
    """Frank-Wolfe with the relative-smoothness step size.   Returns (x, F, Ls, T)."""
    if ls_ratio < 1:
        raise ValueError("ls_ratio must be >= 1")
    if L <= 0:
        raise ValueError("Initial L must be positive")
    if epsilon <= 0:
        raise ValueError("epsilon must be positive")
    if verbose:
        print("\nFW adaptive algorithm")
        print("     k      F(x)         Lk       time")
    lp = _Loop(f, h, x0)
    rt = lp.rt
    F, Ls, T = [], [], []
    delta = 1e-6
    x = lp.x0
    # With config.linear_images the objective's image of every trial point x + alpha (s - x) is the combination
    # (1 - alpha) I(x) + alpha I(s): one image of s per iteration instead of one pass over H / A per trial.  For the
    # simplex LMO over a D-optimal objective I(s) itself needs no pass either (DOptimalObj._img_vertex).
    Ix = lp.img(x)
    for k in range(maxitrs):
        g = lp.enq_start(None, None, None, x, Ix, 2, rt.S_F)
        lp.enq_psi(x)
        s = _call_lmo(lp, lmo, g)
        Is = _vertex_image(lp, f, lmo, s)
        lp.enq_div(s, x, rt.S_DXY)
        lp.enq_dot_diff(g, s, x)                     # <g, s - x>
        vals = lp.fetch()
        fx = vals[rt.S_F]
        F.append(fx + lp.psi(vals))
        T.append(lp.now())
        div = vals[rt.S_DXY]
        if div == 0:
            div = delta
        gdp = vals[rt.S_DOT]
        if 0 < gdp <= delta:
            gdp = 0.0
        if gdp > 0:
            raise ValueError("grad_d_prod must be non-positive")
        if linesearch:
            L = L / ls_ratio
        while True:
            alpha = min((-gdp / (2 * L * div)) ** (1 / (gamma - 1)), 1.0)
            x1 = _step(lp, x, s, alpha)
            Ix1 = lp.img_combo(1 - alpha, Ix, alpha, Is)
            if not linesearch:
                break
            assert not math.isinf(L), "L is infinite"
            lp.enq_f_img(x1, Ix1, rt.S_F2)
            if lp.fetch()[rt.S_F2] <= fx + alpha * gdp + alpha ** gamma * L * div:     # algorithms_fw.py:61
                break
            L = L * ls_ratio
        Ix = lp.img_refresh(k, x1, Ix1)
        x = x1
        Ls.append(L)
        if verbose and k % verbskip == 0:
            print(f"{k:6d}  {F[k]:10.3e}  {L:10.3e}  {T[k]:6.1f}")
        if k > 0 and abs(F[k] - F[k - 1]) < epsilon:
            break
    return lp.result(x), np.array(F), np.array(Ls), np.array(T)


def FW_alg_descent_step(f, h, x0, maxitrs, lmo, epsilon=1e-14, verbose=True, verbskip=1):
    """Frank-Wolfe with the 2/(k+2) step.   Returns (x, F, T, G); G stays zero as in the reference."""
    if verbose:
        print("\nFW descent step size algorithm")
        print("     k      F(x)         alpha_k       time")
    lp = _Loop(f, h, x0)
    rt = lp.rt
    F = np.zeros(maxitrs)
    G = np.zeros(maxitrs)
    T = np.zeros(maxitrs)
    x = lp.x0
    Ix = lp.img(x)
    g = lp.enq_start(None, None, None, x, Ix, 2, rt.S_F)
    lp.enq_psi(x)
    vals = lp.fetch()
    F[0] = vals[rt.S_F] + lp.psi(vals)
    T[0] = lp.now()
    k = 0
    for k in range(1, maxitrs):
        s = _call_lmo(lp, lmo, g)
        Is = _vertex_image(lp, f, lmo, s)
        alpha = 2 / (k + 2)
        x = _step(lp, x, s, alpha)
        Ix = lp.img_refresh(k, x, lp.img_combo(1 - alpha, Ix, alpha, Is))
        g = lp.enq_start(None, None, None, x, Ix, 2, rt.S_F)
        lp.enq_psi(x)
        nat.check(lib.accbpg_vec_dot(rt.ctx, rt.stream, lp.n, g.data_ptr(), g.data_ptr(), rt.slot(rt.S_DOT)))
        vals = lp.fetch()                                # sharded: the partial slots are summed over the ranks here
        F[k] = vals[rt.S_F] + lp.psi(vals)
        T[k] = lp.now()
        if verbose and (k % verbskip == 0 or k == 1):
            print(f"{k:6d}  {F[k]:10.3e}  {alpha:10.3e}  {T[k]:6.1f}")
        if abs(F[k] - F[k - 1]) < epsilon or math.sqrt(vals[rt.S_DOT]) < epsilon:
            break
    return lp.result(x), F[:k + 1], T[:k + 1], G[:k + 1]


# ---------------------------------------------------------------------------------------------- (L0, L1)-smooth variants
def _enq_gnorm2(lp, g, slot):
    """||g||^2 -> slot (one of the per-rank partial slots S_DXY..S_PSI: summed over the ranks at fetch)."""
    rt = lp.rt
    nat.check(lib.accbpg_vec_dot(rt.ctx, rt.stream, lp.n, g.data_ptr(), g.data_ptr(), rt.slot(slot)))


def _enq_dnorm2(lp, s, x, slot):
    """||s - x||^2 -> slot."""
    rt = lp.rt
    nat.check(lib.accbpg_vec_sqdist(rt.ctx, rt.stream, lp.n, s.data_ptr(), x.data_ptr(), rt.slot(slot)))


def FW_alg_L0_L1_shortest_step(f, h, L0, L1, x0, maxitrs, gamma, lmo, epsilon=1e-14,
                               linesearch=True, ls_ratio=2, verbose=True, verbskip=1):
    """Frank-Wolfe for (L0,L1)-smooth f with the shortest-step rule.   algorithms_fw.py:78-207.   Returns (x, F, Ls, T)."""
    if ls_ratio < 1:
        raise ValueError("ls_ratio must be >= 1")
    if L0 < 0 or L1 < 0:
        raise ValueError("Initial L must be positive")
    if epsilon <= 0:
        raise ValueError("epsilon must be positive")
    if verbose:
        print("\nFW (L0,L1)-smooth algorithm with shortest-step rule")
        print("     k        F(x)          a_k           L0            L1        alpha        time")
    lp = _Loop(f, h, x0)
    rt = lp.rt
    F, Ls, T = [], [], []
    delta = 1e-8
    x = lp.x0
    Ix = lp.img(x)
    toggle = 0
    for k in range(maxitrs):
        g = lp.enq_start(None, None, None, x, Ix, 2, rt.S_F)
        lp.enq_psi(x)
        s = _call_lmo(lp, lmo, g)
        Is = _vertex_image(lp, f, lmo, s)
        lp.enq_dot_diff(g, s, x)                     # <g, s - x>
        lp.enq_div(s, x, rt.S_DXY)
        _enq_gnorm2(lp, g, rt.S_DZZ)
        vals = lp.fetch()
        fx, gdp, g_norm, div = vals[rt.S_F], vals[rt.S_DOT], math.sqrt(vals[rt.S_DZZ]), vals[rt.S_DXY]
        F.append(fx + lp.psi(vals))
        T.append(lp.now())
        if div == 0:
            div = delta
        if 0 < gdp <= delta:
            gdp = 0
        if gdp > 0:
            raise ValueError("\u27e8\u2207f(x), d\u27e9 must be nonpositive (LMO issue).")
        a_k = L0 + L1 * g_norm
        if linesearch:
            L0 /= ls_ratio + L0 / a_k
            L1 /= ls_ratio + (L1 * g_norm) / a_k
        while True:
            a_k = L0 + L1 * g_norm
            alpha_k = min((-gdp / (a_k * div * np.e)) ** (1 / (gamma - 1)), 1)
            x1 = _step(lp, x, s, alpha_k)
            Ix1 = lp.img_combo(1 - alpha_k, Ix, alpha_k, Is)
            if not linesearch:
                break
            lp.enq_f_img(x1, Ix1, rt.S_F2)
            if lp.fetch()[rt.S_F2] <= fx + alpha_k * gdp + alpha_k ** gamma * (a_k / 2) * np.e * div:
                break
            if toggle == 0:
                L0 *= ls_ratio - L0 / a_k
                toggle = 1
            else:
                L1 *= ls_ratio - (L1 * g_norm) / a_k
                toggle = 0
        x = x1
        Ix = lp.img_refresh(k, x1, Ix1)
        Ls.append(a_k)
        if verbose and k % verbskip == 0:
            print(f"{k:6d}   {F[k]:10.3e}   {Ls[k]:10.3e}   {L0:10.3e}   {L1:10.3e}   {alpha_k:10.3e}   {T[k]:6.1f}")
        if k > 0 and abs(F[k] - F[k - 1]) < epsilon:
            break
    return lp.result(x), np.array(F), np.array(Ls), np.array(T)


def _fw_l0l1_log(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon, L0_max, L1_max, linesearch, verbose, verbskip,
                 log_only):
    if ls_ratio < 1:
        raise ValueError("ls_ratio must be >= 1")
    if L0 <= 0 or L1 <= 0:
        raise ValueError("Initial L0 and L1 must be positive")
    if epsilon <= 0:
        raise ValueError("epsilon must be positive")
    if verbose:
        print("\nFW L0,L1 smooth algorithm with fixed L1" if log_only else "\nFW L0,L1 smooth logarithmic algorithm")
        print("     k      F(x)         L         L0         L1     log step count       time")
    lp = _Loop(f, h, x0)
    rt = lp.rt
    F, Ls, T, LOG_STEPS = [], [], [], []
    delta = 1e-8
    toggle = 0
    x = lp.x0
    Ix = lp.img(x)
    for k in range(maxitrs):
        g = lp.enq_start(None, None, None, x, Ix, 2, rt.S_F)
        lp.enq_psi(x)
        s = _call_lmo(lp, lmo, g)
        Is = _vertex_image(lp, f, lmo, s)
        lp.enq_dot_diff(g, s, x)
        _enq_gnorm2(lp, g, rt.S_DXY)
        _enq_dnorm2(lp, s, x, rt.S_DZZ)
        vals = lp.fetch()
        fx, gdp = vals[rt.S_F], vals[rt.S_DOT]
        gx_norm, d_norm = math.sqrt(vals[rt.S_DXY]), math.sqrt(vals[rt.S_DZZ])
        F.append(fx + lp.psi(vals))
        T.append(lp.now())
        if 0 < gdp <= delta:
            gdp = 0
        if gdp > 0:
            raise ValueError("grad_d_prod must be non-positive (we minimize)")
        if linesearch:
            L0 /= ls_ratio
            L1 /= ls_ratio
        if log_only:
            L1 = max(math.log(2) / d_norm, L1)
        if k == 0:
            LOG_STEPS.append(0)
        while True:
            assert L0 >= 0 and L1 >= 0, "Smoothness parameters must stay positive"
            a_k = L0 + L1 * gx_norm
            if log_only:
                z = L1 * d_norm
                if z >= math.log(2) - 1e-5:
                    alpha_k = (1 / (L1 * d_norm)) * math.log(1 - (L1 * gdp) / (a_k * d_norm))
                    LOG_STEPS.append(LOG_STEPS[-1] + 1)
                else:
                    assert False, "No use for the second step!"
            elif L1 * d_norm >= np.log(2):
                alpha_k = (1 / (L1 * d_norm)) * np.log(1 - (L1 * gdp) / (a_k * d_norm))
                LOG_STEPS.append(LOG_STEPS[-1] + 1)
            else:
                alpha_k = L1 * (-gdp) / (a_k * d_norm)
                LOG_STEPS.append(LOG_STEPS[-1])
            x1 = _step(lp, x, s, alpha_k)
            Ix1 = lp.img_combo(1 - alpha_k, Ix, alpha_k, Is)
            if not linesearch:
                break
            lp.enq_f_img(x1, Ix1, rt.S_F2)
            fx1 = lp.fetch()[rt.S_F2]
            z = L1 * alpha_k * d_norm
            exp_term = np.expm1(z) - z if z < 50 else 0.5 * z ** 2
            rhs = fx + alpha_k * gdp + (a_k / L1 ** 2) * exp_term
            if fx1 <= rhs:
                break
            if log_only:
                if toggle == 0:
                    L0 = min(L0 * ls_ratio, L0_max) if L0_max else L0 * ls_ratio
                    toggle = 1
                else:
                    L1 = min(L1 * ls_ratio, L1_max) if L1_max else L1 * ls_ratio
                    toggle = 0
            else:
                L0 = min(L0 * ls_ratio, L0_max) if L0_max else L0 * ls_ratio
                L1 = min(L1 * ls_ratio, L1_max) if L1_max else L1 * ls_ratio
            a_k = L0 + L1 * gx_norm
        x = x1
        Ix = lp.img_refresh(k, x1, Ix1)
        Ls.append(a_k)
        if verbose and k % verbskip == 0:
            print(f"{k:6d}   {F[k]:10.3e}   {Ls[k]:10.3e}   {L0:10.3e}   {L1:10.3e}   {LOG_STEPS[k]:6d}      {T[k]:6.1f}")
        if k > 0 and abs(F[k] - F[k - 1]) < epsilon:
            break
    return lp.result(x), np.array(F), np.array(Ls), np.array(LOG_STEPS), np.array(T)


def FW_l0l1_log_and_linear_step(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon=1e-14, L0_max=None, L1_max=None,
                                linesearch=True, verbose=True, verbskip=50):
    """Logarithmic step when L1*||d|| >= ln 2, linear otherwise.   algorithms_fw.py:250-349.
    Returns (x, F, Ls, LOG_STEPS, T)."""
    return _fw_l0l1_log(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon, L0_max, L1_max, linesearch, verbose, verbskip,
                        False)


def FW_l0l1_log_only(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon=1e-14, L0_max=None, L1_max=None,
                     linesearch=True, verbose=True, verbskip=50):
    """L1 is raised to ln 2 / ||d|| so that the step is always logarithmic.   algorithms_fw.py:352-453.
    Returns (x, F, Ls, LOG_STEPS, T)."""
    return _fw_l0l1_log(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon, L0_max, L1_max, linesearch, verbose, verbskip,
                        True)
