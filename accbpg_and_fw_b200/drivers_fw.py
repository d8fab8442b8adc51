"""Generic Frank-Wolfe drivers over GPU-resident iterates.

FW_alg_div_step (Bregman-divergence step size with backtracking on L) and FW_alg_descent_step
(2/(k+2) step) with the signatures, defaults, error behaviour and return tuples of
accbpg/algorithms_fw.py:6-75 and :210-247.  `lmo` is a callable g -> s; the simplex LMO of this
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
    for k in range(maxitrs):
        g = lp.enq_fg(x, rt.S_F)
        lp.enq_psi(x)
        s = _call_lmo(lp, lmo, g)
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
            if not linesearch:
                break
            assert not math.isinf(L), "L is infinite"
            lp.enq_f(x1, rt.S_F2)
            if lp.fetch()[rt.S_F2] <= fx + alpha * gdp + alpha ** gamma * L * div:     # algorithms_fw.py:61
                break
            L = L * ls_ratio
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
    g = lp.enq_fg(x, rt.S_F)
    lp.enq_psi(x)
    vals = lp.fetch()
    F[0] = vals[rt.S_F] + lp.psi(vals)
    T[0] = lp.now()
    k = 0
    for k in range(1, maxitrs):
        s = _call_lmo(lp, lmo, g)
        alpha = 2 / (k + 2)
        x = _step(lp, x, s, alpha)
        g = lp.enq_fg(x, rt.S_F)
        lp.enq_psi(x)
        nat.check(lib.accbpg_vec_dot(rt.ctx, rt.stream, lp.n, g.data_ptr(), g.data_ptr(), rt.slot(rt.S_DOT)))
        if lp.shard is not None and lp.shard.world > 1:
            lp.shard.sum_(rt.scal[rt.S_DOT:rt.S_DOT + 1])
        vals = lp.fetch()
        F[k] = vals[rt.S_F] + lp.psi(vals)
        T[k] = lp.now()
        if verbose and (k % verbskip == 0 or k == 1):
            print(f"{k:6d}  {F[k]:10.3e}  {alpha:10.3e}  {T[k]:6.1f}")
        if abs(F[k] - F[k - 1]) < epsilon or math.sqrt(vals[rt.S_DOT]) < epsilon:
            break
    return lp.result(x), F[:k + 1], T[:k + 1], G[:k + 1]
