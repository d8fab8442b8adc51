"""
CPU oracle for the accbpg hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a NumPy restatement of the arithmetic that the reference
(DredderGun/accbpg_and_fw, a pure NumPy package) performs on the per-iteration
hot path.  It exists so that the CUDA path can be checked on the GPU box, where
/root/reference is not available.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  Nothing under
accbpg_and_fw_b200/ imports it, and the product path raises when the CUDA
library is missing instead of falling back to this file.

Pinning: the reference's own tests do not cover this path (accbpg/functions_test.py
tests an unrelated class).  The oracle is therefore pinned two ways:
  * oracle/gen_golden.py imports the real reference (with stub modules for the
    absent cvxpy/jax/matplotlib) in the build container and stores its outputs in
    tests/golden/*.npz; tests/test_oracle_golden.py replays them against this file;
  * the 4-significant-digit values printed in the reference's stored notebook
    outputs are asserted in the same test file.

Every function cites the reference lines it restates (paths relative to the
reference root).  Expression order is kept identical to the reference wherever a
different order would change rounding (Python builtin ``sum`` = strict
left-to-right adds; ``ndarray.sum(axis=0)`` = NumPy pairwise).

Layout of this file: stateless operator functions first, then small adapter
objects that give them the duck-typed f / h / lmo protocol the drivers use, then
the driver loops, then the instance builders.
"""
import math
import time

import numpy as np

_pysum = sum  # the reference relies on the builtin (sequential) sum


# --------------------------------------------------------------------------
# objective oracles
# --------------------------------------------------------------------------

def dopt_eval(H, x, flag=2):
    """-log det(H diag(x) H^T) and its gradient.  accbpg/functions.py:43-59."""
    n = H.shape[1]
    assert x.size == n, "DOptimalObj: x.size not equal to n"
    assert x.min() >= 0, "DOptimalObj: x needs to be nonnegative"
    M = np.dot(H * x, H.T)                                   # :46
    sign, logdet = np.linalg.slogdet(M)                      # :48
    if sign <= 0:                                            # :49-50
        raise ValueError("HXHT is singular or not positive definite")
    fval = -logdet
    if flag == 0:
        return fval
    Z = np.linalg.solve(M, H)                                # :57
    grad = -np.sum(H * Z, axis=0)                            # :58
    return grad if flag == 1 else (fval, grad)


def poisson_eval(A, b, x, flag=2):
    """D_KL(b, Ax) and gradient A^T(1 - b/Ax).  accbpg/functions.py:102-120."""
    m, n = A.shape
    assert x.size == n, "PoissonRegression: x.size not equal to n."
    Ax = np.dot(A, x)                                        # :104
    if flag == 0:
        return _pysum(b * np.log(b / Ax) + Ax - b)           # :106
    grad = ((1 - b / Ax).reshape(m, 1) * A).sum(axis=0)      # :110
    if flag == 1:
        return grad
    return _pysum(b * np.log(b / Ax) + Ax - b), grad         # :119


def kl_eval(A, b, x, flag=2):
    """D_KL(Ax, b) and gradient A^T log(Ax/b).  accbpg/functions.py:140-158."""
    m, n = A.shape
    assert x.size == n, "NonnegRegression: x.size not equal to n."
    Ax = np.dot(A, x)                                        # :142
    if flag == 0:
        return _pysum(Ax * np.log(Ax / b) - Ax + b)          # :144
    grad = (np.log(Ax / b).reshape(m, 1) * A).sum(axis=0)    # :148
    if flag == 1:
        return grad
    return _pysum(Ax * np.log(Ax / b) - Ax + b), grad        # :157


# --------------------------------------------------------------------------
# Burg entropy  h(x) = -sum log x     accbpg/functions.py:238-356
# --------------------------------------------------------------------------

def burg_value(x):
    assert x.min() > 0, "BurgEntropy only takes positive arguments."
    return -_pysum(np.log(x))                                # :244


def burg_gradient(x):
    assert x.min() > 0, "BurgEntropy only takes positive arguments."
    return -1 / x                                            # :248


def burg_divergence(x, y):
    assert x.shape == y.shape, "Vectors x and y are of different sizes."
    assert x.min() > 0 and y.min() > 0, "Entries of x or y not positive."
    return _pysum(x / y - np.log(x / y) - 1)                 # :253


def burg_prox(g, L):
    assert L > 0, "BurgEntropy prox_map only takes positive L value."
    assert g.min() > 0, "BurgEntropy prox_map only takes positive value."
    return L / g                                             # :262


def burg_l1_prox(g, L, lamda):
    assert L > 0, "BurgEntropyL1: prox_map only takes positive L."
    assert g.min() > -lamda, "Not getting positive solution."
    return L / (lamda + g)                                   # :298


def burg_l2_prox(g, L, lamda):
    assert L > 0, "BurgEntropyL2: prox_map only takes positive L value."
    gg = g / L
    lam_L = lamda / L
    return (np.sqrt(gg * gg + 4 * lam_L) - gg) / (2 * lam_L)  # :321-323


def burg_simplex_prox(g, L, eps=1e-8, trace=None):
    """Newton root-find on sum 1/(g/L + c) = 1.  accbpg/functions.py:336-356.

    ``trace`` (optional list) receives (bisections, newton_steps, c) so tests can
    check that the CUDA kernel stops at the same iteration.
    """
    assert L > 0, "BergEntropySimplex prox_map only takes positive L."
    gg = g / L
    cmin = -gg.min()
    c = cmin + 1
    nbis = 0
    while _pysum(1 / (gg + c)) - 1 < 0:                      # :345
        c = (cmin + c) / 2.0
        nbis += 1
    fc = _pysum(1 / (gg + c)) - 1                            # :348
    nnewton = 0
    while abs(fc) > eps:                                     # :349
        fpc = _pysum(-1.0 / (gg + c) ** 2)                   # :350
        if (c - (c - fc / fpc)) == 0:                        # :351
            break
        c = c - fc / fpc
        fc = _pysum(1 / (gg + c)) - 1                        # :354
        nnewton += 1
    if trace is not None:
        trace.append((nbis, nnewton, c))
    return 1.0 / (gg + c)                                    # :355


def burg_div_prox(prox, y, g, L):
    """prox(g - L*grad_h(y), L) for every Burg class.  accbpg/functions.py:264-271."""
    assert y.shape == g.shape, "Vectors y and g are of different sizes."
    assert y.min() > 0 and L > 0, "Either y or L is not positive."
    return prox(g - L * burg_gradient(y), L)


# --------------------------------------------------------------------------
# Shannon entropy  h(x) = sum x log x     accbpg/functions.py:398-490
# --------------------------------------------------------------------------

def shannon_value(x, delta):
    assert x.min() >= 0, "ShannonEntropy takes nonnegative arguments."
    xx = np.maximum(x, delta)
    return _pysum(xx * np.log(xx))                           # :408


def shannon_gradient(x, delta):
    assert x.min() >= 0, "ShannonEntropy takes nonnegative arguments."
    xx = np.maximum(x, delta)
    return 1.0 + np.log(xx)                                  # :413


def shannon_divergence(x, y, delta):
    assert x.shape == y.shape, "Vectors x and y are of different shapes."
    assert x.min() >= 0 and y.min() >= 0, "Some entries are negative."
    return _pysum(x * np.log((x + delta) / (y + delta))) + (_pysum(y) - _pysum(x))  # :421


def shannon_prox(g, L):
    assert L > 0, "ShannonEntropy prox_map require L > 0."
    return np.exp(-g / L - 1)                                # :428


def shannon_div_prox(y, g, L):
    assert y.shape == g.shape, "Vectors y and g are of different sizes."
    assert y.min() >= 0 and L > 0, "Some entries of y are negavie."
    return y * np.exp(-g / L)                                # :438


def shannon_simplex_prox(g, L):
    assert L > 0, "ShannonEntropy prox_map require L > 0."
    x = np.exp(-g / L - 1)
    return x / _pysum(x)                                     # :480-481


def shannon_simplex_div_prox(y, g, L):
    assert y.shape == g.shape, "Vectors y and g are of different shapes."
    assert y.min() > 0 and L > 0, "prox_map needs positive arguments."
    x = y * np.exp(-g / L)
    return x / _pysum(x)                                     # :489-490


# --------------------------------------------------------------------------
# linear minimisation oracles     accbpg/functions_lmo.py
# --------------------------------------------------------------------------

def lmo_simplex_eval(g, radius=1):
    """functions_lmo.py:153-158: 1e-15 off the vertex, first index among ties."""
    s = np.zeros(g.shape)
    s += 1e-15
    s[np.where(g == np.min(g))[0][0]] = radius
    return s


def lmo_l2_ball_eval(g, radius, center=None):
    """functions_lmo.py:34-49."""
    cp_ = np.zeros_like(g) if center is None else np.broadcast_to(center, g.shape)
    gn = np.linalg.norm(g)
    if gn < 1e-10:
        return cp_
    s = cp_ - radius * g / gn
    assert abs(np.linalg.norm(s - cp_) - radius) <= 1e-10, "Solution does not lie on ball boundary"
    return s


def lmo_l2_ball_positive_orthant_eval(g, radius, center=None, epsilon=0.0):
    """functions_lmo.py:77-100."""
    g = np.asarray(g)
    cp_ = np.zeros_like(g) if center is None else np.asarray(center)
    assert cp_.shape == g.shape, "Shape mismatch between g and center"
    mask = g < 0
    if not np.any(mask):
        return np.maximum(cp_, epsilon)
    gneg = g[mask]
    direction = np.zeros_like(g)
    direction[mask] = -gneg / np.linalg.norm(gneg)
    s = np.maximum(cp_ + radius * direction, epsilon)
    assert np.all(s >= epsilon), "Output violates epsilon-nonnegativity"
    assert np.linalg.norm(s - cp_) <= radius + 1e-8, "Output outside L2 ball"
    return s


def lmo_linf_ball_eval(g, radius, center=None):
    """functions_lmo.py:122-132."""
    cp_ = np.zeros_like(g) if center is None else np.array(center)
    return cp_ - radius * np.sign(g)


def lmo_matrix_simplex_eval(G, radius=1.0):
    """functions_lmo.py:180-185: 1e-60 off the vertex."""
    S = np.zeros_like(G)
    S += 1e-60
    S[np.unravel_index(np.argmin(G), G.shape)] = radius
    return S


def lmo_matrix_box_eval(G, lower, upper):
    """functions_lmo.py:208-210."""
    return np.where(G < 0, upper, lower)


# --------------------------------------------------------------------------
# adapters: the duck-typed protocol of accbpg/functions.py:10-24 and :199-235
# --------------------------------------------------------------------------

class _Objective:
    def __init__(self, evalfn, *mats):
        self._ev = evalfn
        self._mats = mats

    def __call__(self, x):
        return self._ev(*self._mats, x, 0)

    def gradient(self, x):
        return self._ev(*self._mats, x, 1)

    def func_grad(self, x, flag=2):
        return self._ev(*self._mats, x, flag)


def make_dopt(H):
    """accbpg/functions.py:31-35."""
    assert H.shape[0] < H.shape[1], "DOptimalObj: need m < n"
    f = _Objective(dopt_eval, H)
    f.H, (f.m, f.n) = H, H.shape
    return f


def make_poisson(A, b):
    """accbpg/functions.py:89-94."""
    assert A.shape[0] == b.shape[0], "A and b sizes not matching"
    f = _Objective(poisson_eval, A, b)
    f.A, f.b, (f.m, f.n) = A, b, A.shape
    return f


def make_kl(A, b):
    """accbpg/functions.py:127-132."""
    assert A.shape[0] == b.shape[0], "A and b size not matching"
    f = _Objective(kl_eval, A, b)
    f.A, f.b, (f.m, f.n) = A, b, A.shape
    return f


class _Burg:
    """kind in {'plain','l1','l2','simplex'}; accbpg/functions.py:238-356."""

    def __init__(self, kind="plain", lamda=0.0, eps=1e-8):
        assert kind in ("plain", "l1", "l2", "simplex")
        self.kind, self.lamda, self.eps = kind, lamda, eps
        self.newton_trace = []

    def __call__(self, x):
        return burg_value(x)

    def gradient(self, x):
        return burg_gradient(x)

    def divergence(self, x, y):
        return burg_divergence(x, y)

    def extra_Psi(self, x):
        if self.kind == "l1":
            return self.lamda * x.sum()                      # :288
        if self.kind == "l2":
            return (self.lamda / 2) * np.dot(x, x)           # :314
        return 0                                             # :210-211

    def prox_map(self, g, L):
        if self.kind == "l1":
            return burg_l1_prox(g, L, self.lamda)
        if self.kind == "l2":
            return burg_l2_prox(g, L, self.lamda)
        if self.kind == "simplex":
            return burg_simplex_prox(g, L, self.eps, self.newton_trace)
        return burg_prox(g, L)

    def div_prox_map(self, y, g, L):
        return burg_div_prox(self.prox_map, y, g, L)


class _Shannon:
    """kind in {'plain','l1','simplex'}; accbpg/functions.py:398-490."""

    def __init__(self, kind="plain", lamda=0.0, delta=1e-20):
        assert kind in ("plain", "l1", "simplex")
        self.kind, self.lamda, self.delta = kind, lamda, delta

    def __call__(self, x):
        return shannon_value(x, self.delta)

    def gradient(self, x):
        return shannon_gradient(x, self.delta)

    def divergence(self, x, y):
        return shannon_divergence(x, y, self.delta)

    def extra_Psi(self, x):
        return self.lamda * x.sum() if self.kind == "l1" else 0   # :456

    def prox_map(self, g, L):
        if self.kind == "simplex":
            return shannon_simplex_prox(g, L)
        if self.kind == "l1":
            return shannon_prox(self.lamda + g, L)           # :462
        return shannon_prox(g, L)

    def div_prox_map(self, y, g, L):
        if self.kind == "simplex":
            return shannon_simplex_div_prox(y, g, L)
        if self.kind == "l1":
            return shannon_div_prox(y, self.lamda + g, L)    # :466
        return shannon_div_prox(y, g, L)


def make_burg(kind="plain", lamda=0.0, eps=1e-8):
    return _Burg(kind, lamda, eps)


def make_shannon(kind="plain", lamda=0.0, delta=1e-20):
    return _Shannon(kind, lamda, delta)


def make_lmo_simplex(radius=1):
    return lambda g: lmo_simplex_eval(g, radius)


# --------------------------------------------------------------------------
# driver loops.  Host control flow restated from accbpg/algorithms.py:11-514,
# accbpg/algorithms_fw.py:6-75, :210-247 and accbpg/D_opt_alg.py:9-185.
# verbose printing is dropped (it does not influence results).
# --------------------------------------------------------------------------

def solve_theta(theta, gamma, gainratio=1):
    """Newton solve of (1-t)/t^gamma = gainratio/theta^gamma.  algorithms.py:75-91."""
    ckg = theta ** gamma / gainratio
    cta = theta
    tol = 1e-6 * theta
    phi = cta ** gamma - ckg * (1 - cta)
    while abs(phi) > tol:
        drv = gamma * cta ** (gamma - 1) + ckg
        cta = cta - phi / drv
        phi = cta ** gamma - ckg * (1 - cta)
    return cta


def _restart_hit(rule, F, k, g, x, x_prev):
    """algorithms.py:168 / :279 / :406 (same predicate in the three drivers)."""
    return (rule == 'f' and F[k] > F[k - 1]) or (rule == 'g' and np.dot(g, x - x_prev) > 0)


def BPG(f, h, L, x0, maxitrs, epsilon=1e-14, linesearch=True, ls_ratio=1.2,
        verbose=False, verbskip=1):
    """algorithms.py:11-72.  Returns (x, F, Ls, T)."""
    t0 = time.time()
    F = np.zeros(maxitrs)
    Ls = np.ones(maxitrs) * L
    T = np.zeros(maxitrs)
    x = np.copy(x0)
    for k in range(maxitrs):
        fx, g = f.func_grad(x)                               # :46
        F[k] = fx + h.extra_Psi(x)
        T[k] = time.time() - t0
        if linesearch:
            L = L / ls_ratio                                 # :51
            x1 = h.div_prox_map(x, g, L)
            while f(x1) > fx + np.dot(g, x1 - x) + L * h.divergence(x1, x):   # :53
                L = L * ls_ratio
                x1 = h.div_prox_map(x, g, L)
            x = x1
        else:
            x = h.div_prox_map(x, g, L)
        Ls[k] = L
        if k > 0 and abs(F[k] - F[k - 1]) < epsilon:         # :66
            break
    return x, F[:k + 1], Ls[:k + 1], T[:k + 1]


def ABPG(f, h, L, x0, gamma, maxitrs, epsilon=1e-14, theta_eq=False,
         restart=False, restart_rule='g', verbose=False, verbskip=1):
    """algorithms.py:94-180.  Returns (x, F, G, T)."""
    t0 = time.time()
    F = np.zeros(maxitrs)
    G = np.zeros(maxitrs)
    T = np.zeros(maxitrs)
    x = np.copy(x0)
    z = np.copy(x0)
    theta = 1.0
    kk = 0
    for k in range(maxitrs):
        fx = f(x)                                            # :135
        F[k] = fx + h.extra_Psi(x)
        T[k] = time.time() - t0
        z_1 = z
        x_1 = x
        if theta_eq and kk > 0:
            theta = solve_theta(theta, gamma)
        else:
            theta = gamma / (kk + gamma)                     # :145
        y = (1 - theta) * x + theta * z_1                    # :147
        g = f.gradient(y)
        z = h.div_prox_map(z_1, g, theta ** (gamma - 1) * L)  # :149
        x = (1 - theta) * x + theta * z
        dxy = h.divergence(x, y)
        dzz = h.divergence(z, z_1)
        G[k] = dxy / dzz / theta ** gamma                    # :155
        kk += 1
        if restart and k > 0:                                # :165
            if _restart_hit(restart_rule, F, k, g, x, x_1):
                theta = 1.0
                kk = 0
                z = x
        if dzz < epsilon:
            break
    return x, F[:k + 1], G[:k + 1], T[:k + 1]


def ABPG_expo(f, h, L, x0, gamma0, maxitrs, epsilon=1e-14, delta=0.2,
              theta_eq=True, checkdiv=False, Gmargin=10, restart=False,
              restart_rule='g', verbose=False, verbskip=1):
    """algorithms.py:183-292.  Returns (x, F, Gamma, G, T)."""
    t0 = time.time()
    F = np.zeros(maxitrs)
    G = np.zeros(maxitrs)
    Gamma = np.ones(maxitrs) * gamma0
    T = np.zeros(maxitrs)
    gamma = gamma0
    x = np.copy(x0)
    z = np.copy(x0)
    theta = 1.0
    kk = 0
    for k in range(maxitrs):
        fx = f(x)
        F[k] = fx + h.extra_Psi(x)
        T[k] = time.time() - t0
        z_1 = z
        x_1 = x
        if theta_eq and kk > 0:
            theta = solve_theta(theta, gamma)
        else:
            theta = gamma / (kk + gamma)
        y = (1 - theta) * x_1 + theta * z_1                  # :243
        fy, g = f.func_grad(y)                               # :245
        again = True
        while again:                                         # :248
            z = h.div_prox_map(z_1, g, theta ** (gamma - 1) * L)
            x = (1 - theta) * x_1 + theta * z
            dxy = h.divergence(x, y)
            dzz = h.divergence(z, z_1)
            Gdr = dxy / dzz / theta ** gamma
            if checkdiv:
                again = (dxy > Gmargin * (theta ** gamma) * dzz)                      # :258
            else:
                again = (f(x) > fy + np.dot(g, x - y) + theta ** gamma * L * dzz)     # :260
            if again and gamma > 1:
                gamma = max(gamma - delta, 1)                # :263
            else:
                again = False
        G[k] = Gdr
        Gamma[k] = gamma
        kk += 1
        if restart:                                          # :276 (no k>0 guard)
            if _restart_hit(restart_rule, F, k, g, x, x_1):
                theta = 1.0
                kk = 0
                z = x
        if dzz < epsilon:
            break
    return x, F[:k + 1], Gamma[:k + 1], G[:k + 1], T[:k + 1]


def ABPG_gain(f, h, L, x0, gamma, maxitrs, epsilon=1e-14, G0=1,
              ls_inc=1.2, ls_dec=1.2, theta_eq=True, checkdiv=False,
              restart=False, restart_rule='g', verbose=False, verbskip=1):
    """algorithms.py:295-420.  Returns (x, F, Gain, Gdiv, Gavg, T)."""
    t0 = time.time()
    F = np.zeros(maxitrs)
    Gain = np.ones(maxitrs) * G0
    Gdiv = np.zeros(maxitrs)
    Gavg = np.zeros(maxitrs)
    T = np.zeros(maxitrs)
    x = np.copy(x0)
    z = np.copy(x0)
    G = G0
    sumlogG = gamma * np.log(G)                              # :340
    theta = 1.0
    kk = 0
    for k in range(maxitrs):
        fx = f(x)
        F[k] = fx + h.extra_Psi(x)
        T[k] = time.time() - t0
        z_1 = z
        x_1 = x
        G_1 = G
        theta_1 = theta
        G = G / ls_dec                                       # :358
        again = True
        while again:                                         # :361
            if kk > 0:
                if theta_eq:
                    theta = solve_theta(theta_1, gamma, G / G_1)
                else:
                    alpha = G / G_1
                    theta = theta_1 * ((1 + alpha * (gamma - 1)) / (gamma * alpha + theta_1))   # :367
            y = (1 - theta) * x_1 + theta * z_1
            fy, g = f.func_grad(y)
            z = h.div_prox_map(z_1, g, theta ** (gamma - 1) * G * L)   # :373
            x = (1 - theta) * x_1 + theta * z
            dxy = h.divergence(x, y)
            dzz = h.divergence(z, z_1)
            if dzz < epsilon:                                # :379
                break
            Gdr = dxy / dzz / theta ** gamma
            if checkdiv:
                again = (Gdr > G)
            else:
                again = (f(x) > fy + np.dot(g, x - y) + theta ** gamma * G * L * dzz)   # :387
            if again:
                G = G * ls_inc
        Gain[k] = G
        Gdiv[k] = Gdr                                        # :394 (stale value if the break above fired)
        sumlogG += np.log(G)
        Gavg[k] = np.exp(sumlogG / (gamma + k))
        kk += 1
        if restart:                                          # :403 (no k>0 guard)
            if _restart_hit(restart_rule, F, k, g, x, x_1):
                theta = 1.0
                kk = 0
                z = x
        if dzz < epsilon:
            break
    return x, F[:k + 1], Gain[:k + 1], Gdiv[:k + 1], Gavg[:k + 1], T[:k + 1]


def ABDA(f, h, L, x0, gamma, maxitrs, epsilon=1e-14, theta_eq=True,
         verbose=False, verbskip=1):
    """algorithms.py:423-514 (its restart branch is dead code: restart=False at :447)."""
    t0 = time.time()
    F = np.zeros(maxitrs)
    G = np.zeros(maxitrs)
    T = np.zeros(maxitrs)
    x = np.copy(x0)
    z = np.copy(x0)
    theta = 1.0
    kk = 0
    gavg = np.zeros(x.size)
    csum = 0
    for k in range(maxitrs):
        fx = f(x)
        F[k] = fx + h.extra_Psi(x)
        T[k] = time.time() - t0
        z_1 = z
        x_1 = x
        if theta_eq and kk > 0:
            theta = solve_theta(theta, gamma)
        else:
            theta = gamma / (kk + gamma)
        y = (1 - theta) * x_1 + theta * z_1
        g = f.gradient(y)
        gavg = gavg + theta ** (1 - gamma) * g               # :480
        csum = csum + theta ** (1 - gamma)
        z = h.prox_map(gavg / csum, L / csum)                # :482
        x = (1 - theta) * x_1 + theta * z
        dxy = h.divergence(x, y)
        dzz = h.divergence(z, z_1)
        G[k] = dxy / dzz / theta ** gamma
        kk += 1
        if dzz < epsilon:
            break
    return x, F[:k + 1], G[:k + 1], T[:k + 1]


def FW_alg_div_step(f, h, L, x0, maxitrs, gamma, lmo, epsilon=1e-14,
                    linesearch=True, ls_ratio=2, verbose=False, verbskip=1,
                    vertex_log=None):
    """algorithms_fw.py:6-75.  Returns (x, F, Ls, T).  ``vertex_log`` (list) gets argmax(s_k)."""
    if ls_ratio < 1:
        raise ValueError("ls_ratio must be >= 1")
    if L <= 0:
        raise ValueError("Initial L must be positive")
    if epsilon <= 0:
        raise ValueError("epsilon must be positive")
    t0 = time.time()
    F, Ls, T = [], [], []
    delta = 1e-6                                             # :24
    x = np.copy(x0)
    for k in range(maxitrs):
        fx, g = f.func_grad(x)
        F.append(fx + h.extra_Psi(x))
        T.append(time.time() - t0)
        s = lmo(g)                                           # :33
        if vertex_log is not None:
            vertex_log.append(int(np.argmax(s)))
        d = s - x
        div = h.divergence(s, x)
        if div == 0:
            div = delta
        gdp = np.dot(g.ravel(), d.ravel())                   # :39
        if 0 < gdp <= delta:
            gdp = 0.0
        if gdp > 0:
            raise ValueError("grad_d_prod must be non-positive")
        if linesearch:
            L = L / ls_ratio                                 # :46-47
        while True:
            alpha = min((-gdp / (2 * L * div)) ** (1 / (gamma - 1)), 1.0)   # :50-53
            x1 = x + alpha * d
            if not linesearch:
                break
            assert not math.isinf(L), "L is infinite"
            if f.func_grad(x1, flag=0) <= fx + alpha * gdp + alpha ** gamma * L * div:   # :61
                break
            L = L * ls_ratio
        x = x1
        Ls.append(L)
        if k > 0 and abs(F[k] - F[k - 1]) < epsilon:
            break
    return x, np.array(F), np.array(Ls), np.array(T)


def FW_alg_descent_step(f, h, x0, maxitrs, lmo, epsilon=1e-14, verbose=False, verbskip=1):
    """algorithms_fw.py:210-247.  Returns (x, F, T, G) -- G stays all-zero as in the reference."""
    t0 = time.time()
    F = np.zeros(maxitrs)
    G = np.zeros(maxitrs)
    T = np.zeros(maxitrs)
    x = np.copy(x0)
    fx, g = f.func_grad(x)
    F[0] = fx + h.extra_Psi(x)
    T[0] = time.time() - t0
    k = 0
    for k in range(1, maxitrs):
        s = lmo(g)
        x = x + (2 / (k + 2)) * (s - x)                      # :227-231
        fx, g = f.func_grad(x)
        F[k] = fx + h.extra_Psi(x)
        T[k] = time.time() - t0
        if abs(F[k] - F[k - 1]) < epsilon or np.linalg.norm(g) < epsilon:   # :240
            break
    return x, F[:k + 1], T[:k + 1], G[:k + 1]


def FW_alg_L0_L1_shortest_step(f, h, L0, L1, x0, maxitrs, gamma, lmo, epsilon=1e-14,
                               linesearch=True, ls_ratio=2, verbose=False, verbskip=1):
    """algorithms_fw.py:78-207.  Returns (x, F, Ls, T)."""
    if ls_ratio < 1:
        raise ValueError("ls_ratio must be >= 1")
    if L0 < 0 or L1 < 0:
        raise ValueError("Initial L must be positive")
    if epsilon <= 0:
        raise ValueError("epsilon must be positive")
    t0 = time.time()
    F, Ls, T = [], [], []
    delta = 1e-8                                             # :153
    x = np.copy(x0)
    toggle = 0
    for k in range(maxitrs):
        fx, g = f.func_grad(x)
        F.append(fx + h.extra_Psi(x))
        T.append(time.time() - t0)
        s = lmo(g)
        d = s - x
        div = h.divergence(s, x)
        if div == 0:
            div = delta
        gdp = np.dot(g.ravel(), d.ravel())                   # :167
        if 0 < gdp <= delta:
            gdp = 0
        if gdp > 0:
            raise ValueError("<grad f(x), d> must be nonpositive (LMO issue).")
        g_norm = np.linalg.norm(g)
        a_k = L0 + L1 * g_norm
        if linesearch:                                       # :176-178
            L0 /= ls_ratio + L0 / a_k
            L1 /= ls_ratio + (L1 * g_norm) / a_k
        while True:
            a_k = L0 + L1 * g_norm
            alpha = min((-gdp / (a_k * div * np.e)) ** (1 / (gamma - 1)), 1)      # :184
            x1 = x + alpha * d
            if not linesearch:
                break
            if f.func_grad(x1, flag=0) <= fx + alpha * gdp + alpha ** gamma * (a_k / 2) * np.e * div:   # :190
                break
            if toggle == 0:
                L0 *= ls_ratio - L0 / a_k
                toggle = 1
            else:
                L1 *= ls_ratio - (L1 * g_norm) / a_k
                toggle = 0
        x = x1
        Ls.append(a_k)
        if k > 0 and abs(F[k] - F[k - 1]) < epsilon:
            break
    return x, np.array(F), np.array(Ls), np.array(T)


def _fw_l0l1_log(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon, L0_max, L1_max, linesearch, log_only):
    """Shared body of FW_l0l1_log_and_linear_step (algorithms_fw.py:250-349) and FW_l0l1_log_only (:352-453)."""
    if ls_ratio < 1:
        raise ValueError("ls_ratio must be >= 1")
    if L0 <= 0 or L1 <= 0:
        raise ValueError("Initial L0 and L1 must be positive")
    if epsilon <= 0:
        raise ValueError("epsilon must be positive")
    t0 = time.time()
    F, Ls, T, LOG_STEPS = [], [], [], []
    delta = 1e-8
    toggle = 0
    x = np.copy(x0)
    for k in range(maxitrs):
        fx, g = f.func_grad(x)
        gx_norm = np.linalg.norm(g)
        F.append(fx + h.extra_Psi(x))
        T.append(time.time() - t0)
        s = lmo(g)
        d = s - x
        d_norm = np.linalg.norm(d)
        gdp = np.vdot(g, d)
        if 0 < gdp <= delta:
            gdp = 0
        if gdp > 0:
            raise ValueError("grad_d_prod must be non-positive (we minimize)")
        if linesearch:
            L0 /= ls_ratio
            L1 /= ls_ratio
        if log_only:
            L1 = max(math.log(2) / d_norm, L1)               # :401
        if k == 0:
            LOG_STEPS.append(0)
        while True:
            assert L0 >= 0 and L1 >= 0, "Smoothness parameters must stay positive"
            a_k = L0 + L1 * gx_norm
            if log_only:
                z = L1 * d_norm
                if z >= math.log(2) - 1e-5:                  # :410
                    alpha = (1 / (L1 * d_norm)) * math.log(1 - (L1 * gdp) / (a_k * d_norm))
                    LOG_STEPS.append(LOG_STEPS[-1] + 1)
                else:
                    assert False, "No use for the second step!"
            elif L1 * d_norm >= np.log(2):                   # :312
                alpha = (1 / (L1 * d_norm)) * np.log(1 - (L1 * gdp) / (a_k * d_norm))
                LOG_STEPS.append(LOG_STEPS[-1] + 1)
            else:
                alpha = L1 * (-gdp) / (a_k * d_norm)
                LOG_STEPS.append(LOG_STEPS[-1])
            x1 = x + alpha * d
            if not linesearch:
                break
            fx1 = f.func_grad(x1, flag=0)
            z = L1 * alpha * d_norm
            exp_term = np.expm1(z) - z if z < 50 else 0.5 * z ** 2          # :327-330
            rhs = fx + alpha * gdp + (a_k / L1 ** 2) * exp_term
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
        Ls.append(a_k)
        if k > 0 and abs(F[k] - F[k - 1]) < epsilon:
            break
    return x, np.array(F), np.array(Ls), np.array(LOG_STEPS), np.array(T)


def FW_l0l1_log_and_linear_step(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon=1e-14, L0_max=None, L1_max=None,
                                linesearch=True, verbose=False, verbskip=50):
    """algorithms_fw.py:250-349.  Returns (x, F, Ls, LOG_STEPS, T)."""
    return _fw_l0l1_log(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon, L0_max, L1_max, linesearch, False)


def FW_l0l1_log_only(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon=1e-14, L0_max=None, L1_max=None,
                     linesearch=True, verbose=False, verbskip=50):
    """algorithms_fw.py:352-453.  Returns (x, F, Ls, LOG_STEPS, T)."""
    return _fw_l0l1_log(f, h, L0, L1, x0, maxitrs, lmo, ls_ratio, epsilon, L0_max, L1_max, linesearch, True)


def _fw_setup(V, x0):
    """D_opt_alg.py:39-45 / :123-129."""
    x = np.copy(x0)
    VXVT = np.dot(V * x, V.T)
    det0 = np.linalg.det(VXVT)
    Hinv = np.linalg.inv(VXVT)
    w = np.sum(V * np.dot(Hinv, V), axis=0)
    return x, det0, Hinv, w


def D_opt_FW(V, x0, eps, maxitrs, verbose=False, verbskip=1, index_log=None):
    """D_opt_alg.py:9-88.  Returns (x, F, SP, SN, T).  ``index_log`` gets (i, j_full)."""
    t0 = time.time()
    m, n = V.shape
    F = np.zeros(maxitrs)
    SP = np.zeros(maxitrs)
    SN = np.zeros(maxitrs)
    T = np.zeros(maxitrs)
    x, detM, Hinv, w = _fw_setup(V, x0)
    for k in range(maxitrs):
        F[k] = -np.log(detM)                                 # :52
        T[k] = time.time() - t0
        i = np.argmax(w)
        support = x > 0                                      # :60
        w_pos = w[support]
        j = np.argmin(w_pos)
        SP[k] = pos = w[i] / m - 1
        SN[k] = neg = 1 - w_pos[j] / m
        if index_log is not None:
            index_log.append((int(i), int(np.flatnonzero(support)[j])))
        if pos <= eps and neg <= eps:
            break
        t = (w[i] / m - 1) / (w[i] - 1)                      # :75
        x *= (1 - t)
        x[i] += t
        u = np.dot(Hinv, V[:, i])
        coef = t / (1 + t * (w[i] - 1))
        Hinv = (Hinv - coef * np.outer(u, u)) / (1 - t)      # :79
        detM *= np.power(1 - t, m - 1) * (1 + t * (w[i] - 1))
        w = (w - coef * np.dot(u, V) ** 2) / (1 - t)         # :82
    return x, F[:k + 1], SP[:k + 1], SN[:k + 1], T[:k + 1]


def D_opt_FW_away(V, x0, eps, maxitrs, verbose=False, verbskip=1, index_log=None):
    """D_opt_alg.py:91-185.  Returns (x, F, SP, SN, T).  ``index_log`` gets (i, j, away?)."""
    t0 = time.time()
    m, n = V.shape
    F = np.zeros(maxitrs)
    SP = np.zeros(maxitrs)
    SN = np.zeros(maxitrs)
    T = np.zeros(maxitrs)
    x, detM, Hinv, w = _fw_setup(V, x0)
    for k in range(maxitrs):
        F[k] = np.log(np.linalg.det(Hinv))                   # :136 (fresh LU every iteration)
        T[k] = time.time() - t0
        i = np.argmax(w)
        ww = w - w[i]
        j = np.argmin(ww * [x > 1.0e-8])                     # :147 ((1,n) product, flattened argmin)
        SP[k] = pos = w[i] / m - 1
        SN[k] = neg = 1 - w[j] / m
        if index_log is not None:
            index_log.append((int(i), int(j), bool(pos < neg)))
        if pos <= eps and neg <= eps:
            break
        if pos >= neg:                                       # :162 toward vertex i
            t = (w[i] / m - 1) / (w[i] - 1)
            x *= (1 - t)
            x[i] += t
            u = np.dot(Hinv, V[:, i])
            coef = t / (1 - t + t * w[i])
            Hinv = (Hinv - coef * np.outer(u, u)) / (1 - t)
            w = (w - coef * np.dot(u, V) ** 2) / (1 - t)
        else:                                                # :171 away from vertex j
            t = min((1 - w[j] / m) / (w[j] - 1), x[j] / (1 - x[j]))
            x *= (1 + t)
            x[j] -= t
            u = np.dot(Hinv, V[:, j])
            coef = t / (1 + t - t * w[j])
            Hinv = (Hinv + coef * np.outer(u, u)) / (1 + t)
            w = (w + coef * np.dot(u, V) ** 2) / (1 + t)
    return x, F[:k + 1], SP[:k + 1], SN[:k + 1], T[:k + 1]


# --------------------------------------------------------------------------
# instance builders     accbpg/applications.py:17-206
# (legacy global NumPy RNG, same draw order as the reference)
# --------------------------------------------------------------------------

def _seed(randseed):
    if randseed > 0:
        np.random.seed(randseed)


def D_opt_design(m, n, randseed=-1):
    """applications.py:36-56."""
    _seed(randseed)
    H = np.random.randn(m, n)
    return make_dopt(H), make_burg("simplex"), 1.0, (1.0 / n) * np.ones(n)


def D_opt_from_matrix(X):
    """applications.py:22-33 after the loader: densify, put the long side on columns."""
    H = X.T.copy(order='C') if X.shape[0] > X.shape[1] else np.ascontiguousarray(X)
    n = H.shape[1]
    return make_dopt(H), make_burg("simplex"), 1.0, (1.0 / n) * np.ones(n)


def _poisson_instance(m, n, noise, randseed, normalizeA):
    """applications.py:116-125 / :155-164."""
    _seed(randseed)
    A = np.random.rand(m, n)
    if normalizeA:
        A = A / A.sum(axis=0)
    x = np.random.rand(n) / n
    xavg = x.sum() / x.size
    x = np.maximum(x - xavg, 0) * 10
    b = np.dot(A, x) + noise * (np.random.rand(m) - 0.5)
    assert b.min() > 0, "need b > 0 for nonnegative regression."
    return A, b


def Poisson_regrL1(m, n, noise=0.01, lamda=0, randseed=-1, normalizeA=True):
    """applications.py:98-134."""
    A, b = _poisson_instance(m, n, noise, randseed, normalizeA)
    return make_poisson(A, b), make_burg("l1", lamda), b.sum(), (1.0 / n) * np.ones(n) * 10


def Poisson_regrL2(m, n, noise=0.01, lamda=0, randseed=-1, normalizeA=True):
    """applications.py:137-172."""
    A, b = _poisson_instance(m, n, noise, randseed, normalizeA)
    return make_poisson(A, b), make_burg("l2", lamda), b.sum(), (1.0 / n) * np.ones(n)


def KL_nonneg_regr(m, n, noise=0.01, lamdaL1=0, randseed=-1, normalizeA=True):
    """applications.py:175-206."""
    _seed(randseed)
    A = np.random.rand(m, n)
    if normalizeA:
        A = A / A.sum(axis=0)
    x = np.random.rand(n)
    b = np.dot(A, x) + noise * (np.random.rand(m) - 0.5)
    assert b.min() > 0, "need b > 0 for nonnegative regression."
    L = max(A.sum(axis=0))
    return make_kl(A, b), make_shannon("l1", lamdaL1), L, 0.5 * np.ones(n)


def D_opt_KYinit(V):
    """Kumar-Yildirim sparse start.  applications.py:59-95 (draws m*rand(m) from the global RNG)."""
    m, n = V.shape
    if n <= 2 * m:
        return (1.0 / n) * np.ones(n)
    picked = []
    Q = np.zeros((m, m))
    for i in range(m):
        b = np.random.rand(m)
        q = np.copy(b)
        for j in range(i):
            q = q - np.dot(Q[:, j], b) * Q[:, j]
        qV = np.dot(q, V)
        kmax = np.argmax(qV)
        kmin = np.argmin(qV)
        picked.append(kmax)
        picked.append(kmin)
        v = V[:, kmin] - V[:, kmax]
        q = np.copy(v)
        for j in range(i):
            q = q - np.dot(Q[:, j], v) * Q[:, j]
        Q[:, i] = q / np.linalg.norm(q)
    x0 = np.zeros(n)
    x0[picked] = np.ones(len(picked)) / len(picked)
    x0 /= x0.sum()
    return x0
