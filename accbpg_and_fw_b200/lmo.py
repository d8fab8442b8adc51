"""Linear minimisation oracles of accbpg/functions_lmo.py as GPU-resident drop-ins.

Each factory returns a callable g -> s, exactly like the reference.  The returned callables
accept NumPy arrays (NumPy out) or CUDA tensors (CUDA out).  `lmo_simplex` is the one the
benchmark configurations use; its callable also exposes `.last_index(sync)` and an
asynchronous `_enq(gd, out, slot)` form for the drivers.  `lmo_nuclear_norm_ball` replaces the
reference's full SVD by a power iteration for the leading singular pair (two GEMVs per step).
"""
import numpy as np
import torch

from . import _native as nat
from .runtime import Runtime, is_host, like_input

lib = nat.lib


class _SimplexLMO:
    """s = 1e-15 everywhere, s[first argmin g] = radius.   functions_lmo.py:137-160."""

    def __init__(self, radius=1, shard=None, device=None):
        self.radius = float(radius)
        self.shard = shard
        self._device = device

    @property
    def rt(self):
        return Runtime.get(self._device)

    def _enq(self, gd, out, slot):
        """Device form: writes s into `out`, (min g, global index) into scal[slot:slot+2]."""
        rt = self.rt
        n = gd.numel()
        if self.shard is None or self.shard.world == 1:
            nat.check(lib.accbpg_lmo_simplex(rt.ctx, rt.stream, n, gd.data_ptr(), self.radius, out.data_ptr(),
                                             rt.slot(slot)))
            return
        sh = self.shard
        nat.check(lib.accbpg_vec_argext(rt.ctx, rt.stream, n, gd.data_ptr(), 0, rt.slot(slot)))
        pair = rt.scal[slot:slot + 2]
        pair[1] += float(sh.lo)                                  # local -> global column index
        sh.argmin_(pair)
        gidx = int(rt.read(slot, 2)[1])
        local = gidx - sh.lo if sh.lo <= gidx < sh.hi else -1    # -1: the vertex lives on another rank
        nat.check(lib.accbpg_lmo_fill_vertex(rt.ctx, rt.stream, n, 1e-15, local, self.radius, out.data_ptr()))

    def __call__(self, g):
        rt = self.rt
        host = is_host(g)
        gd = rt.to_device(g).reshape(-1)
        out = rt.empty(gd.numel())
        self._enq(gd, out, rt.S_TMP)
        res = like_input(out, host)
        return res.reshape(g.shape) if hasattr(g, "shape") else res

    def last_index(self):
        """Global index of the vertex chosen by the most recent synchronous call."""
        return int(self.rt.read(self.rt.S_TMP, 2)[1])


def lmo_simplex(radius=1, shard=None, device=None):
    return _SimplexLMO(radius, shard, device)


def lmo_matrix_simplex(radius=1.0, device=None):
    """functions_lmo.py:163-187: argmin over the flattened matrix, 1e-60 elsewhere."""
    def f(G):
        rt = Runtime.get(device)
        host = is_host(G)
        gd = rt.to_device(G).reshape(-1)
        out = rt.empty(gd.numel())
        nat.check(lib.accbpg_vec_argext(rt.ctx, rt.stream, gd.numel(), gd.data_ptr(), 0, rt.slot(rt.S_TMP)))
        idx = int(rt.read(rt.S_TMP, 2)[1])
        nat.check(lib.accbpg_lmo_fill_vertex(rt.ctx, rt.stream, gd.numel(), 1e-60, idx, float(radius),
                                             out.data_ptr()))
        return like_input(out, host).reshape(G.shape)
    return f


def lmo_linf_ball(radius, center=None, device=None):
    """functions_lmo.py:106-134: center - radius*sign(g)."""
    def f(g):
        rt = Runtime.get(device)
        host = is_host(g)
        gd = rt.to_device(g).reshape(-1)
        cd = rt.to_device(np.broadcast_to(center, g.shape) if is_host(center) and center is not None else center
                          ).reshape(-1) if center is not None else None
        out = rt.empty(gd.numel())
        nat.check(lib.accbpg_lmo_linf(rt.ctx, rt.stream, gd.numel(), gd.data_ptr(), float(radius),
                                      cd.data_ptr() if cd is not None else None, out.data_ptr()))
        return like_input(out, host).reshape(g.shape)
    return lambda g: f(g)


def lmo_l2_ball(radius, center=None, device=None):
    """functions_lmo.py:16-51: center - radius*g/||g||, or center itself when ||g|| < 1e-10."""
    def f(g):
        rt = Runtime.get(device)
        host = is_host(g)
        gd = rt.to_device(g).reshape(-1)
        n = gd.numel()
        cd = None
        if center is not None:
            cd = rt.to_device(np.broadcast_to(center, g.shape) if is_host(center) else center).reshape(-1)
        nat.check(lib.accbpg_vec_dot(rt.ctx, rt.stream, n, gd.data_ptr(), gd.data_ptr(), rt.slot(rt.S_TMP)))
        gnorm = float(np.sqrt(rt.read(rt.S_TMP, 1)[0]))
        if gnorm < 1e-10:
            res = cd.clone() if cd is not None else torch.zeros_like(gd)
            return like_input(res, host).reshape(g.shape)
        out = rt.empty(n)
        nat.check(lib.accbpg_lmo_l2(rt.ctx, rt.stream, n, gd.data_ptr(), float(radius), gnorm,
                                    cd.data_ptr() if cd is not None else None, out.data_ptr()))
        return like_input(out, host).reshape(g.shape)
    return lambda g: f(g)


def lmo_matrix_box(lower, upper, device=None):
    """functions_lmo.py:190-212: where(G < 0, upper, lower)."""
    def f(G):
        rt = Runtime.get(device)
        host = is_host(G)
        gd = rt.to_device(G).reshape(-1)
        lo = rt.to_device(np.broadcast_to(lower, G.shape) if is_host(lower) else lower).reshape(-1)
        hi = rt.to_device(np.broadcast_to(upper, G.shape) if is_host(upper) else upper).reshape(-1)
        out = rt.empty(gd.numel())
        nat.check(lib.accbpg_lmo_box(rt.ctx, rt.stream, gd.numel(), gd.data_ptr(), lo.data_ptr(), hi.data_ptr(),
                                     out.data_ptr()))
        return like_input(out, host).reshape(G.shape)
    return f


def lmo_l2_ball_positive_orthant(radius, center=None, epsilon=0.0, device=None):
    """functions_lmo.py:54-102: move radius along -g restricted to the coordinates with g < 0, floor at epsilon.

    Composed from the device primitives (masked norm by a dot product of the negative part)."""
    def f(g):
        rt = Runtime.get(device)
        host = is_host(g)
        gd = rt.to_device(g).reshape(-1)
        cd = rt.to_device(center).reshape(-1) if center is not None else torch.zeros_like(gd)
        assert cd.shape == gd.shape, "Shape mismatch between g and center"
        gneg = torch.clamp(gd, max=0.0)                       # g where g < 0, else 0
        n = gd.numel()
        nat.check(lib.accbpg_vec_dot(rt.ctx, rt.stream, n, gneg.data_ptr(), gneg.data_ptr(), rt.slot(rt.S_TMP)))
        nrm = float(np.sqrt(rt.read(rt.S_TMP, 1)[0]))
        if nrm == 0.0:
            res = torch.clamp(cd, min=float(epsilon))
            return like_input(res, host).reshape(g.shape)
        out = rt.empty(n)
        nat.check(lib.accbpg_lmo_l2(rt.ctx, rt.stream, n, gneg.data_ptr(), float(radius), nrm, cd.data_ptr(),
                                    out.data_ptr()))
        res = torch.clamp(out, min=float(epsilon))
        return like_input(res, host).reshape(g.shape)
    return f


def lmo_nuclear_norm_ball(device=None, tol=1e-14, maxit=50000):
    """Rank-one vertex u1 v1^T of the leading singular pair of g.   functions_lmo.py:4-13 (which takes a full SVD and
    returns np.outer(U[:, 0], Vh[0]) -- no sign flip, no radius; mirrored as is).

    Here: power iteration v <- A^T(A v)/||.|| on the device (K5/K6 GEMV kernels), stopped when ||v_new - v||^2 <= tol^2
    (the host reads two scalars per step), then u = A v/||A v||.  u v^T does not depend on the sign of the pair, so it
    matches the SVD's answer whenever sigma_1 is simple; convergence is geometric with ratio (sigma_2/sigma_1)^2."""
    def f(G):
        rt = Runtime.get(device)
        host = is_host(G)
        A = rt.to_device(G)
        assert A.dim() == 2, "lmo_nuclear_norm_ball takes a matrix"
        p, q = int(A.shape[0]), int(A.shape[1])
        ws = rt.workspace(("linreg", p, q), lib.accbpg_linreg_workspace_bytes(p, q))
        # deterministic start with components along every right singular vector
        v = torch.tensor(1.0 + 0.37 * np.cos(1.7 * np.arange(q)), dtype=torch.float64, device=rt.device)
        v /= float(np.sqrt(float((v * v).sum().item())))
        u = rt.empty(p)
        v2 = rt.empty(q)
        s0 = rt.S_TMP
        for it in range(maxit):
            nat.check(lib.accbpg_linreg_matvec(rt.ctx, rt.stream, A.data_ptr(), p, q, A.stride(0), v.data_ptr(),
                                               ws.data_ptr(), u.data_ptr()))
            nat.check(lib.accbpg_linreg_rmatvec(rt.ctx, rt.stream, A.data_ptr(), p, q, A.stride(0), u.data_ptr(),
                                                ws.data_ptr(), v2.data_ptr()))
            nat.check(lib.accbpg_vec_dot(rt.ctx, rt.stream, q, v2.data_ptr(), v2.data_ptr(), rt.slot(s0)))
            nrm = float(np.sqrt(rt.read(s0, 1)[0]))
            if nrm == 0.0:
                break
            vn = rt.empty(q)
            nat.check(lib.accbpg_vec_divide(rt.ctx, rt.stream, q, v2.data_ptr(), nrm, vn.data_ptr()))
            nat.check(lib.accbpg_vec_sqdist(rt.ctx, rt.stream, q, vn.data_ptr(), v.data_ptr(), rt.slot(s0)))
            diff2 = rt.read(s0, 1)[0]
            v = vn
            if diff2 <= tol * tol:
                break
        nat.check(lib.accbpg_linreg_matvec(rt.ctx, rt.stream, A.data_ptr(), p, q, A.stride(0), v.data_ptr(),
                                           ws.data_ptr(), u.data_ptr()))
        nat.check(lib.accbpg_vec_dot(rt.ctx, rt.stream, p, u.data_ptr(), u.data_ptr(), rt.slot(s0)))
        un = float(np.sqrt(rt.read(s0, 1)[0]))
        uu = rt.empty(p)
        nat.check(lib.accbpg_vec_divide(rt.ctx, rt.stream, p, u.data_ptr(), un if un > 0 else 1.0, uu.data_ptr()))
        out = torch.empty(p, q, dtype=torch.float64, device=rt.device)
        nat.check(lib.accbpg_mat_outer(rt.ctx, rt.stream, p, uu.data_ptr(), q, v.data_ptr(), out.data_ptr()))
        return like_input(out, host)

    return lambda g: f(g)
