"""GPU probe (run under gpurun): FP64 GEMM roof, per-kernel timings of the hot path at benchmark shapes.
Writes JSON to gpurun_out/probe.json.  Not part of the product or the tests."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accbpg_and_fw_b200 as acc
from accbpg_and_fw_b200 import _native as nat

lib = nat.lib
out = {}
dev = torch.device("cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


# FP64 roof: cuBLAS DGEMM
N = 8192
A = torch.randn(N, N, dtype=torch.float64, device=dev)
B = torch.randn(N, N, dtype=torch.float64, device=dev)
med, best = timeit(lambda: torch.matmul(A, B), iters=5, warm=2)
out["dgemm_8192_tflops_best"] = 2 * N ** 3 / best / 1e9
out["dgemm_8192_tflops_median"] = 2 * N ** 3 / med / 1e9
t0 = time.time(); cnt = 0
torch.cuda.synchronize()
while time.time() - t0 < 3.0:
    torch.matmul(A, B); cnt += 1
    if cnt % 4 == 0:
        torch.cuda.synchronize()
torch.cuda.synchronize()
out["dgemm_8192_tflops_sustained"] = cnt * 2 * N ** 3 / (time.time() - t0) / 1e12
del A, B
print("dgemm", out, flush=True)

rt = acc.Runtime.get()
for (m, n) in [(500, 50000), (2000, 200000), (80, 200)]:
    g0 = torch.Generator(device="cuda").manual_seed(1)
    H = torch.randn(m, n, dtype=torch.float64, device=dev, generator=g0)
    x = torch.full((n,), 1.0 / n, dtype=torch.float64, device=dev)
    f = acc.DOptimalObj(H)
    ws = f._ws.data_ptr()
    M = torch.empty(m, m, dtype=torch.float64, device=dev)
    L = torch.empty(m, m, dtype=torch.float64, device=dev)
    g = torch.empty(n, dtype=torch.float64, device=dev)
    tag = f"dopt_{m}x{n}"
    t_gram = timeit(lambda: nat.check(lib.accbpg_dopt_gram(rt.ctx, rt.stream, H.data_ptr(), m, n, n, x.data_ptr(), ws, M.data_ptr())))
    t_fac = timeit(lambda: nat.check(lib.accbpg_dopt_factor(rt.ctx, rt.stream, m, M.data_ptr(), L.data_ptr(), 1, ws, rt.slot(40))))
    t_grad = timeit(lambda: nat.check(lib.accbpg_dopt_grad(rt.ctx, rt.stream, H.data_ptr(), m, n, n, ws, g.data_ptr())))
    t_all = timeit(lambda: f._enqueue(x, 2, 40, g))
    out[tag] = {"gram_ms": t_gram, "factor_ms": t_fac, "grad_ms": t_grad, "func_grad_ms": t_all,
                "gram_tflops": m * m * n / t_gram[0] / 1e9, "grad_tflops": m * m * n / t_grad[0] / 1e9}
    print(tag, out[tag], flush=True)
    h = acc.BurgEntropySimplex()
    o = torch.empty(n, dtype=torch.float64, device=dev)
    gg = torch.randn(n, dtype=torch.float64, device=dev)
    t_prox = timeit(lambda: h._enq_div_prox(x, gg, 0.5, o))
    t_div = timeit(lambda: h._enq_divergence(o, x, 41))
    out[tag]["burg_simplex_divprox_ms"] = t_prox
    out[tag]["burg_div_ms"] = t_div
    print(tag, "prox", t_prox, "div", t_div, flush=True)
    if m <= 500:
        t0 = time.time()
        xx, F, SP, SN, T = acc.D_opt_FW_away(H, x, 1e-12, 2000, verbose=False)
        torch.cuda.synchronize()
        el = time.time() - t0
        out[tag]["fw_away_it_per_s_wall"] = len(F) / el
        out[tag]["fw_away_it_per_s_devclock"] = (len(T) - 1) / (T[-1] - T[0]) if len(T) > 1 else None
        print(tag, "fw_away", out[tag]["fw_away_it_per_s_wall"], out[tag]["fw_away_it_per_s_devclock"], flush=True)
    del H, f, M, L, g
    torch.cuda.empty_cache()

# n = 1e6 Burg simplex prox
n = 1000000
y = torch.full((n,), 1.0 / n, dtype=torch.float64, device=dev)
gg = torch.randn(n, dtype=torch.float64, device=dev)
o = torch.empty(n, dtype=torch.float64, device=dev)
h = acc.BurgEntropySimplex()
out["burg_simplex_divprox_1e6_ms"] = timeit(lambda: h._enq_div_prox(y, gg, 0.5, o))
print("burg 1e6", out["burg_simplex_divprox_1e6_ms"], flush=True)

# GEMV pair bandwidth
m, n = 8192, 262144
A = torch.rand(m, n, dtype=torch.float64, device=dev)
b = torch.rand(m, dtype=torch.float64, device=dev) + 0.5
x = torch.rand(n, dtype=torch.float64, device=dev)
f = acc.KLdivRegression(A, b)
g = torch.empty(n, dtype=torch.float64, device=dev)
t_mv = timeit(lambda: nat.check(lib.accbpg_linreg_matvec(rt.ctx, rt.stream, A.data_ptr(), m, n, n, x.data_ptr(), f._ws.data_ptr(), f._Ax.data_ptr())))
t_rmv = timeit(lambda: nat.check(lib.accbpg_linreg_rmatvec(rt.ctx, rt.stream, A.data_ptr(), m, n, n, f._Ax.data_ptr(), f._ws.data_ptr(), g.data_ptr())))
t_fg = timeit(lambda: f._enqueue(x, 2, 40, g))
out["linreg_8192x262144"] = {"matvec_ms": t_mv, "rmatvec_ms": t_rmv, "func_grad_ms": t_fg,
                             "matvec_GBs": 8.0 * m * n / t_mv[0] / 1e6, "rmatvec_GBs": 8.0 * m * n / t_rmv[0] / 1e6}
print(out["linreg_8192x262144"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1, default=str)
