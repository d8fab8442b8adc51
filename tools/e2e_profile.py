"""Where the host-vector (e2e) ABPG iteration of bench.py spends its time: per public call kind, wall clock with a
synchronise on both sides, then a cProfile of the same loop.  python tools/e2e_profile.py [iters]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import accbpg_and_fw_b200 as acc      # noqa: E402
import bench                           # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Hh = bench.make_slab(0, bench.N_PER_GPU)
f = acc.DOptimalObj(Hh)
h = acc.BurgEntropySimplex()
n = bench.N_PER_GPU
x0 = np.ones(n) / n
bench.host_vector_abpg(f, h, 1.0, x0, 2.0, 5)
acc_t = {}


def timed(name, fn, *a):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn(*a)
    torch.cuda.synchronize()
    acc_t[name] = acc_t.get(name, 0.0) + time.perf_counter() - t0
    return r


x, z = x0.copy(), x0.copy()
t_all = time.perf_counter()
for k in range(iters):
    fx = timed("f(x)", f, x) + timed("extra_Psi", h.extra_Psi, x)
    theta = 2.0 / (k + 2.0)
    t0 = time.perf_counter()
    y = (1 - theta) * x + theta * z
    acc_t["host axpby"] = acc_t.get("host axpby", 0.0) + time.perf_counter() - t0
    g = timed("f.gradient(y)", f.gradient, y)
    z1 = timed("div_prox_map", h.div_prox_map, z, g, theta * 1.0)
    t0 = time.perf_counter()
    x = (1 - theta) * x + theta * z1
    acc_t["host axpby"] += time.perf_counter() - t0
    timed("divergence x2", lambda: (h.divergence(x, y), h.divergence(z1, z)))
    z = z1
t_all = time.perf_counter() - t_all
print(f"iteration {t_all / iters * 1e3:.3f} ms (with the extra synchronises)")
for k, v in acc_t.items():
    print(f"  {k:16s} {v / iters * 1e3:8.3f} ms")
pr = cProfile.Profile()
pr.enable()
bench.host_vector_abpg(f, h, 1.0, x0, 2.0, iters)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
