"""Small end-to-end exercise of every kernel family for compute-sanitizer (run under gpurun; not part of the tests)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accbpg_and_fw_b200 as acc      # noqa: E402

for (m, n) in [(150, 1030), (257, 2050), (64, 200)]:
    f, h, L, x0 = acc.D_opt_design(m, n, randseed=3)
    fx, g = f.func_grad(x0)
    x, F, G, T = acc.ABPG(f, h, L, x0, gamma=2, maxitrs=4, verbose=False)
    out = acc.ABPG_gain(f, h, L, x0, gamma=2, maxitrs=3, verbose=False)
    xa, Fa, SPa, SNa, Ta = acc.D_opt_FW_away(f.H, x0, 1e-8, 6, verbose=False)
    print(m, n, fx, F[-1], Fa[-1])
fk, hk, Lk, xk = acc.KL_nonneg_regr(120, 300, noise=0.01, lamdaL1=0.001, randseed=1)
print(acc.BPG(fk, hk, Lk, xk, maxitrs=5, verbose=False)[1][-1])
fp, hp, Lp, xp = acc.Poisson_regrL2(90, 70, noise=1e-3, lamda=1e-3, randseed=1)
print(acc.ABPG_gain(fp, hp, Lp, xp, gamma=2.0, maxitrs=5, verbose=False)[1][-1])
hs = acc.ShannonEntropySimplex()
print(float(hs.div_prox_map(np.ones(50) / 50, np.linspace(-1, 1, 50), 0.7).sum()))
print("sanitize ok")
