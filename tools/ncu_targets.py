"""Small driver for ncu captures of the HBM-bound kernels (run under gpurun + ncu; not part of the tests):
KL func_grad at 8192 x 262144 (GEMV pair), D_opt_FW_away at 500 x 50000 (pass over V), Burg-simplex prox at n = 10^6."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accbpg_and_fw_b200 as acc      # noqa: E402

dev = torch.device("cuda")
gen = torch.Generator(device=dev)
gen.manual_seed(7)
m, n = 8192, 262144
A = torch.rand(m, n, dtype=torch.float64, device=dev, generator=gen)
b = torch.rand(m, dtype=torch.float64, device=dev, generator=gen) + 0.5
x = torch.rand(n, dtype=torch.float64, device=dev, generator=gen) / n
f = acc.KLdivRegression(A, b)
for _ in range(3):
    f.func_grad(x)
del f, A
torch.cuda.empty_cache()
fd, h, L, x0 = acc.D_opt_design(500, 50000, randseed=1)
acc.D_opt_FW_away(fd._Hd, torch.tensor(x0, device=dev), 1e-12, 12, verbose=False)
hs = acc.BurgEntropySimplex()
g = torch.randn(1000000, dtype=torch.float64, device=dev, generator=gen)
y = torch.full((1000000,), 1e-6, dtype=torch.float64, device=dev)
for _ in range(2):
    hs.div_prox_map(y, g, 0.5)
print("ok")
