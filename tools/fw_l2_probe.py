"""D_opt_FW_away iterations/s at 500x50000 (or argv m n) - used to tune the L2 access-policy window of the pass kernel
(ACCBPG_FW_L2_MB / ACCBPG_FW_L2_HIT are read once per process, so run one process per setting)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import accbpg_and_fw_b200 as acc      # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 500
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
its = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
gen = torch.Generator(device="cuda")
gen.manual_seed(1)
V = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=gen)
x0 = torch.full((n,), 1.0 / n, dtype=torch.float64, device="cuda")
for fn in (acc.D_opt_FW_away, acc.D_opt_FW):
    fn(V, x0, 1e-12, 200, verbose=False)
    x, F, SP, SN, T = fn(V, x0, 1e-12, its, verbose=False)
    print(f"{fn.__name__} {m}x{n}: {(len(T) - 1) / (T[-1] - T[0]):.1f} it/s  F_last {F[-1]:.15e}  "
          f"L2_MB={os.environ.get('ACCBPG_FW_L2_MB', 'max')} HIT={os.environ.get('ACCBPG_FW_L2_HIT', 'auto')}", flush=True)
