"""C1 (D_opt_design(80, 200), BPG with line search, 1000 iterations): fused single-CTA solve against the operator path."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import accbpg_and_fw_b200 as acc
from accbpg_and_fw_b200 import config

f, h, L, x0 = acc.D_opt_design(80, 200, randseed=10)
for fused in (True, False, True):
    config.fused_small = fused
    acc.BPG(f, h, L, x0, maxitrs=50, verbose=False)
    t = time.time()
    x, F, Ls, T = acc.BPG(f, h, L, x0, maxitrs=1000, linesearch=True, ls_ratio=1.2, verbose=False)
    dt = time.time() - t
    print(f"fused={fused}: {len(F)} iterations in {dt * 1e3:.1f} ms = {len(F) / dt:.0f} it/s; F_last {F[-1]:.12f} L_last {Ls[-1]:.6f}")
    if fused:
        from accbpg_and_fw_b200 import drivers
        inf = drivers._bpg_small.last_info
        names = ["diag inverses", "gradient", "prox", "div+dot", "gram", "factor"]
        print("   trials %d newton %d; clocks per iteration: " % (inf[1], inf[3]) +
              ", ".join(f"{nm} {inf[4 + i] / len(F):.0f}" for i, nm in enumerate(names)))
        print("   inside one factorisation: " + ", ".join(f"{nm} {inf[10 + i] / (inf[1] + 1):.0f}" for i, nm in
                                                        enumerate(["diagonal tile", "tile inverse", "panel", "trailing"])))
