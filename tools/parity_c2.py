"""Full-size parity record for BASELINE.json configs[1] (run under gpurun): ABPG gamma=2 and ABPG_gain on
D_opt_design(500, 50000, randseed=1), GPU path against the CPU oracle port, objective trajectory F_k.
CPU time dominates (about 0.6 s per oracle iteration on 16 cores).  Writes gpurun_out/parity_c2.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import accbpg_and_fw_b200 as acc                      # noqa: E402
from oracle import accbpg_oracle as orc                # noqa: E402

K1, K2, K3 = int(sys.argv[1]) if len(sys.argv) > 1 else 200, 60, 300
f, h, L, x0 = acc.D_opt_design(500, 50000, randseed=1)
fo, ho = orc.make_dopt(f.H), orc.make_burg("simplex")
out = {}
t = time.time()
xg, Fg, Gg, Tg = acc.ABPG(f, h, L, x0, gamma=2, maxitrs=K1, verbose=False)
xo, Fo, Go, To = orc.ABPG(fo, ho, L, x0, gamma=2, maxitrs=K1, theta_eq=False)
n = min(len(Fg), len(Fo))
out["abpg"] = {"iterations": n, "max_rel_err_F": float(np.max(np.abs(Fg[:n] - Fo[:n]) / np.abs(Fo[:n]))),
               "max_rel_err_x": float(np.max(np.abs(xg - xo) / np.maximum(np.abs(xo), 1e-300))),
               "F_last": [float(Fg[n - 1]), float(Fo[n - 1])]}
print("abpg", out["abpg"], time.time() - t, flush=True)
rg = acc.ABPG_gain(f, h, L, x0, gamma=2, maxitrs=K2, verbose=False)
ro = orc.ABPG_gain(fo, ho, L, x0, gamma=2, maxitrs=K2)
n = min(len(rg[1]), len(ro[1]))
fork = int(np.argmax(rg[2][:n] != ro[2][:n])) if np.any(rg[2][:n] != ro[2][:n]) else n
out["abpg_gain"] = {"iterations": n, "max_rel_err_F": float(np.max(np.abs(rg[1][:n] - ro[1][:n]) / np.abs(ro[1][:n]))),
                    "first_gain_fork": fork}
print("abpg_gain", out["abpg_gain"], flush=True)
ig, io = [], []
xa, Fa, SPa, SNa, Ta = acc.D_opt_FW_away(f.H, x0, 1e-8, K3, verbose=False)
xb, Fb, SPb, SNb, Tb = orc.D_opt_FW_away(f.H, x0, 1e-8, K3)
n = min(len(Fa), len(Fb))
out["fw_away"] = {"iterations": n, "max_rel_err_F": float(np.max(np.abs(Fa[:n] - Fb[:n]) / np.abs(Fb[:n]))),
                  "max_abs_err_x": float(np.max(np.abs(xa - xb)))}
print("fw_away", out["fw_away"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_c2.json"), "w"), indent=1)
