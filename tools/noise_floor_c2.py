"""The reference path's own sensitivity at BASELINE.json configs[1] (CPU only): ABPG_gain and ABPG on
D_opt_design(500, 50000, randseed=1), oracle vs the same oracle with every operator output perturbed by +-1 ulp.
Shows when two correct FP64 implementations of this run stop agreeing to 1e-9.  Writes profiles/noise_floor_c2_r01.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import accbpg_oracle as orc          # noqa: E402
from test_noise_floor import _noisy              # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 140
f, h, L, x0 = orc.D_opt_design(500, 50000, randseed=1)
out = {}
base = orc.ABPG_gain(f, h, L, x0, gamma=2, maxitrs=K)
nf, nh = _noisy(f, h, np.random.RandomState(0))
pert = orc.ABPG_gain(nf, nh, L, x0, gamma=2, maxitrs=K)
n = min(len(base[1]), len(pert[1]))
d = np.abs(base[1][:n] - pert[1][:n]) / np.abs(base[1][:n])
fork = int(np.argmax(base[2][:n] != pert[2][:n])) if np.any(base[2][:n] != pert[2][:n]) else n
cross = int(np.argmax(d > 1e-9)) if np.any(d > 1e-9) else n
out["abpg_gain"] = {"iterations": n, "first_gain_fork": fork, "first_k_with_dF_above_1e-9": cross,
                    "dF_at": {str(k): float(d[k]) for k in range(0, n, 10)}}
print(out["abpg_gain"], flush=True)
base = orc.ABPG(f, h, L, x0, gamma=2, maxitrs=K, theta_eq=False)
pert = orc.ABPG(nf, nh, L, x0, gamma=2, maxitrs=K, theta_eq=False)
n = min(len(base[1]), len(pert[1]))
d = np.abs(base[1][:n] - pert[1][:n]) / np.abs(base[1][:n])
out["abpg"] = {"iterations": n, "max_dF": float(d.max())}
print(out["abpg"], flush=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "noise_floor_c2_r01.json"), "w"), indent=1)
