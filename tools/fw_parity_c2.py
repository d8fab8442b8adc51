import sys, numpy as np, time
sys.path.insert(0, "/root/repo")
import accbpg_and_fw_b200 as acc
from oracle import accbpg_oracle as orc
f, h, L, x0 = acc.D_opt_design(500, 50000, randseed=1)
for K in (300, 1200):
    xa, Fa, SPa, SNa, Ta = acc.D_opt_FW_away(f.H, x0, 1e-8, K, verbose=False)
    xb, Fb, SPb, SNb, Tb = orc.D_opt_FW_away(f.H, x0, 1e-8, K)
    n = min(len(Fa), len(Fb))
    d = np.abs(Fa[:n] - Fb[:n]) / np.abs(Fb[:n])
    print(K, n, "max rel dF", d.max(), "at", int(d.argmax()), "x err", np.max(np.abs(xa - xb)), "it/s", (len(Ta)-1)/(Ta[-1]-Ta[0]))
