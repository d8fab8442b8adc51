import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
import accbpg_and_fw_b200 as acc
from accbpg_and_fw_b200 import config
f, h, L, x0 = acc.D_opt_design(500, 50000, randseed=1)
x0d = torch.tensor(x0, device="cuda")
res = {}
for mode in (True, False):
    config.linear_images = mode
    res[mode] = acc.ABPG_gain(f, h, L, x0d, gamma=2, maxitrs=400, verbose=False)
    ra = acc.ABPG(f, h, L, x0d, gamma=2, maxitrs=400, verbose=False)
    res[(mode, "abpg")] = ra
a, b = res[True], res[False]
n = min(len(a[1]), len(b[1]))
d = np.abs(a[1][:n] - b[1][:n]) / np.abs(b[1][:n])
fork = int(np.argmax(a[2][:n] != b[2][:n])) if np.any(a[2][:n] != b[2][:n]) else n
print("ABPG_gain lin on vs off: n", n, "first gain fork", fork, "max rel dF up to fork", d[:max(fork,1)].max(), "at 50/100/200:", d[50], d[100], d[min(200,n-1)])
a, b = res[(True, "abpg")], res[(False, "abpg")]
n = min(len(a[1]), len(b[1]))
d = np.abs(a[1][:n] - b[1][:n]) / np.abs(b[1][:n])
print("ABPG lin on vs off: n", n, "max rel dF", d.max(), "at 50/100/200/399:", d[50], d[100], d[200], d[n-1])
