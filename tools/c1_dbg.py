import sys; sys.path.insert(0, ".")
import accbpg_and_fw_b200 as acc
from accbpg_and_fw_b200 import drivers
f, h, L, x0 = acc.D_opt_design(80, 200, randseed=10)
try:
    acc.BPG(f, h, L, x0, maxitrs=20, linesearch=False, verbose=False)
except Exception as e:
    print("raised", type(e).__name__)
inf = drivers._bpg_small.last_info if hasattr(drivers._bpg_small, "last_info") else None
print(inf)
