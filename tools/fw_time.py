import sys, time, numpy as np, torch, os
sys.path.insert(0, ".")
import accbpg_and_fw_b200 as acc
f,h,L,x0 = acc.D_opt_design(500,50000,randseed=1)
x0d = torch.tensor(x0, device="cuda")
xa,Fa,SPa,SNa,Ta = acc.D_opt_FW_away(f._Hd, x0d, 1e-12, 64, verbose=False)
for rep in range(2):
    xa,Fa,SPa,SNa,Ta = acc.D_opt_FW_away(f._Hd, x0d, 1e-12, 1500, verbose=False)
    print("persist", os.environ.get("ACCBPG_FW_PERSIST"), "it/s", (len(Ta)-1)/(Ta[-1]-Ta[0]), Fa[-1])
