"""ABPG_gain at the benchmark shape (500 x 50000): wall time per iteration against the sum of kernel time, to see how much
of an iteration the GPU waits for the host's line-search decisions."""
import ctypes
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import accbpg_and_fw_b200 as acc
from accbpg_and_fw_b200 import _native as nat

lib = nat.lib


def prof_read():
    out = {}
    tot, cnt = ctypes.c_double(0), ctypes.c_int64(0)
    for i in range(lib.accbpg_prof_count()):
        nat.check(lib.accbpg_prof_read(i, ctypes.byref(tot), ctypes.byref(cnt)))
        if cnt.value:
            out[lib.accbpg_prof_name(i).decode()] = (tot.value, cnt.value)
    return out


f, h, L, x0 = acc.D_opt_design(500, 50000, randseed=1)
x0d = torch.as_tensor(x0).cuda()
its = 60
for name, run in (("ABPG", lambda: acc.ABPG(f, h, L, x0d, gamma=2, maxitrs=its, verbose=False)),
                  ("ABPG_gain", lambda: acc.ABPG_gain(f, h, L, x0d, gamma=2, maxitrs=its, verbose=False))):
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = run()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / its * 1e3
    lib.accbpg_prof_enable(1)
    prof_read()
    t0 = time.perf_counter()
    r = run()
    torch.cuda.synchronize()
    wallp = (time.perf_counter() - t0) / its * 1e3
    pr = prof_read()
    lib.accbpg_prof_enable(0)
    print(f"{name}: {wall:.3f} ms per iteration ({wallp:.3f} with per-kernel events)")
    tot = 0.0
    for k, (ms, cnt) in sorted(pr.items(), key=lambda kv: -kv[1][0]):
        print(f"   {k:70s} {ms / its:8.4f} ms/it  {cnt / its:5.2f} launches/it  {ms / cnt * 1e3:8.1f} us each")
        if "interval" not in k and "all block columns" not in k:
            tot += ms / its
    print(f"   sum of kernel time (intervals counted once): {tot:.3f} ms/it")
