"""Data-flow Cholesky chain against the launch-per-column chain and torch (run under gpurun).
    python tools/chol_df_check.py [m ...]"""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import accbpg_and_fw_b200 as acc
from accbpg_and_fw_b200 import _native as nat
from accbpg_and_fw_b200.runtime import Runtime
lib = nat.lib
rt = Runtime.get()
dev = rt.device
ms = [int(a) for a in sys.argv[1:]] or [65, 127, 128, 200, 500, 501, 777, 1000, 2000]
for m in ms:
    g = torch.Generator(device=dev); g.manual_seed(m)
    A = torch.randn(m, m + 37, dtype=torch.float64, device=dev, generator=g)
    M = (A @ A.T) / (m + 37) + 0.05 * torch.eye(m, dtype=torch.float64, device=dev)
    ws = torch.zeros(lib.accbpg_dopt_workspace_bytes(m, m), dtype=torch.uint8, device=dev)
    L = torch.empty(m, m, dtype=torch.float64, device=dev)
    for want in (1, 0):
        for rep in range(3):
            nat.check(lib.accbpg_dopt_factor(rt.ctx, rt.stream, m, M.data_ptr(), L.data_ptr(), want, ws.data_ptr(), rt.slot(40)))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 20
        for _ in range(reps):
            nat.check(lib.accbpg_dopt_factor(rt.ctx, rt.stream, m, M.data_ptr(), None, want, ws.data_ptr(), rt.slot(40)))
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / reps * 1e3
        f = rt.read_raw(40, 1)[0][0]
        Lt = torch.linalg.cholesky(M)
        fref = -2 * torch.log(torch.diagonal(Lt)).sum().item()
        eL = (L - Lt).abs().max().item() / Lt.abs().max().item()
        msg = f"m {m:5d} want_inv {want}: {us:8.1f} us  f rel err {abs(f - fref) / abs(fref):.1e}  L err {eL:.1e}"
        if want:
            mp = (m + 127) // 128 * 128
            # Linv lives in the workspace: recover through accbpg_dopt_grad on H = I (g_j = -||Linv e_j||^2)
            H = torch.eye(m, dtype=torch.float64, device=dev).contiguous()
            ws2 = ws
            gvec = torch.empty(m, dtype=torch.float64, device=dev)
            nat.check(lib.accbpg_dopt_grad(rt.ctx, rt.stream, H.data_ptr(), m, m, m, ws2.data_ptr(), gvec.data_ptr()))
            Li = torch.linalg.inv(Lt)
            gref = -(Li * Li).sum(dim=0)
            msg += f"  colnorm(Linv) err {((gvec - gref).abs() / gref.abs()).max().item():.1e}"
        print(msg, flush=True)
