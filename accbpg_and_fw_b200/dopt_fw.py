"""D_opt_FW and D_opt_FW_away (Wolfe-Atwood) with the whole iteration on the GPU.

Same call shape and return tuple as accbpg/D_opt_alg.py:9-88 and :91-185:
    x, F, SP, SN, T = D_opt_FW_away(V, x0, eps, maxitrs, verbose=True, verbskip=1)
V may be the host matrix `f.H` the reference notebooks pass, or a CUDA tensor.  The loop runs in
batches of `batch` iterations enqueued by one C call (five kernels per iteration, no host
round trip); the device stop flag is inspected once per batch, and the history arrays
(F, SP, SN, device-clock T) are copied back at the end.
"""
import time

import numpy as np
import torch

from . import _native as nat
from .runtime import Runtime, is_host, like_input

lib = nat.lib

# layout of the control block: enum FwCtrl in csrc/fw.cu
C_STOP, C_KSTOP, C_LOGDET_HI, C_LOGDET_LO, C_WMAX, C_IMAX, C_WMIN, C_JMIN = range(8)
C_MODE, C_T, C_CS, C_DEN, C_IDX, C_TSIGN, C_NITER = range(8, 15)
NCTRL = nat.MACROS["ACCBPG_FW_CTRL_DOUBLES"]


def _run(V, x0, eps, maxitrs, away, verbose, verbskip, batch, device, index_log):
    rt = Runtime.get(device)
    t_start = time.time()
    host = is_host(x0)
    Vd = rt.to_device(V)
    m, n = int(Vd.shape[0]), int(Vd.shape[1])
    x = rt.to_device(x0).clone()
    assert x.numel() == n, "x0.size not equal to the number of columns of V"
    dev = rt.device
    with rt.on_device():
        ws = rt.workspace(("fw", m, n), lib.accbpg_fw_workspace_bytes(m, n))
    Hinv = torch.empty(m, m, dtype=torch.float64, device=dev)
    w = torch.empty(n, dtype=torch.float64, device=dev)
    ctrl = torch.zeros(NCTRL, dtype=torch.float64, device=dev)
    hist = torch.zeros(4, max(int(maxitrs), 1), dtype=torch.float64, device=dev)
    nat.check(lib.accbpg_fw_setup(rt.ctx, rt.stream, Vd.data_ptr(), m, n, Vd.stride(0), x.data_ptr(),
                                  ws.data_ptr(), Hinv.data_ptr(), w.data_ptr(), ctrl.data_ptr()))
    rt.read(0, 0)                                   # x0 < 0 or singular V X V^T surface here
    if verbose:
        print("\nSolving D-opt design problem using Frank-Wolfe method" + (" with away steps" if away else ""))
        print("     k      F(x)     pos_slack   neg_slack    time")
    k = 0
    done = 0
    t_setup = time.time() - t_start
    ns0 = None
    while k < maxitrs:
        cnt = min(batch, maxitrs - k)
        if index_log is not None:
            cnt = 1
        nat.check(lib.accbpg_fw_run(rt.ctx, rt.stream, Vd.data_ptr(), m, n, Vd.stride(0), int(away), float(eps),
                                    k, cnt, ws.data_ptr(), Hinv.data_ptr(), x.data_ptr(), w.data_ptr(),
                                    ctrl.data_ptr(), hist[0].data_ptr(), hist[1].data_ptr(), hist[2].data_ptr(),
                                    hist[3].data_ptr()))
        c = ctrl.cpu().numpy()                      # one synchronising read per batch
        done = int(c[C_NITER])
        if index_log is not None and done == k + 1:
            # (argmax index, masked argmin index, step kind: 0 toward / 1 away / None on the stopping iteration)
            index_log.append((int(c[C_IMAX]), int(c[C_JMIN]), None if c[C_STOP] != 0 else int(c[C_MODE])))
        if verbose and done > k:
            hb = hist[:, k:done].cpu().numpy()
            if ns0 is None:
                ns0 = hb[3, 0]
            for kk in range(k, done):
                if kk % verbskip == 0:
                    print("{0:6d}  {1:10.3e}  {2:10.3e}  {3:10.3e}  {4:6.1f}".format(
                        kk, hb[0, kk - k], hb[1, kk - k], hb[2, kk - k], t_setup + (hb[3, kk - k] - ns0) * 1e-9))
        k += cnt
        if c[C_STOP] != 0:
            break
    hh = hist[:, :done].cpu().numpy()
    F, SP, SN = hh[0].copy(), hh[1].copy(), hh[2].copy()
    # T[k]: host seconds spent in setup + device clock (ns) since the first iteration's decision
    T = t_setup + (hh[3] - hh[3][0]) * 1e-9 if done > 0 else np.zeros(0)
    return like_input(x, host), F, SP, SN, T


def _run_sharded(V, x0, eps, maxitrs, away, verbose, verbskip, batch, device, index_log, shard):
    """Column-sharded loop: V is this rank's column slab, x0 its slice; Hinv, the control block and the histories are
    replicated.  Per iteration: decision from all ranks' selection records, all-reduce of the chosen column (only its
    owner contributes), u = Hinv v + local pass + rank-one update, all-gather of the new records.  No host round trip
    except one read of the control block per batch."""
    import torch.distributed as dist
    rt = Runtime.get(device)
    t_start = time.time()
    host = is_host(x0)
    Vd = rt.to_device(V)
    m, n = int(Vd.shape[0]), int(Vd.shape[1])
    assert n == shard.n_local, "V must be this rank's column slab (ColumnShard.cols)"
    x = rt.to_device(x0).clone()
    assert x.numel() == n, "x0 must be this rank's slice (ColumnShard.part)"
    dev = rt.device
    with rt.on_device():
        ws = rt.workspace(("fw", m, n), lib.accbpg_fw_workspace_bytes(m, n))
    Hinv = torch.empty(m, m, dtype=torch.float64, device=dev)
    M = torch.empty(m, m, dtype=torch.float64, device=dev)
    w = torch.empty(n, dtype=torch.float64, device=dev)
    vcol = torch.zeros(m, dtype=torch.float64, device=dev)
    ctrl = torch.zeros(NCTRL, dtype=torch.float64, device=dev)
    hist = torch.zeros(4, max(int(maxitrs), 1), dtype=torch.float64, device=dev)
    rec_doubles = lib.accbpg_fw_record_bytes() // 8
    rec = torch.zeros(rec_doubles, dtype=torch.float64, device=dev)
    recs = torch.zeros(shard.world * rec_doubles, dtype=torch.float64, device=dev)
    st = Vd.stride(0)
    nat.check(lib.accbpg_dopt_gram(rt.ctx, rt.stream, Vd.data_ptr(), m, n, st, x.data_ptr(), ws.data_ptr(), M.data_ptr()))
    shard.sum_(M)
    nat.check(lib.accbpg_fw_setup_from_gram(rt.ctx, rt.stream, Vd.data_ptr(), m, n, st, M.data_ptr(), ws.data_ptr(),
                                            Hinv.data_ptr(), w.data_ptr(), ctrl.data_ptr()))
    rt.read(0, 0)
    # both per-iteration exchanges over NVLink peer memory (accbpg_fw_run_peer) when symmetric buffers can be mapped
    from .dist import peer_buffers
    peer = peer_buffers(shard, dev, [(2 * shard.world * rec_doubles, torch.float64), (2 * m, torch.float64),
                                     (shard.world + 1, torch.int64)], cache_key="dopt_fw")
    if peer is None:
        nat.check(lib.accbpg_fw_select_local(rt.ctx, rt.stream, n, shard.lo, int(away), x.data_ptr(), w.data_ptr(),
                                             ws.data_ptr(), m, rec.data_ptr()))
        shard.all_gather_equal(recs, rec)
    if verbose and shard.rank == 0:
        print("\nSolving D-opt design problem using Frank-Wolfe method" + (" with away steps" if away else ""))
        print("     k      F(x)     pos_slack   neg_slack    time")
    t_setup = time.time() - t_start
    k = 0
    done = 0
    while k < maxitrs:
        cnt = 1 if index_log is not None else min(batch, maxitrs - k)
        if peer is not None:
            nat.check(lib.accbpg_fw_run_peer(rt.ctx, rt.stream, Vd.data_ptr(), m, n, st, shard.lo, int(away), float(eps),
                                             k, cnt, shard.rank, shard.world, peer.tables[0], peer.tables[1], peer.tables[2],
                                             ws.data_ptr(), Hinv.data_ptr(), x.data_ptr(), w.data_ptr(), ctrl.data_ptr(),
                                             hist[0].data_ptr(), hist[1].data_ptr(), hist[2].data_ptr(),
                                             hist[3].data_ptr()))
        for kk in range(k, k + cnt) if peer is None else ():
            nat.check(lib.accbpg_fw_decide(rt.ctx, rt.stream, Vd.data_ptr(), m, n, st, shard.lo, int(away), float(eps), kk,
                                           recs.data_ptr(), shard.world, ws.data_ptr(), ctrl.data_ptr(),
                                           hist[0].data_ptr(), hist[1].data_ptr(), hist[2].data_ptr(), hist[3].data_ptr(),
                                           vcol.data_ptr()))
            shard.sum_(vcol)
            nat.check(lib.accbpg_fw_step(rt.ctx, rt.stream, Vd.data_ptr(), m, n, st, shard.lo, int(away), kk, ws.data_ptr(),
                                         Hinv.data_ptr(), vcol.data_ptr(), x.data_ptr(), w.data_ptr(), ctrl.data_ptr(),
                                         rec.data_ptr()))
            shard.all_gather_equal(recs, rec)
        c = ctrl.cpu().numpy()
        done = int(c[C_NITER])
        if index_log is not None and done == k + 1:
            index_log.append((int(c[C_IMAX]), int(c[C_JMIN]), None if c[C_STOP] != 0 else int(c[C_MODE])))
        if verbose and shard.rank == 0 and done > k:
            hb = hist[:, k:done].cpu().numpy()
            for q in range(k, done):
                if q % verbskip == 0:
                    print("{0:6d}  {1:10.3e}  {2:10.3e}  {3:10.3e}".format(q, hb[0, q - k], hb[1, q - k], hb[2, q - k]))
        k += cnt
        if c[C_STOP] != 0:
            break
    hh = hist[:, :done].cpu().numpy()
    F, SP, SN = hh[0].copy(), hh[1].copy(), hh[2].copy()
    T = t_setup + (hh[3] - hh[3][0]) * 1e-9 if done > 0 else np.zeros(0)
    if peer is not None:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(shard.group)        # nobody frees its receive buffers while a peer may still store into them
    return like_input(x, host), F, SP, SN, T


def D_opt_FW(V, x0, eps, maxitrs, verbose=True, verbskip=1, batch=64, device=None, index_log=None, shard=None):
    """Frank-Wolfe for D-optimal design.   accbpg/D_opt_alg.py:9-88.   Returns (x, F, SP, SN, T).
    shard: a ColumnShard when V / x0 are this rank's column slab / slice (the returned x is the local slice)."""
    if shard is not None and shard.world > 1:
        return _run_sharded(V, x0, eps, maxitrs, 0, verbose, verbskip, batch, device, index_log, shard)
    return _run(V, x0, eps, maxitrs, 0, verbose, verbskip, batch, device, index_log)


def D_opt_FW_away(V, x0, eps, maxitrs, verbose=True, verbskip=1, batch=64, device=None, index_log=None, shard=None):
    """Frank-Wolfe with away steps (Wolfe-Atwood).   accbpg/D_opt_alg.py:91-185.   Returns (x, F, SP, SN, T)."""
    if shard is not None and shard.world > 1:
        return _run_sharded(V, x0, eps, maxitrs, 1, verbose, verbskip, batch, device, index_log, shard)
    return _run(V, x0, eps, maxitrs, 1, verbose, verbskip, batch, device, index_log)
