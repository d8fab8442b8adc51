"""Column-sharded parity check, one process per GPU (launched by tests/test_gpu_multi.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/multigpu_check.py

Every rank holds a column slab of H / A and the matching slices of the iterates; results are compared with the
single-process oracle on the full problem (trajectory F to 1e-9, LMO vertices exact, gathered iterates)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import accbpg_and_fw_b200 as acc
    from oracle import accbpg_oracle as orc

    def ferr(a, b):
        n = min(len(a), len(b))
        return float(np.max(np.abs(a[:n] - b[:n]) / np.maximum(np.abs(b[:n]), 1e-3)))

    def gathered(sh, xloc):
        return sh.gather(torch.as_tensor(xloc, device="cuda") if not isinstance(xloc, torch.Tensor) else xloc).cpu().numpy()

    report = {}
    # ---- D-optimal design, 80 x 200 (golden instance) and 300 x 6002 --------------------------------------
    for (m, n, seed, its) in [(80, 200, 10, 300), (300, 6002, 3, 40)]:
        fo, ho, Lo, x0 = orc.D_opt_design(m, n, randseed=seed)
        sh = acc.ColumnShard(n)
        f = acc.DOptimalObj(sh.cols(fo.H), shard=sh)
        h = acc.BurgEntropySimplex(shard=sh)
        # operator level
        fx, g = f.func_grad(sh.part(x0))
        fxo, go = fo.func_grad(x0)
        assert abs(fx - fxo) <= 1e-10 * abs(fxo), (fx, fxo)
        assert np.max(np.abs(g - sh.part(go)) / np.abs(sh.part(go))) <= 1e-9
        assert abs(h.divergence(sh.part(x0 * 1.5), sh.part(x0)) - ho.divergence(x0 * 1.5, x0)) <= 1e-12 * n
        z = h.div_prox_map(sh.part(x0), g, 0.7)
        zo = ho.div_prox_map(x0, go, 0.7)
        assert np.max(np.abs(z - sh.part(zo)) / np.abs(sh.part(zo))) <= 1e-10
        # trajectories
        x, F, Ls, T = acc.BPG(f, h, Lo, sh.part(x0), maxitrs=its, verbose=False)
        xo, Fo, Lso, To = orc.BPG(fo, ho, Lo, x0, maxitrs=its)
        assert ferr(F, Fo) <= 1e-9 and np.array_equal(Ls, Lso), (ferr(F, Fo),)
        assert np.max(np.abs(gathered(sh, x) - xo) / np.abs(xo)) <= 1e-6
        out = acc.ABPG_gain(f, h, Lo, sh.part(x0), gamma=2, maxitrs=its, verbose=False)
        outo = orc.ABPG_gain(fo, ho, Lo, x0, gamma=2, maxitrs=its)
        assert ferr(out[1], outo[1]) <= 1e-9 and np.mean(out[2] != outo[2]) <= 0.02
        # Frank-Wolfe with the sharded simplex LMO: vertex sequence must not depend on the GPU count
        lmo = acc.lmo_simplex(shard=sh)
        log_o = []
        xs, Fs, Lss, Ts = acc.FW_alg_div_step(f, h, Lo, sh.part(x0), maxitrs=min(its, 60), gamma=2.0, lmo=lmo, verbose=False)
        xr, Fr, Lsr, Tr = orc.FW_alg_div_step(fo, ho, Lo, x0, min(its, 60), 2.0, orc.make_lmo_simplex(), vertex_log=log_o)
        assert ferr(Fs, Fr) <= 1e-9 and np.array_equal(Lss, Lsr)
        # D_opt_FW / D_opt_FW_away with V column-sharded: owner-contributed column, replicated Hinv
        for away, fn_g, fn_o in ((1, acc.D_opt_FW_away, orc.D_opt_FW_away), (0, acc.D_opt_FW, orc.D_opt_FW)):
            glog, olog = [], []
            xa, Fa, SPa, SNa, Ta = fn_g(sh.cols(fo.H), sh.part(x0), 1e-8, min(its, 80), verbose=False, shard=sh,
                                        index_log=glog)
            xb, Fb, SPb, SNb, Tb = fn_o(fo.H, x0, 1e-8, min(its, 80), index_log=olog)
            assert len(Fa) == len(Fb) and ferr(Fa, Fb) <= 1e-9, (len(Fa), len(Fb))
            assert [(a, b) for a, b, *_ in glog[:len(olog) - 1]] == [(a, b) for a, b, *_ in olog[:len(olog) - 1]]
            assert np.max(np.abs(gathered(sh, xa) - xb)) <= 1e-9
        xa, Fa, SPa, SNa, Ta = acc.D_opt_FW_away(sh.cols(fo.H), sh.part(x0), 1e-8, its, verbose=False, shard=sh)   # batched
        xb, Fb, SPb, SNb, Tb = orc.D_opt_FW_away(fo.H, x0, 1e-8, its)
        assert len(Fa) == len(Fb) and ferr(Fa, Fb) <= 1e-9
        report[f"dopt_{m}x{n}"] = (ferr(F, Fo), ferr(out[1], outo[1]), ferr(Fs, Fr), ferr(Fa, Fb))
        # Gram all-reduce through NVLink peer memory (the default) against the NCCL all-reduce of the same matrices
        from accbpg_and_fw_b200 import config
        report[f"peer_allreduce_{m}"] = f._peer is not None
        if f._peer is not None:
            config.peer_allreduce = False
            f_nccl = acc.DOptimalObj(sh.cols(fo.H), shard=sh)
            config.peer_allreduce = True
            assert f_nccl._peer is None
            for rep in range(20):                      # repeated calls: flag epochs and buffer reuse
                xt = sh.part(x0 * (1.0 + 0.01 * rep))
                a, ga = f.func_grad(xt)
                b, gb = f_nccl.func_grad(xt)
                assert abs(a - b) <= 1e-13 * abs(b), (rep, a, b)
                assert np.max(np.abs(ga - gb) / np.abs(gb)) <= 1e-11, rep
    # LMO tie across ranks: lowest global index wins
    n = 1000
    sh = acc.ColumnShard(n)
    gfull = np.ones(n)
    gfull[[900, 17, 600]] = -3.0
    lmo = acc.lmo_simplex(2.0, shard=sh)
    s = lmo(sh.part(gfull))
    sfull = gathered(sh, s)
    assert np.array_equal(sfull, orc.lmo_simplex_eval(gfull, 2.0)), np.nonzero(sfull > 1)[0]

    # ---- KL regression + Shannon simplex (config C3 family) -------------------------------------------------
    np.random.seed(5)
    m, n = 150, 400
    A = np.random.rand(m, n)
    A = A / A.sum(axis=0)
    xs = np.random.rand(n)
    xs /= xs.sum()
    b = np.dot(A, xs) * (1 + 0.01 * (np.random.rand(m) - 0.5))
    sh = acc.ColumnShard(n)
    f = acc.KLdivRegression(sh.cols(A), b, shard=sh)
    h = acc.ShannonEntropySimplex(shard=sh)
    fo, ho = orc.make_kl(A, b), orc.make_shannon("simplex")
    x0 = np.ones(n) / n
    out = acc.ABPG_gain(f, h, 1.0, sh.part(x0), gamma=2.0, maxitrs=120, verbose=False)
    outo = orc.ABPG_gain(fo, ho, 1.0, x0, gamma=2.0, maxitrs=120)
    assert ferr(out[1], outo[1]) <= 1e-9, ferr(out[1], outo[1])
    report["kl_shannon_simplex"] = ferr(out[1], outo[1])
    # Poisson + Burg L1 (config C4 family)
    fo, ho, Lo, x0 = orc.Poisson_regrL1(200, 100, noise=1e-4, lamda=0.01, randseed=1)
    sh = acc.ColumnShard(100)
    f = acc.PoissonRegression(sh.cols(fo.A), fo.b, shard=sh)
    h = acc.BurgEntropyL1(0.01, shard=sh)
    out = acc.BPG(f, h, Lo, sh.part(x0), maxitrs=150, verbose=False)
    outo = orc.BPG(fo, ho, Lo, x0, maxitrs=150)
    assert ferr(out[1], outo[1]) <= 1e-9 and np.array_equal(out[2], outo[2]), ferr(out[1], outo[1])
    report["poisson_burgL1"] = ferr(out[1], outo[1])
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_OK world=%d %s" % (world, report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
