"""The other BASELINE.json configurations on one B200 (run under gpurun; writes gpurun_out/configs.json).

  C1  D-opt 80 x 200, BPG with line search, 1000 iterations (latency bound)
  C3  KL regression 20000 x 200000 (32 GB) + ShannonEntropySimplex, ABPG_gain
  C4  Poisson 100000 x 125000 (100 GB: the per-GPU column slab of the 8-GPU configuration) + BurgEntropy, BPG / ABPG_gain
  C5  D-opt 2000 x 1000000 (16 GB), ABPG_gain and D_opt_FW_away  (north_star: >= 60 % of FP64 tensor peak per iteration)

Instances are generated on the device (torch.Generator seeded per config): no host copy of these sizes exists, so
parity for these shapes is established on reduced shapes through the same code path (tests/).  Not part of the tests.
"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import accbpg_and_fw_b200 as acc                      # noqa: E402
from accbpg_and_fw_b200 import _native as nat         # noqa: E402

lib = nat.lib
dev = torch.device("cuda")
which = set(sys.argv[1:]) or {"c1", "c2", "c3", "c4", "c5"}
out = {}


def prof_read():
    res = {}
    tot, cnt = ctypes.c_double(0), ctypes.c_int64(0)
    for i in range(lib.accbpg_prof_count()):
        nat.check(lib.accbpg_prof_read(i, ctypes.byref(tot), ctypes.byref(cnt)))
        if cnt.value:
            res[lib.accbpg_prof_name(i).decode()] = {"ms_total": tot.value, "launches": cnt.value,
                                                     "ms_avg": tot.value / cnt.value}
    return res


def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    torch.cuda.synchronize()
    return r, a.elapsed_time(b)


def fp64_peak():
    N = 8192
    A = torch.randn(N, N, dtype=torch.float64, device=dev)
    B = torch.randn(N, N, dtype=torch.float64, device=dev)
    torch.matmul(A, B)
    best = min(timed(lambda: torch.matmul(A, B))[1] for _ in range(3))
    return 2 * N ** 3 / best / 1e9


peak = fp64_peak()
out["fp64_dgemm_tflops"] = peak
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6551.0)

if "c1" in which:
    f, h, L, x0 = acc.D_opt_design(80, 200, randseed=10)
    acc.BPG(f, h, L, x0, maxitrs=50, verbose=False)
    t0 = time.time()
    x, F, Ls, T = acc.BPG(f, h, L, x0, maxitrs=1000, verbose=False)
    wall = time.time() - t0
    out["c1_dopt_80x200_bpg_ls"] = {"iterations": len(F), "it_per_s": (len(T) - 1) / (T[-1] - T[0]), "wall_s": wall,
                                    "F_last": float(F[-1])}
    print("c1", out["c1_dopt_80x200_bpg_ls"], flush=True)

if "c2" in which:
    f, h, L, x0 = acc.D_opt_design(500, 50000, randseed=1)
    x0d = torch.tensor(x0, device=dev)
    res = {}
    for name, fn in [("ABPG_gain", lambda k: acc.ABPG_gain(f, h, L, x0d, gamma=2, maxitrs=k, verbose=False)),
                     ("ABPG_expo", lambda k: acc.ABPG_expo(f, h, L, x0d, gamma0=3, maxitrs=k, verbose=False)),
                     ("ABPG", lambda k: acc.ABPG(f, h, L, x0d, gamma=2, maxitrs=k, verbose=False)),
                     ("ABDA", lambda k: acc.ABDA(f, h, L, x0d, gamma=2, maxitrs=k, verbose=False)),
                     ("BPG_LS", lambda k: acc.BPG(f, h, L, x0d, maxitrs=k, verbose=False)),
                     ("FW_alg_div_step", lambda k: acc.FW_alg_div_step(f, h, L, x0d, k, 2.0, acc.lmo_simplex(), verbose=False))]:
        fn(5)
        (out_, ms) = timed(lambda: fn(60))
        res[name] = {"iterations": len(out_[1]), "ms_per_iteration": ms / len(out_[1]), "it_per_s": len(out_[1]) / (ms * 1e-3),
                     "F_last": float(out_[1][-1])}
    out["c2_dopt_500x50000_drivers"] = res
    print("c2", res, flush=True)
    del f
    torch.cuda.empty_cache()

if "c5" in which:
    m, n = 2000, 1000000
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    H = torch.randn(m, n, dtype=torch.float64, device=dev, generator=gen)
    f = acc.DOptimalObj(H)
    h = acc.BurgEntropySimplex()
    x0 = torch.full((n,), 1.0 / n, dtype=torch.float64, device=dev)
    acc.ABPG_gain(f, h, 1.0, x0, gamma=2, maxitrs=2, verbose=False)
    iters = 8
    # per-kernel durations with the triangular GEMM serialised behind the Cholesky chain ...
    os.environ["ACCBPG_OVERLAP"] = "0"
    lib.accbpg_prof_enable(1)
    prof_read()
    acc.ABPG_gain(f, h, 1.0, x0, gamma=2, maxitrs=iters, verbose=False)
    kern = prof_read()
    lib.accbpg_prof_enable(0)
    # ... and the iteration time as shipped (overlapped, no per-kernel events)
    os.environ["ACCBPG_OVERLAP"] = "1"
    (res, ms) = timed(lambda: acc.ABPG_gain(f, h, 1.0, x0, gamma=2, maxitrs=iters, verbose=False))
    x, F, Gain, Gdiv, Gavg, T = res
    syrk, trmm = kern.get("syrk_tma_kernel"), kern.get("trmm_persistent_kernel")
    flops_kernel = float(m) * m * n
    # executed oracle work of the run: every SYRK and every triangular GEMM is m^2 n, every factorisation m^3/3 (x2 with the inverse)
    executed = (syrk["launches"] + trmm["launches"]) * flops_kernel
    out["c5_dopt_2000x1e6_abpg_gain"] = {
        "iterations": len(F), "ms_per_iteration": ms / len(F), "it_per_s": len(F) / (ms * 1e-3),
        "gain_trips_per_iteration": trmm["launches"] / len(F),
        "syrk_ms": syrk["ms_avg"], "syrk_tflops": flops_kernel / syrk["ms_avg"] / 1e9, "syrk_frac": flops_kernel / syrk["ms_avg"] / 1e9 / peak,
        "trmm_ms": trmm["ms_avg"], "trmm_tflops": flops_kernel / trmm["ms_avg"] / 1e9, "trmm_frac": flops_kernel / trmm["ms_avg"] / 1e9 / peak,
        "chol_ms_total": kern.get("chol_inv_step_kernel(all block columns)", {}).get("ms_total"),
        "executed_tflops_whole_iteration": executed / (ms * 1e-3) / 1e12,
        "executed_frac_of_fp64_peak_whole_iteration": executed / (ms * 1e-3) / 1e12 / peak,
        "F": [float(v) for v in F], "kernels": kern}
    print("c5 abpg_gain", {k: v for k, v in out["c5_dopt_2000x1e6_abpg_gain"].items() if k not in ("kernels", "F")}, flush=True)
    lib.accbpg_prof_enable(1)
    prof_read()
    xa, Fa, SPa, SNa, Ta = acc.D_opt_FW_away(H, x0, 1e-12, 300, verbose=False)
    kfw = prof_read()
    lib.accbpg_prof_enable(0)
    pss = kfw.get("fw_pass_kernel")
    out["c5_dopt_2000x1e6_fw_away"] = {"iterations": len(Ta), "it_per_s": (len(Ta) - 1) / (Ta[-1] - Ta[0]),
                                       "pass_ms": pss["ms_avg"], "pass_GBs": 8.0 * m * n / pss["ms_avg"] / 1e6,
                                       "pass_frac_hbm": 8.0 * m * n / pss["ms_avg"] / 1e6 / hbm}
    print("c5 fw_away", out["c5_dopt_2000x1e6_fw_away"], flush=True)
    del f, H, x, res, xa
    torch.cuda.empty_cache()

if "c3" in which:
    m, n = 20000, 200000
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    A = torch.rand(m, n, dtype=torch.float64, device=dev, generator=gen)
    A /= A.sum(dim=0, keepdim=True)                                   # columns sum to one (applications.py:194-196)
    xs = torch.rand(n, dtype=torch.float64, device=dev, generator=gen)
    xs /= xs.sum()
    b = (A @ xs) * (1 + 0.01 * (torch.rand(m, dtype=torch.float64, device=dev, generator=gen) - 0.5))
    f = acc.KLdivRegression(A, b)
    h = acc.ShannonEntropySimplex()
    x0 = torch.full((n,), 1.0 / n, dtype=torch.float64, device=dev)
    acc.ABPG_gain(f, h, 1.0, x0, gamma=2.0, maxitrs=3, verbose=False)
    lib.accbpg_prof_enable(1)
    prof_read()
    iters = 30
    (res, ms) = timed(lambda: acc.ABPG_gain(f, h, 1.0, x0, gamma=2.0, maxitrs=iters, verbose=False))
    kern = prof_read()
    lib.accbpg_prof_enable(0)
    F = res[1]
    mv, rmv = kern.get("matvec_kernel"), kern.get("rmatvec_kernel")
    out["c3_kl_20000x200000_abpg_gain"] = {
        "iterations": len(F), "ms_per_iteration": ms / len(F), "it_per_s": len(F) / (ms * 1e-3),
        "matvec_ms": mv["ms_avg"], "matvec_GBs": 8.0 * m * n / mv["ms_avg"] / 1e6, "matvec_frac_hbm": 8.0 * m * n / mv["ms_avg"] / 1e6 / hbm,
        "rmatvec_ms": rmv["ms_avg"], "rmatvec_GBs": 8.0 * m * n / rmv["ms_avg"] / 1e6, "rmatvec_frac_hbm": 8.0 * m * n / rmv["ms_avg"] / 1e6 / hbm,
        "passes_over_A_per_iteration": (mv["launches"] + rmv["launches"]) / len(F),
        "F_first_last": [float(F[0]), float(F[-1])]}
    print("c3", out["c3_kl_20000x200000_abpg_gain"], flush=True)
    del f, A, res
    torch.cuda.empty_cache()

if "c4" in which:
    m, n = 100000, 125000
    gen = torch.Generator(device=dev)
    gen.manual_seed(4)
    A = torch.rand(m, n, dtype=torch.float64, device=dev, generator=gen)
    A /= A.sum(dim=0, keepdim=True)
    xt = torch.clamp(torch.rand(n, dtype=torch.float64, device=dev, generator=gen) / n - 0.5 / n, min=0) * 10
    b = A @ xt + 1e-6 * torch.rand(m, dtype=torch.float64, device=dev, generator=gen)
    f = acc.PoissonRegression(A, b)
    h = acc.BurgEntropyL1(lamda=1e-3)
    L = float(b.sum())
    x0 = torch.full((n,), 10.0 / n, dtype=torch.float64, device=dev)
    acc.ABPG_gain(f, h, L, x0, gamma=2.0, maxitrs=2, verbose=False)
    lib.accbpg_prof_enable(1)
    prof_read()
    iters = 10
    (res, ms) = timed(lambda: acc.ABPG_gain(f, h, L, x0, gamma=2.0, maxitrs=iters, verbose=False))
    kern = prof_read()
    lib.accbpg_prof_enable(0)
    F = res[1]
    mv, rmv = kern.get("matvec_kernel"), kern.get("rmatvec_kernel")
    out["c4_poisson_slab_100000x125000_abpg_gain"] = {
        "iterations": len(F), "ms_per_iteration": ms / len(F), "it_per_s": len(F) / (ms * 1e-3),
        "matvec_ms": mv["ms_avg"], "matvec_GBs": 8.0 * m * n / mv["ms_avg"] / 1e6, "matvec_frac_hbm": 8.0 * m * n / mv["ms_avg"] / 1e6 / hbm,
        "rmatvec_ms": rmv["ms_avg"], "rmatvec_GBs": 8.0 * m * n / rmv["ms_avg"] / 1e6, "rmatvec_frac_hbm": 8.0 * m * n / rmv["ms_avg"] / 1e6 / hbm,
        "passes_over_A_per_iteration": (mv["launches"] + rmv["launches"]) / len(F),
        "F_first_last": [float(F[0]), float(F[-1])]}
    print("c4", out["c4_poisson_slab_100000x125000_abpg_gain"], flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
print("wrote gpurun_out/configs.json")
