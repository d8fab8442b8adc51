"""BASELINE.json configs[3] on N GPUs of one box (torchrun): Poisson linear inverse problem, A 100000 x (125000 N)
column-sharded - one 100 GB slab per GPU (weak scaling; 8 GPUs give the 100000 x 1000000 instance), PoissonRegression +
BurgEntropyL1, ABPG_gain.  Every rank generates its slab on the device.  Rank 0 prints one JSON line.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c4_sharded.py [rows] [cols_per_gpu]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import accbpg_and_fw_b200 as acc      # noqa: E402
from accbpg_and_fw_b200 import _native as nat      # noqa: E402
import ctypes      # noqa: E402

lib = nat.lib
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
per = int(sys.argv[2]) if len(sys.argv) > 2 else 125000
n = per * world
dev = torch.device("cuda", local)
sh = acc.ColumnShard(n) if world > 1 else None
assert sh is None or sh.hi - sh.lo == per
gen = torch.Generator(device=dev)
gen.manual_seed(40 + rank)
A = torch.rand(m, per, dtype=torch.float64, device=dev, generator=gen)
A /= A.sum(dim=0, keepdim=True)                                   # columns sum to one (applications.py:116-132)
xt = torch.clamp(torch.rand(per, dtype=torch.float64, device=dev, generator=gen) / n - 0.5 / n, min=0) * 10
b = A @ xt
if world > 1:
    dist.all_reduce(b)
gen0 = torch.Generator(device=dev)
gen0.manual_seed(4)                                               # the same noise on every rank: b is replicated
b = b + 1e-6 * torch.rand(m, dtype=torch.float64, device=dev, generator=gen0)
f = acc.PoissonRegression(A, b, shard=sh)
h = acc.BurgEntropyL1(lamda=1e-3, shard=sh)
L = float(b.sum())
x0 = torch.full((per,), 10.0 / n, dtype=torch.float64, device=dev)


def prof_read():
    out = {}
    tot, cnt = ctypes.c_double(0), ctypes.c_int64(0)
    for i in range(lib.accbpg_prof_count()):
        nat.check(lib.accbpg_prof_read(i, ctypes.byref(tot), ctypes.byref(cnt)))
        if cnt.value:
            out[lib.accbpg_prof_name(i).decode()] = {"ms_avg": tot.value / cnt.value, "launches": cnt.value}
    return out


acc.ABPG_gain(f, h, L, x0, gamma=2.0, maxitrs=2, verbose=False)
lib.accbpg_prof_enable(1)
prof_read()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = acc.ABPG_gain(f, h, L, x0, gamma=2.0, maxitrs=10, verbose=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
kern = prof_read()
F = res[1]
if rank == 0:
    mv, rmv = kern.get("matvec_kernel"), kern.get("rmatvec_kernel")
    print(json.dumps({"config": f"Poisson {m}x{n}, A column-sharded ({per} columns per GPU)", "n_gpus": world,
                      "abpg_gain_ms_per_iteration": ms / len(F), "it_per_s": len(F) / (ms * 1e-3),
                      "matvec_ms": mv["ms_avg"], "matvec_GBs_per_gpu": 8.0 * m * per / mv["ms_avg"] / 1e6,
                      "rmatvec_ms": rmv["ms_avg"], "rmatvec_GBs_per_gpu": 8.0 * m * per / rmv["ms_avg"] / 1e6,
                      "passes_over_A_per_iteration": (mv["launches"] + rmv["launches"]) / len(F),
                      "F_first_last": [float(F[0]), float(F[-1])]}), flush=True)
if world > 1:
    dist.destroy_process_group()
