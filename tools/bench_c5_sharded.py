"""BASELINE.json configs[4] on N GPUs of one box (torchrun): D-opt 2000 x 1 000 000 with H column-sharded, ABPG_gain and
D_opt_FW_away.  Strong scaling: the same 16 GB instance split over the ranks (each rank generates its own slab on the
device from a seed that depends on the slab only).  Rank 0 prints one JSON line."""
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

m = int(sys.argv[1]) if len(sys.argv) > 1 else 2000          # optional: rows, total columns, FW iterations
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
fw_iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
SLABS = 8                               # the instance is defined as 8 slabs of 125000 columns, whatever the rank count
dev = torch.device("cuda", local)
sh = acc.ColumnShard(n) if world > 1 else None
lo, hi = (sh.lo, sh.hi) if sh is not None else (0, n)
per = n // SLABS
parts = []
for sl in range(SLABS):
    a, b = sl * per, (sl + 1) * per
    if b <= lo or a >= hi:
        continue
    gen = torch.Generator(device=dev)
    gen.manual_seed(100 + sl)
    slab = torch.randn(m, per, dtype=torch.float64, device=dev, generator=gen)
    parts.append(slab[:, max(lo, a) - a: min(hi, b) - a])
H = torch.cat(parts, dim=1).contiguous() if len(parts) > 1 else parts[0].contiguous()
del parts
f = acc.DOptimalObj(H, shard=sh)
h = acc.BurgEntropySimplex(shard=sh)
x0 = torch.full((hi - lo,), 1.0 / n, dtype=torch.float64, device=dev)


def timed(fn):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return r, ms


acc.ABPG_gain(f, h, 1.0, x0, gamma=2, maxitrs=2, verbose=False)
iters = 8
res, ms = timed(lambda: acc.ABPG_gain(f, h, 1.0, x0, gamma=2, maxitrs=iters, verbose=False))
F = res[1]
acc.D_opt_FW_away(H, x0, 1e-12, 20, verbose=False, shard=sh)
resf, msf = timed(lambda: acc.D_opt_FW_away(H, x0, 1e-12, fw_iters, verbose=False, shard=sh))
Tf = resf[4]
if rank == 0:
    print(json.dumps({"config": f"D-opt {m}x{n}, H column-sharded", "n_gpus": world,
                      "abpg_gain_ms_per_iteration": ms / len(F), "abpg_gain_it_per_s": len(F) / (ms * 1e-3),
                      "F": [float(v) for v in F],
                      "fw_away_it_per_s": len(resf[1]) / (msf * 1e-3),        # whole call, setup included
                      "fw_away_it_per_s_from_T": (len(Tf) - 1) / (Tf[-1] - Tf[0]),  # the reference's convention
                      "peer_memory": bool(__import__("accbpg_and_fw_b200.config", fromlist=["x"]).peer_allreduce), "fw_away_F_last": float(resf[1][-1])}), flush=True)
if world > 1:
    dist.destroy_process_group()
