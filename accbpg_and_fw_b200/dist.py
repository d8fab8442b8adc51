"""Column (n) sharding of the design / data matrix over one-process-per-GPU ranks.

The reference is single-process; the path shards naturally by columns:
rank r owns columns [lo, hi) of H / A and the matching slice of every length-n vector, while
m-sized objects (M, L, Ax, b) and all scalars are replicated.  The only exchange steps are
  * all-reduce(sum) of the m x m Gram matrix (D-opt) or of the m-vector Ax (Poisson / KL),
  * all-reduce of a handful of scalars per driver step (divergences, dot products, Newton sums),
  * (value, index) extremum with lowest-index tie-break for the simplex LMO.
On GPUs under NCCL these exchanges run over NVLink peer memory inside this library's own kernels (payload stores into
every rank's symmetric receive buffer + a released flag word; peer_buffers below hands out the buffers and
config.peer_allreduce switches the scheme off); torch.distributed collectives are the fallback and carry everything
else (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


class ColumnShard:
    def __init__(self, n, group=None, rank=None, world=None):
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n, self.rank, self.world = int(n), int(rank), int(world)
        self.offsets = self.partition(self.n, self.world)
        self.lo, self.hi = self.offsets[self.rank], self.offsets[self.rank + 1]
        self.n_local = self.hi - self.lo
        self.width = max(self.offsets[i + 1] - self.offsets[i] for i in range(self.world))

    @staticmethod
    def partition(n, world):
        """Contiguous, balanced, every boundary even (keeps 16-byte alignment of row slabs)."""
        pairs, odd = divmod(n, 2)
        q, r = divmod(pairs, world)
        out = [0]
        for i in range(world):
            out.append(out[-1] + 2 * (q + (1 if i < r else 0)))
        out[-1] += odd            # an odd last column goes to the last rank
        return out

    # ---- data movement ------------------------------------------------------------------
    def cols(self, M):
        """Local column slab of a host or device matrix (copy, C-contiguous)."""
        if isinstance(M, torch.Tensor):
            return M[:, self.lo:self.hi].contiguous()
        import numpy as np
        return np.ascontiguousarray(M[:, self.lo:self.hi])

    def part(self, v):
        if isinstance(v, torch.Tensor):
            return v[self.lo:self.hi].contiguous()
        import numpy as np
        return np.ascontiguousarray(v[self.lo:self.hi])

    def gather(self, t):
        """All ranks receive the full length-n vector assembled from the local slices."""
        if self.world == 1:
            return t.clone()
        width = max(self.offsets[i + 1] - self.offsets[i] for i in range(self.world))
        pad = torch.zeros(width, dtype=t.dtype, device=t.device)
        pad[: t.numel()] = t
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([bufs[i][: self.offsets[i + 1] - self.offsets[i]] for i in range(self.world)])

    def all_gather_equal(self, out, t):
        """out[r*len(t):(r+1)*len(t)] = rank r's t (every rank passes the same length): one collective."""
        if self.world == 1:
            out.copy_(t)
            return out
        if t.is_cuda:
            dist.all_gather_into_tensor(out, t, group=self.group)
        else:                                   # gloo (CPU tests)
            bufs = list(out.view(self.world, -1).unbind(0))
            dist.all_gather(bufs, t, group=self.group)
        return out

    # ---- collectives ----------------------------------------------------------------------
    def sum_(self, t):
        """In-place sum over the ranks.  float64 CUDA tensors up to 2**21 elements go through NVLink peer memory
        (accbpg_peer_sum_scalars / accbpg_peer_sum_vector: rank-ordered sums, identical bits on every rank); anything
        else, and every case where symmetric memory is unavailable, is an NCCL / gloo all-reduce."""
        if self.world > 1 and not self._peer_sum(t):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def _peer_sum(self, t):
        n = t.numel()
        if not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous() or n == 0 or n > (1 << 21):
            return False
        from . import _native as nat
        from .runtime import Runtime
        rt = Runtime.get(t.device)
        if n <= 15:
            return self.sum_scalars_into(t, t)
        cap = 1 << (n - 1).bit_length()
        pb = peer_buffers(self, t.device, [(2 * self.world * cap, torch.float64), (self.world, torch.int64)],
                          cache_key=("sumvec", cap), reset=False)
        if pb is None:
            return False
        nat.check(nat.lib.accbpg_peer_sum_vector(rt.ctx, rt.stream, t.data_ptr(), n, cap, self.rank, self.world,
                                                 pb.tables[0], pb.tables[1], pb.next_epoch()))
        return True

    def sum_scalars_into(self, src, dst):
        """dst <- sum over the ranks of src (<= 15 float64 CUDA scalars; dst may be src) through peer memory, with every
        rank's device status word OR-ed into every other rank's (an assertion that fails on one rank's slice is then
        raised by all ranks at the same read instead of leaving the others waiting in the next exchange).  Returns False
        when peer memory is unavailable: the caller falls back to collectives."""
        if self.world < 2 or not src.is_cuda:
            return False
        from . import _native as nat
        from .runtime import Runtime
        pb = peer_buffers(self, src.device, [(2 * self.world * 16, torch.float64), (self.world, torch.int64)],
                          cache_key="sum16", reset=False)
        if pb is None:
            return False
        rt = Runtime.get(src.device)
        nat.check(nat.lib.accbpg_peer_sum_scalars(rt.ctx, rt.stream, src.data_ptr(), dst.data_ptr(), src.numel(),
                                                  self.rank, self.world, pb.tables[0], pb.tables[1], pb.next_epoch()))
        return True

    def min_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return t

    def max_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def argmin_(self, pair):
        """pair = tensor [value, global_index] (float64).  In place: the minimum value over ranks and,
        among ranks attaining it, the lowest global index -- independent of the GPU count."""
        if self.world == 1:
            return pair
        if pair.is_cuda and pair.dtype == torch.float64 and pair.is_contiguous() and pair.numel() == 2:
            pb = peer_buffers(self, pair.device, [(2 * self.world * 16, torch.float64), (self.world, torch.int64)],
                              cache_key="argmin", reset=False)
            if pb is not None:
                from . import _native as nat
                from .runtime import Runtime
                rt = Runtime.get(pair.device)
                nat.check(nat.lib.accbpg_peer_argmin_pair(rt.ctx, rt.stream, pair.data_ptr(), self.rank, self.world,
                                                          pb.tables[0], pb.tables[1], pb.next_epoch()))
                return pair
        bufs = [torch.empty_like(pair) for _ in range(self.world)]
        dist.all_gather(bufs, pair, group=self.group)
        allp = torch.stack(bufs)                          # [world, 2]
        vmin = allp[:, 0].min()
        cand = torch.where(allp[:, 0] == vmin, allp[:, 1], torch.full_like(allp[:, 1], float("inf")))
        pair[0] = vmin
        pair[1] = cand.min()
        return pair

    def owner(self, col):
        """Rank that owns global column `col`."""
        for r in range(self.world):
            if self.offsets[r] <= col < self.offsets[r + 1]:
                return r
        raise IndexError(col)


_peer_cache = {}


class PeerBuffers:
    """Symmetric buffers of one exchange: `tensors` (this rank's), `tables` (per buffer, a ctypes array with every
    rank's device address of it) and a call counter for the kernels' flag epochs."""

    def __init__(self, tensors, tables, handles):
        self.tensors, self.tables, self.handles = tensors, tables, handles
        self.epoch = 0

    def next_epoch(self):
        self.epoch += 1
        return self.epoch


def peer_buffers(shard, device, specs, cache_key=None, reset=True):
    """Symmetric (NVLink peer-mapped) buffers for the kernels that exchange data without NCCL.  specs: list of
    (numel, dtype).  Returns a PeerBuffers, or None when symmetric memory is not available (not NCCL, one rank, switched
    off by config.peer_allreduce, or torch cannot map the buffers).  Collective: every rank of the group must call it.
    New buffers come back zero-filled and a barrier has passed, so flag words are zero on every rank.
    cache_key: reuse the buffers of an earlier call with the same key (mapping symmetric memory costs ~0.1 s).  With
    reset=True a reused set is zeroed again behind a barrier and its epoch restarts (the caller guarantees that no peer
    still stores into it: a barrier after its last kernel); with reset=False it is handed back as it is and the epochs
    simply continue (exchanges whose every store is awaited by the receiving kernel of the same call)."""
    import ctypes
    import warnings
    from . import config
    if shard is None or shard.world < 2 or shard.world > 16 or not config.peer_allreduce or not dist.is_initialized():
        return None
    if dist.get_backend(shard.group) != "nccl":
        return None
    key = None
    if cache_key is not None:
        key = (cache_key, id(shard.group), shard.world, str(device), tuple((int(a), str(b)) for a, b in specs))
        hit = _peer_cache.get(key)
        if hit is False:                                  # mapping failed before: do not try (and warn) on every call
            return None
        if hit is not None:
            if reset:
                for b in hit.tensors:
                    b.zero_()
                hit.epoch = 0
                torch.cuda.synchronize()
                dist.barrier(shard.group)
            return hit
    pb, err = None, None
    try:
        import torch.distributed._symmetric_memory as symm
        grp = shard.group if shard.group is not None else dist.group.WORLD
        bufs = [symm.empty(int(numel), dtype=dtype, device=device) for numel, dtype in specs]
        for b in bufs:
            b.zero_()
        hdls = [symm.rendezvous(b, grp) for b in bufs]
        arrs = [(ctypes.c_void_p * shard.world)(*[int(p) for p in hd.buffer_ptrs]) for hd in hdls]
        torch.cuda.synchronize()
        pb = PeerBuffers(bufs, arrs, hdls)
    except Exception as exc:                              # no symmetric memory on this system: the callers use NCCL
        err = exc
    # the ranks must take the same path: one that could not map its buffers sends everybody to the collectives
    ok = torch.tensor([1 if pb is not None else 0], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=shard.group)
    if int(ok.item()) == 0:
        warnings.warn(f"peer-memory buffers unavailable on some rank ({err!r}); using NCCL")
        if key is not None:
            _peer_cache[key] = False
        return None
    dist.barrier(shard.group)
    if key is not None:
        _peer_cache[key] = pb
    return pb
