"""Per-device runtime: the C context, result slots, workspaces and host<->device vector plumbing.

PyTorch is used for device memory, streams and (in dist.py) NCCL only; every numerical
operation goes through the C ABI in include/accbpg_b200.h.
"""
import ctypes

import numpy as np
import torch

from . import _native as nat

lib = nat.lib
F64 = torch.float64

# host-visible meaning of the device status bits (the reference's assertion texts)
_STATUS_ERRORS = [
    ("X_NEGATIVE", AssertionError, "DOptimalObj: x needs to be nonnegative"),
    ("NOT_PD", ValueError, "HXHT is singular or not positive definite"),
    ("ARG_NOT_POS", AssertionError, "BurgEntropy: entries of the argument are not positive"),
    ("PROX_NOT_POS", AssertionError, "BurgEntropy prox_map: shifted gradient is not positive"),
    ("ARG_NEGATIVE", AssertionError, "ShannonEntropy takes nonnegative arguments"),
    ("Y_NOT_POS", AssertionError, "prox_map needs positive arguments"),
    ("NEWTON_MAXIT", RuntimeError, "BurgEntropySimplex: Newton iteration guard reached"),
]


class Runtime:
    """One per CUDA device.  Holds the C context and a float64 tensor of result slots."""

    _by_device = {}

    # slot map (indices into self.scal) used by the drivers so one read fetches a whole line-search trip
    S_F, S_F2, S_DXY, S_DZZ, S_DOT, S_PSI, S_AUX0, S_AUX1, S_AUX2, S_AUX3 = range(10)
    S_SUM = 16          # S_SUM .. S_SUM+3: sums over the ranks of S_DXY .. S_PSI (column-sharded runs); +4: status
    S_TMP = 32          # scratch for the synchronous operator methods
    N_SLOTS = 64

    @classmethod
    def get(cls, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("accbpg_and_fw_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        rt = cls._by_device.get(idx)
        if rt is None:
            rt = cls(idx)
            cls._by_device[idx] = rt
        return rt

    def __init__(self, index):
        self.index = index
        self.device = torch.device("cuda", index)
        with torch.cuda.device(index):
            handle = ctypes.c_void_p()
            nat.check(lib.accbpg_ctx_create(ctypes.byref(handle)))
            self.ctx = handle
            self.sm_count = lib.accbpg_ctx_sm_count(handle)
            self.scal = torch.zeros(self.N_SLOTS, dtype=F64, device=self.device)
        self._scal_ptr = self.scal.data_ptr()
        self._hbuf = (ctypes.c_double * 256)()
        self._hbuf_addr = ctypes.addressof(self._hbuf)
        self._hstat = ctypes.c_uint32(0)
        self._ws = {}
        self.dist = None            # set by dist.ColumnShard when the problem is column-sharded
        self._pin = {}              # numel -> (ring of pinned staging buffers, next index) for host-vector uploads

    def on_device(self):
        """Context manager that makes this runtime's device current (for the few calls that take no context)."""
        return torch.cuda.device(self.index)

    # ---- streams / pointers ------------------------------------------------------------
    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def slot(self, i):
        return self._scal_ptr + 8 * i

    def read(self, first, count):
        """Fetch scal[first:first+count] (synchronises the stream) and raise on a status bit."""
        nat.check(lib.accbpg_ctx_read(self.ctx, self.stream, self.slot(first), count, self._hbuf_addr,
                                      ctypes.byref(self._hstat)))
        st = self._hstat.value
        if st:
            self.raise_status(st)
        return self._hbuf[0:count]

    def read_raw(self, first, count):
        """Like read() but returns (values, status) without raising."""
        nat.check(lib.accbpg_ctx_read(self.ctx, self.stream, self.slot(first), count, self._hbuf_addr,
                                      ctypes.byref(self._hstat)))
        return self._hbuf[0:count], self._hstat.value

    def read_async(self, first, count):
        """Enqueue the fetch of scal[first:first+count] behind the work already on the stream; returns a ticket."""
        t = ctypes.c_int(0)
        nat.check(lib.accbpg_ctx_read_async(self.ctx, self.stream, self.slot(first), count, ctypes.byref(t)))
        return t.value

    def read_wait(self, ticket, count):
        """Wait for a deferred fetch; raises on a status bit like read()."""
        nat.check(lib.accbpg_ctx_read_wait(self.ctx, ticket, count, self._hbuf_addr, ctypes.byref(self._hstat)))
        st = self._hstat.value
        if st:
            self.raise_status(st)
        return self._hbuf[0:count]

    @staticmethod
    def raise_status(st):
        for name, exc, msg in _STATUS_ERRORS:
            if st & nat.ST[name]:
                raise exc(msg)
        raise RuntimeError(f"unknown device status 0x{st:x}")

    # ---- memory ------------------------------------------------------------------------
    def workspace(self, key, nbytes):
        """Byte workspace cached per (key); grown on demand, never shrunk.  Zero-filled when allocated: the D-opt
        workspace keeps progress counters and the (never written) upper part of L^-1 there from call to call."""
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    def empty(self, n):
        return torch.empty(int(n), dtype=F64, device=self.device)

    def to_device(self, a):
        """numpy / torch (any device) -> contiguous float64 CUDA tensor on this device (no copy if already so)."""
        if isinstance(a, torch.Tensor):
            if a.device == self.device and a.dtype == F64 and a.is_contiguous():
                return a
            return a.to(device=self.device, dtype=F64).contiguous()
        arr = np.ascontiguousarray(a, dtype=np.float64)
        if arr.ndim == 1 and 0 < arr.size <= (1 << 24):
            base = arr.base
            if isinstance(base, torch.Tensor) and base.is_pinned() and arr.flags.writeable:
                # a vector this package returned (like_input) and the caller passes back unchanged in type: its memory
                # is page-locked already, so the DMA reads it in place (the calling operator synchronises before it
                # returns, so the caller cannot modify it under the copy)
                out = torch.empty(arr.size, dtype=F64, device=self.device)
                out.copy_(torch.from_numpy(arr), non_blocking=True)
                return out
            # host vectors go through a small ring of pinned staging buffers: one CPU memcpy, then an asynchronous DMA
            # on the current stream (a pageable cudaMemcpy would block and stage through the driver's own buffer).
            # Every public operator call ends with a synchronising read, so a ring of 8 cannot be overrun.
            ring = self._pin.get(arr.size)
            if ring is None:
                ring = [[torch.empty(arr.size, dtype=F64).pin_memory() for _ in range(8)], 0]
                self._pin[arr.size] = ring
            buf = ring[0][ring[1]]
            ring[1] = (ring[1] + 1) % 8
            buf.numpy()[:] = arr
            out = torch.empty(arr.size, dtype=F64, device=self.device)
            out.copy_(buf, non_blocking=True)
            return out
        if not arr.flags.writeable:
            arr = arr.copy()
        return torch.from_numpy(arr).to(self.device)


def is_host(a):
    return not isinstance(a, torch.Tensor)


def like_input(t, host):
    """Return a device result in the form the caller used: NumPy for NumPy inputs, the tensor otherwise.
    Host results are fresh arrays in page-locked memory (torch's caching host allocator owns the block and takes it back
    when the array is dropped): the download is one DMA with no pageable staging copy."""
    if not host:
        return t
    if t.dim() == 1 and 0 < t.numel() <= (1 << 24) and t.dtype == F64:
        out = torch.empty(t.numel(), dtype=F64, pin_memory=True)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return out.numpy()
    return t.cpu().numpy()
