"""Thin helpers between torch (device memory, streams, torch.distributed) and the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.RsxError("rsx needs a CUDA device: there is no CPU implementation of the hot path")
    _lib.load()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (or NULL for None)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "rsx kernels need contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def hptr(a):
    """Host pointer of a C-contiguous numpy array (or NULL for None); keep `a` alive during the call."""
    if a is None:
        return C.c_void_p(0)
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def to_device(a: np.ndarray, pinned: bool = True):
    """H2D copy of a numpy array through pinned memory on the current stream."""
    t = torch.from_numpy(np.ascontiguousarray(a))
    if pinned:
        t = t.pin_memory()
    return t.cuda(non_blocking=True)


_FETCH_STAGE = {}


def fetch(t: torch.Tensor) -> np.ndarray:
    """Small device tensor -> new numpy array; synchronises the current stream.  A kernel stores the bytes straight into
    page-locked host memory (rsx_store_to_host), so the read-back never waits behind a large transfer on the copy engine."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes == 0:
        return t.cpu().numpy()
    cap = max(4096, 1 << (nbytes - 1).bit_length())
    stage = _FETCH_STAGE.get(cap)
    if stage is None:
        stage = _FETCH_STAGE[cap] = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    _lib.call("rsx_store_to_host", ptr(t), C.c_void_p(stage.data_ptr()), nbytes, stream_ptr())
    torch.cuda.current_stream().synchronize()
    return stage[:nbytes].view(t.dtype).reshape(t.shape).numpy().copy()


def stage_to_host(t: torch.Tensor) -> torch.Tensor:
    """Asynchronous fetch: starts the kernel store of a small device tensor into a NEW page-locked tensor and returns it; the
    contents are valid after the next synchronisation of the current stream."""
    t = t.contiguous()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    if t.numel():
        _lib.call("rsx_store_to_host", ptr(t), C.c_void_p(host.data_ptr()), t.numel() * t.element_size(), stream_ptr())
    return host


class _UploadRing:
    """Small host arrays reach the device through a ring of page-locked chunks and cudaMemcpyAsync: torch's `.to(device)` of a
    pageable tensor WAITS for the stream (staging copy + synchronise), which drains every kernel queued before it - in the middle
    of a scene that is a hidden full synchronisation."""
    CHUNK, N = 1 << 17, 24

    def __init__(self):
        self.chunks = [torch.empty(self.CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(self.N)]
        self.events = [None] * self.N
        self.i = 0

    def upload(self, a: np.ndarray, device) -> torch.Tensor:
        a = np.ascontiguousarray(a)
        nbytes = a.nbytes
        src = torch.from_numpy(a.reshape(-1).view(np.uint8)) if nbytes else None
        out = torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, a.dtype)).dtype, device=device)
        if not nbytes:
            return out
        if nbytes > self.CHUNK:                                   # large: a one-off pinned copy
            pinned = src.pin_memory()
            out.view(torch.uint8).reshape(-1).copy_(pinned, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out
        k = self.i
        self.i = (k + 1) % self.N
        if self.events[k] is not None:
            self.events[k].synchronize()                            # the copy that last used this chunk (long done)
        self.chunks[k][:nbytes].copy_(src)
        out.view(torch.uint8).reshape(-1).copy_(self.chunks[k][:nbytes], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev
        return out


_UPLOAD = {}


def upload_small(a: np.ndarray, device=None) -> torch.Tensor:
    """Host array -> device tensor of the same shape and dtype without synchronising the stream (see _UploadRing)."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if device.type != "cuda":                                       # host-side callers (the gloo tests)
        return torch.from_numpy(np.ascontiguousarray(a)).to(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    ring = _UPLOAD.get(key)
    if ring is None:
        ring = _UPLOAD[key] = _UploadRing()
    return ring.upload(a, device)


class MinMaxTracker:
    """uint32 [n][2] device tracker (include/rsx.h 'min/max trackers')."""

    def __init__(self, n: int, device=None):
        self.n = n
        self.buf = torch.empty((n, 2), dtype=torch.int32, device=device or "cuda")
        _lib.call("rsx_minmax_init", ptr(self.buf), n, stream_ptr())

    def slot(self, i: int):
        return C.c_void_p(self.buf.data_ptr() + 8 * i)

    def merged(self, comm=None) -> torch.Tensor:
        """The tracker buffer (uint32 [n][2] stored as int32) on the device; with a multi-rank `comm` the trackers of all ranks
        merged first (the encoding is order preserving as unsigned integers), collectively.  No synchronisation."""
        buf = self.buf
        if comm is not None and comm.world > 1:
            enc = buf.to(torch.int64) & 0xFFFFFFFF
            lo, hi = enc[:, 0].contiguous(), enc[:, 1].contiguous()
            comm.all_reduce(lo, "min")
            comm.all_reduce(hi, "max")
            buf = torch.stack([lo, hi], dim=1).to(torch.int32).contiguous()          # truncation keeps the 32 bits
        return buf

    def read(self, comm=None):
        """(min float32[n], max float32[n]); synchronises (collective with a multi-rank `comm`, see merged())."""
        h = fetch(self.merged(comm)).view(np.uint32)
        mn, mx = np.zeros(self.n, np.float32), np.zeros(self.n, np.float32)
        _lib.load().rsx_minmax_decode(hptr(np.ascontiguousarray(h)), self.n, hptr(mn), hptr(mx))
        return mn, mx


class StageTimer:
    """CUDA-event sections on the current stream: `with timer("name"):`; totals after a synchronise."""

    def __init__(self, enabled: bool = True, only=None):
        self.enabled = enabled
        self.only = set(only) if only is not None else None      # record these sections only (an event pair costs ~3 us)
        self.events = {}

    class _Section:
        def __init__(self, timer, name):
            self.t, self.name = timer, name

        def __enter__(self):
            self.on = self.t.enabled and (self.t.only is None or self.name in self.t.only)
            if self.on:
                self.a = torch.cuda.Event(enable_timing=True)
                self.b = torch.cuda.Event(enable_timing=True)
                self.a.record()
            return self

        def __exit__(self, *exc):
            if self.on:
                self.b.record()
                self.t.events.setdefault(self.name, []).append((self.a, self.b))
            return False

    def __call__(self, name: str):
        return StageTimer._Section(self, name)

    def reset(self):
        self.events = {}

    def totals_ms(self):
        """name -> (total ms, launches); call after torch.cuda.synchronize()."""
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.events.items()}


NO_TIMER = StageTimer(enabled=False)
