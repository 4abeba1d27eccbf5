"""Row-strip sharding of a raster over the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU tests).  The raster is cut into
contiguous row strips; every stage is per-pixel or per-window, so the only exchanges are
  * tiny all-reduces: band histograms (int64 sum), PCA moments (float64 sum), feature min/max,
    KMeans partial sums + counts (int64 sum - associative, so results do not depend on the partition);
  * the GLCM halo: windows are top-left anchored (modules/features/indices.py:285), so a strip needs the
    first rows of the strip(s) below it, plus the few rows the bilinear upsample (indices.py:308) reaches
    across a strip boundary.  Expressed here as "fetch global rows [a, b)" from whoever owns them.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from typing import Callable, List, Optional, Sequence, Tuple

import torch

PEER_PASS_ELEMS = 2560                              # include/rsx.h RSX_PEER_PASS_ELEMS
PEER_BLOCK_BYTES = 2 * PEER_PASS_ELEMS * 8 + 8 * 8  # RSX_PEER_BLOCK_BYTES

Range = Tuple[int, int]


def strip_bounds(H: int, G: int, align: int = 1) -> List[Range]:
    """Contiguous row strips of ceil(H/G) rows, rounded up to a multiple of `align`; trailing strips may be empty."""
    rows = int(math.ceil(H / G))
    rows = int(math.ceil(rows / align) * align)
    out = []
    for g in range(G):
        r0 = min(H, g * rows)
        out.append((r0, min(H, r0 + rows)))
    return out


def resize_src_rows(dst_r0: int, dst_r1: int, dst_h: int, src_h: int) -> Range:
    """Source rows [a, b) that cv2.resize(INTER_LINEAR) reads to produce destination rows [dst_r0, dst_r1)
    (rows clamped to the image as OpenCV does)."""
    if dst_r1 <= dst_r0:
        return (0, 0)
    scale = 1.0 / (dst_h / src_h)

    def sy(dy):
        return int(math.floor((dy + 0.5) * scale - 0.5))

    a = min(max(sy(dst_r0), 0), src_h - 1)
    b = min(max(sy(dst_r1 - 1) + 1, 0), src_h - 1) + 1
    return (a, b)


def glcm_rows_needed(own: Range, H: int, window: int, step: int) -> Tuple[Range, Range]:
    """For the strip owning image rows `own`: (props_rows, q_rows) = the rows of the GLCM property map it must
    compute to upsample its own rows, and the rows of the quantised band those windows cover."""
    out_rows = (H - window) // step + 1
    p = resize_src_rows(own[0], own[1], H, out_rows)
    if p[1] <= p[0]:
        return (0, 0), (0, 0)
    return p, (p[0] * step, (p[1] - 1) * step + window)


def exchange_plan(bounds: Sequence[Range], needs: Sequence[Range]) -> List[Tuple[int, int, int, int]]:
    """(src, dst, a, b): src sends its rows [a, b) to dst.  Deterministic, identical on every rank."""
    plan = []
    for dst, (na, nb) in enumerate(needs):
        for src, (r0, r1) in enumerate(bounds):
            if src == dst:
                continue
            a, b = max(na, r0), min(nb, r1)
            if b > a:
                plan.append((src, dst, a, b))
    return plan


class PeerBlocks:
    """This rank's peer-mapped block and the other ranks' mappings of theirs (include/rsx.h, rsx_kmeans_update_peers).
    The sequence number is shared by every KMeans instance of the process: all ranks advance it in lockstep."""

    def __init__(self, own: int, ptrs, rank: int, world: int):
        self.own, self.ptrs, self.rank, self.world = own, ptrs, rank, world
        self.seq = 0

    def next_pass(self) -> Tuple[int, C.c_void_p]:
        """(sequence number of the coming update, device pointer of the pass buffer its assign pass accumulates into)."""
        self.seq += 1
        return self.seq, C.c_void_p(self.own + (self.seq & 1) * PEER_PASS_ELEMS * 8)

    def zero(self, stream):
        from . import _lib
        _lib.call("rsx_peer_zero", C.c_void_p(self.own), 2 * PEER_PASS_ELEMS * 8, stream)


_PEER_CACHE = {}


def release_peer_blocks():
    """Unmaps the peers' blocks and frees this rank's own (rsx_peer_close / rsx_peer_free); registered with atexit."""
    if not _PEER_CACHE:
        return
    try:
        from . import _lib
        lib = _lib.load()
        for pb in _PEER_CACHE.values():
            for p in range(pb.world):
                if p != pb.rank and pb.ptrs[p]:
                    lib.rsx_peer_close(C.c_void_p(pb.ptrs[p]))
            lib.rsx_peer_free(C.c_void_p(pb.own))
    except Exception:
        pass
    _PEER_CACHE.clear()


import atexit

atexit.register(release_peer_blocks)


class Comm:
    """torch.distributed wrapper; `Comm()` without an initialised process group is the 1-GPU case."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.active = dist.is_available() and dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group) if self.active else 0
        self.world = dist.get_world_size(group) if self.active else 1

    def peers(self) -> Optional["PeerBlocks"]:
        """Peer-mapped blocks for the KMeans sums (rsx_kmeans_update_peers), or None: one rank, a CPU backend, more than 8
        ranks, RSX_PEER_REDUCE=0, or CUDA IPC not available on some rank (then every rank stays on the all-reduce).
        Collective on first use."""
        if getattr(self, "_peers_tried", False):
            return self._peers
        self._peers_tried, self._peers = True, None
        if not (self.active and 2 <= self.world <= 8) or os.environ.get("RSX_PEER_REDUCE", "1") == "0":
            return None
        # one block (+ its IPC mappings) per process and group, shared by every Comm / KMeans instance and released at exit:
        # a fresh Comm() per scene must neither leak device memory nor repeat the collective IPC set-up
        key = id(self.group) if self.group is not None else "world"
        if key in _PEER_CACHE:
            self._peers = _PEER_CACHE[key]
            return self._peers
        if self.dist.get_backend(self.group) != "nccl" or not torch.cuda.is_available():
            return None
        from . import _lib
        lib = _lib.load()
        own, handle = C.c_void_p(), (C.c_uint8 * 64)()
        ok = lib.rsx_peer_alloc(PEER_BLOCK_BYTES, C.byref(own), handle) == 0
        mine = torch.tensor(list(handle) + [1 if ok else 0], dtype=torch.uint8, device="cuda")
        every = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(every, mine, group=self.group)
        every = [e.cpu().numpy() for e in every]
        ptrs = (C.c_void_p * self.world)()
        ok = all(int(e[64]) == 1 for e in every)
        if ok:
            for p, e in enumerate(every):
                if p == self.rank:
                    ptrs[p] = own.value
                    continue
                peer = C.c_void_p()
                h = (C.c_uint8 * 64)(*[int(v) for v in e[:64]])
                if lib.rsx_peer_open(h, C.byref(peer)) != 0:
                    ok = False
                    break
                ptrs[p] = peer.value
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            warnings.warn("rsx: CUDA IPC peer mapping is not available; KMeans sums go through all-reduce")
            return None
        self._peers = _PEER_CACHE[key] = PeerBlocks(own.value, ptrs, self.rank, self.world)
        return self._peers

    def all_reduce(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
        if self.active and self.world > 1:
            ops = {"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN, "max": self.dist.ReduceOp.MAX}
            self.dist.all_reduce(t, op=ops[op], group=self.group)
        return t

    def barrier(self):
        if self.active and self.world > 1:
            self.dist.barrier(group=self.group)

    def fetch_rows(self, local: torch.Tensor, bounds: Sequence[Range], needs: Sequence[Range]) -> torch.Tensor:
        """local: this rank's rows (bounds[rank]) of a (rows, W) array.  Returns rows needs[rank] of the global
        array; all ranks call this collectively with the same bounds/needs."""
        r0, r1 = bounds[self.rank]
        na, nb = needs[self.rank]
        assert local.shape[0] == r1 - r0
        if na >= r0 and nb <= r1:                      # everything is local (always so on one GPU): a view, no copy
            return local[na - r0:nb - r0]
        out = torch.empty((max(nb - na, 0),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        a, b = max(na, r0), min(nb, r1)
        if b > a:
            out[a - na:b - na].copy_(local[a - r0:b - r0])
        if not (self.active and self.world > 1):
            assert na >= r0 and nb <= r1, "rows outside the only strip requested"
            return out
        ops, keep = [], []
        for src, dst, pa, pb in exchange_plan(bounds, needs):
            if src == self.rank:
                buf = local[pa - r0:pb - r0].contiguous()
                keep.append(buf)
                ops.append(self.dist.P2POp(self.dist.isend, buf, dst, group=self.group))
            elif dst == self.rank:
                ops.append(self.dist.P2POp(self.dist.irecv, out[pa - na:pb - na], src, group=self.group))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        return out
