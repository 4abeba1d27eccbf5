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

import math
from typing import Callable, List, Sequence, Tuple

import torch

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


class Comm:
    """torch.distributed wrapper; `Comm()` without an initialised process group is the 1-GPU case."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.active = dist.is_available() and dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group) if self.active else 0
        self.world = dist.get_world_size(group) if self.active else 1

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
