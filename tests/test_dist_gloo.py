"""Host-side sharding logic (rs_image_segmentation_b200/dist.py) on CPU: strip bounds, GLCM halo requirements, the
exchange plan, and - with world_size 2 and 3 over gloo - the collective `Comm.fetch_rows` / `Comm.all_reduce`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from rs_image_segmentation_b200 import dist as D


def test_strip_bounds_cover_and_align():
    for H, G, align in [(7000, 8, 1), (600, 4, 21), (13, 8, 1), (40000, 8, 1), (5, 8, 1), (100, 3, 7)]:
        b = D.strip_bounds(H, G, align)
        assert len(b) == G and b[0][0] == 0 and b[-1][1] == H
        for (a0, a1), (b0, b1) in zip(b, b[1:]):
            assert a1 == b0 and a0 <= a1
        for r0, r1 in b[:-1]:
            assert r1 == H or (r1 - r0) % align == 0


def _cv_rows(dst_h, src_h, r0, r1):
    """rows cv2.resize(INTER_LINEAR) touches for destination rows [r0, r1): brute force from its coordinate rule."""
    scale = 1.0 / (dst_h / src_h)
    used = set()
    for dy in range(r0, r1):
        s = int(np.floor((dy + 0.5) * scale - 0.5))
        used.add(min(max(s, 0), src_h - 1))
        used.add(min(max(s + 1, 0), src_h - 1))
    return used


@pytest.mark.parametrize("H,w,s,G", [(600, 21, 21, 4), (97, 7, 1, 2), (233, 11, 1, 3), (64, 5, 2, 8), (7000, 7, 1, 8)])
def test_glcm_rows_needed_cover_resize_footprint(H, w, s, G):
    out_rows = (H - w) // s + 1
    for own in D.strip_bounds(H, G):
        (p0, p1), (q0, q1) = D.glcm_rows_needed(own, H, w, s)
        if own[1] <= own[0]:
            assert p1 <= p0
            continue
        used = _cv_rows(H, out_rows, own[0], own[1])
        assert p0 <= min(used) and max(used) < p1                    # every property row the upsample reads is computed
        assert q0 == p0 * s and q1 == (p1 - 1) * s + w and q1 <= H   # and the windows of those rows are covered


def test_exchange_plan_is_exact():
    bounds = D.strip_bounds(100, 4)
    needs = [(0, 31), (25, 56), (50, 81), (70, 100)]
    plan = D.exchange_plan(bounds, needs)
    for dst, (na, nb) in enumerate(needs):
        got = set(range(max(na, bounds[dst][0]), min(nb, bounds[dst][1])))
        for src, d, a, b in plan:
            if d == dst:
                assert bounds[src][0] <= a < b <= bounds[src][1]
                assert not got & set(range(a, b))
                got |= set(range(a, b))
        assert got == set(range(na, nb))


# ------------------------------------------------------------------------------------------- gloo, world_size > 1
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, w, s, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = D.Comm()
        assert comm.rank == rank and comm.world == world
        full = torch.arange(H * W, dtype=torch.int32).reshape(H, W).to(torch.uint8)        # the "quantised band"
        bounds = D.strip_bounds(H, world)
        r0, r1 = bounds[rank]
        needs = [D.glcm_rows_needed(b, H, w, s)[1] for b in bounds]
        got = comm.fetch_rows(full[r0:r1].clone(), bounds, needs)
        na, nb = needs[rank]
        ok_rows = bool(torch.equal(got, full[na:nb]))
        # integer partial sums: the all-reduce of per-strip accumulators equals the whole-image accumulator
        acc = full[r0:r1].to(torch.int64).sum(dim=0)
        comm.all_reduce(acc)
        ok_sum = bool(torch.equal(acc, full.to(torch.int64).sum(dim=0)))
        mn = torch.tensor([float(r0)])
        mx = torch.tensor([float(r1)])
        comm.all_reduce(mn, "min")
        comm.all_reduce(mx, "max")
        ok_mm = mn.item() == 0.0 and mx.item() == float(H)
        # initial-centroid rows: every rank ends up with the rows of all requested global pixels
        from rs_image_segmentation_b200.pipeline import gather_rows_device
        planes_full = torch.arange(3 * H * W, dtype=torch.float32).reshape(3, H * W)
        idx = np.array([0, W * bounds[0][1] - 1, H * W - 1, (H // 2) * W + 3], dtype=np.int64)   # the same on every rank
        rows = gather_rows_device(planes_full[:, r0 * W:r1 * W].contiguous(), 3, (r1 - r0) * W, idx, r0 * W, comm)
        ok_gather = bool(torch.equal(rows, planes_full[:, torch.from_numpy(idx)].t().to(torch.float64)))
        comm.barrier()
        ret[rank] = (ok_rows, ok_sum, ok_mm, ok_gather)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W,w,s", [(2, 97, 33, 7, 1), (3, 100, 17, 11, 1), (2, 84, 21, 21, 21)])
def test_fetch_rows_and_allreduce_gloo(world, H, W, w, s):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, w, s, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    for p in procs:
        assert p.exitcode == 0, "a gloo worker failed"
    assert len(ret) == world and all(all(v) for v in ret.values()), dict(ret)


def test_peer_blocks_sequence_and_layout():
    """dist.PeerBlocks: pass buffers alternate with the parity of the sequence number; the layout constants are those of
    include/rsx.h; a CPU / single-process Comm never asks for peer memory."""
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "rsx.h")).read()
    elems = int(re.search(r"#define RSX_PEER_PASS_ELEMS (\d+)", hdr).group(1))
    max_k = int(re.search(r"#define RSX_MAX_CLUSTERS (\d+)", hdr).group(1))
    max_d = int(re.search(r"#define RSX_MAX_FEATURES (\d+)", hdr).group(1))
    max_p = int(re.search(r"#define RSX_MAX_PEERS (\d+)", hdr).group(1))
    assert elems == D.PEER_PASS_ELEMS and elems >= max_k * (max_d + 1) + 2
    assert D.PEER_BLOCK_BYTES == 2 * elems * 8 + max_p * 8
    pb = D.PeerBlocks(0x10000, None, 1, 2)
    seqs = [pb.next_pass() for _ in range(4)]
    assert [s for s, _ in seqs] == [1, 2, 3, 4]
    assert [p.value - 0x10000 for _, p in seqs] == [elems * 8, 0, elems * 8, 0]
    assert D.Comm().peers() is None
