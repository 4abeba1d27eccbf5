"""Multi-GPU self check: the row-strip sharded path must reproduce the single-GPU result bit for bit (feature planes,
level-1 context planes, initial centroids, centroids, labels; inertia to 1e-12 relative).

Run collectively by every rank of an initialised NCCL process group (tools/mgpu_check.py under torchrun,
tests/test_gpu_multi.py, and `bench.py --gpus N` which reports the outcome as "mgpu_parity").  The reference has no
multi-GPU path: the sharded result is checked against this repo's own single-GPU run, which the parity
tests tie to the reference / oracle.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import pipeline as P
from .dist import Comm, strip_bounds
from .synth import synth_raster_numpy

CASES = [dict(H=301, W=517, B=7, dtype=np.uint8, win=7, step=1, K=8, D=13, T=6),
         dict(H=256, W=300, B=7, dtype=np.uint8, win=21, step=21, K=5, D=13, T=4),
         dict(H=203, W=260, B=13, dtype=np.uint16, win=0, step=0, K=16, D=13, T=5),
         dict(H=97, W=1100, B=7, dtype=np.uint8, win=11, step=1, K=32, D=13, T=4)]


def single_rank_comm(dist) -> Comm:
    """A 1-rank communicator although a process group exists (rank 0's reference run)."""
    one = Comm.__new__(Comm)
    one.dist, one.active, one.group, one.rank, one.world = dist, False, None, 0, 1
    return one


def sharded_equals_single(comm: Comm, cases: Optional[List[dict]] = None, log=None) -> List[str]:
    """Returns the list of failures as seen by rank 0 (empty = bit-identical); every rank must call it."""
    dist = comm.dist
    rank, world = comm.rank, comm.world
    failures: List[str] = []
    for c in (cases if cases is not None else CASES):
        H, W = c["H"], c["W"]
        full = synth_raster_numpy(H, W, c["B"], c["dtype"], seed=H + W, cell=16)
        if c["dtype"] == np.uint16:
            cfg = P.FeatureConfig(band_map=(1, 2, 3, 7, 11), n_components=6, glcm=False)
            full_t = full.view(np.int16)
        else:
            cfg = P.FeatureConfig(glcm_window=c["win"], glcm_step=c["step"])
            full_t = full
        align = c["step"] if c["step"] > 1 else 1
        bounds = strip_bounds(H, world, align)
        r0, r1 = bounds[rank]
        strip = torch.from_numpy(np.ascontiguousarray(full_t[r0:r1])).cuda()
        fr = P.extract_features(strip, cfg, comm, H, bounds)
        res, km, c0 = P.kmeans_on_features(fr, c["D"], c["K"], c["T"], seed=3, comm=comm, H_total=H, first_row=r0)
        lab = torch.zeros(H * W, dtype=torch.int32, device="cuda")
        lab[r0 * W:r1 * W] = res.labels
        comm.all_reduce(lab)
        planes = torch.zeros((len(fr.names), H * W), dtype=torch.float32, device="cuda")
        planes[:, r0 * W:r1 * W] = fr.planes[:, :fr.n_px]
        comm.all_reduce(planes)
        l1, l1_names, _ = P.level1_with_context(fr, 7, comm, H, bounds)     # 3 halo rows from the neighbouring strips
        ctx = torch.zeros((len(l1_names), H * W), dtype=torch.float32, device="cuda")
        ctx[:, r0 * W:r1 * W] = l1[:, :fr.n_px]
        comm.all_reduce(ctx)
        if rank == 0:
            one = single_rank_comm(dist)
            fr1 = P.extract_features(torch.from_numpy(np.ascontiguousarray(full_t)).cuda(), cfg, one)
            res1, km1, c01 = P.kmeans_on_features(fr1, c["D"], c["K"], c["T"], seed=3, comm=one)
            name = f"{H}x{W}x{c['B']} win={c['win']} step={c['step']} K={c['K']}"
            before = len(failures)
            same = torch.isclose(planes, fr1.planes[:, :fr1.n_px], rtol=0, atol=0, equal_nan=True)
            if not bool(same.all()):
                bad = [n for i, n in enumerate(fr1.names) if not bool(same[i].all())]
                failures.append(f"{name}: feature planes differ: {bad}")
            l11, _, _ = P.level1_with_context(fr1, 7, one)
            if not bool(torch.isclose(ctx, l11[:, :fr1.n_px], rtol=0, atol=0, equal_nan=True).all()):
                failures.append(f"{name}: level-1 context planes differ")
            if not np.array_equal(c0, c01):
                failures.append(f"{name}: initial centroids differ")
            if not np.array_equal(res.centroids, res1.centroids):
                failures.append(f"{name}: centroids differ by {np.abs(res.centroids - res1.centroids).max()}")
            if not torch.equal(lab, res1.labels):
                failures.append(f"{name}: {(lab != res1.labels).sum().item()} labels differ")
            if abs(res.inertia - res1.inertia) > 1e-12 * abs(res1.inertia):
                failures.append(f"{name}: inertia {res.inertia} vs {res1.inertia}")
            if log is not None:
                log(f"[mgpu] {name}: world={world} ok={len(failures) == before} inertia={res.inertia:.9f} near_ties={res.near_ties}")
        comm.barrier()
    return failures
