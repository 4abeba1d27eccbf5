"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-strip sharded path must reproduce the
single-GPU results bit for bit - feature planes, quantised band, centroids, labels, inertia to 1e-12 relative.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.dist import Comm, strip_bounds
from rs_image_segmentation_b200.synth import synth_raster_numpy


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm()
    failures = []
    cases = [dict(H=301, W=517, B=7, dtype=np.uint8, win=7, step=1, K=8, D=13, T=6),
             dict(H=256, W=300, B=7, dtype=np.uint8, win=21, step=21, K=5, D=13, T=4),
             dict(H=203, W=260, B=13, dtype=np.uint16, win=0, step=0, K=16, D=13, T=5),
             dict(H=97, W=1100, B=7, dtype=np.uint8, win=11, step=1, K=32, D=13, T=4)]
    for c in cases:
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
        # gather the strips on rank 0
        lab = torch.zeros(H * W, dtype=torch.int32, device="cuda")
        lab[r0 * W:r1 * W] = res.labels
        dist.all_reduce(lab)
        planes = torch.zeros((len(fr.names), H * W), dtype=torch.float32, device="cuda")
        planes[:, r0 * W:r1 * W] = fr.planes[:, :fr.n_px]
        dist.all_reduce(planes)
        l1, l1_names, _ = P.level1_with_context(fr, 7, comm, H, bounds)     # 3 halo rows from the neighbouring strips
        ctx = torch.zeros((len(l1_names), H * W), dtype=torch.float32, device="cuda")
        ctx[:, r0 * W:r1 * W] = l1[:, :fr.n_px]
        dist.all_reduce(ctx)
        if rank == 0:
            one = Comm.__new__(Comm)                        # a 1-rank communicator although a process group exists
            one.dist, one.active, one.group, one.rank, one.world = dist, False, None, 0, 1
            fr1 = P.extract_features(torch.from_numpy(np.ascontiguousarray(full_t)).cuda(), cfg, one)
            res1, km1, c01 = P.kmeans_on_features(fr1, c["D"], c["K"], c["T"], seed=3, comm=one)
            name = f"{H}x{W}x{c['B']} win={c['win']} step={c['step']} K={c['K']}"
            if not torch.equal(planes, fr1.planes[:, :fr1.n_px]):
                bad = [n for i, n in enumerate(fr1.names) if not torch.equal(planes[i], fr1.planes[i, :fr1.n_px])]
                failures.append(f"{name}: feature planes differ: {bad}")
            l11, _, _ = P.level1_with_context(fr1, 7, one)
            if not torch.equal(ctx, l11[:, :fr1.n_px]):
                failures.append(f"{name}: level-1 context planes differ")
            if not np.array_equal(c0, c01):
                failures.append(f"{name}: initial centroids differ")
            if not np.array_equal(res.centroids, res1.centroids):
                failures.append(f"{name}: centroids differ by {np.abs(res.centroids - res1.centroids).max()}")
            if not torch.equal(lab, res1.labels):
                failures.append(f"{name}: {(lab != res1.labels).sum().item()} labels differ")
            if abs(res.inertia - res1.inertia) > 1e-12 * abs(res1.inertia):
                failures.append(f"{name}: inertia {res.inertia} vs {res1.inertia}")
            print(f"[mgpu] {name}: world={world} ok={not failures} inertia={res.inertia:.9f} near_ties={res.near_ties}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        if failures:
            print("\n".join(failures))
            sys.exit(1)
        print("MGPU PARITY OK")


if __name__ == "__main__":
    main()
