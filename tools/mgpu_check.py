"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-strip sharded path must reproduce the
single-GPU results bit for bit - feature planes, quantised band, centroids, labels, inertia to 1e-12 relative
(rs_image_segmentation_b200/selfcheck.py holds the cases; bench.py --gpus N runs the same check as "mgpu_parity").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from rs_image_segmentation_b200.dist import Comm
from rs_image_segmentation_b200.selfcheck import sharded_equals_single


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = sharded_equals_single(Comm(), log=lambda m: print(m, flush=True))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        if failures:
            print("\n".join(failures))
            sys.exit(1)
        print("MGPU PARITY OK")


if __name__ == "__main__":
    main()
