"""GPU idle gaps inside one multi-GPU bench step (rank 0 profiled with torch.profiler); run under torchrun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.dist import Comm, strip_bounds
from rs_image_segmentation_b200.synth import synth_strip_torch
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = Comm()
H = W = 7000
H_total = H * world
bounds = strip_bounds(H_total, world)
own = bounds[rank]
raster = synth_strip_torch(H_total, W, 7, own[0], own[1] - own[0], "uint8", seed=7000, device="cuda")
cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
def step():
    fr = P.extract_features(raster, cfg, comm, H_total, bounds)
    res, km, c0 = P.kmeans_on_features(fr, 13, 8, 20, 7000, comm, H_total, own[0])
    return res
for _ in range(3): step()
torch.cuda.synchronize(); dist.barrier()
ts = []
for _ in range(6):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
print(f"rank {rank} wall ms per step:", " ".join(f"{x:.1f}" for x in ts), flush=True)
dist.barrier()
if rank == 0:
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
else:
    step(); torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    print(f"span {(t1 - t0) / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, {len(evs)} device activities")
    end = evs[0].time_range.end; gaps = []
    for prev, e in zip(evs, evs[1:]):
        g = e.time_range.start - end
        if g > 15: gaps.append((g, prev.name[:50], e.name[:50]))
        end = max(end, e.time_range.end)
    print(f"total idle in gaps > 15 us: {sum(g[0] for g in gaps) / 1e3:.2f} ms over {len(gaps)} gaps")
    for g, a, b in sorted(gaps, reverse=True)[:14]: print(f"  {g:8.0f} us  after {a}  before {b}")
    long = sorted(evs, key=lambda e: -(e.time_range.end - e.time_range.start))[:8]
    for e in long: print(f"  long: {(e.time_range.end - e.time_range.start)/1e3:.2f} ms {e.name[:70]}")
dist.destroy_process_group()
