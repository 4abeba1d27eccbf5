"""GPU idle gaps inside one bench step (torch.profiler / CUPTI): where does the device wait for the host?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.synth import synth_strip_torch
H = W = 7000
raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
def step():
    fr = P.extract_features(raster, cfg)
    res, km, c0 = P.kmeans_on_features(fr, 13, 8, 20, 7000)
    return res
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"span {(t1 - t0) / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, {len(evs)} device activities")
end = evs[0].time_range.end
gaps = []
for prev, e in zip(evs, evs[1:]):
    g = e.time_range.start - end
    if g > 15:
        gaps.append((g, prev.name[:60], e.name[:60]))
    end = max(end, e.time_range.end)
print(f"total idle in gaps > 15 us: {sum(g[0] for g in gaps) / 1e3:.2f} ms over {len(gaps)} gaps")
for g, a, b in sorted(gaps, reverse=True)[:25]:
    print(f"  {g:8.0f} us  after {a}  before {b}")
