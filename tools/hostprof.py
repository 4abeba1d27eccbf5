import sys, time, cProfile, pstats, io
sys.path.insert(0, '/root/repo')
import torch
from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.synth import synth_strip_torch
H=W=7000
raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
def step():
    fr = P.extract_features(raster, cfg)
    res, km, c0 = P.kmeans_on_features(fr, 13, 8, 20, 7000)
    return res
for _ in range(3): step()
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(3): step()
torch.cuda.synchronize()
print('wall ms/step', (time.perf_counter()-t)/3*1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(35); print(s.getvalue()[:6000])
