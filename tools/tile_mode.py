import sys; sys.path.insert(0, '/root/repo')
import torch, json
from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.device import StageTimer
from rs_image_segmentation_b200.synth import synth_strip_torch
H = W = 7000
raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
for win, step in ((21, 21), (21, 7), (7, 7), (11, 4)):
    cfg = P.FeatureConfig(glcm_window=win, glcm_step=step)
    P.extract_features(raster, cfg)
    t = StageTimer(True); P.extract_features(raster, cfg, timer=t); torch.cuda.synchronize()
    st = {k: round(v[0], 3) for k, v in t.totals_ms().items()}
    print(win, step, st["glcm_props"], st["glcm_resize"])
