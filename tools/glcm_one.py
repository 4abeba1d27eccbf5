"""One config-B feature extraction, repeated: per-stage times with default options (profiling target for the GLCM kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.device import StageTimer
from rs_image_segmentation_b200.synth import synth_strip_torch

H = W = int(os.environ.get("SIZE", 7000))
raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
for rep in range(int(os.environ.get("REPS", 2))):
    t = StageTimer(True)
    fr = P.extract_features(raster, cfg, timer=t)
    if os.environ.get("KM"):
        res, km, c0 = P.kmeans_on_features(fr, 13, 8, 20, 7000, timer=t)
    torch.cuda.synchronize()
    print({k: round(v[0], 3) for k, v in t.totals_ms().items()}, flush=True)
    del fr
