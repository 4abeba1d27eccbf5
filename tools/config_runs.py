"""Stage timings for the other BASELINE.json configurations on one B200:
  C  GLCM sweep on the 7000x7000 scene: windows 5/7/11, grey levels 16/32/64, dense (step 1)
  D  Sentinel-2-like tile 10980x10980x13 uint16: indices + PCA(6) + KMeans k=16, 20 iterations
  E  one GPU's share of the 40k x 40k x 7 mosaic at 8 GPUs (5000 x 40000 strip): full stack + KMeans k=32, 20 iterations
Prints one JSON line per run (CUDA-event stage times, ms)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rs_image_segmentation_b200 import _lib, pipeline as P
from rs_image_segmentation_b200.device import StageTimer, ptr, stream_ptr
from rs_image_segmentation_b200.synth import synth_strip_torch

which = sys.argv[1:] or ["C", "D", "E"]


def timed_step(fn, reps=3):
    fn(StageTimer(False))
    best = None
    for _ in range(reps):
        t = StageTimer(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(t); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if best is None or ms < best[0]:
            best = (ms, {k: round(v[0], 3) for k, v in t.totals_ms().items()})
    return best


if "C" in which:
    H = W = 7000
    raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
    fr = P.extract_features(raster, P.FeatureConfig(glcm=False))
    for levels in (16, 32, 64):
        for win in (5, 7, 11):
            cfg = P.FeatureConfig(glcm_window=win, glcm_step=1, glcm_levels=levels)
            def run(t, cfg=cfg):
                return P.extract_features(raster, cfg, timer=t)
            ms, st = timed_step(run, reps=2)
            n_win = (H - win + 1) * (W - win + 1)
            print(json.dumps({"config": "C", "window": win, "levels": levels, "glcm_props_ms": st.get("glcm_props"), "glcm_resize_ms": st.get("glcm_resize"),
                              "Mwindows_per_s": round(n_win / st["glcm_props"] / 1e3, 1), "feature_stack_ms": round(ms, 2)}), flush=True)
    del raster, fr
if "D" in which:
    H = W = 10980
    raster = synth_strip_torch(H, W, 13, 0, H, "uint16", seed=10980, device="cuda")
    cfg = P.FeatureConfig(band_map=(1, 2, 3, 7, 11), n_components=6, glcm=False)
    def run(t):
        fr = P.extract_features(raster, cfg, timer=t)
        return P.kmeans_on_features(fr, 13, 16, 20, 10980, timer=t)
    ms, st = timed_step(run)
    n = H * W
    print(json.dumps({"config": "D", "pixels": n, "ms_per_step": round(ms, 2), "Mpixel_per_s": round(n / ms / 1e3, 1),
                      "algorithmic_GBps": round(1200 * n / ms / 1e6, 1), "stages_ms": st}), flush=True)
    del raster
if "E" in which:
    H, W = 5000, 40000
    raster = synth_strip_torch(40000, W, 7, 0, H, "uint8", seed=40000, device="cuda")
    cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
    def run(t):
        fr = P.extract_features(raster, cfg, timer=t)
        return P.kmeans_on_features(fr, 13, 32, 20, 40000, timer=t)
    ms, st = timed_step(run)
    n = H * W
    print(json.dumps({"config": "E (one of 8 strips, no halo exchange)", "pixels": n, "ms_per_step": round(ms, 2), "Mpixel_per_s": round(n / ms / 1e3, 1),
                      "algorithmic_GBps": round(1149 * n / ms / 1e6, 1), "frac_of_6548": round(1149 * n / ms / 1e6 / 6548.2, 3), "stages_ms": st}), flush=True)
