"""Small end-to-end invocation of every rsx kernel for compute-sanitizer (memcheck / racecheck / synccheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rs_image_segmentation_b200 import pipeline as P, indices as I
from rs_image_segmentation_b200.synth import synth_raster_numpy

for (H, W, win, step, K, D) in [(75, 131, 7, 1, 5, 13), (90, 70, 21, 21, 12, 9), (64, 203, 11, 1, 8, 13)]:
    bip = synth_raster_numpy(H, W, 7, np.uint8, H, cell=16)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), P.FeatureConfig(glcm_window=win, glcm_step=step))
    res, km, c0 = P.kmeans_on_features(fr, D, K, 4, seed=1)
    l1, names, mm = P.level1_with_context(fr)
    c0[K - 1] = c0[0]                                       # empty cluster -> relocation path
    mn, mx = fr.minmax.read()
    km2 = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, fr.W)
    r2 = km2.fit_converge(c0, max_iter=6, tol=0.0)
    torch.cuda.synchronize()
    print(H, W, win, step, K, "ok", res.inertia, r2.n_iter, flush=True)
bip16 = synth_raster_numpy(60, 90, 13, np.uint16, 5, cell=16)
fr = P.extract_features(torch.from_numpy(bip16.view(np.int16)).cuda(), P.FeatureConfig(band_map=(1, 2, 3, 7, 11), n_components=6, glcm=False))
res, km, c0 = P.kmeans_on_features(fr, 13, 16, 3, seed=1)
x = np.random.default_rng(0).random((50, 60), dtype=np.float32)
I.calculate_glcm_features(x, levels=16, window_size=5, step_size=3)
I.robust_normalize(x); I.calculate_evi(x, x * 0.5, x * 0.25); I.calculate_msavi(x, x * 0.5); I.calculate_bsi(x, x, x, x)
torch.cuda.synchronize()
print("SANITIZE CASE DONE")
