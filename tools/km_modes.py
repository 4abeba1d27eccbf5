"""KMeans pass timing per iteration for config B: changed fraction and ms per pass, full vs delta passes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.device import StageTimer
from rs_image_segmentation_b200.synth import synth_strip_torch
H = W = int(os.environ.get("SIZE", 7000))
K = int(os.environ.get("K", 8)); T = int(os.environ.get("T", 20)); D = 13
raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
fr = P.extract_features(raster, P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32))
for delta in (False, True):
    for rep in range(2):
        timer = StageTimer(True)
        mn, mx = fr.minmax.read()
        km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, int(os.environ.get('ROWLEN', fr.W)), timer=timer, delta=delta)
        c0 = km.scale_rows(km.gather_rows(P.draw_init_indices(fr.n_px, K, 7000), 0))
        km.setup(c0)
        ch = []
        for it in range(T):
            km.step(track_labels=True)
            ch.append(km.changed_count() / fr.n_px)
        lab = km.finish(True)
        torch.cuda.synchronize()
    ev = timer.events.get("kmeans_assign_full", []) + timer.events.get("kmeans_assign_delta", [])
    ms = [a.elapsed_time(b) for a, b in ev]
    print(f"delta={delta} total assign {sum(ms):.2f} ms; final {timer.totals_ms()['kmeans_final'][0]:.3f} ms")
    print("  ms/pass:", " ".join(f"{m:.3f}" for m in ms))
    print("  near ties:", km.near_ties())
    print("  changed:", " ".join(f"{c:.4f}" for c in ch))
