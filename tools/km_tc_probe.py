"""A/B of the K > 8 KMeans passes on one B200: fp32 FFMA2 kernel (km_tc=0) against the tensor-core kernel (km_tc=1) on the
config-B stack (49 Mpx, D = 13): ms per pass, near-tie counts, and bit-identity of labels, centroids and integer totals."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rs_image_segmentation_b200 import _lib, pipeline as P
from rs_image_segmentation_b200.device import StageTimer
from rs_image_segmentation_b200.synth import synth_strip_torch

H = W = int(os.environ.get("SIZE", 7000))
T = int(os.environ.get("T", 8))
D = 13
raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
fr = P.extract_features(raster, P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32))
mn, mx = fr.minmax.read()
for K in [int(k) for k in os.environ.get("KS", "16,32,64,9").split(",")]:
    out = {}
    for tc in (0, 1):
        _lib.set_option("km_tc", tc)
        for rep in range(2):
            timer = StageTimer(True)
            km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, fr.W, timer=timer)
            c0 = km.scale_rows(km.gather_rows(P.draw_init_indices(fr.n_px, K, 7000), 0))
            km.setup(c0)
            for it in range(T):
                km.step(track_labels=True)
            lab = km.finish(True)
            torch.cuda.synchronize()
        ev = timer.events.get("kmeans_assign_full", []) + timer.events.get("kmeans_assign_delta", [])
        ms = [a.elapsed_time(b) for a, b in ev]
        cent, shift, empty = km.read()
        n = km.n_acc
        out[tc] = dict(labels=lab.clone(), cent=cent, totals=km.acc[n:n + K * D + K].clone(), inertia=float(km.inertia.item()), ms=ms,
                       final=timer.totals_ms()["kmeans_final"][0], ties=km.near_ties())
    a, b = out[0], out[1]
    print(json.dumps({"K": K, "labels_equal": bool(torch.equal(a["labels"], b["labels"])), "n_label_diff": int((a["labels"] != b["labels"]).sum()),
                      "centroids_equal": bool(np.array_equal(a["cent"], b["cent"])), "totals_equal": bool(torch.equal(a["totals"], b["totals"])),
                      "inertia_rel_diff": abs(a["inertia"] - b["inertia"]) / max(abs(a["inertia"]), 1e-30),
                      "fp32_ms_per_pass": [round(m, 3) for m in a["ms"]], "tc_ms_per_pass": [round(m, 3) for m in b["ms"]],
                      "fp32_final_ms": round(a["final"], 3), "tc_final_ms": round(b["final"], 3), "near_ties_fp32": a["ties"], "near_ties_tc": b["ties"],
                      "near_tie_share_tc": b["ties"] / ((T + 1) * fr.n_px)}), flush=True)
