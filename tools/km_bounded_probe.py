"""Per-pass timing of the KMeans loop on one B200, bounded (Hamerly test + pixel-interleaved gather) against unbounded delta passes,
and a bit-identity check of labels / centroids / inertia between the two.

    python tools/km_bounded_probe.py [size] [K] [T]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rs_image_segmentation_b200 import _lib, pipeline as P
from rs_image_segmentation_b200.synth import synth_strip_torch

size = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
D = 13
raster = synth_strip_torch(size, size, 7, 0, size, "uint8", seed=7000, device="cuda")
cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
fr = P.extract_features(raster, cfg)
del raster
P.kmeans_on_features(fr, D, K, 1, 7000)          # NaN pass etc.
mn, mx = fr.minmax.read()


def run(bounded, full_passes, q16_from=-1):
    _lib.set_option("km_q16", 1 if q16_from >= 0 else 0)
    _lib.set_option("km_q16_from", max(q16_from, 0))
    km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, fr.W, bounded=bounded, full_passes=full_passes)
    c0 = km.scale_rows(km.gather_rows(P.draw_init_indices(fr.n_px, K, 7000), 0))
    for rep in range(2):
        km.setup(c0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
        changed = []
        ev[0].record()
        for it in range(T):
            mode = km.assign_pass()
            changed.append(km.acc[km.n_acc - 1:km.n_acc].clone())
            km.update(mode)
            ev[it + 1].record()
        torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(T)]
    lab = km.finish(True)
    res = km._result(lab, T)
    return ms, [int(c.item()) for c in changed], res


base = None
FP = [int(v) for v in os.environ.get("PROBE_FULL_PASSES", "3").split(",") if v]
UFP = [int(v) for v in os.environ.get("PROBE_UNBOUNDED_FULL_PASSES", "").split(",") if v]
Q16 = [int(v) for v in os.environ.get("PROBE_Q16_FROM", "").split(",") if v]
for name, bounded, fp, qf in [("unbounded", False, 3, -1)] + [(f"unbounded_fp{v}", False, v, -1) for v in UFP] + [(f"bounded_fp{v}", True, v, -1) for v in FP] + \
        [(f"q16_from{v}", False, 3, v) for v in Q16]:
    ms, changed, res = run(bounded, fp, qf)
    out = {"variant": name, "sum_ms": round(sum(ms), 3), "ms": [round(m, 3) for m in ms], "inertia": res.inertia, "near_ties": res.near_ties}
    if base is None:
        base = res
        out["changed"] = changed
    else:
        out["labels_equal"] = bool(torch.equal(res.labels, base.labels))
        out["centroids_equal"] = bool(np.array_equal(res.centroids, base.centroids))
        out["inertia_equal"] = res.inertia == base.inertia
    print(json.dumps(out), flush=True)
