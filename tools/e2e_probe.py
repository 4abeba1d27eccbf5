"""Per-scene intervals of pipeline.segment_stream (steady state vs fill/drain).  usage: python tools/e2e_probe.py [n_scenes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rs_image_segmentation_b200 import pipeline as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
H = W = 7000
from rs_image_segmentation_b200.synth import synth_strip_torch
pinned = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda").cpu().pin_memory()
cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
for rep in range(2):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    for labels, res in P.segment_stream((pinned for _ in range(n)), cfg, 8, 20, 7000, 13):
        t.append(time.perf_counter())
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    del labels, res
    d = np.diff(np.array(t)) * 1e3
    print("rep", rep, "total/scene %.2f ms" % ((t[-1] - t[0]) * 1e3 / n), "intervals", np.round(d, 2).tolist(), flush=True)
