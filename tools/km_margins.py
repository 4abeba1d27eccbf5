"""Change rate per Lloyd pass and the distribution of best/second-best distance margins on the bench scene
(sizing data for a reduced-precision pre-filter).  usage: python tools/km_margins.py [size]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.synth import synth_strip_torch

S = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
raster = synth_strip_torch(S, S, 7, 0, S, "uint8", seed=7000, device="cuda")
cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
fr = P.extract_features(raster, cfg)
D, K = 13, 8
idx = P.draw_init_indices(S * S, K, 7000)
km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, None, None, S * S, fr.W)
rows = P.gather_rows_device(fr.planes, D, fr.n_px, idx, 0, km.comm)
mn, mx = fr.minmax.read()
km.configure(mn[:D], mx[:D])
c0 = km.scale_rows(rows.cpu().numpy())
km.setup(c0)
scale, min_ = torch.from_numpy(km.scale).cuda(), torch.from_numpy(km.min_).cuda()
n = fr.n_px
sel = torch.randint(0, n, (1 << 21,), device="cuda")
Xs = fr.planes[:D].index_select(1, sel).t().to(torch.float64) * scale + min_
for it in range(20):
    km.step(True)
    ch = km.changed_count()
    cent, _, _ = km.read()
    C = torch.from_numpy(cent).cuda()
    d = ((Xs[:, None, :] - C[None]) ** 2).sum(-1)
    top = torch.topk(d, 2, dim=1, largest=False).values
    m = (top[:, 1] - top[:, 0])
    l1 = (C[:, None, :] - C[None]).abs().sum(-1).max().item()
    fr_ = [(m < t).double().mean().item() for t in (1e-5, 3e-5, 1e-4, 3e-4, 1e-3, 3e-3)]
    print(f"pass {it + 1:2d} changed {ch / n * 100:7.3f}%  max|ca-cb|_1 {l1:.3f}  P(margin<1e-5,3e-5,1e-4,3e-4,1e-3,3e-3) = "
          + " ".join(f"{x * 100:.3f}%" for x in fr_), flush=True)
