"""How much would Hamerly-style bounds let the KMeans passes skip?  Simulation on config B (torch float64 distances)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.synth import synth_strip_torch
H = W = int(os.environ.get("SIZE", 7000)); K = int(os.environ.get("K", 8)); T = int(os.environ.get("T", 20)); D = 13
raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
fr = P.extract_features(raster, P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32))
mn, mx = fr.minmax.read()
km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, fr.W)
c0 = km.scale_rows(km.gather_rows(P.draw_init_indices(fr.n_px, K, 7000), 0))
n = fr.n_px
scale = torch.from_numpy(km.scale).cuda().float(); min_ = torch.from_numpy(km.min_).cuda().float()
C = torch.from_numpy(c0).cuda().float()
def dists(C):
    best = torch.full((n,), 1e30, device="cuda"); second = torch.full((n,), 1e30, device="cuda"); lab = torch.zeros(n, dtype=torch.int64, device="cuda")
    chunk = 1 << 22
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        X = fr.planes[:D, a:b].t() * scale + min_
        d = torch.cdist(X, C)
        v, i = torch.topk(d, 2, dim=1, largest=False)
        best[a:b], second[a:b], lab[a:b] = v[:, 0], v[:, 1], i[:, 0]
    return best, second, lab
u, l, lab = dists(C)
for it in range(T):
    # update
    Cn = torch.zeros_like(C); cnt = torch.bincount(lab, minlength=K).float()
    for a in range(0, n, 1 << 22):
        b = min(n, a + (1 << 22)); X = fr.planes[:D, a:b].t() * scale + min_
        Cn.index_add_(0, lab[a:b], X)
    Cn /= cnt[:, None].clamp(min=1)
    delta = (Cn - C).norm(dim=1); C = Cn
    dsort, didx = torch.sort(delta, descending=True)
    u = u + delta[lab]
    l = l - torch.where(lab == didx[0], dsort[1], dsort[0])
    need = ~(u < l)
    f_px = need.float().mean().item()
    npad = (n + 511) // 512 * 512
    nd = torch.zeros(npad, dtype=torch.bool, device="cuda"); nd[:n] = need
    f512 = nd.view(-1, 512).any(1).float().mean().item(); f128 = nd.view(-1, 128).any(1).float().mean().item()
    # evaluate where needed (here: everywhere, then keep exact bounds only for the needed ones to mimic Hamerly)
    ub, lb, labn = dists(C)
    changed = (labn != lab).float().mean().item()
    wrong = ((labn != lab) & ~need).sum().item()
    u = torch.where(need, ub, u); l = torch.where(need, lb, l); lab = torch.where(need, labn, lab)
    print(f"it {it + 2}: need px {f_px:.4f}  blocks512 {f512:.4f}  segments128 {f128:.4f}  changed {changed:.4f}  skipped-but-changed {wrong}", flush=True)
