"""Large-raster check (> 2^31 stack elements, > 2^32 bytes per plane set): size-independent properties of the whole path on one
GPU - histogram totals, label/count consistency, delta passes == full passes, inertia == recomputed inertia on a sample."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rs_image_segmentation_b200 import pipeline as P
from rs_image_segmentation_b200.synth import synth_strip_torch

H, W = int(os.environ.get("H", 16000)), int(os.environ.get("W", 40000))
K, T, D = int(os.environ.get("K", 8)), int(os.environ.get("T", 4)), 13
n = H * W
print(f"raster {H}x{W}x7 = {n / 1e6:.0f} Mpx; stack elements {n * D / 2**31:.2f} x 2^31", flush=True)
raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=40000, device="cuda")
t0 = time.time()
fr = P.extract_features(raster, P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32))
torch.cuda.synchronize()
print(f"features {time.time() - t0:.2f} s; free mem {torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB", flush=True)
assert int(fr.stats.hist[0].sum()) == n
# spot checks of planes far beyond the 2^31-th element: recompute a few pixels on the host from the raster
from oracle import features as of
idx = np.array([0, n // 2 + 12345, n - 1, n - W - 7, (H - 3) * W + 5], dtype=np.int64)
rows = raster.view(-1, 7)[torch.from_numpy(idx).cuda()].cpu().numpy().astype(np.float32)
norm = fr.stats.norm
nb = [(np.clip(rows[:, b], norm[b, 0], norm[b, 1]) - norm[b, 0]) / norm[b, 2] for b in range(7)]
ix = of.all_indices([x.reshape(-1, 1) for x in nb])
for k in P.INDEX_NAMES:
    got = fr.planes[fr.names.index(k)][torch.from_numpy(idx).cuda()].cpu().numpy()
    assert np.array_equal(got, ix[k].ravel()), k
out = {}
for delta in (True, False):
    res, km, c0 = P.kmeans_on_features(fr, D, K, T, seed=40000, delta=delta, labels_i32=False)
    tot = km.acc[km.n_acc:km.n_acc + K * D + K].cpu().numpy()
    counts = tot[K * D:]
    lab = res.labels
    assert int(counts.sum()) == n
    bc = torch.bincount(lab.to(torch.int64), minlength=K).cpu().numpy()
    out[delta] = (tot, res.centroids, res.inertia, bc)
    print(f"delta={delta}: inertia {res.inertia:.6f} near_ties {res.near_ties} counts {counts.tolist()}", flush=True)
    del res, lab
assert np.array_equal(out[True][0], out[False][0]) and np.array_equal(out[True][1], out[False][1])
assert abs(out[True][2] - out[False][2]) <= 1e-12 * out[True][2]
print("LARGE CHECK OK")
