"""How large do the per-cell pair counts of a dense 7x7 GLCM window get?  (Sizing data for narrower private counters in
glcm_dense_kernel, whose uint8 counters - 528 B per window and angle at 32 levels - are what limits its occupancy.)
CPU, numpy; synthetic scene as in bench.py.  usage: python tools/glcm_count_stats.py [size] [window] [levels]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rs_image_segmentation_b200.synth import synth_raster_numpy

S = int(sys.argv[1]) if len(sys.argv) > 1 else 600
w = int(sys.argv[2]) if len(sys.argv) > 2 else 7
L = int(sys.argv[3]) if len(sys.argv) > 3 else 32
r = synth_raster_numpy(S, S, 7, np.uint8, seed=7000)
nir = r[..., 3].astype(np.float32)
lo, hi = np.percentile(nir, 2), np.percentile(nir, 98)
q = (np.clip((nir - lo) / (hi - lo + 1e-10), 0, 1) * (L - 1)).astype(np.uint8)     # close enough to the two-step normalisation
rng = np.random.default_rng(1)
offs = {0: (0, 1), 1: (1, 1), 2: (1, 0), 3: (1, -1)}
hist_max = np.zeros(64, np.int64)
distinct = []
n_win = 20000
for _ in range(n_win):
    i, j = rng.integers(0, S - w), rng.integers(0, S - w)
    win = q[i:i + w, j:j + w].astype(np.int64)
    for ang, (dr, dc) in offs.items():
        a = win[:w - dr, max(0, -dc):w - max(0, dc)]
        b = win[dr:, max(0, dc):w + min(0, dc)]
        lo_, hi_ = np.minimum(a, b).ravel(), np.maximum(a, b).ravel()
        cells, counts = np.unique(hi_ * (hi_ + 1) // 2 + lo_, return_counts=True)
        hist_max[counts.max()] += 1
        distinct.append(len(cells))
tot = hist_max.sum()
cum = np.cumsum(hist_max[::-1])[::-1]
print(f"{n_win} random {w}x{w} windows x 4 angles at {L} levels: distinct cells per (window, angle): mean {np.mean(distinct):.1f}, max {np.max(distinct)}")
for t in (4, 8, 16, 32):
    print(f"  share of (window, angle) whose largest cell count is >= {t}: {100 * cum[t] / tot:.3f} %")

# Folded counters: cell (a, b) -> (a mod M, b mod M) with a tag for the high parts; a (window, angle) is "clean" when no two
# present cells share a folded slot.  How often would a fold conflict force the exact fallback?
for M in (8, 16):
    bad = 0
    span = []
    rng = np.random.default_rng(1)
    for _ in range(n_win):
        i, j = rng.integers(0, S - w), rng.integers(0, S - w)
        win = q[i:i + w, j:j + w].astype(np.int64)
        span.append(int(win.max() - win.min()))
        for ang, (dr, dc) in offs.items():
            a = win[:w - dr, max(0, -dc):w - max(0, dc)].ravel()
            b = win[dr:, max(0, dc):w + min(0, dc)].ravel()
            lo_, hi_ = np.minimum(a, b), np.maximum(a, b)
            full = np.unique(hi_ * 64 + lo_)
            fl, fh = (full % 64) % M, (full // 64) % M
            folded = np.unique(np.maximum(fl, fh) * 64 + np.minimum(fl, fh))
            bad += len(folded) != len(full)
    print(f"fold mod {M}: {100 * bad / (4 * n_win):.2f} % of (window, angle) have a conflict; level span of a window: mean {np.mean(span):.1f}, "
          f"share >= {M}: {100 * np.mean(np.array(span) >= M):.2f} %")
