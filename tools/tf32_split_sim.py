"""Numerical model (CPU, numpy) of the planned two-term TF32 distance GEMM for K > 8: the error of
A_hi.W_hi + A_hi.W_lo + A_lo.W_hi against float64, the share of pixels inside the widened near-tie band, and the same for a
single TF32 term.  kind::tf32 keeps 10 mantissa bits of an fp32 operand (truncation of the low 13).
usage: python tools/tf32_split_sim.py [n_px] [K] [D]"""
import sys
import numpy as np

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
D = int(sys.argv[3]) if len(sys.argv) > 3 else 13
rng = np.random.default_rng(7)


def trunc13(a):
    return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


# correlated features scaled to [0, 1] (MinMaxScaler), centred at 0.5 as DeviceKMeans.setup does
base = rng.normal(size=(n, 4))
X = base @ rng.normal(size=(4, D)) + 0.4 * rng.normal(size=(n, D))
X = ((X - X.min(0)) / (X.max(0) - X.min(0))).astype(np.float32)
C = X[rng.choice(n, K, replace=False)].astype(np.float64)
for _ in range(6):                                    # a few Lloyd iterations so the centroids sit where they would
    d = ((X[:, None, :].astype(np.float64) - C[None]) ** 2).sum(-1)
    lab = d.argmin(1)
    C = np.stack([X[lab == k].mean(0) if (lab == k).any() else C[k] for k in range(K)])
Xc = X.astype(np.float64) - 0.5
Cc = C - 0.5
W = (-2.0 * Cc).astype(np.float32)                    # dist_j = |c_j|^2 - 2 x.c_j  (|x|^2 is common to all j)
bias = (Cc ** 2).sum(1)
exact = bias[None] + Xc @ (-2.0 * Cc).T
Xf = Xc.astype(np.float32)
xh, wh = trunc13(Xf), trunc13(W)
xl, wl = trunc13(Xf - xh), trunc13(W - wh)            # the low terms are truncated by the tensor core as well


def mm(a, b):                                         # fp32 accumulation (order unspecified in hardware: model with float32 matmul)
    return (a.astype(np.float32) @ b.astype(np.float32).T).astype(np.float64)


one = bias[None] + mm(xh, wh)
three = bias[None] + (mm(xh, wh) + mm(xh, wl) + mm(xl, wh))
fp32 = bias[None] + mm(Xf, W)
mag = np.abs(bias).max() + (np.abs(Xf).max(0)[None] * np.abs(W)).sum(1).max()
srt = np.sort(exact, axis=1)
margin = srt[:, 1] - srt[:, 0]
print(f"n={n} K={K} D={D}  magnitude {mag:.3f}")
for name, approx in (("fp32 FMA chain", fp32), ("TF32 x1", one), ("TF32 x3 (hi.hi + hi.lo + lo.hi)", three)):
    err = np.abs(approx - exact).max()
    band = 2.0 * err
    wrong = (approx.argmin(1) != exact.argmin(1)).mean()
    print(f"{name:34s} max |error| {err:.3e} = {err / mag:.2e} of the magnitude = 2^{np.log2(err / mag):.1f};  "
          f"pixels within 2x that error of a tie {100 * (margin < band).mean():.3f} %;  argmin differs for {100 * wrong:.4f} %")
bound = (D + 3) * mag * 2.0 ** -20
print(f"planned bound (D+3) 2^-20 magnitude = {bound:.3e}: pixels inside a 2x band {100 * (margin < 2 * bound).mean():.3f} %")
