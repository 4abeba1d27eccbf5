"""Oracle (test infrastructure): restatement of the GLCM texture path.

PARITY UNPINNED by the reference: the arithmetic is scikit-image's
``skimage.feature.graycomatrix`` / ``graycoprops`` (skimage/feature/texture.py and
the ``_glcm_loop`` of skimage/feature/_texture.pyx), a dependency that is absent
from /root/reference, unpinned in its requirements.txt and not installed in this
image.  What is restated here is its published algorithm:

  graycomatrix(image, distances, angles, levels, symmetric, normed)
    for every angle a, distance d:   dr = round(sin(a)*d), dc = round(cos(a)*d)
    for every pixel (r,c) with (r+dr, c+dc) in bounds:
        P[image[r,c], image[r+dr,c+dc], d, a] += 1            (uint32)
    symmetric: P += P^T over the two level axes
    normed:    P = P.astype(float64) / P.sum(axes 0,1)         (sum 0 -> 1)
  graycoprops(P, prop): re-normalise, then
    contrast       sum P (i-j)^2        dissimilarity  sum P |i-j|
    homogeneity    sum P / (1+(i-j)^2)  energy         sqrt(sum P^2)
    correlation    sum P (i-mu_i)(j-mu_j) / (sd_i sd_j),  1 where sd < 1e-15

and it is anchored on the reference's own call site,
modules/features/indices.py:264-316 (quantise, top-left anchored windows,
``.mean()`` over the 1x4 property array, float32 map, cv2.resize back to HxW).
The known-answer vector is the 4x4 example of the graycomatrix docstring
(tests/test_oracle_glcm.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .features import robust_normalize

PROPS = ("contrast", "dissimilarity", "homogeneity", "energy", "correlation")
# Property known answers published by scikit-image's own test suite (skimage/feature/tests/test_texture.py: test_contrast,
# test_dissimilarity, test_homogeneity, test_energy, test_correlation) for the docstring image [[0,0,1,1],[0,0,1,1],[0,2,2,2],
# [2,2,3,3]], levels 4, symmetric=True, normed=True, entry (distance 1, angle 0); correlation also at distance 2.  The contrast
# and dissimilarity tests of scikit-image round the normalised matrix to 3 decimals first.
SKIMAGE_PROP_KATS = dict(contrast_rounded=0.585, dissimilarity_rounded=0.418, homogeneity=0.80833333, energy=0.38188131,
                         correlation=0.71953255, correlation_d2=0.41176470)
DEFAULT_ANGLES = (0.0, np.pi / 4, np.pi / 2, 3 * np.pi / 4)


def offsets_for(distances=(1,), angles=DEFAULT_ANGLES):
    """(d_idx, a_idx, dr, dc) exactly as graycomatrix rounds them."""
    out = []
    for di, d in enumerate(distances):
        for ai, a in enumerate(angles):
            out.append((di, ai, int(round(np.sin(a) * d)), int(round(np.cos(a) * d))))
    return out


def graycomatrix(img, distances=(1,), angles=DEFAULT_ANGLES, levels=32, symmetric=False, normed=False):
    """Vectorised restatement of skimage graycomatrix (counts as uint32 unless normed)."""
    img = np.asarray(img)
    rows, cols = img.shape
    P = np.zeros((levels, levels, len(distances), len(angles)), dtype=np.uint32)
    for di, ai, dr, dc in offsets_for(distances, angles):
        r0, r1 = max(0, -dr), min(rows, rows - dr)
        c0, c1 = max(0, -dc), min(cols, cols - dc)
        if r1 <= r0 or c1 <= c0:
            continue
        a = img[r0:r1, c0:c1].astype(np.intp).ravel()
        b = img[r0 + dr:r1 + dr, c0 + dc:c1 + dc].astype(np.intp).ravel()
        ok = (a >= 0) & (a < levels) & (b >= 0) & (b < levels)
        P[:, :, di, ai] = np.bincount(a[ok] * levels + b[ok], minlength=levels * levels).reshape(levels, levels)
    if symmetric:
        P = P + P.transpose(1, 0, 2, 3)
    if normed:
        P = P.astype(np.float64)
        s = P.sum(axis=(0, 1), keepdims=True)
        s[s == 0] = 1
        P = P / s
    return P


def graycoprops(P, prop):
    """Restatement of skimage graycoprops for the five properties the reference uses."""
    L = P.shape[0]
    P = P.astype(np.float64)
    s = P.sum(axis=(0, 1), keepdims=True)
    s[s == 0] = 1
    P = P / s
    I, J = np.ogrid[0:L, 0:L]
    if prop == "contrast":
        return (P * ((I - J) ** 2)[:, :, None, None]).sum(axis=(0, 1))
    if prop == "dissimilarity":
        return (P * np.abs(I - J)[:, :, None, None]).sum(axis=(0, 1))
    if prop == "homogeneity":
        return (P * (1.0 / (1.0 + (I - J) ** 2))[:, :, None, None]).sum(axis=(0, 1))
    if prop == "energy":
        return np.sqrt((P ** 2).sum(axis=(0, 1)))
    if prop == "correlation":
        I4 = I.reshape(L, 1, 1, 1).astype(np.float64)
        J4 = J.reshape(1, L, 1, 1).astype(np.float64)
        di = I4 - (I4 * P).sum(axis=(0, 1))
        dj = J4 - (J4 * P).sum(axis=(0, 1))
        sd_i = np.sqrt((P * di ** 2).sum(axis=(0, 1)))
        sd_j = np.sqrt((P * dj ** 2).sum(axis=(0, 1)))
        cov = (P * (di * dj)).sum(axis=(0, 1))
        res = np.empty(P.shape[2:], dtype=np.float64)
        flat = (sd_i < 1e-15) | (sd_j < 1e-15)
        res[flat] = 1.0
        res[~flat] = cov[~flat] / (sd_i[~flat] * sd_j[~flat])
        return res
    raise ValueError(prop)


def quantize(band, levels=32):
    """indices.py:265-268 - second robust_normalize, scale, truncate to uint8."""
    return (robust_normalize(band) * (levels - 1)).astype(np.uint8)


def window_counts(q, i, j, window, levels, angles=DEFAULT_ANGLES):
    """Directed (non-symmetrised) uint32 counts of the window anchored at (i, j)."""
    return graycomatrix(q[i:i + window, j:j + window], (1,), angles, levels)[:, :, 0, :]


def props_map_numpy(q, levels, window, step, angles=DEFAULT_ANGLES, distances=(1,)):
    """indices.py:270-305 - the Python double loop, one graycomatrix per window."""
    H, W = q.shape
    oh, ow = (H - window) // step + 1, (W - window) // step + 1
    out = np.zeros((5, oh, ow), dtype=np.float32)
    for i in range(0, H - window + 1, step):
        for j in range(0, W - window + 1, step):
            P = graycomatrix(q[i:i + window, j:j + window], distances, angles, levels, symmetric=True, normed=True)
            for k, name in enumerate(PROPS):
                out[k, i // step, j // step] = graycoprops(P, name).mean()
    return out


def resize_to(img, H, W):
    """indices.py:308 - cv2.resize(img, (W, H), INTER_LINEAR)."""
    import cv2

    return cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)


def glcm_features(band, levels=32, window_size=21, step_size=21, use_c=True):
    """indices.py:248-318 end to end (distance 1, the four default angles)."""
    q = quantize(band, levels)
    H, W = q.shape
    maps = props_map_c(q, levels, window_size, step_size) if use_c else props_map_numpy(q, levels, window_size, step_size)
    return {name: resize_to(maps[k], H, W) for k, name in enumerate(PROPS)}


# ---------------------------------------------------------------- plain-C port
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_glcm.so")
_lib = None


def build_c(force=False):
    src = os.path.join(_HERE, "glcm_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _c():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        _lib.oracle_glcm_props.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _lib.oracle_glcm_props.restype = ctypes.c_int
        _lib.oracle_glcm_counts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_glcm_counts.restype = ctypes.c_int
    return _lib


def props_map_c(q, levels, window, step, threads=0):
    """Same result as props_map_numpy, histogram route in C (OpenMP over window rows)."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    H, W = q.shape
    oh, ow = (H - window) // step + 1, (W - window) // step + 1
    out = np.zeros((5, oh, ow), dtype=np.float32)
    rc = _c().oracle_glcm_props(q.ctypes.data, H, W, levels, window, step, out.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f"oracle_glcm_props failed: {rc}")
    return out


def counts_map_c(q, levels, window, step):
    """Directed uint32 counts for every window: (oh, ow, 4, L, L)."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    H, W = q.shape
    oh, ow = (H - window) // step + 1, (W - window) // step + 1
    out = np.zeros((oh, ow, 4, levels, levels), dtype=np.uint32)
    rc = _c().oracle_glcm_counts(q.ctypes.data, H, W, levels, window, step, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_glcm_counts failed: {rc}")
    return out


# ------------------------------------------------------------------ integer pair moments (what the dense GPU kernel keeps)
MOMENT_FIELDS = ("n", "s1", "sa", "sq", "sab", "e", "neq", "hom_fx")


def hom_bits(levels):
    """Fixed-point scale of the homogeneity terms in the GPU kernels: 2^40 up to 32 levels, 2^36 above (rsx.h rsx_glcm_moments)."""
    return 40 if levels <= 32 else 36


def pair_moments(q, levels, window, step=1):
    """(oh, ow, 4, 8) int64: for every window and angle the exact integers from which the five properties follow without a
    histogram (include/rsx.h rsx_glcm_moments; derived from the directed counts of counts_map_c, i.e. from the same
    graycomatrix restatement that the docstring known answer pins):
      n    pair instances            s1  sum |a-b|          sa  sum (a+b)        sq  sum (a^2+b^2)      sab  sum a*b
      e    sum over the cells of the SYMMETRIC count matrix P = C + C^T of P^2 (so energy = sqrt(e) / 2n)
      neq  pairs with a == b         hom_fx  sum of round(2^h / (1 + (a-b)^2)), h = hom_bits(levels) (the kernels' fixed-point terms)"""
    C = counts_map_c(q, levels, window, step).astype(np.int64)            # (oh, ow, 4, L, L)
    a = np.arange(levels, dtype=np.int64).reshape(levels, 1)
    b = np.arange(levels, dtype=np.int64).reshape(1, levels)
    d = np.abs(a - b)
    hom = np.floor(2.0 ** hom_bits(levels) / (1.0 + (d * d).astype(np.float64)) + 0.5).astype(np.int64)
    S = C + C.transpose(0, 1, 2, 4, 3)
    out = np.zeros(C.shape[:3] + (8,), dtype=np.int64)
    out[..., 0] = C.sum(axis=(3, 4))
    out[..., 1] = (C * d).sum(axis=(3, 4))
    out[..., 2] = (C * (a + b)).sum(axis=(3, 4))
    out[..., 3] = (C * (a * a + b * b)).sum(axis=(3, 4))
    out[..., 4] = (C * (a * b)).sum(axis=(3, 4))
    out[..., 5] = (S * S).sum(axis=(3, 4))
    out[..., 6] = (C * (d == 0)).sum(axis=(3, 4))
    out[..., 7] = (C * hom).sum(axis=(3, 4))
    return out


def props_from_moments(m, levels=32):
    """The five graycoprops of one angle from its pair moments, float64 (rsx_glcm.cu header comment)."""
    n, s1, sa, sq, sab, e, neq, hom_fx = (int(v) for v in m)
    var_num = 2 * n * sq - sa * sa
    cov_num = 4 * n * sab - sa * sa
    return dict(contrast=(sq - 2 * sab) / n, dissimilarity=s1 / n, homogeneity=hom_fx / 2.0 ** hom_bits(levels) / n, energy=np.sqrt(float(e)) / (2.0 * n),
                correlation=1.0 if var_num <= 0 else cov_num / var_num)


# ------------------------------------------------------------------ closed-form known answers (implementation independent)
def analytic_stripes(a, b, win):
    """Closed-form graycoprops means for vertical stripes of levels a, b (columns alternate a, b, a, ...; window starts on an
    `a` column): what the textbook definition gives, independent of any implementation.
    0 deg / 45 deg / 135 deg: every pair joins an a and a b column -> symmetric P[a,b] = P[b,a] = 1/2.
    90 deg: pairs stay inside a column -> P[a,a] = na/(na+nb), P[b,b] = nb/(na+nb) (na, nb = columns of each level)."""
    d = abs(a - b)
    cross = dict(contrast=d * d, dissimilarity=d, homogeneity=1.0 / (1.0 + d * d), energy=np.sqrt(0.5), correlation=-1.0)
    na, nb = (win + 1) // 2, win // 2
    pa, pb = na / (na + nb), nb / (na + nb)
    vert = dict(contrast=0.0, dissimilarity=0.0, homogeneity=1.0, energy=np.sqrt(pa * pa + pb * pb), correlation=1.0)
    return {k: (3 * cross[k] + vert[k]) / 4 for k in cross}


def stripes_image(a, b, H, W):
    q = np.empty((H, W), np.uint8)
    q[:, 0::2], q[:, 1::2] = a, b
    return q
