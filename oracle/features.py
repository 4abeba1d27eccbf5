"""Oracle (test infrastructure): numpy restatement of the per-pixel feature stack.

Each function cites the reference lines it restates (paths relative to
/root/reference).  Arithmetic is kept in the reference's dtypes (float32 maps,
np.percentile on float32, sklearn in float32) so that results can be compared
bit-for-bit where the north star demands it.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- a1
def robust_normalize(band, lower_percentile=2, upper_percentile=98):
    """modules/features/indices.py:25-48 - percentile clip then affine map to [0,1]."""
    lo = np.percentile(band, lower_percentile)
    hi = np.percentile(band, upper_percentile)
    return (np.clip(band, lo, hi) - lo) / (hi - lo + 1e-10)


# ---------------------------------------------------------------------- a2..a5
def _masked_ratio(num, den):
    """Shared tail of indices.py:62-69 / 128-135 / 150-156 / 171-177 / 194-201."""
    out = np.zeros_like(den, dtype=F32)
    ok = den > 0.001
    out[ok] = num[ok] / den[ok]
    return np.clip(out, -1.0, 1.0)


def ndvi(nir, red):
    """indices.py:50-71"""
    return _masked_ratio(nir - red, nir + red)


def ndwi(green, nir):
    """indices.py:116-137"""
    return _masked_ratio(green - nir, green + nir)


def mndwi(green, swir):
    """indices.py:139-158"""
    return _masked_ratio(green - swir, green + swir)


def ndbi(swir, nir):
    """indices.py:160-179"""
    return _masked_ratio(swir - nir, swir + nir)


def evi(nir, red, blue, L=1, C1=6, C2=7.5, G=2.5):
    """indices.py:73-95 - note the evaluation order G*(nir-red) then the divide."""
    den = nir + C1 * red - C2 * blue + L
    out = np.zeros_like(nir, dtype=F32)
    ok = den > 0.001
    out[ok] = G * (nir[ok] - red[ok]) / den[ok]
    return np.clip(out, -1.0, 1.0)


def msavi(nir, red):
    """indices.py:97-114"""
    t = 2 * nir + 1
    return np.clip((t - np.sqrt(t ** 2 - 8 * (nir - red))) / 2, -1.0, 1.0)


def bsi(blue, red, nir, swir):
    """indices.py:181-203"""
    a, b = swir + red, nir + blue
    return _masked_ratio(a - b, a + b)


INDEX_ORDER = ("ndvi", "evi", "msavi", "ndwi", "mndwi", "ndbi", "bsi")


def all_indices(nb):
    """The seven calls of scripts/2_feature_extraction.py:63-73 on normalised bands
    nb = [blue, green, red, nir, swir1, ...]."""
    blue, green, red, nir, swir1 = nb[0], nb[1], nb[2], nb[3], nb[4]
    return {
        "ndvi": ndvi(nir, red),
        "evi": evi(nir, red, blue),
        "msavi": msavi(nir, red),
        "ndwi": ndwi(green, nir),
        "mndwi": mndwi(green, swir1),
        "ndbi": ndbi(swir1, nir),
        "bsi": bsi(blue, red, nir, swir1),
    }


# --------------------------------------------------------------------------- a6
def perform_pca(bands, n_components=None, use_robust_scaling=True, promote=False):
    """indices.py:205-246 - (N,B) float32 matrix -> RobustScaler -> sklearn PCA.

    promote=True runs the PCA itself on the float64 promotion of the (float32) scaled matrix: the same
    mathematical object without the float32 Gram-matrix noise of sklearn's covariance_eigh solver, which
    for the minor components (tiny eigenvalue gaps) exceeds the 1e-5 parity bar on its own."""
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import RobustScaler

    h, w = bands[0].shape
    X = np.stack([np.asarray(b, dtype=F32).ravel() for b in bands], axis=1)
    if use_robust_scaling:
        X = RobustScaler().fit_transform(X)
    else:
        X = (X - X.min(axis=0)) / (X.max(axis=0) - X.min(axis=0) + 1e-10)
    if promote:
        X = X.astype(np.float64)
    model = PCA(n_components=n_components)
    Y = model.fit_transform(X)
    return [Y[:, i].reshape(h, w) for i in range(Y.shape[1])], model.explained_variance_ratio_, model


# --------------------------------------------------------------------------- a8
def level1_stack(ix, pcs):
    """indices.py:808-835 - [ndwi, mndwi, ndvi, evi, ndbi, bsi, pc0]."""
    maps = [ix["ndwi"], ix["mndwi"], ix["ndvi"], ix["evi"], ix["ndbi"], ix["bsi"]]
    if pcs:
        maps.append(pcs[0])
    return np.stack(maps, axis=-1)


def spatial_context(stack, window_size=7):
    """indices.py:760-776 - 7x7 box mean (BORDER_REFLECT) appended, float64 result."""
    import cv2

    ctx = np.zeros(stack.shape)
    for i in range(stack.shape[2]):
        ctx[:, :, i] = cv2.boxFilter(stack[:, :, i], -1, (window_size, window_size),
                                     normalize=True, borderType=cv2.BORDER_REFLECT)
    return np.concatenate([stack, ctx], axis=-1)


# ------------------------------------------------------------------- stage 1
TM_GAIN = (0.671339, 1.322205, 1.043976, 0.876024, 0.120354, 0.055376, 0.065551)
TM_BIAS = (-2.19, -4.16, -2.21, -2.39, -0.49, 1.18, -0.22)


def stage1_preprocess(raw_bands):
    """modules/features/preprocessing.py:54-125 - gain/bias, (identity warp), min-max
    stretch to uint8; scripts/2_feature_extraction.py:157-161 then re-reads as float32."""
    out = []
    for i, b in enumerate(raw_bands):
        r = TM_GAIN[i] * b + TM_BIAS[i]
        e = ((r - r.min()) * 255.0 / (r.max() - r.min())).astype(np.uint8)
        out.append(e)
    return out


# ---------------------------------------------------------------- level 2 (N2)
def morph_gradient(band, size=5):
    """indices.py:412-440 - calculate_morphological_features(band)['gradient_<size>']."""
    import cv2

    b8 = (robust_normalize(band) * 255).astype(np.uint8)
    return cv2.morphologyEx(b8, cv2.MORPH_GRADIENT, np.ones((size, size), np.uint8)) / 255.0


def std_dev_scale(band, scale=5):
    """indices.py:531-548 - calculate_multi_scale_features(band)['std_dev_scale_<scale>']."""
    import cv2

    b = robust_normalize(band)
    mean = cv2.blur(b, (scale, scale))
    mean_sq = cv2.blur(b * b, (scale, scale))
    var = mean_sq - mean * mean
    var[var < 0] = 0
    return np.sqrt(var)


def sobel_mag(band):
    """indices.py:455-480 - calculate_filter_responses(band)['sobel_mag']."""
    import cv2

    b8 = (robust_normalize(band) * 255).astype(np.uint8)
    sx = cv2.Sobel(b8, cv2.CV_32F, 1, 0) / 255.0
    sy = cv2.Sobel(b8, cv2.CV_32F, 0, 1) / 255.0
    m = np.sqrt(sx ** 2 + sy ** 2)
    return m / (m.max() + 1e-10)


def level2_stack(glcm, band):
    """indices.py:837-865 with the three stencil maps of the texture band."""
    return np.stack([glcm["contrast"], glcm["homogeneity"], morph_gradient(band), std_dev_scale(band), sobel_mag(band)], axis=-1)
