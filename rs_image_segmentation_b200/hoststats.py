"""Host-side order statistics from per-band histograms.

For an integer-valued raster every statistic the reference obtains by sorting
(np.percentile in robust_normalize, modules/features/indices.py:38-39; np.nanmedian /
np.nanpercentile inside RobustScaler, sklearn/preprocessing/_data.py:1722,1738-1743) is a
function of the 256- or 65536-bin histogram.  The GPU builds the histograms (K1); this module
turns them into the handful of scalars the next kernels need.  It is O(levels) work on a few
kilobytes, not a fallback for the pixel path.

The interpolation arithmetic follows numpy's `_quantile` for method="linear" with the dtypes numpy
ends up using for a float32 array: np.percentile(a, 2) divides q by float32(100), so the virtual
index (n-1)*q, gamma and the lerp are all float32; np.nanpercentile(a, (25.0, 75.0)) keeps q in
float64, so gamma is float64 and the result is float64 (which is why RobustScaler.scale_ is
float64 for float32 data).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _lerp(a, b, t):
    """numpy/lib/_function_base_impl.py `_lerp`, same expressions so the rounding is the same."""
    a, b, t = np.asanyarray(a), np.asanyarray(b), np.asanyarray(t)
    diff = np.subtract(b, a)
    out = np.asanyarray(np.add(a, diff * t))
    np.subtract(b, diff * (1 - t), out=out, where=t >= 0.5, casting="unsafe", dtype=type(out.dtype))
    return out[()]


class OrderBase:
    """numpy's quantile arithmetic on top of `at(k)` = k-th smallest sample; subclasses provide n, dtype, at."""

    n: int
    dtype: np.dtype

    def at(self, k: int):
        raise NotImplementedError

    def _quantile(self, q):
        # q is a numpy scalar / 0-d array whose dtype decides the arithmetic, as in numpy
        vi = np.asanyarray((self.n - 1) * q)
        prev = np.floor(vi)
        nxt = prev + 1
        if vi >= self.n - 1:
            prev = nxt = np.asanyarray(-1.0)
        if vi < 0:
            prev = nxt = np.asanyarray(0.0)
        pi, ni = int(prev), int(nxt)
        a = self.at(pi if pi >= 0 else self.n - 1)
        b = self.at(ni if ni >= 0 else self.n - 1)
        if a == b:  # both order statistics on one level (the usual case for a histogram): lerp(a, a, t) == a exactly
            return np.result_type(a, vi).type(a)
        gamma = np.asanyarray(vi - np.intp(pi), dtype=vi.dtype)
        return _lerp(a, b, gamma)

    def percentile_scalar(self, q):
        """np.percentile(band_float32, q) for a Python-number q: float32 arithmetic, float32 result."""
        q32 = np.true_divide(q, self.dtype.type(100) if self.dtype.kind == "f" else 100)
        return self._quantile(q32)

    def nanpercentile_pair(self, qs=(25.0, 75.0)):
        """np.nanpercentile(column_float32, (25.0, 75.0)): float64 quantiles, float64 result."""
        q = np.true_divide(np.asanyarray(qs), self.dtype.type(100) if self.dtype.kind == "f" else 100)
        return np.array([self._quantile(qi) for qi in q])

    def median(self):
        """np.nanmedian of a float32 column without NaNs: middle element, or the float32 mean of the two."""
        n = self.n
        if n % 2 == 1:
            return self.dtype.type(self.at(n // 2))
        pair = np.array([self.at(n // 2 - 1), self.at(n // 2)], dtype=self.dtype)
        return np.mean(pair)


class LevelOrder(OrderBase):
    """Sorted view of one band: cumulative counts over grey levels + the value each level maps to."""

    def __init__(self, hist, values, cum=None):
        hist = np.asarray(hist, dtype=np.int64)
        self.cum = np.cumsum(hist) if cum is None else cum
        self.n = int(self.cum[-1]) if hist.size else 0
        self.values = np.asarray(values)
        self.dtype = self.values.dtype
        if self.n <= 0:
            raise ValueError("empty histogram")
        if self.values.shape != hist.shape:
            raise ValueError("values/hist shape mismatch")

    def at(self, k: int):
        """k-th smallest sample (0-based) of the band."""
        k = int(k)
        k = 0 if k < 0 else (self.n - 1 if k >= self.n else k)
        return self.values[self.cum.searchsorted(k, "right")]


def norm_params(lo, hi):
    """(lo, hi, den) of robust_normalize, indices.py:42-46: den = hi - lo + 1e-10 in the band dtype."""
    lo, hi = F32(lo), F32(hi)
    den = hi - lo + 1e-10
    return lo, hi, F32(den)


def normalize_levels(values, lo, hi, den):
    """robust_normalize applied to the value of every level: (clip(x, lo, hi) - lo) / den, float32."""
    values = np.asarray(values, dtype=F32)
    return (np.clip(values, lo, hi) - lo) / den


class RasterStats:
    """Everything the kernels need that comes from the histograms of a (.., B) integer raster.

    norm      float32 [B][3]   lo, hi, den of robust_normalize per band
    norm_lut  float32 [B][L]   normalised value of every grey level
    qnorm     float32 [3]      second robust_normalize inside calculate_glcm_features (indices.py:265) on `glcm_band`
    center    float32 [B]      RobustScaler.center_ of the normalised bands
    scale     float64 [B]      RobustScaler.scale_
    x_lut     float32 [B][L]   RobustScaler().fit_transform value of every grey level (uint8 rasters)
    """

    def __init__(self, hist, glcm_band=3, lower=2, upper=98, native=True):
        hist = np.ascontiguousarray(hist, dtype=np.int64)
        B, L = hist.shape
        self.B, self.L = B, L
        self.hist = hist
        self.hist_raw = hist                       # pipeline.extract_features overwrites it with the DN histogram when stage 1 is fused
        self.n = int(hist[0].sum())
        self.norm = np.zeros((B, 3), F32)
        self.norm_lut = np.zeros((B, L), F32)
        self.qnorm = np.array([0, 1, 1], F32)
        self._scaler = None
        if native and L in (256, 65536):
            # the same arithmetic in C (librsx.so, rsx_raster_stats): microseconds instead of a millisecond between two kernels
            from . import _lib
            center, scale, x_lut = np.zeros(B, F32), np.ones(B, np.float64), np.zeros((B, L), F32)
            p = lambda a: a.ctypes.data_as(_lib.C.c_void_p)
            rc = _lib.load().rsx_raster_stats(p(hist), B, L, int(glcm_band) if 0 <= glcm_band < B else -1, float(lower), float(upper),
                                              p(self.norm), p(self.qnorm), p(center), p(scale), p(self.norm_lut), p(x_lut))
            if rc != 0:
                raise ValueError(_lib.load().rsx_last_error().decode())
            self._scaler = (center, scale, x_lut)
            return
        levels = np.arange(L, dtype=F32)
        # phase 1 (what K2 needs): robust_normalize parameters of every band, and of the normalised texture band
        for b in range(B):
            raw = LevelOrder(hist[b], levels)
            lo, hi, den = norm_params(raw.percentile_scalar(lower), raw.percentile_scalar(upper))
            self.norm[b] = (lo, hi, den)
            self.norm_lut[b] = normalize_levels(levels, lo, hi, den)
            if b == glcm_band:
                nb = LevelOrder(hist[b], self.norm_lut[b], cum=raw.cum)
                lo2, hi2, den2 = norm_params(nb.percentile_scalar(lower), nb.percentile_scalar(upper))
                self.qnorm = np.array([lo2, hi2, den2], F32)

    def _ensure_scaler(self):
        """phase 2 (what K3 needs): RobustScaler statistics of the normalised bands; computed on first use so that the
        caller can launch K2 first and do this while the device is busy."""
        if self._scaler is None:
            B, L = self.B, self.L
            center = np.zeros(B, F32)
            scale = np.ones(B, np.float64)
            x_lut = np.zeros((B, L), F32)
            for b in range(B):
                f = self.norm_lut[b]
                nb = LevelOrder(self.hist[b], f)
                center[b] = nb.median()
                q = nb.nanpercentile_pair((25.0, 75.0))
                s = np.float64(q[1] - q[0])
                if s < 10 * np.finfo(np.float64).eps:  # sklearn _handle_zeros_in_scale
                    s = 1.0
                scale[b] = s
                # X -= center_ (float32); X /= scale_ (float64 divisor -> evaluated in float64, stored float32)
                x_lut[b] = ((f - center[b]).astype(np.float64) / s).astype(F32)
            self._scaler = (center, scale, x_lut)
        return self._scaler

    @property
    def center(self):
        return self._ensure_scaler()[0]

    @property
    def scale(self):
        return self._ensure_scaler()[1]

    @property
    def x_lut(self):
        return self._ensure_scaler()[2]

    def quant_lut(self, band, levels):
        """(robust_normalize(norm) * (levels-1)).astype(uint8) per grey level (indices.py:265-268)."""
        g = normalize_levels(self.norm_lut[band], *self.qnorm)
        return (g * (levels - 1)).astype(np.uint8)


TM_GAIN = (0.671339, 1.322205, 1.043976, 0.876024, 0.120354, 0.055376, 0.065551)   # preprocessing.py:65
TM_BIAS = (-2.19, -4.16, -2.21, -2.39, -0.49, 1.18, -0.22)                           # preprocessing.py:66


def stage1_level_tables(hist_raw, gain=TM_GAIN, bias=TM_BIAS):
    """Stage 1 of the reference for 8-bit DN input as one table per band (modules/features/preprocessing.py:54-125):
    radiance = gain*DN + bias (float64), identity warp, (radiance - min) * 255.0 / (max - min) -> astype(uint8).
    hist_raw: int64 [B][256] DN histograms.  Returns (remap uint8 [B][256], hist_stage1 int64 [B][256])."""
    hist_raw = np.asarray(hist_raw, dtype=np.int64)
    B, L = hist_raw.shape
    assert L == 256, "the fused stage-1 chain is defined for 8-bit rasters"
    if len(gain) < B or len(bias) < B:
        raise ValueError(f"{B} bands but {len(gain)} calibration coefficients")
    remap = np.zeros((B, 256), np.uint8)
    hist1 = np.zeros((B, 256), np.int64)
    dn = np.arange(256, dtype=np.uint8)
    for b in range(B):
        present = np.flatnonzero(hist_raw[b])
        if present.size == 0:
            raise ValueError("empty histogram")
        radiance = gain[b] * dn + bias[b]                      # the reference's expression, elementwise on the 256 levels
        lo, hi = radiance[present[0]], radiance[present[-1]]   # np.min / np.max of the band's radiance (gain > 0 or < 0 alike)
        lo, hi = min(lo, hi), max(lo, hi)
        with np.errstate(divide="ignore", invalid="ignore"):
            e = (radiance - lo) * 255.0 / (hi - lo)
        ok = np.zeros(256, bool)
        ok[present] = True
        remap[b, ok] = e[ok].astype(np.uint8)
        np.add.at(hist1[b], remap[b, ok], hist_raw[b, ok])
    return remap, hist1


def float_band_order(band):
    """LevelOrder for an arbitrary float32 band (host np.unique; used only by the per-function drop-ins
    when they are handed small float arrays that are not integer valued)."""
    vals, counts = np.unique(np.asarray(band, dtype=F32).ravel(), return_counts=True)
    return LevelOrder(counts, vals)


def pca_from_moments(moments, n, n_components=None):
    """sklearn PCA(svd_solver='covariance_eigh') from the Gram matrix and column sums
    (sklearn/decomposition/_pca.py:587-640), evaluated in float64.

    moments: float64 [B + B(B+1)/2] (column sums, then upper-triangular X^T X row-major).
    Returns dict(mean, components, explained_variance, explained_variance_ratio, singular_values).
    """
    moments = np.asarray(moments, dtype=np.float64)
    B = int((np.sqrt(9 + 8 * moments.size) - 3) / 2 + 0.5)
    assert B + B * (B + 1) // 2 == moments.size
    mean = moments[:B] / n
    G = np.zeros((B, B))
    G[np.triu_indices(B)] = moments[B:]
    G = G + np.triu(G, 1).T
    Cov = (G - n * np.outer(mean, mean)) / (n - 1)
    w, v = np.linalg.eigh(Cov)
    w, v = w[::-1].copy(), v[:, ::-1]
    w[w < 0.0] = 0.0
    Vt = v.T.copy()
    # svd_flip(u_based_decision=False): largest-|.| entry of every component row is positive
    idx = np.argmax(np.abs(Vt), axis=1)
    signs = np.sign(Vt[np.arange(B), idx])
    signs[signs == 0] = 1.0
    Vt *= signs[:, None]
    k = B if n_components is None else int(n_components)
    total = w.sum()
    return dict(mean=mean, components=Vt[:k], explained_variance=w[:k],
                explained_variance_ratio=(w / total)[:k], singular_values=np.sqrt(w * (n - 1))[:k],
                noise_variance=float(w[k:].mean()) if k < B else 0.0)
