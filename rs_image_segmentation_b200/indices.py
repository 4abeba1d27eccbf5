"""Drop-in for the hot-path functions of the reference's modules/features/indices.py.

Same names, positional order, keyword names, defaults and return types as the reference, so that
`from rs_image_segmentation_b200.indices import *` can replace `from modules.features.indices import *`
(scripts/2_feature_extraction.py:20) for the functions on the hot path.  Inputs and outputs are host numpy
arrays; the arithmetic runs in the CUDA kernels of librsx.so (no CPU fallback: without a GPU every function
raises).  The fused, device-resident route for whole scenes is `run_feature_extraction_stage` /
`pipeline.extract_features`; the per-function entry points below pay one host<->device round trip each.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, hoststats
from .device import hptr, ptr, require_cuda, stream_ptr

__all__ = ["robust_normalize", "calculate_ndvi", "calculate_evi", "calculate_msavi", "calculate_ndwi", "calculate_mndwi",
           "calculate_ndbi", "calculate_bsi", "perform_pca", "calculate_glcm_features", "prepare_level_1_features",
           "add_spatial_context", "morphological_gradient", "local_std_dev", "sobel_magnitude", "prepare_level_2_features",
           "run_feature_extraction_stage", "RsxPCA"]

_DEFAULT_ANGLES = [0, np.pi / 4, np.pi / 2, 3 * np.pi / 4]


# ------------------------------------------------------------------------------------------- helpers
def _dev32(a) -> torch.Tensor:
    """float32 device copy of a 2-D array (the reference feeds float32 maps: scripts/2_feature_extraction.py:158)."""
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return torch.from_numpy(a).cuda(non_blocking=False)


def _out_like(t: torch.Tensor) -> torch.Tensor:
    return torch.empty(t.numel(), dtype=torch.float32, device=t.device)


class _DeviceOrder(hoststats.OrderBase):
    """Order statistics of an arbitrary float32 band: one device sort, numpy's interpolation on the host."""

    def __init__(self, flat: torch.Tensor):
        self.sorted = torch.sort(flat).values
        self.n = int(flat.numel())
        self.dtype = np.dtype(np.float32)

    def at(self, k: int):
        k = min(max(int(k), 0), self.n - 1)
        return np.float32(self.sorted[k].item())


def _norm_params_device(flat: torch.Tensor, lower, upper):
    order = _DeviceOrder(flat)
    return hoststats.norm_params(order.percentile_scalar(lower), order.percentile_scalar(upper))


# ------------------------------------------------------------------------------------------- a1
def robust_normalize(band, lower_percentile=2, upper_percentile=98):
    """modules/features/indices.py:25-48."""
    require_cuda()
    band = np.asarray(band)
    x = _dev32(band)
    flat = x.reshape(-1)
    if bool(torch.isnan(flat).any()):
        return np.full(band.shape, np.nan, dtype=np.float32)      # np.percentile of a NaN band is NaN
    lo, hi, den = _norm_params_device(flat, lower_percentile, upper_percentile)
    out = _out_like(flat)
    _lib.call("rsx_normalize_f32", ptr(flat), flat.numel(), float(lo), float(hi), float(den), ptr(out), stream_ptr())
    return out.cpu().numpy().reshape(band.shape)


# ------------------------------------------------------------------------------------------- a2..a5
def _ratio(a, b):
    require_cuda()
    shape = np.asarray(a).shape
    A, B = _dev32(a).reshape(-1), _dev32(b).reshape(-1)
    out = _out_like(A)
    _lib.call("rsx_index_ratio_f32", ptr(A), ptr(B), A.numel(), ptr(out), stream_ptr())
    return out.cpu().numpy().reshape(shape)


def calculate_ndvi(nir_band, red_band):
    """indices.py:50-71."""
    return _ratio(nir_band, red_band)


def calculate_ndwi(green_band, nir_band):
    """indices.py:116-137."""
    return _ratio(green_band, nir_band)


def calculate_mndwi(green_band, swir_band):
    """indices.py:139-158."""
    return _ratio(green_band, swir_band)


def calculate_ndbi(swir_band, nir_band):
    """indices.py:160-179."""
    return _ratio(swir_band, nir_band)


def calculate_evi(nir_band, red_band, blue_band, L=1, C1=6, C2=7.5, G=2.5):
    """indices.py:73-95."""
    require_cuda()
    shape = np.asarray(nir_band).shape
    N, R, Bl = _dev32(nir_band).reshape(-1), _dev32(red_band).reshape(-1), _dev32(blue_band).reshape(-1)
    out = _out_like(N)
    _lib.call("rsx_index_evi_f32", ptr(N), ptr(R), ptr(Bl), N.numel(), float(L), float(C1), float(C2), float(G), ptr(out), stream_ptr())
    return out.cpu().numpy().reshape(shape)


def calculate_msavi(nir_band, red_band):
    """indices.py:97-114."""
    require_cuda()
    shape = np.asarray(nir_band).shape
    N, R = _dev32(nir_band).reshape(-1), _dev32(red_band).reshape(-1)
    out = _out_like(N)
    _lib.call("rsx_index_msavi_f32", ptr(N), ptr(R), N.numel(), ptr(out), stream_ptr())
    return out.cpu().numpy().reshape(shape)


def calculate_bsi(blue_band, red_band, nir_band, swir_band):
    """indices.py:181-203."""
    require_cuda()
    shape = np.asarray(blue_band).shape
    Bl, R, N, S = (_dev32(x).reshape(-1) for x in (blue_band, red_band, nir_band, swir_band))
    out = _out_like(Bl)
    _lib.call("rsx_index_bsi_f32", ptr(Bl), ptr(R), ptr(N), ptr(S), Bl.numel(), ptr(out), stream_ptr())
    return out.cpu().numpy().reshape(shape)


# ------------------------------------------------------------------------------------------- a6
class RsxPCA:
    """What perform_pca returns in place of the fitted sklearn PCA (indices.py:246): same attribute names."""

    def __init__(self, pca: dict, n_samples: int, n_features: int):
        self.components_ = pca["components"].astype(np.float32)
        self.mean_ = pca["mean"].astype(np.float32)
        self.explained_variance_ = pca["explained_variance"].astype(np.float32)
        self.explained_variance_ratio_ = pca["explained_variance_ratio"].astype(np.float32)
        self.singular_values_ = pca["singular_values"].astype(np.float32)
        self.noise_variance_ = pca["noise_variance"]
        self.n_components_ = self.components_.shape[0]
        self.n_samples_, self.n_features_in_ = n_samples, n_features

    def transform(self, X):
        """sklearn/decomposition/_base.py:151-159 on the device (cuBLAS through torch; not on the reference's path)."""
        require_cuda()
        Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()
        V = torch.from_numpy(self.components_).cuda()
        mu = torch.from_numpy(self.mean_).cuda()
        return (Xd @ V.t() - (mu.reshape(1, -1) @ V.t())).cpu().numpy()


_COMPILED_BANDS = (5, 6, 7, 8, 10, 12, 13)


def perform_pca(bands_data, n_components=None, use_robust_scaling=True):
    """indices.py:205-246.

    The bands the reference passes here are robust-normalised uint8 bands, i.e. float maps with at most 256
    distinct values each.  Each band is rank-coded on the device (torch.unique), which turns the call into the
    uint8 raster path: histogram -> RobustScaler statistics -> per-level table of X -> Gram/moments kernel ->
    eigh (host, BxB) -> projection kernel.  Bands with more than 256 distinct values (e.g. derived from 16-bit data) or a band
    count outside the compiled raster kernels go through the planar float32 kernels (_perform_pca_planar), any B <= 16.
    """
    require_cuda()
    H, W = np.asarray(bands_data[0]).shape
    B = len(bands_data)
    if B > 16:
        raise _lib.RsxError(f"perform_pca: {B} bands; the kernels take up to 16 (RSX_MAX_BANDS)")
    n = H * W
    st = stream_ptr()
    levels, value_tables, planes = [], [], []
    level_coded = B in _COMPILED_BANDS
    for b in bands_data:
        x = _dev32(b).reshape(-1)
        if bool(torch.isnan(x).any()):
            raise ValueError("Input X contains NaN.")                      # what sklearn raises
        planes.append(x)
        if level_coded:
            vals, inv = torch.unique(x, return_inverse=True)
            if vals.numel() > 256:
                level_coded = False                                        # arbitrary float band: the planar path below
            else:
                levels.append(inv.to(torch.uint8))
                value_tables.append(vals.cpu().numpy().astype(np.float32))
    if not level_coded:
        return _perform_pca_planar(planes, H, W, n_components, use_robust_scaling)
    raster = torch.stack(levels, dim=1).contiguous()                       # (N, B) uint8, pixel interleaved
    hist = torch.zeros((B, 256), dtype=torch.int32, device="cuda")
    _lib.call("rsx_hist_u8", ptr(raster), n, B, ptr(hist), st)
    hist = hist.cpu().numpy().astype(np.int64)
    lut = np.zeros((B, 256), np.float32)
    for b in range(B):
        v = value_tables[b]
        k = len(v)
        if use_robust_scaling:                                             # RobustScaler().fit_transform
            order = hoststats.LevelOrder(hist[b, :k], v)
            center = order.median()
            q = order.nanpercentile_pair((25.0, 75.0))
            s = np.float64(q[1] - q[0])
            if s < 10 * np.finfo(np.float64).eps:
                s = 1.0
            lut[b, :k] = ((v - center).astype(np.float64) / s).astype(np.float32)
        else:                                                              # indices.py:234
            lut[b, :k] = (v - v.min()) / (v.max() - v.min() + 1e-10)
    n_comp = B if n_components is None else int(n_components)
    dlut = torch.from_numpy(lut).cuda()
    M = B + B * (B + 1) // 2
    moments = torch.zeros(M, dtype=torch.float64, device="cuda")
    scratch = torch.empty(int(_lib.load().rsx_pca_scratch_elems(B)), dtype=torch.float64, device="cuda")
    _lib.call("rsx_pca_moments_u8", ptr(raster), n, B, ptr(dlut), ptr(moments), ptr(scratch), st)
    pca = hoststats.pca_from_moments(moments.cpu().numpy(), n, n_comp)
    comps = np.ascontiguousarray(pca["components"], dtype=np.float32)
    mean_proj = np.ascontiguousarray((pca["mean"].astype(np.float32).reshape(1, -1) @ comps.T).ravel(), dtype=np.float32)
    stride = (n + 31) // 32 * 32
    out = torch.empty((n_comp, stride), dtype=torch.float32, device="cuda")
    _lib.call("rsx_pca_project_u8", ptr(raster), n, B, ptr(dlut), hptr(comps), hptr(mean_proj), n_comp, ptr(out), stride, None, st)
    maps = out[:, :n].cpu().numpy()
    model = RsxPCA(pca, n, B)
    return [maps[i].reshape(H, W) for i in range(n_comp)], model.explained_variance_ratio_, model


def _perform_pca_planar(planes, H, W, n_components, use_robust_scaling):
    """perform_pca for bands with arbitrary float32 values / any band count up to 16: order statistics of every band from one
    device sort, then rsx_pca_moments_planar_f32 -> eigh (host) -> rsx_pca_project_planar_f32."""
    B, n = len(planes), H * W
    st = stream_ptr()
    stride = (n + 31) // 32 * 32
    stack = torch.zeros((B, stride), dtype=torch.float32, device="cuda")
    a, den, scale = np.zeros(B, np.float32), np.ones(B, np.float32), np.ones(B, np.float64)
    for b, x in enumerate(planes):
        stack[b, :n] = x
        if use_robust_scaling:                                             # RobustScaler().fit (sklearn/preprocessing/_data.py:1722-1743)
            order = _DeviceOrder(x)
            a[b] = order.median()
            q = order.nanpercentile_pair((25.0, 75.0))
            s = np.float64(q[1] - q[0])
            scale[b] = 1.0 if s < 10 * np.finfo(np.float64).eps else s
        else:                                                              # indices.py:234, float32 throughout (NEP 50)
            mn, mx = np.float32(x.min().item()), np.float32(x.max().item())
            a[b], den[b] = mn, np.float32(mx - mn + 1e-10)
    n_comp = B if n_components is None else int(n_components)
    M = B + B * (B + 1) // 2
    moments = torch.zeros(M, dtype=torch.float64, device="cuda")
    scratch = torch.empty(int(_lib.load().rsx_pca_planar_scratch_elems(B)), dtype=torch.float64, device="cuda")
    robust = 1 if use_robust_scaling else 0
    _lib.call("rsx_pca_moments_planar_f32", ptr(stack), stride, n, B, robust, hptr(a), hptr(den), hptr(scale), ptr(moments), ptr(scratch), st)
    pca = hoststats.pca_from_moments(moments.cpu().numpy(), n, n_comp)
    comps = np.ascontiguousarray(pca["components"], dtype=np.float32)
    mean_proj = np.ascontiguousarray((pca["mean"].astype(np.float32).reshape(1, -1) @ comps.T).ravel(), dtype=np.float32)
    out = torch.empty((n_comp, stride), dtype=torch.float32, device="cuda")
    _lib.call("rsx_pca_project_planar_f32", ptr(stack), stride, n, B, robust, hptr(a), hptr(den), hptr(scale), hptr(comps), hptr(mean_proj), n_comp,
              ptr(out), stride, None, st)
    maps = out[:, :n].cpu().numpy()
    model = RsxPCA(pca, n, B)
    return [maps[i].reshape(H, W) for i in range(n_comp)], model.explained_variance_ratio_, model


# ------------------------------------------------------------------------------------------- a7
def calculate_glcm_features(band, distances=[1], angles=[0, np.pi / 4, np.pi / 2, 3 * np.pi / 4], levels=32, window_size=21, step_size=21):
    """indices.py:248-318.  distances=[1] with the four default angles (all the reference ever asks for) runs the dense / tiled
    fast kernels; any other list of up to 16 (distance, angle) pairs runs the general kernel with graycomatrix's offsets
    round(sin(angle) * d), round(cos(angle) * d)."""
    require_cuda()
    offsets = [(int(round(np.sin(a) * d)), int(round(np.cos(a) * d))) for d in distances for a in angles]
    custom = offsets != [(0, 1), (1, 1), (1, 0), (1, -1)]
    if custom and not 1 <= len(offsets) <= 16:
        raise _lib.RsxError("calculate_glcm_features: between 1 and 16 (distance, angle) pairs")
    band = np.asarray(band)
    H, W = band.shape
    if window_size > H or window_size > W:
        raise ValueError("window larger than the band")
    st = stream_ptr()
    x = _dev32(band).reshape(-1)
    lo, hi, den = _norm_params_device(x, 2, 98)                             # indices.py:265
    q = torch.empty((H * W + 3) // 4 * 4, dtype=torch.uint8, device="cuda")
    _lib.call("rsx_quantize_f32", ptr(x), H * W, float(lo), float(hi), float(den), int(levels), ptr(q), st)   # :268
    oh, ow = (H - window_size) // step_size + 1, (W - window_size) // step_size + 1
    pstride = (oh * ow + 31) // 32 * 32
    props = torch.empty((5, pstride), dtype=torch.float32, device="cuda")
    if custom:
        h_off = np.ascontiguousarray(offsets, dtype=np.int32)
        _lib.call("rsx_glcm_props_offsets", ptr(q), H, W, int(levels), int(window_size), int(step_size), oh, ow, hptr(h_off), len(offsets),
                  ptr(props), pstride, st)
    else:
        _lib.call("rsx_glcm_props", ptr(q), H, W, int(levels), int(window_size), int(step_size), oh, ow, ptr(props), pstride, st)
    full = torch.empty((5, H * W), dtype=torch.float32, device="cuda")
    _lib.call("rsx_resize_bilinear_f32", ptr(props), oh, ow, 0, oh, pstride, ptr(full), H, W, 0, H, H * W, 5, None, st)   # :308
    maps = full.cpu().numpy().reshape(5, H, W)
    names = ("contrast", "dissimilarity", "homogeneity", "energy", "correlation")
    return {k: maps[i] for i, k in enumerate(names)}


# ------------------------------------------------------------------------------------------- a8 (glue)
def prepare_level_1_features(features_dict):
    """indices.py:808-835: [ndwi, mndwi, ndvi, evi, ndbi, bsi, pc0] stacked on the last axis (host glue)."""
    maps = [features_dict[k] for k in ("ndwi", "mndwi", "ndvi", "evi", "ndbi", "bsi")]
    if "pca_result" in features_dict and len(features_dict["pca_result"]) > 0:
        maps.append(features_dict["pca_result"][0])
    return np.stack(maps, axis=-1)


def add_spatial_context(features_array, window_size=7):
    """indices.py:760-776: the (H, W, n) stack followed by the window_size x window_size box mean (cv2.boxFilter,
    normalize=True, BORDER_REFLECT) of every channel; float64 (H, W, 2n) like the reference (np.zeros default dtype)."""
    require_cuda()
    features_array = np.asarray(features_array)
    H, W, n = features_array.shape
    planar = np.ascontiguousarray(np.moveaxis(features_array, -1, 0), dtype=np.float32)
    src = torch.from_numpy(planar).cuda().reshape(n, H * W)
    dst = torch.empty_like(src)
    _lib.call("rsx_box_mean_f32", ptr(src), H, W, 0, H, H * W, ptr(dst), 0, H, H * W, n, int(window_size), None, stream_ptr())
    ctx = np.moveaxis(dst.cpu().numpy().reshape(n, H, W), 0, -1)
    return np.concatenate([features_array, ctx.astype(np.float64)], axis=-1)


# ------------------------------------------------------------------------------------------- N2: level-2 stencil channels
def _second_normalisation(band):
    """robust_normalize(band) as every texture function of the reference does first; returns (float32 device plane, uint8
    device plane = (plane * 255).astype(uint8), shape)."""
    band = np.asarray(band)
    x = _dev32(band).reshape(-1)
    lo, hi, den = _norm_params_device(x, 2, 98)
    xf = _out_like(x)
    _lib.call("rsx_normalize_f32", ptr(x), x.numel(), float(lo), float(hi), float(den), ptr(xf), stream_ptr())
    xq = torch.empty((x.numel() + 3) // 4 * 4, dtype=torch.uint8, device="cuda")
    _lib.call("rsx_quantize_f32", ptr(x), x.numel(), float(lo), float(hi), float(den), 256, ptr(xq), stream_ptr())
    return xf, xq, band.shape


def morphological_gradient(band, size=5):
    """calculate_morphological_features(band)['gradient_<size>'] (indices.py:408-440): float64 map gradient / 255.0."""
    require_cuda()
    xf, xq, (H, W) = _second_normalisation(band)
    g = torch.empty(H * W, dtype=torch.uint8, device="cuda")
    _lib.call("rsx_morph_gradient_u8", ptr(xq), H, W, 0, H, ptr(g), 0, H, int(size), stream_ptr())
    return g.cpu().numpy().reshape(H, W) / 255.0


def local_std_dev(band, scale=5):
    """calculate_multi_scale_features(band)['std_dev_scale_<scale>'] (indices.py:531-548): float32 map."""
    require_cuda()
    xf, xq, (H, W) = _second_normalisation(band)
    out = _out_like(xf)
    _lib.call("rsx_local_std_f32", ptr(xf), H, W, 0, H, ptr(out), 0, H, int(scale), None, stream_ptr())
    return out.cpu().numpy().reshape(H, W)


def sobel_magnitude(band):
    """calculate_filter_responses(band)['sobel_mag'] (indices.py:455-480): float32 map in [0, 1]."""
    require_cuda()
    from .device import MinMaxTracker
    xf, xq, (H, W) = _second_normalisation(band)
    out = _out_like(xf)
    mm = MinMaxTracker(1)
    _lib.call("rsx_sobel_mag_u8", ptr(xq), H, W, 0, H, ptr(out), 0, H, ptr(mm.buf), stream_ptr())
    den = np.float32(mm.read()[1][0]) + 1e-10
    _lib.call("rsx_divide_f32", ptr(out), H * W, float(np.float32(den)), stream_ptr())
    return out.cpu().numpy().reshape(H, W)


def prepare_level_2_features(features_dict):
    """indices.py:837-865 (host glue): [glcm contrast, glcm homogeneity, gradient_5, std_dev_scale_5, sobel_mag]."""
    maps = []
    if "glcm_features" in features_dict:
        maps += [features_dict["glcm_features"]["contrast"], features_dict["glcm_features"]["homogeneity"]]
    if "gradient_5" in features_dict.get("morphological_features", {}):
        maps.append(features_dict["morphological_features"]["gradient_5"])
    if "std_dev_scale_5" in features_dict.get("multi_scale_features", {}):
        maps.append(features_dict["multi_scale_features"]["std_dev_scale_5"])
    if "sobel_mag" in features_dict.get("filter_features", {}):
        maps.append(features_dict["filter_features"]["sobel_mag"])
    return np.stack(maps, axis=-1) if maps else np.zeros((1, 1, 1))


def _run_stage_unfused(bands_data):
    """scripts/2_feature_extraction.py:27-133 with preprocessing=False: the bands are used as they are (arbitrary float maps), so
    the stage is the reference's own sequence of calls, each through its drop-in (one host <-> device round trip per call)."""
    blue, green, red, nir, swir1 = bands_data[0], bands_data[1], bands_data[2], bands_data[3], bands_data[4]
    f = {"ndvi": calculate_ndvi(nir, red), "evi": calculate_evi(nir, red, blue), "msavi": calculate_msavi(nir, red),
         "ndwi": calculate_ndwi(green, nir), "mndwi": calculate_mndwi(green, swir1), "ndbi": calculate_ndbi(swir1, nir),
         "bsi": calculate_bsi(blue, red, nir, swir1)}
    valid = [b for b in bands_data if b is not None]
    f["pca_result"], f["variance_ratio"], _ = perform_pca(valid, use_robust_scaling=True)
    f["glcm_features"] = calculate_glcm_features(nir)
    f["morphological_features"] = {"gradient_5": morphological_gradient(nir, 5)}
    f["multi_scale_features"] = {"std_dev_scale_5": local_std_dev(nir, 5)}
    f["filter_features"] = {"sobel_mag": sobel_magnitude(nir)}
    level1 = add_spatial_context(prepare_level_1_features(f))
    level2 = prepare_level_2_features(f)
    return f, {"level_1": level1, "level_2": level2, "all": np.concatenate([level1, level2], axis=-1)}


def run_feature_extraction_stage(bands_data, preprocessing=True, texture_band_index=3):
    """scripts/2_feature_extraction.py:27-133, hot-path part, fused on the device.

    bands_data: list of 2-D integer-valued arrays in TM order (what stage 1 writes: uint8 levels stored as
    float32).  One H2D copy of the packed raster, K1-K4 on the device, one D2H copy of the maps.  Returns
    (features_dict, hierarchical_features) with the reference's keys for the stages on the hot path: the seven
    indices, 'pca_result', 'variance_ratio', 'glcm_features'; hierarchical_features['level_1'] is the 14-channel
    float64 stack add_spatial_context(prepare_level_1_features(...)); for 8-bit input 'level_2' (5 channels) and 'all'
    (19 channels, what all_hierarchical_features.npy holds) are produced too.  Of the LBP / multi-scale / morphology /
    filter dictionaries (scripts/2...:93-107) only the three maps that feed level 2 are computed.
    `texture_band_index` is accepted and ignored, like in the reference (the texture band is always NIR).
    """
    from . import pipeline as P
    require_cuda()
    if not preprocessing:
        return _run_stage_unfused(bands_data)
    arrs = [np.asarray(b) for b in bands_data]
    H, W = arrs[0].shape
    stack = np.stack(arrs, axis=-1)
    if np.isnan(stack).any() or (stack != np.rint(stack)).any() or stack.min() < 0 or stack.max() > 65535:
        raise _lib.RsxError("run_feature_extraction_stage: bands must hold integer levels in [0, 65535] (stage-1 output)")
    packed = stack.astype(np.uint8) if stack.max() <= 255 else stack.astype(np.uint16).view(np.int16)
    raster = torch.from_numpy(np.ascontiguousarray(packed)).cuda()
    fr = P.extract_features(raster, P.FeatureConfig())
    host = fr.planes[:, :fr.n_px].cpu().numpy().reshape(len(fr.names), H, W)
    get = lambda name: host[fr.names.index(name)]
    features = {k: get(k) for k in P.INDEX_NAMES}
    n_comp = fr.pca["components"].shape[0]
    features["pca_result"] = [get(f"pc{i}") for i in range(n_comp)]
    features["variance_ratio"] = fr.pca["explained_variance_ratio"].astype(np.float32)
    features["glcm_features"] = {k: get("glcm_" + k) for k in P.GLCM_NAMES}
    # hierarchical['level_1'] = add_spatial_context(prepare_level_1_features(...)) (scripts/2...:112-119), on the device
    l1, l1_names, _ = P.level1_with_context(fr)
    level1 = np.moveaxis(l1[:, :fr.n_px].cpu().numpy().reshape(len(l1_names), H, W), 0, -1).astype(np.float64)
    hier = {"level_1": level1}
    if packed.dtype == np.uint8:
        # level 2 (scripts/2...:115): the two GLCM maps + the three stencil channels; the reference's arrays are float64
        # (uint8 / 255.0) except std_dev / sobel (float32) - np.stack promotes the stack to float64 either way
        l2, l2_names, _ = P.level2_planes(raster, fr, P.FeatureConfig())
        l2h = l2[:, :fr.n_px].cpu().numpy().reshape(len(l2_names), H, W)
        features["morphological_features"] = {"gradient_5": np.rint(l2h[2].astype(np.float64) * 255.0) / 255.0}
        features["multi_scale_features"] = {"std_dev_scale_5": l2h[3]}
        features["filter_features"] = {"sobel_mag": l2h[4]}
        hier["level_2"] = prepare_level_2_features(features)
        hier["all"] = np.concatenate([level1, hier["level_2"]], axis=-1)
    return features, hier
