"""Device-resident hot path: raster strip -> feature stack -> KMeans labels.

Mirrors the data flow of scripts/2_feature_extraction.py:27-127 (robust_normalize x B, the seven indices,
perform_pca, calculate_glcm_features on the NIR band, stack assembly) and of
modules/features/extract.py:568-577 (MinMaxScaler + KMeans), but keeps everything in HBM:

    raster (h, W, B) uint8/uint16, pixel interleaved            B*e bytes / pixel
      K1 histograms  -> host order statistics (hoststats.RasterStats)
      K2 fused normalise + 7 indices (+ quantised NIR)           planes 0..6 of the stack, q plane (1 B/px)
      K3 PCA moments -> host eigh -> projection                  planes after the GLCM planes
      K4 GLCM props (dense or tiled) -> bilinear upsample        planes 7..11
      K5 KMeans assign/update passes over the first D planes

The stack is planar float32: plane k at planes[k] (stride = padded pixel count).  Every producer also
folds its output's min/max into a tracker so MinMaxScaler.fit costs no extra pass.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, hoststats
from .device import NO_TIMER, MinMaxTracker, StageTimer, fetch, hptr, ptr, require_cuda, stage_to_host, stream_ptr, upload_small
from .dist import Comm, glcm_rows_needed, strip_bounds

INDEX_NAMES = ("ndvi", "evi", "msavi", "ndwi", "mndwi", "ndbi", "bsi")       # RSX plane order (rsx.h)
GLCM_NAMES = ("contrast", "dissimilarity", "homogeneity", "energy", "correlation")


@dataclass
class FeatureConfig:
    band_map: Tuple[int, int, int, int, int] = (0, 1, 2, 3, 4)   # blue, green, red, nir, swir1 (scripts/2...:50-54)
    evi: Tuple[float, float, float, float] = (1.0, 6.0, 7.5, 2.5)  # L, C1, C2, G (indices.py:73)
    n_components: Optional[int] = None                           # perform_pca(n_components=None) -> all bands
    glcm: bool = True
    glcm_levels: int = 32                                        # indices.py:248-249 defaults
    glcm_window: int = 21
    glcm_step: int = 21
    percentiles: Tuple[float, float] = (2, 98)
    # fuse stage 1 (modules/features/preprocessing.py:54-125: gain/bias -> min-max stretch -> uint8) into the load path: the
    # raster holds raw 8-bit DNs; (gain, bias) per band, e.g. (hoststats.TM_GAIN, hoststats.TM_BIAS).  None = raster is
    # already the stage-1 output (what scripts/2_feature_extraction.py reads).
    stage1: Optional[Tuple[Sequence[float], Sequence[float]]] = None
    u16_level_table: bool = False                                # uint16 PCA: exact per-level table instead of reciprocal arithmetic


@dataclass
class FeatureResult:
    planes: torch.Tensor                 # (P, stride) float32
    names: List[str]
    n_px: int
    H: int                               # rows of this strip
    W: int
    stats: hoststats.RasterStats
    pca: dict
    minmax: MinMaxTracker
    quant: Optional[torch.Tensor] = None
    timings: dict = field(default_factory=dict)
    nan_replaced: bool = False           # kmeans_on_features has replaced NaN by 0 in the MSAVI plane (extract.py:548-556)

    def plane(self, name: str) -> torch.Tensor:
        return self.planes[self.names.index(name), :self.n_px].view(self.H, self.W)


def _pad4(n: int) -> int:
    return (n + 31) // 32 * 32


def extract_features(raster: torch.Tensor, cfg: FeatureConfig = FeatureConfig(), comm: Optional[Comm] = None,
                     H_total: Optional[int] = None, bounds: Optional[Sequence[Tuple[int, int]]] = None,
                     timer: StageTimer = NO_TIMER) -> FeatureResult:
    """raster: (h, W, B) uint8 or uint16 CUDA tensor = this rank's row strip of an (H_total, W, B) image."""
    require_cuda()
    comm = comm or Comm()
    assert raster.is_cuda and raster.is_contiguous() and raster.dim() == 3
    h, W, B = raster.shape
    H_total = H_total if H_total is not None else h
    bounds = list(bounds) if bounds is not None else [(0, h)]
    own = bounds[comm.rank]
    assert own[1] - own[0] == h
    n_px = h * W
    n_global = H_total * W
    is16 = raster.dtype in (getattr(torch, 'uint16', torch.int16), torch.int16)
    assert is16 or raster.dtype == torch.uint8
    sfx = "u16" if is16 else "u8"
    L = 65536 if is16 else 256
    dev = raster.device
    st = stream_ptr()

    # ---- K1 histograms (+ all-reduce), host order statistics (phase 1: robust_normalize parameters)
    hist = torch.zeros((B, L), dtype=torch.int32, device=dev)
    if n_px:
        with timer("hist"):
            _lib.call(f"rsx_hist_{sfx}", ptr(raster), n_px, B, ptr(hist), st)
    # allocations first: they overlap the histogram kernel, the host is needed again right after the read-back
    n_comp = B if cfg.n_components is None else int(cfg.n_components)
    names = list(INDEX_NAMES) + (["glcm_" + g for g in GLCM_NAMES] if cfg.glcm else []) + [f"pc{i}" for i in range(n_comp)]
    stride = _pad4(max(n_px, 1))
    planes = torch.empty((len(names), stride), dtype=torch.float32, device=dev)
    mm = MinMaxTracker(len(names), device=dev)
    quant = torch.empty(_pad4(max(n_px, 1)), dtype=torch.uint8, device=dev) if cfg.glcm else None

    # uint8 rasters: the order statistics are computed by a kernel from the device histograms and stay on the device for K2 / K3;
    # the host copy of the histograms (for FeatureResult.stats) is read when the host next has to wait anyway (the PCA moments).
    # (uint16 rasters and the fused stage-1 chain keep the host path: 65536-level tables / level remaps are built there.)
    dev_stats = None
    hist_dev, hist_is64 = hist, 0
    if comm.world > 1:
        with timer("hist_allreduce"):
            hist64 = hist.to(torch.int64) & 0xFFFFFFFF      # the kernel's counters are uint32
            comm.all_reduce(hist64)
        hist_dev, hist_is64 = hist64, 1
    remap = None
    stats = None
    band_map = np.asarray(cfg.band_map, dtype=np.int32)
    evi = np.asarray(cfg.evi, dtype=np.float32)
    if not is16 and cfg.stage1 is None and _lib.get_option("device_stats", 1):
        dev_stats = torch.empty(int(_lib.load().rsx_raster_stats_device_bytes()), dtype=torch.uint8, device=dev)
        gb = int(cfg.band_map[3])
        _lib.call("rsx_raster_stats_u8_device", ptr(hist_dev), hist_is64, B, gb if 0 <= gb < B else -1, float(cfg.percentiles[0]),
                  float(cfg.percentiles[1]), ptr(dev_stats), st)
        hist_pinned = stage_to_host(hist_dev)
        if n_px:
            with timer("indices"):
                _lib.call("rsx_indices_fused_u8_dev", ptr(raster), n_px, B, hptr(band_map), ptr(dev_stats), hptr(evi), ptr(planes), stride,
                          mm.slot(0), ptr(quant), cfg.glcm_levels, st)
    else:
        hist_host = fetch(hist_dev) if hist_is64 else fetch(hist).view(np.uint32).astype(np.int64)
        hist_raw = hist_host
        if cfg.stage1 is not None:
            if is16:
                raise _lib.RsxError("the fused stage-1 chain is defined for 8-bit rasters")
            remap, hist_host = hoststats.stage1_level_tables(hist_host, cfg.stage1[0], cfg.stage1[1])
            remap = np.ascontiguousarray(remap)
        stats = hoststats.RasterStats(hist_host, glcm_band=cfg.band_map[3], lower=cfg.percentiles[0], upper=cfg.percentiles[1])
        stats.hist_raw = hist_raw

        # ---- K2 fused normalise + indices (+ quantised NIR)
        norm = np.ascontiguousarray(stats.norm, dtype=np.float32)
        qnorm = np.ascontiguousarray(stats.qnorm, dtype=np.float32)
        if n_px:
            with timer("indices"):
                if is16:
                    _lib.call("rsx_indices_fused_u16", ptr(raster), n_px, B, hptr(band_map), hptr(norm), hptr(evi), ptr(planes), stride,
                              mm.slot(0), ptr(quant), hptr(qnorm), cfg.glcm_levels, st)
                else:
                    _lib.call("rsx_indices_fused_u8", ptr(raster), n_px, B, hptr(band_map), hptr(norm), hptr(evi), ptr(planes), stride,
                              mm.slot(0), ptr(quant), hptr(qnorm), cfg.glcm_levels, hptr(remap), st)

    # ---- K3a PCA moments (+ all-reduce).  The RobustScaler statistics (host, phase 2) are computed while K2 runs; the
    #      moments come back through pinned memory so that the host can do the eigen-decomposition under the GLCM kernel.
    M = B + B * (B + 1) // 2
    moments = torch.zeros(M, dtype=torch.float64, device=dev)
    scratch = torch.empty(int(_lib.load().rsx_pca_scratch_elems(B)), dtype=torch.float64, device=dev)
    lut = lut_inputs = None
    norm = center = scale = None
    if dev_stats is not None:
        lut = dev_stats[int(_lib.load().rsx_raster_stats_device_lut_offset()):].view(torch.float32)      # [16][256], rows 0..B-1 used
    else:
        center = np.ascontiguousarray(stats.center, dtype=np.float32)
        scale = np.ascontiguousarray(stats.scale, dtype=np.float64)
        norm = np.ascontiguousarray(stats.norm, dtype=np.float32)
        if is16 and cfg.u16_level_table:
            # X per 16-bit level, tabulated on the device with the exact per-sample arithmetic (3.4 MB for 13 bands, L2
            # resident).  Off by default: 13 scattered 4-byte gathers per pixel cost more than the reciprocal arithmetic.
            lut = torch.empty((B, 65536), dtype=torch.float32, device=dev)
            d_norm, d_center, d_scale = upload_small(norm, dev), upload_small(center, dev), upload_small(scale, dev)
            _lib.call("rsx_pca_build_lut_u16", ptr(d_norm), ptr(d_center), ptr(d_scale), B, ptr(lut), st)
            lut_inputs = (d_norm, d_center, d_scale)            # stay referenced until the stream has consumed them
        if not is16:
            x_lut = np.ascontiguousarray(stats.x_lut, dtype=np.float32)
            if remap is not None:                                   # table of the RAW level: x_lut[b][remap[b][v]]
                x_lut = np.ascontiguousarray(np.take_along_axis(x_lut, remap.astype(np.int64), axis=1))
            lut = upload_small(x_lut, dev)
    if n_px:
        with timer("pca_moments"):
            if is16:
                _lib.call("rsx_pca_moments_u16", ptr(raster), n_px, B, hptr(norm), hptr(center), hptr(scale), ptr(lut), ptr(moments),
                          ptr(scratch), st)
            else:
                _lib.call("rsx_pca_moments_u8", ptr(raster), n_px, B, ptr(lut), ptr(moments), ptr(scratch), st)
    if comm.world > 1:
        with timer("pca_allreduce"):
            comm.all_reduce(moments)
    moments_host = _pinned_f64(M)
    _lib.call("rsx_store_to_host", ptr(moments), C.c_void_p(moments_host.data_ptr()), M * 8, st)   # not through the copy engine
    moments_ready = torch.cuda.Event()
    moments_ready.record()

    # ---- K4 GLCM texture on the quantised NIR band (+ halo rows from the strips below/above), upsample
    if cfg.glcm:
        w, s = cfg.glcm_window, cfg.glcm_step
        if H_total < w or W < w:
            raise ValueError(f"GLCM window {w} larger than the {H_total}x{W} image")
        out_rows_total = (H_total - w) // s + 1
        out_cols = (W - w) // s + 1
        needs = [glcm_rows_needed(b, H_total, w, s) for b in bounds]
        (p0, p1), _ = needs[comm.rank]
        q_local = quant[:n_px].view(h, W)
        with timer("glcm_halo"):
            q_ext = comm.fetch_rows(q_local, bounds, [qr for _, qr in needs])
        if p1 > p0:
            props = torch.empty((5, _pad4((p1 - p0) * out_cols)), dtype=torch.float32, device=dev)
            with timer("glcm_props"):
                _lib.call("rsx_glcm_props", ptr(q_ext), q_ext.shape[0], W, cfg.glcm_levels, w, s, p1 - p0, out_cols, ptr(props),
                          props.shape[1], st)
            with timer("glcm_resize"):
                _lib.call("rsx_resize_bilinear_f32", ptr(props), out_rows_total, out_cols, p0, p1 - p0, props.shape[1],
                          C.c_void_p(planes[7].data_ptr()), H_total, W, own[0], h, stride, 5, mm.slot(7), st)

    # ---- K3b eigh on the host (the device is busy with K4), projection
    moments_ready.synchronize()
    if stats is None:                                   # the histograms arrived with (before) the moments
        hp = hist_pinned.numpy()
        hist_host = hp.astype(np.int64) if hist_is64 else hp.view(np.uint32).astype(np.int64)
        stats = hoststats.RasterStats(hist_host, glcm_band=cfg.band_map[3], lower=cfg.percentiles[0], upper=cfg.percentiles[1])
        stats.hist_raw = hist_host
    pca = hoststats.pca_from_moments(moments_host.numpy().copy(), n_global, n_comp)
    comps = np.ascontiguousarray(pca["components"], dtype=np.float32)
    mean32 = pca["mean"].astype(np.float32)
    # sklearn: X_transformed -= mean_ @ components_.T, both float32 for float32 data
    mean_proj = np.ascontiguousarray((mean32.reshape(1, -1) @ comps.T).ravel(), dtype=np.float32)
    pc0 = len(names) - n_comp
    if n_px:
        with timer("pca_project"):
            if is16:
                _lib.call("rsx_pca_project_u16", ptr(raster), n_px, B, hptr(norm), hptr(center), hptr(scale), ptr(lut), hptr(comps),
                          hptr(mean_proj), n_comp, C.c_void_p(planes[pc0].data_ptr()), stride, mm.slot(pc0), st)
            else:
                _lib.call("rsx_pca_project_u8", ptr(raster), n_px, B, ptr(lut), hptr(comps), hptr(mean_proj), n_comp,
                          C.c_void_p(planes[pc0].data_ptr()), stride, mm.slot(pc0), st)
    del lut_inputs
    return FeatureResult(planes=planes, names=names, n_px=n_px, H=h, W=W, stats=stats, pca=pca, minmax=mm, quant=quant)


# ============================================================================================ level-1 stack with spatial context
LEVEL1_NAMES = ("ndwi", "mndwi", "ndvi", "evi", "ndbi", "bsi", "pc0")     # prepare_level_1_features (indices.py:808-835)


def level1_with_context(fr: FeatureResult, window_size: int = 7, comm: Optional[Comm] = None, H_total: Optional[int] = None,
                        bounds: Optional[Sequence[Tuple[int, int]]] = None, timer: StageTimer = NO_TIMER):
    """The reference's real level-1 stack (scripts/2_feature_extraction.py:112-119): the seven maps of
    prepare_level_1_features followed by their window_size x window_size box means (add_spatial_context, indices.py:760-776;
    BORDER_REFLECT), as a planar float32 device stack.  Returns (planes (14, stride), names, MinMaxTracker).
    Row strips: the box filter needs window_size // 2 halo rows from the neighbouring strips."""
    comm = comm or Comm()
    H_total = H_total if H_total is not None else fr.H
    bounds = list(bounds) if bounds is not None else [(0, fr.H)]
    r0, r1 = bounds[comm.rank]
    names = list(LEVEL1_NAMES) + ["ctx_" + n for n in LEVEL1_NAMES]
    n = len(LEVEL1_NAMES)
    dev = fr.planes.device
    out = torch.empty((2 * n, fr.planes.shape[1]), dtype=torch.float32, device=dev)
    mm = MinMaxTracker(2 * n, device=dev)
    src_mn, src_mx = fr.minmax.buf, None
    half = window_size // 2
    needs = [(max(a - half, 0), min(b + half, H_total)) if b > a else (0, 0) for a, b in bounds]
    for k, name in enumerate(LEVEL1_NAMES):
        i = fr.names.index(name)
        out[k].copy_(fr.planes[i])
        mm.buf[k].copy_(fr.minmax.buf[i])
        if fr.n_px == 0:
            continue
        local = fr.planes[i, :fr.n_px].view(fr.H, fr.W)
        ext = comm.fetch_rows(local, bounds, needs)
        if not ext.is_contiguous():
            ext = ext.contiguous()
        with timer("spatial_context"):
            _lib.call("rsx_box_mean_f32", ptr(ext), H_total, fr.W, needs[comm.rank][0], ext.shape[0], ext.numel(),
                      C.c_void_p(out[n + k].data_ptr()), r0, fr.H, out.shape[1], 1, int(window_size), mm.slot(n + k), stream_ptr())
    return out, names, mm


# ============================================================================================ level-2 stack
LEVEL2_NAMES = ("glcm_contrast", "glcm_homogeneity", "gradient_5", "std_dev_scale_5", "sobel_mag")   # indices.py:837-865


def level2_planes(raster: torch.Tensor, fr: FeatureResult, cfg: FeatureConfig = FeatureConfig(), comm: Optional[Comm] = None,
                  H_total: Optional[int] = None, bounds: Optional[Sequence[Tuple[int, int]]] = None, timer: StageTimer = NO_TIMER):
    """prepare_level_2_features (indices.py:837-865) as a planar float32 device stack: GLCM contrast and homogeneity (already
    in `fr`), morphological gradient 5x5, local standard deviation 5x5 and the normalised Sobel magnitude of the texture band
    (calculate_morphological_features :421-440, calculate_multi_scale_features :537-548, calculate_filter_responses :477-480;
    each starts with its own robust_normalize of the already normalised band, like the GLCM).  uint8 rasters only.
    Returns (planes (5, stride), names, MinMaxTracker)."""
    comm = comm or Comm()
    H_total = H_total if H_total is not None else fr.H
    bounds = list(bounds) if bounds is not None else [(0, fr.H)]
    r0, r1 = bounds[comm.rank]
    if raster.dtype != torch.uint8:
        raise _lib.RsxError("level2_planes: the stencil features are implemented for 8-bit rasters")
    if not cfg.glcm:
        raise _lib.RsxError("level2_planes needs the GLCM planes (FeatureConfig.glcm=True)")
    dev, st = raster.device, stream_ptr()
    h, W, B = raster.shape
    n_px = h * W
    nir = cfg.band_map[3]
    # tables of the texture band: second robust_normalize (float32) and its 8-bit quantisation (band * 255).astype(uint8)
    band2 = hoststats.normalize_levels(fr.stats.norm_lut[nir], *fr.stats.qnorm).astype(np.float32)
    q255 = (band2 * 255).astype(np.uint8)
    if cfg.stage1 is not None:
        remap, _ = hoststats.stage1_level_tables(fr.stats.hist_raw, cfg.stage1[0], cfg.stage1[1])
        band2, q255 = band2[remap[nir]], q255[remap[nir]]
    d_band2 = torch.from_numpy(np.ascontiguousarray(band2)).to(dev)
    d_q255 = torch.from_numpy(np.ascontiguousarray(q255)).to(dev)
    out = torch.empty((5, fr.planes.shape[1]), dtype=torch.float32, device=dev)
    mm = MinMaxTracker(5, device=dev)
    for k, name in enumerate(LEVEL2_NAMES[:2]):
        i = fr.names.index(name)
        out[k].copy_(fr.planes[i])
        mm.buf[k].copy_(fr.minmax.buf[i])
    if n_px == 0:
        return out, list(LEVEL2_NAMES), mm
    xf = torch.empty(n_px, dtype=torch.float32, device=dev)
    xq = torch.empty(n_px, dtype=torch.uint8, device=dev)
    with timer("level2_stencils"):
        _lib.call("rsx_band_lut_f32", ptr(raster), n_px, B, nir, ptr(d_band2), ptr(xf), st)
        _lib.call("rsx_band_lut_u8", ptr(raster), n_px, B, nir, ptr(d_q255), ptr(xq), st)
    need2 = [(max(a - 2, 0), min(b + 2, H_total)) if b > a else (0, 0) for a, b in bounds]
    need1 = [(max(a - 1, 0), min(b + 1, H_total)) if b > a else (0, 0) for a, b in bounds]
    xq2 = comm.fetch_rows(xq.view(h, W), bounds, need2).contiguous()
    xf2 = comm.fetch_rows(xf.view(h, W), bounds, need2).contiguous()
    xq1 = comm.fetch_rows(xq.view(h, W), bounds, need1).contiguous()
    grad = torch.empty(n_px, dtype=torch.uint8, device=dev)
    with timer("level2_stencils"):
        _lib.call("rsx_morph_gradient_u8", ptr(xq2), H_total, W, need2[comm.rank][0], xq2.shape[0], ptr(grad), r0, h, 5, st)
        _lib.call("rsx_u8_over_255_f32", ptr(grad), n_px, C.c_void_p(out[2].data_ptr()), st)
        _lib.call("rsx_minmax_planes_f32", C.c_void_p(out[2].data_ptr()), n_px, out.shape[1], 1, mm.slot(2), st)
        _lib.call("rsx_local_std_f32", ptr(xf2), H_total, W, need2[comm.rank][0], xf2.shape[0], C.c_void_p(out[3].data_ptr()), r0, h, 5,
                  mm.slot(3), st)
        _lib.call("rsx_sobel_mag_u8", ptr(xq1), H_total, W, need1[comm.rank][0], xq1.shape[0], C.c_void_p(out[4].data_ptr()), r0, h,
                  mm.slot(4), st)
    # sobel_mag / (sobel_mag.max() + 1e-10): the maximum is global (all strips)
    mn, mx = mm.read()
    smax = torch.tensor([float(mx[4])], dtype=torch.float32, device=dev)
    comm.all_reduce(smax, "max")
    den = np.float32(smax.item()) + 1e-10                      # float32 + Python float stays float32 (NEP 50), as in the reference
    with timer("level2_stencils"):
        _lib.call("rsx_divide_f32", C.c_void_p(out[4].data_ptr()), n_px, float(np.float32(den)), st)
    _lib.call("rsx_minmax_init", mm.slot(4), 1, st)
    _lib.call("rsx_minmax_planes_f32", C.c_void_p(out[4].data_ptr()), n_px, out.shape[1], 1, mm.slot(4), st)
    return out, list(LEVEL2_NAMES), mm


# ============================================================================================ KMeans
def draw_init_indices(n: int, k: int, seed: int) -> np.ndarray:
    """k distinct pixel indices in [0, n), reproducible, without materialising a permutation of n."""
    rng = np.random.default_rng(seed)
    seen, out = set(), []
    while len(out) < k:
        v = int(rng.integers(0, n))
        if v not in seen:
            seen.add(v)
            out.append(v)
    return np.asarray(out, dtype=np.int64)


def minmax_scale_params(fmin: np.ndarray, fmax: np.ndarray):
    """MinMaxScaler.fit in float64 (sklearn/preprocessing/_data.py:527-541)."""
    fmin, fmax = np.asarray(fmin, np.float64), np.asarray(fmax, np.float64)
    rng = fmax - fmin
    rng[rng < 10 * np.finfo(np.float64).eps] = 1.0
    scale = 1.0 / rng
    return scale, 0.0 - fmin * scale


@dataclass
class KMeansResult:
    labels: torch.Tensor          # (n_px,) int32 on the device (this strip)
    centroids: np.ndarray         # (K, D) float64, MinMax-scaled coordinates
    inertia: float
    n_iter: int
    near_ties: int
    shift_sq: float


def gather_rows_device(planes: torch.Tensor, D: int, n_px: int, global_idx: np.ndarray, first_px: int, comm: Comm) -> torch.Tensor:
    """Raw float32 feature rows of the given GLOBAL pixel indices as a (len, D) float64 device tensor (all-reduced so
    every rank has all of them); asynchronous."""
    loc = upload_small(np.asarray(global_idx - first_px, dtype=np.int64), planes.device)
    mine = (loc >= 0) & (loc < n_px)
    sel = torch.where(mine, loc, torch.zeros_like(loc))
    rows = (planes[:D].index_select(1, sel).t().to(torch.float64) * mine.to(torch.float64).unsqueeze(1)).contiguous()
    comm.all_reduce(rows)
    return rows


def _rowsum_like_numpy(sq: torch.Tensor) -> torch.Tensor:
    """(n, D) float64 -> row sums in the order numpy's pairwise summation uses for a contiguous row of fewer than 128 elements
    (np.add.reduce along the last axis): sequential below 8 elements, else eight interleaved partial sums combined as
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) followed by the remaining elements - so that ((X - c)**2).sum(axis=1) of sklearn's
    _relocate_empty_clusters_dense is reproduced bit for bit."""
    n, D = sq.shape
    if D < 8:
        out = sq[:, 0].clone()
        for d in range(1, D):
            out += sq[:, d]
        return out
    r = [sq[:, j].clone() for j in range(8)]
    full = D - D % 8
    for i in range(8, full, 8):
        for j in range(8):
            r[j] += sq[:, i + j]
    out = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    for i in range(full, D):
        out += sq[:, i]
    return out


class DeviceKMeans:
    """Lloyd iterations on a planar float32 stack that stays in HBM."""

    def __init__(self, planes: torch.Tensor, n_px: int, D: int, K: int, feat_min, feat_max, n_global: int, row_len: int,
                 comm: Optional[Comm] = None, timer: StageTimer = NO_TIMER, delta: bool = True, full_passes: Optional[int] = None,
                 bounded: Optional[bool] = None):
        require_cuda()
        self.comm = comm or Comm()
        self.timer = timer
        self.planes, self.n_px, self.D, self.K = planes, int(n_px), int(D), int(K)
        self.stride = planes.stride(0)
        self.row_len = int(row_len)
        self.n_global = int(n_global)
        if feat_min is not None:
            self.configure(feat_min, feat_max)
        dev = planes.device
        self.state = torch.zeros(int(_lib.load().rsx_kmeans_state_bytes()), dtype=torch.uint8, device=dev)
        # pass sums/counts + near-tie and changed-label counters (all-reduced), then the running totals (rsx.h)
        self.n_acc = K * D + K + 2
        self.acc = torch.zeros(2 * (K * D + K + 2), dtype=torch.int64, device=dev)
        self.inertia = torch.zeros(1, dtype=torch.float64, device=dev)
        # delta passes: after the first (full) pass only the pixels whose label changed move their sample between the
        # clusters' integer sums - the same totals as a full pass, without touching the accumulators for the rest
        self.delta = bool(delta)
        # K <= 8: the first passes relabel > 7 % of the pixels, where recomputing the sums with the per-thread accumulators
        # (0.72 ms at 49 Mpx) beats moving that many samples; any split gives the same integers
        # K <= 8: the delta passes carry Hamerly's bound test (rsx_kmeans_assign_bounded): pixels that provably keep their label
        # are skipped unread; the others are gathered from a pixel-interleaved copy of the stack (scratch: 4 * (aos_stride + 1)
        # bytes per pixel).  Same labels, sums and counters as the unbounded passes.
        if bounded is None:
            bounded = bool(_lib.get_option("km_bounded", 0))
        self.bounded = bool(bounded) and self.delta and K <= 8
        if full_passes is None:
            full_passes = _lib.get_option("km_full_passes", 3)
        self.full_passes = max(1, int(full_passes))
        self._aos = self._slack = self._bound_plane = None
        self._bounds_valid = False
        # K <= 8: late delta passes read a 16-bit copy of the stack (half the bytes; rsx_kmeans_assign_q16) once few enough labels
        # move that fetching the float32 samples of the moving pixels costs less than the bytes saved
        self.q16_from = int(_lib.get_option("km_q16_from", 6)) if (_lib.get_option("km_q16", 0) and self.delta and K <= 8) else -1
        self._q16 = None
        self._q16_valid = False
        self._labels = None
        self._passes = 0
        # several GPUs: the update kernels reduce the ranks' sums themselves through peer-mapped memory (no collective per
        # iteration); fit_converge, which inspects the reduced counts on the host before the update, keeps the all-reduce
        self.peers = self.comm.peers() if self.comm.world > 1 else None
        self._peer_seq = 0

    def configure(self, feat_min, feat_max):
        """MinMaxScaler.fit result (per-feature min / max of the raw stack).  May follow the constructor (feat_min=None) so
        that the device buffers are allocated while the feature kernels are still running."""
        self.fmin = np.ascontiguousarray(feat_min, dtype=np.float64)
        self.fmax = np.ascontiguousarray(feat_max, dtype=np.float64)
        self.scale, self.min_ = minmax_scale_params(self.fmin, self.fmax)

    def scale_rows(self, raw_rows: np.ndarray) -> np.ndarray:
        """MinMaxScaler.transform of raw feature rows in float64: X*scale + min_."""
        return np.asarray(raw_rows, np.float64) * self.scale + self.min_

    def gather_rows(self, global_idx: np.ndarray, first_px: int) -> np.ndarray:
        """Raw float32 feature rows of the given GLOBAL pixel indices (all-reduced so every rank has all of them)."""
        return gather_rows_device(self.planes, self.D, self.n_px, global_idx, first_px, self.comm).cpu().numpy()

    def setup(self, init_centroids_scaled: np.ndarray, mean_scaled: Optional[np.ndarray] = None):
        c0 = np.ascontiguousarray(init_centroids_scaled, dtype=np.float64)
        assert c0.shape == (self.K, self.D)
        # any centring origin gives the same labels in exact arithmetic; 0.5 minimises the fp32 rounding bound
        mu = np.full(self.D, 0.5) if mean_scaled is None else np.ascontiguousarray(mean_scaled, dtype=np.float64)
        self._mu = mu
        self.acc.zero_()
        self._passes = 0
        self._bounds_valid = False
        self._q16_valid = False
        if self.peers is not None:
            self.peers.zero(stream_ptr())
        _lib.call("rsx_kmeans_setup", ptr(self.state), self.D, self.K, hptr(self.fmin), hptr(self.fmax), hptr(mu), hptr(c0),
                  self.n_global, stream_ptr())

    def setup_device(self, minmax_buf: torch.Tensor, rows_raw: torch.Tensor, mean_scaled: Optional[np.ndarray] = None):
        """setup() without the host in between (rsx_kmeans_setup_device): per-feature range from the device min/max trackers
        (uint32 [>= D][2], already merged over the ranks), initial centroids from raw feature rows on the device ((K, D)
        float64).  No synchronisation; configure_from_device() fills fmin / fmax / scale afterwards."""
        assert rows_raw.dtype == torch.float64 and rows_raw.is_contiguous() and tuple(rows_raw.shape) == (self.K, self.D)
        mu = np.full(self.D, 0.5) if mean_scaled is None else np.ascontiguousarray(mean_scaled, dtype=np.float64)
        self._mu = mu
        self.acc.zero_()
        self._passes = 0
        self._bounds_valid = False
        self._q16_valid = False
        if self.peers is not None:
            self.peers.zero(stream_ptr())
        _lib.call("rsx_kmeans_setup_device", ptr(self.state), self.D, self.K, ptr(minmax_buf), ptr(rows_raw), hptr(mu), self.n_global,
                  stream_ptr())

    def configure_from_device(self):
        """fmin / fmax / scale / min_ as the device set-up derived them (from the last read() of the state)."""
        self.configure(*self._dev_range)

    def fit_device_init(self, minmax_buf: torch.Tensor, rows_raw: torch.Tensor, n_iter: int, labels_i32: bool = True, first_px: int = 0):
        """fit() with the set-up done on the device: nothing between the feature kernels and the last KMeans pass waits for the
        host.  Returns (result, scaled initial centroids)."""
        self.setup_device(minmax_buf, rows_raw)
        rows_host = stage_to_host(rows_raw)
        for _ in range(n_iter):
            self.step()
        res = self._result(self.finish(labels_i32), n_iter)                  # the one synchronisation
        self.configure_from_device()
        c0 = self.scale_rows(rows_host.numpy())
        if res is None:
            res = self._fit_relocating(c0, n_iter, labels_i32, first_px)
        return res, c0

    def _label_planes(self):
        if self._labels is None:
            npad = (self.n_px + 3) // 4 * 4
            self._labels = [torch.full((max(npad, 4),), 255, dtype=torch.uint8, device=self.planes.device) for _ in range(2)]
        return self._labels

    def assign_pass(self, track_labels: bool = False, peer_reduce: bool = False) -> int:
        """One fused assign + partial-sum pass and the all-reduce of its K*(D+1)+2 integers.  The first pass accumulates from
        scratch; later passes are delta passes (if enabled).  With track_labels (or delta) every pass writes uint8 labels
        and counts the pixels whose label changed.  Returns the mode used (1 = full, 2 = delta)."""
        use_labels = track_labels or self.delta
        # K <= 8: the first full_passes passes recompute the sums (per-thread accumulators); K > 8: only the very first one
        # (the streaming kernel's "every pixel moves in" mode, which adds runs of equal labels before touching the sums)
        mode = 2 if (self.delta and self._passes >= (self.full_passes if self.K <= 8 else 1)) else 1
        cur = prev = None
        if use_labels:
            planes = self._label_planes()
            if self._passes == 0:
                planes[1].fill_(255)
            cur, prev = planes[self._passes % 2], planes[(self._passes + 1) % 2]
        self._cur_labels = cur
        self._peer_seq = 0
        d_acc = ptr(self.acc)
        if peer_reduce and self.peers is not None:
            self._peer_seq, d_acc = self.peers.next_pass()       # this pass accumulates into the peer-mapped block
        if mode == 2 and self.bounded:
            # labels in place: the plane that holds the previous pass (the first bounded pass), the same one from then on
            first = not self._bounds_valid
            lab = prev if first else self._bound_plane
            self._bound_plane = self._cur_labels = cur = lab
            if self.n_px:
                if self._aos is None:
                    stride = int(_lib.load().rsx_kmeans_aos_stride(self.D))
                    self._aos = torch.empty(self.n_px * stride, dtype=torch.float32, device=self.planes.device)
                    self._slack = torch.empty((self.n_px + 3) // 4 * 4, dtype=torch.float32, device=self.planes.device)
                with self.timer("kmeans_assign_bounded_first" if first else "kmeans_assign_delta"):
                    _lib.call("rsx_kmeans_assign_bounded", ptr(self.planes), self.stride, self.n_px, ptr(self.state), d_acc, ptr(lab),
                              ptr(self._aos), ptr(self._slack), 1 if first else 0, self.D, self.K, stream_ptr())
            self._bounds_valid = True
        elif mode == 2 and 0 <= self.q16_from <= self._passes and self.n_px:
            self._bounds_valid = False
            if not self._q16_valid:
                qs = (self.n_px + 7) // 8 * 8
                if self._q16 is None:
                    self._q16 = torch.empty((self.D, qs), dtype=torch.int16, device=self.planes.device)
                with self.timer("kmeans_quantize"):
                    _lib.call("rsx_kmeans_quantize_u16", ptr(self.planes), self.stride, self.n_px, ptr(self.state), ptr(self._q16), qs, self.D,
                              stream_ptr())
                self._q16_valid = True
            with self.timer("kmeans_assign_delta"):
                _lib.call("rsx_kmeans_assign_q16", ptr(self.planes), self.stride, self.n_px, ptr(self.state), d_acc, ptr(cur), ptr(prev),
                          ptr(self._q16), self._q16.shape[1], self.D, self.K, stream_ptr())
        elif self.n_px:
            self._bounds_valid = False
            with self.timer("kmeans_assign_delta" if mode == 2 else "kmeans_assign_full"):
                _lib.call("rsx_kmeans_assign", ptr(self.planes), self.stride, self.n_px, self.row_len, ptr(self.state), d_acc,
                          ptr(cur), ptr(prev), None, None, mode, self.D, self.K, stream_ptr())
        if self.comm.world > 1 and not self._peer_seq:
            with self.timer("kmeans_allreduce"):
                self.comm.all_reduce(self.acc[:self.n_acc])
        return mode

    def update(self, mode: int, adjust: Optional[torch.Tensor] = None):
        """Centroids <- totals / counts (the pass block is folded into the totals first)."""
        with self.timer("kmeans_update"):
            if self._peer_seq:
                _lib.call("rsx_kmeans_update_peers", ptr(self.state), ptr(self.acc), 1 if mode == 2 else 0, self.D, ptr(adjust),
                          self.peers.ptrs, self.peers.rank, self.peers.world, self._peer_seq, stream_ptr())
            else:
                _lib.call("rsx_kmeans_update", ptr(self.state), ptr(self.acc), 1 if mode == 2 else 0, self.D, ptr(adjust), stream_ptr())
        self._passes += 1

    def step(self, track_labels: bool = False):
        """assign_pass + update, without host synchronisation (empty clusters are only reported, see fit_converge)."""
        self.update(self.assign_pass(track_labels, peer_reduce=True))

    def relocation_adjust(self, empty: np.ndarray, first_px: int = 0) -> torch.Tensor:
        """sklearn's _relocate_empty_clusters_dense (_k_means_common.pyx:167-211) for the pass that has just been all-reduced:
        every empty cluster takes the sample that is farthest from its own (old) centre; that sample leaves its cluster's
        sum.  Returns the int64 adjustment [K*D + K] for rsx_kmeans_update.  Rare, host assisted, synchronises.
        The distances are float64 with numpy's summation order; with one rank several simultaneous empties are handed out in
        sklearn's own order (np.argpartition on the whole array, on the host).  With several ranks no rank holds the whole
        array: the farthest samples go to the empty clusters in order of decreasing distance (ties by pixel index)."""
        K, D, dev = self.K, self.D, self.planes.device
        empties = np.flatnonzero(empty)
        n_e = len(empties)
        cent_old, _, _ = self.read()
        mu = torch.from_numpy(self._mu).to(dev)
        C = torch.from_numpy(cent_old).to(dev) - mu                       # centred coordinates, like sklearn's X and centers
        scale, min_ = torch.from_numpy(self.scale).to(dev), torch.from_numpy(self.min_).to(dev)
        labels = self._cur_labels
        cand = torch.full((n_e, 3 + D), -1.0, dtype=torch.float64, device=dev)     # distance, global index, label, raw sample
        dist = None
        if self.n_px:
            dist = torch.empty(self.n_px, dtype=torch.float64, device=dev)
            chunk = 1 << 22
            for a in range(0, self.n_px, chunk):
                b = min(self.n_px, a + chunk)
                Xs = (self.planes[:D, a:b].t().to(torch.float64) * scale + min_) - mu
                dist[a:b] = _rowsum_like_numpy((Xs - C[labels[a:b].long()]) ** 2)
        if self.comm.world == 1 and n_e > 1 and dist is not None:
            # Several empty clusters: sklearn hands the far samples out in the order np.argpartition leaves them in, which is a
            # property of numpy's introselect on the WHOLE distance array - reproduced by running it on the host (rare event:
            # one D2H of the distances).  Several ranks: no rank holds the whole array; see below.
            far = np.argpartition(dist.cpu().numpy(), -n_e)[:-n_e - 1:-1]
            idx = torch.from_numpy(np.ascontiguousarray(far)).to(dev)
            cand[:, 0] = torch.arange(n_e, 0, -1, dtype=torch.float64, device=dev)        # keeps this order in the sort below
            cand[:, 1] = (idx + first_px).to(torch.float64)
            cand[:, 2] = labels[idx].to(torch.float64)
            cand[:, 3:] = self.planes[:D].index_select(1, idx).t().to(torch.float64)
        elif dist is not None:
            k = min(n_e, self.n_px)
            vals, idx = torch.topk(dist, k)
            cand[:k, 0] = vals
            cand[:k, 1] = (idx + first_px).to(torch.float64)
            cand[:k, 2] = labels[idx].to(torch.float64)
            cand[:k, 3:] = self.planes[:D].index_select(1, idx).t().to(torch.float64)
        if self.comm.world > 1:
            gathered = [torch.empty_like(cand) for _ in range(self.comm.world)]
            self.comm.dist.all_gather(gathered, cand, group=self.comm.group)
            cand = torch.cat(gathered, dim=0)
        c = cand.cpu().numpy()
        c = c[c[:, 0] >= 0]
        order = np.lexsort((c[:, 1], -c[:, 0]))[:n_e]                               # farthest first, ties by pixel index
        pow2 = np.zeros(D, np.float64)
        _lib.call("rsx_kmeans_fixed_point_scales", ptr(self.state), hptr(pow2), stream_ptr())
        adj = np.zeros(K * D + K, np.int64)
        for new, row in zip(empties, c[order]):
            old = int(row[2])
            q = np.rint(row[3:].astype(np.float32) * pow2.astype(np.float32)).astype(np.int64)   # as the kernels: rint(x * 2^shift)
            adj[old * D:(old + 1) * D] -= q
            adj[K * D + old] -= 1
            adj[new * D:(new + 1) * D] += q
            adj[K * D + new] += 1
        return torch.from_numpy(adj).to(dev)

    def changed_count(self) -> int:
        """Pixels (all ranks) whose label changed in the last update pass; synchronises."""
        return int(fetch(self.acc[2 * self.n_acc - 1:])[0])

    def near_ties(self) -> int:
        """Pixels (all ranks, all passes) decided by the float64 re-evaluation; collective, synchronises."""
        t = self.acc[self.n_acc - 2:self.n_acc - 1].clone()          # the final assign-only pass is not folded by an update
        self.comm.all_reduce(t)
        return int(fetch(t)[0]) + int(fetch(self.acc[2 * self.n_acc - 2:2 * self.n_acc - 1])[0])

    def _result(self, labels, n_iter: int) -> KMeansResult:
        """Collects centroids, inertia and counters with ONE synchronisation (the stores into page-locked memory are queued
        first; reading the state waits for the stream)."""
        t = self.acc[self.n_acc - 2:self.n_acc - 1].clone()          # near ties of the final assign-only pass
        self.comm.all_reduce(t)
        h_inertia, h_last, h_sofar = stage_to_host(self.inertia), stage_to_host(t), stage_to_host(self.acc[2 * self.n_acc - 2:2 * self.n_acc - 1])
        failure = None
        try:
            cent, shift, empty = self.read()
        except _lib.RsxError as ex:                                    # e.g. a peer never reached the update barrier (time-out)
            failure = ex
        if self.comm.world > 1:
            # a failed exchange must fail on EVERY rank: the late rank itself sees all flags and would carry on alone
            flag = torch.tensor([1 if failure is not None else 0], dtype=torch.int32, device=self.planes.device)
            self.comm.all_reduce(flag, "max")
            if int(flag.item()) and failure is None:
                failure = _lib.RsxError("KMeans: the peer-memory exchange failed on another rank (rsx_kmeans_update_peers timed out)")
        if failure is not None:
            raise failure
        if empty:
            return None                                                # the caller reruns with relocation (fit -> _fit_relocating)
        return KMeansResult(labels=labels, centroids=cent, inertia=float(h_inertia[0]), n_iter=n_iter,
                            near_ties=int(h_last[0]) + int(h_sofar[0]), shift_sq=shift)

    def finish(self, labels_i32: bool = True):
        """The extra assignment pass of sklearn (_kmeans.py:742-754) + inertia."""
        dev = self.planes.device
        lab = torch.empty((self.n_px + 3) // 4 * 4, dtype=torch.int32 if labels_i32 else torch.uint8, device=dev)
        self.inertia.zero_()
        if self.n_px:
            with self.timer("kmeans_final"):
                _lib.call("rsx_kmeans_assign", ptr(self.planes), self.stride, self.n_px, self.row_len, ptr(self.state), ptr(self.acc),
                          None if labels_i32 else ptr(lab), None, ptr(lab) if labels_i32 else None, ptr(self.inertia), 0, self.D,
                          self.K, stream_ptr())
        self.comm.all_reduce(self.inertia)
        return lab[:self.n_px]

    def read(self):
        cent = np.zeros((self.K, self.D), np.float64)
        shift = np.zeros(1, np.float64)
        empty = np.zeros(1, np.int32)
        fmin, fmax = np.zeros(self.D, np.float64), np.zeros(self.D, np.float64)
        _lib.call("rsx_kmeans_read_all", ptr(self.state), hptr(cent), hptr(shift), hptr(empty), hptr(fmin), hptr(fmax), stream_ptr())
        self._dev_range = (fmin, fmax)
        return cent, float(shift[0]), int(empty[0])

    def fit_converge(self, init_centroids_scaled: np.ndarray, max_iter: int = 300, tol: float = 0.0,
                     mean_scaled: Optional[np.ndarray] = None, first_px: int = 0) -> KMeansResult:
        """sklearn's _kmeans_single_lloyd stopping rules (_kmeans.py:703-754): stop when no label changed between two
        passes (strict convergence) or when the squared centre shift is <= tol; then the final assignment + inertia."""
        self.setup(init_centroids_scaled, mean_scaled)
        n_iter = 0
        KD = self.K * self.D
        for it in range(max_iter):
            mode = self.assign_pass(track_labels=True)
            blocks = self.acc.cpu().numpy()                    # the one synchronisation of an iteration
            counts = blocks[KD:KD + self.K] + (blocks[self.n_acc + KD:self.n_acc + KD + self.K] if mode == 2 else 0)
            changed = int(blocks[self.n_acc - 1])
            adjust = self.relocation_adjust(counts == 0, first_px) if (counts == 0).any() else None
            self.update(mode, adjust)
            n_iter = it + 1
            if changed == 0:
                break
            _, shift, _ = self.read()
            if shift <= tol:
                break
        res = self._result(self.finish(True), n_iter)
        if res is None:
            raise _lib.RsxError("KMeans: a cluster stayed empty after relocation (fewer distinct samples than clusters?)")
        return res

    def fit(self, init_centroids_scaled: np.ndarray, n_iter: int, labels_i32: bool = True, first_px: int = 0) -> KMeansResult:
        """Fixed-iteration protocol: exactly n_iter update passes without a host synchronisation, then the final assignment.
        If a cluster ran empty on the way (reported by the state, collectively on every rank), the run is repeated with
        sklearn's empty-cluster relocation, which needs the reduced counts on the host after every pass."""
        self.setup(init_centroids_scaled)
        for _ in range(n_iter):
            self.step()
        res = self._result(self.finish(labels_i32), n_iter)
        return res if res is not None else self._fit_relocating(init_centroids_scaled, n_iter, labels_i32, first_px)

    def _fit_relocating(self, init_centroids_scaled: np.ndarray, n_iter: int, labels_i32: bool, first_px: int) -> KMeansResult:
        self.setup(init_centroids_scaled)
        KD = self.K * self.D
        for _ in range(n_iter):
            mode = self.assign_pass(track_labels=True)
            blocks = self.acc.cpu().numpy()                    # the one synchronisation of an iteration
            counts = blocks[KD:KD + self.K] + (blocks[self.n_acc + KD:self.n_acc + KD + self.K] if mode == 2 else 0)
            adjust = self.relocation_adjust(counts == 0, first_px) if (counts == 0).any() else None
            self.update(mode, adjust)
        res = self._result(self.finish(labels_i32), n_iter)
        if res is None:
            raise _lib.RsxError("KMeans: a cluster stayed empty after relocation (fewer distinct samples than clusters?)")
        return res


def kmeans_on_features(fr: FeatureResult, D: int, K: int, n_iter: int, seed: int, comm: Optional[Comm] = None,
                       H_total: Optional[int] = None, first_row: int = 0, labels_i32: bool = True, timer: StageTimer = NO_TIMER,
                       delta: bool = True):
    """Benchmark protocol: MinMax from the fused trackers, K initial centroids = seeded pixel
    rows of the scaled stack, exactly n_iter update passes, then the final assignment + inertia."""
    comm = comm or Comm()
    H_total = H_total if H_total is not None else fr.H
    n_global = H_total * fr.W
    # NaN -> 0 before MinMaxScaler / KMeans (extract.py:548-556).  Of the planes this pipeline makes only MSAVI can be NaN for a
    # finite raster (float32 radicand rounding below zero, indices.py:109-112); the plane is modified in place and 0 joins its
    # tracked range when a NaN was replaced.
    if fr.n_px and "msavi" in fr.names[:D] and not getattr(fr, "nan_replaced", False):
        i = fr.names.index("msavi")
        with timer("nan_to_zero"):
            _lib.call("rsx_nan_to_zero_minmax_f32", C.c_void_p(fr.planes[i].data_ptr()), fr.n_px, fr.minmax.slot(i), stream_ptr())
        fr.nan_replaced = True
    idx = draw_init_indices(n_global, K, seed)
    km = DeviceKMeans(fr.planes, fr.n_px, D, K, None, None, n_global, fr.W, comm, timer, delta)   # buffers first, ...
    if km.delta:
        km._label_planes()
    rows = gather_rows_device(fr.planes, D, fr.n_px, idx, first_row * fr.W, comm)                      # ... all asynchronous
    if _lib.get_option("km_device_setup", 1):
        # MinMaxScaler.fit and the scaling of the initial centroids happen in the set-up kernel, from the trackers the feature
        # kernels maintained: the device never waits for the host between the last feature kernel and the last KMeans pass
        res, c0 = km.fit_device_init(fr.minmax.merged(comm), rows, n_iter, labels_i32, first_px=first_row * fr.W)
        return res, km, c0
    rows_host = stage_to_host(rows)
    mn, mx = fr.minmax.read(comm)                   # the one synchronisation between the feature kernels and KMeans
    km.configure(mn[:D], mx[:D])
    c0 = km.scale_rows(rows_host.numpy())
    res = km.fit(c0, n_iter, labels_i32, first_px=first_row * fr.W)
    return res, km, c0


# ============================================================================================ public entry: host raster -> labels
def segment_raster(raster_host: np.ndarray, cfg: FeatureConfig = FeatureConfig(), n_clusters: int = 8, n_iter: int = 20, seed: int = 42,
                   stack_depth: Optional[int] = None, comm: Optional[Comm] = None, H_total: Optional[int] = None,
                   bounds: Optional[Sequence[Tuple[int, int]]] = None, timer: StageTimer = NO_TIMER, pinned: Optional[torch.Tensor] = None,
                   out_pinned: Optional[torch.Tensor] = None):
    """End-to-end call with HOST buffers: (h, W, B) uint8/uint16 numpy strip in, (h, W) int32 numpy labels out.

    Copies the strip to the device (pinned staging), runs extract_features + the fixed-iteration KMeans protocol,
    copies the labels back.  Returns (labels, KMeansResult, FeatureResult); `labels` lives in a reusable host buffer that is
    valid until the next call with a raster of the same size."""
    require_cuda()
    comm = comm or Comm()
    if pinned is None:
        a = np.ascontiguousarray(raster_host)
        if a.dtype == np.uint16:
            a = a.view(np.int16)
        pinned = torch.from_numpy(a).pin_memory()
    dev_raster = pinned.cuda(non_blocking=True)
    fr = extract_features(dev_raster, cfg, comm, H_total, bounds, timer)
    D = stack_depth if stack_depth is not None else (13 if cfg.glcm else 7 + min(6, len(fr.names) - 7))
    first_row = bounds[comm.rank][0] if bounds is not None else 0
    res, km, c0 = kmeans_on_features(fr, D, n_clusters, n_iter, seed, comm, H_total, first_row, True, timer)
    # labels come back through page-locked memory (a pageable destination costs several times the copy itself)
    if out_pinned is None or out_pinned.numel() != fr.n_px or out_pinned.dtype != torch.int32:
        out_pinned = _pinned_labels(fr.n_px)
    out_pinned.copy_(res.labels, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out_pinned.numpy().reshape(fr.H, fr.W), res, fr


_PINNED_LABELS = {}
_PINNED_F64 = {}
_STREAM_LABELS = {}


def _stream_label_buffers(n: int, pinned_i32: bool = False):
    """Three page-locked uint8 download buffers + three int32 host arrays for segment_stream (scene i is widened by a worker while
    scene i-1 is with the caller and scene i-2 may still be referenced), kept between calls (first-touch page faults of a fresh
    196 MB array cost more than the widening itself)."""
    bufs = _STREAM_LABELS.get((n, pinned_i32))
    if bufs is None:
        _STREAM_LABELS.clear()
        out8 = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(3)]
        out = [torch.zeros(n, dtype=torch.int32, pin_memory=True) if pinned_i32 else np.zeros(n, dtype=np.int32) for _ in range(3)]
        bufs = _STREAM_LABELS[(n, pinned_i32)] = (out8, out)
    return bufs


def _pinned_f64(n: int) -> torch.Tensor:
    """Reusable page-locked float64 staging buffer (valid until the next call of the same size)."""
    buf = _PINNED_F64.get(n)
    if buf is None:
        buf = _PINNED_F64[n] = torch.empty(n, dtype=torch.float64, pin_memory=True)
    return buf


def _pinned_labels(n: int) -> torch.Tensor:
    """Reusable page-locked int32 staging buffer for the label image (valid until the next call of the same size)."""
    buf = _PINNED_LABELS.get(n)
    if buf is None:
        _PINNED_LABELS.clear()
        buf = _PINNED_LABELS[n] = torch.empty(n, dtype=torch.int32, pin_memory=True)
    return buf


def segment_stream(rasters, cfg: FeatureConfig = FeatureConfig(), n_clusters: int = 8, n_iter: int = 20, seed: int = 42,
                   stack_depth: Optional[int] = None, comm: Optional[Comm] = None, H_total: Optional[int] = None,
                   bounds: Optional[Sequence[Tuple[int, int]]] = None, labels: str = "int32"):
    """Pipelined end-to-end path for a sequence of scenes (or of this rank's strips of them), all of one shape: an iterable of
    page-locked host rasters in, a generator of (labels (h, W) int32 numpy, KMeansResult) out, in order (KMeansResult.labels is
    the uint8 device plane here).

    The host-to-device copy of scene i+1 and the device-to-host copy of the labels of scene i-1 run on their own streams
    (the two copy engines) under the kernels of scene i, so a step costs max(compute, copies) instead of their sum.
    A yielded label array lives in one of three reusable host buffers: it is valid until two more scenes have been yielded.

    labels: "int32" (default) - the dtype the reference returns (extract.py:577), written by the device and copied as such;
    "uint8" - the device's own uint8 plane (K <= 64), a quarter of the bytes, for callers that go on to the uint8 GeoTIFF
    (scripts/3_classification.py:394) anyway; "int32_host_widen" - uint8 over PCIe, widened to int32 by host threads in a worker
    (rsx_widen_u8_to_i32): wins where PCIe is the limit, loses where host memory bandwidth is (8 ranks on one socket: the widening
    adds 2 GB of host traffic per step to a box that is already bound by it - 34 -> 50 ms per step measured)."""
    if labels not in ("int32", "uint8", "int32_host_widen"):
        raise ValueError("labels must be 'int32', 'uint8' or 'int32_host_widen'")
    require_cuda()
    comm = comm or Comm()
    compute = torch.cuda.current_stream()
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    it = iter(rasters)

    def as_pinned(r):
        if isinstance(r, np.ndarray):
            a = np.ascontiguousarray(r)
            r = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
        return r if r.is_pinned() else r.pin_memory()

    first = next(it, None)
    if first is None:
        return
    first = as_pinned(first)
    dev = [torch.empty(first.shape, dtype=first.dtype, device="cuda") for _ in range(2)]       # double-buffered device rasters
    n_lab = first.shape[0] * first.shape[1]
    # labels leave the device as uint8 (a quarter of the int32 bytes over PCIe) and are widened into the int32 array the
    # reference's callers expect (extract.py:577) by a few host threads while the next scene's kernels run
    out8, out = _stream_label_buffers(n_lab, pinned_i32=labels == "int32")
    widen_threads = max(1, min(8, (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", comm.world)))))

    import threading

    def start_widen(k):
        """Worker: waits for the uint8 download of slot k and widens it into out[k], off the launching thread's critical path
        (both the event wait and the C call release the GIL)."""
        def run():
            downloaded[k].synchronize()
            _lib.call("rsx_widen_u8_to_i32", C.c_void_p(out8[k].data_ptr()), hptr(out[k]), n_lab, widen_threads)
        if labels != "int32_host_widen":
            return None
        th = threading.Thread(target=run, daemon=True)
        th.start()
        return th

    def host_labels(k, worker):
        if worker is not None:
            worker.join()
        else:
            downloaded[k].synchronize()
        return out8[k].numpy() if labels == "uint8" else (out[k].numpy() if labels == "int32" else out[k])

    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]       # compute no longer reads dev[k]
    downloaded = [torch.cuda.Event() for _ in range(3)]

    # The copy engines serve one transfer at a time, in order: nothing small may queue there while a scene is being computed.
    # Every per-scene read-back (histograms, moments, min/max, KMeans state) is therefore a kernel store into page-locked
    # memory (device.fetch / rsx_store_to_host); behind a 196 MB label download it would wait 3.5 ms.
    def upload(host, k, wait_consumed):
        with torch.cuda.stream(up):
            if wait_consumed:
                up.wait_event(consumed[k])
            dev[k].copy_(host, non_blocking=True)
            uploaded[k].record(up)

    upload(first, 0, False)
    pending = None                                          # (slot, KMeansResult, H, W) whose labels are on their way to the host
    i = 0
    cur = first
    while cur is not None:
        k = i & 1
        nxt = next(it, None)
        if nxt is not None:
            nxt = as_pinned(nxt)
            upload(nxt, k ^ 1, i >= 1)                     # dev[k^1] was read by scene i-1
        compute.wait_event(uploaded[k])
        fr = extract_features(dev[k], cfg, comm, H_total, bounds)
        D = stack_depth if stack_depth is not None else (13 if cfg.glcm else 7 + min(6, len(fr.names) - 7))
        first_row = bounds[comm.rank][0] if bounds is not None else 0
        res, km, c0 = kmeans_on_features(fr, D, n_clusters, n_iter, seed, comm, H_total, first_row, labels == "int32")
        consumed[k].record(compute)
        done = torch.cuda.Event()
        done.record(compute)
        slot = i % 3
        with torch.cuda.stream(down):
            down.wait_event(done)
            (out[slot] if labels == "int32" else out8[slot]).copy_(res.labels, non_blocking=True)
            downloaded[slot].record(down)
        res.labels.record_stream(down)
        worker = start_widen(slot)
        if pending is not None:                             # hand out scene i-1 (its download overlapped this scene's kernels)
            pk, pres, ph, pw, pth = pending
            yield host_labels(pk, pth).reshape(ph, pw), pres
        pending = (slot, res, fr.H, fr.W, worker)
        cur = nxt
        i += 1
    pk, pres, ph, pw, pth = pending
    yield host_labels(pk, pth).reshape(ph, pw), pres


def segment_stream_d2h_bytes(n_px: int, labels: str = "int32") -> int:
    """Bytes per scene that segment_stream copies from the device to the host."""
    return int(n_px) * (4 if labels == "int32" else 1)
