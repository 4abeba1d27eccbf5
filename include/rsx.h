/* rsx - C ABI of the B200-native feature-stack + KMeans hot path.
 *
 * The reference (beilsme/rs-image-segmentation) is pure Python and has no FFI; the
 * boundary it offers is the set of module-level functions of
 * modules/features/indices.py and modules/features/extract.py.  This header is what a
 * ctypes binding for those functions links against; each entry
 * point names the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (the Python shim passes
 *     torch tensor data pointers); h_* is host memory; nothing is allocated behind the
 *     caller's back except where stated.
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it and no call
 *     synchronises the device.
 *   - return value 0 = ok; otherwise an RSX_ERR_* code and rsx_last_error() describes it
 *     (thread-local string).
 *   - rasters are PIXEL-INTERLEAVED ("BIP"): sample (pixel p, band b) at d_raster[p*B + b].
 *   - feature maps are PLANAR float32: plane k starts at d_out + k*plane_stride (elements).
 *   - min/max trackers are uint32 pairs holding order-preserving encodings of float32
 *     (rsx_minmax_init / rsx_minmax_decode); they are updated with atomics so that several
 *     producers and several GPUs' strips can share one tracker.
 */
#ifndef RSX_H
#define RSX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSX_OK 0
#define RSX_ERR_ARG 1      /* bad argument / unsupported configuration */
#define RSX_ERR_CUDA 2     /* a CUDA runtime call or kernel launch failed */
#define RSX_ERR_UNSUPPORTED 3

#define RSX_MAX_BANDS 16
#define RSX_MAX_FEATURES 32
#define RSX_MAX_CLUSTERS 64
#define RSX_MAX_PEERS 8             /* ranks of one node that reduce KMeans sums through peer memory */
#define RSX_PEER_PASS_ELEMS 2560     /* int64 per pass buffer: >= MAX_CLUSTERS * (MAX_FEATURES + 1) + 2 */
#define RSX_PEER_BLOCK_BYTES (2 * RSX_PEER_PASS_ELEMS * 8 + RSX_MAX_PEERS * 8)
#define RSX_NUM_INDICES 7  /* ndvi, evi, msavi, ndwi, mndwi, ndbi, bsi (scripts/2_feature_extraction.py:63-73) */
#define RSX_NUM_GLCM_PROPS 5 /* contrast, dissimilarity, homogeneity, energy, correlation (indices.py:292-296) */

typedef void* rsx_stream_t;

const char* rsx_last_error(void);
int rsx_abi_version(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t rsx_launch_count(void);
/* Integer tuning knobs (kernel-variant switches for the profiling tools and A/B tests; never needed for correct results).
 * A knob that was never set falls back to the environment variable RSX_<NAME IN UPPER CASE>, then to the built-in default. */
int rsx_set_option(const char* name, int value);
int rsx_get_option(const char* name, int dflt); /* the value a call site with this default would see */
/* Copies `bytes` from device memory into PAGE-LOCKED host memory (cudaHostAlloc / torch pin_memory: under unified addressing
 * the device writes through the same pointer) with a kernel, on `stream`; the caller synchronises.  For the small per-scene
 * results (histograms, moments, min/max, KMeans state): unlike cudaMemcpyAsync it does not queue behind a large transfer on
 * the device-to-host copy engine. */
int rsx_store_to_host(const void* d_src, void* h_mapped_dst, int64_t bytes, rsx_stream_t stream);

/* ---- K1: per-band histograms -------------------------------------------------------------
 * Replaces the sorts inside np.percentile (indices.py:38-39) and RobustScaler's
 * nanmedian/nanpercentile (sklearn/preprocessing/_data.py:1722,1738-1743): for integer-valued
 * rasters every order statistic is derivable exactly from the histogram.
 * d_hist: uint32 [B][256] (u8) or [B][65536] (u16); ACCUMULATED into (caller zeroes). */
int rsx_hist_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, uint32_t* d_hist, rsx_stream_t stream);
int rsx_hist_u16(const uint16_t* d_raster, int64_t n_px, int n_bands, uint32_t* d_hist, rsx_stream_t stream);

/* Host-side order statistics from the K1 histograms (no device work): numpy's float32 percentile arithmetic for
 * robust_normalize (indices.py:38-46; also the second normalisation of the texture band, :265) and RobustScaler's
 * median / inter-quartile range (sklearn/preprocessing/_data.py:1722,1738-1743).
 * h_hist int64 [B][L] (L = 256 or 65536); h_norm float [B][3] = lo, hi, den; h_qnorm float [3] for band texture_band
 * (-1: none); h_center float [B], h_scale double [B] (both may be NULL); h_norm_lut / h_x_lut float [B][L] (may be NULL):
 * the normalised value / the RobustScaler-transformed value of every level. */
int rsx_raster_stats(const int64_t* h_hist, int n_bands, int n_levels, int texture_band, double lower, double upper,
                     float* h_norm, float* h_qnorm, float* h_center, double* h_scale, float* h_norm_lut, float* h_x_lut);

/* The same statistics computed ON THE DEVICE for uint8 rasters (L = 256), so that the histograms need not visit the host between
 * two kernels: d_hist = the [B][256] counters of rsx_hist_u8 (uint32; int64 after a multi-rank all-reduce: hist_is_int64 = 1);
 * d_stats = a device block of rsx_raster_stats_device_bytes() bytes holding norm [16][3], qnorm [4], center [16], scale (double)
 * [16] and, at byte offset rsx_raster_stats_device_lut_offset(), x_lut float [16][256] (rows 0..B-1 = the table rsx_pca_*_u8 take).
 * Bit-identical to rsx_raster_stats (same operations, individually rounded).  rsx_indices_fused_u8_dev reads the block. */
int64_t rsx_raster_stats_device_bytes(void);
int64_t rsx_raster_stats_device_lut_offset(void);
int rsx_raster_stats_u8_device(const void* d_hist, int hist_is_int64, int n_bands, int texture_band, double lower, double upper,
                               void* d_stats, rsx_stream_t stream);

/* ---- K2: fused normalise + spectral indices (+ GLCM quantisation) ---------------------------
 * Replaces robust_normalize x B (indices.py:25-48 via scripts/2...:43-47) and the seven index
 * functions (indices.py:50-203) in one pass over the raster.
 * h_norm:      float [B][3] = lo, hi, fl32(fl32(hi-lo)+1e-10f) per band
 * band_map:    raster band numbers of blue, green, red, nir, swir1
 * d_indices:   7 planes in RSX index order; plane_stride >= n_px
 * d_minmax:    uint32 [7][2] tracker or NULL
 * d_quant:     uint8 [n_px] or NULL: (robust_normalize(nir_norm) * (levels-1)).astype(uint8)
 *              (indices.py:265-268); h_qnorm = lo, hi, den of that second normalisation.
 * evi:         L, C1, C2, G of calculate_evi (indices.py:73)
 * h_remap:     (uint8 only, may be NULL) uint8 [B][256]: level v of band b is first replaced by h_remap[b][v].  This is
 *              how stage 1 (gain/bias -> min-max stretch -> uint8, modules/features/preprocessing.py:54-125) is fused
 *              into the load path: for 8-bit input the whole chain is one table per band. */
int rsx_indices_fused_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, const int* band_map, const float* h_norm,
                         const float* evi, float* d_indices, int64_t plane_stride, uint32_t* d_minmax, uint8_t* d_quant,
                         const float* h_qnorm, int levels, const uint8_t* h_remap, rsx_stream_t stream);
/* rsx_indices_fused_u8 with the normalisation parameters (h_norm, h_qnorm) read from the device block of rsx_raster_stats_u8_device */
int rsx_indices_fused_u8_dev(const uint8_t* d_raster, int64_t n_px, int n_bands, const int* band_map, const void* d_stats,
                             const float* evi, float* d_indices, int64_t plane_stride, uint32_t* d_minmax, uint8_t* d_quant, int levels,
                             rsx_stream_t stream);
int rsx_indices_fused_u16(const uint16_t* d_raster, int64_t n_px, int n_bands, const int* band_map, const float* h_norm,
                          const float* evi, float* d_indices, int64_t plane_stride, uint32_t* d_minmax, uint8_t* d_quant,
                          const float* h_qnorm, int levels, rsx_stream_t stream);

/* ---- planar float32 element-wise entry points (the per-function drop-ins) ------------------ */
/* robust_normalize given lo/hi (indices.py:42-46) */
int rsx_normalize_f32(const float* d_in, int64_t n, float lo, float hi, float den, float* d_out, rsx_stream_t stream);
/* (a-b)/(a+b) masked at a+b>0.001, clipped: ndvi/ndwi/mndwi/ndbi (indices.py:50-71,116-179) */
int rsx_index_ratio_f32(const float* d_a, const float* d_b, int64_t n, float* d_out, rsx_stream_t stream);
int rsx_index_evi_f32(const float* d_nir, const float* d_red, const float* d_blue, int64_t n, float L, float C1, float C2,
                      float G, float* d_out, rsx_stream_t stream); /* indices.py:73-95 */
int rsx_index_msavi_f32(const float* d_nir, const float* d_red, int64_t n, float* d_out, rsx_stream_t stream); /* :97-114 */
int rsx_index_bsi_f32(const float* d_blue, const float* d_red, const float* d_nir, const float* d_swir, int64_t n,
                      float* d_out, rsx_stream_t stream); /* :181-203 */
/* quantise a float32 band: (norm(x) * (levels-1)).astype(uint8) (indices.py:265-268) */
int rsx_quantize_f32(const float* d_in, int64_t n, float lo, float hi, float den, int levels, uint8_t* d_q, rsx_stream_t stream);

/* ---- K3: PCA -----------------------------------------------------------------------------------
 * Replaces perform_pca (indices.py:205-246): X = RobustScaler(normalised bands) in float32, with the
 * division by the float64 scale_ evaluated in double and rounded (sklearn/preprocessing/_data.py:
 * 1738-1743,1782-1784); Gram matrix X^T X and column sums (sklearn/decomposition/_pca.py:587-613)
 * accumulated in float64 with warp-shuffle tree reductions and a fixed-order final sum; then the
 * projection X @ components^T - mean @ components^T (sklearn/decomposition/_base.py:151-159).
 * uint8:  d_lut float [B][256] (DEVICE) tabulates X for every grey level of every band.
 * uint16: h_norm float [B][3] as in K2; h_center float [B] (median), h_scale double [B] (IQR, 0 -> 1).
 * d_moments: double [B + B*(B+1)/2]: sums then upper-triangular cross sums, row-major (a<=b);
 *            ACCUMULATED into (caller zeroes); a multi-GPU caller all-reduces it.
 * d_scratch: double [rsx_pca_scratch_elems(B)] */
int64_t rsx_pca_scratch_elems(int n_bands);
int rsx_pca_moments_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, const float* d_lut, double* d_moments,
                       double* d_scratch, rsx_stream_t stream);
/* d_lut16 (may be NULL): float [B][65536] on the device made by rsx_pca_build_lut_u16 - X tabulated per 16-bit level with
 * the per-sample arithmetic, so that a sample costs a look-up instead of a float32 and a float64 division */
int rsx_pca_build_lut_u16(const float* d_norm, const float* d_center, const double* d_scale, int n_bands, float* d_lut,
                          rsx_stream_t stream);   /* d_norm [B][3], d_center [B], d_scale [B]: DEVICE copies of h_norm/h_center/h_scale */
int rsx_pca_moments_u16(const uint16_t* d_raster, int64_t n_px, int n_bands, const float* h_norm, const float* h_center,
                        const double* h_scale, const float* d_lut16, double* d_moments, double* d_scratch, rsx_stream_t stream);
/* h_components: float [n_comp][B]; h_mean_proj: float [n_comp] = mean_ @ components^T */
int rsx_pca_project_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, const float* d_lut, const float* h_components,
                       const float* h_mean_proj, int n_comp, float* d_out, int64_t plane_stride, uint32_t* d_minmax,
                       rsx_stream_t stream);
int rsx_pca_project_u16(const uint16_t* d_raster, int64_t n_px, int n_bands, const float* h_norm, const float* h_center,
                        const double* h_scale, const float* d_lut16, const float* h_components, const float* h_mean_proj, int n_comp,
                        float* d_out, int64_t plane_stride, uint32_t* d_minmax, rsx_stream_t stream);

/* The same two steps for PLANAR float32 bands with arbitrary values and any band count up to RSX_MAX_BANDS (band b at d_bands +
 * b * plane_stride): what the per-function drop-in of perform_pca uses when the bands are not 8-bit levels.
 * robust != 0: X = (float)((double)(v - h_a[b]) / h_scale[b]) with h_a = RobustScaler.center_, h_scale = scale_ (h_den unused);
 * robust == 0: X = (v - h_a[b]) / h_den[b] in float32 with h_a = min, h_den = fl32(max - min + 1e-10) (indices.py:234).
 * d_moments as for the raster variants; d_scratch: double [rsx_pca_planar_scratch_elems(B)]. */
int64_t rsx_pca_planar_scratch_elems(int n_bands);
int rsx_pca_moments_planar_f32(const float* d_bands, int64_t plane_stride, int64_t n_px, int n_bands, int robust, const float* h_a,
                               const float* h_den, const double* h_scale, double* d_moments, double* d_scratch, rsx_stream_t stream);
int rsx_pca_project_planar_f32(const float* d_bands, int64_t plane_stride, int64_t n_px, int n_bands, int robust, const float* h_a,
                               const float* h_den, const double* h_scale, const float* h_components, const float* h_mean_proj, int n_comp,
                               float* d_out, int64_t out_plane_stride, uint32_t* d_minmax, rsx_stream_t stream);

/* ---- K4: GLCM texture --------------------------------------------------------------------------
 * Replaces the window double loop of calculate_glcm_features (indices.py:283-305): for every
 * window anchored at (i*step, j*step), distance 1, angles 0/45/90/135 deg, symmetric, normed;
 * the five graycoprops averaged over the four angles, stored float32.
 * d_q:      uint8 [rows_avail][W] quantised band; rows_avail >= (out_rows-1)*step + window
 *           (a strip plus its halo rows; the windows of output row r start at row r*step).
 * d_props:  5 planes [out_rows][out_cols], plane stride in elements. */
int rsx_glcm_props(const uint8_t* d_q, int rows_avail, int W, int levels, int window, int step, int out_rows, int out_cols,
                   float* d_props, int64_t plane_stride, rsx_stream_t stream);
/* The same for an arbitrary list of pair offsets: h_offsets int32 [n_offsets][2] = (round(sin(angle)*d), round(cos(angle)*d)) for
 * every (distance, angle) the caller of calculate_glcm_features asked for (indices.py:248-249,288-289); the properties are
 * averaged over all of them (.mean() of the graycoprops array).  1 <= n_offsets <= 16; general kernel (one warp per window). */
int rsx_glcm_props_offsets(const uint8_t* d_q, int rows_avail, int W, int levels, int window, int step, int out_rows, int out_cols,
                           const int32_t* h_offsets, int n_offsets, float* d_props, int64_t plane_stride, rsx_stream_t stream);
/* rsx_glcm_props with the same kernels ALSO storing, for every window and angle, the exact integers the five properties are
 * made from (validation of the production kernel's integer stage against graycomatrix counts, bit for bit):
 * d_moments int64 [out_rows*out_cols][4][8] = n (pair instances), sum|a-b|, sum(a+b), sum(a^2+b^2), sum(a*b),
 * E = sum over the cells of the symmetric count matrix P = C + C^T of P^2, pairs with a == b, sum of round(2^h/(1+(a-b)^2))
 * with h = 40 for levels <= 32 and h = 36 above (the fixed-point homogeneity terms of the kernels).
 * contrast = (sq-2sab)/n, dissimilarity = s1/n, homogeneity = hom/2^h/n, energy = sqrt(E)/2n,
 * correlation = (4n*sab - sa^2)/(2n*sq - sa^2) (1 for a constant window). */
int rsx_glcm_moments(const uint8_t* d_q, int rows_avail, int W, int levels, int window, int step, int out_rows, int out_cols,
                     float* d_props, int64_t plane_stride, int64_t* d_moments, rsx_stream_t stream);
/* Directed (un-symmetrised) uint32 co-occurrence counts, [n_win][4][levels][levels], for the
 * windows whose top-left corners are listed in d_anchors (int32 [n_win][2] = row, col):
 * skimage graycomatrix(symmetric=False, normed=False) of each window.  Validation entry point. */
int rsx_glcm_counts(const uint8_t* d_q, int H, int W, int levels, int window, const int32_t* d_anchors, int n_win,
                    uint32_t* d_counts, rsx_stream_t stream);
/* cv2.resize(src, (dst_w, dst_h), INTER_LINEAR) on n_planes float32 planes (indices.py:308).
 * For row-strip sharding: src rows [src_row0, src_row0+src_rows_avail) of a src_h_total-row
 * image are present; dst rows [dst_row0, dst_row0+dst_rows) of a dst_h_total-row image are made. */
int rsx_resize_bilinear_f32(const float* d_src, int src_h_total, int src_w, int src_row0, int src_rows_avail,
                            int64_t src_plane_stride, float* d_dst, int dst_h_total, int dst_w, int dst_row0, int dst_rows,
                            int64_t dst_plane_stride, int n_planes, uint32_t* d_minmax, rsx_stream_t stream);

/* N1: add_spatial_context (indices.py:760-776): cv2.boxFilter(plane, -1, (ksize, ksize), normalize=True,
 * borderType=BORDER_REFLECT) on n_planes float32 planes (double sums, (float)(sum * 1/k^2) like OpenCV).
 * Row-strip sharding as for the resize: source rows [src_row0, src_row0+rows_avail) are present (the strip plus
 * ksize/2 halo rows where the image continues), destination rows [dst_row0, dst_row0+dst_rows) are made. */
int rsx_box_mean_f32(const float* d_src, int H_total, int W, int src_row0, int rows_avail, int64_t src_plane_stride,
                     float* d_dst, int dst_row0, int dst_rows, int64_t dst_plane_stride, int n_planes, int ksize,
                     uint32_t* d_minmax, rsx_stream_t stream);

/* N2: the stencil channels of the level-2 stack (prepare_level_2_features, indices.py:837-865) on the texture band.
 * Same row-strip convention as rsx_box_mean_f32.  rsx_band_lut_*: plane[p] = lut[raster[p*B + band]] (d_lut on the DEVICE,
 * 256 entries) - how the second robust_normalize of the texture band and its 8-bit quantisation are applied. */
int rsx_band_lut_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, int band, const uint8_t* d_lut, uint8_t* d_out, rsx_stream_t stream);
int rsx_band_lut_f32(const uint8_t* d_raster, int64_t n_px, int n_bands, int band, const float* d_lut, float* d_out, rsx_stream_t stream);
/* cv2.morphologyEx(src, MORPH_GRADIENT, ones(ksize, ksize)) on uint8 (indices.py:421-432), ksize 3, 5 or 7 */
int rsx_morph_gradient_u8(const uint8_t* d_src, int H_total, int W, int src_row0, int rows_avail, uint8_t* d_dst, int dst_row0,
                          int dst_rows, int ksize, rsx_stream_t stream);
/* sqrt(max(cv2.blur(x*x) - cv2.blur(x)^2, 0)) in the reference's float32 arithmetic (indices.py:537-548), ksize 3, 5 or 7;
 * d_minmax: one tracker slot or NULL */
int rsx_local_std_f32(const float* d_src, int H_total, int W, int src_row0, int rows_avail, float* d_dst, int dst_row0, int dst_rows,
                      int ksize, uint32_t* d_minmax, rsx_stream_t stream);
/* sqrt((Sobel_x/255)^2 + (Sobel_y/255)^2), 3x3, BORDER_REFLECT_101 (indices.py:477-479); the division by the global maximum
 * (:480) is rsx_divide_f32 with max + 1e-10 once the tracker has been reduced over all strips */
int rsx_sobel_mag_u8(const uint8_t* d_src, int H_total, int W, int src_row0, int rows_avail, float* d_dst, int dst_row0, int dst_rows,
                     uint32_t* d_minmax, rsx_stream_t stream);
int rsx_divide_f32(float* d_plane, int64_t n, float denominator, rsx_stream_t stream);
/* plane[i] = (float)(in[i] / 255.0): uint8 maps enter the float32 stack (the reference keeps them as float64 k/255.0) */
int rsx_u8_over_255_f32(const uint8_t* d_in, int64_t n, float* d_out, rsx_stream_t stream);

/* N4: output layouts.  (H, W, C) float64 C-order = the payload of the reference's .npy files (scripts/2_feature_extraction.py:
 * 193-214) from the planar float32 stack; label image + 1 as uint8 for the GeoTIFF writer (scripts/3_classification.py:394,
 * extract.py:795-807). */
int rsx_planes_to_hwc_f64(const float* d_planes, int64_t plane_stride, int64_t n_px, int n_channels, double* d_out, rsx_stream_t stream);
int rsx_labels_plus1_u8(const int32_t* d_labels, int64_t n, uint8_t* d_out, rsx_stream_t stream);
/* HOST helper (no device work): widens a downloaded uint8 label plane into the caller's int32 array (the dtype the reference
 * returns, extract.py:577) on n_threads host threads.  The device keeps labels as uint8; downloading those moves a quarter of
 * the bytes of an int32 download over PCIe. */
int rsx_widen_u8_to_i32(const uint8_t* h_src, int32_t* h_dst, int64_t n, int n_threads);

/* ---- min/max trackers (MinMaxScaler.fit, sklearn/preprocessing/_data.py:527-541) --------------- */
int rsx_minmax_init(uint32_t* d_minmax, int n, rsx_stream_t stream);
int rsx_minmax_planes_f32(const float* d_planes, int64_t n_px, int64_t plane_stride, int n_planes, uint32_t* d_minmax,
                          rsx_stream_t stream);
void rsx_minmax_decode(const uint32_t* h_minmax, int n, float* h_min, float* h_max);
void rsx_minmax_encode(const float* h_min, const float* h_max, int n, uint32_t* h_minmax);
/* NaN -> 0 in place (extract.py:548-556) */
int rsx_nan_to_zero_f32(float* d_planes, int64_t n, rsx_stream_t stream);
/* The same on ONE plane whose tracker slot was filled by its producer (the producers skip NaN): when a NaN is replaced, 0
 * joins the tracked range, because the reference fits MinMaxScaler after the replacement (extract.py:548-570).  The device
 * pipeline calls it on the MSAVI plane (the only index that can be NaN for finite input: sqrt of a float32 radicand that
 * rounds below zero, indices.py:109-112) before KMeans. */
int rsx_nan_to_zero_minmax_f32(float* d_plane, int64_t n, uint32_t* d_minmax_slot, rsx_stream_t stream);

/* ---- K5: KMeans -------------------------------------------------------------------------------
 * Replaces MinMaxScaler.transform + the Lloyd iteration of sklearn KMeans as driven by
 * unsupervised_kmeans_classification (extract.py:568-577; sklearn/cluster/_k_means_lloyd.pyx:
 * 168-218, _k_means_common.pyx:274-311, _kmeans.py:1488-1493 for the centring).
 *
 * rsx_kmeans_state is a small DEVICE block owned by the caller, rsx_kmeans_state_bytes() long.
 *   rsx_kmeans_setup     fills it: per-feature MinMax scale/min, data mean (centring), the
 *                        fixed-point shifts for the exact int64 partial sums, the initial
 *                        centroids (scaled, un-centred coordinates, double [K][D]).
 *   rsx_kmeans_assign    ONE pass over the stack: argmin over centroids (fp32 fast path with a
 *                        float64 re-evaluation of near ties, first minimum wins), and, fused,
 *                        per-cluster int64 fixed-point sums + counts (update = 1: from scratch;
 *                        update = 2: only the pixels whose label differs from the previous pass
 *                        move their sample between clusters), labels (d_labels_u8 /
 *                        d_labels_i32, each may be NULL), inertia (d_inertia != NULL).
 *   rsx_kmeans_update    totals <- acc (delta = 0) or totals += acc (delta = 1); centroids <-
 *                        totals/counts (multiply by 1/count), centre shift, empty-cluster flag;
 *                        zeroes the pass accumulators.
 * d_acc: int64 [2*(K*D + K + 2)] = the PASS block: sums [K][D], counts [K], near-tie counter, changed-label
 * counter; then the TOTALS block of the same shape: running sums and counts, near ties so far, labels
 * changed in the last pass.  A multi-GPU caller all-reduces the pass block between assign and update
 * (integer adds: bit-identical for any partition); rsx_kmeans_update folds it into the totals and zeroes it. */
int64_t rsx_kmeans_state_bytes(void);
/* h_feat_min/max: per-feature min/max of the raw stack (MinMaxScaler.fit); h_mean_scaled: per-feature
 * mean of the scaled stack (the centring of _kmeans.py:1488-1490 - any value gives the same labels in
 * exact arithmetic, it only sets the origin the fp32 fast path works around);
 * h_init_centroids: double [K][D], MinMax-scaled, un-centred coordinates.
 * Not re-entrant: the state is mirrored in one __constant__ block per process. */
int rsx_kmeans_setup(void* d_state, int D, int K, const double* h_feat_min, const double* h_feat_max,
                     const double* h_mean_scaled, const double* h_init_centroids, int64_t n_px_global, rsx_stream_t stream);
/* The same set-up without a host round trip: the per-feature range is taken from D min/max trackers on the device (d_minmax:
 * uint32 [D][2] as maintained by the feature kernels, see "min/max trackers"), the initial centroids from K raw feature rows on the
 * device (d_init_rows_raw: double [K][D]; scaled by the kernel exactly like MinMaxScaler.transform, X * scale_ + min_).
 * h_mean_scaled may be NULL (0.5 for every feature).  The state is bit-identical to rsx_kmeans_setup's for the same numbers.
 * rsx_kmeans_read_all returns the range that was used next to what rsx_kmeans_read returns. */
int rsx_kmeans_setup_device(void* d_state, int D, int K, const uint32_t* d_minmax, const double* d_init_rows_raw,
                            const double* h_mean_scaled, int64_t n_px_global, rsx_stream_t stream);
/* The caller zeroes d_acc once.  update: 0 = assign only, 1 = full update pass, 2 = delta update pass (needs d_labels_u8 and
 * d_labels_prev_u8, distinct buffers; entries >= RSX_MAX_CLUSTERS in the previous labels mean "no previous label").
 * d_labels_prev_u8 (may be NULL): labels of the previous pass; the number of pixels whose label differs is
 * added to the changed-label counter (sklearn's strict-convergence test, _kmeans.py:723).
 * row_len: length (pixels) of an image row; only steers the traversal order (threads walk down columns
 * so that runs of equal labels stay in registers); any value gives the same result. */
int rsx_kmeans_assign(const float* d_stack, int64_t plane_stride, int64_t n_px, int row_len, const void* d_state,
                      int64_t* d_acc, uint8_t* d_labels_u8, const uint8_t* d_labels_prev_u8, int32_t* d_labels_i32,
                      double* d_inertia, int update, int D, int K, rsx_stream_t stream);
/* A delta update pass (update = 2 above) with Hamerly's bound test in front of it (K <= 8): pixels whose label provably cannot
 * have changed since their distances were last evaluated - their slack (distance to the second nearest centre minus distance to
 * their own) exceeds what the centres have moved since, tracked by rsx_kmeans_update inside the state - are skipped unread; the
 * rest is evaluated exactly like rsx_kmeans_assign does.  Labels, sums, counters: identical to the unbounded pass
 * (sklearn/cluster/_k_means_lloyd.pyx:168-218 semantics; the bound is the one of sklearn's algorithm="elkan", _k_means_elkan.pyx).
 * d_labels_u8: labels of the previous pass, updated IN PLACE.  d_aos: float [n_px][rsx_kmeans_aos_stride(D)] scratch, d_slack:
 * float [n_px rounded up to 4] scratch; both are written by the call with first = 1 (which reads every pixel from the planes) and
 * maintained by the calls with first = 0.  A pass of rsx_kmeans_assign in between invalidates them (start again with first = 1). */
int64_t rsx_kmeans_aos_stride(int D);
int rsx_kmeans_assign_bounded(const float* d_stack, int64_t plane_stride, int64_t n_px, const void* d_state, int64_t* d_acc,
                              uint8_t* d_labels_u8, float* d_aos, float* d_slack, int first, int D, int K, rsx_stream_t stream);
/* 16-bit screening passes (K <= 8).  rsx_kmeans_quantize_u16 writes a uint16 copy of the stack (d_q16: [D] planes q_stride samples
 * apart, q_stride a multiple of 8 and >= n_px rounded up to 8) on the grid rint((x - min_d) * 65535 / range_d) of the state's
 * feature ranges.  rsx_kmeans_assign_q16 is a delta pass (update = 2 of rsx_kmeans_assign: same labels, sums, counters) that reads
 * the copy - half the bytes of a pass - and goes back to the float32 planes for the pixels whose two nearest centroids are closer
 * than the quantisation can resolve (a rigorous bound kept in the state) and for the samples that change cluster.  The copy stays
 * valid as long as the state's feature ranges do (one rsx_kmeans_setup*). */
int rsx_kmeans_quantize_u16(const float* d_stack, int64_t plane_stride, int64_t n_px, const void* d_state, uint16_t* d_q16,
                            int64_t q_stride, int D, rsx_stream_t stream);
int rsx_kmeans_assign_q16(const float* d_stack, int64_t plane_stride, int64_t n_px, const void* d_state, int64_t* d_acc,
                          uint8_t* d_labels_u8, const uint8_t* d_labels_prev_u8, const uint16_t* d_q16, int64_t q_stride, int D, int K,
                          rsx_stream_t stream);
/* d_adjust (may be NULL): int64 [K*D + K] added to the totals for this centroid computation only - the host-assisted
 * empty-cluster relocation of sklearn (_k_means_common.pyx:167-211); the running totals are not modified by it. */
int rsx_kmeans_update(void* d_state, int64_t* d_acc, int delta, int D, const int64_t* d_adjust, rsx_stream_t stream);
/* Multi-GPU (one node, one process per GPU) without a collective: every rank allocates one block of RSX_PEER_BLOCK_BYTES with
 * rsx_peer_alloc, exchanges the 64-byte handle (e.g. torch.distributed.all_gather) and maps the others with rsx_peer_open.
 * The assign pass of update number `seq` (1, 2, ... - the same on every rank, never reused) accumulates into the rank's own
 * block at int64 offset (seq & 1) * RSX_PEER_PASS_ELEMS (pass it as d_acc of rsx_kmeans_assign); rsx_kmeans_update_peers then
 * waits for all ranks at a flag barrier in peer memory, sums their blocks over NVLink into d_acc's pass block and continues
 * as rsx_kmeans_update (d_acc = the local [pass | totals] array).  h_peer_blocks: HOST array of `world` device pointers
 * (own block at [rank]).  A rank that never arrives makes the others time out after 4 s (option peer_timeout_ms; reported by
 * rsx_kmeans_read - the caller must then fail on every rank, pipeline.DeviceKMeans._result all-reduces the failure). */
int rsx_peer_alloc(int64_t bytes, void** d_ptr, uint8_t* h_handle64);
int rsx_peer_open(const uint8_t* h_handle64, void** d_peer);
int rsx_peer_zero(void* d_ptr, int64_t bytes, rsx_stream_t stream);   /* before a new rsx_kmeans_setup: the two pass buffers */
int rsx_peer_close(void* d_peer);
int rsx_peer_free(void* d_ptr);
int rsx_kmeans_update_peers(void* d_state, int64_t* d_acc, int delta, int D, const int64_t* d_adjust, void* const* h_peer_blocks,
                            int rank, int world, int64_t seq, rsx_stream_t stream);
/* ---- k-means++ seeding (sklearn/cluster/_kmeans.py:180-255) on the planar float32 stack, float64 arithmetic, 8 B/sample of
 * state (the running closest squared distance); the host draws the random numbers and drives the rounds.
 * h_scale / h_min: MinMaxScaler scale_, min_; h_mean: per-feature mean of the scaled stack (the centring of KMeans.fit).
 * rsx_kpp_feature_moments: d_out double [2*D] = (sum, sum of squares) of every scaled feature.
 * rsx_kpp_distances: h_cand double [T][D] = candidate centres in centred scaled coordinates.  mode 0: d_closest <- distance to
 *   candidate 0, d_pot[0] = its sum; mode 1: d_pot[t] = sum_i min(d_closest_i, distance_i,t), T <= 8 candidates, d_closest
 *   untouched; mode 2: d_closest <- min(d_closest, distance to candidate 0).  d_pot double [8].
 * rsx_kpp_block_sums: d_sums double [ceil(n / rsx_kpp_block())] = sums of d_closest over blocks of rsx_kpp_block() samples.
 * d_scratch: double [rsx_kpp_scratch_elems()]. */
int64_t rsx_kpp_scratch_elems(void);
int64_t rsx_kpp_block(void);
int rsx_kpp_feature_moments(const float* d_stack, int64_t plane_stride, int64_t n, int D, const double* h_scale, const double* h_min,
                            double* d_out, double* d_scratch, rsx_stream_t stream);
int rsx_kpp_distances(const float* d_stack, int64_t plane_stride, int64_t n, int D, const double* h_scale, const double* h_min,
                      const double* h_mean, const double* h_cand, int T, int mode, double* d_closest, double* d_pot, double* d_scratch,
                      rsx_stream_t stream);
int rsx_kpp_block_sums(const double* d_closest, int64_t n, double* d_sums, rsx_stream_t stream);
/* SYNCHRONISES: the fixed-point scale 2^shift_d per feature (a raw sample enters the sums as rint(x * scale)). */
int rsx_kmeans_fixed_point_scales(const void* d_state, double* h_pow2, rsx_stream_t stream);
/* SYNCHRONISES the stream; centroids come back in scaled, un-centred coordinates, double [K][D];
 * h_shift_sq = squared centre shift of the last update; h_empty = empty clusters met so far. */
int rsx_kmeans_read(const void* d_state, double* h_centroids, double* h_shift_sq, int32_t* h_empty, rsx_stream_t stream);
int rsx_kmeans_read_all(const void* d_state, double* h_centroids, double* h_shift_sq, int32_t* h_empty, double* h_feat_min,
                        double* h_feat_max, rsx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RSX_H */
