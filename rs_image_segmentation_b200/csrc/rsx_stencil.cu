// N2: the three level-2 channels of the reference's hierarchical stack that are small stencils on the texture band
// (prepare_level_2_features, modules/features/indices.py:837-865):
//   gradient_5       cv2.morphologyEx(band_uint8, MORPH_GRADIENT, ones(5,5)) / 255.0        (indices.py:421-440)
//   std_dev_scale_5  sqrt(max(blur(x*x, 5) - blur(x, 5)^2, 0)) on the float32 band           (indices.py:537-548)
//   sobel_mag        sqrt(Sx^2 + Sy^2) / (max + 1e-10), Sx, Sy = cv2.Sobel(band_uint8, CV_32F) / 255   (indices.py:477-480)
// plus the table look-ups that produce their inputs from the interleaved raster.
// All kernels follow the row-strip convention of the resize / box filter: source rows [src_row0, src_row0 + rows_avail) of an
// H_total-row image are present, destination rows [dst_row0, dst_row0 + dst_rows) are produced.
#include "rsx_common.cuh"

__device__ __forceinline__ int reflect101(int i, int n) {  // BORDER_REFLECT_101 (cv2 BORDER_DEFAULT): gfedcb|abcdefgh|gfedcba
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

// ----------------------------------------------------------------------------- band extraction through a table
template <typename OUT>
__global__ void __launch_bounds__(256) band_lut_kernel(const uint8_t* __restrict__ raster, int64_t n_px, int B, int band, const OUT* __restrict__ lut,
                                                       OUT* __restrict__ out) {
    __shared__ OUT t[256];
    t[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_px; p += (int64_t)gridDim.x * blockDim.x) out[p] = t[raster[p * B + band]];
}

extern "C" int rsx_band_lut_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, int band, const uint8_t* d_lut, uint8_t* d_out, rsx_stream_t stream) {
    RSX_REQUIRE(d_raster && d_lut && d_out && n_px > 0 && band >= 0 && band < n_bands, "rsx_band_lut_u8: bad arguments");
    band_lut_kernel<uint8_t><<<(int)min((int64_t)rsx_num_sms() * 8, ceil_div(n_px, (int64_t)256)), 256, 0, (cudaStream_t)stream>>>(d_raster, n_px, n_bands, band,
                                                                                                                                   d_lut, d_out);
    return rsx_check_launch("band_lut_u8");
}
extern "C" int rsx_band_lut_f32(const uint8_t* d_raster, int64_t n_px, int n_bands, int band, const float* d_lut, float* d_out, rsx_stream_t stream) {
    RSX_REQUIRE(d_raster && d_lut && d_out && n_px > 0 && band >= 0 && band < n_bands, "rsx_band_lut_f32: bad arguments");
    band_lut_kernel<float><<<(int)min((int64_t)rsx_num_sms() * 8, ceil_div(n_px, (int64_t)256)), 256, 0, (cudaStream_t)stream>>>(d_raster, n_px, n_bands, band,
                                                                                                                                 d_lut, d_out);
    return rsx_check_launch("band_lut_f32");
}

// ----------------------------------------------------------------------------- column-walking stencil frame
// 128 threads = 128 columns; the CTA streams image rows through a shared row buffer; every thread keeps the last KS
// horizontal results of its column in registers.
struct Strip {
    int H_total, W, src_row0, rows_avail, dst_row0, dst_rows, rows_per_cta;
};

// morphological gradient: max - min over the KS x KS window, pixels outside the image ignored (cv2's default border value
// for erode/dilate)
template <int KS>
__global__ void __launch_bounds__(128) morph_gradient_kernel(const uint8_t* __restrict__ src, Strip g, uint8_t* __restrict__ dst) {
    constexpr int R = KS / 2;
    __shared__ uint8_t row[128 + 2 * R];
    const int x0 = blockIdx.x * 128, x = x0 + threadIdx.x;
    const int ly0 = blockIdx.y * g.rows_per_cta, ly1 = min(g.dst_rows, ly0 + g.rows_per_cta);
    uint8_t rmin[KS], rmax[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) rmin[k] = 255, rmax[k] = 0;
    for (int gy = g.dst_row0 + ly0 - R; gy < g.dst_row0 + ly1 + R; ++gy) {
        const bool row_in = gy >= 0 && gy < g.H_total;
        const int sy = gy - g.src_row0;
        __syncthreads();
        for (int i = threadIdx.x; i < 128 + 2 * R; i += 128) {
            const int sx = x0 - R + i;
            row[i] = (row_in && sx >= 0 && sx < g.W && sy >= 0 && sy < g.rows_avail) ? src[(int64_t)sy * g.W + sx] : (uint8_t)0;
        }
        __syncthreads();
        int hmin = 255, hmax = 0;
        if (row_in) {
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const int sx = x - R + k;
                if (sx >= 0 && sx < g.W) hmin = min(hmin, (int)row[threadIdx.x + k]), hmax = max(hmax, (int)row[threadIdx.x + k]);
            }
        }
#pragma unroll
        for (int k = 0; k < KS - 1; ++k) rmin[k] = rmin[k + 1], rmax[k] = rmax[k + 1];
        rmin[KS - 1] = (uint8_t)hmin, rmax[KS - 1] = (uint8_t)hmax;
        const int oy = gy - R;
        if (oy >= g.dst_row0 + ly0 && x < g.W) {
            int lo = 255, hi = 0;
#pragma unroll
            for (int k = 0; k < KS; ++k) lo = min(lo, (int)rmin[k]), hi = max(hi, (int)rmax[k]);
            dst[(int64_t)(oy - g.dst_row0) * g.W + x] = (uint8_t)(hi - lo);
        }
    }
}

// local standard deviation: cv2.blur (double sums, (float)(sum / KS^2), BORDER_REFLECT_101) of x and of fl32(x*x), then the
// reference's float32 expression sqrt(max(mean_sq - mean*mean, 0))
template <int KS>
__global__ void __launch_bounds__(128) local_std_kernel(const float* __restrict__ src, Strip g, float* __restrict__ dst, uint32_t* __restrict__ minmax) {
    constexpr int R = KS / 2;
    __shared__ float row[128 + 2 * R];
    const int x0 = blockIdx.x * 128, x = x0 + threadIdx.x;
    const int ly0 = blockIdx.y * g.rows_per_cta, ly1 = min(g.dst_rows, ly0 + g.rows_per_cta);
    double r1[KS], r2[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) r1[k] = r2[k] = 0.0;
    const double scale = 1.0 / (double)(KS * KS);
    float mn = INFINITY, mx = -INFINITY;
    for (int gy = g.dst_row0 + ly0 - R; gy < g.dst_row0 + ly1 + R; ++gy) {
        const int sy = reflect101(gy, g.H_total) - g.src_row0;
        __syncthreads();
        for (int i = threadIdx.x; i < 128 + 2 * R; i += 128) {
            const int sx = reflect101(x0 - R + i, g.W);
            row[i] = (sy >= 0 && sy < g.rows_avail) ? src[(int64_t)sy * g.W + sx] : 0.f;
        }
        __syncthreads();
        double h1 = 0.0, h2 = 0.0;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const float v = row[threadIdx.x + k];
            h1 += (double)v;
            h2 += (double)f_mul(v, v);  // band * band is a float32 array in the reference
        }
#pragma unroll
        for (int k = 0; k < KS - 1; ++k) r1[k] = r1[k + 1], r2[k] = r2[k + 1];
        r1[KS - 1] = h1, r2[KS - 1] = h2;
        const int oy = gy - R;
        if (oy >= g.dst_row0 + ly0 && x < g.W) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int k = 0; k < KS; ++k) s1 += r1[k], s2 += r2[k];
            const float mean = (float)(s1 * scale), mean_sq = (float)(s2 * scale);
            float var = f_sub(mean_sq, f_mul(mean, mean));
            if (var < 0.f) var = 0.f;
            const float sd = f_sqrt(var);
            dst[(int64_t)(oy - g.dst_row0) * g.W + x] = sd;
            mn = fminf(mn, sd), mx = fmaxf(mx, sd);
        }
    }
    if (minmax) warp_minmax_commit(mn, mx, minmax);
}

// Sobel magnitude (before the division by its maximum): Sx, Sy exact integers in float32, divided by 255 in float32
__global__ void __launch_bounds__(128) sobel_mag_kernel(const uint8_t* __restrict__ src, Strip g, float* __restrict__ dst, uint32_t* __restrict__ minmax) {
    __shared__ uint8_t row[128 + 2];
    const int x0 = blockIdx.x * 128, x = x0 + threadIdx.x;
    const int ly0 = blockIdx.y * g.rows_per_cta, ly1 = min(g.dst_rows, ly0 + g.rows_per_cta);
    int d[3] = {0, 0, 0}, s[3] = {0, 0, 0};  // per row: horizontal difference (right - left) and smooth (l + 2c + r)
    float mn = INFINITY, mx = -INFINITY;
    for (int gy = g.dst_row0 + ly0 - 1; gy < g.dst_row0 + ly1 + 1; ++gy) {
        const int sy = reflect101(gy, g.H_total) - g.src_row0;
        __syncthreads();
        for (int i = threadIdx.x; i < 128 + 2; i += 128) {
            const int sx = reflect101(x0 - 1 + i, g.W);
            row[i] = (sy >= 0 && sy < g.rows_avail) ? src[(int64_t)sy * g.W + sx] : (uint8_t)0;
        }
        __syncthreads();
        const int l = row[threadIdx.x], c = row[threadIdx.x + 1], r = row[threadIdx.x + 2];
        d[0] = d[1], d[1] = d[2], d[2] = r - l;
        s[0] = s[1], s[1] = s[2], s[2] = l + 2 * c + r;
        const int oy = gy - 1;
        if (oy >= g.dst_row0 + ly0 && x < g.W) {
            const float sx_ = f_div((float)(d[0] + 2 * d[1] + d[2]), 255.f);
            const float sy_ = f_div((float)(s[2] - s[0]), 255.f);
            const float m = f_sqrt(f_add(f_mul(sx_, sx_), f_mul(sy_, sy_)));
            dst[(int64_t)(oy - g.dst_row0) * g.W + x] = m;
            mn = fminf(mn, m), mx = fmaxf(mx, m);
        }
    }
    if (minmax) warp_minmax_commit(mn, mx, minmax);
}

__global__ void __launch_bounds__(256) divide_kernel(float* __restrict__ p, int64_t n, float den) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = f_div(p[i], den);
}

__global__ void __launch_bounds__(256) u8_over_255_kernel(const uint8_t* __restrict__ in, int64_t n, float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)((double)in[i] / 255.0);  // the reference value is float64 k/255.0; rounded once to the stack's float32
}

static int strip_setup(Strip& g, dim3& grid, int H_total, int W, int src_row0, int rows_avail, int dst_row0, int dst_rows, int halo, const char* what) {
    RSX_REQUIRE(H_total >= 1 && W >= 1 && dst_rows >= 1 && rows_avail >= 1, "%s: bad geometry", what);
    RSX_REQUIRE(max(dst_row0 - halo, 0) >= src_row0 && min(dst_row0 + dst_rows - 1 + halo, H_total - 1) < src_row0 + rows_avail, "%s: halo rows missing", what);
    const int gx = ceil_div(W, 128);
    g.H_total = H_total, g.W = W, g.src_row0 = src_row0, g.rows_avail = rows_avail, g.dst_row0 = dst_row0, g.dst_rows = dst_rows;
    g.rows_per_cta = max(32, ceil_div(dst_rows, max(1, rsx_num_sms() * 8 / max(1, gx))));
    grid = dim3(gx, ceil_div(dst_rows, g.rows_per_cta), 1);
    return RSX_OK;
}

extern "C" int rsx_morph_gradient_u8(const uint8_t* d_src, int H_total, int W, int src_row0, int rows_avail, uint8_t* d_dst, int dst_row0, int dst_rows,
                                     int ksize, rsx_stream_t stream) {
    RSX_REQUIRE(d_src && d_dst && (ksize == 3 || ksize == 5 || ksize == 7), "rsx_morph_gradient_u8: kernel size must be 3, 5 or 7");
    Strip g;
    dim3 grid;
    if (int rc = strip_setup(g, grid, H_total, W, src_row0, rows_avail, dst_row0, dst_rows, ksize / 2, "rsx_morph_gradient_u8")) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (ksize == 3) morph_gradient_kernel<3><<<grid, 128, 0, s>>>(d_src, g, d_dst);
    if (ksize == 5) morph_gradient_kernel<5><<<grid, 128, 0, s>>>(d_src, g, d_dst);
    if (ksize == 7) morph_gradient_kernel<7><<<grid, 128, 0, s>>>(d_src, g, d_dst);
    return rsx_check_launch("morph_gradient");
}

extern "C" int rsx_local_std_f32(const float* d_src, int H_total, int W, int src_row0, int rows_avail, float* d_dst, int dst_row0, int dst_rows, int ksize,
                                 uint32_t* d_minmax, rsx_stream_t stream) {
    RSX_REQUIRE(d_src && d_dst && (ksize == 3 || ksize == 5 || ksize == 7), "rsx_local_std_f32: window must be 3, 5 or 7");
    RSX_REQUIRE(H_total > ksize / 2 && W > ksize / 2, "rsx_local_std_f32: window larger than the image");
    Strip g;
    dim3 grid;
    if (int rc = strip_setup(g, grid, H_total, W, src_row0, rows_avail, dst_row0, dst_rows, ksize / 2, "rsx_local_std_f32")) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (ksize == 3) local_std_kernel<3><<<grid, 128, 0, s>>>(d_src, g, d_dst, d_minmax);
    if (ksize == 5) local_std_kernel<5><<<grid, 128, 0, s>>>(d_src, g, d_dst, d_minmax);
    if (ksize == 7) local_std_kernel<7><<<grid, 128, 0, s>>>(d_src, g, d_dst, d_minmax);
    return rsx_check_launch("local_std");
}

extern "C" int rsx_sobel_mag_u8(const uint8_t* d_src, int H_total, int W, int src_row0, int rows_avail, float* d_dst, int dst_row0, int dst_rows,
                                uint32_t* d_minmax, rsx_stream_t stream) {
    RSX_REQUIRE(d_src && d_dst, "rsx_sobel_mag_u8: null argument");
    Strip g;
    dim3 grid;
    if (int rc = strip_setup(g, grid, H_total, W, src_row0, rows_avail, dst_row0, dst_rows, 1, "rsx_sobel_mag_u8")) return rc;
    sobel_mag_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_src, g, d_dst, d_minmax);
    return rsx_check_launch("sobel_mag");
}

extern "C" int rsx_divide_f32(float* d_plane, int64_t n, float denominator, rsx_stream_t stream) {
    RSX_REQUIRE(d_plane && n > 0, "rsx_divide_f32: bad arguments");
    divide_kernel<<<(int)min((int64_t)rsx_num_sms() * 8, ceil_div(n, (int64_t)256)), 256, 0, (cudaStream_t)stream>>>(d_plane, n, denominator);
    return rsx_check_launch("divide");
}

extern "C" int rsx_u8_over_255_f32(const uint8_t* d_in, int64_t n, float* d_out, rsx_stream_t stream) {
    RSX_REQUIRE(d_in && d_out && n > 0, "rsx_u8_over_255_f32: bad arguments");
    u8_over_255_kernel<<<(int)min((int64_t)rsx_num_sms() * 8, ceil_div(n, (int64_t)256)), 256, 0, (cudaStream_t)stream>>>(d_in, n, d_out);
    return rsx_check_launch("u8_over_255");
}

// ----------------------------------------------------------------------------- N4: output layouts of the reference's files
// all_hierarchical_features.npy / level{1,2}_features.npy are (H, W, C) float64 C-order (scripts/2_feature_extraction.py:
// 193-214): planar float32 planes -> pixel-interleaved float64, i.e. the exact payload bytes of the .npy file.
// A CTA transposes 32 pixels x C channels through shared memory so that both sides are coalesced.
__global__ void __launch_bounds__(256) planes_to_hwc_f64_kernel(const float* __restrict__ planes, int64_t plane_stride, int64_t n_px, int C,
                                                                double* __restrict__ out) {
    extern __shared__ float tile[];  // [C][256 + 1]
    const int64_t tiles = (n_px + 255) / 256;
    for (int64_t tb = blockIdx.x; tb < tiles; tb += gridDim.x) {
        const int64_t p0 = tb * 256;
        const int np = (int)min((int64_t)256, n_px - p0);
        for (int c = 0; c < C; ++c)
            if ((int)threadIdx.x < np) tile[c * 257 + threadIdx.x] = planes[c * plane_stride + p0 + threadIdx.x];
        __syncthreads();
        for (int i = threadIdx.x; i < np * C; i += 256) {
            const int p = i / C, c = i - p * C;
            out[p0 * C + i] = (double)tile[c * 257 + p];
        }
        __syncthreads();
    }
}

extern "C" int rsx_planes_to_hwc_f64(const float* d_planes, int64_t plane_stride, int64_t n_px, int n_channels, double* d_out, rsx_stream_t stream) {
    RSX_REQUIRE(d_planes && d_out && n_px > 0 && n_channels >= 1 && n_channels <= 64, "rsx_planes_to_hwc_f64: bad arguments");
    const size_t smem = (size_t)n_channels * 257 * 4;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(planes_to_hwc_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            rsx_set_error("rsx_planes_to_hwc_f64: %s", cudaGetErrorString(e));
            return RSX_ERR_CUDA;
        }
        configured = smem;
    }
    const int grid = (int)min((int64_t)rsx_num_sms() * 4, ceil_div(n_px, (int64_t)256));
    planes_to_hwc_f64_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(d_planes, plane_stride, n_px, n_channels, d_out);
    return rsx_check_launch("planes_to_hwc_f64");
}

// final_classification_map = kmeans_result + 1 (scripts/3_classification.py:394) cast to uint8 for the label GeoTIFF
// (extract.py:795-807, nodata 0)
__global__ void __launch_bounds__(256) labels_plus1_u8_kernel(const int32_t* __restrict__ lab, int64_t n, uint8_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (uint8_t)(lab[i] + 1);
}

extern "C" int rsx_labels_plus1_u8(const int32_t* d_labels, int64_t n, uint8_t* d_out, rsx_stream_t stream) {
    RSX_REQUIRE(d_labels && d_out && n > 0, "rsx_labels_plus1_u8: bad arguments");
    labels_plus1_u8_kernel<<<(int)min((int64_t)rsx_num_sms() * 8, ceil_div(n, (int64_t)256)), 256, 0, (cudaStream_t)stream>>>(d_labels, n, d_out);
    return rsx_check_launch("labels_plus1_u8");
}
