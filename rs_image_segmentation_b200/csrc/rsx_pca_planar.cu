// K3 for PLANAR float32 bands with arbitrary values and any band count up to RSX_MAX_BANDS: the per-function drop-in of
// perform_pca (modules/features/indices.py:205-246) when the bands are not 8-bit levels (a caller that hands in bands derived
// from uint16 data, or its own float maps).  Same arithmetic as the raster kernels of rsx_raster_kernels.cu:
//   X = RobustScaler().fit_transform(bands):  X -= center_ (float32), X /= scale_ (float64 divisor: the quotient is evaluated in
//       double and rounded once, sklearn/preprocessing/_data.py:1738-1743,1782-1784)
//   or, use_robust_scaling=False (indices.py:234):  X = (band - min) / (max - min + 1e-10) in float32
//   moments  sum_p X_a, sum_p X_a X_b in float64 (sklearn/decomposition/_pca.py:587-613 needs X^T X and the column sums)
//   project  Y = X @ components^T - mean @ components^T in float32 (sklearn/decomposition/_base.py:151-159)
// HBM layout: bands planar, band b at d_bands + b * plane_stride; 4 B read per sample and pass.
#include "rsx_common.cuh"

struct PlanarScale {
    float a[RSX_MAX_BANDS];       // center_ (robust) or min
    float den[RSX_MAX_BANDS];     // fl32(max - min + 1e-10) (non-robust)
    double scale[RSX_MAX_BANDS];  // scale_ (robust)
    int robust, n_bands;
};

__device__ __forceinline__ float planar_scaled(float v, const PlanarScale& P, int b) {
    const float c = f_sub(v, P.a[b]);
    return P.robust ? __double2float_rn(__ddiv_rn((double)c, P.scale[b])) : f_div(c, P.den[b]);
}

constexpr int PL_THREADS = 256, PL_CHUNK = 256, PL_PITCH = PL_CHUNK + 1;

// One CTA walks chunks of 256 pixels: every thread scales one pixel of the chunk into shared memory (planar, padded pitch), then
// thread m < M adds the chunk's 256 products of ITS moment (a sum or a pair (a, b)) to a private float64 accumulator in pixel
// order - a fixed summation order per CTA.  Partial moments per CTA go to scratch; pca_planar_finish_kernel adds them in CTA order.
__global__ void __launch_bounds__(PL_THREADS) pca_planar_moments_kernel(const float* __restrict__ bands, int64_t plane_stride, int64_t n_px,
                                                                         const __grid_constant__ PlanarScale P, double* __restrict__ scratch) {
    __shared__ float xs[RSX_MAX_BANDS * PL_PITCH];
    const int B = P.n_bands, M = B + B * (B + 1) / 2;
    int ma = -1, mb = -1;  // this thread's moment: m < B: sum of band m; else pair (ma, mb)
    if ((int)threadIdx.x < B) {
        ma = threadIdx.x;
    } else if ((int)threadIdx.x < M) {
        int idx = threadIdx.x - B;
        for (int a = 0; a < B; ++a) {
            if (idx < B - a) {
                ma = a, mb = a + idx;
                break;
            }
            idx -= B - a;
        }
    }
    double acc = 0.0;
    const int64_t n_chunks = (n_px + PL_CHUNK - 1) / PL_CHUNK;
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const int64_t p = ch * PL_CHUNK + threadIdx.x;
        for (int b = 0; b < B; ++b) xs[b * PL_PITCH + threadIdx.x] = p < n_px ? planar_scaled(__ldg(bands + b * plane_stride + p), P, b) : 0.f;
        __syncthreads();
        if (ma >= 0) {
            const float* xa = xs + ma * PL_PITCH;
            if (mb < 0) {
#pragma unroll 8
                for (int i = 0; i < PL_CHUNK; ++i) acc += (double)xa[i];
            } else {
                const float* xb = xs + mb * PL_PITCH;
#pragma unroll 8
                for (int i = 0; i < PL_CHUNK; ++i) acc = fma((double)xa[i], (double)xb[i], acc);
            }
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < M) scratch[(size_t)blockIdx.x * M + threadIdx.x] = acc;
}

__global__ void pca_planar_finish_kernel(const double* __restrict__ scratch, int n_blocks, int M, double* __restrict__ moments) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double s = 0.0;
    for (int i = 0; i < n_blocks; ++i) s += scratch[(size_t)i * M + m];
    moments[m] += s;
}

static int planar_grid(int64_t n_px) { return (int)max((int64_t)1, min((int64_t)rsx_num_sms() * 4, ceil_div(n_px, (int64_t)PL_CHUNK))); }

static int fill_planar(PlanarScale& P, int n_bands, int robust, const float* h_a, const float* h_den, const double* h_scale, const char* who) {
    RSX_REQUIRE(n_bands >= 1 && n_bands <= RSX_MAX_BANDS, "%s: 1..%d bands", who, RSX_MAX_BANDS);
    RSX_REQUIRE(h_a && (robust ? h_scale != nullptr : h_den != nullptr), "%s: missing scaling parameters", who);
    memset(&P, 0, sizeof(P));
    P.robust = robust ? 1 : 0, P.n_bands = n_bands;
    for (int b = 0; b < n_bands; ++b) {
        P.a[b] = h_a[b];
        P.den[b] = h_den ? h_den[b] : 1.f;
        P.scale[b] = h_scale ? h_scale[b] : 1.0;
    }
    return RSX_OK;
}

extern "C" int64_t rsx_pca_planar_scratch_elems(int n_bands) { return (int64_t)rsx_num_sms() * 4 * (n_bands + n_bands * (n_bands + 1) / 2); }

extern "C" int rsx_pca_moments_planar_f32(const float* d_bands, int64_t plane_stride, int64_t n_px, int n_bands, int robust, const float* h_a,
                                          const float* h_den, const double* h_scale, double* d_moments, double* d_scratch, rsx_stream_t stream) {
    RSX_REQUIRE(d_bands && d_moments && d_scratch && n_px > 0 && plane_stride >= n_px, "rsx_pca_moments_planar_f32: bad arguments");
    PlanarScale P;
    if (int rc = fill_planar(P, n_bands, robust, h_a, h_den, h_scale, "rsx_pca_moments_planar_f32")) return rc;
    const int M = n_bands + n_bands * (n_bands + 1) / 2;
    const int grid = planar_grid(n_px);
    cudaStream_t s = (cudaStream_t)stream;
    pca_planar_moments_kernel<<<grid, PL_THREADS, 0, s>>>(d_bands, plane_stride, n_px, P, d_scratch);
    if (int rc = rsx_check_launch("pca_planar_moments")) return rc;
    pca_planar_finish_kernel<<<ceil_div(M, 128), 128, 0, s>>>(d_scratch, grid, M, d_moments);
    return rsx_check_launch("pca_planar_finish");
}

struct PlanarProj {
    PlanarScale sc;
    float comp[RSX_MAX_BANDS][RSX_MAX_BANDS];  // [component][band]
    float mean_proj[RSX_MAX_BANDS];
    int n_comp;
};

__global__ void __launch_bounds__(256) pca_planar_project_kernel(const float* __restrict__ bands, int64_t plane_stride, int64_t n_px,
                                                                  const __grid_constant__ PlanarProj P, float* __restrict__ out, int64_t out_stride,
                                                                  uint32_t* __restrict__ minmax) {
    const int B = P.sc.n_bands;
    float mn[RSX_MAX_BANDS], mx[RSX_MAX_BANDS];
#pragma unroll
    for (int c = 0; c < RSX_MAX_BANDS; ++c) mn[c] = INFINITY, mx[c] = -INFINITY;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += (int64_t)gridDim.x * blockDim.x) {
        float x[RSX_MAX_BANDS];
#pragma unroll
        for (int b = 0; b < RSX_MAX_BANDS; ++b) x[b] = b < B ? planar_scaled(__ldg(bands + b * plane_stride + p), P.sc, b) : 0.f;
#pragma unroll
        for (int c = 0; c < RSX_MAX_BANDS; ++c) {
            if (c < P.n_comp) {
                float a = 0.f;
#pragma unroll
                for (int b = 0; b < RSX_MAX_BANDS; ++b)
                    if (b < B) a = fmaf(x[b], P.comp[c][b], a);
                const float y = f_sub(a, P.mean_proj[c]);
                out[c * out_stride + p] = y;
                mn[c] = fminf(mn[c], y), mx[c] = fmaxf(mx[c], y);
            }
        }
    }
    if (minmax) {
#pragma unroll
        for (int c = 0; c < RSX_MAX_BANDS; ++c)
            if (c < P.n_comp) warp_minmax_commit(mn[c], mx[c], minmax + 2 * c);
    }
}

extern "C" int rsx_pca_project_planar_f32(const float* d_bands, int64_t plane_stride, int64_t n_px, int n_bands, int robust, const float* h_a,
                                          const float* h_den, const double* h_scale, const float* h_components, const float* h_mean_proj, int n_comp,
                                          float* d_out, int64_t out_plane_stride, uint32_t* d_minmax, rsx_stream_t stream) {
    RSX_REQUIRE(d_bands && d_out && h_components && h_mean_proj && n_px > 0 && plane_stride >= n_px && out_plane_stride >= n_px,
                "rsx_pca_project_planar_f32: bad arguments");
    RSX_REQUIRE(n_comp >= 1 && n_comp <= n_bands, "rsx_pca_project_planar_f32: n_comp must be in [1, n_bands]");
    PlanarProj P;
    memset(&P, 0, sizeof(P));
    if (int rc = fill_planar(P.sc, n_bands, robust, h_a, h_den, h_scale, "rsx_pca_project_planar_f32")) return rc;
    for (int c = 0; c < n_comp; ++c) {
        for (int b = 0; b < n_bands; ++b) P.comp[c][b] = h_components[c * n_bands + b];
        P.mean_proj[c] = h_mean_proj[c];
    }
    P.n_comp = n_comp;
    const int grid = (int)max((int64_t)1, min((int64_t)rsx_num_sms() * 8, ceil_div(n_px, (int64_t)256)));
    pca_planar_project_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_bands, plane_stride, n_px, P, d_out, out_plane_stride, d_minmax);
    return rsx_check_launch("pca_planar_project");
}
