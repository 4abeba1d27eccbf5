// k-means++ seeding (sklearn/cluster/_kmeans.py:180-255, _kmeans_plusplus) on the planar float32 stack, in float64, without a
// float64 copy of the (N, D) matrix and without the (trials, N) distance matrix: the only per-sample state is the running
// closest squared distance (8 B / sample).  The host draws the random numbers (numpy RandomState, the reference's stream) and
// drives the rounds; the kernels do the O(N) work:
//   rsx_kpp_feature_moments   per-feature sum and sum of squares of the MinMax-scaled stack (mean for the centring of
//                             _kmeans.py:1488-1490, variance for the tolerance of _kmeans.py:285-294)
//   rsx_kpp_distances         one pass per round: squared distances of every sample to T candidate centres (the formula of
//                             sklearn's euclidean_distances: |x|^2 + |y|^2 - 2 x.y, clipped at 0), potentials sum_i min(closest_i,
//                             d_it) per candidate (fixed-order block sums), or the commit of the chosen candidate into closest
//   rsx_kpp_block_sums        sums of closest over blocks of 8192 samples: the host finds the block a random threshold falls
//                             into from their running sum and finishes the search on that block's 8192 values
// HBM traffic per round: 4 D B/sample for the stack + 8 (evaluate) or 16 (commit) for closest.
#include "rsx_common.cuh"

struct KppScale {
    double scale[RSX_MAX_FEATURES], min_[RSX_MAX_FEATURES], mean[RSX_MAX_FEATURES];
    int D;
};
constexpr int KPP_MAXT = 8;
struct KppCand {
    double c[KPP_MAXT][RSX_MAX_FEATURES];  // centred, scaled coordinates
    double norm[KPP_MAXT];
    int T;
};

__device__ __forceinline__ double block_sum_fixed(double v, double* red) {  // all threads; fixed order: warp tree, then warps in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(256) kpp_moments_kernel(const float* __restrict__ stack, int64_t stride, int64_t n, const __grid_constant__ KppScale P,
                                                          double* __restrict__ scratch) {
    __shared__ double red[8];
    for (int d = 0; d < P.D; ++d) {
        double s1 = 0.0, s2 = 0.0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const double x = __dadd_rn(__dmul_rn((double)__ldg(stack + d * stride + i), P.scale[d]), P.min_[d]);  // MinMaxScaler.transform
            s1 += x, s2 = fma(x, x, s2);
        }
        s1 = block_sum_fixed(s1, red), s2 = block_sum_fixed(s2, red);
        if (threadIdx.x == 0) scratch[((size_t)blockIdx.x * P.D + d) * 2] = s1, scratch[((size_t)blockIdx.x * P.D + d) * 2 + 1] = s2;
    }
}

__global__ void kpp_finish_kernel(const double* __restrict__ scratch, int n_blocks, int M, double* __restrict__ out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double s = 0.0;
    for (int b = 0; b < n_blocks; ++b) s += scratch[(size_t)b * M + m];
    out[m] = s;
}

// mode 0: closest = d(., cand 0), pot[0]; mode 1: pot[t] = sum min(closest, d(., cand t)) for t < T; mode 2: closest = min(closest, d(., cand 0))
__global__ void __launch_bounds__(256) kpp_distances_kernel(const float* __restrict__ stack, int64_t stride, int64_t n, const __grid_constant__ KppScale P,
                                                            const __grid_constant__ KppCand Cn, int mode, double* __restrict__ closest,
                                                            double* __restrict__ scratch) {
    __shared__ double red[8];
    double pot[KPP_MAXT];
#pragma unroll
    for (int t = 0; t < KPP_MAXT; ++t) pot[t] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double xx = 0.0, dot[KPP_MAXT];
#pragma unroll
        for (int t = 0; t < KPP_MAXT; ++t) dot[t] = 0.0;
        for (int d = 0; d < P.D; ++d) {
            const double x = __dsub_rn(__dadd_rn(__dmul_rn((double)__ldg(stack + d * stride + i), P.scale[d]), P.min_[d]), P.mean[d]);
            xx = fma(x, x, xx);
#pragma unroll
            for (int t = 0; t < KPP_MAXT; ++t)
                if (t < Cn.T) dot[t] = fma(x, Cn.c[t][d], dot[t]);
        }
        const double cl = mode == 0 ? INFINITY : closest[i];
#pragma unroll
        for (int t = 0; t < KPP_MAXT; ++t) {
            if (t < Cn.T) {
                const double dist = fmax(fma(-2.0, dot[t], xx + Cn.norm[t]), 0.0);
                const double m = fmin(cl, dist);
                pot[t] += m;
                if (t == 0 && mode != 1) closest[i] = m;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < KPP_MAXT; ++t) {
        if (t < Cn.T) {  // block-uniform
            const double s = block_sum_fixed(pot[t], red);
            if (threadIdx.x == 0) scratch[(size_t)blockIdx.x * KPP_MAXT + t] = s;
        } else if (threadIdx.x == 0) {
            scratch[(size_t)blockIdx.x * KPP_MAXT + t] = 0.0;
        }
    }
}

constexpr int KPP_BLOCK = 8192;
__global__ void __launch_bounds__(256) kpp_block_sums_kernel(const double* __restrict__ closest, int64_t n, double* __restrict__ sums) {
    __shared__ double red[8];
    const int64_t base = (int64_t)blockIdx.x * KPP_BLOCK;
    double s = 0.0;
    for (int k = threadIdx.x; k < KPP_BLOCK; k += 256) s += base + k < n ? closest[base + k] : 0.0;
    s = block_sum_fixed(s, red);
    if (threadIdx.x == 0) sums[blockIdx.x] = s;
}

static int kpp_grid(int64_t n) { return (int)max((int64_t)1, min((int64_t)rsx_num_sms() * 8, ceil_div(n, (int64_t)256))); }
static int fill_scale(KppScale& P, int D, const double* s, const double* m, const double* mu, const char* who) {
    RSX_REQUIRE(D >= 1 && D <= RSX_MAX_FEATURES && s && m, "%s: bad scaling arguments", who);
    memset(&P, 0, sizeof(P));
    P.D = D;
    for (int d = 0; d < D; ++d) P.scale[d] = s[d], P.min_[d] = m[d], P.mean[d] = mu ? mu[d] : 0.0;
    return RSX_OK;
}

extern "C" int64_t rsx_kpp_scratch_elems(void) { return (int64_t)rsx_num_sms() * 8 * 2 * RSX_MAX_FEATURES; }
extern "C" int64_t rsx_kpp_block(void) { return KPP_BLOCK; }

extern "C" int rsx_kpp_feature_moments(const float* d_stack, int64_t plane_stride, int64_t n, int D, const double* h_scale, const double* h_min,
                                       double* d_out, double* d_scratch, rsx_stream_t stream) {
    RSX_REQUIRE(d_stack && d_out && d_scratch && n > 0, "rsx_kpp_feature_moments: bad arguments");
    KppScale P;
    if (int rc = fill_scale(P, D, h_scale, h_min, nullptr, "rsx_kpp_feature_moments")) return rc;
    const int grid = kpp_grid(n);
    cudaStream_t s = (cudaStream_t)stream;
    kpp_moments_kernel<<<grid, 256, 0, s>>>(d_stack, plane_stride, n, P, d_scratch);
    if (int rc = rsx_check_launch("kpp_moments")) return rc;
    kpp_finish_kernel<<<ceil_div(2 * D, 64), 64, 0, s>>>(d_scratch, grid, 2 * D, d_out);
    return rsx_check_launch("kpp_finish");
}

extern "C" int rsx_kpp_distances(const float* d_stack, int64_t plane_stride, int64_t n, int D, const double* h_scale, const double* h_min,
                                 const double* h_mean, const double* h_cand, int T, int mode, double* d_closest, double* d_pot, double* d_scratch,
                                 rsx_stream_t stream) {
    RSX_REQUIRE(d_stack && d_closest && d_pot && d_scratch && h_cand && n > 0 && T >= 1 && T <= KPP_MAXT && mode >= 0 && mode <= 2,
                "rsx_kpp_distances: bad arguments");
    KppScale P;
    if (int rc = fill_scale(P, D, h_scale, h_min, h_mean, "rsx_kpp_distances")) return rc;
    KppCand Cn;
    memset(&Cn, 0, sizeof(Cn));
    Cn.T = mode == 1 ? T : 1;
    for (int t = 0; t < Cn.T; ++t) {
        double nn = 0.0;
        for (int d = 0; d < D; ++d) Cn.c[t][d] = h_cand[t * D + d], nn += h_cand[t * D + d] * h_cand[t * D + d];
        Cn.norm[t] = nn;
    }
    const int grid = kpp_grid(n);
    cudaStream_t s = (cudaStream_t)stream;
    kpp_distances_kernel<<<grid, 256, 0, s>>>(d_stack, plane_stride, n, P, Cn, mode, d_closest, d_scratch);
    if (int rc = rsx_check_launch("kpp_distances")) return rc;
    kpp_finish_kernel<<<1, 64, 0, s>>>(d_scratch, grid, KPP_MAXT, d_pot);
    return rsx_check_launch("kpp_finish");
}

extern "C" int rsx_kpp_block_sums(const double* d_closest, int64_t n, double* d_sums, rsx_stream_t stream) {
    RSX_REQUIRE(d_closest && d_sums && n > 0, "rsx_kpp_block_sums: bad arguments");
    kpp_block_sums_kernel<<<(int)ceil_div(n, (int64_t)KPP_BLOCK), 256, 0, (cudaStream_t)stream>>>(d_closest, n, d_sums);
    return rsx_check_launch("kpp_block_sums");
}
