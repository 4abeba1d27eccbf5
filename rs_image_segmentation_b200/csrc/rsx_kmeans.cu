// K5: KMeans (Lloyd) assign + partial-sum pass, centroid update.
//
// Replaces sklearn's lloyd_iter_chunked_dense as driven by extract.py:571-577:
//   MinMaxScaler.transform  X*scale + min_            (sklearn/preprocessing/_data.py:574-575)
//   centring                X -= X.mean(axis=0)       (sklearn/cluster/_kmeans.py:1488-1490)
//   E step                  argmin_j |c_j|^2 - 2 x.c_j, first minimum wins (_k_means_lloyd.pyx:198-212)
//   M step                  per-cluster sums / counts  (_k_means_lloyd.pyx:214-218, _k_means_common.pyx:274-296)
//
// HBM layout: the feature stack is planar float32, D planes of n_px; one pass reads 4*D bytes/pixel and
// nothing else (centroids live in constant memory and reach FFMA2 as uniform-register operands).
//
// Exactness: the parity target is sklearn run on the float64 promotion of the same stack.  The fast
// path evaluates the K distances in fp32 and keeps the best and second best; when their gap is below
// a rigorous rounding bound tau (computed per update from the centroid magnitudes) the pixel is
// re-evaluated in float64 with sklearn's operation order.  Partial sums are accumulated as int64
// fixed point (x * 2^shift_d, rounded once per sample), so they are associative: any tiling, any
// number of GPUs and any atomic ordering give bit-identical sums, hence bit-identical centroids.
#include "rsx_common.cuh"

#define KM_MAXD RSX_MAX_FEATURES
#define KM_MAXK RSX_MAX_CLUSTERS

struct KmState {
    int D, K;
    long long n_global;
    double scale64[KM_MAXD], min64[KM_MAXD], mean64[KM_MAXD];  // MinMax scale_, min_; centring mean (scaled coords)
    double absmax[KM_MAXD];                                     // max |raw x_d|
    float scale32[KM_MAXD], off32[KM_MAXD];                     // x' ~= fma(x, scale32, off32), off = min_ - mean
    float pow2[KM_MAXD];                                        // 2^shift_d (fixed-point scale of the raw feature)
    double inv_pow2[KM_MAXD];
    double cent64[KM_MAXK * KM_MAXD];                           // centred, scaled coordinates [K][D]
    double cnorm64[KM_MAXK];
    // fp32 fast path works on RAW features: dist_j = bias32[j] + sum_d x_d * w32[j][d] with
    // w = -2 c_jd scale_d and bias = |c_j|^2 - 2 sum_d c_jd (min_d - mean_d): scaling and centring are folded in
    float w32[KM_MAXK * KM_MAXD];
    float bias32[KM_MAXK];
    float cent32[KM_MAXK * KM_MAXD];                            // centred, scaled coordinates in fp32 (inertia only)
    float tau;
    float pad0;
    double shift_sq;
    int n_empty;
    int n_updates;
};

__constant__ KmState g_km;  // refreshed (device-to-device) after every setup/update

extern "C" int64_t rsx_kmeans_state_bytes(void) { return (int64_t)sizeof(KmState); }

// ----------------------------------------------------------------------------- derived tables (device, 1 CTA)
__device__ void km_derive(KmState* st) {
    // called by one CTA; thread j < K handles centroid j
    const int D = st->D, K = st->K;
    __shared__ double e_arr[KM_MAXK];
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        double cn = 0.0, bias = 0.0, mag = 0.0;
        for (int d = 0; d < D; ++d) {
            double c = st->cent64[j * KM_MAXD + d];
            double w = -2.0 * c * st->scale64[d];
            cn += c * c;
            bias += -2.0 * c * (st->min64[d] - st->mean64[d]);
            st->w32[j * KM_MAXD + d] = (float)w;
            st->cent32[j * KM_MAXD + d] = (float)c;
            mag += st->absmax[d] * fabs(w) + fabs(2.0 * c * (st->min64[d] - st->mean64[d]));
        }
        st->cnorm64[j] = cn;
        st->bias32[j] = (float)(cn + bias);
        // rounding bound of the fp32 path (DESIGN.md "KMeans near-tie bound"): every term of the D+1 term sum and
        // every partial sum is below mag + cn in magnitude; w32, bias32 and each FMA round once (u = 2^-24)
        e_arr[j] = (D + 3) * (mag + cn);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double e_max = 0.0;
        for (int j = 0; j < K; ++j) e_max = fmax(e_max, e_arr[j]);
        const double e_max_mag = e_max / (D + 3);  // largest |distance| the fp32 path can produce
        // two distances, 1.5x safety; plus the index tag written over the low mantissa bits of each distance
        st->tau = (float)(2.0 * 1.5 * e_max * 5.9604644775390625e-08 + 2.0 * e_max_mag * (K <= 8 ? 8.0 : 64.0) * 1.1920928955078125e-07);
        // never-chosen padding centroids: the kernel evaluates KM_SLOTS (K <= 8) or an even number of centroids
        for (int j = K; j < KM_MAXK && j < ((K <= 8) ? 8 : ((K + 1) & ~1)); ++j) {
            st->bias32[j] = 1e30f;  // finite: the index tag must not turn it into a NaN
            for (int d = 0; d < D; ++d) st->w32[j * KM_MAXD + d] = 0.f;
        }
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        st->scale32[d] = (float)st->scale64[d];
        st->off32[d] = (float)(st->min64[d] - st->mean64[d]);
    }
}

__global__ void km_setup_kernel(KmState* st) {
    km_derive(st);
    if (threadIdx.x == 0) {
        st->shift_sq = 0.0;
        st->n_empty = 0;
        st->n_updates = 0;
    }
}

static int km_publish(void* d_state, cudaStream_t s) {
    cudaError_t e = cudaMemcpyToSymbolAsync(g_km, d_state, sizeof(KmState), 0, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) {
        rsx_set_error("kmeans: publishing state to constant memory failed: %s", cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    return RSX_OK;
}

extern "C" int rsx_kmeans_setup(void* d_state, int D, int K, const double* h_feat_min, const double* h_feat_max, const double* h_mean_scaled,
                                const double* h_init_centroids, int64_t n_px_global, rsx_stream_t stream) {
    RSX_REQUIRE(d_state && h_feat_min && h_feat_max && h_mean_scaled && h_init_centroids, "rsx_kmeans_setup: null argument");
    RSX_REQUIRE(D >= 1 && D <= KM_MAXD && K >= 1 && K <= KM_MAXK && n_px_global > 0, "rsx_kmeans_setup: need 1<=D<=%d, 1<=K<=%d", KM_MAXD, KM_MAXK);
    static thread_local KmState h;  // staging copy; cudaMemcpyAsync from pageable memory returns after staging
    memset(&h, 0, sizeof(h));
    h.D = D, h.K = K, h.n_global = n_px_global;
    // bits available per sample so that n_global samples cannot overflow int64
    int nbits = 0;
    while (((int64_t)1 << nbits) < n_px_global) ++nbits;
    int budget = 62 - nbits;
    for (int d = 0; d < D; ++d) {
        // MinMaxScaler.fit (sklearn/preprocessing/_data.py:527-541): scale_ = 1/range (range 0 -> 1), min_ = -data_min*scale_
        double range = h_feat_max[d] - h_feat_min[d];
        if (range < 10.0 * 2.220446049250313e-16) range = 1.0;
        h.scale64[d] = 1.0 / range;
        h.min64[d] = 0.0 - h_feat_min[d] * h.scale64[d];
        h.mean64[d] = h_mean_scaled[d];
        double am = fmax(fabs(h_feat_min[d]), fabs(h_feat_max[d]));
        h.absmax[d] = am;
        int e = 0;
        if (am > 0) frexp(am, &e);  // am < 2^e
        int shift = budget - e;
        if (shift > 100) shift = 100;
        if (shift < -100) shift = -100;
        h.pow2[d] = (float)ldexp(1.0, shift);
        h.inv_pow2[d] = ldexp(1.0, -shift);
    }
    for (int j = 0; j < K; ++j)
        for (int d = 0; d < D; ++d) h.cent64[j * KM_MAXD + d] = h_init_centroids[j * D + d] - h.mean64[d];
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(d_state, &h, sizeof(h), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) {
        rsx_set_error("rsx_kmeans_setup: %s", cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    km_setup_kernel<<<1, 64, 0, s>>>((KmState*)d_state);
    if (int rc = rsx_check_launch("km_setup")) return rc;
    return km_publish(d_state, s);
}

// ----------------------------------------------------------------------------- exact (float64) re-evaluation
// Cold path: re-loads the pixel's features (L1/L2 hits) so the hot loop keeps nothing in local memory.
template <int D>
__device__ __noinline__ int km_exact_argmin(const float* __restrict__ stack, int64_t plane_stride, int64_t p, double* dist_out) {
    double X[D];
#pragma unroll
    for (int d = 0; d < D; ++d)
        X[d] = __dsub_rn(__dadd_rn(__dmul_rn((double)stack[d * plane_stride + p], g_km.scale64[d]), g_km.min64[d]), g_km.mean64[d]);
    double best = 0.0, xx = 0.0;
    int bi = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) xx = fma(X[d], X[d], xx);
    for (int j = 0; j < g_km.K; ++j) {
        double dot = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) dot = fma(X[d], g_km.cent64[j * KM_MAXD + d], dot);
        double v = fma(-2.0, dot, g_km.cnorm64[j]);
        if (j == 0 || v < best) best = v, bi = j;
    }
    *dist_out = fmax(xx + best, 0.0);
    return bi;
}

// ----------------------------------------------------------------------------- assign + partial sums
// Work decomposition: the flat pixel array is viewed as rows of row_len pixels; a tile is KM_TILE_W columns x
// KM_TILE_R rows; a thread owns 4 adjacent columns (one float4 per plane) and walks down the tile rows, with the
// next row's 4*D floats already in flight in a second register set (software prefetch).
//
// Partial sums: every thread owns KM_SLOTS accumulator slots in shared memory, laid out [slot][feature][thread]
// as int64 so that a warp's accesses are 256 contiguous bytes (2 wavefronts, no bank conflicts, no atomics).
// With K <= KM_SLOTS the slot is the label itself; for larger K the slots form a direct-mapped cache keyed by
// label % KM_SLOTS whose evictions go to a CTA-wide [K][D+1] accumulator with 32-bit shared atomics (carry
// propagated by hand).  At the end of the CTA the slots are summed over the threads and added to the global
// int64 accumulators with one 64-bit reduction per cell.
constexpr int KM_THREADS = 128;
constexpr int KM_TILE_W = KM_THREADS * 4;  // pixels per tile row
constexpr int KM_TILE_R = 32;              // rows per tile
constexpr int KM_SLOTS = 8;

template <int D>
struct KmSmem {
    static constexpr int CELLS = KM_SLOTS * (D + 1);
    static constexpr int CACHE_BYTES = CELLS * KM_THREADS * 8;
};

struct KmWalk {  // (tile, row) iteration over the tiles a CTA owns
    int64_t v_rows, tiles_total, tile;
    int64_t r, r_end, n4;
    int tiles_x, row_len, col;
    __device__ __forceinline__ bool open_tile() {
        while (tile < tiles_total) {
            const int tx = (int)(tile % tiles_x);
            const int64_t ty = tile / tiles_x;
            col = tx * KM_TILE_W + threadIdx.x * 4;
            r = ty * KM_TILE_R;
            r_end = min(v_rows, r + KM_TILE_R);
            if (col < row_len && r * row_len + col < n4) return true;
            tile += gridDim.x;
        }
        return false;
    }
    __device__ __forceinline__ int64_t start(int64_t n4_, int row_len_) {
        n4 = n4_, row_len = row_len_;
        v_rows = (n4 + row_len - 1) / row_len;
        tiles_x = (row_len + KM_TILE_W - 1) / KM_TILE_W;
        tiles_total = ((v_rows + KM_TILE_R - 1) / KM_TILE_R) * tiles_x;
        tile = blockIdx.x;
        return open_tile() ? r * row_len + col : -1;
    }
    __device__ __forceinline__ int64_t next() {
        ++r;
        if (r < r_end) {
            const int64_t p = r * row_len + col;
            if (p < n4) return p;
        }
        tile += gridDim.x;
        return open_tile() ? r * row_len + col : -1;
    }
};

template <int D, bool DIRECT>
struct KmAcc {
    long long* cache;      // [KM_SLOTS][D+1][KM_THREADS], this thread's column pre-applied
    unsigned* s_lo;        // CTA-wide [K][D+1] limbs (only when !DIRECT)
    int* s_hi;
    unsigned long long tags;  // 8 x (label+1) bytes, 0 = empty (only when !DIRECT)

    __device__ __forceinline__ long long* cell(int slot, int d) const { return cache + (slot * (D + 1) + d) * KM_THREADS; }

    __device__ __noinline__ void evict(int slot) {
        const int lab = (int)((tags >> (8 * slot)) & 0xff) - 1;
        if (lab < 0) return;
        const int base = lab * (D + 1);
#pragma unroll 1
        for (int d = 0; d <= D; ++d) {
            long long v = *cell(slot, d);
            *cell(slot, d) = 0;
            unsigned lo = (unsigned)v;
            int hi = (int)(v >> 32);
            unsigned old = atomicAdd(&s_lo[base + d], lo);
            hi += (old + lo < old) ? 1 : 0;
            if (hi) atomicAdd(&s_hi[base + d], hi);
        }
    }

    __device__ __forceinline__ void add(int label, const float (&x)[D]) {
        int slot = label;
        if (!DIRECT) {
            slot = label & (KM_SLOTS - 1);
            const int tag = (int)((tags >> (8 * slot)) & 0xff);
            if (tag != label + 1) {
                evict(slot);
                tags = (tags & ~(0xffull << (8 * slot))) | ((unsigned long long)(label + 1) << (8 * slot));
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) *cell(slot, d) += __float2ll_rn(x[d] * g_km.pow2[d]);
        *cell(slot, D) += 1;
    }
};

template <int D, bool UPDATE, bool INERTIA, bool DIRECT>
__device__ __forceinline__ int km_finish_pixel(const float* __restrict__ stack, int64_t plane_stride, int64_t p, const float (&x)[D], float best,
                                               float second, int bi, KmAcc<D, DIRECT>& acc, double& inertia, unsigned& ties) {
    double dist_exact = -1.0;
    if (!(second - best > g_km.tau)) {  // near tie (or NaN): decide in float64
        bi = km_exact_argmin<D>(stack, plane_stride, p, &dist_exact);
        ++ties;
    }
    if (INERTIA) {
        if (dist_exact < 0.0) {
            // sum of squares of (x' - c): all terms positive, relative error ~D * 2^-24
            float dd = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                float df = fmaf(x[d], g_km.scale32[d], g_km.off32[d]) - g_km.cent32[bi * KM_MAXD + d];
                dd = fmaf(df, df, dd);
            }
            dist_exact = (double)dd;
        }
        inertia += dist_exact;
    }
    if (UPDATE) acc.add(bi, x);
    return bi;
}

#define KM_ARGMIN_STEP(A, B_, S_, I_, J)          \
    S_ = fminf(S_, fmaxf(A, B_));                 \
    if (A < B_) B_ = A, I_ = J;

// Index-in-mantissa variant: the low KM_IDX_BITS bits of the fp32 distance are replaced by the centroid index, so
// best/runner-up tracking is three FMNMX and no index bookkeeping.  The perturbation (< 2^KM_IDX_BITS ulp) is part
// of the near-tie bound tau; anything closer than tau is decided in float64 anyway, so fp32 ties never pick a label.
constexpr int KM_IDX_BITS_SMALL = 3;  // K <= 8 (unrolled path)
constexpr int KM_IDX_BITS_LARGE = 6;  // K <= 64
template <int BITS>
__device__ __forceinline__ float km_tag(float a, int j) { return __uint_as_float((__float_as_uint(a) & ~((1u << BITS) - 1u)) | (unsigned)j); }
#define KM_ARGMIN_TAGGED(A, B_, S_, J)                \
    {                                                 \
        const float t_ = km_tag<KM_IDX_BITS>(A, J);   \
        S_ = fminf(S_, fmaxf(t_, B_));                \
        B_ = fminf(B_, t_);                           \
    }

template <int D, bool UPDATE, bool INERTIA, int KU>
__global__ void __launch_bounds__(KM_THREADS, 2) km_assign_kernel(const float* __restrict__ stack, int64_t plane_stride, int64_t n_px, int row_len,
                                                                  long long* __restrict__ gacc, uint8_t* __restrict__ lab8,
                                                                  const uint8_t* __restrict__ prev8, int32_t* __restrict__ lab32,
                                                                  double* __restrict__ inertia_out) {
    constexpr bool DIRECT = KU > 0;  // K <= KM_SLOTS: an accumulator slot per label, no evictions
    extern __shared__ __align__(16) unsigned char km_smem[];
    const int K = g_km.K;
    KmAcc<D, DIRECT> acc;
    acc.cache = reinterpret_cast<long long*>(km_smem) + threadIdx.x;
    acc.s_lo = reinterpret_cast<unsigned*>(km_smem + KmSmem<D>::CACHE_BYTES);
    acc.s_hi = reinterpret_cast<int*>(acc.s_lo + K * (D + 1));
    acc.tags = 0;
    if (UPDATE) {
        for (int i = threadIdx.x; i < KmSmem<D>::CELLS * KM_THREADS; i += KM_THREADS) reinterpret_cast<long long*>(km_smem)[i] = 0;
        if (!DIRECT)
            for (int i = threadIdx.x; i < 2 * K * (D + 1); i += KM_THREADS) acc.s_lo[i] = 0;
        __syncthreads();
    }
    double inertia = 0.0;
    unsigned ties = 0, changed = 0;
    const int64_t n4 = n_px & ~(int64_t)3;
    const int Kp = (K + 1) & ~1;  // centroids are processed in pairs; slot K (if K is odd) holds bias=+inf

    KmWalk walk;
    int64_t p = walk.start(n4, row_len);
    float4 va[D], vb[D];  // ping-pong register sets: one is consumed while the other is being loaded
    auto load_row = [&](float4 (&dst)[D], int64_t at) {
#pragma unroll
        for (int d = 0; d < D; ++d) dst[d] = ldg_stream4(stack + d * plane_stride + at);
    };
    // L2 prefetch of the row two steps ahead (no registers, no scoreboard): HBM latency is paid there, the
    // register loads of the next row then hit L2
    const int lane_pf = threadIdx.x & 31;
    const bool do_pf = (lane_pf & 7) == 0 || lane_pf == 31;
    auto prefetch_l2 = [&](int64_t at) {
        if (do_pf && at + 4 <= n4) {
#pragma unroll
            for (int d = 0; d < D; ++d) asm volatile("prefetch.global.L2 [%0];" ::"l"(stack + d * plane_stride + at));
        }
    };
    auto process_row = [&](const float4 (&v)[D], int64_t p) {
        prefetch_l2(p + 2 * (int64_t)row_len);
        constexpr int KM_IDX_BITS = KU > 0 ? KM_IDX_BITS_SMALL : KM_IDX_BITS_LARGE;
        float b0 = INFINITY, b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
        float s0 = INFINITY, s1 = INFINITY, s2 = INFINITY, s3 = INFINITY;
        if (KU > 0) {
            // K <= KU: fully unrolled, the weights are immediate constant-bank operands of FFMA2 (no loads at all)
#pragma unroll
            for (int j = 0; j < KU; ++j) {
                const float cA = g_km.bias32[j];
                float2 a01 = make_float2(cA, cA), a23 = a01;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float wa = g_km.w32[j * KM_MAXD + d];
                    a01 = __ffma2_rn(make_float2(v[d].x, v[d].y), make_float2(wa, wa), a01);
                    a23 = __ffma2_rn(make_float2(v[d].z, v[d].w), make_float2(wa, wa), a23);
                }
                KM_ARGMIN_TAGGED(a01.x, b0, s0, j) KM_ARGMIN_TAGGED(a01.y, b1, s1, j)
                KM_ARGMIN_TAGGED(a23.x, b2, s2, j) KM_ARGMIN_TAGGED(a23.y, b3, s3, j)
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < Kp; j += 2) {
                const float cA = g_km.bias32[j], cB = g_km.bias32[j + 1];
                float2 a01 = make_float2(cA, cA), a23 = a01, e01 = make_float2(cB, cB), e23 = e01;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float wa = g_km.w32[j * KM_MAXD + d], wb = g_km.w32[(j + 1) * KM_MAXD + d];
                    const float2 x01 = make_float2(v[d].x, v[d].y), x23 = make_float2(v[d].z, v[d].w);
                    a01 = __ffma2_rn(x01, make_float2(wa, wa), a01);
                    a23 = __ffma2_rn(x23, make_float2(wa, wa), a23);
                    e01 = __ffma2_rn(x01, make_float2(wb, wb), e01);
                    e23 = __ffma2_rn(x23, make_float2(wb, wb), e23);
                }
                KM_ARGMIN_TAGGED(a01.x, b0, s0, j) KM_ARGMIN_TAGGED(a01.y, b1, s1, j)
                KM_ARGMIN_TAGGED(a23.x, b2, s2, j) KM_ARGMIN_TAGGED(a23.y, b3, s3, j)
                KM_ARGMIN_TAGGED(e01.x, b0, s0, j + 1) KM_ARGMIN_TAGGED(e01.y, b1, s1, j + 1)
                KM_ARGMIN_TAGGED(e23.x, b2, s2, j + 1) KM_ARGMIN_TAGGED(e23.y, b3, s3, j + 1)
            }
        }
        const int i0 = (int)(__float_as_uint(b0) & ((1u << KM_IDX_BITS) - 1u)), i1 = (int)(__float_as_uint(b1) & ((1u << KM_IDX_BITS) - 1u));
        const int i2 = (int)(__float_as_uint(b2) & ((1u << KM_IDX_BITS) - 1u)), i3 = (int)(__float_as_uint(b3) & ((1u << KM_IDX_BITS) - 1u));
        int l0, l1, l2, l3;
        {
            float x[D];
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = v[d].x;
            l0 = km_finish_pixel<D, UPDATE, INERTIA, DIRECT>(stack, plane_stride, p, x, b0, s0, i0, acc, inertia, ties);
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = v[d].y;
            l1 = km_finish_pixel<D, UPDATE, INERTIA, DIRECT>(stack, plane_stride, p + 1, x, b1, s1, i1, acc, inertia, ties);
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = v[d].z;
            l2 = km_finish_pixel<D, UPDATE, INERTIA, DIRECT>(stack, plane_stride, p + 2, x, b2, s2, i2, acc, inertia, ties);
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = v[d].w;
            l3 = km_finish_pixel<D, UPDATE, INERTIA, DIRECT>(stack, plane_stride, p + 3, x, b3, s3, i3, acc, inertia, ties);
        }
        const uint32_t packed = (uint32_t)l0 | ((uint32_t)l1 << 8) | ((uint32_t)l2 << 16) | ((uint32_t)l3 << 24);
        if (prev8) {  // strict-convergence test of sklearn (_kmeans.py:723): count labels that differ from the previous pass
            const uint32_t x = packed ^ *reinterpret_cast<const uint32_t*>(prev8 + p);
            changed += ((x & 0xffu) != 0) + ((x & 0xff00u) != 0) + ((x & 0xff0000u) != 0) + ((x & 0xff000000u) != 0);
        }
        if (lab8) *reinterpret_cast<uint32_t*>(lab8 + p) = packed;
        if (lab32) *reinterpret_cast<int4*>(lab32 + p) = make_int4(l0, l1, l2, l3);
    };
    if (p >= 0) load_row(va, p);
    while (p >= 0) {
        const int64_t pn = walk.next();
        if (pn >= 0) load_row(vb, pn);
        process_row(va, p);
        if (pn < 0) break;
        p = walk.next();
        if (p >= 0) load_row(va, p);
        process_row(vb, pn);
    }
    // ragged tail (n_px % 4 pixels): one thread, scalar
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t q = n4; q < n_px; ++q) {
            float x[D];
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = stack[d * plane_stride + q];
            float b = INFINITY, s = INFINITY;
            int bi = 0;
            for (int j = 0; j < K; ++j) {
                float a = g_km.bias32[j];
#pragma unroll
                for (int d = 0; d < D; ++d) a = fmaf(x[d], g_km.w32[j * KM_MAXD + d], a);
                KM_ARGMIN_STEP(a, b, s, bi, j)
            }
            int l = km_finish_pixel<D, UPDATE, INERTIA, DIRECT>(stack, plane_stride, q, x, b, s, bi, acc, inertia, ties);
            if (prev8) changed += prev8[q] != (uint8_t)l;
            if (lab8) lab8[q] = (uint8_t)l;
            if (lab32) lab32[q] = l;
        }
    }
    if (UPDATE) {
        if (DIRECT) {
            __syncthreads();
            // column sums over the CTA's threads: warp w takes cells w, w+4, ...
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const long long* base = reinterpret_cast<const long long*>(km_smem);
            for (int c = warp; c < K * (D + 1); c += KM_THREADS / 32) {
                long long t = 0;
#pragma unroll
                for (int i = 0; i < KM_THREADS / 32; ++i) t += base[c * KM_THREADS + lane + 32 * i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0 && t) {
                    const int j = c / (D + 1), d = c % (D + 1);
                    long long* dst = d < D ? &gacc[j * D + d] : &gacc[K * D + j];
                    atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)t);
                }
            }
        } else {
#pragma unroll 1
            for (int slot = 0; slot < KM_SLOTS; ++slot) acc.evict(slot);
            __syncthreads();
            for (int i = threadIdx.x; i < K * (D + 1); i += KM_THREADS) {
                long long t = ((long long)acc.s_hi[i] << 32) + (long long)acc.s_lo[i];
                if (t) {
                    const int j = i / (D + 1), d = i % (D + 1);
                    long long* dst = d < D ? &gacc[j * D + d] : &gacc[K * D + j];
                    atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)t);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ties += __shfl_xor_sync(0xffffffffu, ties, o);
        changed += __shfl_xor_sync(0xffffffffu, changed, o);
        if (INERTIA) inertia += __shfl_xor_sync(0xffffffffu, inertia, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (ties && gacc) atomicAdd(reinterpret_cast<unsigned long long*>(&gacc[K * D + K]), (unsigned long long)ties);
        if (changed && gacc) atomicAdd(reinterpret_cast<unsigned long long*>(&gacc[K * D + K + 1]), (unsigned long long)changed);
        if (INERTIA && inertia_out) atomicAdd(inertia_out, inertia);
    }
}

template <int D, bool UPDATE, bool INERTIA, int KU>
static int km_launch2(const float* d_stack, int64_t plane_stride, int64_t n_px, int row_len, long long* acc, uint8_t* l8, const uint8_t* p8,
                      int32_t* l32, double* inertia, int K, int grid, cudaStream_t s) {
    int smem = UPDATE ? KmSmem<D>::CACHE_BYTES + (KU > 0 ? 0 : 2 * K * (D + 1) * 4) : 0;
    auto kern = km_assign_kernel<D, UPDATE, INERTIA, KU>;
    static int configured = -1;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max(smem, 48 * 1024));
        if (e != cudaSuccess) {
            rsx_set_error("km_assign: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
            return RSX_ERR_CUDA;
        }
        configured = max(smem, 48 * 1024);
    }
    kern<<<grid, KM_THREADS, smem, s>>>(d_stack, plane_stride, n_px, row_len, acc, l8, p8, l32, inertia);
    return rsx_check_launch("km_assign");
}

template <int D>
static int km_launch(const float* d_stack, int64_t plane_stride, int64_t n_px, int row_len, long long* acc, uint8_t* l8, const uint8_t* p8,
                     int32_t* l32, double* inertia, int update, int K, cudaStream_t s) {
    const int64_t n4 = n_px & ~(int64_t)3;
    const int64_t v_rows = (n4 + row_len - 1) / row_len;
    const int64_t n_tiles = ceil_div(v_rows, (int64_t)KM_TILE_R) * ceil_div(row_len, KM_TILE_W);
    const int per_sm = (update && KmSmem<D>::CACHE_BYTES > 113 * 1024) ? 1 : 2;
    const int grid = (int)max((int64_t)1, min(n_tiles, (int64_t)rsx_num_sms() * per_sm));
    const bool direct = K <= KM_SLOTS;
    if (update)  // inertia is only produced by the final (assign-only) pass
        return direct ? km_launch2<D, true, false, KM_SLOTS>(d_stack, plane_stride, n_px, row_len, acc, l8, p8, l32, nullptr, K, grid, s)
                      : km_launch2<D, true, false, 0>(d_stack, plane_stride, n_px, row_len, acc, l8, p8, l32, nullptr, K, grid, s);
    if (inertia)
        return direct ? km_launch2<D, false, true, KM_SLOTS>(d_stack, plane_stride, n_px, row_len, acc, l8, p8, l32, inertia, K, grid, s)
                      : km_launch2<D, false, true, 0>(d_stack, plane_stride, n_px, row_len, acc, l8, p8, l32, inertia, K, grid, s);
    return direct ? km_launch2<D, false, false, KM_SLOTS>(d_stack, plane_stride, n_px, row_len, acc, l8, p8, l32, inertia, K, grid, s)
                  : km_launch2<D, false, false, 0>(d_stack, plane_stride, n_px, row_len, acc, l8, p8, l32, inertia, K, grid, s);
}

extern "C" int rsx_kmeans_assign(const float* d_stack, int64_t plane_stride, int64_t n_px, int row_len, const void* d_state, int64_t* d_acc,
                                 uint8_t* d_labels_u8, const uint8_t* d_labels_prev_u8, int32_t* d_labels_i32, double* d_inertia, int update,
                                 int D, int K, rsx_stream_t stream) {
    RSX_REQUIRE(d_stack && d_state && n_px > 0, "rsx_kmeans_assign: bad arguments");
    RSX_REQUIRE(D >= 1 && D <= KM_MAXD && K >= 1 && K <= KM_MAXK, "rsx_kmeans_assign: D/K out of range");
    RSX_REQUIRE(!update || d_acc, "rsx_kmeans_assign: update pass needs d_acc");
    RSX_REQUIRE(!(update && d_inertia), "rsx_kmeans_assign: inertia is produced by the assign-only pass (update == 0)");
    RSX_REQUIRE(((uintptr_t)d_stack & 15) == 0 && (plane_stride & 3) == 0, "rsx_kmeans_assign: stack planes must be 16-byte aligned");
    RSX_REQUIRE((((uintptr_t)d_labels_u8 | (uintptr_t)d_labels_prev_u8) & 3) == 0 && (((uintptr_t)d_labels_i32) & 15) == 0,
                "rsx_kmeans_assign: label buffers must be aligned");
    RSX_REQUIRE(!d_labels_prev_u8 || d_acc, "rsx_kmeans_assign: the changed-label counter lives in d_acc");
    if (row_len <= 0) row_len = 4096;
    row_len = (row_len + 3) & ~3;
    cudaStream_t s = (cudaStream_t)stream;
    long long* acc = reinterpret_cast<long long*>(d_acc);
#define CASE(DD) \
    case DD: return km_launch<DD>(d_stack, plane_stride, n_px, row_len, acc, d_labels_u8, d_labels_prev_u8, d_labels_i32, d_inertia, update, K, s);
    switch (D) {
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
        CASE(17) CASE(18) CASE(19) CASE(20)
        default:
            rsx_set_error("rsx_kmeans_assign: D=%d not compiled (1..20)", D);
            return RSX_ERR_UNSUPPORTED;
    }
#undef CASE
}

// ----------------------------------------------------------------------------- centroid update (1 CTA)
__global__ void km_update_kernel(KmState* st, long long* acc) {
    const int D = st->D, K = st->K;
    __shared__ double shift_part[KM_MAXK];
    __shared__ int empty_part[KM_MAXK];
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        long long cnt = acc[K * D + j];
        double sh = 0.0;
        int empty = 0;
        if (cnt > 0) {
            // _average_centers: centers *= 1/weight  (_k_means_common.pyx:274-296)
            double alpha = 1.0 / (double)cnt;
            for (int d = 0; d < D; ++d) {
                double mean_raw = ((double)acc[j * D + d] * st->inv_pow2[d]) * alpha;
                double c = (mean_raw * st->scale64[d] + st->min64[d]) - st->mean64[d];
                double old = st->cent64[j * KM_MAXD + d];
                sh += (c - old) * (c - old);
                st->cent64[j * KM_MAXD + d] = c;
            }
        } else {
            empty = 1;  // relocation (_k_means_common.pyx:167-211) is not done on the device: reported to the host
        }
        shift_part[j] = sh;
        empty_part[j] = empty;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        int e = 0;
        for (int j = 0; j < K; ++j) s += shift_part[j], e += empty_part[j];
        st->shift_sq = s;
        st->n_empty += e;
        st->n_updates += 1;
    }
    __syncthreads();
    km_derive(st);
    __syncthreads();
    for (int i = threadIdx.x; i < K * (D + 1); i += blockDim.x) acc[i] = 0;  // ties counter (last slot) keeps accumulating
}

extern "C" int rsx_kmeans_update(void* d_state, int64_t* d_acc, rsx_stream_t stream) {
    RSX_REQUIRE(d_state && d_acc, "rsx_kmeans_update: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    km_update_kernel<<<1, 64, 0, s>>>((KmState*)d_state, reinterpret_cast<long long*>(d_acc));
    if (int rc = rsx_check_launch("km_update")) return rc;
    return km_publish(d_state, s);
}

extern "C" int rsx_kmeans_read(const void* d_state, double* h_centroids, double* h_shift_sq, int32_t* h_empty, rsx_stream_t stream) {
    RSX_REQUIRE(d_state, "rsx_kmeans_read: bad arguments");
    static thread_local KmState h;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(&h, d_state, sizeof(h), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        rsx_set_error("rsx_kmeans_read: %s", cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    if (h_centroids)
        for (int j = 0; j < h.K; ++j)
            for (int d = 0; d < h.D; ++d) h_centroids[j * h.D + d] = h.cent64[j * KM_MAXD + d] + h.mean64[d];
    if (h_shift_sq) *h_shift_sq = h.shift_sq;
    if (h_empty) *h_empty = h.n_empty;
    return RSX_OK;
}
