// K5: KMeans (Lloyd): host-side API, state set-up and the centroid update kernel.  The assign kernels live in
// rsx_kmeans_part.cu (one translation unit per range of D).
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
#include <cstdlib>

#include "rsx_kmeans_state.cuh"

extern "C" int64_t rsx_kmeans_state_bytes(void) { return (int64_t)sizeof(KmState); }

// ----------------------------------------------------------------------------- derived tables (device, 1 CTA)
// The 16-bit grid of the screening passes depends on the feature ranges only: set-up kernels, once (the CTA synchronises after it).
__device__ void km_derive_grid(KmState* st) {
    for (int d = threadIdx.x; d < st->D; d += blockDim.x) {
        // grid over [fmin, fmax] (a constant feature: step 0, every sample reads back as fmin)
        const double range = st->fmax64[d] - st->fmin64[d];
        const bool flat = !(range >= 10.0 * 2.220446049250313e-16);
        const float step = flat ? 0.f : (float)(range / 65535.0);
        st->qstep32[d] = step;
        st->qinv32[d] = flat ? 0.f : (float)(65535.0 / range);
        st->qmin32[d] = (float)st->fmin64[d];
        st->qoff32[d] = (float)(st->fmin64[d] - 8388608.0 * (double)step);
        // |x~ - x|: 0.51 steps from the rounding to the grid (the fp32 evaluation of (x - fmin) * qinv32 included), 0.5 steps from the
        // rounding of qoff32 (its magnitude is 2^23 steps), 2^-24 relative from qstep32 and from the fma: 1.02 steps + 2^-22 (range + |x|)
        st->qerr32[d] = __double2float_ru(1.02 * (double)step + 2.384185791015625e-07 * ((flat ? 0.0 : range) + st->absmax[d]));
    }
    __syncthreads();
}

__device__ void km_derive(KmState* st) {
    // called by one CTA; thread j < K handles centroid j
    const int D = st->D, K = st->K;
    __shared__ double e_arr[KM_MAXK], q_arr[KM_MAXK];
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        double cn = 0.0, bias = 0.0, mag = 0.0, qe = 0.0;
        for (int d = 0; d < D; ++d) {
            double c = st->cent64[j * KM_MAXD + d];
            double w = -2.0 * c * st->scale64[d];
            cn += c * c;
            bias += -2.0 * c * (st->min64[d] - st->mean64[d]);
            st->w32[j * KM_MAXD + d] = (float)w;
            st->cent32[j * KM_MAXD + d] = (float)c;
            mag += st->absmax[d] * fabs(w) + fabs(2.0 * c * (st->min64[d] - st->mean64[d]));
            qe += fabs(w) * (double)st->qerr32[d];
        }
        st->cnorm64[j] = cn;
        st->bias32[j] = (float)(cn + bias);
        // rounding bound of the fp32 path (DESIGN.md "KMeans near-tie bound"): every term of the D+1 term sum and
        // every partial sum is below mag + cn in magnitude; w32, bias32 and each FMA round once (u = 2^-24)
        e_arr[j] = (D + 3) * (mag + cn);
        q_arr[j] = qe;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double e_max = 0.0, q_max = 0.0;
        for (int j = 0; j < K; ++j) e_max = fmax(e_max, e_arr[j]), q_max = fmax(q_max, q_arr[j]);
        const double e_max_mag = e_max / (D + 3);  // largest |distance| the fp32 path can produce
        // two distances, 1.5x safety; plus the index tag written over the low mantissa bits of each distance
        st->tau_tight = (float)(2.0 * 1.5 * e_max * 5.9604644775390625e-08);
        const int tag_bits = K <= 8 ? 3 : K <= 16 ? 4 : K <= 32 ? 5 : 6;
        st->tag_bits = tag_bits;
        const double tag_term = 2.0 * e_max_mag * (double)(1 << tag_bits) * 1.1920928955078125e-07;
        st->tau = (float)(2.0 * 1.5 * e_max * 5.9604644775390625e-08 + tag_term);
        // screening passes: what the 16-bit grid can move the difference of two distances, on top of the fp32 band
        st->tau_q = __double2float_ru((double)st->tau + 2.0 * 1.01 * q_max);
        // Tensor-core path (rsx_kmeans_part.cu km_tc_kernel): x = xh + xl + xr, w = wh + wl + wr with 11-bit pieces (|xr| < 2^-20 |x|);
        // the three products kept (xh wh + xh wl + xl wh) miss < 3 * 2^-20 |x||w| per feature, the bias is carried in three pieces
        // (< 2^-30).  Accumulation in the tensor core: measured 2^-22.8 of sum|terms| for 16 terms (tools/tc_probe.cu on the B200),
        // budgeted here as 2 * 2^-20 for the 48 terms.  Total 5 * 2^-20 of the magnitude per distance; two distances, 1.5x safety.
        st->tau_tc = (float)(2.0 * 1.5 * 5.0 * 9.5367431640625e-07 * e_max_mag + tag_term);
        // Bounded passes: squared distances q_j = |x'|^2 + dist_j.  dist_j is off by at most e_max * 2^-24 (above).  x'_d is evaluated
        // as fma(x, scale32, off32): scale32, off32 and the fma round once each, all on magnitudes below m_d = absmax scale + |off|,
        // so |x'_d - exact| <= 3 u m_d =: eps_d and |x'_d| <= m_d; the D squares and their running sum add (D + 1) u sum m_d^2.
        double xs_err = 0.0, m2 = 0.0;
        for (int d = 0; d < D; ++d) {
            const double m = st->absmax[d] * st->scale64[d] + fabs(st->min64[d] - st->mean64[d]);
            const double eps = 3.0 * 5.9604644775390625e-08 * m;
            xs_err += 2.0 * m * eps + eps * eps;
            m2 += m * m;
        }
        xs_err += (D + 1) * 5.9604644775390625e-08 * m2;
        st->bound_err = __double2float_ru(1.5 * (e_max * 5.9604644775390625e-08 + xs_err));
        // never-chosen padding centroids: the kernels evaluate centroids in groups of 8
        for (int j = K; j < KM_MAXK && j < ((K + 7) & ~7); ++j) {
            st->bias32[j] = 1e30f;  // finite: the index tag must not turn it into a NaN
            for (int d = 0; d < D; ++d) st->w32[j * KM_MAXD + d] = 0.f;
        }
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        st->scale32[d] = (float)st->scale64[d];
        st->off32[d] = (float)(st->min64[d] - st->mean64[d]);
    }
}

// The set-up / update kernels are one CTA of dependent scalar work; run against a shared-memory mirror of the state their
// per-centroid loops cost shared-memory latency instead of a chain of global-memory round trips.
constexpr int KM_CTRL_THREADS = 256;
// Only the rows of the per-centroid tables that the K centroids (rounded up to the group of 8 the kernels evaluate) occupy are
// copied: at K = 8 that is 11 of the 40 KB, and the copies in and out are most of what an update costs.
__device__ __forceinline__ void km_state_copy(void* dst, const void* src, int K) {
    static_assert(sizeof(KmState) % 16 == 0, "KmState is copied in 16-byte pieces");
    static_assert(offsetof(KmState, cent64) % 16 == 0 && offsetof(KmState, w32) % 16 == 0 && offsetof(KmState, cent32) % 16 == 0 &&
                      offsetof(KmState, cnorm64) % 16 == 0 && offsetof(KmState, bias32) % 16 == 0 && offsetof(KmState, tau) % 16 == 0,
                  "the per-centroid tables start on 16-byte boundaries");
    const int KP = min(KM_MAXK, (K + 7) & ~7);
    // [begin, end) of the unused tails, in 16-byte units
    const int a0 = (int)(offsetof(KmState, cent64) + (size_t)KP * KM_MAXD * 8) / 16, a1 = (int)offsetof(KmState, cnorm64) / 16;
    const int b0 = (int)(offsetof(KmState, w32) + (size_t)KP * KM_MAXD * 4) / 16, b1 = (int)offsetof(KmState, bias32) / 16;
    const int c0 = (int)(offsetof(KmState, cent32) + (size_t)KP * KM_MAXD * 4) / 16, c1 = (int)offsetof(KmState, tau) / 16;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(KmState) / 16); i += blockDim.x) {
        if ((i >= a0 && i < a1) || (i >= b0 && i < b1) || (i >= c0 && i < c1)) continue;
        d4[i] = s4[i];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(KM_CTRL_THREADS) km_setup_kernel(KmState* gst) {
    __shared__ __align__(16) KmState sst;
    KmState* st = &sst;
    km_state_copy(st, gst, KM_MAXK);
    km_derive_grid(st);
    km_derive(st);
    if (threadIdx.x == 0) {
        st->shift_sq = 0.0;
        st->n_empty = 0;
        st->n_updates = 0;
    }
    for (int j = threadIdx.x; j < KM_MAXK; j += blockDim.x) st->drift64[j] = 0.0, st->drift_up[j] = 0.f, st->drift_dn[j] = 0.f;
    __syncthreads();
    km_state_copy(gst, st, KM_MAXK);
}

// rsx_kmeans_setup without the host in between: the per-feature range comes straight from the min/max trackers the feature kernels
// maintained, the initial centroids from raw feature rows gathered on the device.  Same IEEE operations in the same order as
// rsx_kmeans_setup / MinMaxScaler (every product and sum rounded on its own), so the state is bit-identical.
struct KmMean {
    double v[KM_MAXD];
};
__global__ void __launch_bounds__(KM_CTRL_THREADS) km_setup_device_kernel(KmState* gst, int D, int K, const uint32_t* __restrict__ minmax,
                                                                          const double* __restrict__ rows, const KmMean mean, long long n_global) {
    __shared__ __align__(16) KmState sst;
    KmState* st = &sst;
    {
        int4* z = reinterpret_cast<int4*>(st);
        for (int i = threadIdx.x; i < (int)(sizeof(KmState) / 16); i += blockDim.x) z[i] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    if (threadIdx.x == 0) st->D = D, st->K = K, st->n_global = n_global;
    int nbits = 0;
    while ((1ll << nbits) < n_global) ++nbits;
    const int budget = 62 - nbits;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const double fmin = (double)ord2f(minmax[2 * d]), fmax_ = (double)ord2f(minmax[2 * d + 1]);
        double range = __dsub_rn(fmax_, fmin);
        if (!(range >= 10.0 * 2.220446049250313e-16)) range = 1.0;  // also for an empty tracker (+inf / -inf -> NaN)
        const double scale = __ddiv_rn(1.0, range);
        st->fmin64[d] = fmin, st->fmax64[d] = fmax_;
        st->scale64[d] = scale;
        st->min64[d] = __dsub_rn(0.0, __dmul_rn(fmin, scale));
        st->mean64[d] = mean.v[d];
        const double am = fmax(fabs(fmin), fabs(fmax_));
        st->absmax[d] = am;
        int e = 0;
        if (am > 0) frexp(am, &e);
        int shift = budget - e;
        shift = shift > 100 ? 100 : (shift < -100 ? -100 : shift);
        st->pow2[d] = (float)ldexp(1.0, shift);
        st->inv_pow2[d] = ldexp(1.0, -shift);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
        const int j = i / D, d = i - j * D;
        // MinMaxScaler.transform: X * scale_ + min_, then the centring
        st->cent64[j * KM_MAXD + d] = __dsub_rn(__dadd_rn(__dmul_rn(rows[i], st->scale64[d]), st->min64[d]), st->mean64[d]);
    }
    __syncthreads();
    km_derive_grid(st);
    km_derive(st);
    __syncthreads();
    km_state_copy(gst, st, KM_MAXK);
}

static const km_assign_fn g_part_assign[KM_NUM_PARTS] = {rsx_km_part0_assign, rsx_km_part1_assign, rsx_km_part2_assign, rsx_km_part3_assign,
                                                          rsx_km_part4_assign, rsx_km_part5_assign};
static const km_publish_fn g_part_publish[KM_NUM_PARTS] = {rsx_km_part0_publish, rsx_km_part1_publish, rsx_km_part2_publish, rsx_km_part3_publish,
                                                            rsx_km_part4_publish, rsx_km_part5_publish};

// mirror the state into the __constant__ block of the translation unit that owns the kernels for this D
static int km_publish(void* d_state, int D, cudaStream_t s) {
    if (D < 1 || D > km_part_hi(KM_NUM_PARTS - 1)) return RSX_OK;  // no kernels compiled for this depth: assign will report it
    return g_part_publish[km_part_of(D)](d_state, s);
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
        h.fmin64[d] = h_feat_min[d], h.fmax64[d] = h_feat_max[d];
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
    km_setup_kernel<<<1, KM_CTRL_THREADS, 0, s>>>((KmState*)d_state);
    if (int rc = rsx_check_launch("km_setup")) return rc;
    return km_publish(d_state, D, s);
}

extern "C" int rsx_kmeans_setup_device(void* d_state, int D, int K, const uint32_t* d_minmax, const double* d_init_rows_raw,
                                       const double* h_mean_scaled, int64_t n_px_global, rsx_stream_t stream) {
    RSX_REQUIRE(d_state && d_minmax && d_init_rows_raw, "rsx_kmeans_setup_device: null argument");
    RSX_REQUIRE(D >= 1 && D <= KM_MAXD && K >= 1 && K <= KM_MAXK && n_px_global > 0, "rsx_kmeans_setup_device: need 1<=D<=%d, 1<=K<=%d", KM_MAXD, KM_MAXK);
    KmMean mean;
    for (int d = 0; d < KM_MAXD; ++d) mean.v[d] = (h_mean_scaled && d < D) ? h_mean_scaled[d] : 0.5;
    cudaStream_t s = (cudaStream_t)stream;
    km_setup_device_kernel<<<1, KM_CTRL_THREADS, 0, s>>>((KmState*)d_state, D, K, d_minmax, d_init_rows_raw, mean, (long long)n_px_global);
    if (int rc = rsx_check_launch("km_setup_device")) return rc;
    return km_publish(d_state, D, s);
}

extern "C" int rsx_kmeans_assign(const float* d_stack, int64_t plane_stride, int64_t n_px, int row_len, const void* d_state, int64_t* d_acc,
                                 uint8_t* d_labels_u8, const uint8_t* d_labels_prev_u8, int32_t* d_labels_i32, double* d_inertia, int update,
                                 int D, int K, rsx_stream_t stream) {
    RSX_REQUIRE(d_stack && d_state && n_px > 0, "rsx_kmeans_assign: bad arguments");
    RSX_REQUIRE(D >= 1 && D <= KM_MAXD && K >= 1 && K <= KM_MAXK, "rsx_kmeans_assign: D/K out of range");
    RSX_REQUIRE(update >= 0 && update <= 2, "rsx_kmeans_assign: update must be 0 (assign), 1 (full) or 2 (delta)");
    RSX_REQUIRE(!update || d_acc, "rsx_kmeans_assign: update pass needs d_acc");
    RSX_REQUIRE(!(update && d_inertia), "rsx_kmeans_assign: inertia is produced by the assign-only pass (update == 0)");
    RSX_REQUIRE(update != KM_DELTA || (d_labels_u8 && d_labels_prev_u8 && d_labels_u8 != d_labels_prev_u8),
                "rsx_kmeans_assign: a delta pass needs distinct previous and current uint8 label buffers");
    RSX_REQUIRE(((uintptr_t)d_stack & 15) == 0 && (plane_stride & 3) == 0, "rsx_kmeans_assign: stack planes must be 16-byte aligned");
    RSX_REQUIRE((((uintptr_t)d_labels_u8 | (uintptr_t)d_labels_prev_u8) & 3) == 0 && (((uintptr_t)d_labels_i32) & 15) == 0,
                "rsx_kmeans_assign: label buffers must be aligned");
    RSX_REQUIRE(!d_labels_prev_u8 || d_acc, "rsx_kmeans_assign: the changed-label counter lives in d_acc");
    if (D > km_part_hi(KM_NUM_PARTS - 1)) {
        rsx_set_error("rsx_kmeans_assign: D=%d not compiled (1..%d)", D, km_part_hi(KM_NUM_PARTS - 1));
        return RSX_ERR_UNSUPPORTED;
    }
    if (row_len <= 0) row_len = 4096;
    row_len = (row_len + 3) & ~3;
    KmLaunch a;
    a.stack = d_stack, a.plane_stride = plane_stride, a.n_px = n_px, a.row_len = row_len;
    a.acc = reinterpret_cast<long long*>(d_acc), a.lab8 = d_labels_u8, a.prev8 = d_labels_prev_u8, a.lab32 = d_labels_i32;
    a.inertia = d_inertia, a.mode = update, a.D = D, a.K = K;
    a.pf_rows = max(0, rsx_option("km_pf", 2));
    a.n_stages = min(4, max(0, rsx_option("km_stages", 0)));  // the stream kernel's ticket area holds 4 stages
    // 0: fp32 FFMA2 path.  1: tensor-core distances (km_tc_kernel): bit-identical labels / sums, measured at parity with the fp32 path
    // at K = 32 and slower at K = 16 on the B200 (the per-pixel argmin over K distances, not the FMAs, is what both are bound by;
    // DESIGN.md 4.2), so it stays an option.
    a.use_tc = rsx_option("km_tc", 0);
    a.bounded = 0, a.aos = nullptr, a.slack = nullptr, a.q16 = nullptr, a.q_stride = 0;
    a.full_stream = rsx_option("km_full_stream", 0);
    return g_part_assign[km_part_of(D)](a, (cudaStream_t)stream);
}

extern "C" int64_t rsx_kmeans_aos_stride(int D) { return D >= 1 && D <= KM_MAXD ? (int64_t)km_aos_stride(D) : 0; }

// A delta pass with Hamerly's test in front (K <= 8; km_bounded_kernel in rsx_kmeans_part.cu).  Labels are updated IN PLACE.
extern "C" int rsx_kmeans_assign_bounded(const float* d_stack, int64_t plane_stride, int64_t n_px, const void* d_state, int64_t* d_acc,
                                         uint8_t* d_labels_u8, float* d_aos, float* d_slack, int first, int D, int K, rsx_stream_t stream) {
    RSX_REQUIRE(d_stack && d_state && d_acc && d_labels_u8 && d_aos && d_slack && n_px > 0, "rsx_kmeans_assign_bounded: null argument");
    RSX_REQUIRE(D >= 1 && D <= KM_MAXD && K >= 1 && K <= KM_MAXK, "rsx_kmeans_assign_bounded: D/K out of range");
    RSX_REQUIRE((((uintptr_t)d_labels_u8) & 3) == 0 && (((uintptr_t)d_aos | (uintptr_t)d_slack) & 15) == 0, "rsx_kmeans_assign_bounded: buffers must be aligned");
    if (D > km_part_hi(KM_NUM_PARTS - 1)) {
        rsx_set_error("rsx_kmeans_assign_bounded: D=%d not compiled (1..%d)", D, km_part_hi(KM_NUM_PARTS - 1));
        return RSX_ERR_UNSUPPORTED;
    }
    KmLaunch a;
    memset(&a, 0, sizeof(a));
    a.stack = d_stack, a.plane_stride = plane_stride, a.n_px = n_px, a.row_len = 4096;
    a.acc = reinterpret_cast<long long*>(d_acc), a.lab8 = d_labels_u8;
    a.mode = KM_DELTA, a.D = D, a.K = K;
    a.bounded = first ? 1 : 2, a.aos = d_aos, a.slack = d_slack;
    return g_part_assign[km_part_of(D)](a, (cudaStream_t)stream);
}

// 16-bit screening passes (K <= 8): rsx_kmeans_quantize_u16 writes the uint16 copy of the stack on the grid the state derives from
// the feature ranges; rsx_kmeans_assign_q16 is a delta pass (update = 2 of rsx_kmeans_assign) that reads the copy instead of the
// float32 planes and falls back to them for the pixels it cannot decide and for the samples that move.
extern "C" int rsx_kmeans_quantize_u16(const float* d_stack, int64_t plane_stride, int64_t n_px, const void* d_state, uint16_t* d_q16, int64_t q_stride,
                                       int D, rsx_stream_t stream) {
    RSX_REQUIRE(d_stack && d_state && d_q16 && n_px > 0 && D >= 1 && D <= KM_MAXD, "rsx_kmeans_quantize_u16: bad arguments");
    RSX_REQUIRE(((uintptr_t)d_stack & 15) == 0 && (plane_stride & 3) == 0 && ((uintptr_t)d_q16 & 15) == 0 && (q_stride & 7) == 0 && q_stride >= ((n_px + 7) & ~(int64_t)7),
                "rsx_kmeans_quantize_u16: planes must be 16-byte aligned, q_stride a multiple of 8 and >= n_px rounded up to 8");
    if (D > km_part_hi(KM_NUM_PARTS - 1)) {
        rsx_set_error("rsx_kmeans_quantize_u16: D=%d not compiled (1..%d)", D, km_part_hi(KM_NUM_PARTS - 1));
        return RSX_ERR_UNSUPPORTED;
    }
    KmLaunch a;
    memset(&a, 0, sizeof(a));
    a.stack = d_stack, a.plane_stride = plane_stride, a.n_px = n_px, a.D = D, a.q16 = d_q16, a.q_stride = q_stride, a.mode = KM_QUANTIZE;
    return g_part_assign[km_part_of(D)](a, (cudaStream_t)stream);
}

extern "C" int rsx_kmeans_assign_q16(const float* d_stack, int64_t plane_stride, int64_t n_px, const void* d_state, int64_t* d_acc, uint8_t* d_labels_u8,
                                     const uint8_t* d_labels_prev_u8, const uint16_t* d_q16, int64_t q_stride, int D, int K, rsx_stream_t stream) {
    RSX_REQUIRE(d_stack && d_state && d_acc && d_labels_u8 && d_labels_prev_u8 && d_q16 && n_px > 0, "rsx_kmeans_assign_q16: null argument");
    RSX_REQUIRE(d_labels_u8 != d_labels_prev_u8, "rsx_kmeans_assign_q16: a delta pass needs distinct previous and current label buffers");
    RSX_REQUIRE(D >= 1 && D <= KM_MAXD && K >= 1 && K <= 8, "rsx_kmeans_assign_q16: needs 1 <= K <= 8");
    RSX_REQUIRE(((uintptr_t)d_stack & 15) == 0 && (plane_stride & 3) == 0 && ((uintptr_t)d_q16 & 15) == 0 && (q_stride & 7) == 0,
                "rsx_kmeans_assign_q16: planes must be 16-byte aligned");
    RSX_REQUIRE((((uintptr_t)d_labels_u8 | (uintptr_t)d_labels_prev_u8) & 3) == 0, "rsx_kmeans_assign_q16: label buffers must be aligned");
    if (D > km_part_hi(KM_NUM_PARTS - 1)) {
        rsx_set_error("rsx_kmeans_assign_q16: D=%d not compiled (1..%d)", D, km_part_hi(KM_NUM_PARTS - 1));
        return RSX_ERR_UNSUPPORTED;
    }
    KmLaunch a;
    memset(&a, 0, sizeof(a));
    a.stack = d_stack, a.plane_stride = plane_stride, a.n_px = n_px, a.row_len = 4096;
    a.acc = reinterpret_cast<long long*>(d_acc), a.lab8 = d_labels_u8, a.prev8 = d_labels_prev_u8;
    a.mode = KM_DELTA, a.D = D, a.K = K, a.q16 = d_q16, a.q_stride = q_stride;
    a.n_stages = min(4, max(0, rsx_option("km_stages", 0)));
    return g_part_assign[km_part_of(D)](a, (cudaStream_t)stream);
}

// ----------------------------------------------------------------------------- centroid update (1 CTA)
// acc = pass block [sums K*D][counts K][near ties][changed] ++ totals block [sums K*D][counts K][near ties so far][changed in
// the last pass].  A full pass produced the sums themselves (delta == 0: totals <- pass), a delta pass their change (totals +=
// pass); the centroids come from the totals; the pass block is zeroed for the next pass.
// adjust (optional, [K*D + K]): empty-cluster relocation of sklearn (_k_means_common.pyx:167-211) - added to the totals for
// THIS centroid computation only; the running totals stay "sums by label", which is what the next delta pass builds on.
// Peer reduction (multi-GPU, one node): every rank's assign pass accumulated into ITS block of peer-mapped memory; the update
// kernels of all ranks meet at a flag barrier over NVLink, then each sums the ranks' blocks itself (integer sums: the same
// result on every rank) - the all-reduce is part of the update kernel, no collective is launched.
//   block layout (RSX_PEER_BLOCK_BYTES): two pass buffers [RSX_PEER_PASS_ELEMS] int64 (parity of the sequence number), then
//   flags [RSX_MAX_PEERS] uint64: flags[p] = last sequence number rank p has published.
// Safety of the buffers: rank r zeroes its buffer of parity (seq+1)&1 after the barrier of seq; a peer read it during its
// update seq-1, which precedes that peer's flag seq in stream order.
struct KmPeers {
    const long long* pass[RSX_MAX_PEERS];         // every rank's pass buffer of this parity (own included)
    unsigned long long* flag_out[RSX_MAX_PEERS];  // &flags[rank] inside every peer's block
    const unsigned long long* flag_in;            // own flags
    long long* zero_next;                         // own pass buffer of the other parity
    unsigned long long seq;
    unsigned long long timeout_ns;  // a peer that has not published after this long is reported instead of waited for
    int rank, world;
};

__device__ __forceinline__ unsigned long long km_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// returns false on time-out (a peer never arrived): the caller records it instead of hanging the GPU
__device__ bool km_peer_reduce(const KmPeers& pr, long long* __restrict__ pass_out, int n) {
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    if (threadIdx.x < pr.world && threadIdx.x != pr.rank) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pr.flag_out[threadIdx.x]), "l"(pr.seq) : "memory");
        const unsigned long long t0 = km_globaltimer();
        unsigned long long seen = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(pr.flag_in + threadIdx.x) : "memory");
            if (seen >= pr.seq) break;
            if (km_globaltimer() - t0 > pr.timeout_ns) {
                timed_out = 1;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (timed_out) return false;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        long long v = 0;
        for (int p = 0; p < pr.world; ++p) v += *reinterpret_cast<const volatile long long*>(pr.pass[p] + i);
        pass_out[i] = v;
        pr.zero_next[i] = 0;
    }
    __syncthreads();
    return true;
}

template <bool PEERS>
__global__ void __launch_bounds__(KM_CTRL_THREADS) km_update_kernel(KmState* gst, long long* acc, int delta, const long long* __restrict__ adjust,
                                                                    const KmPeers peers) {
    __shared__ __align__(16) KmState sst;
    KmState* st = &sst;
    km_state_copy(st, gst, gst->K);
    const int D = st->D, K = st->K;
    if (PEERS) {
        if (!km_peer_reduce(peers, acc, K * D + K + 2)) {
            if (threadIdx.x == 0) gst->n_empty = -1000000;  // reported by rsx_kmeans_read as a failed exchange
            return;
        }
    }
    long long* tot = acc + K * D + K + 2;
    for (int i = threadIdx.x; i < K * D + K; i += blockDim.x) {
        const long long v = acc[i] + (delta ? tot[i] : 0ll);
        tot[i] = v;
    }
    if (threadIdx.x == 0) {
        tot[K * D + K] += acc[K * D + K];
        tot[K * D + K + 1] = acc[K * D + K + 1];
    }
    __syncthreads();
    acc = tot;
    __shared__ double shift_part[KM_MAXK];
    __shared__ int empty_part[KM_MAXK];
    // one warp per centroid, lane = feature: the D sums of a centroid arrive with one round trip to memory instead of D dependent
    // ones, and the squared shift is added up by a fixed shuffle tree (the same bits on every rank and in every run)
    for (int j = threadIdx.x >> 5; j < K; j += blockDim.x >> 5) {
        const int d = threadIdx.x & 31;
        const long long cnt = acc[K * D + j] + (adjust ? adjust[K * D + j] : 0ll);
        double sq = 0.0;
        if (cnt > 0 && d < D) {
            // _average_centers: centers *= 1/weight  (_k_means_common.pyx:274-296)
            const double alpha = 1.0 / (double)cnt;
            const long long sum_q = acc[j * D + d] + (adjust ? adjust[j * D + d] : 0ll);
            const double mean_raw = ((double)sum_q * st->inv_pow2[d]) * alpha;
            const double c = (mean_raw * st->scale64[d] + st->min64[d]) - st->mean64[d];
            const double old = st->cent64[j * KM_MAXD + d];
            sq = (c - old) * (c - old);
            st->cent64[j * KM_MAXD + d] = c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (d == 0) {
            shift_part[j] = sq;
            empty_part[j] = cnt > 0 ? 0 : 1;  // no relocation was supplied: the centre keeps its position and the event is reported
        }
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
    // Hamerly drift of every label: its own centre's move plus the largest move among the others (rounded outwards: the moves
    // carry ~1e-15 of relative error, the float copies are pushed one further ulp apart)
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        double other = 0.0;
        for (int i = 0; i < K; ++i)
            if (i != j) other = fmax(other, shift_part[i]);
        const double inc = (sqrt(shift_part[j]) + sqrt(other)) * (1.0 + 1e-9);
        const double c = st->drift64[j] + inc;
        st->drift64[j] = c;
        st->drift_up[j] = nextafterf(__double2float_ru(c), INFINITY);
        st->drift_dn[j] = fmaxf(nextafterf(__double2float_rd(c), -INFINITY), 0.f);
    }
    __syncthreads();
    km_derive(st);
    __syncthreads();
    for (int i = threadIdx.x; i < K * D + K + 2; i += blockDim.x) (tot - (K * D + K + 2))[i] = 0;
    km_state_copy(gst, st, K);
}

extern "C" int rsx_kmeans_update(void* d_state, int64_t* d_acc, int delta, int D, const int64_t* d_adjust, rsx_stream_t stream) {
    RSX_REQUIRE(d_state && d_acc && D >= 1 && D <= KM_MAXD, "rsx_kmeans_update: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    km_update_kernel<false><<<1, KM_CTRL_THREADS, 0, s>>>((KmState*)d_state, reinterpret_cast<long long*>(d_acc), delta,
                                                          reinterpret_cast<const long long*>(d_adjust), KmPeers());
    if (int rc = rsx_check_launch("km_update")) return rc;
    return km_publish(d_state, D, s);
}

extern "C" int rsx_kmeans_update_peers(void* d_state, int64_t* d_acc, int delta, int D, const int64_t* d_adjust, void* const* h_peer_blocks, int rank,
                                       int world, int64_t seq, rsx_stream_t stream) {
    RSX_REQUIRE(d_state && d_acc && D >= 1 && D <= KM_MAXD && h_peer_blocks, "rsx_kmeans_update_peers: bad arguments");
    RSX_REQUIRE(world >= 2 && world <= RSX_MAX_PEERS && rank >= 0 && rank < world && seq >= 1, "rsx_kmeans_update_peers: needs 2..%d ranks, seq >= 1",
                RSX_MAX_PEERS);
    KmPeers pr;
    memset(&pr, 0, sizeof(pr));
    const int parity = (int)(seq & 1);
    for (int p = 0; p < world; ++p) {
        RSX_REQUIRE(h_peer_blocks[p], "rsx_kmeans_update_peers: missing block of rank %d", p);
        char* base = reinterpret_cast<char*>(h_peer_blocks[p]);
        pr.pass[p] = reinterpret_cast<const long long*>(base) + (size_t)parity * RSX_PEER_PASS_ELEMS;
        pr.flag_out[p] = reinterpret_cast<unsigned long long*>(base + 2 * RSX_PEER_PASS_ELEMS * 8) + rank;
    }
    char* own = reinterpret_cast<char*>(h_peer_blocks[rank]);
    pr.flag_in = reinterpret_cast<const unsigned long long*>(own + 2 * RSX_PEER_PASS_ELEMS * 8);
    pr.zero_next = reinterpret_cast<long long*>(own) + (size_t)(parity ^ 1) * RSX_PEER_PASS_ELEMS;
    pr.seq = (unsigned long long)seq, pr.rank = rank, pr.world = world;
    pr.timeout_ns = (unsigned long long)max(1, rsx_option("peer_timeout_ms", 4000)) * 1000000ull;
    cudaStream_t s = (cudaStream_t)stream;
    km_update_kernel<true><<<1, KM_CTRL_THREADS, 0, s>>>((KmState*)d_state, reinterpret_cast<long long*>(d_acc), delta,
                                                         reinterpret_cast<const long long*>(d_adjust), pr);
    if (int rc = rsx_check_launch("km_update_peers")) return rc;
    return km_publish(d_state, D, s);
}

// fixed-point scale 2^shift_d of every feature (what the assign kernels multiply a raw sample by before rounding)
extern "C" int rsx_kmeans_fixed_point_scales(const void* d_state, double* h_pow2, rsx_stream_t stream) {
    RSX_REQUIRE(d_state && h_pow2, "rsx_kmeans_fixed_point_scales: bad arguments");
    static thread_local KmState h;
    if (int rc = rsx_fetch_small(&h, d_state, sizeof(h), (cudaStream_t)stream)) return rc;
    for (int d = 0; d < h.D; ++d) h_pow2[d] = (double)h.pow2[d];
    return RSX_OK;
}

extern "C" int rsx_kmeans_read(const void* d_state, double* h_centroids, double* h_shift_sq, int32_t* h_empty, rsx_stream_t stream) {
    return rsx_kmeans_read_all(d_state, h_centroids, h_shift_sq, h_empty, nullptr, nullptr, stream);
}

// + the per-feature raw range the state's scaling was derived from (what MinMaxScaler.fit saw), in the same fetch
extern "C" int rsx_kmeans_read_all(const void* d_state, double* h_centroids, double* h_shift_sq, int32_t* h_empty, double* h_feat_min,
                                   double* h_feat_max, rsx_stream_t stream) {
    RSX_REQUIRE(d_state, "rsx_kmeans_read: bad arguments");
    static thread_local KmState h;
    if (int rc = rsx_fetch_small(&h, d_state, sizeof(h), (cudaStream_t)stream)) return rc;
    for (int d = 0; d < h.D; ++d) {
        if (h_feat_min) h_feat_min[d] = h.fmin64[d];
        if (h_feat_max) h_feat_max[d] = h.fmax64[d];
    }
    if (h_centroids)
        for (int j = 0; j < h.K; ++j)
            for (int d = 0; d < h.D; ++d) h_centroids[j * h.D + d] = h.cent64[j * KM_MAXD + d] + h.mean64[d];
    if (h.n_empty < -500000) {
        rsx_set_error("rsx_kmeans_read: a peer rank never reached the update barrier (rsx_kmeans_update_peers timed out)");
        return RSX_ERR_CUDA;
    }
    if (h_shift_sq) *h_shift_sq = h.shift_sq;
    if (h_empty) *h_empty = h.n_empty;
    return RSX_OK;
}
