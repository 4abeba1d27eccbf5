// K5 assign kernels: one pass over the feature stack = distances, argmin, partial sums, labels, inertia.
// Compiled once per range of D (-DRSX_KM_PART=n, see rsx_kmeans_state.cuh); the host API lives in rsx_kmeans.cu.
//
// HBM layout: the feature stack is planar float32, D planes of n_px; one pass reads 4*D bytes/pixel (+1 B/px of previous
// labels and 1 B/px of new labels when labels are tracked); centroids live in constant memory and reach FFMA2 as
// uniform-register operands.
//
// Exactness: the parity target is sklearn run on the float64 promotion of the same stack.  The fast path evaluates the K
// distances in fp32 and keeps the best and second best; when their gap is below a rigorous rounding bound tau (computed
// per update from the centroid magnitudes) the pixel is re-evaluated in float64 with sklearn's operation order.  Partial
// sums are int64 fixed point (x * 2^shift_d, rounded once per sample), so they are associative: any tiling, any number of
// GPUs and any atomic ordering give bit-identical sums, hence bit-identical centroids.  For the same reason a DELTA pass
// (only the pixels whose label changed move their sample from one cluster's sum to another's) reproduces exactly the
// sums of a full pass.
#include <cstdio>
#include <cstdlib>

#include "rsx_kmeans_state.cuh"

#ifndef RSX_KM_PART
#error "compile with -DRSX_KM_PART=<0..5>"
#endif

__constant__ KmState g_km;  // refreshed (device-to-device) after every setup/update, per translation unit

// ----------------------------------------------------------------------------- exact (float64) re-evaluation
// Cold path: re-loads the pixel's features (L2 hits) so the hot loop keeps nothing in local memory.
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

// K > 8: the same evaluation by the whole warp - lane j takes centroids j and j + 32 - so that a near tie costs ~100 warp
// instructions instead of a K x D loop in one lane with the other 31 idle.  Same operations in the same order per centroid as
// km_exact_argmin, and the same winner: the smallest value, the lowest index among equals.  c64s: shared copy of the centred
// centroids [d][KP64] (KP64 = K rounded up to 32) followed by their squared norms [KP64].  Called by all 32 lanes.
template <int D>
__device__ __forceinline__ int km_exact_argmin_warp(const float* __restrict__ stack, int64_t plane_stride, int64_t p, int K, const double* __restrict__ c64s,
                                                    double* dist_out) {
    const int lane = threadIdx.x & 31;
    const int KP64 = (K + 31) & ~31;
    double X[D];
#pragma unroll
    for (int d = 0; d < D; ++d)
        X[d] = __dsub_rn(__dadd_rn(__dmul_rn((double)__ldg(stack + d * plane_stride + p), g_km.scale64[d]), g_km.min64[d]), g_km.mean64[d]);
    double xx = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) xx = fma(X[d], X[d], xx);
    double best = INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < K; j += 32) {
        double dot = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) dot = fma(X[d], c64s[d * KP64 + j], dot);
        const double v = fma(-2.0, dot, c64s[D * KP64 + j]);
        if (v < best) best = v, bi = j;  // j ascends within the lane: the first minimum stays
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, best, o);
        const int j2 = __shfl_xor_sync(0xffffffffu, bi, o);
        if (v2 < best || (v2 == best && j2 < bi)) best = v2, bi = j2;
    }
    if (bi >= K) bi = 0;  // every distance NaN (a NaN sample): the single-lane loop answers 0 as well
    *dist_out = fmax(xx + best, 0.0);
    return bi;
}

// ----------------------------------------------------------------------------- fp32 distances + tagged argmin
#define KM_ARGMIN_STEP(A, B_, S_, I_, J)          \
    S_ = fminf(S_, fmaxf(A, B_));                 \
    if (A < B_) B_ = A, I_ = J;

// Index-in-mantissa argmin: the low BITS bits of the fp32 distance are replaced by the centroid index, so best/runner-up
// tracking is three FMNMX and no index bookkeeping.  The perturbation (< 2^BITS ulp) is part of the near-tie bound tau;
// anything closer than tau is decided in float64 anyway, so fp32 ties never pick a label.
__device__ __forceinline__ float km_tag(float a, unsigned keep_mask, int j) { return __uint_as_float((__float_as_uint(a) & keep_mask) | (unsigned)j); }
#define KM_ARGMIN_TAGGED(A, B_, S_, J)            \
    {                                             \
        const float t_ = km_tag(A, KEEP, J);      \
        S_ = fminf(S_, fmaxf(t_, B_));            \
        B_ = fminf(B_, t_);                       \
    }

// tag width: 3 bits for the unrolled K <= 8 path; 4 / 5 / 6 bits for K <= 16 / 32 / 64 (g_km.tag_bits, set with tau)

constexpr int KM_THREADS = 128;
constexpr int KM_SLOTS = 8;
constexpr int KM_CHUNK = 4;  // centroids per inner-loop chunk of the K > 8 path (= one LDS.128 of weights per feature)

// best / second-best tagged distances of the thread's 4 pixels (v[d] = feature d of pixels p..p+3).
// KU = 8: eight centroid slots, fully unrolled, weights as constant-bank (uniform register) operands of FFMA2 (K <= 8;
// padding slots carry bias 1e30).  KU = 0: any K, chunks of 8 centroids with the weights in shared memory.
template <int D, int KU>
__device__ __forceinline__ void km_distances(const float4* v, int K, const float* __restrict__ wsm, float (&b)[4], float (&s)[4]) {
    const unsigned KEEP = KU == 8 ? ~7u : ~((1u << g_km.tag_bits) - 1u);
    b[0] = b[1] = b[2] = b[3] = INFINITY;
    s[0] = s[1] = s[2] = s[3] = INFINITY;
    if (KU > 0) {
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
            KM_ARGMIN_TAGGED(a01.x, b[0], s[0], j) KM_ARGMIN_TAGGED(a01.y, b[1], s[1], j)
            KM_ARGMIN_TAGGED(a23.x, b[2], s[2], j) KM_ARGMIN_TAGGED(a23.y, b[3], s[3], j)
        }
    } else {
        // K > 8: chunks of KM_CHUNK centroids, weights from shared memory (wsm = [chunk][D][KM_CHUNK] then bias [KP], KP = K rounded
        // up to 8; padding slots carry bias 1e30); one broadcast LDS.128 per feature and chunk, loaded one feature ahead of the
        // FFMA2s that consume it (with chunks of 8 the 128-register budget left no room for that and the FFMA2s waited on the
        // shared-memory scoreboard: 24 % of the stall samples of a K = 32 pass)
        const int KP = (K + 7) & ~7;
        const float* bias = wsm + D * KP;
#pragma unroll 1
        for (int jc = 0; jc < KP; jc += KM_CHUNK) {
            float2 a[KM_CHUNK][2];
            const float4* wc = reinterpret_cast<const float4*>(wsm + jc * D);  // this chunk's [D][4] block
            {
                const float4 bA = *reinterpret_cast<const float4*>(bias + jc);
                const float bb[4] = {bA.x, bA.y, bA.z, bA.w};
#pragma unroll
                for (int jj = 0; jj < KM_CHUNK; ++jj) a[jj][0] = a[jj][1] = make_float2(bb[jj], bb[jj]);
            }
            float4 wn = wc[0];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float4 wA = wn;
                if (d + 1 < D) wn = wc[d + 1];
                const float ww[4] = {wA.x, wA.y, wA.z, wA.w};
                const float2 x01 = make_float2(v[d].x, v[d].y), x23 = make_float2(v[d].z, v[d].w);
#pragma unroll
                for (int jj = 0; jj < KM_CHUNK; ++jj) {
                    a[jj][0] = __ffma2_rn(x01, make_float2(ww[jj], ww[jj]), a[jj][0]);
                    a[jj][1] = __ffma2_rn(x23, make_float2(ww[jj], ww[jj]), a[jj][1]);
                }
            }
#pragma unroll
            for (int jj = 0; jj < KM_CHUNK; ++jj) {
                KM_ARGMIN_TAGGED(a[jj][0].x, b[0], s[0], jc + jj) KM_ARGMIN_TAGGED(a[jj][0].y, b[1], s[1], jc + jj)
                KM_ARGMIN_TAGGED(a[jj][1].x, b[2], s[2], jc + jj) KM_ARGMIN_TAGGED(a[jj][1].y, b[3], s[3], jc + jj)
            }
        }
    }
}

// label of one pixel from its tagged best/second distances (+ float64 decision of near ties, + inertia)
// K > 8: the 6-bit index tags widen the near-tie band five-fold.  Before paying for float64, redo the K distances of this one
// pixel in fp32 WITHOUT tags (same FMA order, weights from shared memory) and apply the rounding-only bound tau_tight.
// (cold path: re-loads the pixel's features, an L2 hit, so that the hot loop keeps nothing in local memory)
template <int D>
__device__ __noinline__ bool km_recheck_fp32(const float* __restrict__ stack, int64_t plane_stride, int64_t p, int K, const float* __restrict__ wsm,
                                             int* bi_out) {
    const int KP = (K + 7) & ~7;
    float x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = stack[d * plane_stride + p];
    float b = INFINITY, s = INFINITY;
    int bi = 0;
    for (int j = 0; j < K; ++j) {
        float a = wsm[D * KP + j];
#pragma unroll
        for (int d = 0; d < D; ++d) a = fmaf(x[d], wsm[(j / KM_CHUNK) * (D * KM_CHUNK) + d * KM_CHUNK + (j % KM_CHUNK)], a);
        KM_ARGMIN_STEP(a, b, s, bi, j)
    }
    *bi_out = bi;
    return s - b > g_km.tau_tight;  // false also for NaN
}

template <int D, int KU, bool INERTIA>
__device__ __forceinline__ int km_decide(const float* __restrict__ stack, int64_t plane_stride, int64_t p, const float (&x)[D], float best, float second,
                                         const float* __restrict__ wsm, double& inertia, unsigned& ties) {
    int bi = (int)(__float_as_uint(best) & (KU == 8 ? 7u : ((1u << g_km.tag_bits) - 1u)));
    double dist_exact = -1.0;
    if (!(second - best > g_km.tau)) {  // near tie (or NaN): decide in float64
        bool decided = false;
        if (KU == 0) decided = km_recheck_fp32<D>(stack, plane_stride, p, g_km.K, wsm, &bi);
        if (!decided) {
            bi = km_exact_argmin<D>(stack, plane_stride, p, &dist_exact);
            ++ties;
        }
    }
    if (INERTIA) {
        if (dist_exact < 0.0) {
            // sum of squares of (x' - c): all terms positive, relative error ~D * 2^-24
            float dd = 0.f;
            // the centroid comes from shared memory (lanes hold different labels: a constant-bank read would replay once per label)
            const float* cent = KU == 0 ? wsm + (D + 1) * ((g_km.K + 7) & ~7) + bi * D : wsm + bi * D;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                float df = fmaf(x[d], g_km.scale32[d], g_km.off32[d]) - cent[d];
                dd = fmaf(df, df, dd);
            }
            dist_exact = (double)dd;
        }
        inertia += dist_exact;
    }
    return bi;
}

// scalar path for the ragged tail (n_px % 4 pixels)
template <int D>
__device__ __forceinline__ int km_scalar_pixel(const float* __restrict__ stack, int64_t plane_stride, int64_t q, int K, float (&x)[D], bool want_inertia,
                                               double& inertia, unsigned& ties) {
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
    double dist_exact = -1.0;
    if (!(s - b > g_km.tau)) {
        bi = km_exact_argmin<D>(stack, plane_stride, q, &dist_exact);
        ++ties;
    }
    if (want_inertia) {
        if (dist_exact < 0.0) {
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
    return bi;
}

__device__ __forceinline__ unsigned km_changed_mask(uint32_t diff) {
    return ((diff & 0xffu) != 0) | (((diff & 0xff00u) != 0) << 1) | (((diff & 0xff0000u) != 0) << 2) | (((diff & 0xff000000u) != 0) << 3);
}

__device__ __forceinline__ void km_commit_counters(long long* __restrict__ gacc, int K, int D, unsigned ties, unsigned changed, double inertia,
                                                   double* __restrict__ inertia_out, bool want_inertia) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ties += __shfl_xor_sync(0xffffffffu, ties, o);
        changed += __shfl_xor_sync(0xffffffffu, changed, o);
        if (want_inertia) inertia += __shfl_xor_sync(0xffffffffu, inertia, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (ties && gacc) atomicAdd(reinterpret_cast<unsigned long long*>(&gacc[K * D + K]), (unsigned long long)ties);
        if (changed && gacc) atomicAdd(reinterpret_cast<unsigned long long*>(&gacc[K * D + K + 1]), (unsigned long long)changed);
        if (want_inertia && inertia_out) atomicAdd(inertia_out, inertia);
    }
}

// ============================================================================= kernel A: full pass, K <= 8
// Sums from scratch with per-thread accumulators: every thread owns 8 slots (= labels) in shared memory, laid out
// [slot][feature][thread] as int64 so that a warp's accesses are 256 contiguous bytes (2 wavefronts, no bank conflicts, no
// atomics).  112 KB per CTA at D = 13, hence two CTAs per SM, a ping-pong register prefetch of the next row and an L2
// prefetch further ahead.  The flat pixel array is viewed as rows of row_len pixels; a tile is 512 columns x 32 rows; a
// thread owns 4 adjacent columns (one float4 per plane) and walks down the tile rows (CTA-uniform walk).
constexpr int KM_TILE_W = KM_THREADS * 4;
constexpr int KM_TILE_R = 32;

struct KmWalk {
    int64_t v_rows, tiles_total, tile, r, r_end, n4;
    int tiles_x, row_len, col;
    __device__ __forceinline__ bool open_tile() {
        if (tile >= tiles_total) return false;
        const int tx = (int)(tile % tiles_x);
        const int64_t ty = tile / tiles_x;
        col = tx * KM_TILE_W + threadIdx.x * 4;
        r = ty * KM_TILE_R;
        r_end = min(v_rows, r + KM_TILE_R);
        return true;
    }
    __device__ __forceinline__ bool start(int64_t n4_, int row_len_) {
        n4 = n4_, row_len = row_len_;
        v_rows = (n4 + row_len - 1) / row_len;
        tiles_x = (row_len + KM_TILE_W - 1) / KM_TILE_W;
        tiles_total = ((v_rows + KM_TILE_R - 1) / KM_TILE_R) * tiles_x;
        tile = blockIdx.x;
        return open_tile();
    }
    __device__ __forceinline__ bool advance() {
        if (++r < r_end) return true;
        tile += gridDim.x;
        return open_tile();
    }
    __device__ __forceinline__ int64_t pos() const { return r * row_len + col; }
    __device__ __forceinline__ bool valid() const { return col < row_len && r * row_len + col < n4; }
};

template <int D>
struct KmSmem {
    static constexpr int CELLS = KM_SLOTS * (D + 1);
    static constexpr int CACHE_BYTES = CELLS * KM_THREADS * 8;
};

template <int D>
__global__ void __launch_bounds__(KM_THREADS, 2) km_full_kernel(const float* __restrict__ stack, int64_t plane_stride, int64_t n_px, int row_len,
                                                                long long* __restrict__ gacc, uint8_t* __restrict__ lab8, const uint8_t* __restrict__ prev8,
                                                                int32_t* __restrict__ lab32, int pf_rows) {
    extern __shared__ __align__(16) unsigned char km_smem[];
    const int K = g_km.K;
    long long* cache = reinterpret_cast<long long*>(km_smem) + threadIdx.x;  // [slot][d][thread], this thread's column pre-applied
    for (int i = threadIdx.x; i < KmSmem<D>::CELLS * KM_THREADS; i += KM_THREADS) reinterpret_cast<long long*>(km_smem)[i] = 0;
    __syncthreads();
    double inertia = 0.0;
    unsigned ties = 0, changed = 0;
    const int64_t n4 = n_px & ~(int64_t)3;
    auto accumulate = [&](int label, const float (&x)[D]) {
#pragma unroll
        for (int d = 0; d < D; ++d) cache[(label * (D + 1) + d) * KM_THREADS] += __float2ll_rn(x[d] * g_km.pow2[d]);
        cache[(label * (D + 1) + D) * KM_THREADS] += 1;
    };
    float4 va[D], vb[D];  // ping-pong register sets, one is consumed while the other is being loaded
    uint32_t pva = 0xffffffffu, pvb = 0xffffffffu;
    auto load_row = [&](float4* dst, uint32_t& pv, int64_t at) {
#pragma unroll
        for (int d = 0; d < D; ++d) dst[d] = ldg_stream4(stack + d * plane_stride + at);
        if (prev8) pv = __ldg(reinterpret_cast<const uint32_t*>(prev8 + at));
    };
    const int lane_pf = threadIdx.x & 31;
    const bool do_pf = pf_rows > 0 && ((lane_pf & 7) == 0 || lane_pf == 31);
    auto process_row = [&](const float4* v, uint32_t pv, int64_t p) {
        if (do_pf && p + pf_rows * (int64_t)row_len + 4 <= n4) {
#pragma unroll
            for (int d = 0; d < D; ++d) asm volatile("prefetch.global.L2 [%0];" ::"l"(stack + d * plane_stride + p + pf_rows * (int64_t)row_len));
        }
        float b[4], s[4];
        km_distances<D, 8>(v, K, nullptr, b, s);
        int l[4];
        float x[D];
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = v[d].x;
        l[0] = km_decide<D, 8, false>(stack, plane_stride, p, x, b[0], s[0], nullptr, inertia, ties);
        accumulate(l[0], x);
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = v[d].y;
        l[1] = km_decide<D, 8, false>(stack, plane_stride, p + 1, x, b[1], s[1], nullptr, inertia, ties);
        accumulate(l[1], x);
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = v[d].z;
        l[2] = km_decide<D, 8, false>(stack, plane_stride, p + 2, x, b[2], s[2], nullptr, inertia, ties);
        accumulate(l[2], x);
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = v[d].w;
        l[3] = km_decide<D, 8, false>(stack, plane_stride, p + 3, x, b[3], s[3], nullptr, inertia, ties);
        accumulate(l[3], x);
        const uint32_t packed = (uint32_t)l[0] | ((uint32_t)l[1] << 8) | ((uint32_t)l[2] << 16) | ((uint32_t)l[3] << 24);
        if (prev8) changed += __popc(km_changed_mask(packed ^ pv));  // sklearn's strict-convergence test (_kmeans.py:723)
        if (lab8) *reinterpret_cast<uint32_t*>(lab8 + p) = packed;
        if (lab32) *reinterpret_cast<int4*>(lab32 + p) = make_int4(l[0], l[1], l[2], l[3]);
    };
    KmWalk walk;
    bool more = walk.start(n4, row_len);
    int64_t pa = 0, pb = 0;
    bool oka = false, okb = false;
    if (more) {
        pa = walk.pos(), oka = walk.valid();
        if (oka) load_row(va, pva, pa);
    }
    while (more) {
        more = walk.advance();
        if (more) {
            pb = walk.pos(), okb = walk.valid();
            if (okb) load_row(vb, pvb, pb);
        }
        if (oka) process_row(va, pva, pa);
        if (!more) break;
        more = walk.advance();
        if (more) {
            pa = walk.pos(), oka = walk.valid();
            if (oka) load_row(va, pva, pa);
        }
        if (okb) process_row(vb, pvb, pb);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t q = n4; q < n_px; ++q) {
            float x[D];
            const int l = km_scalar_pixel<D>(stack, plane_stride, q, K, x, false, inertia, ties);
            accumulate(l, x);
            if (prev8) changed += prev8[q] != l;
            if (lab8) lab8[q] = (uint8_t)l;
            if (lab32) lab32[q] = l;
        }
    }
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
    km_commit_counters(gacc, K, D, ties, changed, 0.0, nullptr, false);
}

// ============================================================================= kernel B: TMA-staged streaming pass
// The flat pixel array is cut into blocks of 512 pixels; CTA b handles blocks b, b + grid, ...  One elected thread stages
// each block's D plane segments (2 KB each) into a ring in shared memory with bulk asynchronous copies (TMA engine, no
// tensor map) that complete on the stage's "full" mbarrier; n_stages blocks are in flight per CTA, so HBM latency is
// decoupled from registers and the kernel runs at up to four CTAs per SM.  Every thread takes its 4 pixels of a block
// (one LDS.128 per plane).  The warps of a CTA are NOT kept in lockstep: a warp that is done with a stage arrives on the
// stage's "empty" mbarrier, and the last warp out (it sees the phase complete; a ticket makes it unique) refills the stage.
//   KM_ASSIGN  labels (+ inertia)
//   KM_DELTA   labels + the pixels whose label differs from the previous pass move their fixed-point sample between the
//              clusters' sums: the warp handles its changed pixels together - lane d reads feature d of the pixel from the
//              staged block, converts it and updates the WARP's private [K][D+1] int64 accumulator in shared memory (plain
//              read-modify-write: lanes own distinct cells, control flow is warp-uniform); lane D moves the count.
//   KM_FULL    as KM_DELTA with every pixel "moving in" from nowhere (used for K > 8, where per-thread accumulators do
//              not fit in shared memory)
constexpr int KM_WARPS = KM_THREADS / 32;
constexpr int KM_BLOCK_PX = KM_THREADS * 4;
// staged planes are 4 floats apart from a whole number of bank rows: the delta update reads one pixel of every plane with lane =
// feature (a 14-way bank conflict at a stride of 512 floats, 12 extra wavefronts per moved pixel; 2-way at 516)
constexpr int KM_STAGE_STRIDE = KM_BLOCK_PX + 4;

// Delta update of the sums: the pixels of this warp's 128 whose label changed move their sample from the old cluster's sums to the
// new one's: lane = feature (lane D = the count), one warp-uniform iteration per moved pixel (loops over the set bits, nothing
// predicated off).  (A shared-memory list of the moved pixels walked by two half-warps was measured: its 2 KB cost the fourth
// CTA per SM, 0.46 -> 0.585 ms per late pass.)
template <int D>
__device__ __forceinline__ void km_move_samples(const float* __restrict__ st, int warp_px, unsigned cm, uint32_t old_packed, uint32_t new_packed,
                                                long long* __restrict__ wacc, float my_pow2) {
    unsigned b = __ballot_sync(0xffffffffu, cm != 0);
    const int lane = threadIdx.x & 31;
    while (b) {
        const int src = __ffs(b) - 1;
        b &= b - 1;
        unsigned m = __shfl_sync(0xffffffffu, cm, src);
        const uint32_t ov = __shfl_sync(0xffffffffu, old_packed, src), nv = __shfl_sync(0xffffffffu, new_packed, src);
        if (lane <= D) {
            do {  // warp uniform: m is the same in every lane
                const int i = __ffs(m) - 1;
                m &= m - 1;
                const int from = (int)((ov >> (8 * i)) & 0xffu), to = (int)((nv >> (8 * i)) & 0xffu);
                long long q = 1;
                if (lane < D) q = __float2ll_rn(st[lane * KM_STAGE_STRIDE + warp_px + 4 * src + i] * my_pow2);
                wacc[to * (D + 1) + lane] += q;
                if (from < KM_MAXK) wacc[from * (D + 1) + lane] -= q;
            } while (m);
        }
    }
}

// KM_FULL (K > 8): every pixel moves in, nothing moves out.  Lane d takes the four samples of an owning thread with one load,
// converts them, and adds runs of equal labels together before touching the accumulator (neighbouring pixels mostly share
// their label), so a thread's four pixels cost one or two read-modify-writes instead of four dependent ones.
template <int D>
__device__ __forceinline__ void km_move_in_all(const float* __restrict__ st, int warp_px, unsigned valid_mask, uint32_t new_packed,
                                               long long* __restrict__ wacc, float my_pow2) {
    const int lane = threadIdx.x & 31;
    const float* mine = st + min(lane, D - 1) * KM_STAGE_STRIDE + warp_px;
#pragma unroll 2
    for (int src = 0; src < 32; ++src) {
        if (!((valid_mask >> src) & 1u)) break;  // valid threads are a prefix of the warp
        const uint32_t nv = __shfl_sync(0xffffffffu, new_packed, src);
        if (lane <= D) {
            const float4 xv = *reinterpret_cast<const float4*>(mine + 4 * src);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
            long long q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = lane < D ? __float2ll_rn(xs[i] * my_pow2) : 1ll;
            unsigned cur = nv & 0xffu;
            long long run = q[0];
#pragma unroll
            for (int i = 1; i < 4; ++i) {
                const unsigned t = (nv >> (8 * i)) & 0xffu;
                if (t == cur) {
                    run += q[i];
                } else {
                    wacc[cur * (D + 1) + lane] += run;
                    cur = t, run = q[i];
                }
            }
            wacc[cur * (D + 1) + lane] += run;
        }
    }
}

// ---- 16-bit screening passes (QIN): the stage holds the uint16 copy of the stack (half the bytes of a pass); a sample reads back
// as x~ = fma(2^23 + u, step, off) - the float whose low mantissa bits are u, built with one PRMT - and its distances are within
// KmState::tau_q of the exact ones.  A pixel whose two nearest centroids are closer than that is evaluated from the float32 planes
// (km_uncertain_pixel: fp32 distances, float64 inside their own near-tie band); the pixels that move fetch their float32 sample
// for the integer sums.  Labels, sums, counters: those of the float32 pass.
constexpr int KM_QSTRIDE = KM_BLOCK_PX + 8;  // uint16 elements between staged planes (16-byte aligned rows, 4 banks apart)
__device__ __forceinline__ float km_q_lo(uint32_t w, float step, float off) { return fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)), step, off); }
__device__ __forceinline__ float km_q_hi(uint32_t w, float step, float off) { return fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)), step, off); }

template <int D>
__device__ __noinline__ int km_uncertain_pixel(const float* __restrict__ stack, int64_t plane_stride, int64_t p, unsigned* ties_out) {
    float x[D];
    double inertia = 0.0;
    unsigned ties = 0;
    const int l = km_scalar_pixel<D>(stack, plane_stride, p, g_km.K, x, false, inertia, ties);
    *ties_out += ties;
    return l;
}

// km_move_samples with the samples fetched from the float32 planes (the stage holds the 16-bit copy)
template <int D>
__device__ __forceinline__ void km_move_samples_global(const float* __restrict__ stack, int64_t plane_stride, int64_t warp_p0, unsigned cm,
                                                       uint32_t old_packed, uint32_t new_packed, long long* __restrict__ wacc, float my_pow2) {
    unsigned b = __ballot_sync(0xffffffffu, cm != 0);
    const int lane = threadIdx.x & 31;
    const float* mine = stack + (int64_t)min(lane, D - 1) * plane_stride + warp_p0;
    while (b) {
        const int src = __ffs(b) - 1;
        b &= b - 1;
        unsigned m = __shfl_sync(0xffffffffu, cm, src);
        const uint32_t ov = __shfl_sync(0xffffffffu, old_packed, src), nv = __shfl_sync(0xffffffffu, new_packed, src);
        if (lane <= D) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(mine + 4 * src));  // the owning thread's four pixels of this feature
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (m & (1u << i)) {
                    const int from = (int)((ov >> (8 * i)) & 0xffu), to = (int)((nv >> (8 * i)) & 0xffu);
                    const long long q = lane < D ? __float2ll_rn(xs[i] * my_pow2) : 1ll;
                    wacc[to * (D + 1) + lane] += q;
                    if (from < KM_MAXK) wacc[from * (D + 1) + lane] -= q;
                }
            }
        }
    }
}

template <int D, int MODE, bool INERTIA, int KU, bool WARPX, bool QIN = false>
__global__ void __launch_bounds__(KM_THREADS, km_ctas_per_sm(D)) km_stream_kernel(const float* __restrict__ stack, int64_t plane_stride, int64_t n_px,
                                                                  long long* __restrict__ gacc, uint8_t* __restrict__ lab8,
                                                                  const uint8_t* __restrict__ prev8, int32_t* __restrict__ lab32,
                                                                  double* __restrict__ inertia_out, int n_stages, int c64_offset,
                                                                  const uint16_t* __restrict__ q16, int64_t q_stride) {
    constexpr bool SUMS = MODE != KM_ASSIGN;
    constexpr int STAGE_BYTES = QIN ? D * KM_QSTRIDE * 2 : D * KM_STAGE_STRIDE * 4;
    static_assert(!QIN || (KU == 8 && MODE == KM_DELTA && !INERTIA), "the 16-bit screening pass is a K <= 8 delta pass");
    constexpr bool LOCKSTEP = false;  // true: one CTA barrier per block instead of the empty mbarriers (kept for experiments)
    extern __shared__ __align__(128) unsigned char km_smem[];
    const int K = g_km.K;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* stages = reinterpret_cast<float*>(km_smem);                                                          // [n_stages][D][512]
    long long* wacc_all = reinterpret_cast<long long*>(km_smem + (size_t)n_stages * STAGE_BYTES);                  // [4 warps][K][D+1]
    uint64_t* full = reinterpret_cast<uint64_t*>(wacc_all + (SUMS ? KM_WARPS * K * (D + 1) : 0));               // [n_stages]
    uint64_t* empty = full + n_stages;                                                                          // [n_stages]
    int* ticket = reinterpret_cast<int*>(empty + n_stages);                                                     // [n_stages] (+ pad): refills issued per stage
    float* wsm = reinterpret_cast<float*>(empty + n_stages) + 4;                                                    // KU == 0: [D][KP] weights, [KP] biases
    long long* wacc = wacc_all + warp * K * (D + 1);
    // KU == 0: float64 centroids for the warp-wide near-tie evaluation, [D + 1][K rounded up to 32], after the fp32 tables
    double* c64s = reinterpret_cast<double*>(km_smem + c64_offset);
    if (KU == 0) {
        const int KP = (K + 7) & ~7;
        for (int i = tid; i < D * KP; i += KM_THREADS) {  // [chunk][D][KM_CHUNK]
            const int c = i / (D * KM_CHUNK), r = i - c * (D * KM_CHUNK), d = r / KM_CHUNK, j = c * KM_CHUNK + (r % KM_CHUNK);
            wsm[i] = g_km.w32[j * KM_MAXD + d];
        }
        for (int j = tid; j < KP; j += KM_THREADS) wsm[D * KP + j] = g_km.bias32[j];
        if (INERTIA)
            for (int i = tid; i < K * D; i += KM_THREADS) wsm[(D + 1) * KP + i] = g_km.cent32[(i / D) * KM_MAXD + i % D];
        if (WARPX) {
            const int KP64 = (K + 31) & ~31;
            for (int i = tid; i < D * KP64; i += KM_THREADS) {
                const int d = i / KP64, j = i - d * KP64;
                c64s[i] = j < K ? g_km.cent64[j * KM_MAXD + d] : 0.0;
            }
            for (int j = tid; j < KP64; j += KM_THREADS) c64s[D * KP64 + j] = j < K ? g_km.cnorm64[j] : 0.0;
        }
    }
    if (KU == 8 && INERTIA)  // K <= 8, final pass: the centroids for the inertia term, [K][D]
        for (int i = tid; i < K * D; i += KM_THREADS) wsm[i] = g_km.cent32[(i / D) * KM_MAXD + i % D];
    const int64_t n4 = n_px & ~(int64_t)3;
    const int64_t n_blocks = (n4 + KM_BLOCK_PX - 1) / KM_BLOCK_PX;

    if (SUMS)
        for (int i = tid; i < KM_WARPS * K * (D + 1); i += KM_THREADS) wacc_all[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], KM_WARPS), ticket[s] = 0;
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int64_t blk, int s) {  // one thread: stage block blk
        const int64_t p0 = blk * KM_BLOCK_PX;
        const unsigned npx = (unsigned)min((int64_t)KM_BLOCK_PX, n4 - p0);
        // bulk copies move multiples of 16 bytes: the 16-bit planes are padded to a multiple of 8 samples
        const unsigned bytes = QIN ? ((npx + 7u) & ~7u) * 2u : npx * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of this stage precede the async refill
        mbar_expect_tx(&full[s], bytes * D);
        if (QIN) {
            uint16_t* dst = reinterpret_cast<uint16_t*>(km_smem + (size_t)s * STAGE_BYTES);
#pragma unroll 1
            for (int d = 0; d < D; ++d) bulk_g2s(dst + d * KM_QSTRIDE, q16 + d * q_stride + p0, bytes, &full[s]);
            return;
        }
        float* dst = stages + (size_t)s * D * KM_STAGE_STRIDE;
#pragma unroll 1
        for (int d = 0; d < D; ++d) bulk_g2s(dst + d * KM_STAGE_STRIDE, stack + d * plane_stride + p0, bytes, &full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < n_stages; ++s) {
            const int64_t blk = blockIdx.x + (int64_t)s * gridDim.x;
            if (blk < n_blocks) issue(blk, s);
        }
    const float my_pow2 = SUMS ? g_km.pow2[min(lane, D - 1)] : 0.f;
    double inertia = 0.0;
    unsigned ties = 0, changed = 0;
    uint32_t pv_next = 0xffffffffu;
    if (prev8 && (int64_t)blockIdx.x * KM_BLOCK_PX + 4 * tid < n4) pv_next = __ldg(reinterpret_cast<const uint32_t*>(prev8 + (int64_t)blockIdx.x * KM_BLOCK_PX + 4 * tid));

    int s = 0, use = 0;  // stage of the current block, how many times that stage has been used before
    unsigned parity = 0;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int64_t p = blk * KM_BLOCK_PX + 4 * tid;
        const bool valid = p < n4;
        const uint32_t pv = pv_next;
        {
            const int64_t pn = p + (int64_t)gridDim.x * KM_BLOCK_PX;  // previous labels of the next block: in flight during this one
            pv_next = 0xffffffffu;
            if (prev8 && pn < n4) pv_next = __ldg(reinterpret_cast<const uint32_t*>(prev8 + pn));
        }
        mbar_wait(&full[s], parity);
        const float* st = stages + (size_t)s * D * KM_STAGE_STRIDE;
        uint32_t packed = 0, diff = 0;
        if (KU == 8) {
            if (valid && QIN) {
                const uint16_t* sq = reinterpret_cast<const uint16_t*>(km_smem + (size_t)s * STAGE_BYTES) + 4 * tid;
                float4 v[D];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const uint2 r = *reinterpret_cast<const uint2*>(sq + d * KM_QSTRIDE);
                    const float qs = g_km.qstep32[d], qo = g_km.qoff32[d];
                    v[d] = make_float4(km_q_lo(r.x, qs, qo), km_q_hi(r.x, qs, qo), km_q_lo(r.y, qs, qo), km_q_hi(r.y, qs, qo));
                }
                float b[4], sc[4];
                km_distances<D, KU>(v, K, wsm, b, sc);
                int l[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    l[i] = (int)(__float_as_uint(b[i]) & 7u);
                    if (!(sc[i] - b[i] > g_km.tau_q)) l[i] = km_uncertain_pixel<D>(stack, plane_stride, p + i, &ties);
                }
                packed = (uint32_t)l[0] | ((uint32_t)l[1] << 8) | ((uint32_t)l[2] << 16) | ((uint32_t)l[3] << 24);
                diff = packed ^ pv;
                *reinterpret_cast<uint32_t*>(lab8 + p) = packed;
            } else if (valid) {
                float4 v[D];
#pragma unroll
                for (int d = 0; d < D; ++d) v[d] = *reinterpret_cast<const float4*>(st + d * KM_STAGE_STRIDE + 4 * tid);
                float b[4], sc[4];
                km_distances<D, KU>(v, K, wsm, b, sc);
                int l[4];
                float x[D];
#pragma unroll
                for (int d = 0; d < D; ++d) x[d] = v[d].x;
                l[0] = km_decide<D, KU, INERTIA>(stack, plane_stride, p, x, b[0], sc[0], wsm, inertia, ties);
#pragma unroll
                for (int d = 0; d < D; ++d) x[d] = v[d].y;
                l[1] = km_decide<D, KU, INERTIA>(stack, plane_stride, p + 1, x, b[1], sc[1], wsm, inertia, ties);
#pragma unroll
                for (int d = 0; d < D; ++d) x[d] = v[d].z;
                l[2] = km_decide<D, KU, INERTIA>(stack, plane_stride, p + 2, x, b[2], sc[2], wsm, inertia, ties);
#pragma unroll
                for (int d = 0; d < D; ++d) x[d] = v[d].w;
                l[3] = km_decide<D, KU, INERTIA>(stack, plane_stride, p + 3, x, b[3], sc[3], wsm, inertia, ties);
                packed = (uint32_t)l[0] | ((uint32_t)l[1] << 8) | ((uint32_t)l[2] << 16) | ((uint32_t)l[3] << 24);
                if (prev8) diff = packed ^ pv;
                if (lab8) *reinterpret_cast<uint32_t*>(lab8 + p) = packed;
                if (lab32) *reinterpret_cast<int4*>(lab32 + p) = make_int4(l[0], l[1], l[2], l[3]);
            }
        } else {
            // K > 16: tagged fp32 argmin per thread; the pixels inside the near-tie band are then evaluated in float64 by the
            // whole warp, one after the other (warp-uniform loop over the ballot): 1.55 -> 1.43 ms per pass at K = 32.
            // 8 < K <= 16 keeps the per-lane path (untagged fp32 re-check, then float64): half the lanes would idle (+6 %).
            constexpr bool warpwide = WARPX;
            int l[4] = {0, 0, 0, 0};
            double dex[4] = {-1.0, -1.0, -1.0, -1.0};
            unsigned unsure = 0;
            if (valid) {
                float4 v[D];
#pragma unroll
                for (int d = 0; d < D; ++d) v[d] = *reinterpret_cast<const float4*>(st + d * KM_STAGE_STRIDE + 4 * tid);
                float b[4], sc[4];
                km_distances<D, KU>(v, K, wsm, b, sc);
                if (warpwide) {
                    const unsigned tag_mask = (1u << g_km.tag_bits) - 1u;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        l[i] = (int)(__float_as_uint(b[i]) & tag_mask);
                        if (!(sc[i] - b[i] > g_km.tau)) unsure |= 1u << i;  // also NaN
                    }
                } else {
                    float x[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) x[d] = v[d].x;
                    l[0] = km_decide<D, KU, INERTIA>(stack, plane_stride, p, x, b[0], sc[0], wsm, inertia, ties);
#pragma unroll
                    for (int d = 0; d < D; ++d) x[d] = v[d].y;
                    l[1] = km_decide<D, KU, INERTIA>(stack, plane_stride, p + 1, x, b[1], sc[1], wsm, inertia, ties);
#pragma unroll
                    for (int d = 0; d < D; ++d) x[d] = v[d].z;
                    l[2] = km_decide<D, KU, INERTIA>(stack, plane_stride, p + 2, x, b[2], sc[2], wsm, inertia, ties);
#pragma unroll
                    for (int d = 0; d < D; ++d) x[d] = v[d].w;
                    l[3] = km_decide<D, KU, INERTIA>(stack, plane_stride, p + 3, x, b[3], sc[3], wsm, inertia, ties);
                }
            }
            unsigned any = warpwide ? __ballot_sync(0xffffffffu, unsure != 0) : 0u;
            while (any) {
                const int src = __ffs(any) - 1;
                any &= any - 1;
                const unsigned um = __shfl_sync(0xffffffffu, unsure, src);
                const int64_t pw = blk * KM_BLOCK_PX + warp * 128 + 4 * src;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (um & (1u << i)) {  // warp-uniform
                        double de;
                        const int bi = km_exact_argmin_warp<D>(stack, plane_stride, pw + i, K, c64s, &de);
                        if (lane == src) l[i] = bi, dex[i] = de, ++ties;
                    }
                }
            }
            if (valid) {
                if (INERTIA && warpwide) {
                    const int KP = (K + 7) & ~7;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (dex[i] < 0.0) {  // sum of squares of (x' - c) in fp32: all terms positive, relative error ~D * 2^-24
                            const float* cent = wsm + (D + 1) * KP + l[i] * D;
                            float dd = 0.f;
#pragma unroll
                            for (int d = 0; d < D; ++d) {
                                const float df = fmaf(st[d * KM_STAGE_STRIDE + 4 * tid + i], g_km.scale32[d], g_km.off32[d]) - cent[d];
                                dd = fmaf(df, df, dd);
                            }
                            dex[i] = (double)dd;
                        }
                        inertia += dex[i];
                    }
                }
                packed = (uint32_t)l[0] | ((uint32_t)l[1] << 8) | ((uint32_t)l[2] << 16) | ((uint32_t)l[3] << 24);
                if (prev8) diff = packed ^ pv;
                if (lab8) *reinterpret_cast<uint32_t*>(lab8 + p) = packed;
                if (lab32) *reinterpret_cast<int4*>(lab32 + p) = make_int4(l[0], l[1], l[2], l[3]);
            }
        }
        unsigned cm = km_changed_mask(diff);
        changed += __popc(cm);  // sklearn's strict-convergence test (_kmeans.py:723)
        if (MODE == KM_FULL) cm = valid ? 0xfu : 0u;
        if (MODE == KM_FULL)
            km_move_in_all<D>(st, warp * 128, __ballot_sync(0xffffffffu, valid), packed, wacc, my_pow2);
        else if (SUMS && QIN)
            km_move_samples_global<D>(stack, plane_stride, blk * KM_BLOCK_PX + warp * 128, cm, pv, packed, wacc, my_pow2);
        else if (SUMS)
            km_move_samples<D>(st, warp * 128, cm, pv, packed, wacc, my_pow2);
        if (LOCKSTEP) {
            __syncthreads();  // everyone is done with stage s: refill it
            if (tid == 0) {
                const int64_t nb = blk + (int64_t)n_stages * gridDim.x;
                if (nb < n_blocks) issue(nb, s);
            }
        } else {
            // this warp has left stage s; the warp whose arrival completes the phase (the last one out) refills the stage
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty[s]);
                if (mbar_test(&empty[s], parity) && atomicCAS(&ticket[s], use, use + 1) == use) {
                    const int64_t nb = blk + (int64_t)n_stages * gridDim.x;
                    if (nb < n_blocks) issue(nb, s);
                }
            }
        }
        if (++s == n_stages) s = 0, parity ^= 1u, ++use;
    }
    // ragged tail (n_px % 4 pixels): one thread, scalar
    if (blockIdx.x == 0 && tid == 0) {
        for (int64_t q = n4; q < n_px; ++q) {
            float x[D];
            const int l = km_scalar_pixel<D>(stack, plane_stride, q, K, x, INERTIA, inertia, ties);
            const int old = (MODE == KM_FULL || !prev8) ? 255 : (int)prev8[q];
            if (prev8 && prev8[q] != l) ++changed;
            if (SUMS && (MODE == KM_FULL || old != l)) {  // warp 0's accumulator; its other lanes are past their loops
                for (int d = 0; d < D; ++d) {
                    const long long qv = __float2ll_rn(x[d] * g_km.pow2[d]);
                    wacc[l * (D + 1) + d] += qv;
                    if (old < KM_MAXK) wacc[old * (D + 1) + d] -= qv;
                }
                wacc[l * (D + 1) + D] += 1;
                if (old < KM_MAXK) wacc[old * (D + 1) + D] -= 1;
            }
            if (lab8) lab8[q] = (uint8_t)l;
            if (lab32) lab32[q] = l;
        }
    }
    if (SUMS) {
        __syncthreads();
        for (int i = tid; i < K * (D + 1); i += KM_THREADS) {
            long long t = 0;
#pragma unroll
            for (int w = 0; w < KM_WARPS; ++w) t += wacc_all[w * K * (D + 1) + i];
            if (t) {
                const int j = i / (D + 1), d = i % (D + 1);
                long long* dst = d < D ? &gacc[j * D + d] : &gacc[K * D + j];
                atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)t);
            }
        }
    }
    km_commit_counters(gacc, K, D, ties, changed, inertia, inertia_out, INERTIA);
}

// ============================================================================= kernel C: K > 8 distances on the tensor cores
// At K = 32, D = 13 a pass is 416 FMA per pixel: above the fp32 ridge of the B200, so kernel B runs at a quarter of the HBM rate
// there.  Here the K distances of a pixel come from tcgen05 (5th-generation tensor cores, accumulators in TMEM):
//     dist[128 px x N] = A[128 px x 16] * W[N x 16]^T,   columns 0..D-1 = the raw features, D..D+2 = 1.0 against the bias pieces
// in TF32 with a three-term split so that what the tensor core drops is bounded: x = xh + xl (xh = the 11 leading bits), w = wh +
// wl likewise, dist ~ xh.wh + xh.wl + xl.wh (six M128 N K8 instructions per 128-pixel tile, fp32 accumulation in TMEM).  The result
// carries an error of at most ~5 * 2^-20 of the distance magnitude (KmState::tau_tc, rsx_kmeans.cu); pixels whose two nearest
// centroids are closer than that band are decided in float64 by the whole warp (km_exact_argmin_warp), exactly as in kernel B, so
// the tensor cores never decide a label either.
//   * staging: the same ring of bulk-copied plane segments as kernel B (blocks of 384 pixels = three 128-pixel tiles);
//   * A operand: thread t of the CTA owns pixel t of the tile = TMEM lane t; it reads its D features from the staged block (LDS.32,
//     conflict free), splits them and writes xh / xl to 2 x 16 TMEM columns with tcgen05.st (no shared-memory copy of A);
//   * B operand (weights): K-major, no swizzle, in shared memory, built once per CTA from the state: element (n, k) at
//     (k/4) * (N/8) * 128 + (n/8) * 128 + (n%8) * 16 + (k%4) * 4 bytes (8 x 16 B core matrices; LBO = k-chunk stride,
//     SBO = 128: layout verified on the B200 by tools/tc_probe.cu);
//   * one elected thread issues the six MMAs and tcgen05.commit onto an mbarrier; every thread then reads its pixel's N distances
//     with tcgen05.ld and runs the tagged argmin / delta update of kernel B.  Four CTAs per SM cover each other's MMA round trip
//     (~300 cycles).  Sums: one int64 accumulator per CTA in shared memory (64-bit shared atomics), which is what lets four CTAs
//     fit next to the staging ring.
constexpr int KM_TC_SUBS = 3;
constexpr int KM_TC_BPX = 128 * KM_TC_SUBS;  // pixels per staged block
constexpr int KM_TC_KDIM = 16;               // D features + 3 bias slots <= 16

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&a)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(a[0]),
                 "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]),
                 "r"(a[13]), "r"(a[14]), "r"(a[15])
                 : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&d)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]),
                   "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                 : "r"(taddr)
                 : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle (start address, leading / stride byte offsets in 16-byte units, version 1)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) |
           ((uint64_t)1 << 46);
}
// D[tmem] (+)= A[tmem, 128 lanes x 8 columns tf32] * B[smem descriptor, N x 8]^T
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// 64-bit add on shared memory as two native 32-bit atomics (the carry is passed on by whoever produces it; additions commute, so
// the word pair is exact once all adds have landed).  atomicAdd on a 64-bit shared address compiles to a compare-and-swap loop.
__device__ __forceinline__ void km_smem_add64(long long* p, long long q) {
    unsigned* w = reinterpret_cast<unsigned*>(p);
    const unsigned lo = (unsigned)(unsigned long long)q, hi = (unsigned)((unsigned long long)q >> 32);
    const unsigned old = atomicAdd(w, lo);
    const unsigned carry = (old + lo) < old ? 1u : 0u;
    if (hi + carry) atomicAdd(w + 1, hi + carry);
}

struct KmTcSmem {  // byte offsets inside the dynamic shared memory of km_tc_kernel (host and device agree through this one function)
    int bop, c64, w32, wacc, cent, bars, total;
    __host__ __device__ KmTcSmem(int D, int K, int n_stages, int npad, bool sums, bool inertia) {
        const int KP64 = (K + 31) & ~31;
        int o = n_stages * D * KM_TC_BPX * 4;
        bop = o, o += 2 * npad * KM_TC_KDIM * 4;               // weights for the tensor core: hi, lo
        c64 = o, o += (D + 1) * KP64 * 8;                      // float64 centroids + norms (near ties, second stage)
        w32 = o, o += (D + 1) * KP64 * 4;                      // fp32 weights [d][j] + bias [j] (near ties, first stage)
        wacc = o, o += sums ? K * (D + 1) * 8 : 0;             // the CTA's int64 sums
        cent = o, o += inertia ? ((K * D * 4 + 15) & ~15) : 0;  // fp32 centroids (inertia)
        bars = o, o += (2 * n_stages + 4) * 8 + (n_stages + 2) * 4 + 16;
        total = (o + 15) & ~15;
    }
};

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&d)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]),
          "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]),
          "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
        : "r"(taddr)
        : "memory");
}

// Near tie, first stage: the K distances of ONE pixel in plain fp32 (FMA chain from the bias, no index tags) by the whole warp -
// lane j takes centroids j and j + 32 - from the staged block (xs: feature d at xs[d * stride]).  Returns true when the two
// nearest are farther apart than the rounding-only bound tau_tight; *bi_out = the nearest.
template <int D>
__device__ __forceinline__ bool km_recheck_fp32_warp(const float* __restrict__ xs, int stride, int K, const float* __restrict__ w32s, int* bi_out) {
    const int lane = threadIdx.x & 31;
    const int KP64 = (K + 31) & ~31;
    float b = INFINITY, sc = INFINITY;
    int bi = 0;
    for (int j = lane; j < K; j += 32) {
        float a = w32s[D * KP64 + j];
#pragma unroll
        for (int d = 0; d < D; ++d) a = fmaf(xs[d * stride], w32s[d * KP64 + j], a);
        KM_ARGMIN_STEP(a, b, sc, bi, j)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float b2 = __shfl_xor_sync(0xffffffffu, b, o), s2 = __shfl_xor_sync(0xffffffffu, sc, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
        sc = fminf(fminf(sc, s2), fmaxf(b, b2));  // second smallest of the union
        if (b2 < b) b = b2, bi = i2;
    }
    *bi_out = bi;
    return sc - b > g_km.tau_tight;  // false also for NaN
}

// Near tie, second stage: km_exact_argmin_warp with the pixel's features read from the staged block instead of global memory.
template <int D>
__device__ __forceinline__ int km_exact_argmin_warp_staged(const float* __restrict__ xs, int stride, int K, const double* __restrict__ c64s, double* dist_out) {
    const int lane = threadIdx.x & 31;
    const int KP64 = (K + 31) & ~31;
    double X[D];
#pragma unroll
    for (int d = 0; d < D; ++d) X[d] = __dsub_rn(__dadd_rn(__dmul_rn((double)xs[d * stride], g_km.scale64[d]), g_km.min64[d]), g_km.mean64[d]);
    double xx = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) xx = fma(X[d], X[d], xx);
    double best = INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < K; j += 32) {
        double dot = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) dot = fma(X[d], c64s[d * KP64 + j], dot);
        const double v = fma(-2.0, dot, c64s[D * KP64 + j]);
        if (v < best) best = v, bi = j;  // j ascends within the lane: the first minimum stays
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, best, o);
        const int j2 = __shfl_xor_sync(0xffffffffu, bi, o);
        if (v2 < best || (v2 == best && j2 < bi)) best = v2, bi = j2;
    }
    if (bi >= K) bi = 0;  // every distance NaN (a NaN sample): the single-lane loop answers 0 as well
    *dist_out = fmax(xx + best, 0.0);
    return bi;
}

template <int D, int MODE, bool INERTIA, int NPAD>
__global__ void __launch_bounds__(KM_THREADS, 4) km_tc_kernel(const float* __restrict__ stack, int64_t plane_stride, int64_t n_px,
                                                              long long* __restrict__ gacc, uint8_t* __restrict__ lab8, const uint8_t* __restrict__ prev8,
                                                              int32_t* __restrict__ lab32, double* __restrict__ inertia_out, int n_stages, int tmem_cols) {
    static_assert(D + 3 <= KM_TC_KDIM, "features + bias slots must fit the 16-wide K dimension");
    static_assert(NPAD == 16 || NPAD == 32 || NPAD == 64, "MMA N");
    constexpr bool SUMS = MODE != KM_ASSIGN;
    constexpr int BPX = KM_TC_BPX;
    extern __shared__ __align__(128) unsigned char km_smem[];
    const int K = g_km.K;
    const int KP64 = (K + 31) & ~31;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const KmTcSmem L(D, K, n_stages, NPAD, SUMS, INERTIA);
    float* stages = reinterpret_cast<float*>(km_smem);
    float* bop = reinterpret_cast<float*>(km_smem + L.bop);
    double* c64s = reinterpret_cast<double*>(km_smem + L.c64);
    float* w32s = reinterpret_cast<float*>(km_smem + L.w32);
    long long* wacc = reinterpret_cast<long long*>(km_smem + L.wacc);
    float* cent_s = reinterpret_cast<float*>(km_smem + L.cent);
    uint64_t* full = reinterpret_cast<uint64_t*>(km_smem + L.bars);
    uint64_t* empty = full + n_stages;
    uint64_t* mma_bar = empty + n_stages;  // [2]: the MMAs of TMEM buffer b have completed
    uint64_t* a_ready = mma_bar + 2;       // [2]: the four warps have written their lanes of A buffer b
    int* ticket = reinterpret_cast<int*>(a_ready + 2);  // [n_stages] refills issued per stage, then [2] MMA batches issued per buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ticket + n_stages + 2);

    // ---- weights: W[n][k] = w32[n][k] (k < D), bias in three 11-bit pieces at k = D, D+1, D+2; never-chosen rows carry bias 1e30
    constexpr int CSTR = (NPAD >> 3) * 128;  // bytes between k-chunks
    for (int i = tid; i < NPAD * KM_TC_KDIM; i += KM_THREADS) {
        const int n = i / KM_TC_KDIM, k = i - n * KM_TC_KDIM;
        float hi = 0.f, lo = 0.f;
        if (k < D) {
            const float w = n < K ? g_km.w32[n * KM_MAXD + k] : 0.f;
            hi = tf32_trunc(w), lo = tf32_trunc(w - hi);
        } else if (k < D + 3) {
            const float b = n < K ? g_km.bias32[n] : 1e30f;
            const float b0 = tf32_trunc(b), b1 = tf32_trunc(b - b0), b2 = tf32_trunc(b - b0 - b1);
            hi = k == D ? b0 : (k == D + 1 ? b1 : b2);
        }
        const int off = ((k >> 2) * CSTR + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4) >> 2;
        bop[off] = hi;
        bop[NPAD * KM_TC_KDIM + off] = lo;
    }
    for (int i = tid; i < D * KP64; i += KM_THREADS) {
        const int d = i / KP64, j = i - d * KP64;
        c64s[i] = j < K ? g_km.cent64[j * KM_MAXD + d] : 0.0;
        w32s[i] = j < K ? g_km.w32[j * KM_MAXD + d] : 0.f;
    }
    for (int j = tid; j < KP64; j += KM_THREADS) c64s[D * KP64 + j] = j < K ? g_km.cnorm64[j] : 0.0, w32s[D * KP64 + j] = j < K ? g_km.bias32[j] : 1e30f;
    if (INERTIA)
        for (int i = tid; i < K * D; i += KM_THREADS) cent_s[i] = g_km.cent32[(i / D) * KM_MAXD + i % D];
    if (SUMS)
        for (int i = tid; i < K * (D + 1); i += KM_THREADS) wacc[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], KM_WARPS), ticket[s] = 0;
        for (int b = 0; b < 2; ++b) mbar_init(&mma_bar[b], 1), mbar_init(&a_ready[b], KM_WARPS), ticket[n_stages + b] = 0;
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the weights were written by generic stores, the tensor core reads them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;
    const uint32_t t_lane = tbase + ((uint32_t)(warp * 32) << 16);  // a warp reaches the 32 TMEM lanes of its quarter
    // TMEM columns, two buffers b = 0, 1: A hi at b*32, A lo at b*32 + 16, distances at 64 + b*NPAD
    constexpr uint32_t COL_D = 4 * KM_TC_KDIM;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);  // f32 acc, tf32 x tf32, K-major, N, M = 128
    const uint32_t bop_addr = smem_u32(bop);
    const unsigned tag_mask = (1u << g_km.tag_bits) - 1u, KEEP = ~tag_mask;
    const float tau = g_km.tau_tc;

    const int64_t n4 = n_px & ~(int64_t)3;
    const int64_t n_blocks = (n4 + BPX - 1) / BPX;
    auto issue = [&](int64_t blk, int s) {
        const int64_t p0 = blk * BPX;
        const unsigned bytes = (unsigned)min((int64_t)BPX, n4 - p0) * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[s], bytes * D);
        float* dst = stages + (size_t)s * D * BPX;
#pragma unroll 1
        for (int d = 0; d < D; ++d) bulk_g2s(dst + d * BPX, stack + d * plane_stride + p0, bytes, &full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < n_stages; ++s) {
            const int64_t blk = blockIdx.x + (int64_t)s * gridDim.x;
            if (blk < n_blocks) issue(blk, s);
        }
    const float my_pow2 = SUMS ? g_km.pow2[min(lane, D - 1)] : 0.f;
    double inertia = 0.0;
    unsigned ties = 0, changed = 0;

    // ---- epilogue of one tile whose distances sit in TMEM buffer `buf`: wst = the warp's 32 pixels of feature 0 in the staged
    //      block (feature d at + d * BPX), p = this thread's pixel, old = its previous label
    auto epilogue = [&](const float* __restrict__ wst, int64_t p, unsigned old, bool valid, uint32_t buf) {
        // tagged (best, second) of the N distances: four independent chains (instruction-level parallelism: a single chain is a
        // dependent FMNMX sequence as long as N), merged at the end
        float cb[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, cs[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
#pragma unroll
        for (int c0 = 0; c0 < NPAD; c0 += 32) {
            if constexpr (NPAD >= 32) {
                uint32_t dv[32];
                tc_ld32(t_lane + COL_D + buf * NPAD + c0, dv);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) KM_ARGMIN_TAGGED(__uint_as_float(dv[j]), cb[j & 3], cs[j & 3], c0 + j)
            } else {
                uint32_t dv[16];
                tc_ld16(t_lane + COL_D + buf * NPAD + c0, dv);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) KM_ARGMIN_TAGGED(__uint_as_float(dv[j]), cb[j & 3], cs[j & 3], c0 + j)
            }
        }
        // union of two (best, second) pairs: best = min of the bests, second = min(seconds, the larger best)
        const float b01 = fminf(cb[0], cb[1]), s01 = fminf(fminf(cs[0], cs[1]), fmaxf(cb[0], cb[1]));
        const float b23 = fminf(cb[2], cb[3]), s23 = fminf(fminf(cs[2], cs[3]), fmaxf(cb[2], cb[3]));
        const float b = fminf(b01, b23), sc = fminf(fminf(s01, s23), fmaxf(b01, b23));
        int l = (int)(__float_as_uint(b) & tag_mask);
        double dex = -1.0;
        unsigned any = __ballot_sync(0xffffffffu, valid && !(sc - b > tau));  // near ties (also NaN)
        while (any) {  // the whole warp on one pixel: plain fp32 first, float64 if that is still too close
            const int src = __ffs(any) - 1;
            any &= any - 1;
            int bi;
            double de = -1.0;
            if (!km_recheck_fp32_warp<D>(wst + src, BPX, K, w32s, &bi)) {
                bi = km_exact_argmin_warp_staged<D>(wst + src, BPX, K, c64s, &de);
                if (lane == src) ++ties;
            }
            if (lane == src) l = bi, dex = de;
        }
        if (valid) {
            if (INERTIA) {
                if (dex < 0.0) {  // sum of squares of (x' - c) in fp32: all terms positive, relative error ~D * 2^-24
                    const float* cent = cent_s + l * D;
                    float dd = 0.f;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float df = fmaf(wst[d * BPX + lane], g_km.scale32[d], g_km.off32[d]) - cent[d];
                        dd = fmaf(df, df, dd);
                    }
                    dex = (double)dd;
                }
                inertia += dex;
            }
            if (lab8) lab8[p] = (uint8_t)l;
            if (lab32) lab32[p] = l;
        }
        const bool moved = valid && prev8 && (unsigned)l != old;
        changed += moved ? 1u : 0u;  // sklearn's strict-convergence test (_kmeans.py:723)
        if (MODE == KM_DELTA) {
            // the warp moves its relabelled pixels together: lane d takes feature d (lane D the count) from the staged block
            unsigned mb = __ballot_sync(0xffffffffu, moved);
            while (mb) {
                const int src = __ffs(mb) - 1;
                mb &= mb - 1;
                const int to = __shfl_sync(0xffffffffu, l, src), from = (int)__shfl_sync(0xffffffffu, old, src);
                if (lane <= D) {
                    long long q = 1;
                    if (lane < D) q = __float2ll_rn(wst[lane * BPX + src] * my_pow2);
                    km_smem_add64(&wacc[to * (D + 1) + lane], q);
                    if (from < KM_MAXK) km_smem_add64(&wacc[from * (D + 1) + lane], -q);
                }
            }
        } else if (MODE == KM_FULL) {
            // every pixel moves in: lane d walks the warp's 32 pixels and adds runs of equal labels together
            const unsigned vb = __ballot_sync(0xffffffffu, valid);
            const float* mine = wst + min(lane, D - 1) * BPX;
            int cur = -1;
            long long run = 0;
#pragma unroll 4
            for (int src = 0; src < 32; ++src) {
                if (!((vb >> src) & 1u)) break;  // valid pixels are a prefix of the warp
                const int tl = __shfl_sync(0xffffffffu, l, src);
                if (lane <= D) {
                    const long long q = lane < D ? __float2ll_rn(mine[src] * my_pow2) : 1ll;
                    if (tl == cur) {
                        run += q;
                    } else {
                        if (cur >= 0) km_smem_add64(&wacc[cur * (D + 1) + lane], run);
                        cur = tl, run = q;
                    }
                }
            }
            if (lane <= D && cur >= 0) km_smem_add64(&wacc[cur * (D + 1) + lane], run);
        }
    };
    // this warp has left a block's stage; the warp whose arrival completes the phase (the last one out) refills it
    auto release = [&](int stage, unsigned parity, int use, int64_t blk) {
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty[stage]);
            if (mbar_test(&empty[stage], parity) && atomicCAS(&ticket[stage], use, use + 1) == use) {
                const int64_t nb = blk + (int64_t)n_stages * gridDim.x;
                if (nb < n_blocks) issue(nb, stage);
            }
        }
    };

    // ---- software pipeline over this CTA's tiles (two TMEM buffers): a warp writes its lanes of A for tile i+1, announces them,
    //      and goes on with the epilogue of tile i.  The warp whose announcement completes the four (a ticket makes it unique)
    //      issues the six MMAs of tile i+1, which then run under everybody's epilogue of tile i.  No CTA-wide barrier: a warp
    //      only ever waits for the MMAs of the tile it is about to read.
    const float* wst_prev = nullptr;
    int64_t p_prev = 0, blk_prev = 0;
    unsigned old_prev = 255u, par_prev = 0;
    int rel_prev = -1, use_prev = 0;  // stage to release after the previous tile's epilogue (-1: not the block's last tile)
    bool valid_prev = false, have_prev = false;
    uint32_t tile_idx = 0;
    int s = 0, use = 0;
    unsigned parity = 0;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int64_t p_blk = blk * BPX;
        const int rem = (int)min((int64_t)BPX, n4 - p_blk);  // pixels of this block
        const int n_sub = (rem + 127) >> 7;
        const float* st = stages + (size_t)s * D * BPX;
        mbar_wait(&full[s], parity);
#pragma unroll 1
        for (int sub = 0; sub < n_sub; ++sub) {
            const uint32_t buf = tile_idx & 1u;
            const int off = sub * 128 + tid;
            const bool valid = off < rem;
            const unsigned old = (prev8 && valid) ? (unsigned)__ldg(prev8 + p_blk + off) : 255u;  // in flight under the MMAs
            {
                // A operand: this thread's pixel in two pieces.  kind::tf32 reads the 11 leading bits of an fp32 operand and drops
                // the rest (tools/tc_probe.cu: 128/128 truncation on the B200), so xh is x itself and xl = x - trunc11(x).
                uint32_t ah[16], al[16];
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    if (d < D) {
                        const float x = st[d * BPX + off];
                        ah[d] = __float_as_uint(x);
                        al[d] = __float_as_uint(x - tf32_trunc(x));
                    } else {
                        ah[d] = d < D + 3 ? 0x3f800000u : 0u;  // 1.0 against the bias pieces
                        al[d] = 0u;
                    }
                }
                tc_st16(t_lane + buf * 2 * KM_TC_KDIM, ah);
                tc_st16(t_lane + buf * 2 * KM_TC_KDIM + KM_TC_KDIM, al);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                // Announce this warp's lanes.  It has also finished reading the distances this buffer held two tiles ago (program
                // order), so once all four warps are in, the buffer may be overwritten.
                const int n_use = (int)(tile_idx >> 1);
                mbar_arrive(&a_ready[buf]);
                if (mbar_test(&a_ready[buf], n_use & 1) && atomicCAS(&ticket[n_stages + buf], n_use, n_use + 1) == n_use) {
                    tc_fence_after();
                    const uint32_t a_hi = tbase + buf * 2 * KM_TC_KDIM, d_col = tbase + COL_D + buf * NPAD;
#pragma unroll
                    for (int term = 0; term < 3; ++term) {  // xh.wh, xh.wl, xl.wh
                        const uint32_t a_col = a_hi + (term == 2 ? KM_TC_KDIM : 0);
                        const uint32_t b_base = bop_addr + (term == 1 ? NPAD * KM_TC_KDIM * 4 : 0);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            tc_mma_tf32(d_col, a_col + 8 * ks, tc_smem_desc(b_base + ks * 2 * CSTR, CSTR, 128), IDESC, (term | ks) ? 1u : 0u);
                    }
                    tc_commit(&mma_bar[buf]);
                }
            }
            __syncwarp();
            if (have_prev) {
                const uint32_t pt = tile_idx - 1;
                mbar_wait(&mma_bar[pt & 1u], (pt >> 1) & 1u);
                tc_fence_after();
                epilogue(wst_prev, p_prev, old_prev, valid_prev, pt & 1u);
                if (rel_prev >= 0) release(rel_prev, par_prev, use_prev, blk_prev);
            }
            wst_prev = st + sub * 128 + warp * 32, p_prev = p_blk + off, old_prev = old, valid_prev = valid, have_prev = true;
            rel_prev = sub == n_sub - 1 ? s : -1, par_prev = parity, use_prev = use, blk_prev = blk;
            ++tile_idx;
        }
        if (++s == n_stages) s = 0, parity ^= 1u, ++use;
    }
    if (have_prev) {  // drain
        const uint32_t pt = tile_idx - 1;
        mbar_wait(&mma_bar[pt & 1u], (pt >> 1) & 1u);
        tc_fence_after();
        epilogue(wst_prev, p_prev, old_prev, valid_prev, pt & 1u);
        if (rel_prev >= 0) release(rel_prev, par_prev, use_prev, blk_prev);
    }
    // ragged tail (n_px % 4 pixels): one thread, scalar fp32 path
    if (blockIdx.x == 0 && tid == 0) {
        for (int64_t q = n4; q < n_px; ++q) {
            float x[D];
            const int l = km_scalar_pixel<D>(stack, plane_stride, q, K, x, INERTIA, inertia, ties);
            const int old = (MODE == KM_FULL || !prev8) ? 255 : (int)prev8[q];
            if (prev8 && prev8[q] != l) ++changed;
            if (SUMS && (MODE == KM_FULL || old != l)) {
                for (int d = 0; d <= D; ++d) {
                    const long long qv = d < D ? __float2ll_rn(x[d] * g_km.pow2[d]) : 1ll;
                    km_smem_add64(&wacc[l * (D + 1) + d], qv);
                    if (old < KM_MAXK) km_smem_add64(&wacc[old * (D + 1) + d], -qv);
                }
            }
            if (lab8) lab8[q] = (uint8_t)l;
            if (lab32) lab32[q] = l;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(tmem_cols) : "memory");
    if (SUMS) {
        for (int i = tid; i < K * (D + 1); i += KM_THREADS) {
            const long long t = wacc[i];
            if (t) {
                const int j = i / (D + 1), d = i % (D + 1);
                long long* dst = d < D ? &gacc[j * D + d] : &gacc[K * D + j];
                atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)t);
            }
        }
    }
    km_commit_counters(gacc, K, D, ties, changed, inertia, inertia_out, INERTIA);
}

// ----------------------------------------------------------------------------- launchers
// ============================================================================= kernel D: bounded passes, K <= 8
// Hamerly's test in front of the assignment (J. Hamerly, "Making k-means even faster", SDM 2010; sklearn's own algorithm="elkan" is
// the same idea): a pixel keeps, next to its label, slack' = (distance to the second nearest centre - distance to its own centre)
// + drift[label] at the time both were computed.  Centres move; by the triangle inequality the pixel's label cannot change while
// slack' > drift[label] now (KmState::drift64).  Such a pixel is skipped without reading its features: 5 bytes (slack', label) instead
// of 4 D.  The others are gathered from a pixel-interleaved copy of the stack (one to three whole sectors per pixel instead of D
// sectors from the planes), evaluated exactly as in kernel B (fp32 distances, float64 inside the near-tie band) and their bound is
// renewed.  Labels, sums and counters are what a delta pass of kernel B produces - every skipped pixel is one whose label provably
// stays - so the centroids stay bit-identical.
//   Rounding: the fp32 squared distances q_j = |x'|^2 + dist_j are within bound_err of the exact ones (rsx_kmeans.cu); the slack is
//   sqrt_rd(q_second - err) - sqrt_ru(q_best + err), sums and drift copies are rounded outwards; a pixel decided in float64 gets
//   slack' = -inf (looked at again in the next pass).
//   FIRST = true: the first bounded pass reads the planes (every pixel) and writes the interleaved copy and the slacks.
constexpr int KB_THREADS = 256;
constexpr int KB_TILE = KB_THREADS * 4;

template <int D, bool FIRST>
__global__ void __launch_bounds__(KB_THREADS, D > 16 ? 2 : 4) km_bounded_kernel(const float* __restrict__ stack, int64_t plane_stride, int64_t n_px,
                                                                             long long* __restrict__ gacc, uint8_t* __restrict__ lab8,
                                                                             float* __restrict__ aos, float* __restrict__ slack) {
    constexpr int DP = km_aos_stride(D);
    __shared__ long long sacc[KM_SLOTS * (D + 1)];
    __shared__ unsigned list[KB_TILE];
    __shared__ int n_list;
    __shared__ float drift_up[KM_SLOTS], drift_dn[KM_SLOTS];  // lanes index them by label: shared memory broadcasts, the constant bank replays
    const int K = g_km.K;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < KM_SLOTS * (D + 1); i += KB_THREADS) sacc[i] = 0;
    if (tid < KM_SLOTS) drift_up[tid] = g_km.drift_up[tid], drift_dn[tid] = g_km.drift_dn[tid];
    if (tid == 0) n_list = 0;
    __syncthreads();
    const int64_t n_tiles = (n_px + KB_TILE - 1) / KB_TILE;
    const float err = g_km.bound_err, tau = g_km.tau_tight;
    unsigned ties = 0, changed = 0;

    // slack' and labels of this thread's four pixels in the next tile: in flight while the current tile is evaluated
    float4 sp_next = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t lb_next = 0;
    auto prefetch = [&](int64_t tile) {
        const int64_t p = tile * KB_TILE + 4 * tid;
        if (!FIRST && tile < n_tiles && p < n_px) {
            sp_next = __ldcs(reinterpret_cast<const float4*>(slack + p));
            lb_next = __ldcs(reinterpret_cast<const uint32_t*>(lab8 + p));
        }
    };
    prefetch(blockIdx.x);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * KB_TILE;
        int count;
        if (!FIRST) {
            // ---- phase 1: Hamerly's test, the pixels that fail it are listed (local index | label << 16)
            const float4 sp = sp_next;
            const uint32_t lb = lb_next;
            const int64_t p = p0 + 4 * tid;
            unsigned mask = 0;
            if (p < n_px) {
                const float spv[4] = {sp.x, sp.y, sp.z, sp.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const unsigned l = (lb >> (8 * i)) & 0xffu;
                                        if (p + i < n_px && !(spv[i] > drift_up[l & (KM_SLOTS - 1)])) mask |= 1u << i;
                }
            }
            prefetch(tile + gridDim.x);
            const int mine = __popc(mask);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int base = 0;
            if (lane == 31 && incl) base = atomicAdd(&n_list, incl);
            base = __shfl_sync(0xffffffffu, base, 31);
            int at = base + incl - mine;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (mask & (1u << i)) list[at++] = (unsigned)(4 * tid + i) | (((lb >> (8 * i)) & 0xffu) << 16);
            __syncthreads();
            count = n_list;
        } else {
            count = (int)min((int64_t)KB_TILE, n_px - p0);
        }
        // ---- phase 2: the listed pixels
        for (int e = tid; e < count; e += KB_THREADS) {
            int local, old;
            if (FIRST) {
                local = e;
                old = lab8[p0 + e];
            } else {
                const unsigned it = list[e];
                local = (int)(it & 0xffffu), old = (int)(it >> 16);
            }
            const int64_t p = p0 + local;
            float x[DP];
            if (FIRST) {
#pragma unroll
                for (int d = 0; d < D; ++d) x[d] = __ldcs(stack + d * plane_stride + p);
#pragma unroll
                for (int d = D; d < DP; ++d) x[d] = 0.f;
                float4* row = reinterpret_cast<float4*>(aos + p * DP);
#pragma unroll
                for (int d = 0; d < DP; d += 4) row[d / 4] = make_float4(x[d], x[d + 1], x[d + 2], x[d + 3]);
            } else {
                const float4* row = reinterpret_cast<const float4*>(aos + p * DP);
#pragma unroll
                for (int d = 0; d < DP; d += 4) {
                    const float4 v = __ldg(row + d / 4);
                    x[d] = v.x, x[d + 1] = v.y, x[d + 2] = v.z, x[d + 3] = v.w;
                }
            }
            float xs = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float xc = fmaf(x[d], g_km.scale32[d], g_km.off32[d]);
                xs = fmaf(xc, xc, xs);
            }
            float b = INFINITY, sc = INFINITY;
            int bi = 0;
#pragma unroll
            for (int j = 0; j < KM_SLOTS; ++j) {
                float a = g_km.bias32[j];
#pragma unroll
                for (int d = 0; d < D; ++d) a = fmaf(x[d], g_km.w32[j * KM_MAXD + d], a);
                KM_ARGMIN_STEP(a, b, sc, bi, j)
            }
            float sp_new;
            if (!(sc - b > tau)) {  // near tie (or NaN): float64 decides, no bound
                double de;
                bi = km_exact_argmin<D>(stack, plane_stride, p, &de);
                ++ties;
                sp_new = -INFINITY;
            } else {
                const float lo2 = __fsqrt_rd(fmaxf(__fsub_rd(__fadd_rd(xs, sc), err), 0.f));
                const float hi1 = __fsqrt_ru(fmaxf(__fadd_ru(__fadd_ru(xs, b), err), 0.f));
                sp_new = __fadd_rd(__fsub_rd(lo2, hi1), drift_dn[bi]);
            }
            slack[p] = sp_new;
            if (bi != old) {
                ++changed;
                lab8[p] = (uint8_t)bi;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const long long q = __float2ll_rn(x[d] * g_km.pow2[d]);
                    km_smem_add64(&sacc[bi * (D + 1) + d], q);
                    if (old < KM_MAXK) km_smem_add64(&sacc[old * (D + 1) + d], -q);
                }
                km_smem_add64(&sacc[bi * (D + 1) + D], 1);
                if (old < KM_MAXK) km_smem_add64(&sacc[old * (D + 1) + D], -1);
            }
        }
        if (!FIRST) {
            __syncthreads();
            if (tid == 0) n_list = 0;
            // the next phase 1 adds to n_list only after its own loads; the barrier inside it orders the reset before the reads
            __syncthreads();
        }
    }
    __syncthreads();
    for (int i = tid; i < K * (D + 1); i += KB_THREADS) {
        const long long t = sacc[i];
        if (t) {
            const int j = i / (D + 1), d = i % (D + 1);
            long long* dst = d < D ? &gacc[j * D + d] : &gacc[K * D + j];
            atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)t);
        }
    }
    km_commit_counters(gacc, K, D, ties, changed, 0.0, nullptr, false);
}

template <int D>
static int km_launch_bounded(const KmLaunch& a, cudaStream_t s) {
    static int per_sm[2] = {0, 0}, cfg_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cfg_dev != dev) {
        cfg_dev = dev;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[0], km_bounded_kernel<D, false>, KB_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], km_bounded_kernel<D, true>, KB_THREADS, 0);
    }
    const bool first = a.bounded == 1;
    const int64_t n_tiles = ceil_div(a.n_px, (int64_t)KB_TILE);
    const int grid = (int)max((int64_t)1, min(n_tiles, (int64_t)rsx_num_sms() * max(1, per_sm[first ? 1 : 0])));
    if (first)
        km_bounded_kernel<D, true><<<grid, KB_THREADS, 0, s>>>(a.stack, a.plane_stride, a.n_px, a.acc, a.lab8, a.aos, a.slack);
    else
        km_bounded_kernel<D, false><<<grid, KB_THREADS, 0, s>>>(a.stack, a.plane_stride, a.n_px, a.acc, a.lab8, a.aos, a.slack);
    return rsx_check_launch("km_bounded");
}

// the 16-bit copy of the stack for the screening passes: u = rint((x - fmin) * 65535 / range), planes q_stride apart
static __global__ void __launch_bounds__(256) km_quantize_kernel(const float* __restrict__ stack, int64_t plane_stride, int64_t n_px, uint16_t* __restrict__ q16,
                                                          int64_t q_stride) {
    const int d = blockIdx.y;
    const float fmin = g_km.qmin32[d], inv = g_km.qinv32[d];
    const float* src = stack + d * plane_stride;
    uint16_t* dst = q16 + d * q_stride;
    const int64_t n4 = (n_px + 3) >> 2;  // the planes are padded to a multiple of 4 samples
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        const float4 x = __ldcs(reinterpret_cast<const float4*>(src) + i);
        auto qz = [&](float v) { return (uint32_t)__float2uint_rn(fminf(fmaxf(__fmul_rn(__fsub_rn(v, fmin), inv), 0.f), 65535.f)); };
        const uint2 o = make_uint2(qz(x.x) | (qz(x.y) << 16), qz(x.z) | (qz(x.w) << 16));
        *reinterpret_cast<uint2*>(dst + 4 * i) = o;
    }
}
static int km_launch_quantize(const KmLaunch& a, uint16_t* q16, cudaStream_t s) {
    const int64_t n4 = (a.n_px + 3) >> 2;
    dim3 grid((unsigned)max((int64_t)1, min(ceil_div(n4, (int64_t)256), (int64_t)rsx_num_sms() * 8)), (unsigned)a.D);
    km_quantize_kernel<<<grid, 256, 0, s>>>(a.stack, a.plane_stride, a.n_px, q16, a.q_stride);
    return rsx_check_launch("km_quantize");
}

template <int D>
static int km_launch_full(const KmLaunch& a, cudaStream_t s) {
    const int smem = KmSmem<D>::CACHE_BYTES;
    auto kern = km_full_kernel<D>;
    static int per_sm = 0, cfg_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (per_sm <= 0 || cfg_dev != dev) {
        cfg_dev = dev;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max(smem, 48 * 1024));
        if (e != cudaSuccess) {
            rsx_set_error("km_full: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
            return RSX_ERR_CUDA;
        }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, KM_THREADS, smem);
        per_sm = max(per_sm, 1);
    }
    const int64_t n4 = a.n_px & ~(int64_t)3;
    const int64_t v_rows = (n4 + a.row_len - 1) / a.row_len;
    const int64_t n_tiles = ceil_div(v_rows, (int64_t)KM_TILE_R) * ceil_div(a.row_len, KM_TILE_W);
    const int grid = (int)max((int64_t)1, min(n_tiles, (int64_t)rsx_num_sms() * per_sm));
    kern<<<grid, KM_THREADS, smem, s>>>(a.stack, a.plane_stride, a.n_px, a.row_len, a.acc, a.lab8, a.prev8, a.lab32, a.pf_rows);
    return rsx_check_launch("km_full");
}

template <int D, int MODE, bool INERTIA, int KU, bool WARPX = false, bool QIN = false>
static int km_launch_stream(const KmLaunch& a, cudaStream_t s) {
    auto kern = km_stream_kernel<D, MODE, INERTIA, KU, WARPX, QIN>;
    constexpr int STAGE_BYTES = QIN ? D * KM_QSTRIDE * 2 : D * KM_STAGE_STRIDE * 4;
    const int acc_bytes = MODE == KM_ASSIGN ? 0 : KM_WARPS * a.K * (D + 1) * 8;
    const int w_bytes = KU == 0 ? (((D + 1) * ((a.K + 7) & ~7) + (INERTIA ? a.K * D : 0)) * 4 + 15) / 16 * 16 : (INERTIA ? (a.K * D * 4 + 15) / 16 * 16 : 0);
    const int c64_bytes = WARPX ? (D + 1) * ((a.K + 31) & ~31) * 8 : 0;
    auto c64_off = [&](int stages) { return stages * STAGE_BYTES + acc_bytes + 2 * stages * 8 + 16 + w_bytes; };
    auto smem_for = [&](int stages) { return c64_off(stages) + c64_bytes; };
    // stages: enough blocks in flight per SM to cover HBM latency at full bandwidth (~64 KB/SM), within shared memory
    static int cfg_K = -1, cfg_stages = 0, cfg_per_sm = 0, cfg_dev = -1, cfg_req = -1;  // per kernel instantiation
    int dev = 0;
    cudaGetDevice(&dev);
    if (cfg_K != a.K || cfg_dev != dev || cfg_req != a.n_stages) {
        int best_stages = 2, best_per_sm = 0;
        for (int stages = a.n_stages > 0 ? a.n_stages : 2; stages <= (a.n_stages > 0 ? a.n_stages : (QIN ? 4 : 3)); ++stages) {
            const int smem = smem_for(stages);
            if (smem > 227 * 1024) break;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max(smem, 48 * 1024)) != cudaSuccess) break;
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, KM_THREADS, smem);
            // prefer more resident CTAs (warps hide the compute latency), then more stages
            if (per_sm > best_per_sm || (per_sm == best_per_sm && per_sm > 0)) best_per_sm = per_sm, best_stages = stages;
        }
        if (best_per_sm <= 0) {
            cudaGetLastError();
            rsx_set_error("km_stream: no launch configuration fits (D=%d, K=%d)", D, a.K);
            return RSX_ERR_CUDA;
        }
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max(smem_for(best_stages), 48 * 1024));
        cfg_K = a.K, cfg_stages = best_stages, cfg_per_sm = best_per_sm, cfg_dev = dev, cfg_req = a.n_stages;
        if (getenv("RSX_DEBUG")) fprintf(stderr, "[rsx] km_stream D=%d K=%d mode=%d: %d stages, %d CTAs/SM, %d B smem\n", D, a.K, MODE, best_stages, best_per_sm, smem_for(best_stages));
    }
    const int64_t n4 = a.n_px & ~(int64_t)3;
    const int64_t n_blocks = ceil_div(n4, (int64_t)KM_BLOCK_PX);
    const int grid = (int)max((int64_t)1, min(n_blocks, (int64_t)rsx_num_sms() * cfg_per_sm));
    kern<<<grid, KM_THREADS, smem_for(cfg_stages), s>>>(a.stack, a.plane_stride, a.n_px, a.acc, a.lab8, a.prev8, a.lab32, a.inertia, cfg_stages,
                                                        c64_off(cfg_stages), a.q16, a.q_stride);
    return rsx_check_launch("km_stream");
}

template <int D, int MODE, bool INERTIA, int NPAD>
static int km_launch_tc_n(const KmLaunch& a, cudaStream_t s) {
    auto kern = km_tc_kernel<D, MODE, INERTIA, NPAD>;
    constexpr int need_cols = 4 * KM_TC_KDIM + 2 * NPAD;  // two A buffers (hi, lo) + two distance buffers
    constexpr int tmem_cols = need_cols <= 128 ? 128 : 256;
    static int cfg_K = -1, cfg_stages = 0, cfg_per_sm = 0, cfg_dev = -1, cfg_req = -1;  // per kernel instantiation
    int dev = 0;
    cudaGetDevice(&dev);
    auto smem_for = [&](int stages) { return KmTcSmem(D, a.K, stages, NPAD, MODE != KM_ASSIGN, INERTIA).total; };
    const int forced = rsx_option("km_tc_ctas", 0);
    if (cfg_K != a.K || cfg_dev != dev || cfg_req != a.n_stages * 16 + forced) {
        int best_stages = 2, best_per_sm = 0;
        for (int stages = a.n_stages > 0 ? a.n_stages : 2; stages <= (a.n_stages > 0 ? a.n_stages : 3); ++stages) {
            const int smem = smem_for(stages);
            if (smem > 227 * 1024) break;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max(smem, 48 * 1024)) != cudaSuccess) break;
            // The occupancy calculator answers 1 CTA/SM for a kernel that allocates tensor memory (it cannot know how many columns
            // the kernel will ask for); the real limits are the 128-register launch bound (4 CTAs), shared memory and the TMEM
            // columns every resident CTA holds for its whole life.
            int per_sm = min(min(4, (227 * 1024) / (smem + 1024)), 512 / tmem_cols);
            if (forced > 0) per_sm = min(per_sm, forced);
            if (per_sm > best_per_sm || (per_sm == best_per_sm && per_sm > 0)) best_per_sm = per_sm, best_stages = stages;
        }
        if (best_per_sm <= 0) {
            cudaGetLastError();
            rsx_set_error("km_tc: no launch configuration fits (D=%d, K=%d)", D, a.K);
            return RSX_ERR_CUDA;
        }
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max(smem_for(best_stages), 48 * 1024));
        cfg_K = a.K, cfg_stages = best_stages, cfg_per_sm = best_per_sm, cfg_dev = dev, cfg_req = a.n_stages * 16 + forced;
        if (getenv("RSX_DEBUG"))
            fprintf(stderr, "[rsx] km_tc D=%d K=%d mode=%d: N=%d, %d TMEM columns, %d stages, %d CTAs/SM, %d B smem\n", D, a.K, MODE, NPAD, tmem_cols, best_stages,
                    best_per_sm, smem_for(best_stages));
    }
    const int64_t n4 = a.n_px & ~(int64_t)3;
    const int64_t n_blocks = ceil_div(n4, (int64_t)KM_TC_BPX);
    const int grid = (int)max((int64_t)1, min(n_blocks, (int64_t)rsx_num_sms() * cfg_per_sm));
    kern<<<grid, KM_THREADS, smem_for(cfg_stages), s>>>(a.stack, a.plane_stride, a.n_px, a.acc, a.lab8, a.prev8, a.lab32, a.inertia, cfg_stages, tmem_cols);
    return rsx_check_launch("km_tc");
}

template <int D, int MODE, bool INERTIA>
static int km_launch_tc(const KmLaunch& a, cudaStream_t s) {
    if (a.K <= 16) return km_launch_tc_n<D, MODE, INERTIA, 16>(a, s);
    if (a.K <= 32) return km_launch_tc_n<D, MODE, INERTIA, 32>(a, s);
    return km_launch_tc_n<D, MODE, INERTIA, 64>(a, s);
}

template <int D, int KU, bool WARPX>
static int km_launch2(const KmLaunch& a, cudaStream_t s) {
    if constexpr (KU == 0 && D + 3 <= KM_TC_KDIM && D >= 6) {  // shallower stacks: too few FMAs to matter
        if (a.use_tc) {  // K > 8: distances on the tensor cores
            if (a.mode == KM_FULL) return km_launch_tc<D, KM_FULL, false>(a, s);
            if (a.mode == KM_DELTA) return km_launch_tc<D, KM_DELTA, false>(a, s);
            if (a.inertia) return km_launch_tc<D, KM_ASSIGN, true>(a, s);
            return km_launch_tc<D, KM_ASSIGN, false>(a, s);
        }
    }
    if (a.mode == KM_FULL) {
        if constexpr (KU == 8 && D <= 20) {
            if (!a.full_stream) return km_launch_full<D>(a, s);
        }
        return km_launch_stream<D, KM_FULL, false, KU, WARPX>(a, s);
    }
    if constexpr (KU == 8) {
        if (a.mode == KM_DELTA && a.q16) return km_launch_stream<D, KM_DELTA, false, KU, false, true>(a, s);
    }
    if (a.mode == KM_DELTA) return km_launch_stream<D, KM_DELTA, false, KU, WARPX>(a, s);
    if (a.inertia) return km_launch_stream<D, KM_ASSIGN, true, KU, WARPX>(a, s);
    return km_launch_stream<D, KM_ASSIGN, false, KU, WARPX>(a, s);
}

// K <= 8: unrolled constant-operand body.  8 < K <= 16: chunked body, near ties re-checked by the lane that owns the pixel.
// K > 16: chunked body, near ties evaluated in float64 by the whole warp (1.55 -> 1.43 ms per pass at K = 32; at K = 16 half the
// lanes would idle: +6 %).
template <int D>
static int km_launch(const KmLaunch& a, cudaStream_t s) {
    if (a.bounded) {
        if (a.K > KM_SLOTS) {
            rsx_set_error("rsx_kmeans_assign_bounded: K=%d (bounded passes are compiled for K <= %d)", a.K, KM_SLOTS);
            return RSX_ERR_UNSUPPORTED;
        }
        return km_launch_bounded<D>(a, s);
    }
    if (a.K <= 8) return km_launch2<D, 8, false>(a, s);
    if (a.K <= 16) return km_launch2<D, 0, false>(a, s);
    return km_launch2<D, 0, true>(a, s);
}

#define KM_CAT2(a, b) a##b
#define KM_CAT(a, b) KM_CAT2(a, b)
#define KM_PART_FN(name) KM_CAT(KM_CAT(rsx_km_part, RSX_KM_PART), name)

template <int D>
static int km_dispatch(const KmLaunch& a, cudaStream_t s) {
    if constexpr (D > km_part_hi(RSX_KM_PART)) {
        rsx_set_error("rsx_kmeans_assign: D=%d is not in this translation unit", a.D);
        return RSX_ERR_UNSUPPORTED;
    } else {
        if (a.D == D) return km_launch<D>(a, s);
        return km_dispatch<D + 1>(a, s);
    }
}

int KM_PART_FN(_assign)(const KmLaunch& a, cudaStream_t s) {
    if (a.mode == KM_QUANTIZE) return km_launch_quantize(a, const_cast<uint16_t*>(a.q16), s);  // reads this unit's constant mirror of the state
    return km_dispatch<km_part_lo(RSX_KM_PART)>(a, s);
}

int KM_PART_FN(_publish)(const void* d_state, cudaStream_t s) {
    cudaError_t e = cudaMemcpyToSymbolAsync(g_km, d_state, sizeof(KmState), 0, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) {
        rsx_set_error("kmeans: publishing state to constant memory failed: %s", cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    return RSX_OK;
}
