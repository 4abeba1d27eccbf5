// KMeans device state + constants shared by the host-side API (rsx_kmeans.cu) and the assign-kernel translation units
// (rsx_kmeans_part.cu, compiled once per range of D).
#pragma once
#include "rsx_common.cuh"

#define KM_MAXD RSX_MAX_FEATURES
#define KM_MAXK RSX_MAX_CLUSTERS

struct __align__(16) KmState {
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
    float tau;        // near-tie band of the tagged fp32 distances
    float tau_tight;  // rounding-only bound (untagged fp32 distances)
    double shift_sq;
    int n_empty;
    int n_updates;
    int tag_bits;  // width of the index tag in the fp32 distances: 3 (K <= 8), 4, 5, 6
    float tau_tc;  // near-tie band of the tensor-core (3 x TF32) distances, tag included
    // Bounded passes (km_bounded_kernel): Hamerly's test.  A pixel whose stored slack (distance to the second nearest minus distance
    // to its own centre, a rigorous lower bound) exceeds what its centre and the fastest other centre have moved since cannot
    // change its label and is skipped unread.  drift64[j] = sum over the updates of |dc_j| + max_{i != j} |dc_i|.
    double drift64[KM_MAXK];
    float drift_up[KM_MAXK], drift_dn[KM_MAXK];  // drift64 rounded outwards (one float ulp beyond)
    double fmin64[KM_MAXD], fmax64[KM_MAXD];     // the raw per-feature range the scaling was derived from (rsx_kmeans_read_scaling)
    // 16-bit screening passes (km_stream_kernel<..., QIN>): features quantised to u = rint((x - fmin) * qinv32), read back as
    // x~ = fma(2^23 + u, qstep32, qoff32) with qoff32 = fl32(fmin - 2^23 qstep32); |x~ - x| <= qerr (derived in km_derive), hence
    // distances within sum_d |w_jd| qerr_d of the fp32 ones: tau_q = tau + twice the largest such sum
    float qstep32[KM_MAXD], qoff32[KM_MAXD], qinv32[KM_MAXD], qmin32[KM_MAXD], qerr32[KM_MAXD];
    float tau_q;
    float bound_err;                             // error bound of the fp32 squared distances |x'|^2 + dist_j (tau_tight + that of |x'|^2)
    float pad_[2];
};


constexpr int KM_ASSIGN = 0, KM_FULL = 1, KM_DELTA = 2;  // the `update` argument of rsx_kmeans_assign
constexpr int KM_QUANTIZE = 3;                           // KmLaunch::mode only: write the 16-bit copy of the stack (no pass)

// one translation unit per range of D (compile time); each owns a __constant__ mirror of the state
#define KM_NUM_PARTS 6
#define KM_MAX_COMPILED_D 24  // the reference's own call site stacks 3 + 19 = 22 planes (scripts/3_classification.py:381-391)
__host__ __device__ constexpr int km_part_of(int D) { return D <= 8 ? 0 : D <= 12 ? 1 : D <= 15 ? 2 : D <= 18 ? 3 : D <= 20 ? 4 : 5; }
__host__ __device__ constexpr int km_part_lo(int part) { return part == 0 ? 1 : part == 1 ? 9 : part == 2 ? 13 : part == 3 ? 16 : part == 4 ? 19 : 21; }
__host__ __device__ constexpr int km_part_hi(int part) { return part == 0 ? 8 : part == 1 ? 12 : part == 2 ? 15 : part == 3 ? 18 : part == 4 ? 20 : KM_MAX_COMPILED_D; }
// D > 20: a thread's four pixels x D features no longer fit 128 registers - two CTAs per SM instead of four, and the first
// (full) pass of K <= 8 goes through the streaming kernel (per-thread accumulators would not fit in shared memory)
__host__ __device__ constexpr int km_ctas_per_sm(int D) { return D > 20 ? 2 : 4; }

struct KmLaunch {
    const float* stack;
    int64_t plane_stride, n_px;
    int row_len;
    long long* acc;
    uint8_t* lab8;
    const uint8_t* prev8;
    int32_t* lab32;
    double* inertia;
    int mode, D, K;
    int pf_rows;   // full-pass kernel: L2 prefetch distance in rows (0 = off)
    int n_stages;  // streaming kernel: blocks in flight per CTA (0 = choose)
    int use_tc;    // K > 8, D <= 13: distances on the tensor cores (tcgen05, 3 x TF32 split); 0 = the fp32 FFMA2 path
    // bounded passes (K <= 8): 0 = off; 1 = first one (reads the planes, writes the pixel-interleaved copy + slacks); 2 = later ones
    int bounded;
    const uint16_t* q16;  // [D][q_stride] 16-bit copy of the stack (screening passes), or nullptr
    int64_t q_stride;
    int full_stream;  // K <= 8 full pass through the streaming kernel (every pixel moves in) instead of the per-thread accumulators
    float* aos;    // [n_px][km_aos_stride(D)] pixel-interleaved copy of the stack (scattered reads cost 1-3 sectors instead of D)
    float* slack;  // [n_px rounded up to 4] slack + drift of the label at the time it was computed
};
__host__ __device__ constexpr int km_aos_stride(int D) { return (D + 7) & ~7; }  // whole 32-byte sectors per pixel
typedef int (*km_assign_fn)(const KmLaunch&, cudaStream_t);
typedef int (*km_publish_fn)(const void* d_state, cudaStream_t);
#define KM_DECLARE_PART(N)                                      \
    int rsx_km_part##N##_assign(const KmLaunch&, cudaStream_t); \
    int rsx_km_part##N##_publish(const void* d_state, cudaStream_t);
KM_DECLARE_PART(0) KM_DECLARE_PART(1) KM_DECLARE_PART(2) KM_DECLARE_PART(3) KM_DECLARE_PART(4) KM_DECLARE_PART(5)
