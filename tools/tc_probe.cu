// Probe of the tcgen05 building blocks the K > 8 KMeans distance kernel relies on (sm_100a), run on a B200 before the kernel
// is trusted:  D[128 x N] (TMEM, fp32) = A[128 x 16] (TMEM, tf32, written with tcgen05.st: lane = row, column = k)
//                                        x B[N x 16]^T (shared memory, K-major, no swizzle), as two K = 8 steps.
//   1. layout check with small integers (exact in tf32): which (LBO, SBO) reading of the shared-memory descriptor is right;
//   2. operand rounding: are the low 13 mantissa bits of an fp32 operand truncated or rounded by kind::tf32;
//   3. accumulator behaviour: error of a 16-term dot product of 11-bit operands against float64;
//   4. cycles of one stage -> MMA -> commit -> wait -> tcgen05.ld round trip.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;                // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}

template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Dout, int variant,
                                                    long long* cycles, int reps) {
    __shared__ __align__(128) float bsm[N * 16];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int CSTR = (N / 8) * 128;  // bytes between k-chunks (4 tf32 = 16 B) of the B operand
    // element (n, k) at c * CSTR + g * 128 + r * 16 + e * 4 with g = n >> 3, r = n & 7, c = k >> 2, e = k & 3
    for (int i = tid; i < N * 16; i += 128) {
        const int n = i / 16, k = i % 16;
        const int off = (k >> 2) * CSTR + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4;
        bsm[off / 4] = B[n * 16 + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes of bsm -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    const uint32_t t_lane = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t colA = 0, colD = 16;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    unsigned phase = 0;
    long long t0 = clock64();
    float dacc[N];
    for (int rep = 0; rep < reps; ++rep) {
        uint32_t a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = __float_as_uint(A[tid * 16 + k]);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(t_lane + colA),
            "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]),
            "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t lbo = variant == 0 ? CSTR : 128, sbo = variant == 0 ? 128 : CSTR;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const uint64_t bdesc = make_b_desc(smem_u32(bsm) + s * 2 * CSTR, lbo, sbo);
                const uint32_t acc = s;  // first step overwrites
                asm volatile(
                    "{\n"
                    ".reg .pred p;\n"
                    "setp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
                    "}\n" ::"r"(tbase + colD),
                    "r"(tbase + colA + 8 * s), "l"(bdesc), "r"(idesc), "r"(acc)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        // wait for the MMAs
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "W_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
            "@p bra D_%=;\n"
            "bra W_%=;\n"
            "D_%=:\n"
            "}\n" ::"r"(smem_u32(&mbar)),
            "r"(phase)
            : "memory");
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t d[N];
        if constexpr (N == 32) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]),
                  "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]),
                  "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]),
                  "=r"(d[31])
                : "r"(t_lane + colD)
                : "memory");
        } else {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]),
                  "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                : "r"(t_lane + colD)
                : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < N; ++j) dacc[j] = __uint_as_float(d[j]);
    }
    long long t1 = clock64();
    if (tid == 0 && cycles) *cycles = (t1 - t0) / reps;
#pragma unroll
    for (int j = 0; j < N; ++j) Dout[tid * N + j] = dacc[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64));
}

static float trunc13(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}
static float round13(float x) {  // round to nearest even at 10 mantissa bits
    uint32_t u;
    memcpy(&u, &x, 4);
    const uint32_t lsb = (u >> 13) & 1u;
    u += 0xfffu + lsb;
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

template <int N>
static void run(int variant, const std::vector<float>& A, const std::vector<float>& B, std::vector<float>& D, long long* cyc, int reps) {
    float *dA, *dB, *dD;
    long long* dC;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 128 * N * 4));
    probe_kernel<N><<<1, 128>>>(dA, dB, dD, variant, dC, reps);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    D.resize(128 * N);
    CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
    if (cyc) CK(cudaMemcpy(cyc, dC, 8, cudaMemcpyDeviceToHost));
    cudaFree(dA), cudaFree(dB), cudaFree(dD), cudaFree(dC);
}

template <int N>
static void all_tests() {
    srand(1234 + N);
    std::vector<float> A(128 * 16), B(N * 16), D;
    // 1. layout: small integers
    for (auto& v : A) v = (float)(rand() % 15 - 7);
    for (auto& v : B) v = (float)(rand() % 15 - 7);
    for (int variant = 0; variant < 1; ++variant) {  // variant 1 (LBO/SBO swapped) reads outside the operand: illegal address on the B200
        long long cyc = 0;
        run<N>(variant, A, B, D, &cyc, 1);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double r = 0;
                for (int k = 0; k < 16; ++k) r += (double)A[m * 16 + k] * B[n * 16 + k];
                if (D[m * N + n] != (float)r) ++bad;
            }
        printf("N=%d layout variant %d (%s): %d of %d entries wrong; D[0][0..3] = %g %g %g %g\n", N, variant,
               variant == 0 ? "LBO = k-chunk stride, SBO = 8-row group stride" : "swapped", bad, 128 * N, D[0], D[1], D[2], D[3]);
    }
    // 2. operand rounding: one non-zero product per entry, a has low bits set
    std::vector<float> A2(128 * 16, 0.f), B2(N * 16, 0.f);
    for (int m = 0; m < 128; ++m) A2[m * 16 + (m % 16)] = 1.0f + (float)(rand() % 8191 + 1) / 8388608.0f * 1023.0f + (float)(rand() % 1024) / 1024.0f;
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < 16; ++k) B2[n * 16 + k] = 1.0f;
    for (int variant = 0; variant < 1; ++variant) {  // variant 1 (LBO/SBO swapped) reads outside the operand: illegal address on the B200
        run<N>(variant, A2, B2, D, nullptr, 1);
        int t = 0, r = 0, x = 0;
        for (int m = 0; m < 128; ++m) {
            const float a = A2[m * 16 + (m % 16)];
            if (D[m * N] == trunc13(a)) ++t;
            if (D[m * N] == round13(a)) ++r;
            if (D[m * N] == a) ++x;
        }
        printf("N=%d variant %d operand A rounding: %d/128 match truncation, %d/128 match round-to-nearest, %d/128 keep all 24 bits\n", N, variant, t, r, x);
    }
    // 3. accumulation error with 11-bit operands (exact products), 16 terms of mixed sign and magnitude
    for (auto& v : A) v = trunc13(((float)rand() / RAND_MAX - 0.5f) * 4.f);
    for (auto& v : B) v = trunc13(((float)rand() / RAND_MAX - 0.5f) * 4.f);
    for (int variant = 0; variant < 1; ++variant) {  // variant 1 (LBO/SBO swapped) reads outside the operand: illegal address on the B200
        long long cyc = 0;
        run<N>(variant, A, B, D, &cyc, 200);
        double worst = 0, mag = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double r = 0, s = 0;
                for (int k = 0; k < 16; ++k) r += (double)A[m * 16 + k] * B[n * 16 + k], s += fabs((double)A[m * 16 + k] * B[n * 16 + k]);
                worst = fmax(worst, fabs(D[m * N + n] - r) / s);
                mag = fmax(mag, s);
            }
        printf("N=%d variant %d accumulation: worst |D - exact| / sum|terms| = %.3g (2^%.1f); round trip %lld cycles per stage->mma->ld\n", N, variant,
               worst, log2(worst > 0 ? worst : 1e-30), cyc);
    }
}

int main() {
    all_tests<32>();
    all_tests<16>();
    return 0;
}
