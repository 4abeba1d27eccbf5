// Micro-benchmark: what read bandwidth does the KMeans access pattern reach with trivial compute?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/stream_bench.cu -o tools/bin/stream_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ float4 ldg4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// planar D streams, tile walk as in km_assign (tile = 512 px x R rows of row_len), U rows in flight per thread
template <int D, int U, int MINB>
__global__ void __launch_bounds__(128, MINB) planar_kernel(const float* __restrict__ stack, int64_t stride, int64_t n4, int row_len, int R, int pf, float* out) {
    const int tiles_x = (row_len + 511) / 512;
    const int64_t v_rows = (n4 + row_len - 1) / row_len;
    const int64_t tiles = ((v_rows + R - 1) / R) * tiles_x;
    float acc = 0.f;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const int64_t ty = tile / tiles_x;
        const int col = tx * 512 + threadIdx.x * 4;
        const int64_t r0 = ty * R, r1 = min(v_rows, r0 + R);
        for (int64_t r = r0; r < r1; r += U) {
            float4 v[U][D];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t p = (r + u) * row_len + col;
                const bool ok = col < row_len && r + u < r1 && p < n4;
                if (pf && ok && (threadIdx.x & 7) == 0) {
#pragma unroll
                    for (int d = 0; d < D; ++d) asm volatile("prefetch.global.L2 [%0];" ::"l"(stack + d * stride + p + (int64_t)pf * row_len));
                }
#pragma unroll
                for (int d = 0; d < D; ++d) v[u][d] = ok ? ldg4(stack + d * stride + p) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int d = 0; d < D; ++d) acc += v[u][d].x * v[u][d].y + v[u][d].z * v[u][d].w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// flat: every thread streams float4s of one contiguous array
__global__ void __launch_bounds__(256) flat_kernel(const float4* __restrict__ a, int64_t n, float* out) {
    float acc = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = a[i];
        acc += v.x * v.y + v.z * v.w;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 10) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    f();
    f();
    float best = 1e9;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    constexpr int D = 13;
    const int64_t n = 49000000, stride = (n + 31) / 32 * 32;
    float *stack, *out;
    CK(cudaMalloc(&stack, stride * D * 4));
    CK(cudaMalloc(&out, 4));
    CK(cudaMemset(stack, 0, stride * D * 4));
    const double gb = (double)n * D * 4 / 1e9;
    int sms = 148;
    {
        float ms = time_ms([&] { flat_kernel<<<sms * 8, 256>>>((const float4*)stack, stride * D / 4, out); });
        printf("flat read                       : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    }
#define RUN(U, MINB, RL, R, PF)                                                                                                  \
    {                                                                                                                            \
        float ms = time_ms([&] { planar_kernel<D, U, MINB><<<sms * MINB, 128>>>(stack, stride, n & ~3ll, RL, R, PF, out); });     \
        printf("planar U=%d ctas/sm=%d row_len=%5d R=%3d pf=%d : %.3f ms  %.0f GB/s\n", U, MINB, RL, R, PF, ms, gb / ms * 1e3); \
    }
    RUN(1, 4, 7000, 32, 0) RUN(1, 4, 7000, 32, 2) RUN(1, 4, 512, 32, 0) RUN(1, 4, 512, 32, 2) RUN(1, 4, 512, 128, 0)
    RUN(2, 2, 7000, 32, 0) RUN(2, 2, 512, 32, 0) RUN(2, 2, 512, 32, 4)
    RUN(1, 8, 7000, 32, 0) RUN(1, 8, 512, 32, 0) RUN(1, 8, 512, 128, 0) RUN(1, 6, 512, 32, 0)
    CK(cudaDeviceSynchronize());
    return 0;
}
