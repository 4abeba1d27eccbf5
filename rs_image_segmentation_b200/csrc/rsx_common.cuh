// Shared device/host helpers for the rsx kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rsx.h"

#define RSX_SM_COUNT_FALLBACK 148

// ----------------------------------------------------------------------------- errors
void rsx_set_error(const char* fmt, ...);
int rsx_check_launch(const char* what);
int rsx_num_sms();
// integer tuning knob: rsx_set_option value, else env RSX_<NAME>, else dflt (rsx_core.cu)
int rsx_option(const char* name, int dflt);
// device -> host of a small block through a kernel store into a page-locked staging buffer; synchronises the stream
int rsx_fetch_small(void* h_dst, const void* d_src, size_t bytes, cudaStream_t s);

#define RSX_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            rsx_set_error(__VA_ARGS__); \
            return RSX_ERR_ARG;         \
        }                               \
    } while (0)

// ----------------------------------------------------------------------------- IEEE fp32 without contraction
// The reference is numpy float32: every operation rounds once.  The *_rn intrinsics are never
// fused into FMAs by nvcc, which is what makes the index maps bit-exact.
__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float f_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float f_clip(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
// np.clip keeps NaN (fminf/fmaxf would drop it): for the planar drop-ins, whose inputs are arbitrary caller arrays
__device__ __forceinline__ float f_clip_nan(float x, float lo, float hi) { return x != x ? x : fminf(fmaxf(x, lo), hi); }

// robust_normalize (modules/features/indices.py:42-46): (clip(x, lo, hi) - lo) / (hi - lo + 1e-10)
// den = fl32(fl32(hi - lo) + 1e-10f) is precomputed on the host in float32.
struct NormParam {
    float lo, hi, den;
};
__device__ __forceinline__ float norm_apply(float x, NormParam p) { return f_div(f_sub(f_clip(x, p.lo, p.hi), p.lo), p.den); }

// ----------------------------------------------------------------------------- ordered float <-> uint for atomic min/max
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// min/max tracker: slot layout in global memory is uint32 [n][2] = {ordered min, ordered max}
__device__ __forceinline__ void warp_minmax_commit(float mn, float mx, unsigned* slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // the tracker only ever moves outwards: a plain (possibly stale) read tells most warps that they have nothing to add, which
        // keeps tens of thousands of same-address atomics of the many-CTA kernels out of the L2 atomic unit
        const unsigned omn = f2ord(mn), omx = f2ord(mx);
        if (omn < *reinterpret_cast<volatile unsigned*>(slot)) atomicMin(slot, omn);
        if (omx > *reinterpret_cast<volatile unsigned*>(slot + 1)) atomicMax(slot + 1, omx);
    }
}

// ----------------------------------------------------------------------------- streaming loads / stores
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ----------------------------------------------------------------------------- mbarrier + 1-D bulk copy (TMA engine, no tensor map)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {  // non-blocking: has the phase with this parity completed?
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"  // suspend-time hint: parked warps do not spin
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// bytes and both addresses must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) { return (a + b - 1) / b; }
