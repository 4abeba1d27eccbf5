// Error handling, launch accounting, min/max trackers and the planar float32 element-wise
// entry points (per-function drop-ins for modules/features/indices.py).
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "rsx_common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void rsx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int rsx_check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        rsx_set_error("%s: %s", what, cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    return RSX_OK;
}

int rsx_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = RSX_SM_COUNT_FALLBACK;
    }
    return n;
}

// ----------------------------------------------------------------------------- host side of the label download
// The reference returns int32 labels (extract.py:577); the device keeps them as uint8 (K <= 64).  Downloading the uint8 plane and
// widening it on the host's cores moves a quarter of the bytes over PCIe (49 instead of 196 MB for a 7000 x 7000 scene), which is
// what bounds the end-to-end rate once several GPUs of one box stream at the same time.
#include <thread>
#include <vector>
extern "C" int rsx_widen_u8_to_i32(const uint8_t* h_src, int32_t* h_dst, int64_t n, int n_threads) {
    RSX_REQUIRE(h_src && h_dst && n >= 0, "rsx_widen_u8_to_i32: bad arguments");
    n_threads = (int)max((int64_t)1, min((int64_t)min(n_threads, 64), n / (1 << 16)));
    auto work = [=](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; ++i) h_dst[i] = (int32_t)h_src[i];  // vectorised by the host compiler
    };
    if (n_threads <= 1) {
        work(0, n);
        return RSX_OK;
    }
    std::vector<std::thread> pool;
    const int64_t chunk = ((n + n_threads - 1) / n_threads + 63) & ~(int64_t)63;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t a = min(n, t * chunk), b = min(n, a + chunk);
        if (b > a) pool.emplace_back(work, a, b);
    }
    for (auto& th : pool) th.join();
    return RSX_OK;
}

// ----------------------------------------------------------------------------- tuning options
// Small registry of integer knobs (kernel variant switches used by the profiling tools and the A/B tests).  A knob set through
// rsx_set_option wins; otherwise the environment variable RSX_<NAME IN UPPER CASE> is consulted on every query; otherwise the
// default of the call site.
struct RsxOption {
    char name[32];
    int value;
};
static RsxOption g_options[32];
static int g_n_options = 0;
static std::mutex g_option_mu;

int rsx_option(const char* name, int dflt) {
    {
        std::lock_guard<std::mutex> lock(g_option_mu);
        for (int i = 0; i < g_n_options; ++i)
            if (!strcmp(g_options[i].name, name)) return g_options[i].value;
    }
    char env[48] = "RSX_";
    size_t k = 4;
    for (const char* c = name; *c && k + 1 < sizeof(env); ++c) env[k++] = (char)((*c >= 'a' && *c <= 'z') ? *c - 32 : *c);
    env[k] = 0;
    const char* e = getenv(env);
    return e ? atoi(e) : dflt;
}

extern "C" int rsx_set_option(const char* name, int value) {
    RSX_REQUIRE(name && strlen(name) < sizeof(g_options[0].name), "rsx_set_option: bad name");
    std::lock_guard<std::mutex> lock(g_option_mu);
    for (int i = 0; i < g_n_options; ++i)
        if (!strcmp(g_options[i].name, name)) {
            g_options[i].value = value;
            return RSX_OK;
        }
    RSX_REQUIRE(g_n_options < (int)(sizeof(g_options) / sizeof(g_options[0])), "rsx_set_option: table full");
    strcpy(g_options[g_n_options].name, name);
    g_options[g_n_options++].value = value;
    return RSX_OK;
}

extern "C" int rsx_get_option(const char* name, int dflt) { return name ? rsx_option(name, dflt) : dflt; }

extern "C" const char* rsx_last_error(void) { return g_err; }
extern "C" int rsx_abi_version(void) { return 1; }
extern "C" int64_t rsx_launch_count(void) { return g_launches.load(); }

// ----------------------------------------------------------------------------- small results to the host without a copy engine
// The device-to-host copy engine serves one transfer at a time; a histogram or min/max read-back queued behind a scene-sized
// label download waits for all of it (3.5 ms at 196 MB).  A kernel that stores straight into page-locked (mapped) host memory
// does not queue there.
__global__ void store_to_host_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t n16, int64_t n) {
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = i0; i < n16; i += stride) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (int64_t i = (n16 << 4) + i0; i < n; i += stride) dst[i] = src[i];
}

extern "C" int rsx_store_to_host(const void* d_src, void* h_mapped_dst, int64_t bytes, rsx_stream_t stream) {
    RSX_REQUIRE(d_src && h_mapped_dst && bytes >= 0, "rsx_store_to_host: bad arguments");
    if (bytes == 0) return RSX_OK;
    const bool vec = (((uintptr_t)d_src | (uintptr_t)h_mapped_dst) & 15) == 0;
    const int64_t n16 = vec ? bytes >> 4 : 0;
    const int64_t work = vec ? n16 + 15 : bytes;
    const int grid = (int)max((int64_t)1, min((work + 255) / 256, (int64_t)rsx_num_sms() * 4));
    store_to_host_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)d_src, (uint8_t*)h_mapped_dst, n16, bytes);
    return rsx_check_launch("store_to_host");
}

int rsx_fetch_small(void* h_dst, const void* d_src, size_t bytes, cudaStream_t s) {
    static std::mutex mu;
    static void* stage = nullptr;
    static size_t cap = 0;
    std::lock_guard<std::mutex> lock(mu);
    if (bytes > cap) {
        if (stage) cudaFreeHost(stage);
        stage = nullptr, cap = 0;
        const size_t want = max(bytes, (size_t)1 << 16);
        cudaError_t e = cudaHostAlloc(&stage, want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            stage = nullptr;
            rsx_set_error("rsx_fetch_small: cudaHostAlloc(%zu): %s", want, cudaGetErrorString(e));
            return RSX_ERR_CUDA;
        }
        cap = want;
    }
    if (int rc = rsx_store_to_host(d_src, stage, (int64_t)bytes, (rsx_stream_t)s)) return rc;
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        rsx_set_error("rsx_fetch_small: %s", cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    memcpy(h_dst, stage, bytes);
    return RSX_OK;
}

// ----------------------------------------------------------------------------- peer memory (one node, one process per GPU)
// A small cudaMalloc'ed block that the other ranks map with CUDA IPC: the KMeans update kernel reduces the ranks' partial
// sums through it over NVLink (rsx_kmeans_update_peers) instead of calling a collective.
extern "C" int rsx_peer_alloc(int64_t bytes, void** d_ptr, uint8_t* h_handle64) {
    RSX_REQUIRE(bytes > 0 && d_ptr && h_handle64, "rsx_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        rsx_set_error("rsx_peer_alloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    memcpy(h_handle64, &h, 64);
    *d_ptr = p;
    return RSX_OK;
}

extern "C" int rsx_peer_open(const uint8_t* h_handle64, void** d_peer) {
    RSX_REQUIRE(h_handle64 && d_peer, "rsx_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(d_peer, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        rsx_set_error("rsx_peer_open: %s", cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    return RSX_OK;
}

extern "C" int rsx_peer_zero(void* d_ptr, int64_t bytes, rsx_stream_t stream) {
    RSX_REQUIRE(d_ptr && bytes >= 0, "rsx_peer_zero: bad arguments");
    cudaError_t e = cudaMemsetAsync(d_ptr, 0, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        rsx_set_error("rsx_peer_zero: %s", cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    return RSX_OK;
}

extern "C" int rsx_peer_close(void* d_peer) {
    if (d_peer && cudaIpcCloseMemHandle(d_peer) != cudaSuccess) {
        cudaGetLastError();
        return RSX_ERR_CUDA;
    }
    return RSX_OK;
}

extern "C" int rsx_peer_free(void* d_ptr) {
    if (d_ptr && cudaFree(d_ptr) != cudaSuccess) {
        cudaGetLastError();
        return RSX_ERR_CUDA;
    }
    return RSX_OK;
}

// ----------------------------------------------------------------------------- min/max trackers
__global__ void minmax_init_kernel(uint32_t* mm, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        mm[2 * i] = 0xffffffffu;  // ordered +inf side
        mm[2 * i + 1] = 0u;
    }
}

extern "C" int rsx_minmax_init(uint32_t* d_minmax, int n, rsx_stream_t stream) {
    RSX_REQUIRE(d_minmax && n > 0, "rsx_minmax_init: bad arguments");
    minmax_init_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(d_minmax, n);
    return rsx_check_launch("minmax_init");
}

static inline unsigned h_f2ord(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float h_ord2f(unsigned u) {
    unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &v, 4);
    return f;
}
extern "C" void rsx_minmax_decode(const uint32_t* h, int n, float* h_min, float* h_max) {
    for (int i = 0; i < n; ++i) {
        h_min[i] = h_ord2f(h[2 * i]);
        h_max[i] = h_ord2f(h[2 * i + 1]);
    }
}
extern "C" void rsx_minmax_encode(const float* h_min, const float* h_max, int n, uint32_t* h) {
    for (int i = 0; i < n; ++i) {
        h[2 * i] = h_f2ord(h_min[i]);
        h[2 * i + 1] = h_f2ord(h_max[i]);
    }
}

// grid.y = plane; NaNs are ignored by fminf/fmaxf, matching the NaN->0 replacement being done first
__global__ void __launch_bounds__(256) minmax_planes_kernel(const float* __restrict__ p, int64_t n, int64_t stride, uint32_t* mm) {
    const float* base = p + (int64_t)blockIdx.y * stride;
    float mn = INFINITY, mx = -INFINITY;
    int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = ldg_stream4(base + 4 * i);
        mn = fminf(fminf(mn, fminf(v.x, v.y)), fminf(v.z, v.w));
        mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        float v = base[(n4 << 2) + threadIdx.x];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    warp_minmax_commit(mn, mx, mm + 2 * blockIdx.y);
}

extern "C" int rsx_minmax_planes_f32(const float* d_planes, int64_t n_px, int64_t plane_stride, int n_planes, uint32_t* d_minmax,
                                     rsx_stream_t stream) {
    RSX_REQUIRE(d_planes && d_minmax && n_px > 0 && n_planes > 0, "rsx_minmax_planes_f32: bad arguments");
    RSX_REQUIRE(((uintptr_t)d_planes & 15) == 0 && (plane_stride & 3) == 0, "rsx_minmax_planes_f32: planes must be 16-byte aligned");
    int gx = (int)min((int64_t)rsx_num_sms() * 4 / n_planes + 1, ceil_div(n_px, (int64_t)1024));
    minmax_planes_kernel<<<dim3(gx, n_planes), 256, 0, (cudaStream_t)stream>>>(d_planes, n_px, plane_stride, d_minmax);
    return rsx_check_launch("minmax_planes");
}

__global__ void __launch_bounds__(256) nan_to_zero_kernel(float* p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = p[i];
        if (v != v) p[i] = 0.f;
    }
}
// NaN -> 0 on one plane of a stack whose min/max tracker was filled by the producer (which skips NaN): when a NaN was replaced,
// 0 joins the tracked range - MinMaxScaler is fitted after the replacement in the reference (extract.py:548-570).
__global__ void __launch_bounds__(256) nan_to_zero_minmax_kernel(float* __restrict__ p, int64_t n, uint32_t* __restrict__ slot) {
    bool any = false;
    const int64_t n4 = n >> 2;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = i0; i < n4; i += stride) {
        float4 v = ldg_stream4(p + 4 * i);
        if (v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w) {
            v.x = v.x != v.x ? 0.f : v.x, v.y = v.y != v.y ? 0.f : v.y, v.z = v.z != v.z ? 0.f : v.z, v.w = v.w != v.w ? 0.f : v.w;
            *reinterpret_cast<float4*>(p + 4 * i) = v;
            any = true;
        }
    }
    for (int64_t i = (n4 << 2) + i0; i < n; i += stride)
        if (p[i] != p[i]) p[i] = 0.f, any = true;
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0 && slot) {
        atomicMin(slot, f2ord(0.f));
        atomicMax(slot + 1, f2ord(0.f));
    }
}
extern "C" int rsx_nan_to_zero_minmax_f32(float* d_plane, int64_t n, uint32_t* d_minmax_slot, rsx_stream_t stream) {
    RSX_REQUIRE(d_plane && n > 0 && ((uintptr_t)d_plane & 15) == 0, "rsx_nan_to_zero_minmax_f32: bad arguments (the plane must be 16-byte aligned)");
    nan_to_zero_minmax_kernel<<<(int)min((int64_t)rsx_num_sms() * 8, ceil_div(n, (int64_t)1024)), 256, 0, (cudaStream_t)stream>>>(d_plane, n, d_minmax_slot);
    return rsx_check_launch("nan_to_zero_minmax");
}

extern "C" int rsx_nan_to_zero_f32(float* d, int64_t n, rsx_stream_t stream) {
    RSX_REQUIRE(d && n > 0, "rsx_nan_to_zero_f32: bad arguments");
    nan_to_zero_kernel<<<(int)min((int64_t)rsx_num_sms() * 8, ceil_div(n, (int64_t)256)), 256, 0, (cudaStream_t)stream>>>(d, n);
    return rsx_check_launch("nan_to_zero");
}

// ----------------------------------------------------------------------------- planar element-wise ops
// One functor per reference function; the driver handles vectorised body + scalar tail.
struct OpNormalize {
    NormParam p;
    __device__ float operator()(float x, float, float, float) const { return f_div(f_sub(f_clip_nan(x, p.lo, p.hi), p.lo), p.den); }
};
// indices.py:62-69 and its three siblings: num = a-b, den = a+b
struct OpRatio {
    __device__ float operator()(float a, float b, float, float) const {
        float den = f_add(a, b);
        float r = den > 0.001f ? f_div(f_sub(a, b), den) : 0.f;
        return f_clip_nan(r, -1.f, 1.f);
    }
};
// indices.py:86-93: den = nir + C1*red - C2*blue + L ; G*(nir-red)/den
struct OpEvi {
    float L, C1, C2, G;
    __device__ float operator()(float nir, float red, float blue, float) const {
        float den = f_add(f_sub(f_add(nir, f_mul(C1, red)), f_mul(C2, blue)), L);
        float r = den > 0.001f ? f_div(f_mul(G, f_sub(nir, red)), den) : 0.f;
        return f_clip_nan(r, -1.f, 1.f);
    }
};
// indices.py:109: (2n+1 - sqrt((2n+1)^2 - 8(n-r)))/2 ; NaN propagates through the clip like np.clip
struct OpMsavi {
    __device__ float operator()(float nir, float red, float, float) const {
        float t = f_add(f_mul(2.f, nir), 1.f);
        float rad = f_sub(f_mul(t, t), f_mul(8.f, f_sub(nir, red)));
        float v = f_div(f_sub(t, f_sqrt(rad)), 2.f);
        return v != v ? v : f_clip(v, -1.f, 1.f);
    }
};
// indices.py:194-201
struct OpBsi {
    __device__ float operator()(float blue, float red, float nir, float swir) const {
        float a = f_add(swir, red), b = f_add(nir, blue);
        float den = f_add(a, b);
        float r = den > 0.001f ? f_div(f_sub(a, b), den) : 0.f;
        return f_clip_nan(r, -1.f, 1.f);
    }
};

template <int NIN, typename Op>
__global__ void __launch_bounds__(256) planar_op_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                                        const float* __restrict__ d, float* __restrict__ out, int64_t n, Op op) {
    int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 va = ldg_stream4(a + 4 * i), vb = va, vc = va, vd = va;
        if (NIN > 1) vb = ldg_stream4(b + 4 * i);
        if (NIN > 2) vc = ldg_stream4(c + 4 * i);
        if (NIN > 3) vd = ldg_stream4(d + 4 * i);
        float4 r;
        r.x = op(va.x, vb.x, vc.x, vd.x);
        r.y = op(va.y, vb.y, vc.y, vd.y);
        r.z = op(va.z, vb.z, vc.z, vd.z);
        r.w = op(va.w, vb.w, vc.w, vd.w);
        stg_stream4(out + 4 * i, r);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        int64_t i = (n4 << 2) + threadIdx.x;
        out[i] = op(a[i], NIN > 1 ? b[i] : 0.f, NIN > 2 ? c[i] : 0.f, NIN > 3 ? d[i] : 0.f);
    }
}

template <int NIN, typename Op>
static int launch_planar(const char* name, const float* a, const float* b, const float* c, const float* d, float* out, int64_t n, Op op,
                         rsx_stream_t stream) {
    RSX_REQUIRE(a && out && n > 0, "%s: bad arguments", name);
    RSX_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)out) & 15) == 0, "%s: buffers must be 16-byte aligned", name);
    int grid = (int)min((int64_t)rsx_num_sms() * 8, ceil_div(n, (int64_t)1024));
    planar_op_kernel<NIN, Op><<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, c, d, out, n, op);
    return rsx_check_launch(name);
}

extern "C" int rsx_normalize_f32(const float* d_in, int64_t n, float lo, float hi, float den, float* d_out, rsx_stream_t stream) {
    OpNormalize op{{lo, hi, den}};
    return launch_planar<1>("rsx_normalize_f32", d_in, nullptr, nullptr, nullptr, d_out, n, op, stream);
}
extern "C" int rsx_index_ratio_f32(const float* d_a, const float* d_b, int64_t n, float* d_out, rsx_stream_t stream) {
    RSX_REQUIRE(d_b, "rsx_index_ratio_f32: bad arguments");
    return launch_planar<2>("rsx_index_ratio_f32", d_a, d_b, nullptr, nullptr, d_out, n, OpRatio{}, stream);
}
extern "C" int rsx_index_evi_f32(const float* d_nir, const float* d_red, const float* d_blue, int64_t n, float L, float C1, float C2, float G,
                                 float* d_out, rsx_stream_t stream) {
    RSX_REQUIRE(d_red && d_blue, "rsx_index_evi_f32: bad arguments");
    return launch_planar<3>("rsx_index_evi_f32", d_nir, d_red, d_blue, nullptr, d_out, n, OpEvi{L, C1, C2, G}, stream);
}
extern "C" int rsx_index_msavi_f32(const float* d_nir, const float* d_red, int64_t n, float* d_out, rsx_stream_t stream) {
    RSX_REQUIRE(d_red, "rsx_index_msavi_f32: bad arguments");
    return launch_planar<2>("rsx_index_msavi_f32", d_nir, d_red, nullptr, nullptr, d_out, n, OpMsavi{}, stream);
}
extern "C" int rsx_index_bsi_f32(const float* d_blue, const float* d_red, const float* d_nir, const float* d_swir, int64_t n, float* d_out,
                                 rsx_stream_t stream) {
    RSX_REQUIRE(d_red && d_nir && d_swir, "rsx_index_bsi_f32: bad arguments");
    return launch_planar<4>("rsx_index_bsi_f32", d_blue, d_red, d_nir, d_swir, d_out, n, OpBsi{}, stream);
}

__global__ void __launch_bounds__(256) quantize_f32_kernel(const float* __restrict__ in, int64_t n, NormParam p, float lm1, uint8_t* __restrict__ q) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        q[i] = (uint8_t)(int)f_mul(norm_apply(in[i], p), lm1);
}
extern "C" int rsx_quantize_f32(const float* d_in, int64_t n, float lo, float hi, float den, int levels, uint8_t* d_q, rsx_stream_t stream) {
    RSX_REQUIRE(d_in && d_q && n > 0 && levels >= 2 && levels <= 256, "rsx_quantize_f32: bad arguments");
    quantize_f32_kernel<<<(int)min((int64_t)rsx_num_sms() * 8, ceil_div(n, (int64_t)256)), 256, 0, (cudaStream_t)stream>>>(
        d_in, n, NormParam{lo, hi, den}, (float)(levels - 1), d_q);
    return rsx_check_launch("rsx_quantize_f32");
}

// ----------------------------------------------------------------------------- host: order statistics from band histograms
// The few scalars (and per-level tables) the kernels need from the K1 histograms: robust_normalize's P2/P98
// (modules/features/indices.py:38-46), the second robust_normalize of the texture band (:265), RobustScaler's median / IQR
// (sklearn/preprocessing/_data.py:1722,1738-1743) and the per-level table of the scaled value.  This is numpy's `_quantile`
// (method="linear") restated operation by operation in the dtypes numpy ends up using for float32 data - see
// rs_image_segmentation_b200/hoststats.py, the Python statement of the same arithmetic that the tests hold it against.
// Host code, O(levels) per band; it sits between two kernels, so it has to take microseconds, not a millisecond of Python.
namespace {
struct Order {
    const int64_t* cum;   // inclusive cumulative counts over the levels
    const float* values;  // value of every level (float32)
    int L;
    int64_t n;
    float at(int64_t k) const {  // k-th smallest sample
        k = k < 0 ? 0 : (k >= n ? n - 1 : k);
        int lo = 0, hi = L;  // first level with cum > k  (np.searchsorted(cum, k, "right"))
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cum[mid] > k) hi = mid; else lo = mid + 1;
        }
        return values[lo < L ? lo : L - 1];
    }
    // np.percentile(float32 band, q) with a Python-number q: everything in float32
    float percentile32(double q) const {
        const float q32 = (float)q / 100.0f;
        const float nm1 = (float)(n - 1);
        const volatile float vi = nm1 * q32;
        float prev = floorf(vi), nxt = prev + 1.0f;
        if (vi >= nm1) prev = nxt = -1.0f;
        if (vi < 0.0f) prev = nxt = 0.0f;
        const int64_t pi = (int64_t)prev, ni = (int64_t)nxt;
        const float a = at(pi >= 0 ? pi : n - 1), b = at(ni >= 0 ? ni : n - 1);
        if (a == b) return a;
        const float t = (float)((double)vi - (double)pi);
        const volatile float diff = b - a;
        volatile float prod = diff * t;
        volatile float out = a + prod;
        if (t >= 0.5f) {
            volatile float omt = 1.0f - t;
            prod = diff * omt;
            out = b - prod;
        }
        return out;
    }
    // np.nanpercentile(float32 column, (..., q, ...)) with float64 q: float64 virtual index, float64 result
    double percentile64(double q) const {
        const double qq = q / 100.0;
        const volatile double vi = (double)(n - 1) * qq;
        double prev = floor(vi), nxt = prev + 1.0;
        if (vi >= (double)(n - 1)) prev = nxt = -1.0;
        if (vi < 0.0) prev = nxt = 0.0;
        const int64_t pi = (int64_t)prev, ni = (int64_t)nxt;
        const float a = at(pi >= 0 ? pi : n - 1), b = at(ni >= 0 ? ni : n - 1);
        if (a == b) return (double)a;
        const double t = vi - (double)pi;
        const volatile float diff = b - a;
        volatile double out = (double)a + (double)diff * t;
        if (t >= 0.5) out = (double)b - (double)diff * (1.0 - t);
        return out;
    }
    float median32() const {
        if (n & 1) return at(n / 2);
        const volatile float s = at(n / 2 - 1) + at(n / 2);
        return s / 2.0f;
    }
};
}  // namespace

extern "C" int rsx_raster_stats(const int64_t* h_hist, int n_bands, int n_levels, int texture_band, double lower, double upper, float* h_norm,
                                float* h_qnorm, float* h_center, double* h_scale, float* h_norm_lut, float* h_x_lut) {
    RSX_REQUIRE(h_hist && h_norm && h_qnorm && n_bands >= 1 && n_bands <= RSX_MAX_BANDS && (n_levels == 256 || n_levels == 65536),
                "rsx_raster_stats: bad arguments");
    static thread_local int64_t* cum = nullptr;
    static thread_local float *levels = nullptr, *f = nullptr;
    static thread_local int cap = 0;
    if (cap < n_levels) {
        delete[] cum, delete[] levels, delete[] f;
        cum = new int64_t[n_levels], levels = new float[n_levels], f = new float[n_levels];
        cap = n_levels;
        for (int v = 0; v < n_levels; ++v) levels[v] = (float)v;
    }
    h_qnorm[0] = 0.f, h_qnorm[1] = 1.f, h_qnorm[2] = 1.f;
    for (int b = 0; b < n_bands; ++b) {
        const int64_t* h = h_hist + (size_t)b * n_levels;
        int64_t run = 0;
        for (int v = 0; v < n_levels; ++v) run += h[v], cum[v] = run;
        RSX_REQUIRE(run > 0, "rsx_raster_stats: empty histogram (band %d)", b);
        const Order raw{cum, levels, n_levels, run};
        const float lo = raw.percentile32(lower), hi = raw.percentile32(upper);
        const volatile float range = hi - lo;
        const float den = range + 1e-10f;  // hi - lo + 1e-10 in float32 (indices.py:46)
        h_norm[3 * b] = lo, h_norm[3 * b + 1] = hi, h_norm[3 * b + 2] = den;
        for (int v = 0; v < n_levels; ++v) {  // robust_normalize of every level
            const float c = fminf(fmaxf(levels[v], lo), hi);
            const volatile float num = c - lo;
            f[v] = num / den;
        }
        if (h_norm_lut) memcpy(h_norm_lut + (size_t)b * n_levels, f, sizeof(float) * n_levels);
        const Order nb{cum, f, n_levels, run};
        if (b == texture_band) {
            const float lo2 = nb.percentile32(lower), hi2 = nb.percentile32(upper);
            const volatile float r2 = hi2 - lo2;
            h_qnorm[0] = lo2, h_qnorm[1] = hi2, h_qnorm[2] = r2 + 1e-10f;
        }
        if (h_center && h_scale) {
            const float center = nb.median32();
            double s = nb.percentile64(75.0) - nb.percentile64(25.0);
            if (s < 10.0 * 2.220446049250313e-16) s = 1.0;  // sklearn _handle_zeros_in_scale
            h_center[b] = center, h_scale[b] = s;
            if (h_x_lut)
                for (int v = 0; v < n_levels; ++v) {
                    const volatile float d = f[v] - center;
                    h_x_lut[(size_t)b * n_levels + v] = (float)((double)d / s);
                }
        }
    }
    return RSX_OK;
}
