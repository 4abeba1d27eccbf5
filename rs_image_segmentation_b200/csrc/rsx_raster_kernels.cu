// Kernels that stream the pixel-interleaved raster: K1 histograms, K2 fused normalise + spectral
// indices (+ GLCM quantisation), K3 PCA moments (warp-shuffle tree reductions) and projection.
#include "rsx_raster.cuh"

// ============================================================================ K1: histograms
// A thread takes the B consecutive 32-bit words of four pixels, so the band of every byte is a compile-time constant and a
// sample costs a byte extract and a shared atomic with an immediate offset (3-4 instructions; 14.5 with a run-time band
// per byte: the kernel was bound by instruction issue, 68 % of the slots busy).  All lanes of a warp then count the same
// band at once and neighbouring pixels mostly hold similar values, so lane l counts into copy l % NCOPY of the histograms:
// at most 32 / NCOPY lanes can meet on one address.
template <int B, int NCOPY>
__global__ void __launch_bounds__(256) hist_u8_kernel(const uint8_t* __restrict__ raster, int64_t n_px, uint32_t* __restrict__ hist) {
    using RT = RasterTiles<uint8_t, B, 2>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t* sh = reinterpret_cast<uint32_t*>(smem + RT::SMEM_BYTES);  // [NCOPY][B][256]
    for (int i = threadIdx.x; i < NCOPY * B * 256; i += 256) sh[i] = 0;
    __syncthreads();
    uint32_t* mine = sh + (threadIdx.x % NCOPY) * (B * 256);
    for_each_tile<uint8_t, B, 2>(raster, n_px, smem, [&](const uint32_t* words, int64_t, int npx) {
        const int nquads = npx >> 2;
        for (int g = threadIdx.x; g < nquads; g += 256) {
            const uint32_t* wq = words + g * B;
#pragma unroll
            for (int k = 0; k < B; ++k) {
                const uint32_t w = wq[k];
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(&mine[((4 * k + j) % B) * 256 + ((w >> (8 * j)) & 0xffu)], 1u);
            }
        }
        // the last tile's npx % 4 pixels, byte by byte
        const uint8_t* bytes = reinterpret_cast<const uint8_t*>(words);
        for (int s = nquads * 4 * B + threadIdx.x; s < npx * B; s += 256) atomicAdd(&mine[(s % B) * 256 + bytes[s]], 1u);
    });
    __syncthreads();
    for (int i = threadIdx.x; i < B * 256; i += 256) {
        uint32_t s = 0;
#pragma unroll
        for (int c = 0; c < NCOPY; ++c) s += sh[c * B * 256 + i];
        if (s) atomicAdd(&hist[i], s);
    }
}

// uint16: the low `nb` values of every band are counted in shared memory - 16-bit counters, two per 32-bit word, bumped with
// 32-bit shared atomics and flushed to the global [B][65536] histogram before any of them can wrap (every < 65536 pixels of
// the CTA); values >= nb (rare for reflectance data) go to global atomics directly.  512-pixel tiles leave ~200 KB for the
// counters (nb ~ 7.8 k for 13 bands).  A lane reads one 32-bit word = 2 samples of neighbouring bands, so lanes that can
// collide on a bin are B/2 words apart.
template <int B>
__global__ void __launch_bounds__(256) hist_u16_kernel(const uint16_t* __restrict__ raster, int64_t n_px, uint32_t* __restrict__ hist, int nb) {
    using RT = RasterTiles<uint16_t, B, 2, 1>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + ((RT::SMEM_BYTES + 15) & ~15));
    const int n_words = B * nb / 2;
    for (int i = threadIdx.x; i < n_words; i += 256) cnt[i] = 0;
    __syncthreads();
    auto flush = [&]() {
        __syncthreads();
        for (int w = threadIdx.x; w < n_words; w += 256) {
            const uint32_t x = cnt[w];
            if (x) {
                cnt[w] = 0;
                const int i = 2 * w, band = i / nb, v = i - band * nb;  // nb is even: both halves belong to one band
                if (x & 0xffffu) atomicAdd(&hist[(size_t)band * 65536 + v], x & 0xffffu);
                if (x >> 16) atomicAdd(&hist[(size_t)band * 65536 + v + 1], x >> 16);
            }
        }
        __syncthreads();
    };
    int since_flush = 0;  // pixels (= samples per band) counted since the last flush; CTA-uniform
    for_each_tile<uint16_t, B, 2, 1>(raster, n_px, smem, [&](const uint32_t* words, int64_t, int npx) {
        if (since_flush + npx > 65535) {
            flush();
            since_flush = 0;
        }
        since_flush += npx;
        const int nhalf = npx * B;
        const int nwords = (nhalf + 1) >> 1;
        auto count = [&](int band, unsigned v) {
            if ((int)v < nb) {
                const int i = band * nb + (int)v;
                atomicAdd(&cnt[i >> 1], (i & 1) ? 65536u : 1u);
            } else {
                atomicAdd(&hist[(size_t)band * 65536 + v], 1u);
            }
        };
        for (int wi = threadIdx.x; wi < nwords; wi += 256) {
            const uint32_t w = words[wi];
            int band = (wi * 2) % B;
            count(band, w & 0xffffu);
            band = band + 1 == B ? 0 : band + 1;
            if (wi * 2 + 1 < nhalf) count(band, w >> 16);
        }
    });
    flush();
}

template <typename K>
static int set_smem(K kernel, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        rsx_set_error("cudaFuncSetAttribute(%d bytes): %s", bytes, cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    return RSX_OK;
}

static int persistent_grid(int64_t n_tiles, int ctas_per_sm) { return (int)min(n_tiles, (int64_t)rsx_num_sms() * ctas_per_sm); }

extern "C" int rsx_hist_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, uint32_t* d_hist, rsx_stream_t stream) {
    RSX_REQUIRE(d_raster && d_hist && n_px > 0, "rsx_hist_u8: bad arguments");
    RSX_REQUIRE(((uintptr_t)d_raster & 15) == 0, "rsx_hist_u8: raster must be 16-byte aligned");
#define LAUNCH(BB)                                                                                         \
    {                                                                                                      \
        using RT = RasterTiles<uint8_t, BB, 2>;                                                            \
        constexpr int NCOPY = BB <= 8 ? 8 : 4;                                                             \
        int smem = RT::SMEM_BYTES + NCOPY * BB * 256 * 4;                                                  \
        if (int rc = set_smem(hist_u8_kernel<BB, NCOPY>, smem)) return rc;                                 \
        int grid = persistent_grid(ceil_div(n_px, (int64_t)RT::TILE_PX), 2);                               \
        hist_u8_kernel<BB, NCOPY><<<grid, 256, smem, (cudaStream_t)stream>>>(d_raster, n_px, d_hist);      \
    }
    RSX_DISPATCH_BANDS(n_bands, LAUNCH)
#undef LAUNCH
    return rsx_check_launch("rsx_hist_u8");
}

extern "C" int rsx_hist_u16(const uint16_t* d_raster, int64_t n_px, int n_bands, uint32_t* d_hist, rsx_stream_t stream) {
    RSX_REQUIRE(d_raster && d_hist && n_px > 0, "rsx_hist_u16: bad arguments");
    RSX_REQUIRE(((uintptr_t)d_raster & 15) == 0, "rsx_hist_u16: raster must be 16-byte aligned");
#define LAUNCH(BB)                                                                                    \
    {                                                                                                 \
        using RT = RasterTiles<uint16_t, BB, 2, 1>;                                                   \
        const int tile_bytes = (RT::SMEM_BYTES + 15) & ~15;                                           \
        int nb = ((226 * 1024 - tile_bytes) / (2 * BB)) & ~1;                                         \
        nb = nb > 65536 ? 65536 : nb;                                                                 \
        const int smem = tile_bytes + BB * nb * 2;                                                    \
        if (int rc = set_smem(hist_u16_kernel<BB>, smem)) return rc;                                  \
        int grid = persistent_grid(ceil_div(n_px, (int64_t)RT::TILE_PX), 1);                          \
        hist_u16_kernel<BB><<<grid, 256, smem, (cudaStream_t)stream>>>(d_raster, n_px, d_hist, nb);   \
    }
    RSX_DISPATCH_BANDS(n_bands, LAUNCH)
#undef LAUNCH
    return rsx_check_launch("rsx_hist_u16");
}

// ============================================================================ K1b: order statistics on the device (uint8 rasters)
// rsx_raster_stats (rsx_core.cu: numpy's percentile arithmetic restated for the host) once more for the device, operation by
// operation with explicitly rounded intrinsics, so that the K1 histograms never have to visit the host between two kernels: one
// warp per band, the per-level loops spread over the lanes, the few scalar percentile evaluations done redundantly by every lane.
// The results stay in a device block that K2 / K3 read; tests/test_gpu_kernels.py holds them bit for bit against rsx_raster_stats.
struct RsxDevStats {
    float norm[RSX_MAX_BANDS][3];  // lo, hi, den of robust_normalize
    float qnorm[4];                // second robust_normalize of the texture band (+ pad)
    float center[RSX_MAX_BANDS];
    double scale[RSX_MAX_BANDS];
    float x_lut[RSX_MAX_BANDS][256];  // RobustScaler value of every grey level
};

struct DevOrder {
    const long long* cum;  // inclusive cumulative counts (shared memory)
    const float* values;   // value of every level (shared memory)
    long long n;
    __device__ float at(long long k) const {
        k = k < 0 ? 0 : (k >= n ? n - 1 : k);
        int lo = 0, hi = 256;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cum[mid] > k) hi = mid; else lo = mid + 1;
        }
        return values[lo < 256 ? lo : 255];
    }
    __device__ float percentile32(double q) const {
        const float q32 = __fdiv_rn((float)q, 100.0f);
        const float nm1 = __ll2float_rn(n - 1);
        const float vi = __fmul_rn(nm1, q32);
        float prev = floorf(vi), nxt = __fadd_rn(prev, 1.0f);
        if (vi >= nm1) prev = nxt = -1.0f;
        if (vi < 0.0f) prev = nxt = 0.0f;
        const long long pi = (long long)prev, ni = (long long)nxt;
        const float a = at(pi >= 0 ? pi : n - 1), b = at(ni >= 0 ? ni : n - 1);
        if (a == b) return a;
        const float t = __double2float_rn(__dsub_rn((double)vi, (double)pi));
        const float diff = __fsub_rn(b, a);
        float out = __fadd_rn(a, __fmul_rn(diff, t));
        if (t >= 0.5f) out = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, t)));
        return out;
    }
    __device__ double percentile64(double q) const {
        const double qq = __ddiv_rn(q, 100.0);
        const double vi = __dmul_rn((double)(n - 1), qq);
        double prev = floor(vi), nxt = __dadd_rn(prev, 1.0);
        if (vi >= (double)(n - 1)) prev = nxt = -1.0;
        if (vi < 0.0) prev = nxt = 0.0;
        const long long pi = (long long)prev, ni = (long long)nxt;
        const float a = at(pi >= 0 ? pi : n - 1), b = at(ni >= 0 ? ni : n - 1);
        if (a == b) return (double)a;
        const double t = __dsub_rn(vi, (double)pi);
        const float diff = __fsub_rn(b, a);
        double out = __dadd_rn((double)a, __dmul_rn((double)diff, t));
        if (t >= 0.5) out = __dsub_rn((double)b, __dmul_rn((double)diff, __dsub_rn(1.0, t)));
        return out;
    }
    __device__ float median32() const {
        if (n & 1) return at(n / 2);
        return __fdiv_rn(__fadd_rn(at(n / 2 - 1), at(n / 2)), 2.0f);
    }
};

// hist: [B][256] counters, uint32 (one rank) or int64 (after the all-reduce); one warp per band
__global__ void __launch_bounds__(32 * RSX_MAX_BANDS) raster_stats_u8_kernel(const void* __restrict__ hist, int is64, int texture_band, double lower,
                                                                             double upper, RsxDevStats* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char stats_smem[];  // [B][256] int64 cumulative counts, [B][256] float values, [256] levels
    const int n_bands = blockDim.x >> 5;
    long long(*cum)[256] = reinterpret_cast<long long(*)[256]>(stats_smem);
    float(*f)[256] = reinterpret_cast<float(*)[256]>(stats_smem + (size_t)n_bands * 256 * 8);
    float* levels = reinterpret_cast<float*>(stats_smem + (size_t)n_bands * 256 * 12);
    const int b = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int v = threadIdx.x; v < 256; v += blockDim.x) levels[v] = (float)v;
    if (threadIdx.x == 0 && (texture_band < 0 || texture_band >= n_bands)) out->qnorm[0] = 0.f, out->qnorm[1] = 1.f, out->qnorm[2] = 1.f;
    // inclusive cumulative counts: 8 consecutive levels per lane, then a warp scan of the lane totals
    long long c[8], run = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int v = lane * 8 + k;
        const long long h = is64 ? reinterpret_cast<const long long*>(hist)[b * 256 + v] : (long long)reinterpret_cast<const unsigned*>(hist)[b * 256 + v];
        run += h, c[k] = run;
    }
    long long incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const long long before = incl - run;
#pragma unroll
    for (int k = 0; k < 8; ++k) cum[b][lane * 8 + k] = before + c[k];
    const long long n = __shfl_sync(0xffffffffu, incl, 31);
    __syncthreads();
    if (n <= 0) return;  // empty histogram: the caller's tables stay as they were (rsx_raster_stats reports it on the host)
    const DevOrder raw{cum[b], levels, n};
    const float lo = raw.percentile32(lower), hi = raw.percentile32(upper);
    const float den = __fadd_rn(__fsub_rn(hi, lo), 1e-10f);
    for (int v = lane; v < 256; v += 32) f[b][v] = __fdiv_rn(__fsub_rn(fminf(fmaxf(levels[v], lo), hi), lo), den);
    __syncwarp();
    const DevOrder nb{cum[b], f[b], n};
    if (b == texture_band) {
        const float lo2 = nb.percentile32(lower), hi2 = nb.percentile32(upper);
        if (lane == 0) out->qnorm[0] = lo2, out->qnorm[1] = hi2, out->qnorm[2] = __fadd_rn(__fsub_rn(hi2, lo2), 1e-10f);
    }
    const float center = nb.median32();
    double sc = __dsub_rn(nb.percentile64(75.0), nb.percentile64(25.0));
    if (sc < 10.0 * 2.220446049250313e-16) sc = 1.0;
    if (lane == 0) out->norm[b][0] = lo, out->norm[b][1] = hi, out->norm[b][2] = den, out->center[b] = center, out->scale[b] = sc;
    for (int v = lane; v < 256; v += 32) out->x_lut[b][v] = __double2float_rn(__ddiv_rn((double)__fsub_rn(f[b][v], center), sc));
}

extern "C" int64_t rsx_raster_stats_device_bytes(void) { return (int64_t)sizeof(RsxDevStats); }
extern "C" int64_t rsx_raster_stats_device_lut_offset(void) { return (int64_t)offsetof(RsxDevStats, x_lut); }

extern "C" int rsx_raster_stats_u8_device(const void* d_hist, int hist_is_int64, int n_bands, int texture_band, double lower, double upper, void* d_stats,
                                          rsx_stream_t stream) {
    RSX_REQUIRE(d_hist && d_stats && n_bands >= 1 && n_bands <= RSX_MAX_BANDS, "rsx_raster_stats_u8_device: bad arguments");
    RSX_REQUIRE(((uintptr_t)d_stats & 15) == 0 && ((uintptr_t)d_hist & 7) == 0, "rsx_raster_stats_u8_device: buffers must be aligned");
    const int smem = n_bands * 256 * 12 + 1024;
    if (int rc = set_smem(raster_stats_u8_kernel, smem)) return rc;
    raster_stats_u8_kernel<<<1, 32 * n_bands, smem, (cudaStream_t)stream>>>(d_hist, hist_is_int64, texture_band, lower, upper,
                                                                           reinterpret_cast<RsxDevStats*>(d_stats));
    return rsx_check_launch("rsx_raster_stats_u8_device");
}

// ============================================================================ K2: fused normalise + indices
struct IndexParams {
    NormParam norm[5];  // blue, green, red, nir, swir1 (already gathered through band_map)
    int band[5];
    float evi_L, evi_C1, evi_C2, evi_G;
    NormParam qnorm;
    float q_scale;  // levels - 1
    // uint8 rasters: level remap of the five bands applied before the normalisation (identity by default); carries the
    // fused stage-1 chain gain/bias -> min-max stretch -> uint8 (modules/features/preprocessing.py:54-125)
    uint8_t remap[5][256];
};

// the seven maps of scripts/2_feature_extraction.py:63-73, in RSX index order
__device__ __forceinline__ void seven_indices(float blue, float green, float red, float nir, float swir, const IndexParams& P, float (&o)[7]) {
    // ndvi (indices.py:62-69)
    {
        float den = f_add(nir, red);
        o[0] = f_clip(den > 0.001f ? f_div(f_sub(nir, red), den) : 0.f, -1.f, 1.f);
    }
    // evi (indices.py:86-93)
    {
        float den = f_add(f_sub(f_add(nir, f_mul(P.evi_C1, red)), f_mul(P.evi_C2, blue)), P.evi_L);
        o[1] = f_clip(den > 0.001f ? f_div(f_mul(P.evi_G, f_sub(nir, red)), den) : 0.f, -1.f, 1.f);
    }
    // msavi (indices.py:109-112)
    {
        float t = f_add(f_mul(2.f, nir), 1.f);
        float v = f_mul(f_sub(t, f_sqrt(f_sub(f_mul(t, t), f_mul(8.f, f_sub(nir, red))))), 0.5f);  // x / 2 == x * 0.5 exactly
        o[2] = v != v ? v : f_clip(v, -1.f, 1.f);
    }
    // ndwi (indices.py:128-135)
    {
        float den = f_add(green, nir);
        o[3] = f_clip(den > 0.001f ? f_div(f_sub(green, nir), den) : 0.f, -1.f, 1.f);
    }
    // mndwi (indices.py:150-156)
    {
        float den = f_add(green, swir);
        o[4] = f_clip(den > 0.001f ? f_div(f_sub(green, swir), den) : 0.f, -1.f, 1.f);
    }
    // ndbi (indices.py:171-177)
    {
        float den = f_add(swir, nir);
        o[5] = f_clip(den > 0.001f ? f_div(f_sub(swir, nir), den) : 0.f, -1.f, 1.f);
    }
    // bsi (indices.py:194-201)
    {
        float a = f_add(swir, red), b = f_add(nir, blue);
        float den = f_add(a, b);
        o[6] = f_clip(den > 0.001f ? f_div(f_sub(a, b), den) : 0.f, -1.f, 1.f);
    }
}

template <typename T, int B>
__global__ void __launch_bounds__(256) indices_fused_kernel(const T* __restrict__ raster, int64_t n_px, IndexParams P, float* __restrict__ out,
                                                            int64_t plane_stride, uint32_t* __restrict__ minmax, uint8_t* __restrict__ quant,
                                                            const RsxDevStats* __restrict__ ds) {
    using RT = RasterTiles<T, B, 3>;
    constexpr int PXT = RT::PXT;
    extern __shared__ __align__(128) unsigned char smem[];
    float mn[7], mx[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) mn[k] = INFINITY, mx[k] = -INFINITY;
    // uint8 rasters: the normalised value of every grey level of the five bands (and the quantised NIR level) is tabulated
    // once per CTA with the same float32 expressions, so the per-pixel work has 6 IEEE divisions instead of 12
    __shared__ float nlut[sizeof(T) == 1 ? 5 * 256 : 1];
    __shared__ uint8_t qlut[sizeof(T) == 1 ? 256 : 1];
    if constexpr (sizeof(T) == 1) {
        // ds: the normalisation parameters come from the device block raster_stats_u8_kernel filled (no host round trip)
        for (int i = threadIdx.x; i < 5 * 256; i += 256) {
            const int k = i >> 8;
            const NormParam np = ds ? NormParam{ds->norm[P.band[k]][0], ds->norm[P.band[k]][1], ds->norm[P.band[k]][2]} : P.norm[k];
            nlut[i] = norm_apply((float)P.remap[k][i & 255], np);
        }
        __syncthreads();
        const NormParam qn = ds ? NormParam{ds->qnorm[0], ds->qnorm[1], ds->qnorm[2]} : P.qnorm;
        for (int i = threadIdx.x; i < 256; i += 256) qlut[i] = (uint8_t)(int)f_mul(norm_apply(nlut[3 * 256 + i], qn), P.q_scale);
        __syncthreads();
    }

    for_each_tile<T, B, 3>(raster, n_px, smem, [&](const uint32_t* words, int64_t px0, int npx) {
        const T* samples = reinterpret_cast<const T*>(words);
#pragma unroll
        for (int sub = 0; sub < RT::SUB; ++sub) {
            const int g = sub * 256 + threadIdx.x;
            const int lp = g * PXT;  // first pixel of this group inside the tile
            if (lp >= npx) continue;
            float o[PXT][7];
            uint8_t q[PXT];
#pragma unroll
            for (int p = 0; p < PXT; ++p) {
                const T* px = samples + (lp + p) * B;
                float nb[5];
                if constexpr (sizeof(T) == 1) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) nb[k] = nlut[k * 256 + px[P.band[k]]];
                    q[p] = qlut[px[P.band[3]]];
                } else {
#pragma unroll
                    for (int k = 0; k < 5; ++k) nb[k] = norm_apply((float)px[P.band[k]], P.norm[k]);
                    q[p] = (uint8_t)(int)f_mul(norm_apply(nb[3], P.qnorm), P.q_scale);
                }
                seven_indices(nb[0], nb[1], nb[2], nb[3], nb[4], P, o[p]);
            }
            const int64_t gp = px0 + lp;
            if (lp + PXT <= npx) {
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    float* dst = out + k * plane_stride + gp;
                    if constexpr (PXT == 4)
                        stg_stream4(dst, make_float4(o[0][k], o[1][k], o[2][k], o[3][k]));
                    else
                        *reinterpret_cast<float2*>(dst) = make_float2(o[0][k], o[1][k]);
#pragma unroll
                    for (int p = 0; p < PXT; ++p) mn[k] = fminf(mn[k], o[p][k]), mx[k] = fmaxf(mx[k], o[p][k]);
                }
                if (quant) {
                    if constexpr (PXT == 4)
                        *reinterpret_cast<uint32_t*>(quant + gp) = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24);
                    else
                        *reinterpret_cast<uint16_t*>(quant + gp) = (uint16_t)(q[0] | (q[1] << 8));
                }
            } else {
#pragma unroll
                for (int p = 0; p < PXT; ++p)
                    if (lp + p < npx) {
#pragma unroll
                        for (int k = 0; k < 7; ++k) {
                            out[k * plane_stride + gp + p] = o[p][k];
                            mn[k] = fminf(mn[k], o[p][k]), mx[k] = fmaxf(mx[k], o[p][k]);
                        }
                        if (quant) quant[gp + p] = q[p];
                    }
            }
        }
    });
    if (minmax) {
#pragma unroll
        for (int k = 0; k < 7; ++k) warp_minmax_commit(mn[k], mx[k], minmax + 2 * k);
    }
}

template <typename T>
static int indices_fused_impl(const T* d_raster, int64_t n_px, int n_bands, const int* band_map, const float* h_norm, const float* evi,
                              float* d_indices, int64_t plane_stride, uint32_t* d_minmax, uint8_t* d_quant, const float* h_qnorm, int levels,
                              const uint8_t* h_remap, rsx_stream_t stream, const RsxDevStats* d_stats = nullptr) {
    RSX_REQUIRE(d_raster && band_map && (h_norm || d_stats) && evi && d_indices && n_px > 0, "rsx_indices_fused: bad arguments");
    RSX_REQUIRE(!d_stats || sizeof(T) == 1, "rsx_indices_fused: device statistics are defined for uint8 rasters");
    RSX_REQUIRE(!h_remap || sizeof(T) == 1, "rsx_indices_fused: a level remap is only defined for uint8 rasters");
    RSX_REQUIRE(((uintptr_t)d_raster & 15) == 0 && ((uintptr_t)d_indices & 15) == 0 && (plane_stride & 3) == 0 && plane_stride >= n_px,
                "rsx_indices_fused: raster/planes must be 16-byte aligned, plane_stride a multiple of 4 and >= n_px");
    RSX_REQUIRE(!d_quant || ((h_qnorm || d_stats) && levels >= 2 && levels <= 256 && ((uintptr_t)d_quant & 3) == 0), "rsx_indices_fused: bad quantisation arguments");
    IndexParams P;
    for (int k = 0; k < 5; ++k) {
        int b = band_map[k];
        RSX_REQUIRE(b >= 0 && b < n_bands, "rsx_indices_fused: band_map[%d]=%d out of range", k, b);
        P.band[k] = b;
        P.norm[k] = h_norm ? NormParam{h_norm[3 * b], h_norm[3 * b + 1], h_norm[3 * b + 2]} : NormParam{0.f, 1.f, 1.f};
        for (int v = 0; v < 256; ++v) P.remap[k][v] = h_remap ? h_remap[b * 256 + v] : (uint8_t)v;
    }
    P.evi_L = evi[0], P.evi_C1 = evi[1], P.evi_C2 = evi[2], P.evi_G = evi[3];
    P.qnorm = (d_quant && h_qnorm) ? NormParam{h_qnorm[0], h_qnorm[1], h_qnorm[2]} : NormParam{0.f, 1.f, 1.f};
    P.q_scale = (float)(levels - 1);
#define LAUNCH(BB)                                                                                                                  \
    {                                                                                                                               \
        using RT = RasterTiles<T, BB, 3>;                                                                                           \
        if (int rc = set_smem(indices_fused_kernel<T, BB>, RT::SMEM_BYTES)) return rc;                                              \
        int grid = persistent_grid(ceil_div(n_px, (int64_t)RT::TILE_PX), RT::SMEM_BYTES > 110000 ? 1 : 2);                          \
        indices_fused_kernel<T, BB><<<grid, 256, RT::SMEM_BYTES, (cudaStream_t)stream>>>(d_raster, n_px, P, d_indices, plane_stride, d_minmax, d_quant, \
                                                                                         d_stats);                                            \
    }
    RSX_DISPATCH_BANDS(n_bands, LAUNCH)
#undef LAUNCH
    return rsx_check_launch("rsx_indices_fused");
}

extern "C" int rsx_indices_fused_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, const int* band_map, const float* h_norm, const float* evi,
                                    float* d_indices, int64_t plane_stride, uint32_t* d_minmax, uint8_t* d_quant, const float* h_qnorm, int levels,
                                    const uint8_t* h_remap, rsx_stream_t stream) {
    return indices_fused_impl<uint8_t>(d_raster, n_px, n_bands, band_map, h_norm, evi, d_indices, plane_stride, d_minmax, d_quant, h_qnorm, levels, h_remap,
                                       stream);
}
// normalisation parameters from the device block of rsx_raster_stats_u8_device instead of host arrays
extern "C" int rsx_indices_fused_u8_dev(const uint8_t* d_raster, int64_t n_px, int n_bands, const int* band_map, const void* d_stats, const float* evi,
                                        float* d_indices, int64_t plane_stride, uint32_t* d_minmax, uint8_t* d_quant, int levels, rsx_stream_t stream) {
    RSX_REQUIRE(d_stats, "rsx_indices_fused_u8_dev: null statistics block");
    return indices_fused_impl<uint8_t>(d_raster, n_px, n_bands, band_map, nullptr, evi, d_indices, plane_stride, d_minmax, d_quant, nullptr, levels, nullptr,
                                       stream, reinterpret_cast<const RsxDevStats*>(d_stats));
}
extern "C" int rsx_indices_fused_u16(const uint16_t* d_raster, int64_t n_px, int n_bands, const int* band_map, const float* h_norm, const float* evi,
                                     float* d_indices, int64_t plane_stride, uint32_t* d_minmax, uint8_t* d_quant, const float* h_qnorm, int levels,
                                     rsx_stream_t stream) {
    return indices_fused_impl<uint16_t>(d_raster, n_px, n_bands, band_map, h_norm, evi, d_indices, plane_stride, d_minmax, d_quant, h_qnorm, levels, nullptr,
                                        stream);
}

// ============================================================================ K3: PCA
// X = RobustScaler().fit_transform(normalised bands) as sklearn computes it for float32 input:
//   X -= center_ (float32);  X /= scale_ where scale_ is FLOAT64 (np.nanpercentile of a (25.0, 75.0) tuple), so the
//   division is evaluated in float64 and rounded to float32 (sklearn/preprocessing/_data.py:1738-1743,1782-1784).
// uint8 rasters: X is a function of the 256 grey levels of each band, so the host tabulates it once (same numpy
// expressions as the reference) and the kernels look it up in shared memory: bit-exact by construction, no divisions.
// uint16 rasters: evaluated arithmetically with the same operation order.
template <int B>
struct PcaParams {
    NormParam norm[B];
    float center[B];
    double scale[B];
    double inv_scale[B], inv_den[B];  // reciprocals for the uint16 arithmetic path
    const float* lut;  // device [B][256] (uint8 rasters), [B][65536] (uint16 rasters, optional) or nullptr
};

template <typename T, int B>
__device__ __forceinline__ void scaled_pixel(const int (&raw)[B], const PcaParams<B>& P, const float* lut_s, float (&x)[B]) {
#pragma unroll
    for (int b = 0; b < B; ++b) {
        if constexpr (sizeof(T) == 1)
            x[b] = lut_s[b * 256 + raw[b]];
        else if (P.lut)
            x[b] = __ldg(P.lut + b * 65536 + raw[b]);  // built by pca_lut_u16_kernel with the expression below
        else {
            // uint16 without a table: the two divisions (float32 by den, float64 by scale_) become float64 multiplications
            // by the reciprocals.  The product is within 1 ulp(double) of the quotient, so the float32 result differs from
            // sklearn's only when the quotient sits within 2^-29 (relative) of a float32 rounding boundary - a 1-ulp
            // difference in ~1 sample per 10^8, far below the 1e-5 bar of the PCA outputs (the index maps, which must be
            // bit exact, never use this path).
            const float num = f_sub(f_clip((float)raw[b], P.norm[b].lo, P.norm[b].hi), P.norm[b].lo);
            const float n = __double2float_rn(__dmul_rn((double)num, P.inv_den[b]));
            x[b] = __double2float_rn(__dmul_rn((double)f_sub(n, P.center[b]), P.inv_scale[b]));
        }
    }
}

template <typename T, int B>
__device__ __forceinline__ const float* stage_lut(const PcaParams<B>& P, float* lut_s) {
    if constexpr (sizeof(T) == 1) {
        for (int i = threadIdx.x; i < B * 256; i += blockDim.x) lut_s[i] = P.lut[i];
        __syncthreads();
    }
    return lut_s;
}

template <int B>
struct Moments {
    static constexpr int NPAIR = B * (B + 1) / 2;
    static constexpr int M = B + NPAIR;
    static constexpr int NSPLIT = B <= 8 ? 1 : 4;  // thread groups sharing the pair set (register budget)
    static constexpr int NACC = (NPAIR + NSPLIT - 1) / NSPLIT;
};

template <int B, int PART>
__device__ __forceinline__ void moments_accumulate(const float (&x)[B], double (&sum)[B], double (&acc)[Moments<B>::NACC]) {
    constexpr int NS = Moments<B>::NSPLIT;
    if (PART == 0) {
#pragma unroll
        for (int b = 0; b < B; ++b) sum[b] += (double)x[b];
    }
    int idx = 0;
#pragma unroll
    for (int a = 0; a < B; ++a)
#pragma unroll
        for (int b = a; b < B; ++b) {
            if (idx % NS == PART) acc[idx / NS] = fma((double)x[a], (double)x[b], acc[idx / NS]);
            ++idx;
        }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
constexpr int lut_bytes(int B) { return sizeof(T) == 1 ? B * 256 * 4 : 0; }

// X of every 16-bit level of every band, same arithmetic as the per-sample path (one IEEE float32 division, one float64
// division per entry instead of per sample)
__global__ void __launch_bounds__(256) pca_lut_u16_kernel(const float* __restrict__ norm, const float* __restrict__ center, const double* __restrict__ scale,
                                                          int B, float* __restrict__ lut) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B * 65536) return;
    const int b = i >> 16, v = i & 0xffff;
    const NormParam np_{norm[3 * b], norm[3 * b + 1], norm[3 * b + 2]};
    lut[i] = __double2float_rn(__ddiv_rn((double)f_sub(norm_apply((float)v, np_), center[b]), scale[b]));
}

extern "C" int rsx_pca_build_lut_u16(const float* d_norm, const float* d_center, const double* d_scale, int n_bands, float* d_lut, rsx_stream_t stream) {
    RSX_REQUIRE(d_norm && d_center && d_scale && d_lut && n_bands >= 1 && n_bands <= RSX_MAX_BANDS, "rsx_pca_build_lut_u16: bad arguments");
    pca_lut_u16_kernel<<<n_bands * 256, 256, 0, (cudaStream_t)stream>>>(d_norm, d_center, d_scale, n_bands, d_lut);
    return rsx_check_launch("pca_lut_u16");
}

template <typename T, int B>
__global__ void __launch_bounds__(256) pca_moments_kernel(const T* __restrict__ raster, int64_t n_px, const __grid_constant__ PcaParams<B> P,
                                                          double* __restrict__ scratch) {
    using MM = Moments<B>;
    constexpr int MSUB = MM::NSPLIT > 1 ? 1 : 4;  // B > 8: one pixel group per thread and tile (small tiles, three CTAs per SM)
    using RT = RasterTiles<T, B, 3, MSUB>;
    constexpr int PXT = RT::PXT;
    constexpr int NS = MM::NSPLIT;
    constexpr int TPP = 256 / NS;  // threads per part
    extern __shared__ __align__(128) unsigned char smem[];
    double* red = reinterpret_cast<double*>(smem + RT::SMEM_BYTES);                      // [8 warps][M]
    float* lut_s = reinterpret_cast<float*>(smem + RT::SMEM_BYTES + 8 * MM::M * 8);      // [B][256]
    stage_lut<T, B>(P, lut_s);
    const int part = threadIdx.x / TPP, lt = threadIdx.x % TPP;

    double sum[B], acc[MM::NACC];
#pragma unroll
    for (int b = 0; b < B; ++b) sum[b] = 0.0;
#pragma unroll
    for (int i = 0; i < MM::NACC; ++i) acc[i] = 0.0;

    // B > 8 (NS = 4 thread groups share the 91 pair sums of 13 bands): the scaled values of a chunk of 256 * PXT pixels are
    // computed ONCE by all threads into shared memory, planar [band][pixel]; each group then walks the chunk and accumulates its
    // quarter of the pairs.  (Scaling a uint16 sample costs two float64 multiplications and five conversions: with every group
    // scaling every pixel for itself the kernel spent 4/5 of its instructions there - 11.2 ms on the 120 Mpx x 13 tile.)
    constexpr int CH = 256 * PXT;
    float* xs = reinterpret_cast<float*>(smem + RT::SMEM_BYTES + 8 * MM::M * 8 + lut_bytes<T>(B));  // [B][CH], NS > 1 only
    for_each_tile<T, B, 3, MSUB>(raster, n_px, smem, [&](const uint32_t* words, int64_t, int npx) {
        if constexpr (NS == 1) {
            for (int g = lt; g * PXT < npx; g += TPP) {
                uint32_t w[B];
#pragma unroll
                for (int i = 0; i < B; ++i) w[i] = words[g * B + i];
                int raw[PXT][B];
                unpack_pixels<T, B>(w, raw);
#pragma unroll
                for (int p = 0; p < PXT; ++p) {
                    if (g * PXT + p < npx) {
                        float x[B];
                        scaled_pixel<T, B>(raw[p], P, lut_s, x);
                        moments_accumulate<B, 0>(x, sum, acc);
                    }
                }
            }
        } else {
            for (int c0 = 0; c0 < npx; c0 += CH) {
                const int n_ch = min(CH, npx - c0);
                {
                    const int g = c0 / PXT + threadIdx.x;  // this thread's pixel group of the chunk
                    if (g * PXT < npx) {
                        uint32_t w[B];
#pragma unroll
                        for (int i = 0; i < B; ++i) w[i] = words[g * B + i];
                        int raw[PXT][B];
                        unpack_pixels<T, B>(w, raw);
#pragma unroll
                        for (int p = 0; p < PXT; ++p) {
                            float x[B];
                            scaled_pixel<T, B>(raw[p], P, lut_s, x);
#pragma unroll
                            for (int b = 0; b < B; ++b) xs[b * CH + threadIdx.x * PXT + p] = x[b];
                        }
                    }
                }
                __syncthreads();
                for (int px = lt; px < n_ch; px += TPP) {
                    float x[B];
#pragma unroll
                    for (int b = 0; b < B; ++b) x[b] = xs[b * CH + px];
                    switch (part) {
                        case 0: moments_accumulate<B, 0>(x, sum, acc); break;
                        case 1: moments_accumulate<B, 1 % NS>(x, sum, acc); break;
                        case 2: moments_accumulate<B, 2 % NS>(x, sum, acc); break;
                        default: moments_accumulate<B, 3 % NS>(x, sum, acc); break;
                    }
                }
                __syncthreads();
            }
        }
    });

    // warp-shuffle tree per accumulator, then a fixed-order sum over the CTA's warps
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * MM::M; i += 256) red[i] = 0.0;
    __syncthreads();
#pragma unroll
    for (int b = 0; b < B; ++b) {
        double v = warp_sum(sum[b]);
        if (lane == 0 && part == 0) red[warp * MM::M + b] = v;
    }
#pragma unroll
    for (int i = 0; i < MM::NACC; ++i) {
        double v = warp_sum(acc[i]);
        int idx = i * NS + part;
        if (lane == 0 && idx < MM::NPAIR) red[warp * MM::M + B + idx] = v;
    }
    __syncthreads();
    for (int m = threadIdx.x; m < MM::M; m += 256) {
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += red[wv * MM::M + m];
        scratch[(size_t)blockIdx.x * MM::M + m] = s;
    }
}

__global__ void pca_moments_finish_kernel(const double* __restrict__ scratch, int n_blocks, int M, double* __restrict__ moments) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double s = 0.0;
    for (int i = 0; i < n_blocks; ++i) s += scratch[(size_t)i * M + m];
    moments[m] += s;
}

static int pca_grid() { return rsx_num_sms() * 4; }  // upper bound of the moments kernel's grid (the scratch is sized by it)
extern "C" int64_t rsx_pca_scratch_elems(int n_bands) { return (int64_t)pca_grid() * (n_bands + n_bands * (n_bands + 1) / 2); }

template <int B>
static void fill_pca_params(PcaParams<B>& P, const float* h_norm, const float* h_center, const double* h_scale, const float* d_lut) {
    for (int b = 0; b < B; ++b) {
        P.norm[b] = h_norm ? NormParam{h_norm[3 * b], h_norm[3 * b + 1], h_norm[3 * b + 2]} : NormParam{0.f, 1.f, 1.f};
        P.center[b] = h_center ? h_center[b] : 0.f;
        P.scale[b] = h_scale ? h_scale[b] : 1.0;
        P.inv_scale[b] = 1.0 / P.scale[b];
        P.inv_den[b] = h_norm ? 1.0 / (double)h_norm[3 * b + 2] : 1.0;
    }
    P.lut = d_lut;
}

template <typename T>
static int pca_moments_impl(const T* d_raster, int64_t n_px, int n_bands, const float* h_norm, const float* h_center, const double* h_scale,
                            const float* d_lut, double* d_moments, double* d_scratch, rsx_stream_t stream) {
    RSX_REQUIRE(d_raster && d_moments && d_scratch && n_px > 0, "rsx_pca_moments: bad arguments");
    RSX_REQUIRE(sizeof(T) == 1 ? d_lut != nullptr : (h_norm && h_center && h_scale), "rsx_pca_moments: missing scaling parameters");
    RSX_REQUIRE(((uintptr_t)d_raster & 15) == 0, "rsx_pca_moments: raster must be 16-byte aligned");
#define LAUNCH(BB)                                                                                                     \
    {                                                                                                                  \
        using MM = Moments<BB>;                                                                                        \
        using RT = RasterTiles<T, BB, 3, (MM::NSPLIT > 1 ? 1 : 4)>;                                                    \
        PcaParams<BB> P;                                                                                               \
        fill_pca_params<BB>(P, h_norm, h_center, h_scale, d_lut);                                                      \
        int smem = RT::SMEM_BYTES + 8 * MM::M * 8 + lut_bytes<T>(BB) + (MM::NSPLIT > 1 ? BB * 256 * RT::PXT * 4 : 0);  \
        if (int rc = set_smem(pca_moments_kernel<T, BB>, smem)) return rc;                                             \
        int per_sm = 1;                                                                                                \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pca_moments_kernel<T, BB>, 256, smem);                  \
        per_sm = max(1, min(per_sm, MM::NSPLIT > 1 ? 4 : 2));                                                          \
        int grid = (int)min((int64_t)rsx_num_sms() * per_sm, ceil_div(n_px, (int64_t)RT::TILE_PX));                    \
        pca_moments_kernel<T, BB><<<grid, 256, smem, (cudaStream_t)stream>>>(d_raster, n_px, P, d_scratch);            \
        if (int rc = rsx_check_launch("pca_moments")) return rc;                                                       \
        pca_moments_finish_kernel<<<ceil_div(MM::M, 128), 128, 0, (cudaStream_t)stream>>>(d_scratch, grid, MM::M, d_moments); \
    }
    RSX_DISPATCH_BANDS(n_bands, LAUNCH)
#undef LAUNCH
    return rsx_check_launch("pca_moments_finish");
}
extern "C" int rsx_pca_moments_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, const float* d_lut, double* d_moments, double* d_scratch,
                                  rsx_stream_t stream) {
    return pca_moments_impl<uint8_t>(d_raster, n_px, n_bands, nullptr, nullptr, nullptr, d_lut, d_moments, d_scratch, stream);
}
extern "C" int rsx_pca_moments_u16(const uint16_t* d_raster, int64_t n_px, int n_bands, const float* h_norm, const float* h_center,
                                   const double* h_scale, const float* d_lut16, double* d_moments, double* d_scratch, rsx_stream_t stream) {
    return pca_moments_impl<uint16_t>(d_raster, n_px, n_bands, h_norm, h_center, h_scale, d_lut16, d_moments, d_scratch, stream);
}

// ---- projection: Y = X @ components^T - mean @ components^T  (sklearn/decomposition/_base.py:151-159)
template <int B>
struct ProjParams {
    PcaParams<B> pca;
    float comp[B][B];  // [component][band]
    float mean_proj[B];
    int n_comp;
};

template <typename T, int B>
__global__ void __launch_bounds__(256) pca_project_kernel(const T* __restrict__ raster, int64_t n_px, const __grid_constant__ ProjParams<B> P,
                                                          float* __restrict__ out, int64_t plane_stride, uint32_t* __restrict__ minmax) {
    using RT = RasterTiles<T, B, 3>;
    constexpr int PXT = RT::PXT;
    extern __shared__ __align__(128) unsigned char smem[];
    float* lut_s = reinterpret_cast<float*>(smem + RT::SMEM_BYTES);
    stage_lut<T, B>(P.pca, lut_s);
    float mn[B], mx[B];
#pragma unroll
    for (int c = 0; c < B; ++c) mn[c] = INFINITY, mx[c] = -INFINITY;

    for_each_tile<T, B, 3>(raster, n_px, smem, [&](const uint32_t* words, int64_t px0, int npx) {
#pragma unroll
        for (int sub = 0; sub < RT::SUB; ++sub) {
            const int g = sub * 256 + threadIdx.x;
            const int lp = g * PXT;
            if (lp >= npx) continue;
            uint32_t w[B];
#pragma unroll
            for (int i = 0; i < B; ++i) w[i] = words[g * B + i];
            int raw[PXT][B];
            unpack_pixels<T, B>(w, raw);
            float x[PXT][B];
#pragma unroll
            for (int p = 0; p < PXT; ++p) scaled_pixel<T, B>(raw[p], P.pca, lut_s, x[p]);
            const int64_t gp = px0 + lp;
            const bool full = lp + PXT <= npx;
#pragma unroll
            for (int c = 0; c < B; ++c) {
                if (c < P.n_comp) {
                    float y[PXT];
#pragma unroll
                    for (int p = 0; p < PXT; ++p) {
                        float a = 0.f;
#pragma unroll
                        for (int b = 0; b < B; ++b) a = fmaf(x[p][b], P.comp[c][b], a);
                        y[p] = f_sub(a, P.mean_proj[c]);
                    }
                    float* dst = out + c * plane_stride + gp;
                    if (full) {
                        if constexpr (PXT == 4)
                            stg_stream4(dst, make_float4(y[0], y[1], y[2], y[3]));
                        else
                            *reinterpret_cast<float2*>(dst) = make_float2(y[0], y[1]);
#pragma unroll
                        for (int p = 0; p < PXT; ++p) mn[c] = fminf(mn[c], y[p]), mx[c] = fmaxf(mx[c], y[p]);
                    } else {
#pragma unroll
                        for (int p = 0; p < PXT; ++p)
                            if (lp + p < npx) {
                                dst[p] = y[p];
                                mn[c] = fminf(mn[c], y[p]), mx[c] = fmaxf(mx[c], y[p]);
                            }
                    }
                }
            }
        }
    });
    if (minmax) {
#pragma unroll
        for (int c = 0; c < B; ++c)
            if (c < P.n_comp) warp_minmax_commit(mn[c], mx[c], minmax + 2 * c);
    }
}

template <typename T>
static int pca_project_impl(const T* d_raster, int64_t n_px, int n_bands, const float* h_norm, const float* h_center, const double* h_scale,
                            const float* d_lut, const float* h_components, const float* h_mean_proj, int n_comp, float* d_out, int64_t plane_stride,
                            uint32_t* d_minmax, rsx_stream_t stream) {
    RSX_REQUIRE(d_raster && h_components && h_mean_proj && d_out && n_px > 0, "rsx_pca_project: bad arguments");
    RSX_REQUIRE(sizeof(T) == 1 ? d_lut != nullptr : (h_norm && h_center && h_scale), "rsx_pca_project: missing scaling parameters");
    RSX_REQUIRE(n_comp >= 1 && n_comp <= n_bands, "rsx_pca_project: n_comp must be in [1, n_bands]");
    RSX_REQUIRE(((uintptr_t)d_raster & 15) == 0 && ((uintptr_t)d_out & 15) == 0 && (plane_stride & 3) == 0 && plane_stride >= n_px,
                "rsx_pca_project: raster/planes must be 16-byte aligned, plane_stride a multiple of 4 and >= n_px");
#define LAUNCH(BB)                                                                                                           \
    {                                                                                                                        \
        using RT = RasterTiles<T, BB, 3>;                                                                                    \
        ProjParams<BB> P;                                                                                                    \
        fill_pca_params<BB>(P.pca, h_norm, h_center, h_scale, d_lut);                                                        \
        for (int c = 0; c < BB; ++c) {                                                                                       \
            for (int b = 0; b < BB; ++b) P.comp[c][b] = c < n_comp ? h_components[c * BB + b] : 0.f;                         \
            P.mean_proj[c] = c < n_comp ? h_mean_proj[c] : 0.f;                                                              \
        }                                                                                                                    \
        P.n_comp = n_comp;                                                                                                   \
        int smem = RT::SMEM_BYTES + lut_bytes<T>(BB);                                                                        \
        if (int rc = set_smem(pca_project_kernel<T, BB>, smem)) return rc;                                                   \
        int grid = persistent_grid(ceil_div(n_px, (int64_t)RT::TILE_PX), smem > 110000 ? 1 : 2);                             \
        pca_project_kernel<T, BB><<<grid, 256, smem, (cudaStream_t)stream>>>(d_raster, n_px, P, d_out, plane_stride, d_minmax); \
    }
    RSX_DISPATCH_BANDS(n_bands, LAUNCH)
#undef LAUNCH
    return rsx_check_launch("rsx_pca_project");
}
extern "C" int rsx_pca_project_u8(const uint8_t* d_raster, int64_t n_px, int n_bands, const float* d_lut, const float* h_components,
                                  const float* h_mean_proj, int n_comp, float* d_out, int64_t plane_stride, uint32_t* d_minmax, rsx_stream_t stream) {
    return pca_project_impl<uint8_t>(d_raster, n_px, n_bands, nullptr, nullptr, nullptr, d_lut, h_components, h_mean_proj, n_comp, d_out, plane_stride,
                                     d_minmax, stream);
}
extern "C" int rsx_pca_project_u16(const uint16_t* d_raster, int64_t n_px, int n_bands, const float* h_norm, const float* h_center,
                                   const double* h_scale, const float* d_lut16, const float* h_components, const float* h_mean_proj, int n_comp,
                                   float* d_out, int64_t plane_stride, uint32_t* d_minmax, rsx_stream_t stream) {
    return pca_project_impl<uint16_t>(d_raster, n_px, n_bands, h_norm, h_center, h_scale, d_lut16, h_components, h_mean_proj, n_comp, d_out, plane_stride,
                                      d_minmax, stream);
}
