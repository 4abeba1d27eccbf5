// Streaming access to a pixel-interleaved (BIP) integer raster.
//
// HBM layout: sample (pixel p, band b) at raster[p*B + b], uint8 or uint16.  A pixel is 7 bytes
// (Landsat TM) or 26 bytes (Sentinel-2-like), so no power-of-two vector load lines up with pixel
// boundaries.  Instead a persistent CTA streams contiguous TILE_PX-pixel tiles into shared memory
// with the bulk-copy engine (cp.async.bulk, one elected thread, mbarrier completion, NSTAGE deep),
// and every thread then reads the PXT pixels it owns as B consecutive 32-bit words - a lane stride
// of B words, conflict-free when B is odd - and unpacks the samples with byte permutes.
#pragma once
#include "rsx_common.cuh"

template <typename T>
struct PxPerThread {
    static constexpr int value = 4 / sizeof(T);  // 4 px (u8) or 2 px (u16): PXT*B*sizeof(T) = 4*B bytes = B words
};

template <typename T, int B, int NSTAGE_ = 3, int SUB_ = 4>
struct RasterTiles {
    static constexpr int PXT = PxPerThread<T>::value;
    static constexpr int THREADS = 256;
    static constexpr int SUB = SUB_;                         // pixel groups per thread per tile
    static constexpr int TILE_PX = THREADS * PXT * SUB;      // 4096 (u8) / 2048 (u16)
    static constexpr int TILE_BYTES = TILE_PX * B * (int)sizeof(T);
    static constexpr int NSTAGE = NSTAGE_;
    static constexpr int SMEM_BYTES = NSTAGE * TILE_BYTES + NSTAGE * 8;
    static_assert(TILE_BYTES % 16 == 0, "tile must be a multiple of 16 bytes");
};

// Unpack the PXT pixels whose B words start at w[] into v[p][b].
template <typename T, int B>
__device__ __forceinline__ void unpack_pixels(const uint32_t (&w)[B], int (&v)[PxPerThread<T>::value][B]) {
    constexpr int PXT = PxPerThread<T>::value;
#pragma unroll
    for (int p = 0; p < PXT; ++p)
#pragma unroll
        for (int b = 0; b < B; ++b) {
            if (sizeof(T) == 1) {
                int byte = p * B + b;
                v[p][b] = (int)((w[byte >> 2] >> (8 * (byte & 3))) & 0xffu);
            } else {
                int half = p * B + b;
                v[p][b] = (int)((w[half >> 1] >> (16 * (half & 1))) & 0xffffu);
            }
        }
}

// Persistent tile loop.  body(tile_smem_words, first_pixel_of_tile, pixels_in_tile) is called by all
// threads of the CTA for every tile the CTA owns; a partial last tile is staged with plain loads.
template <typename T, int B, int NSTAGE, int SUB = 4, typename Body>
__device__ __forceinline__ void for_each_tile(const T* __restrict__ raster, int64_t n_px, unsigned char* smem_raw, Body body) {
    using RT = RasterTiles<T, B, NSTAGE, SUB>;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + RT::NSTAGE * RT::TILE_BYTES);
    const int64_t n_tiles = (n_px + RT::TILE_PX - 1) / RT::TILE_PX;
    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < RT::NSTAGE; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int64_t tile, int stage) {
        // full tiles only: the engine needs 16-byte multiples
        int64_t px0 = tile * RT::TILE_PX;
        if (px0 + RT::TILE_PX <= n_px) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of this stage precede the async refill
            mbar_expect_tx(&bars[stage], RT::TILE_BYTES);
            bulk_g2s(smem_raw + stage * RT::TILE_BYTES, raster + px0 * B, RT::TILE_BYTES, &bars[stage]);
        }
    };

    int64_t first = blockIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < RT::NSTAGE; ++s) {
            int64_t t = first + (int64_t)s * gridDim.x;
            if (t < n_tiles) issue(t, s);
        }
    }
    int stage = 0;
    unsigned parity = 0;
    for (int64_t tile = first; tile < n_tiles; tile += gridDim.x) {
        int64_t px0 = tile * RT::TILE_PX;
        int npx = (int)min((int64_t)RT::TILE_PX, n_px - px0);
        unsigned char* buf = smem_raw + stage * RT::TILE_BYTES;
        if (npx == RT::TILE_PX) {
            mbar_wait(&bars[stage], parity);
        } else {
            // ragged tail: byte-granular cooperative copy, zero fill
            const unsigned char* src = reinterpret_cast<const unsigned char*>(raster + px0 * B);
            int nbytes = npx * B * (int)sizeof(T);
            for (int i = tid; i < RT::TILE_BYTES; i += RT::THREADS) buf[i] = i < nbytes ? src[i] : 0;
            __syncthreads();
        }
        body(reinterpret_cast<const uint32_t*>(buf), px0, npx);
        __syncthreads();  // everyone is done with this stage before it is refilled
        if (tid == 0) {
            int64_t nxt = tile + (int64_t)RT::NSTAGE * gridDim.x;
            if (nxt < n_tiles) issue(nxt, stage);
        }
        if (++stage == RT::NSTAGE) {
            stage = 0;
            parity ^= 1;
        }
    }
}

// Dispatch over the (type, band-count) instantiations that are compiled in.
#define RSX_DISPATCH_BANDS(B_RUNTIME, MACRO)                                  \
    switch (B_RUNTIME) {                                                      \
        case 5: MACRO(5); break;                                              \
        case 6: MACRO(6); break;                                              \
        case 7: MACRO(7); break;                                              \
        case 8: MACRO(8); break;                                              \
        case 10: MACRO(10); break;                                            \
        case 12: MACRO(12); break;                                            \
        case 13: MACRO(13); break;                                            \
        default:                                                              \
            rsx_set_error("unsupported band count %d (compiled: 5,6,7,8,10,12,13)", B_RUNTIME); \
            return RSX_ERR_UNSUPPORTED;                                       \
    }
