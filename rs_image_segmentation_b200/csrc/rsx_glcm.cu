// K4: GLCM texture properties, co-occurrence count dump, and the cv2-compatible bilinear upsample.
//
// What the reference does per window (indices.py:285-296 + skimage graycomatrix/graycoprops):
// directed counts C_t[a][b] for the four offsets t = (0,1),(1,1),(1,0),(1,-1); P = C + C^T; normalise;
// five properties; mean over the four angles.  Writing n for the number of pixel pairs of one angle
// and (a_s, b_s) for the grey levels of pair s, the symmetric normalised matrix gives exactly
//     contrast      = sum_s (a-b)^2 / n            dissimilarity = sum_s |a-b| / n
//     homogeneity   = sum_s 1/(1+(a-b)^2) / n
//     mu            = SA / 2n,  SA = sum_s (a+b);   var = (2n SQ - SA^2)/(2n)^2,  SQ = sum_s (a^2+b^2)
//     correlation   = (4n SAB - SA^2) / (2n SQ - SA^2),  SAB = sum_s a b   (1 when var == 0)
//     energy        = sqrt(sum_cells P^2) = sqrt(E)/2n, E = sum_s (a!=b ? 2 U[{a,b}] : 4 U[{a,a}])
// where U[{a,b}] counts the pairs of the window whose UNORDERED levels are {a,b}.  All sums are exact
// integers; only E needs the histogram, and it needs it only at the cells the window touches.
#include "rsx_common.cuh"

__constant__ double g_homog[256];  // 1/(1+k^2)
static bool g_homog_ready = false;

static int ensure_homog() {
    if (g_homog_ready) return RSX_OK;
    double h[256];
    for (int k = 0; k < 256; ++k) h[k] = 1.0 / (1.0 + (double)k * (double)k);
    cudaError_t e = cudaMemcpyToSymbol(g_homog, h, sizeof(h));
    if (e != cudaSuccess) {
        rsx_set_error("glcm: constant upload failed: %s", cudaGetErrorString(e));
        return RSX_ERR_CUDA;
    }
    g_homog_ready = true;
    return RSX_OK;
}

__device__ __forceinline__ int tri_cell(int a, int b) {
    int lo = min(a, b), hi = max(a, b);
    return (hi * (hi + 1) >> 1) + lo;
}

struct AngleSums {
    int s1, sa, sq, sab, e;
    double sh;
};

// finalise one angle (lane 0): returns the five properties
__device__ __forceinline__ void angle_props(const AngleSums& t, int n, double (&acc)[5]) {
    const double dn = (double)n;
    const long long s2 = (long long)t.sq - 2ll * t.sab;
    acc[0] += (double)s2 / dn;
    acc[1] += (double)t.s1 / dn;
    acc[2] += t.sh / dn;
    acc[3] += sqrt((double)t.e) / (2.0 * dn);
    const long long var_num = 2ll * n * (long long)t.sq - (long long)t.sa * t.sa;
    const long long cov_num = 4ll * n * (long long)t.sab - (long long)t.sa * t.sa;
    // graycoprops: std < 1e-15 -> 1.  var = var_num/(2n)^2 is an exact rational: zero iff the window is constant.
    acc[4] += var_num <= 0 ? 1.0 : (double)cov_num / (double)var_num;
}

// ----------------------------------------------------------------------------- general kernel: one warp per window
// Any window/step/levels.  The histogram of unordered cells lives in the warp's slice of shared memory and is
// cleaned by revisiting only the touched cells, so the cost per window is O(pairs), not O(levels^2).
__global__ void __launch_bounds__(128) glcm_props_warp_kernel(const uint8_t* __restrict__ q, int W, int L, int win, int step, int out_rows,
                                                              int out_cols, float* __restrict__ props, int64_t plane_stride) {
    extern __shared__ unsigned glcm_sm[];
    const int ncell = L * (L + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* U = glcm_sm + warp * ncell;
    for (int i = lane; i < ncell; i += 32) U[i] = 0;
    __syncwarp();
    const int64_t n_win = (int64_t)out_rows * out_cols;
    const int64_t warps_total = (int64_t)gridDim.x * 4;
    for (int64_t wdx = (int64_t)blockIdx.x * 4 + warp; wdx < n_win; wdx += warps_total) {
        const int oi = (int)(wdx / out_cols), oj = (int)(wdx % out_cols);
        const uint8_t* base = q + (int64_t)oi * step * W + (int64_t)oj * step;
        double acc[5] = {0, 0, 0, 0, 0};
#pragma unroll 1
        for (int ang = 0; ang < 4; ++ang) {
            const int dr = ang == 0 ? 0 : 1;
            const int dc = ang == 0 ? 1 : (ang == 1 ? 1 : (ang == 2 ? 0 : -1));
            const int nrows = win - dr, ncols = win - (dc != 0 ? 1 : 0), c0 = dc < 0 ? 1 : 0;
            const int n = nrows * ncols;
            const int off = dr * W + dc;
            AngleSums t = {0, 0, 0, 0, 0, 0.0};
            // pass A: count + integer moments
            for (int s = lane; s < n; s += 32) {
                const int r = s / ncols, c = s - r * ncols + c0;
                const uint8_t* p = base + r * W + c;
                const int a = p[0], b = p[off];
                const int d = abs(a - b);
                t.s1 += d;
                t.sa += a + b;
                t.sq += a * a + b * b;
                t.sab += a * b;
                t.sh += g_homog[d];
                if (a < L && b < L) atomicAdd(&U[tri_cell(a, b)], 1u);
            }
            __syncwarp();
            // pass B: E = sum over pairs of the multiplicity of their cell
            for (int s = lane; s < n; s += 32) {
                const int r = s / ncols, c = s - r * ncols + c0;
                const uint8_t* p = base + r * W + c;
                const int a = p[0], b = p[off];
                if (a < L && b < L) t.e += (a != b ? 2 : 4) * (int)U[tri_cell(a, b)];
            }
            __syncwarp();
            // pass C: clean
            for (int s = lane; s < n; s += 32) {
                const int r = s / ncols, c = s - r * ncols + c0;
                const uint8_t* p = base + r * W + c;
                const int a = p[0], b = p[off];
                if (a < L && b < L) U[tri_cell(a, b)] = 0;
            }
            __syncwarp();
            t.s1 = __reduce_add_sync(0xffffffffu, t.s1);
            t.sa = __reduce_add_sync(0xffffffffu, t.sa);
            t.sq = __reduce_add_sync(0xffffffffu, t.sq);
            t.sab = __reduce_add_sync(0xffffffffu, t.sab);
            t.e = __reduce_add_sync(0xffffffffu, t.e);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t.sh += __shfl_xor_sync(0xffffffffu, t.sh, o);
            if (lane == 0) angle_props(t, n, acc);
        }
        if (lane == 0) {
            const int64_t o = (int64_t)oi * out_cols + oj;
#pragma unroll
            for (int k = 0; k < 5; ++k) props[k * plane_stride + o] = (float)(acc[k] * 0.25);  // .mean() over the 1x4 array
        }
    }
}

// ----------------------------------------------------------------------------- dense kernel (step == 1): one thread per window column
// Lane = window column j, the thread slides its window DOWN the rows.  Moving down by one row removes
// the pairs anchored in the row that leaves and adds those of the row that enters (2*(4w-2) pair updates
// instead of 4w(w-1)+... recounts).  Integer moments are maintained incrementally and exactly; the
// energy term needs per-window cell multiplicities, kept as uint8 counters (<= w(w-1) <= 255 for w <= 16)
// in a thread-private slice of shared memory laid out [cell][lane] so that the 32 lanes of a warp
// hit 32 consecutive bytes (8 banks) per cell index.
template <int NTHREADS>
struct DenseCfg {
    static constexpr int TILE_COLS = NTHREADS;  // windows per CTA row
};

template <int NTHREADS>
__global__ void __launch_bounds__(NTHREADS) glcm_props_dense_kernel(const uint8_t* __restrict__ q, int W, int L, int win, int out_rows, int out_cols,
                                                                    int rows_per_cta, float* __restrict__ props, int64_t plane_stride) {
    extern __shared__ unsigned char dsm[];
    // layout: counters [4 angles][ncell][NTHREADS] u8, then the q tile rows ring [(win+1)][tile_w]
    const int ncell = L * (L + 1) / 2;
    unsigned char* cnt = dsm;
    const int tile_w = NTHREADS + win - 1;
    unsigned char* qt = dsm + (size_t)4 * ncell * NTHREADS;  // ring of win+1 rows
    const int ring = win + 1;

    const int j0 = blockIdx.x * NTHREADS;           // first window column of this CTA
    const int i_begin = blockIdx.y * rows_per_cta;  // first window row
    const int i_end = min(out_rows, i_begin + rows_per_cta);
    if (i_begin >= i_end) return;
    const int t = threadIdx.x;
    const int j = j0 + t;
    const bool active = j < out_cols;

    for (int i = t; i < 4 * ncell * NTHREADS / 4; i += NTHREADS) reinterpret_cast<unsigned*>(cnt)[i] = 0;

    auto load_row = [&](int img_row) {
        unsigned char* dst = qt + (img_row % ring) * tile_w;
        const uint8_t* src = q + (int64_t)img_row * W + j0;
        const int valid = min(tile_w, W - j0);
        for (int c = t; c < tile_w; c += NTHREADS) dst[c] = c < valid ? (unsigned char)min((int)src[c], L - 1) : 0;
    };
    auto Q = [&](int img_row, int c) -> int { return qt[(img_row % ring) * tile_w + c]; };

    // per-angle running sums for this thread's window
    int s1[4] = {0, 0, 0, 0}, sa[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0}, sab[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};
    double sh[4] = {0, 0, 0, 0};

    // add (sign=+1) or remove (sign=-1) the pairs of all four angles anchored in image row r (window columns t..t+win-1).
    // angle 0: (r,c)-(r,c+1), c in [0,win-1)        needs row r
    // angle 1: (r,c)-(r+1,c+1), c in [0,win-1)      needs rows r, r+1
    // angle 2: (r,c)-(r+1,c), c in [0,win)          needs rows r, r+1
    // angle 3: (r,c)-(r+1,c-1), c in [1,win)        needs rows r, r+1
    auto pair_update = [&](int ang, int a, int b, int sign) {
        const int d = abs(a - b);
        s1[ang] += sign * d;
        sa[ang] += sign * (a + b);
        sq[ang] += sign * (a * a + b * b);
        sab[ang] += sign * (a * b);
        sh[ang] += sign > 0 ? g_homog[d] : -g_homog[d];
        unsigned char* cell = cnt + ((size_t)ang * ncell + tri_cell(a, b)) * NTHREADS + t;
        const int wgt = a != b ? 2 : 4;
        int u = *cell;
        if (sign > 0) {
            e[ang] += wgt * (2 * u + 1);  // (u+1)^2 - u^2
            *cell = (unsigned char)(u + 1);
        } else {
            e[ang] -= wgt * (2 * u - 1);  // u^2 - (u-1)^2
            *cell = (unsigned char)(u - 1);
        }
    };
    auto row_update = [&](int r, bool with_next, int sign) {
        for (int c = 0; c < win; ++c) {
            const int a = Q(r, t + c);
            if (c + 1 < win) pair_update(0, a, Q(r, t + c + 1), sign);
            if (with_next) {
                if (c + 1 < win) pair_update(1, a, Q(r + 1, t + c + 1), sign);
                pair_update(2, a, Q(r + 1, t + c), sign);
                if (c >= 1) pair_update(3, a, Q(r + 1, t + c - 1), sign);
            }
        }
    };

    // prologue: rows of the first window
    for (int r = 0; r < win; ++r) load_row(i_begin + r);
    __syncthreads();
    if (active)
        for (int r = 0; r < win; ++r) row_update(i_begin + r, r + 1 < win, +1);

    for (int i = i_begin; i < i_end; ++i) {
        if (active) {
            double acc[5] = {0, 0, 0, 0, 0};
#pragma unroll
            for (int ang = 0; ang < 4; ++ang) {
                const int n = (ang == 0 || ang == 2) ? win * (win - 1) : (win - 1) * (win - 1);
                AngleSums ts = {s1[ang], sa[ang], sq[ang], sab[ang], e[ang], sh[ang]};
                angle_props(ts, n, acc);
            }
            const int64_t o = (int64_t)i * out_cols + j;
#pragma unroll
            for (int k = 0; k < 5; ++k) props[k * plane_stride + o] = (float)(acc[k] * 0.25);
        }
        if (i + 1 >= i_end) break;
        // slide down: image row i leaves, image row i+win enters
        __syncthreads();  // everyone finished reading the ring slot that row i+win overwrites (slot of row i-1... see ring = win+1)
        load_row(i + win);
        __syncthreads();
        if (active) {
            // leaving: pairs anchored in row i (angle 0 within row i; angles 1-3 between rows i and i+1)
            row_update(i, true, -1);
            // entering: angle 0 within row i+win; angles 1-3 between rows i+win-1 and i+win
            for (int c = 0; c < win; ++c) {
                const int a = Q(i + win, t + c);
                if (c + 1 < win) pair_update(0, a, Q(i + win, t + c + 1), +1);
            }
            for (int c = 0; c < win; ++c) {
                const int a = Q(i + win - 1, t + c);
                if (c + 1 < win) pair_update(1, a, Q(i + win, t + c + 1), +1);
                pair_update(2, a, Q(i + win, t + c), +1);
                if (c >= 1) pair_update(3, a, Q(i + win, t + c - 1), +1);
            }
        }
    }
}

extern "C" int rsx_glcm_props(const uint8_t* d_q, int rows_avail, int W, int levels, int window, int step, int out_rows, int out_cols,
                              float* d_props, int64_t plane_stride, rsx_stream_t stream) {
    RSX_REQUIRE(d_q && d_props, "rsx_glcm_props: null argument");
    RSX_REQUIRE(levels >= 2 && levels <= 128, "rsx_glcm_props: levels must be in [2,128]");
    RSX_REQUIRE(window >= 2 && window <= 127 && step >= 1, "rsx_glcm_props: window must be in [2,127], step >= 1");
    RSX_REQUIRE(out_rows >= 1 && out_cols >= 1 && (int64_t)(out_rows - 1) * step + window <= rows_avail && (int64_t)(out_cols - 1) * step + window <= W,
                "rsx_glcm_props: windows exceed the %dx%d band", rows_avail, W);
    RSX_REQUIRE(plane_stride >= (int64_t)out_rows * out_cols, "rsx_glcm_props: plane_stride too small");
    if (int rc = ensure_homog()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int ncell = levels * (levels + 1) / 2;
    // dense path: thread-private uint8 counters must hold w(w-1) and fit in shared memory
    constexpr int NT = 64;
    const size_t dense_smem = (size_t)4 * ncell * NT + (size_t)(window + 1) * (NT + window - 1);
    if (step == 1 && window * (window - 1) <= 255 && dense_smem <= 200 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(glcm_props_dense_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) {
                rsx_set_error("rsx_glcm_props: %s", cudaGetErrorString(e));
                return RSX_ERR_CUDA;
            }
            attr_set = true;
        }
        const int gx = ceil_div(out_cols, NT);
        // enough row strips to fill the machine ~4x, but long enough to amortise the (window-1)-row prologue
        int strips = max(1, (rsx_num_sms() * 4) / gx);
        int rows_per_cta = max(8 * window, ceil_div(out_rows, strips));
        rows_per_cta = min(rows_per_cta, out_rows);
        const int gy = ceil_div(out_rows, rows_per_cta);
        glcm_props_dense_kernel<NT><<<dim3(gx, gy), NT, dense_smem, s>>>(d_q, W, levels, window, out_rows, out_cols, rows_per_cta, d_props, plane_stride);
        return rsx_check_launch("glcm_props_dense");
    }
    const size_t smem = (size_t)4 * ncell * 4;
    static bool attr_set2 = false;
    if (!attr_set2) {
        cudaError_t e = cudaFuncSetAttribute(glcm_props_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
        if (e != cudaSuccess) {
            rsx_set_error("rsx_glcm_props: %s", cudaGetErrorString(e));
            return RSX_ERR_CUDA;
        }
        attr_set2 = true;
    }
    const int64_t n_win = (int64_t)out_rows * out_cols;
    const int grid = (int)min(ceil_div(n_win, (int64_t)4), (int64_t)rsx_num_sms() * 8);
    glcm_props_warp_kernel<<<grid, 128, smem, s>>>(d_q, W, levels, window, step, out_rows, out_cols, d_props, plane_stride);
    return rsx_check_launch("glcm_props_warp");
}

// ----------------------------------------------------------------------------- count dump (validation entry point)
__global__ void __launch_bounds__(256) glcm_counts_kernel(const uint8_t* __restrict__ q, int W, int L, int win, const int32_t* __restrict__ anchors,
                                                          uint32_t* __restrict__ counts) {
    const int wdx = blockIdx.x;
    uint32_t* out = counts + (size_t)wdx * 4 * L * L;
    for (int i = threadIdx.x; i < 4 * L * L; i += blockDim.x) out[i] = 0;
    __syncthreads();
    const uint8_t* base = q + (int64_t)anchors[2 * wdx] * W + anchors[2 * wdx + 1];
    for (int ang = 0; ang < 4; ++ang) {
        const int dr = ang == 0 ? 0 : 1;
        const int dc = ang == 0 ? 1 : (ang == 1 ? 1 : (ang == 2 ? 0 : -1));
        const int nrows = win - dr, ncols = win - (dc != 0 ? 1 : 0), c0 = dc < 0 ? 1 : 0;
        for (int s = threadIdx.x; s < nrows * ncols; s += blockDim.x) {
            const int r = s / ncols, c = s - r * ncols + c0;
            const int a = base[r * W + c], b = base[(r + dr) * W + c + dc];
            if (a < L && b < L) atomicAdd(&out[((size_t)ang * L + a) * L + b], 1u);
        }
    }
}

extern "C" int rsx_glcm_counts(const uint8_t* d_q, int H, int W, int levels, int window, const int32_t* d_anchors, int n_win, uint32_t* d_counts,
                               rsx_stream_t stream) {
    RSX_REQUIRE(d_q && d_anchors && d_counts && n_win >= 1, "rsx_glcm_counts: bad arguments");
    RSX_REQUIRE(levels >= 2 && levels <= 256 && window >= 2 && window <= H && window <= W, "rsx_glcm_counts: bad levels/window");
    glcm_counts_kernel<<<n_win, 256, 0, (cudaStream_t)stream>>>(d_q, W, levels, window, d_anchors, d_counts);
    return rsx_check_launch("glcm_counts");
}

// ----------------------------------------------------------------------------- cv2.resize(INTER_LINEAR) for float32 planes
// The oracle is cv2.resize as shipped in opencv-python (built with Intel IPP, which handles float32 linear resize).
// Its arithmetic, established by probing it (DESIGN.md "cv2.resize parity"): source coordinate c = (d+0.5)*scale-0.5
// in DOUBLE with scale = 1/(dst/src); s = floor(c); weight f = (float)(c - s); the two taps are clamped to the image
// (replicated border); horizontal pass first, then vertical, each as fma(q - p, f, p).
// (OpenCV's own non-IPP path rounds c to float before taking the fraction and differs from this by ~6e-5 relative.)
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ src, int src_h, int src_w, int src_row0, int src_rows_avail,
                                                              int64_t src_stride, float* __restrict__ dst, int dst_w, int dst_row0, int dst_rows,
                                                              int64_t dst_stride, double scale_x, double scale_y, uint32_t* __restrict__ minmax) {
    const int plane = blockIdx.z;
    const float* sp = src + plane * src_stride;
    float* dp = dst + plane * dst_stride;
    float mn = INFINITY, mx = -INFINITY;
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    if (dx < dst_w) {
        const double cx = (dx + 0.5) * scale_x - 0.5;
        const double fxd = floor(cx);
        const float fx = (float)(cx - fxd);
        const int sx = (int)fxd;
        const int x0 = min(max(sx, 0), src_w - 1), x1 = min(max(sx + 1, 0), src_w - 1);
        for (int ly = blockIdx.y; ly < dst_rows; ly += gridDim.y) {
            const int dy = dst_row0 + ly;
            const double cy = (dy + 0.5) * scale_y - 0.5;
            const double fyd = floor(cy);
            const float fy = (float)(cy - fyd);
            const int sy = (int)fyd;
            const int y0 = min(max(sy, 0), src_h - 1) - src_row0, y1 = min(max(sy + 1, 0), src_h - 1) - src_row0;
            float r0 = 0.f, r1 = 0.f;
            if (y0 >= 0 && y0 < src_rows_avail) {
                const float p = sp[(int64_t)y0 * src_w + x0], q = sp[(int64_t)y0 * src_w + x1];
                r0 = fmaf(f_sub(q, p), fx, p);
            }
            if (y1 >= 0 && y1 < src_rows_avail) {
                const float p = sp[(int64_t)y1 * src_w + x0], q = sp[(int64_t)y1 * src_w + x1];
                r1 = fmaf(f_sub(q, p), fx, p);
            }
            const float v = fmaf(f_sub(r1, r0), fy, r0);
            dp[(int64_t)ly * dst_w + dx] = v;
            mn = fminf(mn, v), mx = fmaxf(mx, v);
        }
    }
    if (minmax) warp_minmax_commit(mn, mx, minmax + 2 * plane);
}

extern "C" int rsx_resize_bilinear_f32(const float* d_src, int src_h_total, int src_w, int src_row0, int src_rows_avail, int64_t src_plane_stride,
                                       float* d_dst, int dst_h_total, int dst_w, int dst_row0, int dst_rows, int64_t dst_plane_stride, int n_planes,
                                       uint32_t* d_minmax, rsx_stream_t stream) {
    RSX_REQUIRE(d_src && d_dst && src_h_total >= 1 && src_w >= 1 && dst_h_total >= 1 && dst_w >= 1 && dst_rows >= 1 && n_planes >= 1,
                "rsx_resize_bilinear_f32: bad arguments");
    const double scale_x = 1.0 / ((double)dst_w / (double)src_w), scale_y = 1.0 / ((double)dst_h_total / (double)src_h_total);
    dim3 grid(ceil_div(dst_w, 256), min(dst_rows, max(1, rsx_num_sms() * 8 / ceil_div(dst_w, 256))), n_planes);
    resize_bilinear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, src_h_total, src_w, src_row0, src_rows_avail, src_plane_stride, d_dst, dst_w,
                                                                   dst_row0, dst_rows, dst_plane_stride, scale_x, scale_y, d_minmax);
    return rsx_check_launch("resize_bilinear");
}
