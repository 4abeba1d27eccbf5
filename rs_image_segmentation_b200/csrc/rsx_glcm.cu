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
#include <climits>
#include <cstdlib>

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
struct GlcmOffsets {  // (row, column) offsets of the co-occurrence pairs: round(sin(angle) * d), round(cos(angle) * d) per (distance, angle)
    int n;
    int dr[16], dc[16];
};
static GlcmOffsets default_offsets() {
    GlcmOffsets o;
    o.n = 4;
    const int dr[4] = {0, 1, 1, 1}, dc[4] = {1, 1, 0, -1};
    for (int i = 0; i < 16; ++i) o.dr[i] = i < 4 ? dr[i] : 0, o.dc[i] = i < 4 ? dc[i] : 0;
    return o;
}

__global__ void __launch_bounds__(128) glcm_props_warp_kernel(const uint8_t* __restrict__ q, int W, int L, int win, int step, int out_rows,
                                                              int out_cols, float* __restrict__ props, int64_t plane_stride,
                                                              long long* __restrict__ moments, const __grid_constant__ GlcmOffsets offs) {
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
        for (int ang = 0; ang < offs.n; ++ang) {
            const int dr = offs.dr[ang], dc = offs.dc[ang];
            // anchors (r, c) whose partner (r + dr, c + dc) lies inside the window (graycomatrix's bounds test)
            const int r0 = max(0, -dr), c0 = max(0, -dc);
            const int nrows = max(0, min(win, win - dr) - r0), ncols = max(0, min(win, win - dc) - c0);
            const int n = nrows * ncols;
            if (n == 0) {  // offset longer than the window: an all-zero matrix - every property 0, correlation 1 (std < 1e-15)
                acc[4] += 1.0;
                continue;
            }
            const int off = dr * W + dc;
            AngleSums t = {0, 0, 0, 0, 0, 0.0};
            long long hfx = 0;  // moments dump only: the fixed-point homogeneity terms of the dense kernel
            int neq = 0;
            // pass A: count + integer moments
            for (int s = lane; s < n; s += 32) {
                const int r = s / ncols + r0, c = s % ncols + c0;
                const uint8_t* p = base + r * W + c;
                const int a = p[0], b = p[off];
                const int d = abs(a - b);
                if (moments) hfx += (long long)((L <= 32 ? 1099511627776.0 : 68719476736.0) / (1.0 + (double)d * (double)d) + 0.5), neq += d == 0;
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
                const int r = s / ncols + r0, c = s % ncols + c0;
                const uint8_t* p = base + r * W + c;
                const int a = p[0], b = p[off];
                if (a < L && b < L) t.e += (a != b ? 2 : 4) * (int)U[tri_cell(a, b)];
            }
            __syncwarp();
            // pass C: clean
            for (int s = lane; s < n; s += 32) {
                const int r = s / ncols + r0, c = s % ncols + c0;
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
            if (moments) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) hfx += __shfl_xor_sync(0xffffffffu, hfx, o);
                neq = __reduce_add_sync(0xffffffffu, neq);
                if (lane == 0) {
                    long long* m = moments + (((int64_t)oi * out_cols + oj) * 4 + ang) * 8;
                    m[0] = n, m[1] = t.s1, m[2] = t.sa, m[3] = t.sq, m[4] = t.sab, m[5] = t.e, m[6] = neq, m[7] = hfx;
                }
            }
            if (lane == 0) angle_props(t, n, acc);
        }
        if (lane == 0) {
            const int64_t o = (int64_t)oi * out_cols + oj;
#pragma unroll
            for (int k = 0; k < 5; ++k) props[k * plane_stride + o] = (float)(offs.n == 4 ? acc[k] * 0.25 : acc[k] / (double)offs.n);  // .mean() over the (distances x angles) array
        }
    }
}

// ----------------------------------------------------------------------------- dense kernel (step == 1)
// One thread per window column; the CTA slides a band of NTW = NT-(WIN-1) windows DOWN the rows.
//
// Per image row that enters (and per row that leaves) every thread handles ONE image column x = t:
//   * the four pair codes anchored there (unordered cell index | "a==b" flag) go to a shared ring, so that the
//     WIN windows containing a pair never recompute its cell;
//   * the packed integer moments of those pairs update the thread's COLUMN sums (registers):
//        W1 = |a-b| + (a*b << 13)       W2 = (a+b) + ((a^2+b^2) << 14)       SH = round(2^40/(1+(a-b)^2)) (u64)
//     fields never overflow for levels <= 32, WIN <= 12, so one integer add updates two sums.
// Column sums are exchanged through shared memory and each window thread adds the WIN (or WIN-1) columns it
// spans.  Only the energy term needs co-occurrence multiplicities: per window, uint8 counters indexed by
// (angle, unordered cell), thread-private in shared memory laid out [cell][thread]; a row slide does WIN-ish
// decrements and increments per angle and maintains E = sum_cells weight*count^2 incrementally.
// column range [c0, c1) of the anchors of each angle inside a window
__host__ __device__ constexpr int dense_c0(int ang) { return ang == 3 ? 1 : 0; }
__host__ __device__ constexpr int dense_c1(int ang, int win) { return ang <= 1 ? win - 1 : win; }

// Thread (t, ang): warp w serves angle w & 3 (so each SM sub-partition runs one angle's specialised code) and image
// columns (w >> 2) * 32 + lane of the CTA's band.  Each thread owns the column sums of ITS angle at ITS column and - if
// t < NTW - the angle's energy counters and window sums of window t; the four angle threads of a window meet in shared
// memory.
//
// Energy without dependent read-modify-write chains: the WIN-ish pairs of one image row that enter (leave) a window may
// hit the same cell, so a naive loop is a chain of LDS -> add -> STS.  Instead every pair code carries an "equal to the
// pair d columns to the left" mask (computed once by the thread that produces the code, shared by all the windows that
// contain it); a window then loads all its counters at once and corrects each by the number of equal pairs that precede
// it inside the window (popc of the masked bits) - independent loads, ordered stores.
// With weight w = 2 (a != b) or 4 (a == b):  E = sum_cells w U^2 = e + SW,  where e accumulates 2w*U_before on every
// increment (and -2w*U_after on every decrement) and SW = sum over the window's pairs of w = 2n + 2*Neq (Neq = pairs with
// a == b, one more packed column sum).
// Pair code word: bits 0..3 = 2w, bits 4..13 = equality mask (bit 3+d: equal to the pair d columns left), bits 16..31 =
// byte offset of the counter (cell * NTW).
struct DenseShared {
    uint4* xch;            // [4 ang][NT]: w1, w2, sh.lo, sh.hi column sums
    float* outx;           // [4 ang][5][NT] partial properties
    unsigned* codes;       // [4][RING][NT] pair code words
    unsigned* base;        // [4][WIN + NT] cell ids of the row being entered (WIN sentinels on the left)
    unsigned char* qring;  // [RING][NT] u8
    unsigned char* cnt;    // [4][ncell][NTW] u8
    const unsigned long long* homog_fx;
};

// WIDE (levels > 32): sum of a^2+b^2 needs its own 32-bit word, so the sum of a+b moves into the u64:
//   narrow: W1 = s1 | sab << 13,  W2 = sa | sq << 14,  SH = homog * 2^40 | Neq << 52
//   wide:   W1 = s1 | sab << 13,  W2 = sq,             SH = homog * 2^36 | Neq << 43 | sa << 50
// FOLD = 8 or 16 (levels <= 32): the private counters are indexed by the levels modulo FOLD - F(F+1)/2 cells for the pairs with
// a != b plus F cells for the pairs with a == b (44 B per window and angle at FOLD 8 instead of 528) - which is exact for every
// window whose levels span fewer than FOLD values (the fold is injective there).  Windows with a wider span are flagged by
// glcm_span_flag_kernel and get their energy from glcm_energy_patch_kernel; the other four properties never use the counters.
// The a == b pairs have their own cells because the incremental energy e is a function of the counters only as long as all
// pairs of a cell carry the same weight (2 for a != b, 4 for a == b): with mixed cells a window that WAS flagged would leave a
// wrong e behind for the unflagged windows below it.
__host__ __device__ constexpr int fold_cells(int F) { return F * (F + 1) / 2 + F; }
template <int WIN, int NT, int ANG, bool WIDE, int FOLD>
__device__ __forceinline__ void glcm_dense_body(const DenseShared sm, const uint8_t* __restrict__ q, int W, int L, int out_cols, int i_begin, int i_end,
                                                int j0, int t, int NTW, float* __restrict__ props, int64_t plane_stride,
                                                long long* __restrict__ moments) {
    constexpr int RING = WIN + 1;
    constexpr int DR = ANG == 0 ? 0 : 1;
    constexpr int DC = ANG == 0 ? 1 : (ANG == 1 ? 1 : (ANG == 2 ? 0 : -1));
    constexpr int C0 = ANG == 3 ? 1 : 0, C1 = ANG <= 1 ? WIN - 1 : WIN;  // anchor columns of the angle inside a window
    constexpr int NPAIR = (ANG == 0 || ANG == 2) ? WIN * (WIN - 1) : (WIN - 1) * (WIN - 1);
    const int ncell = FOLD ? fold_cells(FOLD) : L * (L + 1) / 2;
    const bool has_win = t < NTW && j0 + t < out_cols;
    const bool col_ok = j0 + t < W;
    const bool pair_ok = t + DC >= 0 && t + DC < NT;  // partner column inside the band
    const float inv_n = 1.f / (float)NPAIR;

    unsigned cw1 = 0, cw2 = 0;  // column sums of this (column, angle): packed moments
    unsigned long long csh = 0;
    int e = 0;
    unsigned* my_codes = sm.codes + ANG * RING * NT + t;
    unsigned* my_base = sm.base + ANG * (WIN + NT) + WIN + t;
    // Folded variants: the counters of a thread sit in ITS bank - cell c of thread t at word (c / 4) * NT + t, byte c % 4 (NT is a
    // multiple of 32) - so a warp's 32 counter accesses never conflict.  (With the unfolded [cell][window] byte layout the bank is
    // a function of the cell: ~3 wavefronts per access on imagery, 29 % of the kernel's shared-memory wavefronts.)
    unsigned char* my_cnt = FOLD ? sm.cnt + ((size_t)ANG * ((ncell + 3) / 4) * NT + t) * 4 : sm.cnt + (size_t)ANG * ncell * NTW + t;
    const unsigned char* qt = sm.qring + t;

    auto slot_add = [](int s, int k) { s += k; return s >= RING ? s - RING : s; };
    auto load_q = [&](int r) -> unsigned char { return col_ok ? (unsigned char)min((int)q[(int64_t)r * W + j0 + t], L - 1) : (unsigned char)0; };
    auto pair_terms = [&](int a, int b, unsigned& w1, unsigned& w2, unsigned long long& sh) {
        const int d = abs(a - b);
        w1 = (unsigned)d + ((unsigned)(a * b) << 13);
        if (WIDE) {
            w2 = (unsigned)(a * a + b * b);
            sh = sm.homog_fx[d] + ((unsigned long long)(a + b) << 50);
        } else {
            w2 = (unsigned)(a + b) + ((unsigned)(a * a + b * b) << 14);
            sh = sm.homog_fx[d];  // scaled 1/(1+d^2), plus the Neq unit when d == 0
        }
    };
    // pairs of this angle anchored in ring slot sa (partner row in slot sb) become complete: column sums + cell id
    auto enter_pairs = [&](int sa, int sb) {
        unsigned cell = 0xfffffff0u;  // never equal to a real cell
        if (pair_ok) {
            const int a = qt[sa * NT], b = qt[sb * NT + DC];
            const int cid = !FOLD ? tri_cell(a, b) : (a == b ? FOLD * (FOLD + 1) / 2 + (a & (FOLD - 1)) : tri_cell(a & (FOLD - 1), b & (FOLD - 1)));
            cell = ((unsigned)(FOLD ? (cid >> 2) * (NT * 4) + (cid & 3) : cid * NTW) << 16) | (a == b ? 8u : 4u);
            unsigned w1, w2;
            unsigned long long sh;
            pair_terms(a, b, w1, w2, sh);
            cw1 += w1, cw2 += w2, csh += sh;
        }
        *my_base = cell;
    };
    // after a barrier: attach the equality mask and publish the code word in ring slot sa
    auto publish_code = [&](int sa) {
        const unsigned cell = *my_base;
        unsigned mask = 0;
#pragma unroll
        for (int d = 1; d < WIN; ++d)
            if (my_base[-d] == cell) mask |= 1u << (3 + d);
        my_codes[sa * NT] = cell | mask;
    };
    auto leave_pairs = [&](int sa, int sb) {
        if (pair_ok) {
            unsigned w1, w2;
            unsigned long long sh;
            pair_terms(qt[sa * NT], qt[sb * NT + DC], w1, w2, sh);
            cw1 -= w1, cw2 -= w2, csh -= sh;
        }
    };
    auto energy_add = [&](int sa) {
        const unsigned* row = my_codes + sa * NT;
        unsigned code[WIN];
        int u[WIN];
#pragma unroll
        for (int c = C0; c < C1; ++c) code[c] = row[c];
#pragma unroll
        for (int c = C0; c < C1; ++c) u[c] = my_cnt[code[c] >> 16];
#pragma unroll
        for (int c = C0; c < C1; ++c) {
            constexpr unsigned all = 0x3ff0u;
            const unsigned kmask = (((1u << (c - C0)) - 1u) << 4) & all;  // pairs of this window to the left of c
            const int uu = u[c] + __popc(code[c] & kmask);
            e += uu * (int)(code[c] & 0xfu);
            my_cnt[code[c] >> 16] = (unsigned char)(uu + 1);
        }
    };
    auto energy_sub = [&](int sa) {
        const unsigned* row = my_codes + sa * NT;
        unsigned code[WIN];
        int u[WIN];
#pragma unroll
        for (int c = C0; c < C1; ++c) code[c] = row[c];
#pragma unroll
        for (int c = C0; c < C1; ++c) u[c] = my_cnt[code[c] >> 16];
#pragma unroll
        for (int c = C0; c < C1; ++c) {
            constexpr unsigned all = 0x3ff0u;
            const unsigned kmask = (((1u << (c - C0)) - 1u) << 4) & all;
            const int uu = u[c] - __popc(code[c] & kmask) - 1;
            e -= uu * (int)(code[c] & 0xfu);
            my_cnt[code[c] >> 16] = (unsigned char)uu;
        }
    };

    // prologue: bring in the WIN rows of the first window (row i_begin + r lives in slot r)
    for (int r = 0; r < WIN; ++r) {
        if (ANG == 0) sm.qring[r * NT + t] = load_q(i_begin + r);
        __syncthreads();
        if (r - DR >= 0) enter_pairs(r - DR, r);
        __syncthreads();
        if (r - DR >= 0) publish_code(r - DR);
        __syncthreads();
        if (has_win && r - DR >= 0) energy_add(r - DR);
    }
    unsigned char q_next = 0;  // register prefetch of the next row's sample
    if (ANG == 0 && i_begin + 1 < i_end) q_next = load_q(i_begin + WIN);

    int s_top = 0;       // ring slot of image row i (the top row of the current window)
    int s_pending = -1;  // ring slot whose pairs were published in the previous iteration and still have to enter E
    for (int i = i_begin; i < i_end; ++i) {
        const bool more = i + 1 < i_end;
        const int s_new = slot_add(s_top, WIN);  // slot that receives image row i + WIN (== slot of row i - 1)
        // publish column sums; image row i + WIN enters the ring
        sm.xch[ANG * NT + t] = make_uint4(cw1, cw2, (unsigned)csh, (unsigned)(csh >> 32));
        if (ANG == 0 && more) sm.qring[s_new * NT + t] = q_next;
        __syncthreads();  // #1
        if (ANG == 0 && more && i + 2 < i_end) q_next = load_q(i + WIN + 1);
        if (has_win) {
            if (s_pending >= 0) energy_add(s_pending);
            // window sums of this angle -> its share of the five properties
            unsigned w1 = 0, w2 = 0;
            unsigned long long sh = 0;
            const uint4* xa = sm.xch + ANG * NT + t;
#pragma unroll
            for (int c = C0; c < C1; ++c) {
                const uint4 v = xa[c];
                w1 += v.x;
                w2 += v.y;
                sh += (unsigned long long)v.z | ((unsigned long long)v.w << 32);
            }
            const int s1 = (int)(w1 & 0x1fffu), sab = (int)(w1 >> 13);
            const int sa = WIDE ? (int)(sh >> 50) : (int)(w2 & 0x3fffu), sq = WIDE ? (int)w2 : (int)(w2 >> 14);
            const int neq = WIDE ? (int)((sh >> 43) & 0x7fu) : (int)(sh >> 52);
            const unsigned long long shom = sh & ((1ull << (WIDE ? 43 : 52)) - 1ull);
            float* o = sm.outx + ANG * 5 * NT + t;
            o[0 * NT] = (float)(sq - 2 * sab) * inv_n;
            o[1 * NT] = (float)s1 * inv_n;
            o[2 * NT] = (float)((double)shom * (WIDE ? 1.4551915228366852e-11 : 9.094947017729282e-13)) * inv_n;  // 2^-36 / 2^-40
            o[3 * NT] = sqrtf((float)(e + 2 * NPAIR + 2 * neq)) * (0.5f * inv_n);
            const int var_num = 2 * NPAIR * sq - sa * sa, cov_num = 4 * NPAIR * sab - sa * sa;
            o[4 * NT] = var_num <= 0 ? 1.f : (float)cov_num / (float)var_num;
            if (moments) {  // validation dump of the exact integers the five values above were made from (rsx_glcm_moments)
                long long* m = moments + (((int64_t)i * out_cols + j0 + t) * 4 + ANG) * 8;
                m[0] = NPAIR, m[1] = s1, m[2] = sa, m[3] = sq, m[4] = sab, m[5] = e + 2 * NPAIR + 2 * neq, m[6] = neq;
                m[7] = (long long)shom;  // 2^40-scaled terms (levels <= 32), 2^36-scaled for the wide variant
            }
        }
        int s_anchor = -1;
        if (more) {
            // the window moves down: pairs anchored in image row i leave, pairs completed by row i + WIN enter
            leave_pairs(s_top, slot_add(s_top, DR));
            if (has_win) energy_sub(s_top);
            s_anchor = DR ? slot_add(s_top, WIN - 1) : s_new;
            enter_pairs(s_anchor, s_new);
        }
        __syncthreads();  // #2: outx complete, cell ids of the entering pairs visible
        if (has_win) {
            // combine the four angles: thread `ang` writes property `ang` (angle-0 threads also property 4)
            const int64_t o = (int64_t)i * out_cols + j0 + t;
            {
                constexpr int k = ANG;
                const float v = (sm.outx[(0 * 5 + k) * NT + t] + sm.outx[(1 * 5 + k) * NT + t]) + (sm.outx[(2 * 5 + k) * NT + t] + sm.outx[(3 * 5 + k) * NT + t]);
                props[k * plane_stride + o] = v * 0.25f;
            }
            if (ANG == 0) {
                constexpr int k = 4;
                const float v = (sm.outx[(0 * 5 + k) * NT + t] + sm.outx[(1 * 5 + k) * NT + t]) + (sm.outx[(2 * 5 + k) * NT + t] + sm.outx[(3 * 5 + k) * NT + t]);
                props[k * plane_stride + o] = v * 0.25f;
            }
        }
        if (!more) break;
        publish_code(s_anchor);  // consumed after the next barrier #1
        s_pending = s_anchor;
        s_top = slot_add(s_top, 1);
    }
}

template <int WIN, int NT, bool WIDE, int FOLD>
__global__ void __launch_bounds__(NT * 4, (FOLD && WIN <= 7) ? 1280 / (NT * 4) : 1) glcm_dense_kernel(const uint8_t* __restrict__ q, int W, int L, int out_rows, int out_cols, int rows_per_cta,
                                                            int NTW, float* __restrict__ props, int64_t plane_stride, long long* __restrict__ moments,
                                                            const int* __restrict__ fold_mode) {
    // several variants are launched back to back; the one the span statistics selected (glcm_fold_select_kernel) does the work
    if (fold_mode && *fold_mode != FOLD) return;
    // NTW = windows per CTA, <= NT - (WIN - 1), capped by what the private counters leave of shared memory
    constexpr int RING = WIN + 1;
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ unsigned long long homog_fx[64];
    const int ncell = FOLD ? fold_cells(FOLD) : L * (L + 1) / 2;
    DenseShared sm;
    sm.xch = reinterpret_cast<uint4*>(dsm);
    sm.outx = reinterpret_cast<float*>(sm.xch + 4 * NT);
    sm.codes = reinterpret_cast<unsigned*>(sm.outx + 20 * NT);
    sm.base = sm.codes + 4 * RING * NT;
    sm.qring = reinterpret_cast<unsigned char*>(sm.base + 4 * (WIN + NT));
    sm.cnt = sm.qring + RING * NT;
    sm.homog_fx = homog_fx;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int ang = warp & 3, t = (warp >> 2) * 32 + (tid & 31);
    const int j0 = blockIdx.x * NTW;  // first window column == first image column of this CTA
    const int i_begin = blockIdx.y * rows_per_cta;
    const int i_end = min(out_rows, i_begin + rows_per_cta);
    if (i_begin >= i_end) return;

    for (int i = tid; i < (FOLD ? 4 * ((ncell + 3) / 4) * NT * 4 : 4 * ncell * NTW); i += 4 * NT) sm.cnt[i] = 0;
    for (int i = tid; i < 4 * (WIN + NT); i += 4 * NT) sm.base[i] = 0xfffffff0u;
    if (tid < 64)  // 2^40 (wide: 2^36) / (1+k^2); the k == 0 entry also counts the pair in the Neq field
        homog_fx[tid] = (unsigned long long)((WIDE ? 68719476736.0 : 1099511627776.0) / (1.0 + (double)tid * (double)tid) + 0.5) +
                        (tid == 0 ? (1ull << (WIDE ? 43 : 52)) : 0ull);
    __syncthreads();
    switch (ang) {  // warp-uniform
        case 0: glcm_dense_body<WIN, NT, 0, WIDE, FOLD>(sm, q, W, L, out_cols, i_begin, i_end, j0, t, NTW, props, plane_stride, moments); break;
        case 1: glcm_dense_body<WIN, NT, 1, WIDE, FOLD>(sm, q, W, L, out_cols, i_begin, i_end, j0, t, NTW, props, plane_stride, moments); break;
        case 2: glcm_dense_body<WIN, NT, 2, WIDE, FOLD>(sm, q, W, L, out_cols, i_begin, i_end, j0, t, NTW, props, plane_stride, moments); break;
        default: glcm_dense_body<WIN, NT, 3, WIDE, FOLD>(sm, q, W, L, out_cols, i_begin, i_end, j0, t, NTW, props, plane_stride, moments); break;
    }
}

static size_t dense_smem_bytes(int win, int nt, int ntw, int ncell, bool fold = false) {
    return (fold ? (size_t)4 * ((ncell + 3) / 4) * nt * 4 : (size_t)4 * ncell * ntw) + (size_t)nt * (16 * (win + 1) + 16 + 64 + 80 + (win + 1)) + 16 * win + 64;
}

template <int WIN, int NT, bool WIDE, int FOLD = 0>
static int launch_dense(const uint8_t* d_q, int W, int levels, int ntw, int out_rows, int out_cols, float* d_props, int64_t plane_stride, long long* d_moments,
                        cudaStream_t s, const int* d_fold_mode = nullptr) {
    const int ncell = FOLD ? fold_cells(FOLD) : levels * (levels + 1) / 2;
    const size_t smem = dense_smem_bytes(WIN, NT, ntw, ncell, FOLD != 0);
    auto kern = glcm_dense_kernel<WIN, NT, WIDE, FOLD>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);  // the counters want shared memory, not L1
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max(smem, (size_t)49152));
        if (e != cudaSuccess) {
            rsx_set_error("rsx_glcm_props: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
            return RSX_ERR_CUDA;
        }
        configured = smem;
    }
    const int gx = ceil_div(out_cols, ntw);
    // rows per CTA: balance whole waves over the SMs against the (WIN-1)-row prologue every CTA pays
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT * 4, smem);
    occ = max(occ, 1);
    const int slots = rsx_num_sms() * occ;
    int best_gy = 1;
    double best_eff = 0.0;
    for (int gy = 1; gy <= max(1, out_rows / (4 * WIN)); ++gy) {
        const int rows = ceil_div(out_rows, gy);
        const int64_t ctas = (int64_t)gx * ceil_div(out_rows, rows);
        const double waves = (double)ctas / slots;
        const double eff = waves / ceil(waves) * rows / (rows + WIN - 1.0);
        if (eff > best_eff + 1e-9) best_eff = eff, best_gy = gy;
    }
    const int rows_per_cta = ceil_div(out_rows, best_gy);
    const int gy = ceil_div(out_rows, rows_per_cta);
    kern<<<dim3(gx, gy), NT * 4, smem, s>>>(d_q, W, levels, out_rows, out_cols, rows_per_cta, ntw, d_props, plane_stride, d_moments, d_fold_mode);
    return rsx_check_launch("glcm_dense");
}

// pick the widest CTA whose private counters fit in shared memory (and whose counter offsets fit the 16-bit code field)
template <int WIN, bool WIDE>
static int dispatch_dense(const uint8_t* d_q, int W, int levels, int out_rows, int out_cols, float* d_props, int64_t plane_stride, long long* d_moments,
                          cudaStream_t s, const int* d_fold_mode = nullptr) {
    const int ncell = levels * (levels + 1) / 2;
    const size_t limit = (size_t)226 * 1024;
    auto windows = [&](int nt) {  // windows per CTA for this width (0 = does not fit)
        int ntw = nt - (WIN - 1);
        while (ntw > 0 && (dense_smem_bytes(WIN, nt, ntw, ncell) > limit || (size_t)ncell * ntw > 65535)) --ntw;
        return ntw;
    };
    // a narrower CTA wastes fewer columns per window but has fewer warps; prefer the widest that keeps >= 3/4 of its windows
    const int forced = rsx_option("glcm_nt", 0);
    const int nts[4] = {256, 128, 96, 64};
    for (int k = 0; k < 4; ++k) {
        const int nt = nts[k], ntw = windows(nt);
        if (forced ? (nt == forced && ntw >= 8) : (ntw * 4 >= (nt - (WIN - 1)) * 3)) {
            if (nt == 256) return launch_dense<WIN, 256, WIDE>(d_q, W, levels, ntw, out_rows, out_cols, d_props, plane_stride, d_moments, s, d_fold_mode);
            if (nt == 128) return launch_dense<WIN, 128, WIDE>(d_q, W, levels, ntw, out_rows, out_cols, d_props, plane_stride, d_moments, s, d_fold_mode);
            if (nt == 96) return launch_dense<WIN, 96, WIDE>(d_q, W, levels, ntw, out_rows, out_cols, d_props, plane_stride, d_moments, s, d_fold_mode);
            return launch_dense<WIN, 64, WIDE>(d_q, W, levels, ntw, out_rows, out_cols, d_props, plane_stride, d_moments, s, d_fold_mode);
        }
    }
    const int ntw = windows(32);
    if (ntw >= 8 && (!forced || forced == 32)) return launch_dense<WIN, 32, WIDE>(d_q, W, levels, ntw, out_rows, out_cols, d_props, plane_stride, d_moments, s, d_fold_mode);
    return -1;
}

// ----------------------------------------------------------------------------- folded counters: span flags + energy patch
// Windows whose levels span 8 or more values go to list 0, 16 or more to list 1 (the fold modulo 8 / 16 may merge two cells
// there): lists of window indices with capacity `cap` each; stats[0], stats[1] = entries appended so far.  A list that is full
// stops growing - its fold is not going to be selected - so a scene where every window is flagged costs no atomics.
// One CTA per tile of SPAN_TR x SPAN_TC windows: the clamped levels of the tile (+ win-1 halo) go to shared memory, a
// horizontal min/max pass over win columns, then a vertical one over win rows.
constexpr int SPAN_TR = 32, SPAN_TC = 128, SPAN_MAXW = 11;
__device__ __forceinline__ void span_append(bool flagged, unsigned idx, unsigned long long* __restrict__ counter, unsigned* __restrict__ list, unsigned cap) {
    const unsigned m = __ballot_sync(0xffffffffu, flagged);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned long long pos = 0;
    if (lane == 0 && *reinterpret_cast<volatile unsigned long long*>(counter) <= cap) pos = atomicAdd(counter, (unsigned long long)__popc(m)) + 1;
    pos = __shfl_sync(0xffffffffu, pos, 0);  // 0: the list was already over its capacity
    if (pos && flagged) {
        const unsigned long long at = pos - 1 + __popc(m & ((1u << lane) - 1u));
        if (at < cap) list[at] = idx;
    }
}
// Four windows per thread and step: the levels sit in shared memory as bytes, a word holds four neighbouring columns, and
// __vminu4 / __vmaxu4 take the running minimum / maximum of four windows at once (the window's column d is the word pair
// funnel-shifted by d bytes).  0.40 -> 0.1 ms for 49 M windows against one byte per thread and step.
constexpr int SPAN_RS = (SPAN_TC + SPAN_MAXW - 1 + 3) / 4 + 1;  // words per tile row (+1: the funnel shift reads one word ahead)
__global__ void __launch_bounds__(256) glcm_span_flag_kernel(const uint8_t* __restrict__ q, int W, int L, int win, int out_rows, int out_cols,
                                                             unsigned* __restrict__ list8, unsigned* __restrict__ list16, unsigned cap,
                                                             unsigned long long* __restrict__ stats) {
    __shared__ unsigned raw[SPAN_TR + SPAN_MAXW - 1][SPAN_RS];
    __shared__ unsigned hmn[SPAN_TR + SPAN_MAXW - 1][SPAN_TC / 4], hmx[SPAN_TR + SPAN_MAXW - 1][SPAN_TC / 4];
    const int i0 = blockIdx.y * SPAN_TR, j0 = blockIdx.x * SPAN_TC;
    const int orows = min(SPAN_TR, out_rows - i0), ocols = min(SPAN_TC, out_cols - j0);
    const int rows = orows + win - 1, cols = ocols + win - 1;
    // the tile: whole words where the image rows are word aligned (W a multiple of 4) and the word lies inside the row
    const bool words_ok = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(q) & 3) == 0;
    const unsigned lmax = (unsigned)(L - 1) * 0x01010101u;
    for (int k = threadIdx.x; k < rows * SPAN_RS; k += 256) {
        const int r = k / SPAN_RS, cw = k - r * SPAN_RS;
        const uint8_t* src = q + (int64_t)(i0 + r) * W + j0 + 4 * cw;
        unsigned w = 0;
        if (4 * cw + 3 < cols && words_ok) {
            w = *reinterpret_cast<const unsigned*>(src);
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (4 * cw + b < cols) w |= (unsigned)src[b] << (8 * b);
        }
        raw[r][cw] = __vminu4(w, lmax);
    }
    __syncthreads();
    // horizontal pass: min / max over win columns, for the columns 4g .. 4g+3 of every tile row
    for (int k = threadIdx.x; k < rows * (SPAN_TC / 4); k += 256) {
        const int r = k / (SPAN_TC / 4), g = k - r * (SPAN_TC / 4);
        unsigned mn = 0xffffffffu, mx = 0u;
        for (int d = 0; d < win; ++d) {
            const unsigned v = __funnelshift_r(raw[r][g + (d >> 2)], raw[r][g + (d >> 2) + 1], 8 * (d & 3));
            mn = __vminu4(mn, v), mx = __vmaxu4(mx, v);
        }
        hmn[r][g] = mn, hmx[r][g] = mx;
    }
    __syncthreads();
    // vertical pass + flags; block-uniform trip count: span_append is a warp collective
    for (int k0 = 0; k0 < SPAN_TR * (SPAN_TC / 4); k0 += 256) {
        const int k = k0 + threadIdx.x;
        const int r = k / (SPAN_TC / 4), g = k - r * (SPAN_TC / 4);
        unsigned span = 0;
        if (r < orows) {
            unsigned mn = 0xffffffffu, mx = 0u;
            for (int d = 0; d < win; ++d) mn = __vminu4(mn, hmn[r + d][g]), mx = __vmaxu4(mx, hmx[r + d][g]);
            span = __vsub4(mx, mn);
        }
        const unsigned base = (unsigned)((int64_t)(i0 + r) * out_cols + j0 + 4 * g);
        // columns past the tile's last window never count
        const int nb = r < orows ? min(4, ocols - 4 * g) : 0;
        if (nb < 4) span &= nb <= 0 ? 0u : (0xffffffffu >> (8 * (4 - nb)));
        if (!__any_sync(0xffffffffu, __vcmpgeu4(span, 0x08080808u) != 0u)) continue;  // nothing to list in this warp (the common case)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const unsigned sp = (span >> (8 * b)) & 0xffu;
            span_append(sp >= 8u, base + b, &stats[0], list8, cap);
            span_append(sp >= 16u, base + b, &stats[1], list16, cap);
        }
    }
}

// mode = 8 / 16: the fold whose flagged share is at most 1/256 of the windows (their energy is recomputed one warp per window, ~50x
// the cost of a dense window); 0: the unfolded kernel.  Levels that fit the fold need no flags at all.
__global__ void glcm_fold_select_kernel(const unsigned long long* __restrict__ stats, unsigned cap, int levels, int force, int* __restrict__ mode) {
    int m = 0;
    if (levels <= 8 || stats[0] <= cap) m = 8;
    else if (levels <= 16 || stats[1] <= cap) m = 16;
    if ((force == 8 && stats[0] <= cap) || (force == 16 && stats[1] <= cap) || force == 0) m = force;  // a forced fold needs its whole list
    *mode = m;
}

// Exact energy of the flagged windows, one warp per window, with the float operations of glcm_dense_body (so that a patched
// value equals what the unfolded dense kernel writes): per angle E = sum over the pairs of w * U[cell of the pair] from a
// warp-private histogram of all L(L+1)/2 unordered cells, energy = sqrtf(E) * (0.5 / n); mean of the four angles.
__global__ void __launch_bounds__(256) glcm_energy_patch_kernel(const uint8_t* __restrict__ q, int W, int L, int win, int out_rows, int out_cols,
                                                                const unsigned* __restrict__ list8, const unsigned* __restrict__ list16,
                                                                const unsigned long long* __restrict__ stats, const int* __restrict__ fold_mode,
                                                                float* __restrict__ energy, long long* __restrict__ moments) {
    extern __shared__ unsigned patch_hist[];  // [8 warps][L (L + 1) / 2]
    const int mode = *fold_mode;
    if (mode == 0) return;
    const unsigned* list = mode == 8 ? list8 : list16;
    const int64_t n_list = (int64_t)stats[mode == 8 ? 0 : 1];  // <= the capacity, or this fold would not have been selected
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncell = L * (L + 1) / 2;
    unsigned* hist = patch_hist + warp * ncell;
    for (int i = lane; i < ncell; i += 32) hist[i] = 0;
    __syncwarp();
    const int64_t warps_total = (int64_t)gridDim.x * 8;
    for (int64_t k = (int64_t)blockIdx.x * 8 + warp; k < n_list; k += warps_total) {
        {
            const int64_t o = (int64_t)list[k];
            const int oi = (int)(o / out_cols), oj = (int)(o - (int64_t)oi * out_cols);
            const uint8_t* wbase = q + (int64_t)oi * W + oj;
            float part[4];
#pragma unroll
            for (int ang = 0; ang < 4; ++ang) {
                const int dr = ang == 0 ? 0 : 1;
                const int dc = ang == 0 ? 1 : (ang == 1 ? 1 : (ang == 2 ? 0 : -1));
                const int nrows = win - dr, ncols = win - (dc != 0 ? 1 : 0), c0 = dc < 0 ? 1 : 0;
                const int n = nrows * ncols, off = dr * W + dc;
                for (int sidx = lane; sidx < n; sidx += 32) {
                    const int r = sidx / ncols, c = sidx - r * ncols + c0;
                    const uint8_t* pp = wbase + r * W + c;
                    const int a = min((int)pp[0], L - 1), b = min((int)pp[off], L - 1);
                    atomicAdd(&hist[tri_cell(a, b)], 1u);
                }
                __syncwarp();
                int e = 0;
                for (int sidx = lane; sidx < n; sidx += 32) {
                    const int r = sidx / ncols, c = sidx - r * ncols + c0;
                    const uint8_t* pp = wbase + r * W + c;
                    const int a = min((int)pp[0], L - 1), b = min((int)pp[off], L - 1);
                    e += (a != b ? 2 : 4) * (int)hist[tri_cell(a, b)];
                }
                __syncwarp();
                for (int sidx = lane; sidx < n; sidx += 32) {
                    const int r = sidx / ncols, c = sidx - r * ncols + c0;
                    const uint8_t* pp = wbase + r * W + c;
                    const int a = min((int)pp[0], L - 1), b = min((int)pp[off], L - 1);
                    hist[tri_cell(a, b)] = 0;
                }
                __syncwarp();
                e = __reduce_add_sync(0xffffffffu, e);
                if (moments && lane == 0) moments[(o * 4 + ang) * 8 + 5] = e;
                const float inv_n = 1.f / (float)n;
                part[ang] = __fmul_rn(sqrtf((float)e), 0.5f * inv_n);  // rounded product, as the dense kernel stores it (no FMA with the sum)
            }
            if (lane == 0) energy[o] = __fmul_rn(__fadd_rn(__fadd_rn(part[0], part[1]), __fadd_rn(part[2], part[3])), 0.25f);
        }
    }
}

// scratch of the folded path, per device: [0,16) two counters, [16,20) the selected mode, [64, ...) two lists of `cap` window indices
static int g_patch_smem_configured = 0;
static uint8_t* g_span_buf = nullptr;
static size_t g_span_cap = 0;
static int g_span_dev = -1;

// Folded dense path: span lists + statistics, fold selection on the device, the three dense variants back to back (two of them
// return at once), energy patch for the listed windows.  Returns -1 when not applicable.
template <int WIN, bool WIDE>
static int dispatch_dense_folded(const uint8_t* d_q, int W, int levels, int out_rows, int out_cols, float* d_props, int64_t plane_stride, long long* d_moments,
                                 cudaStream_t s) {
    const size_t n_win = (size_t)out_rows * out_cols;
    if (n_win >= ((size_t)1 << 32)) return -1;  // window indices are 32-bit
    // a fold is used while at most 1/256 of the windows need the patch (one warp per window, ~50x the cost of a dense window)
    const unsigned cap = (unsigned)(n_win / (size_t)max(1, rsx_option("glcm_fold_cap_div", 256)) + 1);
    const size_t need = 64 + (size_t)2 * cap * sizeof(unsigned);
    int dev = 0;
    cudaGetDevice(&dev);
    if (need > g_span_cap || dev != g_span_dev) {
        if (g_span_buf && dev == g_span_dev) cudaFree(g_span_buf);
        g_span_buf = nullptr, g_span_cap = 0, g_span_dev = dev;
        if (cudaMalloc(&g_span_buf, need) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        g_span_cap = need;
    }
    unsigned long long* stats = reinterpret_cast<unsigned long long*>(g_span_buf);
    int* mode = reinterpret_cast<int*>(g_span_buf + 16);
    unsigned* list8 = reinterpret_cast<unsigned*>(g_span_buf + 64);
    unsigned* list16 = list8 + cap;
    if (cudaMemsetAsync(g_span_buf, 0, 64, s) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    glcm_span_flag_kernel<<<dim3(ceil_div(out_cols, SPAN_TC), ceil_div(out_rows, SPAN_TR)), 256, 0, s>>>(d_q, W, levels, WIN, out_rows, out_cols, list8, list16,
                                                                                                       cap, stats);
    if (int rc = rsx_check_launch("glcm_span_flags")) return rc;
    glcm_fold_select_kernel<<<1, 1, 0, s>>>(stats, cap, levels, rsx_option("glcm_fold_force", -1), mode);
    if (int rc = rsx_check_launch("glcm_fold_select")) return rc;
    const int nt = rsx_option("glcm_fold_nt", 64);
    int rc_d;
    if (nt == 32) {
        rc_d = launch_dense<WIN, 32, WIDE, 8>(d_q, W, levels, 32 - (WIN - 1), out_rows, out_cols, d_props, plane_stride, d_moments, s, mode);
        if (!rc_d) rc_d = launch_dense<WIN, 32, WIDE, 16>(d_q, W, levels, 32 - (WIN - 1), out_rows, out_cols, d_props, plane_stride, d_moments, s, mode);
    } else {
        rc_d = launch_dense<WIN, 64, WIDE, 8>(d_q, W, levels, 64 - (WIN - 1), out_rows, out_cols, d_props, plane_stride, d_moments, s, mode);
        if (!rc_d) rc_d = launch_dense<WIN, 64, WIDE, 16>(d_q, W, levels, 64 - (WIN - 1), out_rows, out_cols, d_props, plane_stride, d_moments, s, mode);
    }
    if (rc_d) return rc_d;
    if (int rc = dispatch_dense<WIN, WIDE>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s, mode)) {  // mode 0
        if (rc < 0) rsx_set_error("rsx_glcm_props: no dense configuration for window %d at %d levels", WIN, levels);
        return rc < 0 ? RSX_ERR_UNSUPPORTED : rc;
    }
    const int grid = (int)min((int64_t)ceil_div((int64_t)cap, (int64_t)8), (int64_t)rsx_num_sms() * 8);
    const int patch_smem = 8 * (levels * (levels + 1) / 2) * 4;
    int& patch_configured = g_patch_smem_configured;  // one kernel, one attribute: shared by every instantiation of this dispatcher
    if (patch_smem > patch_configured) {
        if (cudaFuncSetAttribute(glcm_energy_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max(patch_smem, 48 * 1024)) != cudaSuccess) {
            cudaGetLastError();
            rsx_set_error("rsx_glcm_props: energy patch kernel needs %d B of shared memory", patch_smem);
            return RSX_ERR_CUDA;
        }
        patch_configured = patch_smem;
    }
    glcm_energy_patch_kernel<<<grid, 256, patch_smem, s>>>(d_q, W, levels, WIN, out_rows, out_cols, list8, list16, stats, mode, d_props + 3 * plane_stride,
                                                               d_moments);
    return rsx_check_launch("glcm_energy_patch");
}

static int glcm_run(const uint8_t* d_q, int rows_avail, int W, int levels, int window, int step, int out_rows, int out_cols, float* d_props,
                    int64_t plane_stride, long long* d_moments, rsx_stream_t stream, const GlcmOffsets* custom = nullptr) {
    RSX_REQUIRE(d_q && d_props, "rsx_glcm_props: null argument");
    RSX_REQUIRE(levels >= 2 && levels <= 128, "rsx_glcm_props: levels must be in [2,128]");
    RSX_REQUIRE(window >= 2 && window <= 127 && step >= 1, "rsx_glcm_props: window must be in [2,127], step >= 1");
    RSX_REQUIRE(out_rows >= 1 && out_cols >= 1 && (int64_t)(out_rows - 1) * step + window <= rows_avail && (int64_t)(out_cols - 1) * step + window <= W,
                "rsx_glcm_props: windows exceed the %dx%d band", rows_avail, W);
    RSX_REQUIRE(plane_stride >= (int64_t)out_rows * out_cols, "rsx_glcm_props: plane_stride too small");
    if (int rc = ensure_homog()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int ncell = levels * (levels + 1) / 2;
    // dense fast path: packed integer moments hold for levels <= 64 and window <= 11; uint8 counters hold w(w-1) <= 110
    if (step == 1 && levels <= 64 && !custom) {
        int rc = -1;
        const int fold_env = rsx_option("glcm_fold", 1);  // 1: folded counters where the scene allows (chosen on the device), 0: never
        if (fold_env && levels > 16) {  // up to 16 levels the unfolded counters are as small as the folded ones
            switch (window) {
                case 3: rc = levels <= 32 ? dispatch_dense_folded<3, false>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s) : dispatch_dense_folded<3, true>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s); break;
                case 5: rc = levels <= 32 ? dispatch_dense_folded<5, false>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s) : dispatch_dense_folded<5, true>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s); break;
                case 7: rc = levels <= 32 ? dispatch_dense_folded<7, false>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s) : dispatch_dense_folded<7, true>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s); break;
                case 9: rc = levels <= 32 ? dispatch_dense_folded<9, false>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s) : dispatch_dense_folded<9, true>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s); break;
                case 11: rc = levels <= 32 ? dispatch_dense_folded<11, false>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s) : dispatch_dense_folded<11, true>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s); break;
                default: break;
            }
            if (rc >= 0) return rc;
        }
#define DENSE(WW)                                                                                                    \
    case WW:                                                                                                         \
        rc = levels <= 32 ? dispatch_dense<WW, false>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s)  \
                          : dispatch_dense<WW, true>(d_q, W, levels, out_rows, out_cols, d_props, plane_stride, d_moments, s);  \
        break;
        switch (window) {
            DENSE(3) DENSE(5) DENSE(7) DENSE(9) DENSE(11)
            default: break;
        }
#undef DENSE
        if (rc >= 0) return rc;
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
    glcm_props_warp_kernel<<<grid, 128, smem, s>>>(d_q, W, levels, window, step, out_rows, out_cols, d_props, plane_stride, d_moments,
                                                   custom ? *custom : default_offsets());
    return rsx_check_launch("glcm_props_warp");
}

extern "C" int rsx_glcm_props_offsets(const uint8_t* d_q, int rows_avail, int W, int levels, int window, int step, int out_rows, int out_cols,
                                      const int32_t* h_offsets, int n_offsets, float* d_props, int64_t plane_stride, rsx_stream_t stream) {
    RSX_REQUIRE(h_offsets && n_offsets >= 1 && n_offsets <= 16, "rsx_glcm_props_offsets: 1..16 (row, column) offsets");
    GlcmOffsets o = default_offsets();
    o.n = n_offsets;
    for (int i = 0; i < n_offsets; ++i) {
        o.dr[i] = h_offsets[2 * i], o.dc[i] = h_offsets[2 * i + 1];
        RSX_REQUIRE(abs(o.dr[i]) < 4096 && abs(o.dc[i]) < 4096, "rsx_glcm_props_offsets: offset out of range");
    }
    return glcm_run(d_q, rows_avail, W, levels, window, step, out_rows, out_cols, d_props, plane_stride, nullptr, stream, &o);
}

extern "C" int rsx_glcm_props(const uint8_t* d_q, int rows_avail, int W, int levels, int window, int step, int out_rows, int out_cols,
                              float* d_props, int64_t plane_stride, rsx_stream_t stream) {
    return glcm_run(d_q, rows_avail, W, levels, window, step, out_rows, out_cols, d_props, plane_stride, nullptr, stream);
}

extern "C" int rsx_glcm_moments(const uint8_t* d_q, int rows_avail, int W, int levels, int window, int step, int out_rows, int out_cols,
                                float* d_props, int64_t plane_stride, int64_t* d_moments, rsx_stream_t stream) {
    RSX_REQUIRE(d_moments, "rsx_glcm_moments: null argument");
    return glcm_run(d_q, rows_avail, W, levels, window, step, out_rows, out_cols, d_props, plane_stride, reinterpret_cast<long long*>(d_moments), stream);
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
                                                              int64_t dst_stride, double scale_x, double scale_y, int rows_per_cta,
                                                              uint32_t* __restrict__ minmax) {
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
        // a thread walks down a run of destination rows, four at a time: the 16 taps of four rows are requested before any
        // of them is used (duplicates between neighbouring rows / threads are L1 hits), so that enough loads are in flight
        const int ly0 = blockIdx.y * rows_per_cta, ly1 = min(dst_rows, ly0 + rows_per_cta);
        for (int lyb = ly0; lyb < ly1; lyb += 4) {
            float p0[4], q0[4], p1[4], q1[4], fy[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int dy = dst_row0 + min(lyb + u, ly1 - 1);
                const double cy = (dy + 0.5) * scale_y - 0.5;
                const double fyd = floor(cy);
                fy[u] = (float)(cy - fyd);
                const int sy = (int)fyd;
                const int y0 = min(max(sy, 0), src_h - 1) - src_row0, y1 = min(max(sy + 1, 0), src_h - 1) - src_row0;
                const bool ok0 = y0 >= 0 && y0 < src_rows_avail, ok1 = y1 >= 0 && y1 < src_rows_avail;
                p0[u] = ok0 ? sp[(int64_t)y0 * src_w + x0] : 0.f;
                q0[u] = ok0 ? sp[(int64_t)y0 * src_w + x1] : 0.f;
                p1[u] = ok1 ? sp[(int64_t)y1 * src_w + x0] : 0.f;
                q1[u] = ok1 ? sp[(int64_t)y1 * src_w + x1] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (lyb + u < ly1) {
                    const float r0 = fmaf(f_sub(q0[u], p0[u]), fx, p0[u]);
                    const float r1 = fmaf(f_sub(q1[u], p1[u]), fx, p1[u]);
                    const float v = fmaf(f_sub(r1, r0), fy[u], r0);
                    dp[(int64_t)(lyb + u) * dst_w + dx] = v;
                    mn = fminf(mn, v), mx = fmaxf(mx, v);
                }
            }
        }
    }
    if (minmax) warp_minmax_commit(mn, mx, minmax + 2 * plane);
}

// Fast path for scale factors up to ~1 (the dense GLCM map, (H-w+1) x (W-w+1) -> H x W): a CTA makes a 32 x 256 tile of one
// destination plane from the source rows / columns it touches, staged in shared memory with coalesced row loads (the general
// kernel above fetches four scattered taps per output: 0.72 ms for five 49 Mpx planes against 0.30 ms of HBM time).  Same
// arithmetic, tap for tap: the weights and clamped tap indices of a row are computed once per CTA (double), of a column once per
// thread.
constexpr int RSZ_TR = 32, RSZ_TC = 256, RSZ_SR = RSZ_TR + 4, RSZ_SC = RSZ_TC + 8;
__global__ void __launch_bounds__(256) resize_bilinear_tile_kernel(const float* __restrict__ src, int src_h, int src_w, int src_row0, int src_rows_avail,
                                                                   int64_t src_stride, float* __restrict__ dst, int dst_w, int dst_row0, int dst_rows,
                                                                   int64_t dst_stride, double scale_x, double scale_y, uint32_t* __restrict__ minmax) {
    __shared__ float tile[RSZ_SR][RSZ_SC];
    __shared__ float fy_s[RSZ_TR];
    __shared__ int y0_s[RSZ_TR], y1_s[RSZ_TR];
    const int plane = blockIdx.z;
    const float* sp = src + plane * src_stride;
    float* dp = dst + plane * dst_stride;
    const int dx0 = blockIdx.x * RSZ_TC, ly0 = blockIdx.y * RSZ_TR;
    const int n_rows = min(RSZ_TR, dst_rows - ly0), n_cols = min(RSZ_TC, dst_w - dx0);
    auto tap = [](int d, double scale, int n, int& i0, int& i1) -> float {
        const double c = (d + 0.5) * scale - 0.5;
        const double f = floor(c);
        const int s = (int)f;
        i0 = min(max(s, 0), n - 1), i1 = min(max(s + 1, 0), n - 1);
        return (float)(c - f);
    };
    // source footprint of the tile (taps are monotone in the destination index)
    int a, b, sy_lo, sy_hi, sx_lo, sx_hi;
    tap(dst_row0 + ly0, scale_y, src_h, sy_lo, a);
    tap(dst_row0 + ly0 + n_rows - 1, scale_y, src_h, b, sy_hi);
    tap(dx0, scale_x, src_w, sx_lo, a);
    tap(dx0 + n_cols - 1, scale_x, src_w, b, sx_hi);
    const int f_rows = sy_hi - sy_lo + 1, f_cols = sx_hi - sx_lo + 1;  // the launcher guarantees they fit the tile
    if (threadIdx.x < n_rows) {
        int y0, y1;
        fy_s[threadIdx.x] = tap(dst_row0 + ly0 + threadIdx.x, scale_y, src_h, y0, y1);
        y0_s[threadIdx.x] = y0 - sy_lo, y1_s[threadIdx.x] = y1 - sy_lo;
    }
    for (int r = threadIdx.x >> 5; r < f_rows; r += 8) {  // a warp takes whole rows: all of a row's loads are in flight together
        const int gy = sy_lo + r - src_row0;
        const bool ok = gy >= 0 && gy < src_rows_avail;
        const float* row = sp + (int64_t)gy * src_w + sx_lo;
        constexpr int NCH = (RSZ_SC + 31) / 32;
        float v[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = (threadIdx.x & 31) + 32 * k;
            v[k] = (ok && c < f_cols) ? __ldg(row + c) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = (threadIdx.x & 31) + 32 * k;
            if (c < f_cols) tile[r][c] = v[k];
        }
    }
    __syncthreads();
    float mn = INFINITY, mx = -INFINITY;
    if ((int)threadIdx.x < n_cols) {
        int x0, x1;
        const float fx = tap(dx0 + threadIdx.x, scale_x, src_w, x0, x1);
        x0 -= sx_lo, x1 -= sx_lo;
        float* out = dp + (int64_t)ly0 * dst_w + dx0 + threadIdx.x;
#pragma unroll 4
        for (int r = 0; r < n_rows; ++r) {
            const int y0 = y0_s[r], y1 = y1_s[r];
            const float p0 = tile[y0][x0], q0 = tile[y0][x1], p1 = tile[y1][x0], q1 = tile[y1][x1];
            const float r0 = fmaf(f_sub(q0, p0), fx, p0);
            const float r1 = fmaf(f_sub(q1, p1), fx, p1);
            const float v = fmaf(f_sub(r1, r0), fy_s[r], r0);
            out[(int64_t)r * dst_w] = v;
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
    // source footprint of a 32 x 256 destination tile: (n - 1) * scale + 3 samples at most
    if (rsx_option("resize_tiled", 1) && (RSZ_TR - 1) * scale_y + 3.0 <= RSZ_SR && (RSZ_TC - 1) * scale_x + 3.0 <= RSZ_SC) {
        dim3 tgrid(ceil_div(dst_w, RSZ_TC), ceil_div(dst_rows, RSZ_TR), n_planes);
        resize_bilinear_tile_kernel<<<tgrid, 256, 0, (cudaStream_t)stream>>>(d_src, src_h_total, src_w, src_row0, src_rows_avail, src_plane_stride, d_dst, dst_w,
                                                                           dst_row0, dst_rows, dst_plane_stride, scale_x, scale_y, d_minmax);
        return rsx_check_launch("resize_bilinear_tile");
    }
    const int gx = ceil_div(dst_w, 256);
    const int rows_per_cta = 32;  // many small CTAs: tens of waves, so the last partial wave costs a few percent at most
    dim3 grid(gx, ceil_div(dst_rows, rows_per_cta), n_planes);
    resize_bilinear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, src_h_total, src_w, src_row0, src_rows_avail, src_plane_stride, d_dst, dst_w,
                                                                   dst_row0, dst_rows, dst_plane_stride, scale_x, scale_y, rows_per_cta, d_minmax);
    return rsx_check_launch("resize_bilinear");
}

// ----------------------------------------------------------------------------- N1: add_spatial_context (indices.py:760-776)
// cv2.boxFilter(feature, -1, (k, k), normalize=True, borderType=BORDER_REFLECT) on float32 planes.  OpenCV sums float32
// sources in double (row sums, then column sums) and stores (float)(sum * (1.0 / (k*k))); the same arithmetic here with
// direct k-term sums (OpenCV's running sums differ from them by ~1e-16 relative, below float32 resolution).
// Row-strip sharding: source rows [src_row0, src_row0 + rows_avail) of an H_total-row image are present; destination rows
// [dst_row0, dst_row0 + dst_rows) are produced.
__device__ __forceinline__ int reflect_index(int i, int n) {  // BORDER_REFLECT: fedcba|abcdefgh|hgfedcb
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - i - 1;
    return min(max(i, 0), n - 1);
}

template <int KS>
__global__ void __launch_bounds__(128) box_mean_kernel(const float* __restrict__ src, int H_total, int W, int src_row0, int rows_avail, int64_t src_stride,
                                                       float* __restrict__ dst, int dst_row0, int dst_rows, int64_t dst_stride, int rows_per_cta,
                                                       uint32_t* __restrict__ minmax) {
    constexpr int R = KS / 2;
    __shared__ float row[128 + 2 * R];
    const int plane = blockIdx.z;
    const float* sp = src + plane * src_stride;
    float* dp = dst + plane * dst_stride;
    const int x0 = blockIdx.x * 128, x = x0 + threadIdx.x;
    const int ly0 = blockIdx.y * rows_per_cta, ly1 = min(dst_rows, ly0 + rows_per_cta);
    double ring[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) ring[k] = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    const double scale = 1.0 / (double)(KS * KS);
    // image rows gy0 - R .. gy1 - 1 + R stream through; an output row is complete when KS rows are in the ring
    for (int gy = dst_row0 + ly0 - R; gy < dst_row0 + ly1 + R; ++gy) {
        const int sy = reflect_index(gy, H_total) - src_row0;
        __syncthreads();
        for (int i = threadIdx.x; i < 128 + 2 * R; i += 128) {
            const int sx = reflect_index(x0 - R + i, W);
            row[i] = (sy >= 0 && sy < rows_avail) ? sp[(int64_t)sy * W + sx] : 0.f;
        }
        __syncthreads();
        double hs = 0.0;
#pragma unroll
        for (int k = 0; k < KS; ++k) hs += (double)row[threadIdx.x + k];
#pragma unroll
        for (int k = 0; k < KS - 1; ++k) ring[k] = ring[k + 1];
        ring[KS - 1] = hs;
        const int oy = gy - R;  // output row completed by this input row
        if (oy >= dst_row0 + ly0 && x < W) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < KS; ++k) s += ring[k];
            const float v = (float)(s * scale);
            dp[(int64_t)(oy - dst_row0) * W + x] = v;
            mn = fminf(mn, v), mx = fmaxf(mx, v);
        }
    }
    if (minmax) warp_minmax_commit(mn, mx, minmax + 2 * plane);
}

extern "C" int rsx_box_mean_f32(const float* d_src, int H_total, int W, int src_row0, int rows_avail, int64_t src_plane_stride, float* d_dst,
                                int dst_row0, int dst_rows, int64_t dst_plane_stride, int n_planes, int ksize, uint32_t* d_minmax,
                                rsx_stream_t stream) {
    RSX_REQUIRE(d_src && d_dst && H_total >= 1 && W >= 1 && dst_rows >= 1 && n_planes >= 1, "rsx_box_mean_f32: bad arguments");
    RSX_REQUIRE(ksize >= 3 && ksize <= 11 && (ksize & 1), "rsx_box_mean_f32: window size must be odd, 3..11");
    RSX_REQUIRE(ksize / 2 <= H_total && ksize / 2 <= W, "rsx_box_mean_f32: window larger than the image");
    {
        const int r = ksize / 2;
        const int lo = max(dst_row0 - r, -(r)), hi = dst_row0 + dst_rows - 1 + r;
        // reflected rows stay inside [0, H_total); the non-reflected range must be available
        RSX_REQUIRE(max(dst_row0 - r, 0) >= src_row0 && min(hi, H_total - 1) < src_row0 + rows_avail, "rsx_box_mean_f32: halo rows missing");
        (void)lo;
    }
    const int gx = ceil_div(W, 128);
    const int rows_per_cta = max(32, ceil_div(dst_rows, max(1, rsx_num_sms() * 8 / max(1, gx * n_planes))));
    dim3 grid(gx, ceil_div(dst_rows, rows_per_cta), n_planes);
    cudaStream_t s = (cudaStream_t)stream;
#define BOX(K)                                                                                                                                   \
    case K:                                                                                                                                      \
        box_mean_kernel<K><<<grid, 128, 0, s>>>(d_src, H_total, W, src_row0, rows_avail, src_plane_stride, d_dst, dst_row0, dst_rows, dst_plane_stride, \
                                                 rows_per_cta, d_minmax);                                                                       \
        break;
    switch (ksize) {
        BOX(3) BOX(5) BOX(7) BOX(9) BOX(11)
    }
#undef BOX
    return rsx_check_launch("box_mean");
}
