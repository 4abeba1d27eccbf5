/* Oracle (test infrastructure): plain-C restatement of the windowed GLCM texture loop.
 *
 * Follows modules/features/indices.py:283-305 (top-left anchored windows, distance 1,
 * angles 0, pi/4, pi/2, 3pi/4, symmetric + normed co-occurrence matrix, five properties
 * averaged over the four angles, stored as float32) with the co-occurrence arithmetic of
 * scikit-image's graycomatrix/graycoprops as described in oracle/glcm.py (third-party,
 * absent from /root/reference: parity unpinned by the reference itself).
 *
 * Never linked into the product library; used by tests/ and bench.py's CPU baseline only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* (drow, dcol) for angle 0, pi/4, pi/2, 3pi/4 at distance 1: round(sin), round(cos). */
static const int DR[4] = {0, 1, 1, 1};
static const int DC[4] = {1, 1, 0, -1};

static void window_counts(const uint8_t *q, int W, int i0, int j0, int win, int L, uint32_t *cnt /* [4][L][L] */)
{
    memset(cnt, 0, sizeof(uint32_t) * 4 * L * L);
    for (int a = 0; a < 4; ++a) {
        int dr = DR[a], dc = DC[a];
        int r0 = dr < 0 ? -dr : 0, r1 = dr > 0 ? win - dr : win;
        int c0 = dc < 0 ? -dc : 0, c1 = dc > 0 ? win - dc : win;
        uint32_t *pa = cnt + (size_t)a * L * L;
        for (int r = r0; r < r1; ++r)
            for (int c = c0; c < c1; ++c) {
                int u = q[(size_t)(i0 + r) * W + j0 + c];
                int v = q[(size_t)(i0 + r + dr) * W + j0 + c + dc];
                if (u < L && v < L)
                    pa[u * L + v] += 1;
            }
    }
}

static void window_props(const uint32_t *cnt, int L, double out[5])
{
    double acc[5] = {0, 0, 0, 0, 0};
    for (int a = 0; a < 4; ++a) {
        const uint32_t *pa = cnt + (size_t)a * L * L;
        double total = 0;
        for (int i = 0; i < L; ++i)
            for (int j = 0; j < L; ++j)
                total += (double)pa[i * L + j] + (double)pa[j * L + i];
        if (total == 0)
            total = 1;
        double con = 0, dis = 0, hom = 0, asm_ = 0, mi = 0, mj = 0;
        for (int i = 0; i < L; ++i)
            for (int j = 0; j < L; ++j) {
                double p = ((double)pa[i * L + j] + (double)pa[j * L + i]) / total;
                if (p == 0)
                    continue;
                double d = (double)(i - j);
                con += p * d * d;
                dis += p * fabs(d);
                hom += p / (1.0 + d * d);
                asm_ += p * p;
                mi += i * p;
                mj += j * p;
            }
        double vi = 0, vj = 0, cov = 0;
        for (int i = 0; i < L; ++i)
            for (int j = 0; j < L; ++j) {
                double p = ((double)pa[i * L + j] + (double)pa[j * L + i]) / total;
                if (p == 0)
                    continue;
                vi += p * (i - mi) * (i - mi);
                vj += p * (j - mj) * (j - mj);
                cov += p * (i - mi) * (j - mj);
            }
        double si = sqrt(vi), sj = sqrt(vj);
        double cor = (si < 1e-15 || sj < 1e-15) ? 1.0 : cov / (si * sj);
        acc[0] += con;
        acc[1] += dis;
        acc[2] += hom;
        acc[3] += sqrt(asm_);
        acc[4] += cor;
    }
    for (int k = 0; k < 5; ++k)
        out[k] = acc[k] / 4.0;
}

/* out: float32 [5][oh][ow] */
int oracle_glcm_props(const uint8_t *q, int H, int W, int L, int win, int step, float *out, int threads)
{
    if (win > H || win > W || L < 1 || L > 256 || step < 1)
        return 1;
    int oh = (H - win) / step + 1, ow = (W - win) / step + 1;
#ifdef _OPENMP
    if (threads > 0)
        omp_set_num_threads(threads);
#endif
    int fail = 0;
#pragma omp parallel
    {
        uint32_t *cnt = (uint32_t *)malloc(sizeof(uint32_t) * 4 * L * L);
        if (!cnt) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int oi = 0; oi < oh; ++oi)
                for (int oj = 0; oj < ow; ++oj) {
                    double p[5];
                    window_counts(q, W, oi * step, oj * step, win, L, cnt);
                    window_props(cnt, L, p);
                    for (int k = 0; k < 5; ++k)
                        out[((size_t)k * oh + oi) * ow + oj] = (float)p[k];
                }
            free(cnt);
        }
    }
    return fail ? 2 : 0;
}

/* out: uint32 [oh][ow][4][L][L] directed counts */
int oracle_glcm_counts(const uint8_t *q, int H, int W, int L, int win, int step, uint32_t *out)
{
    if (win > H || win > W || L < 1 || L > 256 || step < 1)
        return 1;
    int oh = (H - win) / step + 1, ow = (W - win) / step + 1;
    for (int oi = 0; oi < oh; ++oi)
        for (int oj = 0; oj < ow; ++oj)
            window_counts(q, W, oi * step, oj * step, win, L, out + ((size_t)oi * ow + oj) * 4 * L * L);
    return 0;
}
