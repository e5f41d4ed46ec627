/* ORACLE (test infrastructure, not product): deterministic plain-C restatement of
 * the reference's lifting arithmetic.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library.
 *
 * Why C next to oracle/ref_lift.py: the torch restatement inherits whatever
 * MKL/ATen kernels the host CPU dispatches to; this file fixes the arithmetic to
 * explicit IEEE-754 binary32 operations (fmaf, +, *, /, sqrtf; build with
 * -ffp-contract=off) so the checker gives the same bits on every host.  It is
 * pinned against fixtures produced by the reference's own utils/pcd.py and
 * kitti_utils.Calibration (tests/golden/, oracle/make_golden.py).
 *
 * Reference lines restated:
 *   rotate / translate            src/nuscenes/utils/pcd.py:159-172
 *   [p,1] @ M.T (KITTI)           src/kitti/kitti_utils.py:212-249
 *   close-point removal           src/nuscenes/2d_to_3d.py:441-446
 *   view_points + normalise       src/nuscenes/utils/pcd.py:262-284
 *   bounds/depth test, floor      src/nuscenes/2d_to_3d.py:597-605
 *   mask lookup with !=0 quirk    src/nuscenes/2d_to_3d.py:608-617
 *   3x3 erosion (cv2.erode)       src/nuscenes/2d_to_3d.py:526-527
 *   medoid = argmin cdist.sum(0)  src/nuscenes/2d_to_3d.py:116-119
 *
 * torch-CPU facts this relies on (probed on torch 2.11, see DESIGN.md "Numerics"):
 *   - matmul(3x3|4x4, kxN) and matmul(Nx4, 4x3) are k-ordered FMA chains;
 *   - cdist(p=2) uses sqrt(clamp_min(x1_ @ x2_.T, 0)) with x1_=[-2a,|a|^2,1],
 *     x2_=[b,1,|b|^2] when M>25, else sqrt(fma(dz,dz,fma(dy,dy,dx*dx)));
 *   - sum(axis=0) of the MxM matrix is ATen's 4-level cascade per column for the
 *     first 32*floor(M/32) columns and a 4-way interleaved cascade for the rest;
 *   - torch's sqrt is MKL vsSqrt (<=1 ulp off IEEE on ~0.7% of inputs); this
 *     oracle uses the IEEE sqrtf, argmin agreement is checked in the tests.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { OP_END = 0, OP_T = 1, OP_R = 2, OP_A = 3, OP_WORDS = 16, MAX_CHAIN = 4 };

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static void apply_chain(const uint32_t *chain, float *x, float *y, float *z)
{
    for (int k = 0; k < MAX_CHAIN; k++) {
        const uint32_t *op = chain + k * OP_WORDS;
        const uint32_t kind = op[0];
        float m[12];
        for (int i = 0; i < 12; i++) m[i] = u2f(op[1 + i]);
        float a = *x, b = *y, c = *z;
        if (kind == OP_END) break;
        if (kind == OP_T) {
            *x = a + m[0]; *y = b + m[1]; *z = c + m[2];
        } else if (kind == OP_R) {
            float r[3];
            for (int i = 0; i < 3; i++) {
                float t = m[i * 3] * a;
                t = fmaf(m[i * 3 + 1], b, t);
                t = fmaf(m[i * 3 + 2], c, t);
                r[i] = t;
            }
            *x = r[0]; *y = r[1]; *z = r[2];
        } else { /* OP_A: [a b c 1] . row_i of the 3x4 */
            float r[3];
            for (int i = 0; i < 3; i++) {
                float t = a * m[i * 4];
                t = fmaf(b, m[i * 4 + 1], t);
                t = fmaf(c, m[i * 4 + 2], t);
                t = fmaf(1.0f, m[i * 4 + 3], t);
                r[i] = t;
            }
            *x = r[0]; *y = r[1]; *z = r[2];
        }
    }
}

/* One sweep: optional close-point removal, chain, append to SoA rows.
 * fourth: 0 none, 1 raw column 3, 2 ones.  Returns points kept. */
long oracle_aggregate_sweep(const float *raw, long n, int stride, const uint32_t *chain,
                            int fourth, int use_close, float close_thresh,
                            float *ox, float *oy, float *oz, float *ow)
{
    long k = 0;
    for (long i = 0; i < n; i++) {
        float x = raw[i * stride], y = raw[i * stride + 1], z = raw[i * stride + 2];
        if (use_close && fabsf(x) < close_thresh && fabsf(y) < close_thresh) continue;
        float w = fourth == 1 ? raw[i * stride + 3] : 1.0f;
        apply_chain(chain, &x, &y, &z);
        ox[k] = x; oy[k] = y; oz[k] = z;
        if (fourth) ow[k] = w;
        k++;
    }
    return k;
}

/* Project the cloud into one camera; pix[i] = fx | fy<<16 for points passing the
 * depth + strict image-bounds test, else -1.  viewpad = rows 0..2 of the 4x4. */
void oracle_project(const float *px, const float *py, const float *pz, long n,
                    const uint32_t *chain, const float *viewpad, float min_depth,
                    int W, int H, int32_t *pix)
{
    const float wlim = (float)(W - 1), hlim = (float)(H - 1);
    for (long i = 0; i < n; i++) {
        float x = px[i], y = py[i], z = pz[i];
        apply_chain(chain, &x, &y, &z);
        const float depth = z;
        float r[3];
        for (int j = 0; j < 3; j++) {
            float t = viewpad[j * 4] * x;
            t = fmaf(viewpad[j * 4 + 1], y, t);
            t = fmaf(viewpad[j * 4 + 2], z, t);
            t = fmaf(viewpad[j * 4 + 3], 1.0f, t);
            r[j] = t;
        }
        const float u = r[0] / r[2], v = r[1] / r[2];
        int32_t code = -1;
        if (depth > min_depth && u > 0.0f && u < wlim && v > 0.0f && v < hlim) {
            const int fx = (int)floorf(u), fy = (int)floorf(v);
            code = fx | (fy << 16);
        }
        pix[i] = code;
    }
}

/* Membership against one eroded (H,W) mask: member iff mask[fy][fx] && fx!=0 && fy!=0.
 * Writes ascending point indices, returns the count. */
long oracle_membership(const int32_t *pix, long n, const uint8_t *mask_hw, int W, int H,
                       int32_t *idx_out)
{
    long k = 0;
    (void)H;
    for (long i = 0; i < n; i++) {
        const int32_t c = pix[i];
        if (c < 0) continue;
        const int fx = c & 0xffff, fy = c >> 16;
        if (fx != 0 && fy != 0 && mask_hw[(long)fy * W + fx]) idx_out[k++] = (int32_t)i;
    }
    return k;
}

/* cv2.erode(mask, ones(3,3)) with the default border (+inf: outside pixels never lower the min). */
void oracle_erode3x3(const uint8_t *in, int W, int H, uint8_t *out)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint8_t m = 255;
            for (int dy = -1; dy <= 1; dy++) {
                const int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                for (int dx = -1; dx <= 1; dx++) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= W) continue;
                    const uint8_t v = in[(long)yy * W + xx];
                    if (v < m) m = v;
                }
            }
            out[(long)y * W + x] = m;
        }
}

/* COCO run lengths (0-run first) -> (H,W) uint8, row-major over the image. */
int oracle_rle_decode(const uint32_t *runs, long nruns, int W, int H, uint8_t *out)
{
    const long total = (long)W * H;
    long p = 0;
    uint8_t v = 0;
    for (long r = 0; r < nruns; r++) {
        const long len = runs[r];
        if (p + len > total) return -1;
        memset(out + p, v, (size_t)len);
        p += len;
        v = !v;
    }
    return p == total ? 0 : -1;
}

/* ---------------------------------------------------------------- medoid */
static long ceil_log2(long x)
{
    if (x <= 2) return 1;
    long l = 0, v = x - 1;
    while (v > 0) { l++; v >>= 1; }
    return l;
}

static inline float pair_dist(int mm, float xi, float yi, float zi, float ni,
                              float xj, float yj, float zj, float nj)
{
    if (mm) {
        float r = (-2.0f * xi) * xj;
        r = fmaf(-2.0f * yi, yj, r);
        r = fmaf(-2.0f * zi, zj, r);
        r = fmaf(ni, 1.0f, r);
        r = fmaf(1.0f, nj, r);
        r = r < 0.0f ? 0.0f : r;
        return sqrtf(r);
    }
    const float dx = xi - xj, dy = yi - yj, dz = zi - zj;
    return sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
}

/* One column sum with ATen's summation order. */
static float column_sum(const float *x, const float *y, const float *z, const float *nrm,
                        long m, long full, int mm, long j)
{
    const int nk = j < full ? 1 : 4;           /* interleave factor */
    const long n = j < full ? m : m / 4;       /* cascade length */
    long lp = ceil_log2(n) / 4;
    if (lp < 4) lp = 4;
    const long step = 1L << lp, mask = step - 1;
    float acc[4][4];
    memset(acc, 0, sizeof acc);
    long i = 0;
#define D(row) pair_dist(mm, x[row], y[row], z[row], nrm[row], x[j], y[j], z[j], nrm[j])
    while (i + step <= n) {
        for (long t = 0; t < step; t++, i++)
            for (int k = 0; k < nk; k++) acc[0][k] += D(i * nk + k);
        for (int l = 1; l < 4; l++) {
            for (int k = 0; k < nk; k++) { acc[l][k] += acc[l - 1][k]; acc[l - 1][k] = 0.0f; }
            if ((i & (mask << (l * lp))) != 0) break;
        }
    }
    for (; i < n; i++)
        for (int k = 0; k < nk; k++) acc[0][k] += D(i * nk + k);
    for (int l = 1; l < 4; l++)
        for (int k = 0; k < nk; k++) acc[0][k] += acc[l][k];
    if (nk == 4) {
        for (long r = n * 4; r < m; r++) acc[0][0] += D(r);
        for (int k = 1; k < 4; k++) acc[0][0] += acc[0][k];
    }
#undef D
    return acc[0][0];
}

typedef struct {
    const float *x, *y, *z, *nrm;
    float *colsum;
    long m, full;
    int mm, tid, nthr;
} medoid_job;

static void *medoid_worker(void *arg)
{
    const medoid_job *jb = (const medoid_job *)arg;
    for (long j = jb->tid; j < jb->m; j += jb->nthr)
        jb->colsum[j] = column_sum(jb->x, jb->y, jb->z, jb->nrm, jb->m, jb->full, jb->mm, j);
    return NULL;
}

/* argmin_j sum_i d(i,j); sums (optional) receives the M column sums.  Columns are
 * independent, so threads (nthreads<=0: 1) change wall time only, never a result. */
long oracle_medoid_mt(const float *x, const float *y, const float *z, long m, float *sums, int nthreads)
{
    if (m <= 0) return -1;
    const int mm = m > 25;
    float *nrm = (float *)malloc(sizeof(float) * (size_t)m);
    for (long i = 0; i < m; i++) nrm[i] = (x[i] * x[i] + y[i] * y[i]) + z[i] * z[i];
    const long full = m >= 8 ? (m / 32) * 32 : (m / 4) * 4;
    float *colsum = sums ? sums : (float *)malloc(sizeof(float) * (size_t)m);
    int nthr = nthreads < 1 || m < 512 ? 1 : (nthreads > 64 ? 64 : nthreads);
    medoid_job jobs[64];
    pthread_t th[64];
    for (int t = 0; t < nthr; t++) {
        medoid_job jb = { x, y, z, nrm, colsum, m, full, mm, t, nthr };
        jobs[t] = jb;
    }
    if (nthr == 1) medoid_worker(&jobs[0]);
    else {
        for (int t = 0; t < nthr; t++) pthread_create(&th[t], NULL, medoid_worker, &jobs[t]);
        for (int t = 0; t < nthr; t++) pthread_join(th[t], NULL);
    }
    long best = 0;
    for (long j = 1; j < m; j++)
        if (colsum[j] < colsum[best]) best = j;   /* first minimum, like torch.argmin */
    if (!sums) free(colsum);
    free(nrm);
    return best;
}

long oracle_medoid(const float *x, const float *y, const float *z, long m, float *sums)
{
    return oracle_medoid_mt(x, y, z, m, sums, 1);
}

/* The matmul formula's squared distance before the clamp, r(i, j) with row i of x1 and row j of x2 (the same
 * chain as pair_dist above), as an m x m row-major matrix: tests of the symmetry conditions the CUDA screen
 * relies on (cm3d_b200/csrc/medoid.cu: k_medoid_classify). */
void oracle_sqdist(const float *x, const float *y, const float *z, long m, float *out)
{
    float *nrm = (float *)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
    for (long i = 0; i < m; i++) nrm[i] = (x[i] * x[i] + y[i] * y[i]) + z[i] * z[i];
    for (long i = 0; i < m; i++)
        for (long j = 0; j < m; j++) {
            float r = (-2.0f * x[i]) * x[j];
            r = fmaf(-2.0f * y[i], y[j], r);
            r = fmaf(-2.0f * z[i], z[j], r);
            r = fmaf(nrm[i], 1.0f, r);
            r = fmaf(1.0f, nrm[j], r);
            out[i * m + j] = r;
        }
    free(nrm);
}
