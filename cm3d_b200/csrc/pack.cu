// Host-side packer: FrameSpecs (flattened by cm3d_b200/batch.py: pack_frames_native) -> the raw / meta /
// mask buffers of a PackedBatch, byte for byte what the Python packer (batch.py: pack_frames) writes.
// Pure host C++: no CUDA call, no allocation of the outputs (the caller hands in pinned buffers sized by
// cm3d_pack_plan), no Python object touched - so ctypes releases the GIL around it and several batches
// are packed at once by Lifter.lift_frame_stream's worker threads.  The descriptor tables are the ones
// of include/cm3d_b200.h; the conservative cull planes follow batch.py: cull_planes (fp64).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "../../include/cm3d_b200.h"

namespace {

constexpr int kSW = CM3D_SW_WORDS, kFR = CM3D_FR_WORDS, kVC = CM3D_VC_WORDS, kIN = CM3D_IN_WORDS;

inline int32_t f32_bits(float x)
{
    int32_t b;
    memcpy(&b, &x, 4);
    return b;
}
inline int64_t align4(int64_t n) { return (n + 3) & ~(int64_t)3; }

struct Key {
    int cam, W, H;
    bool operator<(const Key &o) const { return cam != o.cam ? cam < o.cam : (W != o.W ? W < o.W : H < o.H); }
    bool operator==(const Key &o) const { return cam == o.cam && W == o.W && H == o.H; }
};

// One op of a chain: kind (CM3D_OP_T / _R / _A) and its fp32 matrix (3, 9 or 12 values).
struct OpRef {
    int kind;
    const float *m;
};

int op_size(int kind) { return kind == CM3D_OP_T ? 3 : (kind == CM3D_OP_R ? 9 : 12); }

// nuScenes .bin rows are (x, y, z, intensity, ring index): nothing downstream reads the fifth column
// (src/nuscenes/utils/pcd.py:246-257 keeps four), so it does not travel to the GPU.
// only the columns the kernels read travel: x, y, z (+ column 3 when it becomes row 3 of the cloud, fourth == 1)
int64_t packed_stride(int fourth) { return fourth == 1 ? 4 : 3; }

// Copy the first `cols` (3 or 4) floats of n rows of `stride_in` floats into a dense (n, cols) block.
// The destination is the pinned staging buffer the DMA engine reads next and the packer never reads back:
// non-temporal 16-byte stores keep it out of the caches and skip the read-for-ownership of every line, a
// quarter of the host memory traffic of a packed frame - and host memory bandwidth is what bounds the
// FrameSpec stream once eight packer threads and the H2D copies run at the same time.
// dst must be 16-byte aligned (sweeps start at multiples of four floats of a 4 KiB-aligned buffer).
void pack_rows(float *dst, const float *src, int64_t n, int stride_in, int cols)
{
#if defined(__SSE2__)
    if (((uintptr_t)dst & 15) == 0) {
        int64_t k = 0;
        if (cols == 4) {
            for (; k < n; ++k) _mm_stream_ps(dst + 4 * k, _mm_loadu_ps(src + (int64_t)stride_in * k));
        } else if (stride_in == 5) {            // (N,5) rows -> x, y, z: five loads, four shuffles, three stores per 4 rows
            for (; k + 4 <= n; k += 4) {
                const float *p = src + 5 * k;
                const __m128 v0 = _mm_loadu_ps(p), v1 = _mm_loadu_ps(p + 4), v2 = _mm_loadu_ps(p + 8),
                             v3 = _mm_loadu_ps(p + 12), v4 = _mm_loadu_ps(p + 16);
                const __m128 t = _mm_shuffle_ps(v0, v1, _MM_SHUFFLE(1, 1, 2, 2));                 // f2 f2 f5 f5
                _mm_stream_ps(dst + 3 * k, _mm_shuffle_ps(v0, t, _MM_SHUFFLE(2, 0, 1, 0)));       // f0 f1 f2 f5
                _mm_stream_ps(dst + 3 * k + 4, _mm_shuffle_ps(v1, v2, _MM_SHUFFLE(3, 2, 3, 2)));  // f6 f7 f10 f11
                _mm_stream_ps(dst + 3 * k + 8, _mm_shuffle_ps(v3, v4, _MM_SHUFFLE(1, 0, 3, 0)));  // f12 f15 f16 f17
            }
        } else if (stride_in == 3) {            // (N,3) rows as they are (Waymo's range-image points)
            const int64_t m = 3 * n;
            int64_t j = 0;
            for (; j + 4 <= m; j += 4) _mm_stream_ps(dst + j, _mm_loadu_ps(src + j));
            for (; j < m; ++j) dst[j] = src[j];
            k = n;
        } else if (stride_in == 4) {            // (N,4) rows -> x, y, z
            for (; k + 4 <= n; k += 4) {
                const float *p = src + 4 * k;
                const __m128 a = _mm_loadu_ps(p), b = _mm_loadu_ps(p + 4), c = _mm_loadu_ps(p + 8), d = _mm_loadu_ps(p + 12);
                const __m128 t = _mm_shuffle_ps(a, b, _MM_SHUFFLE(0, 0, 2, 2));                   // a2 a2 b0 b0
                const __m128 u = _mm_shuffle_ps(c, d, _MM_SHUFFLE(0, 0, 2, 2));                   // c2 c2 d0 d0
                _mm_stream_ps(dst + 3 * k, _mm_shuffle_ps(a, t, _MM_SHUFFLE(2, 0, 1, 0)));        // a0 a1 a2 b0
                _mm_stream_ps(dst + 3 * k + 4, _mm_shuffle_ps(b, c, _MM_SHUFFLE(1, 0, 2, 1)));    // b1 b2 c0 c1
                _mm_stream_ps(dst + 3 * k + 8, _mm_shuffle_ps(u, d, _MM_SHUFFLE(2, 1, 2, 0)));    // c2 d0 d1 d2
            }
        } else {
            for (; k + 4 <= n; k += 4) {
                const float *a = src + (int64_t)stride_in * k, *b = a + stride_in, *c = b + stride_in, *d = c + stride_in;
                _mm_stream_ps(dst + 3 * k, _mm_setr_ps(a[0], a[1], a[2], b[0]));
                _mm_stream_ps(dst + 3 * k + 4, _mm_setr_ps(b[1], b[2], c[0], c[1]));
                _mm_stream_ps(dst + 3 * k + 8, _mm_setr_ps(c[2], d[0], d[1], d[2]));
            }
        }
        if (cols != 4) {
            for (; k < n; ++k) memcpy(dst + 3 * k, src + (int64_t)stride_in * k, 12);
        }
        _mm_sfence();
        return;
    }
#endif
    for (int64_t k = 0; k < n; ++k) memcpy(dst + cols * k, src + (int64_t)stride_in * k, (size_t)cols * 4);
}

void encode_chain(const OpRef *ops, int n, uint32_t *out)
{
    memset(out, 0, CM3D_CHAIN_WORDS * 4);
    for (int k = 0; k < n; ++k) {
        out[k * CM3D_OP_WORDS] = (uint32_t)ops[k].kind;
        memcpy(out + k * CM3D_OP_WORDS + 1, ops[k].m, op_size(ops[k].kind) * 4);
    }
}

int chain_sig(const OpRef *ops, int n)
{
    int s = 0;
    for (int i = 0; i < n; ++i) s += ops[i].kind << (2 * i);
    return s;
}

// batch.py: _compose + cull_planes.  planes: 5 x 4 fp32.  Returns the flags word.
int cull_planes(const OpRef *ops, int n_ops, const float *K32, int W, int H, float min_dist32, const float *tref32,
                float *planes)
{
    const float off[4] = {0.f, 0.f, 0.f, 1.f};
    double K[3][3];
    for (int i = 0; i < 9; ++i) K[i / 3][i % 3] = (double)K32[i];
    const bool bottom_ok = K[2][0] == 0 && K[2][1] == 0 && K[2][2] == 1;
    const bool simple = K[0][1] == 0 && K[1][0] == 0 && K[0][0] != 0 && K[1][1] != 0 && bottom_ok;
    const int flags = simple ? 1 : 0;
    double M[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, c[3] = {0, 0, 0}, tau = 0.0;
    bool ok = true;
    for (int k = 0; k < n_ops; ++k) {
        const float *m = ops[k].m;
        if (ops[k].kind == CM3D_OP_T) {
            for (int i = 0; i < 3; ++i) c[i] += (double)m[i];
            if (k) tau += fabs((double)m[0]) + fabs((double)m[1]) + fabs((double)m[2]);
        } else {
            const int ld = ops[k].kind == CM3D_OP_R ? 3 : 4;
            double L[3][3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) L[i][j] = (double)m[i * ld + j];
            double dev = 0.0;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    double s = 0.0;
                    for (int q = 0; q < 3; ++q) s += L[i][q] * L[j][q];
                    dev = fmax(dev, fabs(s - (i == j ? 1.0 : 0.0)));
                }
            ok = ok && (dev < 1e-3);
            double M2[3][3], c2[3];
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) {
                    double s = 0.0;
                    for (int q = 0; q < 3; ++q) s += L[i][q] * M[q][j];
                    M2[i][j] = s;
                }
                double s = 0.0;
                for (int q = 0; q < 3; ++q) s += L[i][q] * c[q];
                c2[i] = s;
            }
            memcpy(M, M2, sizeof(M));
            memcpy(c, c2, sizeof(c));
            if (ops[k].kind == CM3D_OP_A) {
                for (int i = 0; i < 3; ++i) c[i] += (double)m[i * 4 + 3];
                tau += fabs((double)m[3]) + fabs((double)m[7]) + fabs((double)m[11]);
            }
        }
    }
    bool finite = true;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) finite = finite && isfinite(M[i][j]);
    auto switch_off = [&]() {
        for (int p = 0; p < 5; ++p) memcpy(planes + 4 * p, off, 16);
        return flags;
    };
    if (!ok || !bottom_ok || !finite) return switch_off();
    double tref[3] = {(double)tref32[0], (double)tref32[1], (double)tref32[2]};
    if (n_ops && ops[0].kind == CM3D_OP_T)
        tau += fabs((double)ops[0].m[0] - tref[0]) + fabs((double)ops[0].m[1] - tref[1]) + fabs((double)ops[0].m[2] - tref[2]);   // residual t_c - tref
    else
        tau += fabs(tref[0]) + fabs(tref[1]) + fabs(tref[2]);
    double c2[3];
    for (int i = 0; i < 3; ++i) c2[i] = c[i] - (M[i][0] * tref[0] + M[i][1] * tref[1] + M[i][2] * tref[2]);
    const double abc[5][3] = {{0.0, 0.0, 1.0},
                              {K[0][0], K[0][1], K[0][2]},
                              {-K[0][0], -K[0][1], (double)W - K[0][2]},
                              {K[1][0], K[1][1], K[1][2]},
                              {-K[1][0], -K[1][1], (double)H - K[1][2]}};
    const double d0[5] = {-(double)min_dist32, 0.0, 0.0, 0.0, 0.0};
    double out[5][4], dmax = 0.0;
    for (int p = 0; p < 5; ++p) {
        const double kappa = 1.75 * (fabs(abc[p][0]) + fabs(abc[p][1]) + fabs(abc[p][2]));
        double nsum = 0.0;
        for (int j = 0; j < 3; ++j) {
            out[p][j] = (abc[p][0] * M[0][j] + abc[p][1] * M[1][j] + abc[p][2] * M[2][j]) / kappa;
            nsum += fabs(out[p][j]);
        }
        out[p][3] = (abc[p][0] * c2[0] + abc[p][1] * c2[1] + abc[p][2] * c2[2] + d0[p]) / kappa;
        if (!(nsum <= 1.0) || !isfinite(out[p][3])) return switch_off();
        dmax = fmax(dmax, fabs(out[p][3]));
    }
    const double mg0 = ldexp(1.0, -17) * tau + ldexp(1.0, -20) * dmax + 1e-30;
    for (int p = 0; p < 5; ++p) {
        for (int j = 0; j < 3; ++j) planes[4 * p + j] = (float)out[p][j];
        planes[4 * p + 3] = nextafterf((float)(out[p][3] + mg0), INFINITY);      // rounding of d never tightens a plane
    }
    return flags;
}

}  // namespace

extern "C" {

int cm3d_pack_input_size(void) { return (int)sizeof(cm3d_pack_input); }

// plan[]: 0 n_tiles, 1 n_vcams, 2 n_chains, 3 raw floats (incl. 4 of padding), 4 meta words, 5 mask bytes,
// 6..12 word offsets of tile_sweep, sweep_desc, frame_desc, vcam_desc, cam_inst_list, inst_desc, chains in meta,
// 13 max_inst_per_frame, 14 n_raw_points.  Returns 0, or CM3D_ELIMIT / CM3D_EINVAL.
int cm3d_pack_plan(const cm3d_pack_input *in, int64_t *plan)
{
    if (!in || !plan || in->n_frames <= 0) return CM3D_EINVAL;
    int64_t n_tiles = 0, raw = 0, n_vcams = 0, n_pts = 0;
    int si = 0, ii = 0, max_inst = 0;
    std::vector<Key> keys;
    for (int f = 0; f < in->n_frames; ++f) {
        for (int s = 0; s < in->fr_n_sweeps[f]; ++s, ++si) {
            const int64_t n = in->sw_npts[si];
            n_pts += n;
            if (n == 0) continue;
            n_tiles += (n + CM3D_TILE - 1) / CM3D_TILE;
            raw += align4(n * packed_stride(in->fr_fourth[f]));
        }
        const int I = in->fr_n_inst[f];
        if (I > CM3D_MAX_INST) return CM3D_ELIMIT;
        max_inst = std::max(max_inst, I);
        keys.clear();
        for (int i = 0; i < I; ++i) keys.push_back(Key{in->in_cam[ii + i], in->in_W[ii + i], in->in_H[ii + i]});
        std::sort(keys.begin(), keys.end());
        const int nk = (int)(std::unique(keys.begin(), keys.end()) - keys.begin());
        if (nk > CM3D_MAX_VCAMS) return CM3D_ELIMIT;
        n_vcams += nk;
        ii += I;
    }
    const int64_t n_chains = in->n_sweeps + n_vcams;
    int64_t pos = 0;
    const int64_t sizes[7] = {std::max<int64_t>(n_tiles, 1), std::max<int64_t>(in->n_sweeps, 1) * kSW, (int64_t)in->n_frames * kFR,
                              std::max<int64_t>(n_vcams, 1) * kVC, std::max<int64_t>(in->n_inst, 1),
                              std::max<int64_t>(in->n_inst, 1) * kIN, std::max<int64_t>(n_chains, 1) * CM3D_CHAIN_WORDS};
    for (int k = 0; k < 7; ++k) {
        plan[6 + k] = pos;
        pos += align4(sizes[k]);
    }
    plan[0] = n_tiles; plan[1] = n_vcams; plan[2] = n_chains; plan[3] = raw + 4; plan[4] = pos;
    plan[5] = in->in_counts_off[in->n_inst];
    plan[13] = max_inst; plan[14] = n_pts;
    return CM3D_OK;
}

// Fills raw (plan[3] floats), meta (plan[4] words, zeroed here), mask (plan[5] bytes), mask_off (n_inst + 1).
// vcam_keys: plan[1] x 4 ints (frame, camera, W, H).  out[]: 0 cnt_total, 1 bits_words, 2 max_words,
// 3 grid_words, 4 max_cells, 5 max_runs.
int cm3d_pack_fill(const cm3d_pack_input *in, const int64_t *plan, float *raw, int32_t *meta, uint8_t *mask,
                   int64_t *mask_off, int32_t *vcam_keys, int64_t *out)
{
    if (!in || !plan || !raw || !meta || !mask_off || !vcam_keys || !out) return CM3D_EINVAL;
    memset(meta, 0, (size_t)plan[4] * 4);
    int32_t *tile_sweep = meta + plan[6], *sweep_desc = meta + plan[7], *frame_desc = meta + plan[8];
    int32_t *vcam_desc = meta + plan[9], *cam_inst_list = meta + plan[10], *inst_desc = meta + plan[11];
    uint32_t *chains = reinterpret_cast<uint32_t *>(meta + plan[12]);

    int64_t ro = 0, cnt_total = 0, bits_words = 0, max_words = 0, grid_words = 0, max_cells = 0, max_runs = 0;
    int si = 0, ti = 0, ii = 0, ci = 0, n_chain = 0, n_vcam = 0;
    std::vector<OpRef> ops;
    std::vector<Key> keys;
    std::vector<int> key_of;
    for (int f = 0; f < in->n_frames; ++f) {
        const int I = in->fr_n_inst[f];
        const int t_begin = ti;
        for (int s = 0; s < in->fr_n_sweeps[f]; ++s, ++si) {
            const int64_t n = in->sw_npts[si], stride_in = in->sw_stride[si], stride = packed_stride(in->fr_fourth[f]);
            const int64_t o = ro;
            int nt = 0;
            if (n > 0) {
                nt = (int)((n + CM3D_TILE - 1) / CM3D_TILE);
                const float *src = reinterpret_cast<const float *>(in->sw_ptr[si]);
                pack_rows(raw + o, src, n, (int)stride_in, (int)stride);
                for (int64_t k = n * stride; k < align4(n * stride); ++k) raw[o + k] = 0.0f;
                ro += align4(n * stride);
            }
            int32_t *sd = sweep_desc + (size_t)si * kSW;
            sd[0] = (int32_t)(uint32_t)(o & 0xffffffff); sd[1] = (int32_t)(uint32_t)((uint64_t)o >> 32);
            sd[2] = (int32_t)n; sd[3] = (int32_t)stride; sd[4] = f; sd[5] = ti; sd[6] = n_chain; sd[7] = in->fr_fourth[f];
            ops.clear();
            for (int k = in->op_begin[si]; k < in->op_begin[si + 1]; ++k)
                ops.push_back(OpRef{in->op_kind[k], reinterpret_cast<const float *>(in->op_ptr[k])});
            if ((int)ops.size() > CM3D_MAX_CHAIN) return CM3D_EINVAL;
            encode_chain(ops.data(), (int)ops.size(), chains + (size_t)n_chain * CM3D_CHAIN_WORDS);
            ++n_chain;
            for (int k = 0; k < nt; ++k) tile_sweep[ti + k] = si;
            ti += nt;
        }
        const int ntf = ti - t_begin;
        // vcams: unique (camera, W, H) among the frame's instances, sorted
        keys.clear();
        for (int i = 0; i < I; ++i) keys.push_back(Key{in->in_cam[ii + i], in->in_W[ii + i], in->in_H[ii + i]});
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        const int nk = (int)keys.size();
        key_of.assign(I, 0);
        for (int i = 0; i < I; ++i) {
            const Key k{in->in_cam[ii + i], in->in_W[ii + i], in->in_H[ii + i]};
            key_of[i] = (int)(std::lower_bound(keys.begin(), keys.end(), k) - keys.begin());
        }
        const int v_begin = n_vcam;
        float tref[3] = {0.f, 0.f, 0.f};
        int sig = -1;
        if (nk) {
            const int cam0 = in->n_sweeps + ci + keys[0].cam;
            const int b = in->op_begin[cam0], e = in->op_begin[cam0 + 1];
            if (e > b && in->op_kind[b] == CM3D_OP_T) memcpy(tref, reinterpret_cast<const float *>(in->op_ptr[b]), 12);
        }
        int lb = 0;
        bool sig_same = true;
        for (int v = 0; v < nk; ++v) {
            const int cam = in->n_sweeps + ci + keys[v].cam;
            ops.clear();
            for (int k = in->op_begin[cam]; k < in->op_begin[cam + 1]; ++k)
                ops.push_back(OpRef{in->op_kind[k], reinterpret_cast<const float *>(in->op_ptr[k])});
            if ((int)ops.size() > CM3D_MAX_CHAIN) return CM3D_EINVAL;
            const int sg = chain_sig(ops.data(), (int)ops.size());
            if (v == 0) sig = sg; else sig_same = sig_same && sg == sig;
            const float *K = reinterpret_cast<const float *>(in->cam_K[ci + keys[v].cam]);
            int32_t *row = vcam_desc + (size_t)n_vcam * kVC;
            float planes[20];
            row[CM3D_VC_FLAGS] = cull_planes(ops.data(), (int)ops.size(), K, keys[v].W, keys[v].H, in->fr_min_dist[f], tref, planes);
            memcpy(row + CM3D_VC_PLANES, planes, 80);
            row[CM3D_VC_CHAIN] = n_chain;
            encode_chain(ops.data(), (int)ops.size(), chains + (size_t)n_chain * CM3D_CHAIN_WORDS);
            ++n_chain;
            float vp[12] = {K[0], K[1], K[2], 0.f, K[3], K[4], K[5], 0.f, K[6], K[7], K[8], 0.f};
            memcpy(row + CM3D_VC_VIEWPAD, vp, 48);
            int n_mem = 0;
            for (int i = 0; i < I; ++i)
                if (key_of[i] == v) cam_inst_list[ii + lb + n_mem++] = i;
            const int gnx = (keys[v].W + CM3D_CELL - 1) / CM3D_CELL, gny = (keys[v].H + CM3D_CELL - 1) / CM3D_CELL;
            row[CM3D_VC_W] = keys[v].W; row[CM3D_VC_H] = keys[v].H; row[CM3D_VC_LIST_BEGIN] = lb;
            row[CM3D_VC_LIST_COUNT] = n_mem; row[CM3D_VC_FRAME] = f; row[CM3D_VC_GRID_OFF] = (int32_t)grid_words;
            row[CM3D_VC_GRID_NX] = gnx;
            grid_words += (int64_t)gnx * gny * ((n_mem + 31) / 32);
            max_cells = std::max<int64_t>(max_cells, (int64_t)gnx * gny);
            lb += n_mem;
            int32_t *vk = vcam_keys + (size_t)n_vcam * 4;
            vk[0] = f; vk[1] = keys[v].cam; vk[2] = keys[v].W; vk[3] = keys[v].H;
            ++n_vcam;
        }
        int32_t *fd = frame_desc + (size_t)f * kFR;
        fd[CM3D_FR_TILE_BEGIN] = t_begin; fd[CM3D_FR_TILE_END] = ti; fd[CM3D_FR_VCAM_BEGIN] = v_begin; fd[CM3D_FR_NVCAMS] = nk;
        fd[CM3D_FR_INST_BEGIN] = ii; fd[CM3D_FR_NINST] = I;
        fd[CM3D_FR_CLOSE_BITS] = f32_bits(in->fr_use_close[f] ? in->fr_close[f] : 0.0f);
        fd[CM3D_FR_USE_CLOSE] = in->fr_use_close[f] ? 1 : 0;
        fd[CM3D_FR_MIN_DEPTH_BITS] = f32_bits(in->fr_min_dist[f]);
        fd[CM3D_FR_CNT_OFF] = (int32_t)cnt_total;
        fd[CM3D_FR_MIN_MEDOID_PTS] = in->fr_min_pts[f];
        fd[CM3D_FR_LIST_BEGIN] = ii;
        fd[CM3D_FR_TREF] = f32_bits(tref[0]); fd[CM3D_FR_TREF + 1] = f32_bits(tref[1]); fd[CM3D_FR_TREF + 2] = f32_bits(tref[2]);
        fd[CM3D_FR_CHAIN_SIG] = (nk && sig_same) ? sig : -1;
        fd[CM3D_FR_FLOOR_BITS] = f32_bits(in->fr_use_floor[f] ? in->fr_floor[f] : 0.0f);
        fd[CM3D_FR_USE_FLOOR] = in->fr_use_floor[f] ? 1 : 0;
        cnt_total += (int64_t)ntf * I;
        for (int i = 0; i < I; ++i) {
            const int W = in->in_W[ii + i], H = in->in_H[ii + i];
            const int pitch = (W + 31) / 32;
            const int64_t words = (int64_t)pitch * H;
            int32_t *d = inst_desc + (size_t)(ii + i) * kIN;
            d[CM3D_IN_BITS_LO] = (int32_t)(uint32_t)(bits_words & 0xffffffff);
            d[CM3D_IN_BITS_HI] = (int32_t)(uint32_t)((uint64_t)bits_words >> 32);
            d[CM3D_IN_W] = W; d[CM3D_IN_H] = H; d[CM3D_IN_PITCH] = pitch; d[CM3D_IN_VCAM] = v_begin + key_of[i];
            d[CM3D_IN_FRAME] = f; d[CM3D_IN_LOCAL] = i;
            bits_words += words;
            max_words = std::max(max_words, words);
            max_runs = std::max<int64_t>(max_runs, in->in_counts_off[ii + i + 1] - in->in_counts_off[ii + i]);
        }
        ii += I;
        ci += in->fr_n_cams[f];
    }
    for (int k = 0; k < 4; ++k) raw[ro + k] = 0.0f;
    if (plan[5] > 0) memcpy(mask, in->counts, (size_t)plan[5]);
    for (int i = 0; i <= in->n_inst; ++i) mask_off[i] = in->in_counts_off[i];
    out[0] = cnt_total; out[1] = bits_words; out[2] = max_words; out[3] = grid_words; out[4] = max_cells; out[5] = max_runs;
    return CM3D_OK;
}

}  // extern "C"
