// Instance masks on the GPU: COCO run lengths or dense uint8 -> bit planes ->
// 3x3-eroded bit planes + bounding boxes.
//
// Replaces, for a whole batch of frames at once:
//   pycocotools.mask.decode + transpose   src/nuscenes/2d_to_3d.py:425-428, waymo:520-521
//   cv2.erode(mask, ones((3,3)))          src/nuscenes/2d_to_3d.py:526-527, kitti:1149-1150, waymo:531-532
//   .astype(bool) -> (W,H) torch view     src/nuscenes/2d_to_3d.py:543-544
// A bit plane stores pixel (x, y) of the (H,W) image at bit (x & 31) of word
// y*pitch + (x >> 5); the reference's `mask[fx, fy]` on its (W,H) view is that bit.
// HBM-bound byte/bit work: coalesced 32-byte reads per thread, one word out.
#include "common.cuh"

namespace cm3d {

// ------------------------------------------------------------------ dense uint8 -> bits
__device__ __forceinline__ uint32_t nz4(uint32_t v)
{
    // one bit per non-zero byte of v, in byte order
    return ((__vcmpne4(v, 0u) & 0x08040201u) * 0x01010101u) >> 24;
}

__global__ void __launch_bounds__(256)
k_pack_dense(const uint8_t *__restrict__ masks, const int64_t *__restrict__ src_off,
             const int32_t *__restrict__ inst_desc, uint32_t *__restrict__ bits)
{
    const int i = blockIdx.y;
    const int32_t *d = inst_desc + i * CM3D_IN_WORDS;
    const int W = d[CM3D_IN_W], H = d[CM3D_IN_H], pitch = d[CM3D_IN_PITCH];
    const int wd = blockIdx.x * blockDim.x + threadIdx.x;
    if (wd >= H * pitch) return;
    const int y = wd / pitch, xw = wd - y * pitch, x0 = xw * 32;
    const int64_t so = src_off[i];
    const uint8_t *row = masks + so + (int64_t)y * W;
    uint32_t out = 0;
    if (((W & 31) == 0) && ((so & 15) == 0)) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(row + x0));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(row + x0 + 16));
        out = nz4(a.x) | (nz4(a.y) << 4) | (nz4(a.z) << 8) | (nz4(a.w) << 12) |
              (nz4(b.x) << 16) | (nz4(b.y) << 20) | (nz4(b.z) << 24) | (nz4(b.w) << 28);
    } else {
        const int nb = min(32, W - x0);
        for (int b = 0; b < nb; ++b) out |= (uint32_t)(row[x0 + b] != 0) << b;
    }
    bits[join64(d[CM3D_IN_BITS_LO], d[CM3D_IN_BITS_HI]) + wd] = out;
}

// ------------------------------------------------------------------ COCO runs -> bits
// ------------------------------------------------------------------ COCO `counts` string -> runs
// pycocotools' compressed run lengths (maskApi.c rleFrString, restated in cm3d_b200/rle.py): every
// run is 1..7 chars of 5 payload bits (char - 48; bit 0x20 = "more", sign-extended from bit 0x10
// of the last char); from the 4th run on the value is a delta against the run two places back.
// One block per instance.  Pass 1: every char that starts a token decodes it (tokens are found
// with a ballot prefix over "previous char ended a token").  Pass 2: the deltas are two
// interleaved running sums (odd tokens from token 1, even tokens from token 2), done as a block
// scan over token pairs.  Instance i's runs land at runs[byte_off[i] ..), zero-padded up to
// byte_off[i+1] (a run never needs less than one char), so run_off == byte_off downstream.
__global__ void __launch_bounds__(256)
k_rle_decode(const uint8_t *__restrict__ bytes, const int64_t *__restrict__ byte_off, uint32_t *__restrict__ runs)
{
    __shared__ uint32_t s_a[8], s_b[8];
    __shared__ uint32_t s_ca, s_cb;
    const int i = blockIdx.x;
    const int64_t b0 = byte_off[i];
    const int nb = (int)(byte_off[i + 1] - b0);
    const uint8_t *src = bytes + b0;
    uint32_t *dst = runs + b0;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_ca = 0; s_cb = 0; }
    __syncthreads();
    for (int base = 0; base < nb; base += blockDim.x) {
        const int p = base + threadIdx.x;
        const bool start = p < nb && (p == 0 || !((src[p - 1] - 48) & 0x20));
        const unsigned bal = __ballot_sync(0xffffffffu, start);
        if (lane == 0) s_a[warp] = __popc(bal);
        __syncthreads();
        uint32_t t = s_ca + __popc(bal & lanemask_lt());
        for (unsigned w = 0; w < warp; ++w) t += s_a[w];
        if (start) {
            long long x = 0;
            int k = 0, q = p;
            bool more = true;
            while (more && q < nb) {
                const int c = (int)src[q] - 48;
                x |= (long long)(c & 0x1f) << (5 * k);
                more = (c & 0x20) != 0;
                ++q; ++k;
                if (!more && (c & 0x10)) x |= -1ll << (5 * k);
            }
            dst[t] = (uint32_t)x;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; ++w) tot += s_a[w];
            s_ca += tot;
        }
        __syncthreads();
    }
    const int ntok = (int)s_ca;
    for (int t = ntok + threadIdx.x; t < nb; t += blockDim.x) dst[t] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) s_ca = 0;
    __syncthreads();
    for (int base = 0; 2 * base < ntok; base += blockDim.x) {
        const int te = 2 * (base + threadIdx.x), to = te + 1;
        const bool he = te < ntok && te >= 2, ho = to < ntok;
        uint32_t e = he ? dst[te] : 0u, o = ho ? dst[to] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ue = __shfl_up_sync(0xffffffffu, e, d), uo = __shfl_up_sync(0xffffffffu, o, d);
            if (lane >= (unsigned)d) { e += ue; o += uo; }
        }
        if (lane == 31) { s_a[warp] = e; s_b[warp] = o; }
        __syncthreads();
        uint32_t ce = s_ca, co = s_cb;
        for (unsigned w = 0; w < warp; ++w) { ce += s_a[w]; co += s_b[w]; }
        if (he) dst[te] = ce + e;
        if (ho) dst[to] = co + o;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) { s_ca = ce + e; s_cb = co + o; }
        __syncthreads();
    }
}

// run_start[r] = first flat pixel (y*W + x) of run r; one block per instance.
__global__ void __launch_bounds__(256)
k_rle_prefix(const uint32_t *__restrict__ runs, const int64_t *__restrict__ run_off,
             const int32_t *__restrict__ inst_desc, uint32_t *__restrict__ run_start,
             int32_t *__restrict__ row_range, int32_t *__restrict__ errflags)
{
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_carry;
    __shared__ unsigned s_lo, s_hi;           // first / one-past-last set pixel (flat index) of the mask
    if (threadIdx.x == 0) { s_lo = 0xffffffffu; s_hi = 0u; }
    const int i = blockIdx.x;
    const int64_t r0 = run_off[i], r1 = run_off[i + 1];
    const int32_t *d = inst_desc + i * CM3D_IN_WORDS;
    const uint32_t total = (uint32_t)d[CM3D_IN_W] * (uint32_t)d[CM3D_IN_H];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = r0; base < r1; base += blockDim.x) {
        const int64_t r = base + threadIdx.x;
        const uint32_t v = r < r1 ? runs[r] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        uint32_t woff = s_carry;
        for (unsigned w = 0; w < warp; ++w) woff += s_w[w];
        if (r < r1) run_start[r] = woff + inc - v;
        if (r < r1 && v != 0u && ((r - r0) & 1)) {       // a non-empty 1-run
            atomicMin(&s_lo, woff + inc - v);
            atomicMax(&s_hi, min(woff + inc, total));
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (s_carry != total) atomicExch(&errflags[CM3D_ERR_RLE_SIZE], i + 1);
        const uint32_t W = (uint32_t)d[CM3D_IN_W];
        // rows that hold a set pixel: erosion only needs to look at those (everything else stays 0)
        row_range[2 * i] = s_hi ? (int32_t)(s_lo / W) : 0x7fffffff;
        row_range[2 * i + 1] = s_hi ? (int32_t)((s_hi - 1) / W) : -1;
    }
}

// One warp per 1-run: OR its pixels into the (zeroed) bit plane, row by row.
__global__ void __launch_bounds__(256)
k_rle_fill(const uint32_t *__restrict__ runs, const int64_t *__restrict__ run_off,
           const uint32_t *__restrict__ run_start, const int32_t *__restrict__ inst_desc,
           uint32_t *__restrict__ bits)
{
    const int i = blockIdx.y;
    const int64_t r0 = run_off[i], r1 = run_off[i + 1];
    const int64_t r = r0 + 2 * (int64_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) + 1;
    if (r >= r1) return;
    const uint32_t len = runs[r];
    if (len == 0) return;
    const int32_t *d = inst_desc + i * CM3D_IN_WORDS;
    const uint32_t W = (uint32_t)d[CM3D_IN_W], H = (uint32_t)d[CM3D_IN_H];
    const int pitch = d[CM3D_IN_PITCH];
    uint32_t *plane = bits + join64(d[CM3D_IN_BITS_LO], d[CM3D_IN_BITS_HI]);
    const uint32_t s = run_start[r];
    uint32_t e = s + len;                       // exclusive
    if (e > W * H) e = W * H;                   // malformed input is flagged by k_rle_prefix
    if (s >= e) return;
    const uint32_t y0 = s / W, y1 = (e - 1) / W;
    const unsigned lane = lane_id();
    for (uint32_t y = y0; y <= y1; ++y) {
        const uint32_t xa = (y == y0) ? s - y0 * W : 0u;
        const uint32_t xb = (y == y1) ? (e - 1) - y1 * W + 1 : W;   // exclusive
        const uint32_t wa = xa >> 5, wb = (xb - 1) >> 5;
        for (uint32_t w = wa + lane; w <= wb; w += 32) {
            const uint32_t lo = max(xa, w << 5) - (w << 5);
            const uint32_t hi = min(xb, (w + 1) << 5) - (w << 5);   // 1..32
            const uint32_t m = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
            atomicOr(plane + (size_t)y * pitch + w, m);
        }
    }
}

// ------------------------------------------------------------------ 3x3 erosion on bit planes
__global__ void k_bbox_init(int32_t *__restrict__ bbox, int n_inst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_inst) {
        bbox[4 * i + 0] = 0x7fffffff;
        bbox[4 * i + 1] = 0x7fffffff;
        bbox[4 * i + 2] = -1;
        bbox[4 * i + 3] = -1;
    }
}

// Horizontal 3-AND of one row word; pixels outside the image count as set (cv2's default
// erosion border is +inf, so the border never lowers the minimum).
__device__ __forceinline__ uint32_t hmin3(const uint32_t *__restrict__ row, int xw, int pitch, uint32_t tailmask)
{
    uint32_t c = __ldg(row + xw);
    uint32_t l = xw > 0 ? __ldg(row + xw - 1) : 0xffffffffu;
    uint32_t r = xw + 1 < pitch ? __ldg(row + xw + 1) : 0xffffffffu;
    if (xw == pitch - 1) c |= ~tailmask;
    if (xw + 1 == pitch - 1) r |= ~tailmask;
    return c & ((c << 1) | (l >> 31)) & ((c >> 1) | (r << 31));
}

// Six consecutive words of one row around quad xw0: [xw0-1, xw0 .. xw0+3, xw0+4]; words outside the
// row read as all ones, the row's last word gets ones above the image width (see hmin3).
__device__ __forceinline__ void load_row6(const uint32_t *__restrict__ row, int xw0, int pitch, uint32_t tailmask,
                                          bool vec, uint32_t (&w)[6])
{
    if (vec) {
        const uint4 c = __ldg(reinterpret_cast<const uint4 *>(row + xw0));
        w[1] = c.x; w[2] = c.y; w[3] = c.z; w[4] = c.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) w[1 + k] = xw0 + k < pitch ? __ldg(row + xw0 + k) : 0xffffffffu;
    }
    w[0] = xw0 > 0 ? __ldg(row + xw0 - 1) : 0xffffffffu;
    w[5] = xw0 + 4 < pitch ? __ldg(row + xw0 + 4) : 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 6; ++k)
        if (xw0 - 1 + k == pitch - 1) w[k] |= ~tailmask;
}

// cv2.erode(mask, ones(3,3)) on a bit plane, four consecutive words (one 16-byte store) per thread.
// Planes whose pitch and offset are multiples of four words (every 1024-pixel-wide mask) take
// 16-byte loads and stores; others go word by word through the same arithmetic.
__global__ void __launch_bounds__(256)
k_erode3x3(const uint32_t *__restrict__ bits_in, const int32_t *__restrict__ inst_desc,
           const int32_t *__restrict__ row_range, uint32_t *__restrict__ bits_out, int32_t *__restrict__ bbox)
{
    const int i = blockIdx.y;
    const int32_t *d = inst_desc + i * CM3D_IN_WORDS;
    const int W = d[CM3D_IN_W], H = d[CM3D_IN_H], pitch = d[CM3D_IN_PITCH];
    const int64_t off = join64(d[CM3D_IN_BITS_LO], d[CM3D_IN_BITS_HI]);
    const bool vec = ((pitch & 3) == 0) && ((off & 3) == 0);
    const int quads_per_row = (pitch + 3) >> 2;
    // the grid is sized for planes whose pitch is a multiple of four; others take more trips
    for (int q0 = blockIdx.x * blockDim.x; q0 < H * quads_per_row; q0 += gridDim.x * blockDim.x) {
    const int q = q0 + threadIdx.x;
    const bool live = q < H * quads_per_row;
    const int y = live ? q / quads_per_row : 0;
    const int xw0 = live ? (q - y * quads_per_row) * 4 : 0;
    uint32_t out[4] = {0u, 0u, 0u, 0u};
    if (live) {
        const int r0 = row_range ? row_range[2 * i] : 0, r1 = row_range ? row_range[2 * i + 1] : H - 1;
        if (y >= r0 && y <= r1) {          // a row without a set pixel: nothing survives, nothing to read
            const uint32_t *plane = bits_in + off;
            const uint32_t tailmask = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
            uint32_t acc[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = y + dy;
                // rows outside the image do not lower the minimum; rows outside the set range are all zero
                if (yy < 0 || yy >= H) continue;
                if (yy < r0 || yy > r1) { acc[0] = acc[1] = acc[2] = acc[3] = 0u; continue; }
                uint32_t w[6];
                load_row6(plane + (size_t)yy * pitch, xw0, pitch, tailmask, vec, w);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t c = w[1 + k];
                    acc[k] &= c & ((c << 1) | (w[k] >> 31)) & ((c >> 1) | (w[2 + k] << 31));
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                out[k] = acc[k];
                if (xw0 + k == pitch - 1) out[k] &= tailmask;
                if (xw0 + k >= pitch) out[k] = 0u;
            }
        }
        uint32_t *dst = bits_out + off + (size_t)y * pitch + xw0;
        if (vec) {
            *reinterpret_cast<uint4 *>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (xw0 + k < pitch) dst[k] = out[k];
        }
    }
    // bounding box of the surviving pixels: warp-reduce, then 4 atomics per warp at most
    int xmin = 0x7fffffff, ymin = 0x7fffffff, xmax = -1, ymax = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (out[k]) {
            xmin = min(xmin, (xw0 + k) * 32 + (__ffs(out[k]) - 1));
            xmax = max(xmax, (xw0 + k) * 32 + (31 - __clz(out[k])));
            ymin = ymax = y;
        }
    }
    const unsigned full = 0xffffffffu;
    if (__any_sync(full, xmax >= 0)) {
        xmin = __reduce_min_sync(full, xmin);
        ymin = __reduce_min_sync(full, ymin);
        xmax = __reduce_max_sync(full, xmax);
        ymax = __reduce_max_sync(full, ymax);
        if (lane_id() == 0) {
            atomicMin(&bbox[4 * i + 0], xmin);
            atomicMin(&bbox[4 * i + 1], ymin);
            atomicMax(&bbox[4 * i + 2], xmax);
            atomicMax(&bbox[4 * i + 3], ymax);
        }
    }
    }
}

}  // namespace cm3d

using namespace cm3d;

extern "C" int cm3d_masks_pack_dense(const uint8_t *masks, const int64_t *src_off,
                                     const int32_t *inst_desc, int n_inst, int max_words,
                                     uint32_t *bits, void *stream)
{
    if (n_inst < 0 || max_words < 0) return CM3D_EINVAL;
    if (n_inst == 0 || max_words == 0) return CM3D_OK;
    if (!masks || !src_off || !inst_desc || !bits) return CM3D_EINVAL;
    dim3 grid((max_words + 255) / 256, n_inst);
    k_pack_dense<<<grid, 256, 0, (cudaStream_t)stream>>>(masks, src_off, inst_desc, bits);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}

extern "C" int cm3d_masks_decode_counts(const uint8_t *counts, const int64_t *byte_off, int n_inst,
                                        uint32_t *runs, void *stream)
{
    if (n_inst < 0) return CM3D_EINVAL;
    if (n_inst == 0) return CM3D_OK;
    if (!counts || !byte_off || !runs) return CM3D_EINVAL;
    k_rle_decode<<<n_inst, 256, 0, (cudaStream_t)stream>>>(counts, byte_off, runs);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}

extern "C" int cm3d_masks_fill_rle(const uint32_t *runs, const int64_t *run_off, uint32_t *run_start,
                                   const int32_t *inst_desc, int n_inst, int max_runs, uint32_t *bits,
                                   int32_t *row_range, int32_t *errflags, void *stream)
{
    if (n_inst < 0 || max_runs < 0) return CM3D_EINVAL;
    if (n_inst == 0) return CM3D_OK;
    if (!run_off || !inst_desc || !bits || !row_range || !errflags) return CM3D_EINVAL;
    if (!runs || !run_start) return CM3D_EINVAL;
    k_rle_prefix<<<n_inst, 256, 0, (cudaStream_t)stream>>>(runs, run_off, inst_desc, run_start, row_range, errflags);
    CM3D_LAUNCH_CHECK();
    if (max_runs == 0) return CM3D_OK;
    const int one_runs = (max_runs + 1) / 2;          // 1-runs sit at odd positions
    dim3 grid((one_runs + 7) / 8, n_inst);
    k_rle_fill<<<grid, 256, 0, (cudaStream_t)stream>>>(runs, run_off, run_start, inst_desc, bits);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}

extern "C" int cm3d_masks_erode3x3(const uint32_t *bits_in, const int32_t *inst_desc, const int32_t *row_range,
                                   int n_inst, int max_words, uint32_t *bits_out, int32_t *bbox, void *stream)
{
    if (n_inst < 0 || max_words < 0) return CM3D_EINVAL;
    if (n_inst == 0) return CM3D_OK;
    if (!bits_in || !inst_desc || !bits_out || !bbox) return CM3D_EINVAL;
    k_bbox_init<<<(n_inst + 255) / 256, 256, 0, (cudaStream_t)stream>>>(bbox, n_inst);
    CM3D_LAUNCH_CHECK();
    if (max_words == 0) return CM3D_OK;
    dim3 grid((max_words / 4 + 255) / 256 + 1, n_inst);      // four words per thread (the kernel loops if a plane needs more)
    k_erode3x3<<<grid, 256, 0, (cudaStream_t)stream>>>(bits_in, inst_desc, row_range, bits_out, bbox);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}
