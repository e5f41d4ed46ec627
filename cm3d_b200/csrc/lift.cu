// Sweeps -> aggregated cloud -> per-camera projection -> mask membership -> ordered
// per-instance segments, for a whole batch of frames per launch.
//
// Replaces (and runs ONCE per frame and camera what the reference redoes per mask):
//   k_aggregate        src/nuscenes/2d_to_3d.py:437-465  close-point filter, rotate/translate
//                      chain per sweep (utils/pcd.py:159-172), hstack;
//                      src/kitti/2d_to_3d.py:1066-1077 (project_velo_to_ref);
//                      src/waymo/2d_to_3d.py:472-481 (xyz + ones row)
//   k_project_count    src/nuscenes/2d_to_3d.py:553-613  clone, global->camera chain,
//                      view_points (utils/pcd.py:262-284), depth/bounds test, floor,
//                      mask lookup incl. the `logical_and(floored_points, ...)` quirk;
//                      src/kitti/2d_to_3d.py:1238-1351; src/waymo/2d_to_3d.py:557-616
//   k_scan_* / k_compact  src/nuscenes/2d_to_3d.py:615-620  track_points (ascending point
//                      index per instance) and `aggr_pc_points[:, track_points]`
//
// All of it is HBM-bound streaming work: raw tiles are staged with one bulk async copy
// (TMA, UBLKCP) per tile, the cloud lives as SoA (x|y|z|w) read with 128-bit loads, and
// order-preserving compaction uses warp ballots + small scans instead of atomics on HBM.
#include "common.cuh"

namespace cm3d {

constexpr int kGroups = kTile / 32;   // 32-slot groups per tile

// ------------------------------------------------------------------------------------------
// K1: one block per raw tile.
__global__ void __launch_bounds__(kBlock)
k_aggregate(const float *__restrict__ raw, const int32_t *__restrict__ tile_sweep,
            const int32_t *__restrict__ sweep_desc, const int32_t *__restrict__ frame_desc,
            const uint32_t *__restrict__ chains, float *__restrict__ xyzw, int64_t n_slots,
            int32_t *__restrict__ tile_cnt)
{
    __shared__ __align__(128) float s_raw[kTile * 5];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_chain[CM3D_CHAIN_WORDS];
    __shared__ int s_gcnt[kGroups];

    const int t = blockIdx.x;
    const int32_t *sd = sweep_desc + (size_t)tile_sweep[t] * CM3D_SW_WORDS;
    const int stride = sd[CM3D_SW_STRIDE];
    const int p0 = (t - sd[CM3D_SW_TILE_BASE]) * kTile;
    const int npts = min(kTile, sd[CM3D_SW_NPTS] - p0);
    const int fourth = sd[CM3D_SW_FOURTH];
    const int32_t *fd = frame_desc + (size_t)sd[CM3D_SW_FRAME] * CM3D_FR_WORDS;
    const bool use_close = fd[CM3D_FR_USE_CLOSE] != 0;
    const float close_thr = __int_as_float(fd[CM3D_FR_CLOSE_BITS]);
    const float *src = raw + join64(sd[CM3D_SW_RAW_LO], sd[CM3D_SW_RAW_HI]) + (int64_t)p0 * stride;

    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    if (threadIdx.x < CM3D_CHAIN_WORDS)
        s_chain[threadIdx.x] = chains[(size_t)sd[CM3D_SW_CHAIN] * CM3D_CHAIN_WORDS + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        // sweeps start 16-byte aligned and are padded to 16 bytes, tiles are 1024*stride*4 bytes
        const uint32_t bytes = (uint32_t)(((npts * stride * 4) + 15) & ~15);
        mbar_expect_tx(&s_bar, bytes);
        bulk_g2s(s_raw, src, bytes, &s_bar);
    }
    mbar_wait(&s_bar, 0);

    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const bool use_floor = fd[CM3D_FR_USE_FLOOR] != 0;          // default-off ground threshold
    const float floor_thr = __int_as_float(fd[CM3D_FR_FLOOR_BITS]);
    float px[kPerThread], py[kPerThread], pz[kPerThread], pw[kPerThread];
    unsigned ball[kPerThread];
#pragma unroll
    for (int r = 0; r < kPerThread; ++r) {
        const int p = r * kBlock + threadIdx.x;
        bool keep = p < npts;
        float x = 0.f, y = 0.f, z = 0.f, w = 1.0f;
        if (keep) {
            if (stride == 4) {              // one 16-byte load per point (stride 4 scalar loads would conflict 4-way)
                const float4 v = reinterpret_cast<const float4 *>(s_raw)[p];
                x = v.x; y = v.y; z = v.z;
                if (fourth == 1) w = v.w;
            } else {
                const float *q = s_raw + p * stride;
                x = q[0]; y = q[1]; z = q[2];
                if (fourth == 1) w = q[3];
            }
            if (use_close && fabsf(x) < close_thr && fabsf(y) < close_thr) keep = false;
            apply_chain(s_chain, x, y, z);
            // `aggr_pc_points[2] > floor_thresh` of the reference's commented-out ground filter
            // (src/kitti/2d_to_3d.py:1186-1190), on the aggregated (transformed) cloud
            if (use_floor && !(z > floor_thr)) keep = false;
        }
        px[r] = x; py[r] = y; pz[r] = z; pw[r] = w;
        ball[r] = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_gcnt[r * (kBlock / 32) + warp] = __popc(ball[r]);
    }
    __syncthreads();
    if (warp == 0) {
        const int c = s_gcnt[lane];
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += v;
        }
        s_gcnt[lane] = inc - c;
        if (lane == 31) tile_cnt[t] = inc;
    }
    __syncthreads();
    float *ox = xyzw, *oy = xyzw + n_slots, *oz = xyzw + 2 * n_slots, *ow = xyzw + 3 * n_slots;
#pragma unroll
    for (int r = 0; r < kPerThread; ++r) {
        if (ball[r] & (1u << lane)) {
            const int64_t slot = (int64_t)t * kTile + s_gcnt[r * (kBlock / 32) + warp] + __popc(ball[r] & lanemask_lt());
            ox[slot] = px[r]; oy[slot] = py[r]; oz[slot] = pz[r];
            if (fourth) ow[slot] = pw[r];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Per-frame tables staged in shared memory by the membership kernels.
struct VcamS {
    float4 m[CM3D_MAX_CHAIN][3];   // 12 constants per op (16-byte aligned: three broadcast LDS.128)
    float4 vp[3];                  // rows 0..2 of view_points' 4x4 viewpad
    int kind[CM3D_MAX_CHAIN];
    float wlim, hlim;              // float(W-1), float(H-1): the reference's strict upper bounds
    float wcap, hcap;              // float(W), float(H): conservative reject bounds (see project_n)
    int list_begin, list_count;
    const uint32_t *grid;          // instance lookup grid of this vcam
    int grid_nx, grid_nwords;
    float4 plane[5];               // conservative cull planes in q = p + tref coordinates (cm3d_b200.h)
    int flags, pad0, pad1, pad2;   // bit 0: simple viewpad
};
struct InstS {
    int xmin, ymin, xmax, ymax;   // eroded bbox, clamped to >= 1 (the reference drops fx==0 / fy==0)
    const uint32_t *plane;
    int pitch, pad;
};

struct FrameTables {
    VcamS vcam[CM3D_MAX_VCAMS];
    InstS inst[CM3D_MAX_INST + 2];
    uint8_t list[CM3D_MAX_INST + 2];
    int n_vcams, n_inst;
    float min_depth;
    float tref[3];
    int chain_sig;
};

__device__ __forceinline__ void load_frame_tables(FrameTables &ft, const int32_t *__restrict__ fd,
                                                  const int32_t *__restrict__ vcam_desc,
                                                  const int32_t *__restrict__ cam_inst_list,
                                                  const int32_t *__restrict__ inst_desc,
                                                  const int32_t *__restrict__ inst_bbox,
                                                  const uint32_t *__restrict__ chains,
                                                  const uint32_t *__restrict__ bits,
                                                  const uint32_t *__restrict__ vcam_grid)
{
    const int nv = fd[CM3D_FR_NVCAMS], ni = fd[CM3D_FR_NINST];
    const int v0 = fd[CM3D_FR_VCAM_BEGIN], i0 = fd[CM3D_FR_INST_BEGIN], l0 = fd[CM3D_FR_LIST_BEGIN];
    if (threadIdx.x == 0) {
        ft.n_vcams = nv;
        ft.n_inst = ni;
        ft.min_depth = __int_as_float(fd[CM3D_FR_MIN_DEPTH_BITS]);
        ft.tref[0] = __int_as_float(fd[CM3D_FR_TREF]);
        ft.tref[1] = __int_as_float(fd[CM3D_FR_TREF + 1]);
        ft.tref[2] = __int_as_float(fd[CM3D_FR_TREF + 2]);
        ft.chain_sig = fd[CM3D_FR_CHAIN_SIG];
    }
    for (int k = threadIdx.x; k < nv * 21; k += blockDim.x) {
        const int v = k / 21, wd = k - v * 21;
        const int32_t *vd = vcam_desc + (size_t)(v0 + v) * CM3D_VC_WORDS;
        if (wd < 20) reinterpret_cast<float *>(ft.vcam[v].plane)[wd] = __int_as_float(vd[CM3D_VC_PLANES + wd]);
        else ft.vcam[v].flags = vd[CM3D_VC_FLAGS];
    }
    for (int k = threadIdx.x; k < nv * CM3D_CHAIN_WORDS; k += blockDim.x) {
        const int v = k / CM3D_CHAIN_WORDS, wd = k - v * CM3D_CHAIN_WORDS;
        const int op = wd / CM3D_OP_WORDS, q = wd - op * CM3D_OP_WORDS;
        const int32_t *vd = vcam_desc + (size_t)(v0 + v) * CM3D_VC_WORDS;
        const uint32_t val = chains[(size_t)vd[CM3D_VC_CHAIN] * CM3D_CHAIN_WORDS + wd];
        if (q == 0) ft.vcam[v].kind[op] = (int)val;
        else if (q <= 12) reinterpret_cast<float *>(ft.vcam[v].m[op])[q - 1] = __uint_as_float(val);
    }
    for (int k = threadIdx.x; k < nv * 16; k += blockDim.x) {
        const int v = k >> 4, wd = k & 15;
        const int32_t *vd = vcam_desc + (size_t)(v0 + v) * CM3D_VC_WORDS;
        if (wd < 12) reinterpret_cast<float *>(ft.vcam[v].vp)[wd] = __int_as_float(vd[CM3D_VC_VIEWPAD + wd]);
        else if (wd == 12) { ft.vcam[v].wlim = (float)(vd[CM3D_VC_W] - 1); ft.vcam[v].wcap = (float)vd[CM3D_VC_W]; }
        else if (wd == 13) { ft.vcam[v].hlim = (float)(vd[CM3D_VC_H] - 1); ft.vcam[v].hcap = (float)vd[CM3D_VC_H]; }
        else if (wd == 14) ft.vcam[v].list_begin = vd[CM3D_VC_LIST_BEGIN];
        else {
            ft.vcam[v].list_count = vd[CM3D_VC_LIST_COUNT];
            ft.vcam[v].grid = vcam_grid + vd[CM3D_VC_GRID_OFF];
            ft.vcam[v].grid_nx = vd[CM3D_VC_GRID_NX];
            ft.vcam[v].grid_nwords = (vd[CM3D_VC_LIST_COUNT] + 31) >> 5;
        }
    }
    for (int j = threadIdx.x; j < ni; j += blockDim.x) {
        const int32_t *d = inst_desc + (size_t)(i0 + j) * CM3D_IN_WORDS;
        const int32_t *bb = inst_bbox + (size_t)(i0 + j) * 4;
        InstS s;
        s.xmin = max(bb[0], 1); s.ymin = max(bb[1], 1); s.xmax = bb[2]; s.ymax = bb[3];
        s.plane = bits + join64(d[CM3D_IN_BITS_LO], d[CM3D_IN_BITS_HI]);
        s.pitch = d[CM3D_IN_PITCH];
        s.pad = 0;
        ft.inst[j] = s;
        ft.list[j] = (uint8_t)cam_inst_list[l0 + j];
    }
}

// Project NP points into vcam `vc`; code = fx | fy<<16, or -1 when the point fails the
// reference's `depths > min_dist, 0 < u < W-1, 0 < v < H-1` test (nuscenes:597-603).
// The op kinds are uniform over the block, so the constants are loaded once per op and
// applied to all NP points.  Before the two IEEE divisions a division-free test throws out
// points that cannot pass: for r2 > 0, r0 <= 0 gives u <= 0, and r0 >= fl(r2*W) gives a true
// quotient >= W(1-2^-24) > W-1/2, whose rounding is >= W-1; both fail the strict bounds, so the
// filter never changes a result (same for v).
template <int NP>
__device__ __forceinline__ void project_n(const VcamS &vc, float min_depth, const float (&px)[NP],
                                          const float (&py)[NP], const float (&pz)[NP], int32_t (&code)[NP])
{
    float x[NP], y[NP], z[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) { x[p] = px[p]; y[p] = py[p]; z[p] = pz[p]; }
#pragma unroll
    for (int k = 0; k < CM3D_MAX_CHAIN; ++k) {
        const int kind = vc.kind[k];
        if (kind == CM3D_OP_END) break;
        const float4 a = vc.m[k][0], b = vc.m[k][1], c = vc.m[k][2];
        if (kind == CM3D_OP_T) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                x[p] = __fadd_rn(x[p], a.x); y[p] = __fadd_rn(y[p], a.y); z[p] = __fadd_rn(z[p], a.z);
            }
        } else if (kind == CM3D_OP_R) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float u = x[p], v = y[p], w = z[p];
                x[p] = __fmaf_rn(a.z, w, __fmaf_rn(a.y, v, __fmul_rn(a.x, u)));
                y[p] = __fmaf_rn(b.y, w, __fmaf_rn(b.x, v, __fmul_rn(a.w, u)));
                z[p] = __fmaf_rn(c.x, w, __fmaf_rn(b.w, v, __fmul_rn(b.z, u)));
            }
        } else {  // CM3D_OP_A: [u v w 1] . row
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float u = x[p], v = y[p], w = z[p];
                x[p] = __fmaf_rn(1.0f, a.w, __fmaf_rn(w, a.z, __fmaf_rn(v, a.y, __fmul_rn(u, a.x))));
                y[p] = __fmaf_rn(1.0f, b.w, __fmaf_rn(w, b.z, __fmaf_rn(v, b.y, __fmul_rn(u, b.x))));
                z[p] = __fmaf_rn(1.0f, c.w, __fmaf_rn(w, c.z, __fmaf_rn(v, c.y, __fmul_rn(u, c.x))));
            }
        }
    }
    const float4 k0 = vc.vp[0], k1 = vc.vp[1], k2 = vc.vp[2];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        code[p] = -1;
        if (!(z[p] > min_depth)) continue;
        const float r0 = __fmaf_rn(k0.w, 1.0f, __fmaf_rn(k0.z, z[p], __fmaf_rn(k0.y, y[p], __fmul_rn(k0.x, x[p]))));
        const float r1 = __fmaf_rn(k1.w, 1.0f, __fmaf_rn(k1.z, z[p], __fmaf_rn(k1.y, y[p], __fmul_rn(k1.x, x[p]))));
        const float r2 = __fmaf_rn(k2.w, 1.0f, __fmaf_rn(k2.z, z[p], __fmaf_rn(k2.y, y[p], __fmul_rn(k2.x, x[p]))));
        if (r2 > 0.0f && (!(r0 > 0.0f) || !(r1 > 0.0f) || r0 >= __fmul_rn(r2, vc.wcap) || r1 >= __fmul_rn(r2, vc.hcap)))
            continue;
        const float u = __fdiv_rn(r0, r2), v = __fdiv_rn(r1, r2);
        if (u > 0.0f && u < vc.wlim && v > 0.0f && v < vc.hlim)
            code[p] = (int32_t)floorf(u) | ((int32_t)floorf(v) << 16);
    }
}

// Calls f(j) for every instance j of vcam `vc` whose eroded mask has pixel `code` set; ids ascend.
// The cell grid narrows the vcam's instance list to those whose bbox touches the pixel's cell.
template <class F>
__device__ __forceinline__ void hits_in_vcam(const FrameTables &ft, const VcamS &vc, int32_t code, F f)
{
    const int fx = code & 0xffff, fy = code >> 16;
    const uint32_t *cell = vc.grid + (size_t)((fy / CM3D_CELL) * vc.grid_nx + fx / CM3D_CELL) * vc.grid_nwords;
    for (int w = 0; w < vc.grid_nwords; ++w) {
        uint32_t cand = __ldg(cell + w);
        while (cand) {
            const int k = vc.list_begin + w * 32 + (__ffs(cand) - 1);
            cand &= cand - 1u;
            const int j = ft.list[k];
            const InstS &s = ft.inst[j];
            if (fx < s.xmin || fx > s.xmax || fy < s.ymin || fy > s.ymax) continue;
            const uint32_t wd = __ldg(s.plane + (size_t)fy * s.pitch + (fx >> 5));
            if ((wd >> (fx & 31)) & 1u) f(j);
        }
    }
}

// Single-point form (rare paths): f(j) for every instance the point belongs to.
template <class F>
__device__ __forceinline__ void for_each_hit(const FrameTables &ft, float x, float y, float z, int32_t *pix,
                                             int64_t pix_stride, F f)
{
    const float xs[1] = {x}, ys[1] = {y}, zs[1] = {z};
    for (int v = 0; v < ft.n_vcams; ++v) {
        int32_t code[1];
        project_n<1>(ft.vcam[v], ft.min_depth, xs, ys, zs, code);
        if (pix) pix[v * pix_stride] = code[0];
        if (code[0] >= 0) hits_in_vcam(ft, ft.vcam[v], code[0], f);
    }
}

// One thread per grid cell of one vcam: bit k of the cell = "bbox of the vcam's k-th instance
// (eroded pixels, fx/fy >= 1) touches the cell".
__global__ void __launch_bounds__(256)
k_build_vcam_grid(const int32_t *__restrict__ vcam_desc, const int32_t *__restrict__ frame_desc,
                  const int32_t *__restrict__ cam_inst_list, const int32_t *__restrict__ inst_bbox,
                  uint32_t *__restrict__ vcam_grid)
{
    const int32_t *vd = vcam_desc + (size_t)blockIdx.y * CM3D_VC_WORDS;
    const int nx = vd[CM3D_VC_GRID_NX], ny = (vd[CM3D_VC_H] + CM3D_CELL - 1) / CM3D_CELL;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nx * ny) return;
    const int32_t *fd = frame_desc + (size_t)vd[CM3D_VC_FRAME] * CM3D_FR_WORDS;
    const int32_t *list = cam_inst_list + fd[CM3D_FR_LIST_BEGIN] + vd[CM3D_VC_LIST_BEGIN];
    const int n = vd[CM3D_VC_LIST_COUNT], nwords = (n + 31) >> 5;
    const int cy = c / nx, cx = c - cy * nx;
    const int x0 = cx * CM3D_CELL, x1 = x0 + CM3D_CELL - 1, y0 = cy * CM3D_CELL, y1 = y0 + CM3D_CELL - 1;
    uint32_t *out = vcam_grid + vd[CM3D_VC_GRID_OFF] + (size_t)c * nwords;
    for (int w = 0; w < nwords; ++w) {
        uint32_t m = 0;
        for (int b = 0; b < 32 && w * 32 + b < n; ++b) {
            const int32_t *bb = inst_bbox + (size_t)(fd[CM3D_FR_INST_BEGIN] + list[w * 32 + b]) * 4;
            if (bb[0] <= x1 && bb[2] >= x0 && bb[1] <= y1 && bb[3] >= y0) m |= 1u << b;
        }
        out[w] = m;
    }
}

// Chain signatures (sum kind_k * 4^k) the kernels are specialised for; anything else runs the
// generic path that reads the op kinds from shared memory.
constexpr int kSigTRTR = 1 + (2 << 2) + (1 << 4) + (2 << 6);   // nuScenes: global -> ego -> camera
constexpr int kSigTR = 1 + (2 << 2);                           // Waymo: vehicle -> camera
constexpr int kSigAAR = 3 + (3 << 2) + (2 << 4);               // KITTI: ref -> velo -> ref -> rect
constexpr int kSigGeneric = -1;

template <int SIG>
__device__ __forceinline__ int op_kind(const VcamS &vc, int k)
{
    return SIG == kSigGeneric ? vc.kind[k] : ((SIG >> (2 * k)) & 3);
}

// Exact projection of NP points into vcam `vc` (same arithmetic as project_n), with the op kinds
// fixed at compile time and the zero terms of a plain pinhole viewpad dropped: with
// K = [[fx,0,cx],[0,fy,cy],[0,0,1]] the reference's 4-term chains reduce to fma(cx,z,fx*x),
// fma(cy,z,fy*y) and z for every finite point (adding a 0*y or 0*1 term changes at most the sign
// of a zero), and non-finite points are rejected either way.
template <int SIG, int NP>
__device__ __forceinline__ void project_sig(const VcamS &vc, float min_depth, const float (&px)[NP],
                                            const float (&py)[NP], const float (&pz)[NP], int32_t (&code)[NP])
{
    float x[NP], y[NP], z[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) { x[p] = px[p]; y[p] = py[p]; z[p] = pz[p]; }
#pragma unroll
    for (int k = 0; k < CM3D_MAX_CHAIN; ++k) {
        const int kind = op_kind<SIG>(vc, k);
        if (kind == CM3D_OP_END) break;
        const float4 a = vc.m[k][0], b = vc.m[k][1], c = vc.m[k][2];
        if (kind == CM3D_OP_T) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                x[p] = __fadd_rn(x[p], a.x); y[p] = __fadd_rn(y[p], a.y); z[p] = __fadd_rn(z[p], a.z);
            }
        } else if (kind == CM3D_OP_R) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float u = x[p], v = y[p], w = z[p];
                x[p] = __fmaf_rn(a.z, w, __fmaf_rn(a.y, v, __fmul_rn(a.x, u)));
                y[p] = __fmaf_rn(b.y, w, __fmaf_rn(b.x, v, __fmul_rn(a.w, u)));
                z[p] = __fmaf_rn(c.x, w, __fmaf_rn(b.w, v, __fmul_rn(b.z, u)));
            }
        } else {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float u = x[p], v = y[p], w = z[p];
                x[p] = __fmaf_rn(1.0f, a.w, __fmaf_rn(w, a.z, __fmaf_rn(v, a.y, __fmul_rn(u, a.x))));
                y[p] = __fmaf_rn(1.0f, b.w, __fmaf_rn(w, b.z, __fmaf_rn(v, b.y, __fmul_rn(u, b.x))));
                z[p] = __fmaf_rn(1.0f, c.w, __fmaf_rn(w, c.z, __fmaf_rn(v, c.y, __fmul_rn(u, c.x))));
            }
        }
    }
    const float4 k0 = vc.vp[0], k1 = vc.vp[1], k2 = vc.vp[2];
    const bool simple = (vc.flags & 1) != 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        code[p] = -1;
        if (!(z[p] > min_depth)) continue;
        float r0, r1, r2;
        if (simple) {
            r0 = __fmaf_rn(k0.z, z[p], __fmul_rn(k0.x, x[p]));
            r1 = __fmaf_rn(k1.z, z[p], __fmul_rn(k1.y, y[p]));
            r2 = z[p];
        } else {
            r0 = __fmaf_rn(k0.w, 1.0f, __fmaf_rn(k0.z, z[p], __fmaf_rn(k0.y, y[p], __fmul_rn(k0.x, x[p]))));
            r1 = __fmaf_rn(k1.w, 1.0f, __fmaf_rn(k1.z, z[p], __fmaf_rn(k1.y, y[p], __fmul_rn(k1.x, x[p]))));
            r2 = __fmaf_rn(k2.w, 1.0f, __fmaf_rn(k2.z, z[p], __fmaf_rn(k2.y, y[p], __fmul_rn(k2.x, x[p]))));
        }
        if (r2 > 0.0f && (!(r0 > 0.0f) || !(r1 > 0.0f) || r0 >= __fmul_rn(r2, vc.wcap) || r1 >= __fmul_rn(r2, vc.hcap)))
            continue;
        const float u = __fdiv_rn(r0, r2), v = __fdiv_rn(r1, r2);
        if (u > 0.0f && u < vc.wlim && v > 0.0f && v < vc.hlim)
            code[p] = (int32_t)floorf(u) | ((int32_t)floorf(v) << 16);
    }
}

// Up to four ids (1..254) appended in any order, one per byte, zeros on top -> ascending from byte 0
// with the zeros still on top (the layout k_compact reads).  An overflowed word (byte 3 == 255)
// is recomputed by k_compact and passes through.
__device__ __forceinline__ uint32_t hit_sort(uint32_t h)
{
    if (h < 0x100u || (h >> 24) == 0xffu) return h;          // zero or one id, or overflow
    uint32_t b0 = h & 0xffu, b1 = (h >> 8) & 0xffu, b2 = (h >> 16) & 0xffu, b3 = h >> 24;
    // empty bytes sort last: map 0 -> 256
    b0 = b0 ? b0 : 256u; b1 = b1 ? b1 : 256u; b2 = b2 ? b2 : 256u; b3 = b3 ? b3 : 256u;
    uint32_t t;
#define CM3D_CSWAP(a, b) t = min(a, b); b = max(a, b); a = t;
    CM3D_CSWAP(b0, b1) CM3D_CSWAP(b2, b3) CM3D_CSWAP(b0, b2) CM3D_CSWAP(b1, b3) CM3D_CSWAP(b1, b2)
#undef CM3D_CSWAP
    return (b0 & 0xffu) | ((b1 & 0xffu) << 8) | ((b2 & 0xffu) << 16) | ((b3 & 0xffu) << 24);
}

// Sorted insert of id (1..254) into a 4-byte hit word; byte 3 becomes 255 on overflow.
__device__ __forceinline__ uint32_t hit_insert(uint32_t hw, uint32_t id)
{
    if (hw >> 24) return hw | 0xff000000u;          // already four (or overflowed)
    // shift bytes greater than id up by one
    uint32_t out = 0, placed = 0;
    int pos = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const uint32_t cur = (hw >> (8 * b)) & 0xffu;
        if (cur == 0) break;
        if (!placed && id < cur) { out |= id << (8 * pos++); placed = 1; }
        out |= cur << (8 * pos++);
    }
    if (!placed) out |= id << (8 * pos);
    return out;
}

// K2 (count pass): one block per tile of the aggregated cloud, 4 consecutive slots per thread, and
// every WARP on its own: a warp owns 128 consecutive slots (points in firing order: one narrow
// azimuth wedge), its own in-image list and its own hit words, so there is no block barrier between
// cameras and a warp that does not see a camera moves on while another one walks its masks.
// Per vcam: (1) every point is tested against the vcam's five conservative cull planes (3 FMAs + a
// compare each, early exit; whole warps drop a camera together); (2) threads that still hold a
// candidate run the exact chain; in-image points are compacted (ballots, no atomics) into the
// warp's list; (3) the list is walked by all 32 lanes (instance lookup grid -> one bit probe per
// candidate), so the walk is load balanced instead of divergent.
// Hit words are kept in shared memory during the walk (a slot is touched by one lane per vcam
// pass) and leave with one 16-byte store per thread.
constexpr int kWarpSlots = kTile / (kBlock / 32);      // slots (and list entries) per warp

template <int SIG>
__device__ __forceinline__ void project_count_tile(FrameTables &ft, int *s_hist, uint32_t *s_hw, int32_t *s_lcode,
                                                   uint16_t *s_lslot, const float *__restrict__ xyzw,
                                                   int64_t n_slots, int64_t base, int cnt, int32_t *__restrict__ pix)
{
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    int32_t *wl_code = s_lcode + warp * kWarpSlots;
    uint16_t *wl_slot = s_lslot + warp * kWarpSlots;
    const int s0 = threadIdx.x * 4;
    const int np = max(0, min(4, cnt - s0));
    float xs[4] = {0.f, 0.f, 0.f, 0.f}, ys[4] = {0.f, 0.f, 0.f, 0.f}, zs[4] = {0.f, 0.f, 0.f, 0.f};
    float qx[4], qy[4], qz[4], nmg[4];
    if (np > 0) {
        const float4 X = __ldg(reinterpret_cast<const float4 *>(xyzw + base + s0));
        const float4 Y = __ldg(reinterpret_cast<const float4 *>(xyzw + n_slots + base + s0));
        const float4 Z = __ldg(reinterpret_cast<const float4 *>(xyzw + 2 * n_slots + base + s0));
        xs[0] = X.x; xs[1] = X.y; xs[2] = X.z; xs[3] = X.w;
        ys[0] = Y.x; ys[1] = Y.y; ys[2] = Y.z; ys[3] = Y.w;
        zs[0] = Z.x; zs[1] = Z.y; zs[2] = Z.z; zs[3] = Z.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        qx[k] = __fadd_rn(xs[k], ft.tref[0]); qy[k] = __fadd_rn(ys[k], ft.tref[1]); qz[k] = __fadd_rn(zs[k], ft.tref[2]);
        nmg[k] = -((fabsf(qx[k]) + fabsf(qy[k])) + fabsf(qz[k])) * 0x1p-17f;     // -S * 2^-17
    }
    *reinterpret_cast<uint4 *>(s_hw + s0) = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();

    for (int v = 0; v < ft.n_vcams; ++v) {
        const VcamS &vc = ft.vcam[v];
        // (1) cull planes: plane-outer so each plane is one LDS.128 for the thread's four points;
        // a thread leaves as soon as none of its points is left (depth first, then left/right, ...)
        unsigned cand = (1u << np) - 1u;
#pragma unroll
        for (int pl = 0; pl < 5; ++pl) {
            if (cand == 0) break;
            const float4 n = vc.plane[pl];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float val = __fmaf_rn(n.x, qx[k], __fmaf_rn(n.y, qy[k], __fmaf_rn(n.z, qz[k], n.w)));
                if (val < nmg[k]) cand &= ~(1u << k);
            }
        }
        int32_t code[4] = {-1, -1, -1, -1};
        int n_list = 0;
        if (__any_sync(0xffffffffu, cand != 0)) {
            // (2) exact chain for the threads that still hold a candidate
            if (cand != 0) {
                project_sig<SIG, 4>(vc, ft.min_depth, xs, ys, zs, code);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (!((cand >> k) & 1u)) code[k] = -1;       // k >= np, or culled (then it is -1 anyway)
            }
            // in-image points -> the warp's list, in slot order; pixel column / row 0 is never a member
            // (the reference's `logical_and(floored_points, ...)` quirk)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool in = code[k] >= 0 && (code[k] & 0xffff) != 0 && (code[k] >> 16) != 0;
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const int at = n_list + __popc(bal & lanemask_lt());
                    wl_code[at] = code[k];
                    wl_slot[at] = (uint16_t)(s0 + k);
                }
                n_list += __popc(bal);
            }
        }
        if (pix) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < np) pix[(int64_t)v * n_slots + base + s0 + k] = code[k];
        }
        if (n_list == 0) continue;             // uniform over the warp
        __syncwarp();
        // (3) walk the warp's in-image list of this vcam
        for (int e = lane; e < n_list; e += 32) {
            const int32_t code = wl_code[e];
            const int fx = code & 0xffff, fy = code >> 16;
            const uint32_t *cell = vc.grid + (size_t)((fy / CM3D_CELL) * vc.grid_nx + fx / CM3D_CELL) * vc.grid_nwords;
            const uint32_t bit = 1u << (fx & 31);
            uint32_t h = 0;
            bool touched = false;
            for (int w = 0; w < vc.grid_nwords; ++w) {
                uint32_t c = __ldg(cell + w);
                while (c) {
                    const int j = ft.list[vc.list_begin + w * 32 + (__ffs(c) - 1)];
                    c &= c - 1u;
                    const InstS &si = ft.inst[j];
                    if (__ldg(si.plane + (size_t)fy * si.pitch + (fx >> 5)) & bit) {
                        if (!touched) { h = s_hw[wl_slot[e]]; touched = true; }
                        h = (h >> 24) ? (h | 0xff000000u) : ((h << 8) | (uint32_t)(j + 1));   // append, newest in byte 0
                        atomicAdd(&s_hist[j], 1);
                    }
                }
            }
            if (touched) s_hw[wl_slot[e]] = h;
        }
        __syncwarp();                          // the list is reused by the next vcam
    }
}

__global__ void __launch_bounds__(kBlock, 4)
k_project_count(const float *__restrict__ xyzw, int64_t n_slots, const int32_t *__restrict__ tile_cnt,
                const int32_t *__restrict__ tile_sweep, const int32_t *__restrict__ sweep_desc,
                const int32_t *__restrict__ frame_desc, const int32_t *__restrict__ vcam_desc,
                const int32_t *__restrict__ cam_inst_list, const int32_t *__restrict__ inst_desc,
                const int32_t *__restrict__ inst_bbox, const uint32_t *__restrict__ chains,
                const uint32_t *__restrict__ bits, const uint32_t *__restrict__ vcam_grid,
                uint32_t *__restrict__ hits, uint16_t *__restrict__ tile_inst_cnt, int32_t *__restrict__ pix)
{
    __shared__ FrameTables ft;
    __shared__ int s_hist[CM3D_MAX_INST + 2];
    __shared__ __align__(16) uint32_t s_hw[kTile];
    __shared__ int32_t s_lcode[kTile];
    __shared__ uint16_t s_lslot[kTile];

    const int t = blockIdx.x;
    const int f = sweep_desc[(size_t)tile_sweep[t] * CM3D_SW_WORDS + CM3D_SW_FRAME];
    const int32_t *fd = frame_desc + (size_t)f * CM3D_FR_WORDS;
    load_frame_tables(ft, fd, vcam_desc, cam_inst_list, inst_desc, inst_bbox, chains, bits, vcam_grid);
    for (int j = threadIdx.x; j < CM3D_MAX_INST + 2; j += blockDim.x) s_hist[j] = 0;
    __syncthreads();

    const int cnt = tile_cnt[t];
    const int64_t base = (int64_t)t * kTile;
    switch (ft.chain_sig) {
    case kSigTRTR: project_count_tile<kSigTRTR>(ft, s_hist, s_hw, s_lcode, s_lslot, xyzw, n_slots, base, cnt, pix); break;
    case kSigTR:   project_count_tile<kSigTR>(ft, s_hist, s_hw, s_lcode, s_lslot, xyzw, n_slots, base, cnt, pix); break;
    case kSigAAR:  project_count_tile<kSigAAR>(ft, s_hist, s_hw, s_lcode, s_lslot, xyzw, n_slots, base, cnt, pix); break;
    default:       project_count_tile<kSigGeneric>(ft, s_hist, s_hw, s_lcode, s_lslot, xyzw, n_slots, base, cnt, pix); break;
    }
    __syncthreads();
    const int s0 = threadIdx.x * 4;
    if (s0 < cnt) {
        uint4 hw = *reinterpret_cast<const uint4 *>(s_hw + s0);
        hw.x = hit_sort(hw.x); hw.y = hit_sort(hw.y); hw.z = hit_sort(hw.z); hw.w = hit_sort(hw.w);
        *reinterpret_cast<uint4 *>(hits + base + s0) = hw;
    }
    const int tl = t - fd[CM3D_FR_TILE_BEGIN], ntf = fd[CM3D_FR_TILE_END] - fd[CM3D_FR_TILE_BEGIN];
    uint16_t *out = tile_inst_cnt + (size_t)fd[CM3D_FR_CNT_OFF] + tl;
    for (int j = threadIdx.x; j < ft.n_inst; j += blockDim.x) out[(size_t)j * ntf] = (uint16_t)s_hist[j];
}

// ------------------------------------------------------------------------------------------
// Scans.  k_scan_rows: one warp per (frame, row); rows 0..n_inst-1 are instances (prefix of
// tile_inst_cnt over the frame's tiles), row n_inst is the tile occupancy (tile_prefix).
__global__ void __launch_bounds__(256)
k_scan_rows(const int32_t *__restrict__ tile_cnt, const uint16_t *__restrict__ tile_inst_cnt,
            const int32_t *__restrict__ frame_desc, int32_t *__restrict__ tile_prefix,
            int32_t *__restrict__ frame_n, int32_t *__restrict__ tile_inst_base,
            int32_t *__restrict__ seg_count)
{
    const int f = blockIdx.x;
    const int row = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int32_t *fd = frame_desc + (size_t)f * CM3D_FR_WORDS;
    const int ni = fd[CM3D_FR_NINST];
    if (row > ni) return;
    const int tb = fd[CM3D_FR_TILE_BEGIN], ntf = fd[CM3D_FR_TILE_END] - tb;
    const unsigned lane = lane_id();
    const bool is_tiles = row == ni;
    const uint16_t *src16 = tile_inst_cnt + (size_t)fd[CM3D_FR_CNT_OFF] + (size_t)row * ntf;
    int32_t *dst = is_tiles ? tile_prefix + tb : tile_inst_base + (size_t)fd[CM3D_FR_CNT_OFF] + (size_t)row * ntf;
    int run = 0;
    for (int c = 0; c < ntf; c += 32) {
        const int k = c + lane;
        const int v = k < ntf ? (is_tiles ? tile_cnt[tb + k] : (int)src16[k]) : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += u;
        }
        if (k < ntf) dst[k] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
        if (is_tiles) frame_n[f] = run;
        else seg_count[fd[CM3D_FR_INST_BEGIN] + row] = run;
    }
}

// k_scan_batch: one block.  seg_off = exclusive prefix of the member counts over all instances of
// the batch; then the medoid schedule: instances are ordered by their number of work items,
// largest first (64 buckets; the medoid grid is dynamic, so the long items start early and the
// launch has a short tail), item_inst[p] = the p-th instance of that order and item_off = the
// exclusive prefix of the item counts in that order.
__global__ void __launch_bounds__(1024)
k_scan_batch(const int32_t *__restrict__ seg_count, const int32_t *__restrict__ inst_desc,
             const int32_t *__restrict__ frame_desc, int n_inst_total, int64_t seg_cap,
             int32_t *__restrict__ seg_off, int32_t *__restrict__ item_off, int32_t *__restrict__ item_inst,
             unsigned long long *__restrict__ medoid_best, int32_t *__restrict__ errflags)
{
    __shared__ long long s_wa[32];
    __shared__ long long s_ca;
    __shared__ int s_hist[64], s_cursor[64];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_ca = 0;
    if (threadIdx.x < 64) s_hist[threadIdx.x] = 0;
    __syncthreads();
    // ---- seg_off (seg_count is staged in seg_off[1..]: element i+1 is read before it is written)
    for (int base = 0; base < n_inst_total; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const long long a = i < n_inst_total ? seg_count[i] : 0;
        if (i < n_inst_total) medoid_best[i] = ~0ull;
        long long ia = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long ua = __shfl_up_sync(0xffffffffu, ia, o);
            if (lane >= (unsigned)o) ia += ua;
        }
        if (lane == 31) s_wa[warp] = ia;
        __syncthreads();
        long long oa = s_ca;
        for (unsigned w = 0; w < warp; ++w) oa += s_wa[w];
        if (i < n_inst_total) seg_off[i] = (int32_t)min(oa + ia - a, (long long)0x7fffffff);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_ca = oa + ia;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        seg_off[n_inst_total] = (int32_t)min(s_ca, (long long)0x7fffffff);
        if (s_ca > seg_cap) errflags[CM3D_ERR_SEG_OVERFLOW] = (int32_t)min(s_ca, (long long)0x7fffffff);
    }
    __syncthreads();
    // ---- medoid schedule
    auto items_of = [&](int i) {
        const int m = seg_off[i + 1] - seg_off[i];
        const int f = inst_desc[(size_t)i * CM3D_IN_WORDS + CM3D_IN_FRAME];
        return medoid_items(m, frame_desc[(size_t)f * CM3D_FR_WORDS + CM3D_FR_MIN_MEDOID_PTS]);
    };
    for (int i = threadIdx.x; i < n_inst_total; i += blockDim.x) atomicAdd(&s_hist[63 - min(items_of(i), 63)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < 64; ++k) { s_cursor[k] = run; run += s_hist[k]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_inst_total; i += blockDim.x) {
        const int b = items_of(i);
        const int p = atomicAdd(&s_cursor[63 - min(b, 63)], 1);
        item_inst[p] = i;
        item_off[p] = b;                  // count for now, prefix below
    }
    if (threadIdx.x == 0) s_ca = 0;
    __syncthreads();
    for (int base = 0; base < n_inst_total; base += blockDim.x) {
        const int p = base + threadIdx.x;
        const long long a = p < n_inst_total ? item_off[p] : 0;
        long long ia = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long ua = __shfl_up_sync(0xffffffffu, ia, o);
            if (lane >= (unsigned)o) ia += ua;
        }
        if (lane == 31) s_wa[warp] = ia;
        __syncthreads();
        long long oa = s_ca;
        for (unsigned w = 0; w < warp; ++w) oa += s_wa[w];
        if (p < n_inst_total) item_off[p] = (int32_t)(oa + ia - a);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_ca = oa + ia;
        __syncthreads();
    }
    if (threadIdx.x == 0) item_off[n_inst_total] = (int32_t)s_ca;
}

// ------------------------------------------------------------------------------------------
// K3 (write pass): one block per tile.  Ranks inside a tile come from warp ballots taken in
// ascending instance order per 32-slot group, so every segment is in ascending point order.
// The ballot walk runs twice: once to count members per (group, instance), once - after an
// exclusive prefix over the groups - to scatter.
constexpr int kFastInst = 24;    // k_compact: instances per tile the ballot-per-instance path handles

struct PendingHits {
    uint32_t word;       // up to four ids, ascending, one per byte (non-overflow points)
    uint32_t bm[8];      // membership bitmask of an overflow point (more than four masks)
    bool ovf;
    __device__ __forceinline__ uint32_t front() const
    {
        if (!ovf) return word & 0xffu;
        for (int w = 0; w < 8; ++w)
            if (bm[w]) return (uint32_t)(w * 32 + __ffs(bm[w]));   // id = j + 1
        return 0u;
    }
    __device__ __forceinline__ void pop()
    {
        if (!ovf) { word >>= 8; return; }
        for (int w = 0; w < 8; ++w)
            if (bm[w]) { bm[w] &= bm[w] - 1u; return; }
    }
};

__global__ void __launch_bounds__(kBlock)
k_compact(const float *__restrict__ xyzw, int64_t n_slots, const int32_t *__restrict__ tile_cnt,
          const int32_t *__restrict__ tile_prefix, const int32_t *__restrict__ tile_sweep,
          const int32_t *__restrict__ sweep_desc, const int32_t *__restrict__ frame_desc,
          const int32_t *__restrict__ vcam_desc, const int32_t *__restrict__ cam_inst_list,
          const int32_t *__restrict__ inst_desc, const int32_t *__restrict__ inst_bbox,
          const uint32_t *__restrict__ chains, const uint32_t *__restrict__ bits,
          const uint32_t *__restrict__ vcam_grid, const uint32_t *__restrict__ hits,
          const int32_t *__restrict__ tile_inst_base,
          const int32_t *__restrict__ seg_off, int32_t *__restrict__ seg_point_idx,
          float *__restrict__ seg_xyzw, int64_t seg_cap, const int32_t *__restrict__ errflags)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];   // [kGroups][ni] u16, then [ni] int
    __shared__ FrameTables ft;                                    // only filled on the overflow path
    __shared__ int s_any, s_ovf;

    if (errflags[CM3D_ERR_SEG_OVERFLOW] != 0) return;
    const int t = blockIdx.x;
    const int cnt = tile_cnt[t];
    const int32_t *sd = sweep_desc + (size_t)tile_sweep[t] * CM3D_SW_WORDS;
    const int32_t *fd = frame_desc + (size_t)sd[CM3D_SW_FRAME] * CM3D_FR_WORDS;
    const int ni = fd[CM3D_FR_NINST];
    const int fourth = sd[CM3D_SW_FOURTH];
    const int64_t base = (int64_t)t * kTile;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;

    // group g = r*8 + warp covers slots [g*32, g*32+32)
    uint32_t hw[kPerThread];
    bool any = false, ovf = false;
#pragma unroll
    for (int r = 0; r < kPerThread; ++r) {
        const int s = r * kBlock + threadIdx.x;
        hw[r] = s < cnt ? __ldg(hits + base + s) : 0u;
        any |= hw[r] != 0;
        ovf |= (hw[r] >> 24) == 0xffu;
    }
    __shared__ int s_np, s_pbase[kFastInst];
    __shared__ uint8_t s_plist[kFastInst];
    __shared__ uint16_t s_pc[kFastInst * kGroups];
    if (threadIdx.x == 0) { s_any = 0; s_ovf = 0; s_np = 0; }
    __syncthreads();
    if (any) s_any = 1;
    if (ovf) s_ovf = 1;
    __syncthreads();
    if (!s_any) return;

    // ---- fast path: few instances in this tile (points are in firing order, so a tile of 1024
    // consecutive points usually sees a handful) and no overflow word.  For every present instance:
    // one ballot per 32-slot group gives its members, an exclusive scan over the 32 groups their
    // ranks; members leave in slot order, so consecutive lanes write consecutive addresses.
    {
        const int tl0 = t - fd[CM3D_FR_TILE_BEGIN], ntf0 = fd[CM3D_FR_TILE_END] - fd[CM3D_FR_TILE_BEGIN];
        const int i0 = fd[CM3D_FR_INST_BEGIN];
        for (int j = threadIdx.x; j < ni; j += blockDim.x) {
            const int32_t *bj = tile_inst_base + (size_t)fd[CM3D_FR_CNT_OFF] + (size_t)j * ntf0;
            const int b0 = bj[tl0];
            const int b1 = tl0 + 1 < ntf0 ? bj[tl0 + 1] : seg_off[i0 + j + 1] - seg_off[i0 + j];
            if (b1 > b0) {
                const int k = atomicAdd(&s_np, 1);
                if (k < kFastInst) { s_plist[k] = (uint8_t)j; s_pbase[k] = seg_off[i0 + j] + b0; }
            }
        }
        __syncthreads();
        const int np = s_np;
        if (!s_ovf && np <= kFastInst) {
            float px[kPerThread], py[kPerThread], pz[kPerThread], pw[kPerThread];
#pragma unroll
            for (int r = 0; r < kPerThread; ++r) {
                px[r] = py[r] = pz[r] = pw[r] = 0.f;
                if (hw[r] != 0) {
                    const int64_t s = base + r * kBlock + threadIdx.x;
                    px[r] = xyzw[s]; py[r] = xyzw[n_slots + s]; pz[r] = xyzw[2 * n_slots + s];
                    if (fourth) pw[r] = xyzw[3 * n_slots + s];
                }
            }
            for (int k = 0; k < np; ++k) {
                const uint32_t pat = ((uint32_t)s_plist[k] + 1u) * 0x01010101u;
#pragma unroll
                for (int r = 0; r < kPerThread; ++r) {
                    const unsigned bal = __ballot_sync(0xffffffffu, __vcmpeq4(hw[r], pat) != 0u);
                    if (lane == 0) s_pc[k * kGroups + r * (kBlock / 32) + warp] = (uint16_t)__popc(bal);
                }
            }
            __syncthreads();
            for (int k = warp; k < np; k += kBlock / 32) {          // lane = group
                const int c = s_pc[k * kGroups + lane];
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= (unsigned)o) inc += v;
                }
                s_pc[k * kGroups + lane] = (uint16_t)(inc - c);
            }
            __syncthreads();
            float *sx = seg_xyzw, *sy = seg_xyzw + seg_cap, *sz = seg_xyzw + 2 * seg_cap, *sw = seg_xyzw + 3 * seg_cap;
            const int tp0 = tile_prefix[t];
            for (int k = 0; k < np; ++k) {
                const uint32_t pat = ((uint32_t)s_plist[k] + 1u) * 0x01010101u;
                const int kb = s_pbase[k];
#pragma unroll
                for (int r = 0; r < kPerThread; ++r) {
                    const bool mem = __vcmpeq4(hw[r], pat) != 0u;
                    const unsigned bal = __ballot_sync(0xffffffffu, mem);
                    if (mem) {
                        const int64_t dst = (int64_t)kb + s_pc[k * kGroups + r * (kBlock / 32) + warp] + __popc(bal & lanemask_lt());
                        if (dst < seg_cap) {
                            seg_point_idx[dst] = tp0 + r * kBlock + threadIdx.x;
                            sx[dst] = px[r]; sy[dst] = py[r]; sz[dst] = pz[r];
                            if (fourth) sw[dst] = pw[r];
                        }
                    }
                }
            }
            return;
        }
    }

    uint16_t *s_gc = reinterpret_cast<uint16_t *>(dyn_smem);
    int *s_base = reinterpret_cast<int *>(dyn_smem + ((kGroups * ni * 2 + 15) & ~15));
    for (int k = threadIdx.x; k < kGroups * ni; k += blockDim.x) s_gc[k] = 0;
    const int tl = t - fd[CM3D_FR_TILE_BEGIN], ntf = fd[CM3D_FR_TILE_END] - fd[CM3D_FR_TILE_BEGIN];
    for (int j = threadIdx.x; j < ni; j += blockDim.x)
        s_base[j] = seg_off[fd[CM3D_FR_INST_BEGIN] + j] +
                    tile_inst_base[(size_t)fd[CM3D_FR_CNT_OFF] + (size_t)j * ntf + tl];
    const bool have_ovf = s_ovf != 0;
    if (have_ovf) load_frame_tables(ft, fd, vcam_desc, cam_inst_list, inst_desc, inst_bbox, chains, bits, vcam_grid);
    __syncthreads();

    const float *gx = xyzw, *gy = xyzw + n_slots, *gz = xyzw + 2 * n_slots, *gw = xyzw + 3 * n_slots;
    float *sx = seg_xyzw, *sy = seg_xyzw + seg_cap, *sz = seg_xyzw + 2 * seg_cap, *sw = seg_xyzw + 3 * seg_cap;
    const int tp = tile_prefix[t];

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int r = 0; r < kPerThread; ++r) {
            const int s = r * kBlock + threadIdx.x;
            const int g = r * (kBlock / 32) + warp;
            if (!__any_sync(0xffffffffu, hw[r] != 0)) continue;
            float x = 0.f, y = 0.f, z = 0.f, w = 0.f;
            PendingHits ph;
            ph.word = hw[r];
            ph.ovf = (hw[r] >> 24) == 0xffu;
            if (hw[r] != 0 && (pass == 1 || ph.ovf)) {
                x = gx[base + s]; y = gy[base + s]; z = gz[base + s];
                if (fourth) w = gw[base + s];
            }
            if (have_ovf && ph.ovf) {
#pragma unroll
                for (int q = 0; q < 8; ++q) ph.bm[q] = 0u;
                for_each_hit(ft, x, y, z, nullptr, 0, [&](int j) { ph.bm[j >> 5] |= 1u << (j & 31); });
            }
            while (true) {
                const uint32_t mine = ph.front();                 // smallest pending id, 0 = none
                const uint32_t jmin = __reduce_min_sync(0xffffffffu, mine ? mine : 0xffffu);
                if (jmin == 0xffffu) break;
                const int j = (int)jmin - 1;
                const unsigned m = __ballot_sync(0xffffffffu, mine == jmin);
                if (pass == 0) {
                    if (lane == 0) s_gc[g * ni + j] = (uint16_t)__popc(m);
                } else if (mine == jmin) {
                    const int64_t dst = (int64_t)s_base[j] + s_gc[g * ni + j] + __popc(m & lanemask_lt());
                    if (dst < seg_cap) {
                        seg_point_idx[dst] = tp + s;
                        sx[dst] = x; sy[dst] = y; sz[dst] = z;
                        if (fourth) sw[dst] = w;
                    }
                }
                if (mine == jmin) ph.pop();
            }
        }
        if (pass == 0) {
            __syncthreads();
            // exclusive prefix over the groups, per instance
            for (int j = threadIdx.x; j < ni; j += blockDim.x) {
                int run = 0;
                for (int g = 0; g < kGroups; ++g) {
                    const int c = s_gc[g * ni + j];
                    s_gc[g * ni + j] = (uint16_t)run;
                    run += c;
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace cm3d

using namespace cm3d;

extern "C" int cm3d_abi_version(void) { return CM3D_ABI_VERSION; }

extern "C" const char *cm3d_error_string(int code)
{
    if (code == CM3D_OK) return "ok";
    if (code == CM3D_EINVAL) return "invalid argument";
    if (code == CM3D_ELIMIT) return "per-frame limit exceeded (instances > 254 or vcams > 16)";
    if (code <= -1000) return cudaGetErrorString((cudaError_t)(-code - 1000));
    return "unknown cm3d error";
}

extern "C" int cm3d_build_vcam_grid(const int32_t *vcam_desc, int n_vcams, int max_cells, const int32_t *frame_desc,
                                    const int32_t *cam_inst_list, const int32_t *inst_bbox, uint32_t *vcam_grid,
                                    void *stream)
{
    if (n_vcams < 0 || max_cells < 0) return CM3D_EINVAL;
    if (n_vcams == 0 || max_cells == 0) return CM3D_OK;
    if (!vcam_desc || !frame_desc || !cam_inst_list || !inst_bbox || !vcam_grid) return CM3D_EINVAL;
    dim3 grid((max_cells + 255) / 256, n_vcams);
    k_build_vcam_grid<<<grid, 256, 0, (cudaStream_t)stream>>>(vcam_desc, frame_desc, cam_inst_list, inst_bbox, vcam_grid);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}

extern "C" int cm3d_aggregate_sweeps(const float *raw, const int32_t *tile_sweep, int n_tiles,
                                     const int32_t *sweep_desc, const int32_t *frame_desc,
                                     const uint32_t *chains, float *xyzw, int32_t *tile_cnt, void *stream)
{
    if (n_tiles < 0) return CM3D_EINVAL;
    if (n_tiles == 0) return CM3D_OK;
    if (!raw || !tile_sweep || !sweep_desc || !frame_desc || !chains || !xyzw || !tile_cnt) return CM3D_EINVAL;
    if (((uintptr_t)raw & 15) != 0) return CM3D_EINVAL;
    k_aggregate<<<n_tiles, kBlock, 0, (cudaStream_t)stream>>>(raw, tile_sweep, sweep_desc, frame_desc, chains, xyzw,
                                                              (int64_t)n_tiles * kTile, tile_cnt);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}

extern "C" int cm3d_project_membership(const float *xyzw, const int32_t *tile_cnt, const int32_t *tile_sweep,
                                       int n_tiles, const int32_t *sweep_desc, const int32_t *frame_desc,
                                       const int32_t *vcam_desc, const int32_t *cam_inst_list,
                                       const int32_t *inst_desc, const int32_t *inst_bbox,
                                       const uint32_t *chains, const uint32_t *bits, const uint32_t *vcam_grid,
                                       uint32_t *hits, uint16_t *tile_inst_cnt, int32_t *pix, void *stream)
{
    if (n_tiles < 0) return CM3D_EINVAL;
    if (n_tiles == 0) return CM3D_OK;
    if (!xyzw || !tile_cnt || !tile_sweep || !sweep_desc || !frame_desc || !vcam_desc || !cam_inst_list ||
        !inst_desc || !inst_bbox || !chains || !bits || !vcam_grid || !hits || !tile_inst_cnt)
        return CM3D_EINVAL;
    k_project_count<<<n_tiles, kBlock, 0, (cudaStream_t)stream>>>(
        xyzw, (int64_t)n_tiles * kTile, tile_cnt, tile_sweep, sweep_desc, frame_desc, vcam_desc, cam_inst_list,
        inst_desc, inst_bbox, chains, bits, vcam_grid, hits, tile_inst_cnt, pix);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}

extern "C" int cm3d_scan_segments(const int32_t *tile_cnt, const uint16_t *tile_inst_cnt,
                                  const int32_t *frame_desc, int n_frames, int max_inst_per_frame,
                                  int n_inst_total, const int32_t *inst_desc, int64_t seg_cap,
                                  int32_t *tile_prefix, int32_t *frame_n, int32_t *tile_inst_base,
                                  int32_t *seg_off, int32_t *item_off, int32_t *item_inst,
                                  unsigned long long *medoid_best, int32_t *errflags, void *stream)
{
    if (n_frames < 0 || max_inst_per_frame < 0 || n_inst_total < 0 || seg_cap < 0) return CM3D_EINVAL;
    if (max_inst_per_frame > CM3D_MAX_INST) return CM3D_ELIMIT;
    if (n_frames == 0) return CM3D_OK;
    if (!tile_cnt || !tile_inst_cnt || !frame_desc || !tile_prefix || !frame_n || !tile_inst_base || !seg_off ||
        !item_off || !item_inst || !medoid_best || !errflags || (n_inst_total && !inst_desc))
        return CM3D_EINVAL;
    // seg_count is staged in seg_off[1..] (k_scan_batch reads element i before writing it)
    int32_t *seg_count = seg_off + 1;
    dim3 grid(n_frames, (max_inst_per_frame + 1 + 7) / 8);
    k_scan_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(tile_cnt, tile_inst_cnt, frame_desc, tile_prefix, frame_n,
                                                        tile_inst_base, seg_count);
    CM3D_LAUNCH_CHECK();
    k_scan_batch<<<1, 1024, 0, (cudaStream_t)stream>>>(seg_count, inst_desc, frame_desc, n_inst_total, seg_cap,
                                                       seg_off, item_off, item_inst, medoid_best, errflags);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}

extern "C" int cm3d_schedule_segments(const int32_t *frame_desc, const int32_t *inst_desc, int n_inst_total,
                                      int64_t seg_cap, int32_t *seg_off, int32_t *item_off, int32_t *item_inst,
                                      unsigned long long *medoid_best, int32_t *errflags, void *stream)
{
    if (n_inst_total < 0 || seg_cap < 0) return CM3D_EINVAL;
    if (!frame_desc || !seg_off || !item_off || !item_inst || !medoid_best || !errflags || (n_inst_total && !inst_desc))
        return CM3D_EINVAL;
    k_scan_batch<<<1, 1024, 0, (cudaStream_t)stream>>>(seg_off + 1, inst_desc, frame_desc, n_inst_total, seg_cap, seg_off,
                                                       item_off, item_inst, medoid_best, errflags);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}

extern "C" int cm3d_compact_segments(const float *xyzw, const int32_t *tile_cnt, const int32_t *tile_prefix,
                                     const int32_t *tile_sweep, int n_tiles, const int32_t *sweep_desc,
                                     const int32_t *frame_desc, const int32_t *vcam_desc,
                                     const int32_t *cam_inst_list, const int32_t *inst_desc,
                                     const int32_t *inst_bbox, const uint32_t *chains, const uint32_t *bits,
                                     const uint32_t *vcam_grid, const uint32_t *hits,
                                     const int32_t *tile_inst_base,
                                     const int32_t *seg_off, int32_t *seg_point_idx, float *seg_xyzw,
                                     int64_t seg_cap, int max_inst_per_frame, const int32_t *errflags,
                                     void *stream)
{
    if (n_tiles < 0 || seg_cap < 0 || max_inst_per_frame < 0) return CM3D_EINVAL;
    if (max_inst_per_frame > CM3D_MAX_INST) return CM3D_ELIMIT;
    if (n_tiles == 0) return CM3D_OK;
    if (!xyzw || !tile_cnt || !tile_prefix || !tile_sweep || !sweep_desc || !frame_desc || !vcam_desc ||
        !cam_inst_list || !inst_desc || !inst_bbox || !chains || !bits || !vcam_grid || !hits || !tile_inst_base ||
        !seg_off ||
        !seg_point_idx || !seg_xyzw || !errflags)
        return CM3D_EINVAL;
    const size_t smem = ((kGroups * max_inst_per_frame * 2 + 15) & ~15) + max_inst_per_frame * sizeof(int) + 16;
    k_compact<<<n_tiles, kBlock, smem, (cudaStream_t)stream>>>(
        xyzw, (int64_t)n_tiles * kTile, tile_cnt, tile_prefix, tile_sweep, sweep_desc, frame_desc, vcam_desc,
        cam_inst_list, inst_desc, inst_bbox, chains, bits, vcam_grid, hits, tile_inst_base, seg_off, seg_point_idx,
        seg_xyzw, seg_cap, errflags);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}
