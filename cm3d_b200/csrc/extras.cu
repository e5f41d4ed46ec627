// Default-OFF extensions named by the north star but never executed by the reference
// (SURVEY 0.4 / 8f-4): per-instance outlier filtering by neighbour counting, and a block-level
// orientation / extent search for the 3D box.  PARITY UNPINNED against the reference (its
// clustering code, src/kitti/2d_to_3d.py:159-174 `clusters_hdbscan`, is dead: the import is
// commented out; its box is a shape prior + lane yaw); graded against oracle/extras_oracle.py.
//
//   k_neighbor_count    for every member point of every instance: number of members of the SAME
//                       instance within `radius` (itself included).  Shared-memory tiled all-pairs
//                       like the medoid: a work item is (instance, 256 columns), rows stream through
//                       shared memory.  d2 = (dx*dx + dy*dy) + dz*dz with rounded fp32 operations,
//                       compared with fl(radius*radius): bit-exact against numpy float32.
//   k_filter_segments   ordered compaction of the kept points into new instance segments.
//   k_box_search        one block per instance: for n_angles headings in [0, pi/2) the extent of the
//                       points along the rotated ground-plane axes (warp-shuffle min/max
//                       reductions); the heading with the smallest footprint area wins.
#include "common.cuh"

namespace cm3d {

constexpr int kNbTile = 1024;

__global__ void __launch_bounds__(256)
k_neighbor_count(const float *__restrict__ seg_xyzw, int64_t seg_cap, const int32_t *__restrict__ seg_off,
                 const int32_t *__restrict__ item_first /* exclusive prefix of ceil(M/256) */, int n_inst, float r2,
                 int min_neighbors, uint8_t *__restrict__ keep, int32_t *__restrict__ kept_count)
{
    __shared__ float4 s_rows[kNbTile];
    __shared__ int s_kept;
    const int item = blockIdx.x;
    if (item >= item_first[n_inst]) return;
    int lo = 0, hi = n_inst;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (item_first[mid] <= item) lo = mid; else hi = mid;
    }
    const int inst = lo;
    const int o = seg_off[inst], m = seg_off[inst + 1] - o;
    const int j = (item - item_first[inst]) * 256 + threadIdx.x;
    const float *sx = seg_xyzw + o, *sy = seg_xyzw + seg_cap + o, *sz = seg_xyzw + 2 * seg_cap + o;
    const bool valid = j < m;
    const float xj = valid ? sx[j] : 0.f, yj = valid ? sy[j] : 0.f, zj = valid ? sz[j] : 0.f;
    int cnt = 0;
    if (threadIdx.x == 0) s_kept = 0;
    for (int t0 = 0; t0 < m; t0 += kNbTile) {
        const int rows = min(kNbTile, m - t0);
        __syncthreads();
        for (int r = threadIdx.x; r < rows; r += blockDim.x) s_rows[r] = make_float4(sx[t0 + r], sy[t0 + r], sz[t0 + r], 0.f);
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
            const float4 p = s_rows[r];
            const float dx = __fsub_rn(p.x, xj), dy = __fsub_rn(p.y, yj), dz = __fsub_rn(p.z, zj);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            cnt += d2 <= r2;
        }
    }
    const bool k = valid && cnt >= min_neighbors;
    if (valid) keep[o + j] = k ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (lane_id() == 0 && bal) atomicAdd(&s_kept, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0 && s_kept) atomicAdd(kept_count + inst, s_kept);
}

// item_first for k_neighbor_count (one block; exclusive prefix of ceil(M/256)) and zeroed kept counters
__global__ void __launch_bounds__(1024)
k_neighbor_items(const int32_t *__restrict__ seg_off, int n_inst, int32_t *__restrict__ item_first,
                 int32_t *__restrict__ kept_count)
{
    __shared__ int s_w[32];
    __shared__ int s_c;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    for (int base = 0; base < n_inst; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int a = i < n_inst ? (seg_off[i + 1] - seg_off[i] + 255) / 256 : 0;
        if (i < n_inst) kept_count[i] = 0;
        int inc = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (unsigned)d) inc += u;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        int off = s_c;
        for (unsigned w = 0; w < warp; ++w) off += s_w[w];
        if (i < n_inst) item_first[i] = off + inc - a;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_c = off + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) item_first[n_inst] = s_c;
}

// One block per instance: kept points keep their (ascending point index) order.
__global__ void __launch_bounds__(256)
k_filter_segments(const float *__restrict__ seg_xyzw, const int32_t *__restrict__ seg_point_idx, int64_t seg_cap,
                  const int32_t *__restrict__ seg_off, const uint8_t *__restrict__ keep,
                  const int32_t *__restrict__ seg_off2, float *__restrict__ seg_xyzw2,
                  int32_t *__restrict__ seg_point_idx2, const int32_t *__restrict__ errflags)
{
    __shared__ int s_w[8];
    __shared__ int s_c;
    if (errflags[CM3D_ERR_SEG_OVERFLOW] != 0) return;
    const int inst = blockIdx.x;
    const int o = seg_off[inst], m = seg_off[inst + 1] - o, o2 = seg_off2[inst];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    for (int base = 0; base < m; base += blockDim.x) {
        const int j = base + threadIdx.x;
        const bool k = j < m && keep[o + j] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        if (lane == 0) s_w[warp] = __popc(bal);
        __syncthreads();
        int off = s_c + __popc(bal & lanemask_lt());
        for (unsigned w = 0; w < warp; ++w) off += s_w[w];
        if (k) {
            const int64_t d = (int64_t)o2 + off, s = (int64_t)o + j;
            seg_point_idx2[d] = seg_point_idx[s];
#pragma unroll
            for (int c = 0; c < 4; ++c) seg_xyzw2[c * seg_cap + d] = seg_xyzw[c * seg_cap + s];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += s_w[w];
            s_c += t;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ orientation / extent search
struct MinMax4 { float umin, umax, vmin, vmax; };

__device__ __forceinline__ MinMax4 block_minmax4(MinMax4 v, MinMax4 *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.umin = fminf(v.umin, __shfl_xor_sync(0xffffffffu, v.umin, o));
        v.umax = fmaxf(v.umax, __shfl_xor_sync(0xffffffffu, v.umax, o));
        v.vmin = fminf(v.vmin, __shfl_xor_sync(0xffffffffu, v.vmin, o));
        v.vmax = fmaxf(v.vmax, __shfl_xor_sync(0xffffffffu, v.vmax, o));
    }
    __syncthreads();
    if (lane_id() == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    MinMax4 t = s_red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        t.umin = fminf(t.umin, s_red[w].umin); t.umax = fmaxf(t.umax, s_red[w].umax);
        t.vmin = fminf(t.vmin, s_red[w].vmin); t.vmax = fmaxf(t.vmax, s_red[w].vmax);
    }
    return t;
}

// box[8*i..] = centre (in the segment's frame, 3 floats), extent along the heading / across it /
// up, heading theta in [0, pi/2) about the up axis, footprint area.  Ground plane = the two axes
// other than `up_axis`, taken in cyclic order (up=2: (x,y); up=1: (z,x); up=0: (y,z)).
// Every warp takes headings warp, warp+8, ...: points are read from shared memory when the
// instance fits (<= 4096 points), else from L2.
constexpr int kBoxSmemPts = 4096;

__global__ void __launch_bounds__(256)
k_box_search(const float *__restrict__ seg_xyzw, int64_t seg_cap, const int32_t *__restrict__ seg_off, int up_axis,
             int n_angles, int min_pts, float *__restrict__ box, const int32_t *__restrict__ errflags)
{
    __shared__ float s_a[kBoxSmemPts], s_b[kBoxSmemPts];
    __shared__ float s_area[8], s_theta[8], s_ext[8][4];
    __shared__ MinMax4 s_red[8];
    const int inst = blockIdx.x;
    float *out = box + (size_t)inst * 8;
    const int o = seg_off[inst], m = seg_off[inst + 1] - o;
    const float nanv = __int_as_float(0x7fc00000);
    if (errflags[CM3D_ERR_SEG_OVERFLOW] != 0 || m < max(min_pts, 1)) {
        if (threadIdx.x < 8) out[threadIdx.x] = nanv;
        return;
    }
    const int ia = (up_axis + 1) % 3, ib = (up_axis + 2) % 3;
    const float *pa = seg_xyzw + (int64_t)ia * seg_cap + o, *pb = seg_xyzw + (int64_t)ib * seg_cap + o;
    const float *pu = seg_xyzw + (int64_t)up_axis * seg_cap + o;
    const bool in_smem = m <= kBoxSmemPts;
    if (in_smem)
        for (int r = threadIdx.x; r < m; r += blockDim.x) { s_a[r] = pa[r]; s_b[r] = pb[r]; }
    // vertical extent (block reduction, reusing the 4-wide helper)
    MinMax4 z = {3.4e38f, -3.4e38f, 0.f, 0.f};
    for (int r = threadIdx.x; r < m; r += blockDim.x) { const float w = pu[r]; z.umin = fminf(z.umin, w); z.umax = fmaxf(z.umax, w); }
    z = block_minmax4(z, s_red);
    __syncthreads();
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    float best_area = 3.4e38f, best_theta = 0.f;
    MinMax4 best = {0.f, 0.f, 0.f, 0.f};
    for (int k = warp; k < n_angles; k += 8) {
        const float theta = (float)k * (1.57079632679489662f / (float)n_angles);
        float sn, cs;
        sincosf(theta, &sn, &cs);
        MinMax4 e = {3.4e38f, -3.4e38f, 3.4e38f, -3.4e38f};
        for (int r = lane; r < m; r += 32) {
            const float a = in_smem ? s_a[r] : pa[r], b = in_smem ? s_b[r] : pb[r];
            const float u = a * cs + b * sn, v = b * cs - a * sn;
            e.umin = fminf(e.umin, u); e.umax = fmaxf(e.umax, u);
            e.vmin = fminf(e.vmin, v); e.vmax = fmaxf(e.vmax, v);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            e.umin = fminf(e.umin, __shfl_xor_sync(0xffffffffu, e.umin, d));
            e.umax = fmaxf(e.umax, __shfl_xor_sync(0xffffffffu, e.umax, d));
            e.vmin = fminf(e.vmin, __shfl_xor_sync(0xffffffffu, e.vmin, d));
            e.vmax = fmaxf(e.vmax, __shfl_xor_sync(0xffffffffu, e.vmax, d));
        }
        const float area = (e.umax - e.umin) * (e.vmax - e.vmin);
        if (area < best_area) { best_area = area; best_theta = theta; best = e; }     // first minimum per warp
    }
    if (lane == 0) {
        s_area[warp] = best_area; s_theta[warp] = best_theta;
        s_ext[warp][0] = best.umin; s_ext[warp][1] = best.umax; s_ext[warp][2] = best.vmin; s_ext[warp][3] = best.vmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int bw = 0;
        for (int w = 1; w < 8; ++w)       // smallest area, then smallest heading (= first minimum over k)
            if (s_area[w] < s_area[bw] || (s_area[w] == s_area[bw] && s_theta[w] < s_theta[bw])) bw = w;
        const float th = s_theta[bw];
        float sn, cs;
        sincosf(th, &sn, &cs);
        const float uc = 0.5f * (s_ext[bw][0] + s_ext[bw][1]), vc = 0.5f * (s_ext[bw][2] + s_ext[bw][3]);
        float c[3];
        c[ia] = uc * cs - vc * sn;
        c[ib] = uc * sn + vc * cs;
        c[up_axis] = 0.5f * (z.umin + z.umax);
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
        out[3] = s_ext[bw][1] - s_ext[bw][0];
        out[4] = s_ext[bw][3] - s_ext[bw][2];
        out[5] = z.umax - z.umin;
        out[6] = th;
        out[7] = s_area[bw];
    }
}

}  // namespace cm3d

using namespace cm3d;

extern "C" int cm3d_neighbor_filter(const float *seg_xyzw, int64_t seg_cap, const int32_t *seg_off, int n_inst_total,
                                    int max_items, float radius, int min_neighbors, int32_t *item_first,
                                    uint8_t *keep, int32_t *kept_count, void *stream)
{
    if (n_inst_total < 0 || seg_cap < 0 || max_items < 0 || !(radius >= 0.0f)) return CM3D_EINVAL;
    if (n_inst_total == 0) return CM3D_OK;
    if (!seg_xyzw || !seg_off || !item_first || !keep || !kept_count) return CM3D_EINVAL;
    k_neighbor_items<<<1, 1024, 0, (cudaStream_t)stream>>>(seg_off, n_inst_total, item_first, kept_count);
    CM3D_LAUNCH_CHECK();
    if (max_items > 0) {
        const float r2 = radius * radius;      // fl(r*r), like numpy float32
        k_neighbor_count<<<max_items, 256, 0, (cudaStream_t)stream>>>(seg_xyzw, seg_cap, seg_off, item_first, n_inst_total,
                                                                      r2, min_neighbors, keep, kept_count);
        CM3D_LAUNCH_CHECK();
    }
    return CM3D_OK;
}

extern "C" int cm3d_filter_segments(const float *seg_xyzw, const int32_t *seg_point_idx, int64_t seg_cap,
                                    const int32_t *seg_off, const uint8_t *keep, const int32_t *seg_off2,
                                    int n_inst_total, float *seg_xyzw2, int32_t *seg_point_idx2,
                                    const int32_t *errflags, void *stream)
{
    if (n_inst_total < 0 || seg_cap < 0) return CM3D_EINVAL;
    if (n_inst_total == 0) return CM3D_OK;
    if (!seg_xyzw || !seg_point_idx || !seg_off || !keep || !seg_off2 || !seg_xyzw2 || !seg_point_idx2 || !errflags)
        return CM3D_EINVAL;
    k_filter_segments<<<n_inst_total, 256, 0, (cudaStream_t)stream>>>(seg_xyzw, seg_point_idx, seg_cap, seg_off, keep,
                                                                      seg_off2, seg_xyzw2, seg_point_idx2, errflags);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}

extern "C" int cm3d_box_search(const float *seg_xyzw, int64_t seg_cap, const int32_t *seg_off, int n_inst_total,
                               int up_axis, int n_angles, int min_pts, float *box, const int32_t *errflags,
                               void *stream)
{
    if (n_inst_total < 0 || seg_cap < 0 || up_axis < 0 || up_axis > 2 || n_angles < 1) return CM3D_EINVAL;
    if (n_inst_total == 0) return CM3D_OK;
    if (!seg_xyzw || !seg_off || !box || !errflags) return CM3D_EINVAL;
    k_box_search<<<n_inst_total, 256, 0, (cudaStream_t)stream>>>(seg_xyzw, seg_cap, seg_off, up_axis, n_angles, min_pts,
                                                                 box, errflags);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}
