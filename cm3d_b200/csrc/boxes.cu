// Pass-2 helper on the GPU: nearest discretised lane point of every centroid.
//
// Replaces `scipy.spatial.distance.cdist(all_centroids[:, :2], all_lane_pts[:, :2])` followed by
// argmin / min over the lane axis (src/nuscenes/2d_to_3d.py:277-302, src/waymo/2d_to_3d.py:
// lane_yaws_distances_and_coords).  scipy computes sqrt(dx*dx + dy*dy) in binary64 with no FMA;
// the same three rounded operations are used here, and ties go to the first lane point like
// numpy.argmin, so index and distance are bit-identical.  The n x m matrix is never stored.
#include "common.cuh"

namespace cm3d {

__global__ void __launch_bounds__(256)
k_nearest_lane(const double *__restrict__ cxy, int n, const double *__restrict__ lxy, int m,
               int32_t *__restrict__ idx_out, double *__restrict__ dist_out)
{
    __shared__ double s_d[8];
    __shared__ int s_i[8];
    const int c = blockIdx.x;
    if (c >= n) return;
    const double x = cxy[2 * c], y = cxy[2 * c + 1];
    double best_s = 1.0 / 0.0, best_d = 1.0 / 0.0;   // +inf
    int best_i = 0x7fffffff;
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const double2 p = reinterpret_cast<const double2 *>(lxy)[k];
        const double dx = __dsub_rn(x, p.x), dy = __dsub_rn(y, p.y);
        const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        // sqrt is monotone: s >= best_s cannot give a strictly smaller distance
        if (s < best_s || best_i == 0x7fffffff) {
            const double d = __dsqrt_rn(s);
            if (d < best_d || best_i == 0x7fffffff) { best_d = d; best_i = k; }
            best_s = s;
        }
    }
    // lexicographic (distance, index) minimum; NaN distances never win against a number
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best_d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (od < best_d || (od == best_d && oi < best_i) || (best_i == 0x7fffffff && oi != 0x7fffffff)) { best_d = od; best_i = oi; }
    }
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (lane == 0) { s_d[warp] = best_d; s_i[warp] = best_i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            const double od = s_d[w];
            const int oi = s_i[w];
            if (od < best_d || (od == best_d && oi < best_i) || (best_i == 0x7fffffff && oi != 0x7fffffff)) { best_d = od; best_i = oi; }
        }
        idx_out[c] = best_i == 0x7fffffff ? -1 : best_i;
        dist_out[c] = best_d;
    }
}

}  // namespace cm3d

using namespace cm3d;

extern "C" int cm3d_nearest_lane(const double *centroids_xy, int n, const double *lane_xy, int m,
                                 int32_t *idx_out, double *dist_out, void *stream)
{
    if (n < 0 || m < 0) return CM3D_EINVAL;
    if (n == 0) return CM3D_OK;
    if (!centroids_xy || !idx_out || !dist_out || (m && !lane_xy)) return CM3D_EINVAL;
    if (((uintptr_t)lane_xy & 15) != 0) return CM3D_EINVAL;
    k_nearest_lane<<<n, 256, 0, (cudaStream_t)stream>>>(centroids_xy, n, lane_xy, m, idx_out, dist_out);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}
