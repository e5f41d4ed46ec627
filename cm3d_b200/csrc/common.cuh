// Shared device helpers for the cm3d_b200 kernels (sm_100a).
//
// Numerics rule for the whole directory: every floating-point operation that
// feeds a pixel index, a membership decision or a medoid sum is written with an
// explicit IEEE round-to-nearest intrinsic (__fmul_rn, __fmaf_rn, __fadd_rn,
// __fdiv_rn, __fsqrt_rn).  nvcc never contracts or reorders those, so the
// kernels reproduce torch-CPU's k-ordered FMA chains bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cm3d_b200.h"

#define CM3D_LAUNCH_CHECK()                                   \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return -(1000 + (int)e__);    \
    } while (0)

namespace cm3d {

constexpr int kTile = CM3D_TILE;
constexpr int kBlock = 256;               // threads per block in the point kernels
constexpr int kPerThread = kTile / kBlock;  // points per thread per tile

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// One transform chain from shared memory (CM3D_CHAIN_WORDS words, uniform address
// -> broadcast LDS).  Mirrors pcd.py translate/rotate and kitti_utils cart2hom+matmul.
__device__ __forceinline__ void apply_chain(const uint32_t *__restrict__ chain, float &x, float &y, float &z)
{
#pragma unroll
    for (int k = 0; k < CM3D_MAX_CHAIN; ++k) {
        const uint32_t *op = chain + k * CM3D_OP_WORDS;
        const uint32_t kind = op[0];
        if (kind == CM3D_OP_END) break;
        const float *m = reinterpret_cast<const float *>(op + 1);
        const float a = x, b = y, c = z;
        if (kind == CM3D_OP_T) {
            x = __fadd_rn(a, m[0]);
            y = __fadd_rn(b, m[1]);
            z = __fadd_rn(c, m[2]);
        } else if (kind == CM3D_OP_R) {
            x = __fmaf_rn(m[2], c, __fmaf_rn(m[1], b, __fmul_rn(m[0], a)));
            y = __fmaf_rn(m[5], c, __fmaf_rn(m[4], b, __fmul_rn(m[3], a)));
            z = __fmaf_rn(m[8], c, __fmaf_rn(m[7], b, __fmul_rn(m[6], a)));
        } else {  // CM3D_OP_A: [a b c 1] . row
            x = __fmaf_rn(1.0f, m[3], __fmaf_rn(c, m[2], __fmaf_rn(b, m[1], __fmul_rn(a, m[0]))));
            y = __fmaf_rn(1.0f, m[7], __fmaf_rn(c, m[6], __fmaf_rn(b, m[5], __fmul_rn(a, m[4]))));
            z = __fmaf_rn(1.0f, m[11], __fmaf_rn(c, m[10], __fmaf_rn(b, m[9], __fmul_rn(a, m[8]))));
        }
    }
}

// ------------------------------------------------------------------ medoid work items
constexpr int kCols = CM3D_MEDOID_COLS;   // columns of the distance matrix per full medoid item
constexpr int kSmallM = 32;               // instances below this size are one generic item

// Work items of an instance with m members (0 when it has fewer than min_pts):
//   m <  kSmallM : one generic item (every column, normal and tail, in one block);
//   m >= kSmallM : ceil(full/kCols) items over the `full` = (m/32)*32 single-accumulator columns,
//                  plus one tail item for the last m%32 columns (four threads per column).
__host__ __device__ inline int medoid_items(int m, int min_pts)
{
    if (m < (min_pts > 1 ? min_pts : 1)) return 0;
    if (m < kSmallM) return 1;
    const int full = (m / 32) * 32;
    return (full + kCols - 1) / kCols + ((m & 31) ? 1 : 0);
}

__device__ __forceinline__ int64_t join64(int32_t lo, int32_t hi)
{
    return (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
}

}  // namespace cm3d

// ------------------------------------------------------------------ bulk async copy (TMA 1-D) + mbarrier
namespace cm3d {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    // make the init visible to the async proxy before any bulk copy can signal it
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// global -> shared bulk copy (SASS: UBLKCP); dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace cm3d
