// Issue-slot model probes (B200): does FFMA2 / MUFU occupy the dispatch port for more than one cycle?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench3 pipe_bench3.cu && ./pipe_bench3
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITER 2048
__device__ __forceinline__ u64 pk(float lo, float hi){ u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi){ asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ int lop(int a, int b){ int r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ float rsq(float a){ float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float fmx(float a, float b){ float r; asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int NF2, int NF, int NL, int NM, int NX>
__global__ void k(float *out, float a, float b, int ia)
{
    u64 p[8]; float x[8]; int n[8]; float m[4]; float g[4];
    for (int i = 0; i < 8; ++i) { p[i] = pk(a + i + threadIdx.x, b + i); x[i] = a + i; n[i] = ia + i; }
    for (int i = 0; i < 4; ++i) { m[i] = a + i; g[i] = b + i; }
    const u64 bb = pk(b, b), cc = pk(a, a);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < NF2) p[i] = fma2(p[i], bb, cc);
                if (i < NF) x[i] = ffma(x[i], b, a);
                if (i < NL) n[i] = lop(n[i], ia);
                if (i < NM) m[i & 3] = rsq(m[i & 3]);
                if (i < NX) g[i & 3] = fmx(g[i & 3], x[(i + 1) & 7]);
            }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; upk(p[i], lo, hi); s += x[i] + lo + hi + n[i] + m[i & 3] + g[i & 3]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NF2, int NF, int NL, int NM, int NX> void run()
{
    float *d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NF2, NF, NL, NM, NX><<<148 * 4, 512>>>(d, 1.0001f, 0.9999f, 3);
    cudaEventRecord(e0);
    k<NF2, NF, NL, NM, NX><<<148 * 4, 512>>>(d, 1.0001f, 0.9999f, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9 / (16.0 * ITER * 4);
    printf("FFMA2 x%d FFMA x%d LOP x%d MUFU x%d FMNMX x%d : %.2f cycles/group | fma pipe %d, alu %d, xu %d, instr %d\n",
           NF2, NF, NL, NM, NX, cyc, 2 * NF2 + NF, 2 * (NL + NX), 8 * NM, NF2 + NF + NL + NM + NX);
    cudaFree(d);
}
int main()
{
    run<4, 0, 0, 0, 0>(); run<0, 8, 0, 0, 0>(); run<0, 0, 4, 0, 0>(); run<0, 0, 0, 1, 0>(); run<0, 0, 0, 0, 4>();
    run<4, 0, 4, 0, 0>(); run<4, 0, 2, 0, 0>(); run<0, 8, 2, 0, 0>(); run<0, 8, 4, 0, 0>();
    run<0, 6, 0, 1, 0>(); run<0, 7, 0, 1, 0>(); run<0, 7, 1, 1, 0>(); run<3, 0, 2, 1, 0>(); run<4, 0, 0, 1, 0>();
    run<4, 1, 0, 1, 1>(); run<4, 2, 0, 1, 1>(); run<5, 0, 0, 1, 1>(); run<0, 10, 0, 1, 1>(); run<4, 2, 1, 1, 1>();
    run<4, 0, 0, 0, 2>(); run<4, 0, 0, 0, 4>(); run<0, 8, 0, 0, 4>(); run<2, 4, 0, 1, 1>();
    return 0;
}
