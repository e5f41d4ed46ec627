// Micro-benchmark of the issue/pipe rates the medoid kernel depends on (B200, sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench pipe_bench.cu && ./pipe_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITER 4096
#define UN 8
template <int MODE>
__global__ void k(float *out, float a, float b)
{
    float x[UN], y[UN];
    u64 p[UN];
    for (int i = 0; i < UN; ++i) { x[i] = a + i + threadIdx.x; y[i] = b + i; asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(x[i]), "f"(y[i])); }
    u64 bb; asm("mov.b64 %0, {%1,%2};" : "=l"(bb) : "f"(b), "f"(b));
    u64 cc; asm("mov.b64 %0, {%1,%2};" : "=l"(cc) : "f"(a), "f"(a));
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            if (MODE == 0) x[i] = __fmaf_rn(x[i], b, a);                       // FFMA 3-reg
            if (MODE == 1) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));   // FFMA2
            if (MODE == 2) x[i] = fmaxf(x[i], y[i]) ;                          // FMNMX (dependent on y)
            if (MODE == 3) asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));    // MUFU.RSQ
            if (MODE == 4) { x[i] = __fmaf_rn(x[i], b, a); asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(y[i])); }  // 1 FFMA + 1 MUFU
            if (MODE == 5) { asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc)); x[i] = fmaxf(x[i], y[i]); y[i] = fminf(y[i], a); }
            if (MODE == 6) { x[i] = __fmaf_rn(x[i], b, a); y[i] = __fmaf_rn(y[i], a, b); }  // 2 independent FFMA
            if (MODE == 7) { x[i] = __fadd_rn(x[i], b); }                       // FADD
            if (MODE == 8) asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(bb));   // FMUL2
            if (MODE == 9) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));   // FADD2
        }
    }
    float s = 0; for (int i = 0; i < UN; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += x[i] + y[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ u64 pk(float lo, float hi){ u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi){ asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// the medoid inner step on register/shared data: VAR 0 = full, 1 = no MUFU (y=r), 2 = no FMNMX, 3 = no acc
template <int VAR> __device__ __forceinline__ u64 xmul2(u64 a, u64 b){ if (VAR == 4) return fma2(a, b, pk(-0.0f, -0.0f)); return mul2(a, b); }
template <int VAR> __device__ __forceinline__ u64 xadd2(u64 a, u64 b){ if (VAR == 4) return fma2(a, pk(1.0f, 1.0f), b); return add2(a, b); }
template <int VAR>
__global__ void kmed(float *out, const float4 *rows_g, float xj, float yj, float zj, float nj)
{
    __shared__ float4 s_rows[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_rows[i] = rows_g[i];
    __syncthreads();
    const u64 xj2 = pk(xj + threadIdx.x, xj + threadIdx.x), yj2 = pk(yj, yj), zj2 = pk(zj, zj), nj2 = pk(nj, nj);
    float a0 = 0.f;
    for (int it = 0; it < 64; ++it) {
        for (int b = 0; b < 1024; b += 16) {
            float nd[16];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a = s_rows[b + 2 * q], bb = s_rows[b + 2 * q + 1];
                u64 r = xmul2<VAR>(pk(a.x, a.y), xj2);
                r = fma2(pk(a.z, a.w), yj2, r);
                r = fma2(pk(bb.x, bb.y), zj2, r);
                r = xadd2<VAR>(pk(bb.z, bb.w), r);
                r = xadd2<VAR>(nj2, r);
                float r0, r1, y0, y1;
                upk(r, r0, r1);
                float nx0, nx1;
                if (VAR == 2) { nx0 = -r0; nx1 = -r1; } else { nx0 = fminf(-r0, -0.0f); nx1 = fminf(-r1, -0.0f); }
                if (VAR == 1) { y0 = r0; y1 = r1; }
                else if (VAR == 2) { asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(r0)); asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(r1)); }
                else { asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(fmaxf(r0, 0x1p-101f))); asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(fmaxf(r1, 0x1p-101f))); }
                const u64 nx = pk(nx0, nx1), y = pk(y0, y1);
                const u64 ns = xmul2<VAR>(nx, y);
                const u64 h = xmul2<VAR>(y, pk(0.5f, 0.5f));
                const u64 rr = fma2(ns, ns, nx);
                upk(fma2(rr, h, ns), nd[2 * q], nd[2 * q + 1]);
            }
            if (VAR == 3) { float t = 0; 
#pragma unroll
                for (int q = 0; q < 16; ++q) t = fmaxf(t, nd[q]); a0 += t; }
            else {
#pragma unroll
            for (int q = 0; q < 16; ++q) a0 = __fsub_rn(a0, nd[q]); }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0;
}
template <int VAR> void runmed(const char *name, int blocks_per_sm, int threads)
{
    float *d; cudaMalloc(&d, 148 * 16 * 1024 * 4);
    float4 *rows; cudaMalloc(&rows, 1024 * 16);
    float4 *h = new float4[1024];
    for (int i = 0; i < 1024; ++i) h[i] = make_float4(-2.f * (1200 + i * 0.01f), -2.f * (1200 + (i + 1) * 0.01f), -2.f * 950.f, -2.f * 951.f);
    cudaMemcpy(rows, h, 1024 * 16, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kmed<VAR><<<148 * blocks_per_sm, threads>>>(d, rows, 1200.f, 950.f, 1.f, 2342501.f);
    cudaEventRecord(e0);
    kmed<VAR><<<148 * blocks_per_sm, threads>>>(d, rows, 1200.f, 950.f, 1.f, 2342501.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9;
    double warp_rows_per_smsp = (blocks_per_sm * threads / 32 / 4.0) * 64.0 * 1024;
    printf("%-34s %d blk/SM x %d thr: %.3f ms -> %.2f cycles per warp-row per SMSP\n", name, blocks_per_sm, threads, ms, cyc / warp_rows_per_smsp);
    cudaFree(d); cudaFree(rows);
}
template <int MODE> void run(const char *name, double ops_per_iter)
{
    float *d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 512>>>(d, 1.0001f, 0.9999f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(d, 1.0001f, 0.9999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // warps per SMSP = 4 blocks*16 warps / 4 = 16; instr per warp = ITER*UN*ops
    double cyc = ms * 1e-3 * 1.965e9;
    double instr_per_smsp = 16.0 * ITER * UN * ops_per_iter;
    printf("%-28s %.3f ms  -> %.2f cycles per warp-instruction per SMSP (at 1965 MHz)\n", name, ms, cyc / instr_per_smsp);
    cudaFree(d);
}
int main()
{
    run<0>("FFMA (3 reg)", 1); run<1>("FFMA2 (packed)", 1); run<2>("FMNMX", 1); run<3>("MUFU.RSQ", 1);
    run<4>("FFMA + MUFU.RSQ (pair)", 2); run<5>("FFMA2 + 2 FMNMX (triple)", 3); run<6>("2 indep FFMA", 2); run<7>("FADD", 1); run<8>("FMUL2", 1); run<9>("FADD2", 1);
    runmed<0>("medoid step full", 3, 256); runmed<0>("medoid step full", 4, 256); runmed<0>("medoid step full", 8, 128); runmed<0>("medoid step full", 2, 256);
    runmed<1>("medoid step, no MUFU", 4, 256); runmed<2>("medoid step, no FMNMX", 4, 256); runmed<3>("medoid step, no acc chain", 4, 256); runmed<4>("medoid step, all FFMA2", 4, 256);
    return 0;
}
