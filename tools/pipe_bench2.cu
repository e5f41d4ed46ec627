// Second micro-benchmark for the medoid inner step (B200, sm_100a): does FFMA2 leave issue
// slots free for other pipes, and how many cycles do candidate step variants take?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench2 pipe_bench2.cu && ./pipe_bench2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITER 4096
#define UN 8
__device__ __forceinline__ u64 pk(float lo, float hi){ u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi){ asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void k(float *out, float a, float b, int ia)
{
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = a + i;
    __syncthreads();
    float x[UN], y[UN];
    u64 p[UN];
    int n[UN];
    for (int i = 0; i < UN; ++i) { x[i] = a + i + threadIdx.x; y[i] = b + i; p[i] = pk(x[i], y[i]); n[i] = ia + i; }
    const u64 bb = pk(b, b), cc = pk(a, a);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            if (MODE == 0) { p[i] = fma2(p[i], bb, cc); n[i] = n[i] * 3 + ia; }                  // FFMA2 + IMAD
            if (MODE == 1) { x[i] = __fmaf_rn(x[i], b, a); n[i] = n[i] * 3 + ia; }               // FFMA + IMAD
            if (MODE == 2) { p[i] = fma2(p[i], bb, cc); n[i] = (n[i] ^ ia) + it; }               // FFMA2 + LOP3/IADD
            if (MODE == 3) { p[i] = fma2(p[i], bb, cc); y[i] = fmaxf(y[i] + 0.0f, x[i]); x[i] = fminf(x[i], y[i]) ; }  // FFMA2 + mixed
            if (MODE == 4) { p[i] = fma2(p[i], bb, cc); asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }   // FFMA2 + MUFU
            if (MODE == 5) { p[i] = fma2(p[i], bb, cc); p[(i + 1) % UN] = fma2(p[(i + 1) % UN], cc, bb); asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }  // 2 FFMA2 + MUFU
            if (MODE == 6) { p[i] = fma2(p[i], bb, cc); x[i] += sm[(it + i * 32 + threadIdx.x) & 1023]; }  // FFMA2 + LDS + FADD
            if (MODE == 7) { n[i] = (n[i] ^ ia) + it; }                                          // int only
            if (MODE == 8) { p[i] = fma2(p[i], bb, cc); x[i] = fmaxf(x[i], __int_as_float(n[i])); n[i] += it; }  // FFMA2 + FMNMX + IADD
        }
    }
    float s = 0; for (int i = 0; i < UN; ++i) { float lo, hi; upk(p[i], lo, hi); s += x[i] + y[i] + lo + hi + n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------ medoid step variants
// rows in shared memory as pairs: (X0,X1,Y0,Y1)(Z0,Z1,N0,N1)
// VAR 0: current kernel step (2 FMNMX per row, negated result)
// VAR 1: NaN clean-up step: negated chain, rsq(-nr), one FMNMX at the end
// VAR 2: VAR 1 fully scalar (no f32x2)
// VAR 3: VAR 1 with scalar chain, packed sqrt
// VAR 4: VAR 1 with packed chain, scalar sqrt
template <int VAR, int NC>
__global__ void kmed(float *out, const float4 *rows_g, float xj, float yj, float zj, float nj)
{
    __shared__ float4 s_rows[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_rows[i] = rows_g[i];
    __syncthreads();
    u64 xj2[NC], yj2[NC], zj2[NC], nj2[NC];
    float a0[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const float o = (float)(threadIdx.x + c * 256);
        xj2[c] = pk(xj + o, xj + o); yj2[c] = pk(yj, yj); zj2[c] = pk(zj, zj); nj2[c] = pk(nj + 2400.f * o, nj + 2400.f * o);
        a0[c] = 0.f;
    }
    for (int it = 0; it < 64; ++it) {
        for (int b = 0; b < 1024; b += 16) {
#pragma unroll
            for (int hb = 0; hb < 16; hb += 8) {
                float nd[NC][8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 a = s_rows[b + hb + 2 * q], bb = s_rows[b + hb + 2 * q + 1];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        if (VAR == 0) {
                            u64 r = mul2(pk(a.x, a.y), xj2[c]);
                            r = fma2(pk(a.z, a.w), yj2[c], r);
                            r = fma2(pk(bb.x, bb.y), zj2[c], r);
                            r = add2(pk(bb.z, bb.w), r);
                            r = add2(nj2[c], r);
                            float r0, r1, y0, y1;
                            upk(r, r0, r1);
                            const float nx0 = fminf(-r0, -0.0f), nx1 = fminf(-r1, -0.0f);
                            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(fmaxf(r0, 0x1p-101f)));
                            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(fmaxf(r1, 0x1p-101f)));
                            const u64 nx = pk(nx0, nx1), y = pk(y0, y1);
                            const u64 ns = mul2(nx, y);
                            const u64 h = mul2(y, pk(0.5f, 0.5f));
                            const u64 rr = fma2(ns, ns, nx);
                            upk(fma2(rr, h, ns), nd[c][2 * q], nd[c][2 * q + 1]);
                        } else if (VAR == 1 || VAR == 4) {
                            // rows/columns hold the NEGATED constants: chain gives nr = -r
                            u64 nr = mul2(pk(a.x, a.y), xj2[c]);
                            nr = fma2(pk(a.z, a.w), yj2[c], nr);
                            nr = fma2(pk(bb.x, bb.y), zj2[c], nr);
                            nr = add2(pk(bb.z, bb.w), nr);
                            nr = add2(nj2[c], nr);
                            float n0, n1, y0, y1;
                            upk(nr, n0, n1);
                            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(-n0));
                            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(-n1));
                            if (VAR == 1) {
                                const u64 y = pk(y0, y1);
                                const u64 ns = mul2(nr, y);
                                const u64 h = mul2(y, pk(0.5f, 0.5f));
                                const u64 rr = fma2(ns, ns, nr);
                                float d0, d1;
                                upk(fma2(rr, h, ns), d0, d1);
                                nd[c][2 * q] = fminf(d0, -0.0f); nd[c][2 * q + 1] = fminf(d1, -0.0f);
                            } else {
                                const float s0 = __fmul_rn(n0, y0), s1 = __fmul_rn(n1, y1);
                                const float h0 = __fmul_rn(y0, 0.5f), h1 = __fmul_rn(y1, 0.5f);
                                const float e0 = __fmaf_rn(s0, s0, n0), e1 = __fmaf_rn(s1, s1, n1);
                                nd[c][2 * q] = fminf(__fmaf_rn(e0, h0, s0), -0.0f);
                                nd[c][2 * q + 1] = fminf(__fmaf_rn(e1, h1, s1), -0.0f);
                            }
                        } else {   // VAR 2, 3: scalar chain
                            float xs, xd, ys, yd, zs, zd, ns_, nd_;
                            upk(xj2[c], xs, xd); upk(yj2[c], ys, yd); upk(zj2[c], zs, zd); upk(nj2[c], ns_, nd_);
                            float n0 = __fmul_rn(a.x, xs), n1 = __fmul_rn(a.y, xs);
                            n0 = __fmaf_rn(a.z, ys, n0); n1 = __fmaf_rn(a.w, ys, n1);
                            n0 = __fmaf_rn(bb.x, zs, n0); n1 = __fmaf_rn(bb.y, zs, n1);
                            n0 = __fadd_rn(bb.z, n0); n1 = __fadd_rn(bb.w, n1);
                            n0 = __fadd_rn(ns_, n0); n1 = __fadd_rn(ns_, n1);
                            float y0, y1;
                            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(-n0));
                            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(-n1));
                            if (VAR == 2) {
                                const float s0 = __fmul_rn(n0, y0), s1 = __fmul_rn(n1, y1);
                                const float h0 = __fmul_rn(y0, 0.5f), h1 = __fmul_rn(y1, 0.5f);
                                const float e0 = __fmaf_rn(s0, s0, n0), e1 = __fmaf_rn(s1, s1, n1);
                                nd[c][2 * q] = fminf(__fmaf_rn(e0, h0, s0), -0.0f);
                                nd[c][2 * q + 1] = fminf(__fmaf_rn(e1, h1, s1), -0.0f);
                            } else {
                                const u64 nr = pk(n0, n1), y = pk(y0, y1);
                                const u64 ns = mul2(nr, y);
                                const u64 h = mul2(y, pk(0.5f, 0.5f));
                                const u64 rr = fma2(ns, ns, nr);
                                float d0, d1;
                                upk(fma2(rr, h, ns), d0, d1);
                                nd[c][2 * q] = fminf(d0, -0.0f); nd[c][2 * q + 1] = fminf(d1, -0.0f);
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
#pragma unroll
                    for (int c = 0; c < NC; ++c) a0[c] = __fsub_rn(a0[c], nd[c][q]);
            }
        }
    }
    float s = 0;
    for (int c = 0; c < NC; ++c) s += a0[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// Column-packed step: every packed op handles (column c0, column c1) of ONE row; rows in shared
// memory with duplicated halves: s_rows[2r] = (X,X,Y,Y), s_rows[2r+1] = (Z,Z,W,W).  NCP column pairs/thread.
template <int NCP, int ROWS>
__global__ void kmedc(float *out, const float4 *rows_g, float xj, float yj, float zj, float nnj)
{
    __shared__ float4 s_rows[2 * ROWS];
    for (int i = threadIdx.x; i < 2 * ROWS; i += blockDim.x) s_rows[i] = rows_g[i];
    __syncthreads();
    u64 xj2[NCP], yj2[NCP], zj2[NCP], nj2[NCP], a0[NCP];
#pragma unroll
    for (int c = 0; c < NCP; ++c) {
        const float o = (float)(threadIdx.x + c * 256);
        xj2[c] = pk(xj + o, xj + o + 0.5f); yj2[c] = pk(yj, yj + 1.f); zj2[c] = pk(zj, zj); nj2[c] = pk(nnj - 2400.f * o, nnj - 2400.f * o - 1200.f);
        a0[c] = pk(0.f, 0.f);
    }
    for (int it = 0; it < 64 * (1024 / ROWS); ++it) {
        for (int b = 0; b < ROWS; b += 8) {
            u64 nd[NCP][8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a = s_rows[2 * (b + q)], bb = s_rows[2 * (b + q) + 1];
#pragma unroll
                for (int c = 0; c < NCP; ++c) {
                    u64 nr = mul2(pk(a.x, a.y), xj2[c]);
                    nr = fma2(pk(a.z, a.w), yj2[c], nr);
                    nr = fma2(pk(bb.x, bb.y), zj2[c], nr);
                    nr = add2(pk(bb.z, bb.w), nr);
                    nr = add2(nj2[c], nr);
                    float n0, n1, y0, y1;
                    upk(nr, n0, n1);
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(-n0));
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(-n1));
                    const u64 y = pk(y0, y1);
                    const u64 ns = mul2(nr, y);
                    const u64 h = mul2(y, pk(0.5f, 0.5f));
                    const u64 rr = fma2(ns, ns, nr);
                    float d0, d1;
                    upk(fma2(rr, h, ns), d0, d1);
                    nd[c][q] = pk(fminf(d0, -0.0f), fminf(d1, -0.0f));
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int c = 0; c < NCP; ++c) asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(a0[c]) : "l"(nd[c][q]));
        }
    }
    float s = 0;
    for (int c = 0; c < NCP; ++c) { float lo, hi; upk(a0[c], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NCP, int ROWS> void runmedc(const char *name, int blocks_per_sm, int threads)
{
    float *d; cudaMalloc(&d, 148 * 16 * 1024 * 4);
    float4 *rows; cudaMalloc(&rows, 2 * ROWS * 16);
    float4 *h = new float4[2 * ROWS];
    for (int i = 0; i < ROWS; ++i) {
        const float x0 = 1200 + i * 0.01f, y0 = 950.f, z0 = 1.f;
        const float n0 = x0 * x0 + y0 * y0 + z0 * z0;
        h[2 * i] = make_float4(2 * x0, 2 * x0, 2 * y0, 2 * y0);
        h[2 * i + 1] = make_float4(2 * z0, 2 * z0, -n0, -n0);
    }
    cudaMemcpy(rows, h, 2 * ROWS * 16, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const float nj = 1200.f * 1200.f + 950.f * 950.f + 1.f;
    kmedc<NCP, ROWS><<<148 * blocks_per_sm, threads>>>(d, rows, 1200.f, 950.f, 1.f, -nj);
    cudaEventRecord(e0);
    kmedc<NCP, ROWS><<<148 * blocks_per_sm, threads>>>(d, rows, 1200.f, 950.f, 1.f, -nj);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9;
    double warp_rows_per_smsp = (blocks_per_sm * threads / 32 / 4.0) * 64.0 * 1024 * NCP * 2;
    cudaError_t err = cudaGetLastError();
    printf("%-40s NCP=%d rows/tile=%d %d blk/SM x %d thr: %.3f ms -> %.2f cycles per warp-row-col per SMSP %s\n", name, NCP, ROWS, blocks_per_sm, threads, ms,
           cyc / warp_rows_per_smsp, err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(d); cudaFree(rows);
}

template <int VAR, int NC> void runmed(const char *name, int blocks_per_sm, int threads)
{
    float *d; cudaMalloc(&d, 148 * 16 * 1024 * 4);
    float4 *rows; cudaMalloc(&rows, 1024 * 16);
    float4 *h = new float4[1024];
    const float sg = VAR == 0 ? -2.f : 2.f;
    for (int i = 0; i < 1024; i += 2) {
        // pair (X0,X1,Y0,Y1)(Z0,Z1,N0,N1)
        const float x0 = 1200 + i * 0.01f, x1 = 1200 + (i + 1) * 0.01f, y0 = 950.f, y1 = 951.f, z0 = 1.f, z1 = 1.5f;
        const float n0 = x0 * x0 + y0 * y0 + z0 * z0, n1 = x1 * x1 + y1 * y1 + z1 * z1;
        h[i] = make_float4(sg * x0, sg * x1, sg * y0, sg * y1);
        h[i + 1] = make_float4(sg * z0, sg * z1, VAR == 0 ? n0 : -n0, VAR == 0 ? n1 : -n1);
    }
    cudaMemcpy(rows, h, 1024 * 16, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const float nj = 1200.f * 1200.f + 950.f * 950.f + 1.f;
    kmed<VAR, NC><<<148 * blocks_per_sm, threads>>>(d, rows, 1200.f, 950.f, 1.f, VAR == 0 ? nj : -nj);
    cudaEventRecord(e0);
    kmed<VAR, NC><<<148 * blocks_per_sm, threads>>>(d, rows, 1200.f, 950.f, 1.f, VAR == 0 ? nj : -nj);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9;
    double warp_rows_per_smsp = (blocks_per_sm * threads / 32 / 4.0) * 64.0 * 1024 * NC;
    cudaError_t err = cudaGetLastError();
    printf("%-40s NC=%d %d blk/SM x %d thr: %.3f ms -> %.2f cycles per warp-row-col per SMSP %s\n", name, NC, blocks_per_sm, threads, ms,
           cyc / warp_rows_per_smsp, err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(d); cudaFree(rows);
}
template <int MODE> void run(const char *name, double ops_per_iter)
{
    float *d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 512>>>(d, 1.0001f, 0.9999f, 3);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(d, 1.0001f, 0.9999f, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9;
    double groups_per_smsp = 16.0 * ITER * UN;
    printf("%-34s %.3f ms  -> %.2f cycles per group per SMSP (%g instr/group)\n", name, ms, cyc / groups_per_smsp, ops_per_iter);
    cudaFree(d);
}
int main()
{
    run<0>("FFMA2 + IMAD", 2); run<1>("FFMA + IMAD", 2); run<2>("FFMA2 + LOP3 + IADD", 3); run<3>("FFMA2 + FADD + 2 FMNMX", 4);
    run<4>("FFMA2 + MUFU", 2); run<5>("2 FFMA2 + MUFU", 3); run<6>("FFMA2 + LDS + FADD", 3); run<7>("LOP3 + IADD", 2);
    run<8>("FFMA2 + FMNMX + IADD", 3);
    runmed<0, 1>("current step", 4, 256); runmed<0, 2>("current step", 4, 128); runmed<0, 2>("current step", 7, 128);
    runmed<1, 1>("NaN-cleanup packed", 4, 256); runmed<1, 2>("NaN-cleanup packed", 4, 128); runmed<1, 2>("NaN-cleanup packed", 7, 128);
    runmed<1, 4>("NaN-cleanup packed", 4, 128);
    runmed<2, 1>("NaN-cleanup scalar", 4, 256); runmed<2, 2>("NaN-cleanup scalar", 4, 128);
    runmed<3, 2>("scalar chain, packed sqrt", 4, 128); runmed<4, 2>("packed chain, scalar sqrt", 4, 128);
    runmedc<1, 512>("column-packed", 4, 256); runmedc<1, 512>("column-packed", 7, 128); runmedc<1, 512>("column-packed", 8, 128);
    runmedc<2, 512>("column-packed", 4, 128); runmedc<2, 512>("column-packed", 7, 128); runmedc<2, 512>("column-packed", 7, 64); runmedc<2, 512>("column-packed", 14, 64);
    runmedc<1, 1024>("column-packed", 6, 128);
    return 0;
}
