// tools/microbench.cu -- pipe-rate probes for design decisions (not product code).
// Reports lane-ops per clock per SM for: FFMA, FMUL+FADD pairs, packed fma.rn.f32x2,
// packed mul+add f32x2, DFMA, MUFU.EX2, LDS.32/LDS.128.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE> __global__ void k(float* out, float a, float b)
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    if (MODE == 0) {
#pragma unroll 16
        for (int i = 0; i < ITER; ++i) { x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b); x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b); }
    } else if (MODE == 1) {
#pragma unroll 16
        for (int i = 0; i < ITER; ++i) { x0 = __fadd_rn(__fmul_rn(x0, a), b); x1 = __fadd_rn(__fmul_rn(x1, a), b); x2 = __fadd_rn(__fmul_rn(x2, a), b); x3 = __fadd_rn(__fmul_rn(x3, a), b); x4 = __fadd_rn(__fmul_rn(x4, a), b); x5 = __fadd_rn(__fmul_rn(x5, a), b); x6 = __fadd_rn(__fmul_rn(x6, a), b); x7 = __fadd_rn(__fmul_rn(x7, a), b); }
    } else if (MODE == 2) {
        unsigned long long p0, p1, p2, p3, pa, pb;
        asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1)); asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
        asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x4), "f"(x5)); asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x6), "f"(x7));
        asm("mov.b64 %0, {%1,%2};" : "=l"(pa) : "f"(a), "f"(a)); asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
#pragma unroll 16
        for (int i = 0; i < ITER; ++i) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
        }
        asm("mov.b64 {%0,%1}, %2;" : "=f"(x0), "=f"(x1) : "l"(p0)); asm("mov.b64 {%0,%1}, %2;" : "=f"(x2), "=f"(x3) : "l"(p1));
        asm("mov.b64 {%0,%1}, %2;" : "=f"(x4), "=f"(x5) : "l"(p2)); asm("mov.b64 {%0,%1}, %2;" : "=f"(x6), "=f"(x7) : "l"(p3));
    } else if (MODE == 3) {
        unsigned long long p0, p1, p2, p3, pa, pb;
        asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1)); asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
        asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x4), "f"(x5)); asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x6), "f"(x7));
        asm("mov.b64 %0, {%1,%2};" : "=l"(pa) : "f"(a), "f"(a)); asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
#pragma unroll 16
        for (int i = 0; i < ITER; ++i) {
            asm volatile("mul.rn.f32x2 %0, %0, %1; add.rn.f32x2 %0, %0, %2;" : "+l"(p0) : "l"(pa), "l"(pb)); asm volatile("mul.rn.f32x2 %0, %0, %1; add.rn.f32x2 %0, %0, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            asm volatile("mul.rn.f32x2 %0, %0, %1; add.rn.f32x2 %0, %0, %2;" : "+l"(p2) : "l"(pa), "l"(pb)); asm volatile("mul.rn.f32x2 %0, %0, %1; add.rn.f32x2 %0, %0, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
        }
        asm("mov.b64 {%0,%1}, %2;" : "=f"(x0), "=f"(x1) : "l"(p0)); asm("mov.b64 {%0,%1}, %2;" : "=f"(x2), "=f"(x3) : "l"(p1));
        asm("mov.b64 {%0,%1}, %2;" : "=f"(x4), "=f"(x5) : "l"(p2)); asm("mov.b64 {%0,%1}, %2;" : "=f"(x6), "=f"(x7) : "l"(p3));
    } else if (MODE == 4) {
        double d0 = x0, d1 = x1, d2 = x2, d3 = x3, da = a, db = b;
#pragma unroll 16
        for (int i = 0; i < ITER; ++i) { d0 = fma(d0, da, db); d1 = fma(d1, da, db); d2 = fma(d2, da, db); d3 = fma(d3, da, db); d0 = fma(d0, da, db); d1 = fma(d1, da, db); d2 = fma(d2, da, db); d3 = fma(d3, da, db); }
        x0 = d0 + d1 + d2 + d3;
    } else if (MODE == 5) {
#pragma unroll 16
        for (int i = 0; i < ITER; ++i) { x0 = exp2f(x0); x1 = exp2f(x1); x2 = exp2f(x2); x3 = exp2f(x3); x4 = exp2f(x4); x5 = exp2f(x5); x6 = exp2f(x6); x7 = exp2f(x7); }
    } else if (MODE == 6) {
        int idx = threadIdx.x;
#pragma unroll 16
        for (int i = 0; i < ITER; ++i) { x0 += sm[idx]; x1 += sm[(idx + 32) & 4095]; x2 += sm[(idx + 64) & 4095]; x3 += sm[(idx + 96) & 4095]; x4 += sm[(idx + 128) & 4095]; x5 += sm[(idx + 160) & 4095]; x6 += sm[(idx + 192) & 4095]; x7 += sm[(idx + 224) & 4095]; idx = (idx + 256) & 4095; }
    } else if (MODE == 7) {
        int idx = (threadIdx.x * 4) & 4095;
#pragma unroll 8
        for (int i = 0; i < ITER; ++i) { float4 q = *(float4*)&sm[idx]; float4 r = *(float4*)&sm[(idx + 1024) & 4095]; x0 += q.x; x1 += q.y; x2 += q.z; x3 += q.w; x4 += r.x; x5 += r.y; x6 += r.z; x7 += r.w; idx = (idx + 2048) & 4095; }
    } else if (MODE == 8) {   // mixed: 8 FFMA + 1 LDS per group (conv inner-loop shape)
        int idx = threadIdx.x;
#pragma unroll 16
        for (int i = 0; i < ITER; ++i) { float v = sm[idx]; x0 = __fmaf_rn(v, a, x0); x1 = __fmaf_rn(v, b, x1); x2 = __fmaf_rn(v, a, x2); x3 = __fmaf_rn(v, b, x3); x4 = __fmaf_rn(v, a, x4); x5 = __fmaf_rn(v, b, x5); x6 = __fmaf_rn(v, a, x6); x7 = __fmaf_rn(v, b, x7); idx = (idx + 32) & 4095; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <int MODE> void run(const char* name, double ops_per_iter_per_thread)
{
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * 2, threads = 1024;
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 1e-7f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(out, 1.0001f, 1e-7f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double ops = (double)blocks * threads * ITER * ops_per_iter_per_thread;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %8.2f Tops/s  %7.1f lane-ops/clk/SM @%d MHz nominal\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main()
{
    run<0>("FFMA", 8); run<1>("FMUL+FADD (pairs)", 8); run<2>("FFMA2 f32x2 (fma lanes)", 8); run<3>("FMUL2+FADD2 (pairs lanes)", 8);
    run<4>("DFMA", 8); run<5>("MUFU.EX2", 8); run<6>("LDS.32 (+FADD)", 8); run<7>("LDS.128 (floats)", 8); run<8>("8 FFMA + 1 LDS (fma)", 8);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
