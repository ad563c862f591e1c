// frangi2d_kernels.cuh -- the 2-D path of the filter: Frangi::frangi2d
// (reference pnr-vaa3d/frangi.cpp:392-505) over hessian2d (:507-560).  The smoothed image F comes from
// K1 (gauss_xy_kernel on one plane: the 2-D imgaussian, frangi.cpp:562-645, has the same taps, clamp and
// accumulation order).  One thread per pixel: the three second differences with the reference's face
// rules, then the closed-form 2 x 2 eigen analysis and vesselness.  The reference mixes float and double
// (pow(x, 2) of a float is a double square, `.5 * float` a double product, exp / sqrt / abs of floats the
// float overloads); every step is mirrored in the same precision, exp through the double routine rounded
// to float (2-D images are small; this is not a hot path).  Running maximum over scales, direction
// codes and Jmin / Jmax follow :462-503.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace frangi {

struct F2DParams {
    const float* F;          // smoothed image, row pitch fpitch
    int w, h, fpitch;
    float sigma2;            // sig*sig (float)
    float beta, c;           // 2*BetaOne^2, 2*BetaTwo^2 (float, frangi.cpp:411-412)
    int blackwhite;
    int first;               // si == 0
    float* J;                // dense w*h
    uint8_t *Vx, *Vy, *Vz;
    float* D[3];             // stage dump (Dyy, Dxy, Dxx) or NULL
    int* minmax;             // bit patterns of Jmin / Jmax (J >= 0)
};

__device__ __forceinline__ float f2d_dx(const F2DParams& p, int x, int y)   // frangi.cpp:536-540
{
    const float* r = p.F + (long long)y * p.fpitch;
    if (x == 0) return __fsub_rn(r[1], r[0]);
    if (x < p.w - 1) return (float)(.5 * (double)__fsub_rn(r[x + 1], r[x - 1]));
    return __fsub_rn(r[x], r[x - 1]);
}
__device__ __forceinline__ float f2d_dy(const F2DParams& p, int x, int y)   // frangi.cpp:516-520
{
    const float* c = p.F + x;
    if (y == 0) return __fsub_rn(c[(long long)p.fpitch], c[0]);
    if (y < p.h - 1) return (float)(.5 * (double)__fsub_rn(c[(long long)(y + 1) * p.fpitch], c[(long long)(y - 1) * p.fpitch]));
    return __fsub_rn(c[(long long)y * p.fpitch], c[(long long)(y - 1) * p.fpitch]);
}
// the same rule applied to a first-difference field g(x, y) along x or y
template <class G>
__device__ __forceinline__ float f2d_second(G g, int x, int y, int n, bool along_y)
{
    const int c = along_y ? y : x;
    auto at = [&](int q) { return along_y ? g(x, q) : g(q, y); };
    if (c == 0) return __fsub_rn(at(1), at(0));
    if (c < n - 1) return (float)(.5 * (double)__fsub_rn(at(c + 1), at(c - 1)));
    return __fsub_rn(at(c), at(c - 1));
}

__device__ __forceinline__ uint8_t f2d_code(float c, float n)              // frangi.cpp:457-459
{
    const float q = __fdiv_rn(__fadd_rn(__fdiv_rn(c, n), 1.0f), 2.0f);
    const double r = round((double)q * 255.0);
    if (!(r == r)) return 0;                      // 0/0 direction: the reference's int conversion is negative -> 0
    const int v = (int)r;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__global__ void __launch_bounds__(128)
frangi2d_pixel_kernel(const __grid_constant__ F2DParams p)
{
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    float jv = 0.0f;
    bool assigned = false;
    if (x < p.w) {
        auto gx = [&](int a, int b) { return f2d_dx(p, a, b); };
        auto gy = [&](int a, int b) { return f2d_dy(p, a, b); };
        const float Dyy = __fmul_rn(f2d_second(gy, x, y, p.h, true), p.sigma2);
        const float Dxx = __fmul_rn(f2d_second(gx, x, y, p.w, false), p.sigma2);
        const float Dxy = __fmul_rn(f2d_second(gx, x, y, p.h, true), p.sigma2);
        const long long i = (long long)y * p.w + x;
        if (p.D[0]) { p.D[0][i] = Dyy; p.D[1][i] = Dxy; p.D[2][i] = Dxx; }
        else {
        const float d = __fsub_rn(Dxx, Dyy);
        const float tmp = (float)sqrt((double)d * (double)d + 4.0 * ((double)Dxy * (double)Dxy));
        float v2x = __fmul_rn(2.0f, Dxy);
        float v2y = __fadd_rn(__fsub_rn(Dyy, Dxx), tmp);
        const float mag = (float)sqrt((double)v2x * (double)v2x + (double)v2y * (double)v2y);
        if (mag > 0) { v2x = __fdiv_rn(v2x, mag); v2y = __fdiv_rn(v2y, mag); }
        const float v1x = -v2y, v1y = v2x;
        const float mu1 = (float)(0.5 * (double)__fadd_rn(__fadd_rn(Dxx, Dyy), tmp));
        const float mu2 = (float)(0.5 * (double)__fsub_rn(__fadd_rn(Dxx, Dyy), tmp));
        const bool check = fabsf(mu1) < fabsf(mu2);
        float L1 = check ? mu2 : mu1;
        const float L2 = check ? mu1 : mu2;
        const float Vecx = check ? v2x : v1x, Vecy = check ? v2y : v1y;
        L1 = (L1 == 0) ? 1.175494351e-38f : L1;
        const float q = __fdiv_rn(L2, L1);
        const float Rb = (float)((double)q * (double)q);
        const float S2 = (float)((double)L1 * (double)L1 + (double)L2 * (double)L2);
        const float ea = (float)exp((double)__fdiv_rn(-Rb, p.beta));
        const float eb = (float)exp((double)__fdiv_rn(-S2, p.c));
        float v = __fmul_rn(ea, __fsub_rn(1.0f, eb));
        if (p.blackwhite) v = (L1 < 0) ? 0.0f : v; else v = (L1 > 0) ? 0.0f : v;
        if (p.first || v > p.J[i]) {
            p.J[i] = v;
            const float Vecn = __fsqrt_rn(__fadd_rn(__fmul_rn(Vecx, Vecx), __fmul_rn(Vecy, Vecy)));
            p.Vx[i] = f2d_code(Vecx, Vecn);
            p.Vy[i] = f2d_code(Vecy, Vecn);
            p.Vz[i] = 0;
            jv = v; assigned = true;
        }
        }
    }
    if (p.D[0]) return;                           // stage dump (uniform)
    // Jmin / Jmax over the assigned values (frangi.cpp:464-465, 484-485)
    float mn = assigned ? jv : 3.4e38f, mx = assigned ? jv : 0.0f;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mn < 3.0e38f) atomicMin(p.minmax + 0, __float_as_int(mn));
        atomicMax(p.minmax + 1, __float_as_int(mx));
    }
}

}  // namespace frangi
