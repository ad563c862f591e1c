// soma_kernels.cuh -- the data-parallel helpers of the plugin's soma branch (second part),
// only reached when somaradius > 0 (Advantra_plugin.cpp:2426-2440):
//   Frangi::imerode(I,w,h,l,rad,E)   frangi.cpp:879-969    separable xy minimum, window radius ceil(rad), replicate clamp
//   Frangi::imdilate(I,w,h,l,rad)    frangi.cpp:1110-1199  the same with the maximum, in place
//   Frangi::imgaussian(I,w,h,l,sig)  frangi.cpp:786-877    xy Gaussian of a uint8 volume IN PLACE: the x pass accumulates
//       in float32, the y pass accumulates INTO THE unsigned char (`I[i0] += K[i1]*G`), i.e. every tap truncates
//       (uint8)((float)acc + K*G) -- restated tap by tap.
// One thread per voxel, lanes along x, taps straight from global memory (L1 / L2 hits): these are cold helpers
// (one call per image), written for exactness first.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "frangi_kernels.cuh"

namespace frangi {

// one 1-D pass of the minimum (IS_MIN) / maximum over [c - L, c + L] clamped, along x (ALONG_Y = false) or y
template <bool IS_MIN, bool ALONG_Y>
__global__ void __launch_bounds__(256)
morph_pass_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int w, int h, long long planes, int L)
{
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w) return;
    for (long long z = blockIdx.z; z < planes; z += gridDim.z) {
        const uint8_t* pl = in + z * (long long)w * h;
        int v = __ldg(pl + (long long)y * w + x);
        for (int k = -L; k <= L; ++k) {
            const int q = ALONG_Y ? __ldg(pl + (long long)clampi(y + k, 0, h - 1) * w + x)
                                  : __ldg(pl + (long long)y * w + clampi(x + k, 0, w - 1));
            v = IS_MIN ? min(v, q) : max(v, q);
        }
        out[z * (long long)w * h + (long long)y * w + x] = (uint8_t)v;
    }
}

// z pass of the z-scaled erosion Frangi::imerode(I,w,h,l,rad,zdist,E) (frangi.cpp:971-1108): minimum over
// [z - Lz, z + Lz] clamped to the volume, Lz = ceil(rad / zdist); one thread per (x, y) column and plane
__global__ void __launch_bounds__(256)
morph_min_z_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int w, int h, int l, int Lz)
{
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w) return;
    const long long plane = (long long)w * h;
    const uint8_t* col = in + (long long)y * w + x;
    for (int z = blockIdx.z; z < l; z += gridDim.z) {
        int v = __ldg(col + z * plane);
        for (int k = -Lz; k <= Lz; ++k) v = min(v, (int)__ldg(col + clampi(z + k, 0, l - 1) * plane));
        out[z * plane + (long long)y * w + x] = (uint8_t)v;
    }
}

// x pass of the in-place Gaussian: uint8 -> float32, ascending taps from zero, separately rounded (frangi.cpp:806-838)
__global__ void __launch_bounds__(256)
gauss_x_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ K, int w, int h, long long planes, int L, int Lt,
                  const __grid_constant__ GaussTaps taps)
{
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w) return;
    for (long long z = blockIdx.z; z < planes; z += gridDim.z) {
        const uint8_t* row = in + z * (long long)w * h + (long long)y * w;
        float acc = 0.0f;
        for (int k = -L; k <= L; ++k)
            acc = __fadd_rn(acc, __fmul_rn((float)(int)__ldg(row + clampi(x + k, 0, w - 1)), taps.g[k + Lt]));
        K[z * (long long)w * h + (long long)y * w + x] = acc;
    }
}

// y pass: the accumulator IS the unsigned char (frangi.cpp:841-873): acc = (uint8)((float)acc + K * G) per tap
__global__ void __launch_bounds__(256)
gauss_y_trunc_kernel(const float* __restrict__ K, uint8_t* __restrict__ out, int w, int h, long long planes, int L, int Lt,
                     const __grid_constant__ GaussTaps taps)
{
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w) return;
    for (long long z = blockIdx.z; z < planes; z += gridDim.z) {
        const float* pl = K + z * (long long)w * h;
        unsigned acc = 0;
        for (int k = -L; k <= L; ++k) {
            const float t = __fadd_rn((float)(int)acc, __fmul_rn(__ldg(pl + (long long)clampi(y + k, 0, h - 1) * w + x), taps.g[k + Lt]));
            acc = (unsigned)__float2int_rz(t) & 0xffu;          // float -> unsigned char
        }
        out[z * (long long)w * h + (long long)y * w + x] = (uint8_t)acc;
    }
}

}  // namespace frangi
