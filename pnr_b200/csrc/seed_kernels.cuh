// seed_kernels.cuh -- the data-parallel pre-pass of the first consumer of the Frangi outputs,
// SeedExtractor::extractSeeds (reference pnr-vaa3d/seed.cpp:574-632), on the J8 volume that the
// filter leaves on the device.  Per z layer, exactly as the reference:
//   K5a layer range       globalMin / globalMax of the layer                       seed.cpp:578-586
//   K5b candidate maxima  pixels that are not on the layer border, differ from globalMin and have
//                         no strictly greater 8-neighbour                            seed.cpp:590-614
//       ranking key       (int)((v - globalMin) * (float)(2e9 / (globalMax - globalMin))) << 32 | y*w+x
//                                                                                    seed.cpp:616-628
// The keys of a layer are then sorted ascending (seed.cpp:630) by a segmented radix sort.  The
// flood-fill / tolerance analysis that follows (seed.cpp:643-782) is sequential per layer and works on
// this (small) list; it stays with the caller.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace frangi {

// K5a: minmax[2z] = min, minmax[2z+1] = max of layer z (ints, initialised to 255 / 0 by the host)
__global__ void __launch_bounds__(256)
j8_layer_minmax_kernel(const uint8_t* __restrict__ J8, long long plane, int* __restrict__ minmax)
{
    const int z = blockIdx.y;
    const uint8_t* layer = J8 + (long long)z * plane;
    unsigned lo = 255u, hi = 0u;
    // head bytes up to 4-byte alignment, 32-bit body, tail bytes
    const long long mis = (4 - (reinterpret_cast<uintptr_t>(layer) & 3)) & 3;
    const long long head = mis < plane ? mis : plane;
    const long long nwords = (plane - head) / 4;
    const uint32_t* body = reinterpret_cast<const uint32_t*>(layer + head);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
        const uint32_t v = __ldg(body + i);
        const unsigned a = __vminu4(v, (v >> 8) | (v << 24)), b = __vmaxu4(v, (v >> 8) | (v << 24));
        const unsigned a2 = __vminu4(a, (a >> 16) | (a << 16)), b2 = __vmaxu4(b, (b >> 16) | (b << 16));
        lo = min(lo, a2 & 0xffu); hi = max(hi, b2 & 0xffu);
    }
    if (blockIdx.x == 0) {
        const long long tail0 = head + 4 * nwords;
        for (long long i = threadIdx.x; i < head + (plane - tail0); i += blockDim.x) {
            const unsigned v = layer[i < head ? i : tail0 + (i - head)];
            lo = min(lo, v); hi = max(hi, v);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo < 255u) atomicMin(minmax + 2 * z, (int)lo);
        if (hi > 0u) atomicMax(minmax + 2 * z + 1, (int)hi);
    }
}

// K5b: one thread per pixel, a warp = 32 consecutive x of one row.  FILL = false counts the candidates of
// each layer, FILL = true writes their keys at keys[offsets[z] + cursor[z]++] (unordered inside a layer;
// the sort that follows fixes the order, keys are unique).
template <bool FILL>
__global__ void __launch_bounds__(256)
j8_local_maxima_kernel(const uint8_t* __restrict__ J8, int w, int h, const int* __restrict__ minmax,
                       int* __restrict__ counter, const long long* __restrict__ offsets, long long* __restrict__ keys)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int z = blockIdx.z;
    const long long plane = (long long)w * h;
    const uint8_t* layer = J8 + (long long)z * plane;
    const int lo = minmax[2 * z], hi = minmax[2 * z + 1];
    bool is_max = false;
    int v = 0;
    if (x > 0 && x < w - 1 && y > 0 && y < h - 1) {
        const uint8_t* c = layer + (long long)y * w + x;
        v = __ldg(c);
        if (v != lo) {
            int m = max(max(__ldg(c - w - 1), __ldg(c - w)), __ldg(c - w + 1));
            m = max(m, max(__ldg(c - 1), __ldg(c + 1)));
            m = max(m, max(max(__ldg(c + w - 1), __ldg(c + w)), __ldg(c + w + 1)));
            is_max = m <= v;
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, is_max);
    if (ballot == 0) return;
    const int lane = threadIdx.x;
    int base = 0;
    if (lane == 0) base = atomicAdd(counter + z, __popc(ballot));
    if (!FILL) return;
    base = __shfl_sync(0xffffffffu, base, 0);
    if (is_max) {
        // the reference's float arithmetic, operation by operation (seed.cpp:616, 625-626)
        const float gmin = (float)lo, gmax = (float)hi;
        const float factor = (float)(2e9 / (double)__fsub_rn(gmax, gmin));
        const int iv = (int)__fmul_rn(__fsub_rn((float)v, gmin), factor);
        const long long p = (long long)y * w + x;
        keys[offsets[z] + base + __popc(ballot & ((1u << lane) - 1u))] = ((long long)iv << 32) | p;
    }
}

}  // namespace frangi
