// frangi_kernels.cuh -- hand-written sm_100a kernels of the Frangi hot path.
//
// Stages (reference lines are pnr-vaa3d/frangi.cpp):
//   K1 gauss_xy_kernel      u8 -> f32, x then y Gaussian passes            :683-748
//   K2 gauss_z_kernel       f32 -> f32, z Gaussian pass (sigma/zdist)      :751-782
//   K3 hessian_eigen_kernel second differences, 3x3 eigen, vesselness,
//                           running max over scales, direction, min/max    :306-381, :190-273
//   K4 j_to_j8_kernel       min-max normalisation to 8 bit   Advantra_plugin.cpp:2499-2512
//
// Arithmetic contract.  The smoothing accumulates in float32 in ascending tap
// order from zero; in EXACT mode every tap is a separate rounded multiply and
// add (__fmul_rn/__fadd_rn, which nvcc never contracts), which is what the
// reference's x86-64 -O2 build executes, so the smoothed volume and the six
// second differences are bit-identical to the reference.  The finite
// differences always use rounded sub/mul.  The eigen stage is float32 and
// closed-form (the reference runs an iterative double-precision QL); it is
// built so that close eigenvalue pairs are split from a deflated 2x2 problem
// rather than from the trigonometric formula, which keeps the vesselness well
// inside the 1e-4 relative / 1e-6 absolute tolerance of BASELINE.json.
#pragma once
#include <cuda.h>            // CUtensorMap (type only; the encoder is fetched at run time)
#include <cuda_runtime.h>
#include <stdint.h>

#include "frangi_voxel_math.cuh"
#include "ref_eigen.h"

namespace frangi {

constexpr int kMaxRadius = 30;              // largest supported tap radius (sigma <= 10 in xy)
constexpr int kMaxTaps = 2 * kMaxRadius + 1;

struct GaussTaps {
    float g[kMaxTaps + 3];  // taps for template radius L live in g[0 .. 2L]
};

template <bool EXACT>
__device__ __forceinline__ float mac(float acc, float v, float g)
{
    if (EXACT) return __fadd_rn(acc, __fmul_rn(v, g));
    return __fmaf_rn(v, g, acc);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---------------------------------------------------------------------------
// K1: fused x and y Gaussian passes for one z-plane strip.
//
// A CTA owns TW=256 columns of one plane and marches down a y segment in
// batches of RB=16 rows.  Per batch: (1) RB input rows (+ x halo, replicate
// clamped) are converted u8 -> f32 into s_in; (2) the x pass: each thread
// produces 16 consecutive outputs of one row from a 16+2L window read with
// 128-bit shared loads (lanes run along rows, row pitch = 4 mod 32 words, so
// every quarter-warp touches 32 distinct banks) into a ring of x-passed rows;
// (3) the y pass: each thread owns one column and produces the batch's 16
// outputs from 16+2L ring rows (lanes along x: conflict-free, coalesced
// stores).  Both passes are register-blocked: 16*(2L+1) MACs per 16+2L
// shared-memory words.  The x pass of every row is done exactly once per
// y segment (only the segment's 2L halo rows are redundant).
// ---------------------------------------------------------------------------
template <int L>
struct XYCfg {
    static constexpr int TW = 256;
    static constexpr int RB = 16;
    static constexpr int NT = 256;
    static constexpr int LAL = (L + 3) / 4 * 4;  // halo rounded to 4 so that windows are float4-aligned
    static constexpr int WIN = 16 + 2 * LAL;
    static constexpr int PIN0 = TW + 2 * LAL;
    static constexpr int PIN = ((PIN0 / 4) % 2 == 1) ? PIN0 : PIN0 + 4;  // pitch/4 odd
    static constexpr int PR = TW + 4;                                    // 260: /4 odd
    static constexpr int NBLK = 1 + (2 * L + RB - 1) / RB;               // ring blocks of RB rows
    static constexpr int SMEM_BYTES = (RB * PIN + NBLK * RB * PR) * 4;
};

struct XYParams {
    const uint8_t* I;   // input planes of this launch, dense [nz][h][w]
    float* out;         // Fxy, plane 0 of this launch
    int w, h, nz;
    int fpitch;         // floats per output row
    long long fplane;   // floats per output plane
    int seg_h;          // rows per y segment (multiple of 16)
    int nstrips, nsegs;
    int vec_ok;         // rows of I are 4-byte aligned (w % 4 == 0 and base aligned)
    // Two plane ranges in one launch (the two boundary groups of a slab): launch planes >= zsplit are the planes
    // zsplit + zgap, ... of I / out.  One range: zsplit = nz, zgap = 0.
    int zsplit, zgap;
};

template <int L, bool EXACT>
__global__ void __launch_bounds__(256, 2)
gauss_xy_kernel(const __grid_constant__ XYParams p, const __grid_constant__ GaussTaps taps)
{
    using C = XYCfg<L>;
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;
    float* s_ring = smem + C::RB * C::PIN;

    const int tid = threadIdx.x;
    const int bid = blockIdx.x;
    const int strip = bid % p.nstrips;
    const int seg = (bid / p.nstrips) % p.nsegs;
    const int zl = bid / (p.nstrips * p.nsegs);
    const int z = zl >= p.zsplit ? zl + p.zgap : zl;
    const int x0 = strip * C::TW;
    const int ys = seg * p.seg_h;
    const int ye = min(ys + p.seg_h, p.h);
    const uint8_t* __restrict__ Iz = p.I + (long long)z * p.w * p.h;
    float* __restrict__ Oz = p.out + (long long)z * p.fplane;

    const int xr = tid & 15;   // x pass: row within the batch
    const int xc = tid >> 4;   // x pass: 16-column chunk
    const int nb = (ye - ys + C::RB - 1) / C::RB;
    const int nphases = nb + C::NBLK - 1;   // x-pass phases; phase m covers rows ys - L + m*RB + [0, RB)

    // Input staging is software-pipelined: the raw bytes of phase m+1 (4 pixels per 32-bit word,
    // replicate-clamped) are loaded into registers while phase m is being filtered, and are
    // converted to float in shared memory at the top of phase m+1.
    constexpr int GROUPS = C::PIN0 / 4;
    constexpr int NW = (C::RB * GROUPS + C::NT - 1) / C::NT;   // words per thread per phase
    uint32_t raw[NW];
    auto fetch = [&](int ph) {
        const int r0 = ys - L + ph * C::RB;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int idx = tid + k * C::NT;
            uint32_t u = 0;
            if (idx < C::RB * GROUPS) {
                const int r = idx / GROUPS;
                const int g = idx - r * GROUPS;
                const int y = clampi(r0 + r, 0, p.h - 1);
                const int xg = x0 - C::LAL + 4 * g;
                const uint8_t* row = Iz + (long long)y * p.w;
                if (p.vec_ok && xg >= 0 && xg + 3 < p.w) {
                    u = __ldg(reinterpret_cast<const uint32_t*>(row + xg));
                } else {
                    u = (uint32_t)__ldg(row + clampi(xg + 0, 0, p.w - 1)) |
                        ((uint32_t)__ldg(row + clampi(xg + 1, 0, p.w - 1)) << 8) |
                        ((uint32_t)__ldg(row + clampi(xg + 2, 0, p.w - 1)) << 16) |
                        ((uint32_t)__ldg(row + clampi(xg + 3, 0, p.w - 1)) << 24);
                }
            }
            raw[k] = u;
        }
    };
    fetch(0);

    for (int phase = 0; phase < nphases; ++phase) {
        {
            __syncthreads();  // s_in and the ring block about to be overwritten are no longer read
            // ---- stage RB input rows as float ----
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                const int idx = tid + k * C::NT;
                if (idx < C::RB * GROUPS) {
                    const int r = idx / GROUPS;
                    const int g = idx - r * GROUPS;
                    const uint32_t u = raw[k];
                    float4 v;
                    v.x = (float)(u & 0xffu);
                    v.y = (float)((u >> 8) & 0xffu);
                    v.z = (float)((u >> 16) & 0xffu);
                    v.w = (float)(u >> 24);
                    *reinterpret_cast<float4*>(s_in + r * C::PIN + 4 * g) = v;
                }
            }
            __syncthreads();
            if (phase + 1 < nphases) fetch(phase + 1);   // in flight during the x and y passes below
            // ---- x pass: 16 outputs of row xr, columns 16*xc .. 16*xc+15 ----
            {
                float acc[16];
#pragma unroll
                for (int o = 0; o < 16; ++o) acc[o] = 0.0f;
                const float4* src = reinterpret_cast<const float4*>(s_in + xr * C::PIN + 16 * xc);
#pragma unroll
                for (int i4 = 0; i4 < C::WIN / 4; ++i4) {
                    const float4 q = src[i4];
                    const float e[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int i = 4 * i4 + s;  // window index: x = x0 + 16*xc + i - LAL
#pragma unroll
                        for (int o = 0; o < 16; ++o) {
                            const int t = i - o - C::LAL + L;  // tap index
                            if (t >= 0 && t <= 2 * L) acc[o] = mac<EXACT>(acc[o], e[s], taps.g[t]);
                        }
                    }
                }
                float4* dst = reinterpret_cast<float4*>(
                    s_ring + ((phase % C::NBLK) * C::RB + xr) * C::PR + 16 * xc);
                dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                dst[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
                dst[3] = make_float4(acc[12], acc[13], acc[14], acc[15]);
            }
        }
        const int b = phase - (C::NBLK - 1);     // the batch whose 16 + 2L ring rows are now complete
        if (b < 0) continue;
        __syncthreads();
        // ---- y pass: column tid, output rows ys + b*RB + [0, RB) ----
        {
            float acc[C::RB];
#pragma unroll
            for (int o = 0; o < C::RB; ++o) acc[o] = 0.0f;
            const float* bp[C::NBLK];
            int blk = b % C::NBLK;
#pragma unroll
            for (int q = 0; q < C::NBLK; ++q) {
                bp[q] = s_ring + blk * C::RB * C::PR + tid;
                blk = (blk + 1 == C::NBLK) ? 0 : blk + 1;
            }
#pragma unroll
            for (int j = 0; j < C::RB + 2 * L; ++j) {
                const float v = bp[j / C::RB][(j % C::RB) * C::PR];
#pragma unroll
                for (int o = 0; o < C::RB; ++o) {
                    const int t = j - o;
                    if (t >= 0 && t <= 2 * L) acc[o] = mac<EXACT>(acc[o], v, taps.g[t]);
                }
            }
            const int x = x0 + tid;
            if (x < p.w) {
                const int ybase = ys + b * C::RB;
#pragma unroll
                for (int o = 0; o < C::RB; ++o) {
                    const int y = ybase + o;
                    if (y < ye) Oz[(long long)y * p.fpitch + x] = acc[o];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K1, FMA mode: the same strip / batch / ring structure with both passes in packed
// float32x2 (FFMA2 with the tap as a broadcast scalar operand), which halves the issue
// slots of the multiply-adds.  The x pass pairs the SAME output column of two adjacent
// rows, so the staged rows are interleaved in pairs ({row 2p, row 2p+1} per column) and
// a 128-bit shared load yields two ready register pairs; a thread makes 8 columns x 2
// rows.  The y pass pairs two adjacent columns of one row (a 64-bit load from the ring
// is a ready pair); a thread makes 8 rows x 2 columns and stores 64-bit.
// ---------------------------------------------------------------------------
template <int L>
struct XYFmaCfg {
    static constexpr int TW = 256;
    static constexpr int RB = 16;
    static constexpr int NT = 256;
    static constexpr int LAL = (L + 3) / 4 * 4;
    static constexpr int PIN0 = TW + 2 * LAL;                       // staged columns per row
    static constexpr int PINP = ((PIN0 + 13) / 16) * 16 + 2;        // >= PIN0 and = 2 mod 16
    static constexpr int PP = 2 * PINP;                             // floats per row pair: = 4 mod 32 (conflict-free LDS.128)
    static constexpr int PR = TW + 4;
    static constexpr int NBLK = 1 + (2 * L + RB - 1) / RB;
    static constexpr int SMEM_BYTES = ((RB / 2) * PP + NBLK * RB * PR) * 4;
};

template <int L, int YH>
__device__ __forceinline__ void xy_fma_ypass(const float* s_ring, const GaussTaps& taps, int b, int ycp, float* Oz, int x0,
                                             int ys, int ye, int w, int fpitch)
{
    using C = XYFmaCfg<L>;
    float2 acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = make_float2(0.0f, 0.0f);
    const float* bp[C::NBLK];
    int blk = b % C::NBLK;
#pragma unroll
    for (int q = 0; q < C::NBLK; ++q) {
        bp[q] = s_ring + blk * C::RB * C::PR + 2 * ycp;
        blk = (blk + 1 == C::NBLK) ? 0 : blk + 1;
    }
#pragma unroll
    for (int jj = 0; jj < 8 + 2 * L; ++jj) {
        constexpr int dummy = 0; (void)dummy;
        const int rr = jj + 8 * YH;              // ring row of the batch's window
        const float2 v = *reinterpret_cast<const float2*>(bp[rr / C::RB] + (rr % C::RB) * C::PR);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int t = jj - o;
            if (t >= 0 && t <= 2 * L) acc[o] = __ffma2_rn(v, make_float2(taps.g[t], taps.g[t]), acc[o]);
        }
    }
    const int x = x0 + 2 * ycp;
    const int ybase = ys + b * C::RB + 8 * YH;
    if (x + 1 < w) {
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (ybase + o < ye) *reinterpret_cast<float2*>(Oz + (long long)(ybase + o) * fpitch + x) = acc[o];
    } else if (x < w) {
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (ybase + o < ye) Oz[(long long)(ybase + o) * fpitch + x] = acc[o].x;
    }
}

template <int L>
#ifndef XY_FMA_CTAS3_MAXL
#define XY_FMA_CTAS3_MAXL 12     // radii up to this run 3 CTAs per SM (<= 85 registers), larger ones 2
#endif
__global__ void __launch_bounds__(256, (L <= XY_FMA_CTAS3_MAXL ? 3 : 2))
gauss_xy_fma_kernel(const __grid_constant__ XYParams p, const __grid_constant__ GaussTaps taps)
{
    using C = XYFmaCfg<L>;
    extern __shared__ __align__(16) float smem[];
    float* s_in2 = smem;
    float* s_ring = smem + (C::RB / 2) * C::PP;

    const int tid = threadIdx.x;
    const int bid = blockIdx.x;
    const int strip = bid % p.nstrips;
    const int seg = (bid / p.nstrips) % p.nsegs;
    const int zl = bid / (p.nstrips * p.nsegs);
    const int z = zl >= p.zsplit ? zl + p.zgap : zl;
    const int x0 = strip * C::TW;
    const int ys = seg * p.seg_h;
    const int ye = min(ys + p.seg_h, p.h);
    const uint8_t* __restrict__ Iz = p.I + (long long)z * p.w * p.h;
    float* __restrict__ Oz = p.out + (long long)z * p.fplane;

    const int nb = (ye - ys + C::RB - 1) / C::RB;
    const int nphases = nb + C::NBLK - 1;

    // staging: a thread fetches the same 4 columns of both rows of a pair (two 32-bit words)
    constexpr int GROUPS = C::PIN0 / 4;
    constexpr int NW = ((C::RB / 2) * GROUPS + C::NT - 1) / C::NT;
    uint32_t raw[NW][2];
    auto fetch = [&](int ph) {
        const int r0 = ys - L + ph * C::RB;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int idx = tid + k * C::NT;
            uint32_t u[2] = { 0, 0 };
            if (idx < (C::RB / 2) * GROUPS) {
                const int rp = idx / GROUPS;
                const int g = idx - rp * GROUPS;
                const int xg = x0 - C::LAL + 4 * g;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int y = clampi(r0 + 2 * rp + q, 0, p.h - 1);
                    const uint8_t* row = Iz + (long long)y * p.w;
                    if (p.vec_ok && xg >= 0 && xg + 3 < p.w) {
                        u[q] = __ldg(reinterpret_cast<const uint32_t*>(row + xg));
                    } else {
                        u[q] = (uint32_t)__ldg(row + clampi(xg + 0, 0, p.w - 1)) |
                               ((uint32_t)__ldg(row + clampi(xg + 1, 0, p.w - 1)) << 8) |
                               ((uint32_t)__ldg(row + clampi(xg + 2, 0, p.w - 1)) << 16) |
                               ((uint32_t)__ldg(row + clampi(xg + 3, 0, p.w - 1)) << 24);
                    }
                }
            }
            raw[k][0] = u[0]; raw[k][1] = u[1];
        }
    };
    fetch(0);

    const int xrp = tid & 7;    // x pass: row pair
    const int xch = tid >> 3;   // x pass: 8-column chunk
    const int ycp = tid & 127;  // y pass: column pair
    const int yh = tid >> 7;    // y pass: upper / lower 8 rows of the batch

    for (int phase = 0; phase < nphases; ++phase) {
        __syncthreads();  // s_in2 and the ring block about to be overwritten are no longer read
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int idx = tid + k * C::NT;
            if (idx < (C::RB / 2) * GROUPS) {
                const int rp = idx / GROUPS;
                const int g = idx - rp * GROUPS;
                const uint32_t a = raw[k][0], b = raw[k][1];
                float* dst = s_in2 + rp * C::PP + 8 * g;
                *reinterpret_cast<float4*>(dst) = make_float4((float)(a & 0xffu), (float)(b & 0xffu),
                                                              (float)((a >> 8) & 0xffu), (float)((b >> 8) & 0xffu));
                *reinterpret_cast<float4*>(dst + 4) = make_float4((float)((a >> 16) & 0xffu), (float)((b >> 16) & 0xffu),
                                                                  (float)(a >> 24), (float)(b >> 24));
            }
        }
        __syncthreads();
        if (phase + 1 < nphases) fetch(phase + 1);   // in flight during the x and y passes below
        // ---- x pass: rows 2*xrp, 2*xrp+1, columns 8*xch .. 8*xch+7 ----
        {
            float2 acc[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[o] = make_float2(0.0f, 0.0f);
            const float4* src = reinterpret_cast<const float4*>(s_in2 + xrp * C::PP + 16 * xch);
#pragma unroll
            for (int i2 = 0; i2 < (8 + 2 * C::LAL) / 2; ++i2) {
                const float4 q = src[i2];
                const float2 e[2] = { make_float2(q.x, q.y), make_float2(q.z, q.w) };
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const int i = 2 * i2 + s2;   // window index: column = x0 + 8*xch + i - LAL
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const int t = i - o - C::LAL + L;
                        if (t >= 0 && t <= 2 * L) acc[o] = __ffma2_rn(e[s2], make_float2(taps.g[t], taps.g[t]), acc[o]);
                    }
                }
            }
            float* d0 = s_ring + ((phase % C::NBLK) * C::RB + 2 * xrp) * C::PR + 8 * xch;
            *reinterpret_cast<float4*>(d0) = make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
            *reinterpret_cast<float4*>(d0 + 4) = make_float4(acc[4].x, acc[5].x, acc[6].x, acc[7].x);
            *reinterpret_cast<float4*>(d0 + C::PR) = make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
            *reinterpret_cast<float4*>(d0 + C::PR + 4) = make_float4(acc[4].y, acc[5].y, acc[6].y, acc[7].y);
        }
        const int b = phase - (C::NBLK - 1);     // the batch whose 16 + 2L ring rows are now complete
        if (b < 0) continue;
        __syncthreads();
        // ---- y pass: columns 2*ycp, 2*ycp+1, output rows ys + b*RB + 8*yh + [0, 8) ----
        // (yh is warp-uniform; the two instantiations keep every ring row index static)
        if (yh) xy_fma_ypass<L, 1>(s_ring, taps, b, ycp, Oz, x0, ys, ye, p.w, p.fpitch);
        else xy_fma_ypass<L, 0>(s_ring, taps, b, ycp, Oz, x0, ys, ye, p.w, p.fpitch);
    }
}

// ---------------------------------------------------------------------------
// K1, FMA mode, warp-autonomous form (radii <= 18): the CTA-wide form above spends its time in barriers at small
// radii (three __syncthreads per 16-row batch: 3.4 barrier stalls per issue at sigma = 2, FMA pipe 48 % busy).  Here a
// WARP owns a 64-column strip of one plane and marches down its y segment on its own: private staging tile and
// private ring of x-passed rows, __syncwarp only.  Same register blocking, twice as deep: in the x pass a lane
// makes 16 columns x 2 rows (the rows paired in packed registers, staged rows interleaved in pairs so that a 128-bit
// shared load yields two ready pairs), in the y pass 16 rows x 2 adjacent columns; every shared-memory value feeds up
// to 16 FFMA2.  Arithmetic per output is the same ascending chain of fused multiply-adds as in gauss_xy_fma_kernel:
// bit-identical results.  One CTA per SM, as many warps as the per-warp tiles allow (15 / 11 / 8 at radius 6 / 12 / 18).
// ---------------------------------------------------------------------------
template <int L>
struct XYWarpCfg {
    static constexpr int WS = 64;                              // columns per warp strip
    static constexpr int RB = 16;
    static constexpr int LAL = (L + 3) / 4 * 4;
    static constexpr int PIN0 = WS + 2 * LAL;                  // staged columns per row
    static constexpr int PINP = ((PIN0 + 13) / 16) * 16 + 2;   // >= PIN0 and = 2 mod 16
    static constexpr int PP = 2 * PINP;                        // floats per staged row pair: = 4 mod 32 (conflict-free LDS.128)
    static constexpr int PR = WS + 4;                          // ring row pitch
    static constexpr int NBLK = 1 + (2 * L + RB - 1) / RB;
    static constexpr int WARP_FLOATS = (RB / 2) * PP + NBLK * RB * PR;
    static constexpr int NW0 = (220 * 1024) / (WARP_FLOATS * 4);
    static constexpr int NW = NW0 > 16 ? 16 : NW0;             // warps per CTA (one CTA per SM)
    static constexpr int SMEM_BYTES = NW * WARP_FLOATS * 4;
};

struct XYWarpParams {
    XYParams b;          // I, out, w, h, nz, fpitch, fplane, seg_h, nsegs, vec_ok, zsplit, zgap; nstrips = strips of 64 columns
    long long items;     // strips x segments x planes
};

template <int L>
__global__ void __launch_bounds__(32 * XYWarpCfg<L>::NW, 1)
gauss_xy_warp_kernel(const __grid_constant__ XYWarpParams pp, const __grid_constant__ GaussTaps taps)
{
    using C = XYWarpCfg<L>;
    extern __shared__ __align__(16) float smem[];
    const XYParams& p = pp.b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long item = (long long)blockIdx.x * C::NW + warp;
    if (item >= pp.items) return;                              // warp-uniform; there is no CTA-wide barrier in this kernel
    float* s_in2 = smem + warp * C::WARP_FLOATS;
    float* s_ring = s_in2 + (C::RB / 2) * C::PP;
    const int strip = (int)(item % p.nstrips);
    const int seg = (int)((item / p.nstrips) % p.nsegs);
    const int zl = (int)(item / ((long long)p.nstrips * p.nsegs));
    const int z = zl >= p.zsplit ? zl + p.zgap : zl;
    const int x0 = strip * C::WS;
    const int ys = seg * p.seg_h;
    const int ye = min(ys + p.seg_h, p.h);
    const uint8_t* __restrict__ Iz = p.I + (long long)z * p.w * p.h;
    float* __restrict__ Oz = p.out + (long long)z * p.fplane;
    const int nb = (ye - ys + C::RB - 1) / C::RB;
    const int nphases = nb + C::NBLK - 1;

    // staging: a lane fetches the same 4 columns of both rows of a pair (two 32-bit words per unit)
    constexpr int GROUPS = C::PIN0 / 4;
    constexpr int UNITS = (C::RB / 2) * GROUPS;
    constexpr int NU = (UNITS + 31) / 32;
    uint32_t raw[NU][2];
    auto fetch = [&](int ph) {
        const int r0 = ys - L + ph * C::RB;
#pragma unroll
        for (int k = 0; k < NU; ++k) {
            const int idx = lane + 32 * k;
            uint32_t u[2] = { 0, 0 };
            if (idx < UNITS) {
                const int rp = idx / GROUPS;
                const int g = idx - rp * GROUPS;
                const int xg = x0 - C::LAL + 4 * g;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int y = clampi(r0 + 2 * rp + q, 0, p.h - 1);
                    const uint8_t* row = Iz + (long long)y * p.w;
                    if (p.vec_ok) {
                        // w is a multiple of 4 and so is xg: a group lies wholly inside the row or wholly beyond one end,
                        // where all four pixels are the end pixel (replicate clamp) -- ONE load either way, and nothing
                        // consumes it here (the replication is done at staging time), so the prefetch stays in flight
                        if (xg >= 0 && xg + 3 < p.w) u[q] = __ldg(reinterpret_cast<const uint32_t*>(row + xg));
                        else u[q] = (uint32_t)__ldg(row + (xg < 0 ? 0 : p.w - 1));
                    } else {
                        u[q] = (uint32_t)__ldg(row + clampi(xg + 0, 0, p.w - 1)) |
                               ((uint32_t)__ldg(row + clampi(xg + 1, 0, p.w - 1)) << 8) |
                               ((uint32_t)__ldg(row + clampi(xg + 2, 0, p.w - 1)) << 16) |
                               ((uint32_t)__ldg(row + clampi(xg + 3, 0, p.w - 1)) << 24);
                    }
                }
            }
            raw[k][0] = u[0]; raw[k][1] = u[1];
        }
    };
    fetch(0);

    const int xrp = lane & 7;     // x pass: row pair
    const int xch = lane >> 3;    // x pass: 16-column chunk
    for (int phase = 0; phase < nphases; ++phase) {
        __syncwarp();             // the staging tile and the ring block about to be overwritten are no longer read
#pragma unroll
        for (int k = 0; k < NU; ++k) {
            const int idx = lane + 32 * k;
            if (idx < UNITS) {
                const int rp = idx / GROUPS;
                const int g = idx - rp * GROUPS;
                uint32_t a = raw[k][0], b = raw[k][1];
                const int xg = x0 - C::LAL + 4 * g;
                if (p.vec_ok && !(xg >= 0 && xg + 3 < p.w)) { a *= 0x01010101u; b *= 0x01010101u; }   // end pixel replicated
                float* dst = s_in2 + rp * C::PP + 8 * g;
                *reinterpret_cast<float4*>(dst) = make_float4((float)(a & 0xffu), (float)(b & 0xffu),
                                                              (float)((a >> 8) & 0xffu), (float)((b >> 8) & 0xffu));
                *reinterpret_cast<float4*>(dst + 4) = make_float4((float)((a >> 16) & 0xffu), (float)((b >> 16) & 0xffu),
                                                                  (float)(a >> 24), (float)(b >> 24));
            }
        }
        __syncwarp();
        if (phase + 1 < nphases) fetch(phase + 1);   // in flight during the x and y passes below
        // ---- x pass: rows 2*xrp, 2*xrp+1, columns 16*xch .. 16*xch+15 ----
        {
            float2 acc[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[o] = make_float2(0.0f, 0.0f);
            const float4* src = reinterpret_cast<const float4*>(s_in2 + xrp * C::PP + 32 * xch);
#pragma unroll
            for (int i2 = 0; i2 < (16 + 2 * C::LAL) / 2; ++i2) {
                const float4 q = src[i2];
                const float2 e[2] = { make_float2(q.x, q.y), make_float2(q.z, q.w) };
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const int i = 2 * i2 + s2;   // window index: column = x0 + 16*xch + i - LAL
#pragma unroll
                    for (int o = 0; o < 16; ++o) {
                        const int t = i - o - C::LAL + L;
                        if (t >= 0 && t <= 2 * L) acc[o] = __ffma2_rn(e[s2], make_float2(taps.g[t], taps.g[t]), acc[o]);
                    }
                }
            }
            float* d0 = s_ring + ((phase % C::NBLK) * C::RB + 2 * xrp) * C::PR + 16 * xch;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                *reinterpret_cast<float4*>(d0 + 4 * q) = make_float4(acc[4 * q].x, acc[4 * q + 1].x, acc[4 * q + 2].x, acc[4 * q + 3].x);
                *reinterpret_cast<float4*>(d0 + C::PR + 4 * q) = make_float4(acc[4 * q].y, acc[4 * q + 1].y, acc[4 * q + 2].y, acc[4 * q + 3].y);
            }
        }
        const int b = phase - (C::NBLK - 1);     // the batch whose 16 + 2L ring rows are now complete
        if (b < 0) continue;
        __syncwarp();
        // ---- y pass: columns 2*lane, 2*lane+1, output rows ys + b*RB + [0, 16) ----
        {
            float2 acc[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[o] = make_float2(0.0f, 0.0f);
            const float* bp[C::NBLK];
            int blk = b % C::NBLK;
#pragma unroll
            for (int q = 0; q < C::NBLK; ++q) {
                bp[q] = s_ring + blk * C::RB * C::PR + 2 * lane;
                blk = (blk + 1 == C::NBLK) ? 0 : blk + 1;
            }
#pragma unroll
            for (int jj = 0; jj < C::RB + 2 * L; ++jj) {
                const float2 v = *reinterpret_cast<const float2*>(bp[jj / C::RB] + (jj % C::RB) * C::PR);
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    const int t = jj - o;
                    if (t >= 0 && t <= 2 * L) acc[o] = __ffma2_rn(v, make_float2(taps.g[t], taps.g[t]), acc[o]);
                }
            }
            const int x = x0 + 2 * lane;
            const int ybase = ys + b * C::RB;
            if (x + 1 < p.w) {
#pragma unroll
                for (int o = 0; o < 16; ++o)
                    if (ybase + o < ye) *reinterpret_cast<float2*>(Oz + (long long)(ybase + o) * p.fpitch + x) = acc[o];
            } else if (x < p.w) {
#pragma unroll
                for (int o = 0; o < 16; ++o)
                    if (ybase + o < ye) Oz[(long long)(ybase + o) * p.fpitch + x] = acc[o].x;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K2: z Gaussian pass.  One thread per (x, y) column and chunk of RZ=8 output
// planes; the 8+2LZ input planes are read with plane stride (coalesced along
// x) and replicate-clamped against the GLOBAL volume ends (never at slab
// faces: halo planes are resident).  Blocks are ordered z-chunk fastest so the
// window overlap between neighbouring chunks is served by L2.
// ---------------------------------------------------------------------------
struct ZParams {
    const float* in;    // Fxy, plane 0 = global plane in_base
    float* out;         // F,   plane 0 = global plane out_base
    int w, h, l;        // global dims
    int fpitch;
    long long fplane;
    int in_base;
    int in_count;       // planes resident in `in`
    int out_base;
    int out_count;      // planes to produce
    int nzc, nxs;       // z chunks, x strips
    int zchunk;         // output planes per chunk (marching form)
    int tm_base;        // gauss_z_tma_kernel: global plane of plane 0 of the tensor map (the slab's Fxy buffer)
};

template <int LZ, bool EXACT>
__global__ void __launch_bounds__(128)
gauss_z_kernel(const __grid_constant__ ZParams p, const __grid_constant__ GaussTaps taps)
{
    constexpr int RZ = 8;
    const long long bid = blockIdx.x;
    const int zc = (int)(bid % p.nzc);
    const long long rest = bid / p.nzc;
    const int xs = (int)(rest % p.nxs);
    const int y = (int)(rest / p.nxs);
    const int x = xs * 128 + threadIdx.x;
    if (x >= p.w) return;
    const int zg0 = p.out_base + zc * RZ;
    const float* __restrict__ col = p.in + (long long)y * p.fpitch + x;
    float acc[RZ];
#pragma unroll
    for (int o = 0; o < RZ; ++o) acc[o] = 0.0f;
#pragma unroll
    for (int j = 0; j < RZ + 2 * LZ; ++j) {
        // clamp to the volume (replicate border), then to the resident planes (only
        // reached by window entries that feed outputs beyond out_count, which are dropped)
        const int zs = clampi(clampi(zg0 - LZ + j, 0, p.l - 1) - p.in_base, 0, p.in_count - 1);
        const float v = __ldg(col + (long long)zs * p.fplane);
#pragma unroll
        for (int o = 0; o < RZ; ++o) {
            const int t = j - o;
            if (t >= 0 && t <= 2 * LZ) acc[o] = mac<EXACT>(acc[o], v, taps.g[t]);
        }
    }
    float* __restrict__ dst = p.out + (long long)y * p.fpitch + x;
#pragma unroll
    for (int o = 0; o < RZ; ++o) {
        const int zo = zc * RZ + o;
        if (zo < p.out_count) dst[(long long)zo * p.fplane] = acc[o];
    }
}

// ---------------------------------------------------------------------------
// K2 (marching form, radius <= 12): z Gaussian pass as a register sliding
// window.  A thread owns two adjacent x columns and marches along z through
// its chunk; the 2LZ+1 inputs of the current output (plus PF planes of
// prefetch, so that several loads per thread are in flight) live in a register
// ring whose indices are static because the march is unrolled by the ring
// period.  Every input plane is read once per chunk and every output needs one
// 8-byte load, one 8-byte store and 2LZ+1 multiply-adds: the pass runs at the
// HBM rate.  Accumulation order is the reference's (ascending taps from zero,
// frangi.cpp:756-768), replicate clamping against the GLOBAL volume ends only.
// ---------------------------------------------------------------------------
#ifndef ZM_THREADS
#define ZM_THREADS 128
#endif
#ifndef ZM_PF
#define ZM_PF 3
#endif
#ifndef ZM_MINB
#define ZM_MINB 1
#endif
template <int LZ, bool EXACT>
__global__ void __launch_bounds__(ZM_THREADS)
gauss_z_march_kernel(const __grid_constant__ ZParams p, const __grid_constant__ GaussTaps taps)
{
    constexpr int PF = ZM_PF;
    constexpr int Q = 2 * LZ + 1 + PF;            // ring period
    const long long bid = blockIdx.x;
    const int xs = (int)(bid % p.nxs);
    const long long rest = bid / p.nxs;
    const int y = (int)(rest % p.h);
    const int zc = (int)(rest / p.h);
    const int x = xs * (2 * ZM_THREADS) + 2 * threadIdx.x;
    if (x >= p.w) return;
    const int t_begin = zc * p.zchunk;             // first output plane (relative to out_base) of this chunk
    const int nout = min(p.zchunk, p.out_count - t_begin);
    const int zg0 = p.out_base + t_begin;          // global plane of the first output
    const float2* __restrict__ col = reinterpret_cast<const float2*>(p.in + (long long)y * p.fpitch + x);
    float2* __restrict__ dst = reinterpret_cast<float2*>(p.out + ((long long)t_begin * p.fplane + (long long)y * p.fpitch + x));
    const long long plane2 = p.fplane / 2;         // float2 units per plane (fpitch is even)
    auto ld = [&](int j) {                         // input j of the chunk = global plane zg0 - LZ + j, clamped
        const int zs = clampi(clampi(zg0 - LZ + j, 0, p.l - 1) - p.in_base, 0, p.in_count - 1);
        return __ldcs(col + (long long)zs * plane2);
    };
    float2 win[Q];
#pragma unroll
    for (int j = 0; j < Q - 1; ++j) win[j] = ld(j);
    for (int t0 = 0; t0 < nout; t0 += Q) {
#pragma unroll
        for (int s = 0; s < Q; ++s) {
            const int t = t0 + s;
            if (t < nout) {
                win[(s + Q - 1) % Q] = ld(t + Q - 1);
                float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int k = 0; k <= 2 * LZ; ++k) {
                    const float2 v = win[(s + k) % Q];
                    if (EXACT) {
                        acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, taps.g[k]));
                        acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, taps.g[k]));
                    } else {
                        acc = __ffma2_rn(v, make_float2(taps.g[k], taps.g[k]), acc);
                    }
                }
                __stcs(dst + (long long)t * plane2, acc);
            }
        }
    }
}

// ---- TMA tile staging (cp.async.bulk.tensor + mbarrier) ----------------------------------------
// The K3 kernels stage one (PW x PH x 1) box of the smoothed volume F per plane with a single
// instruction issued by one thread; the copy engine zero-fills whatever lies outside the tensor
// (x, y < 0, x >= w, y >= h), which is exactly what the voxels next to a volume face need (they
// never read those entries).  Completion is signalled on a shared-memory mbarrier per ring slot.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#ifndef MBAR_HINT_NS
#define MBAR_HINT_NS 0        // > 0: suspend-time hint of try_wait (measured at 20000: no effect on any kernel)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
#if MBAR_HINT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"((uint32_t)MBAR_HINT_NS) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)      // non-blocking: has the phase completed?
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// The ring of TILE::SLOTS plane tiles of a z-marching CTA.  Plane `zs - 2 + n` lives in slot n % SLOTS;
// its mbarrier completes phase n / SLOTS.  Planes beyond a z face are not fetched (the face rules never
// read them) but still complete their phase, so the parity bookkeeping is uniform.  o[0..4] are the
// float offsets of planes z-2 .. z+2, rotated once per plane (no modulo in the loop).
template <class TILE>
struct TileRing {
    uint32_t ring_s, mbar_s;
    int o[5];
    uint32_t wbar, wpar;          // barrier / parity of the next plane to wait for (plane z + 2)
    uint32_t ibar, idst;          // thread 0: barrier / destination of the next plane to issue
    __device__ __forceinline__ void init(float* ring, uint64_t* mbar, int tid)
    {
        ring_s = smem_u32(ring); mbar_s = smem_u32(mbar);
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < TILE::SLOTS; ++k) mbar_init(mbar_s + 8 * k, 1);
            mbar_init_fence();
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) o[k] = k * TILE::SLOT;
        wbar = mbar_s; wpar = 0; ibar = mbar_s; idst = ring_s;
    }
    // thread 0 only: start the copy of `plane` (tile origin x0, y0) into the next slot
    __device__ __forceinline__ void issue(const CUtensorMap* tm, int x0, int y0, int plane, int base, int l)
    {
        if (plane < 0 || plane > l - 1) mbar_arrive(ibar);
        else {
            mbar_expect_tx(ibar, TILE::PLANE * 4);
            tma_load_3d(idst, tm, ibar, x0, y0, plane - base);
        }
        ibar += 8; idst += TILE::SLOT * 4;
        if (ibar == mbar_s + 8 * TILE::SLOTS) { ibar = mbar_s; idst = ring_s; }
    }
    __device__ __forceinline__ void wait_next()
    {
        mbar_wait(wbar, wpar);
        wbar += 8;
        if (wbar == mbar_s + 8 * TILE::SLOTS) { wbar = mbar_s; wpar ^= 1u; }
    }
    __device__ __forceinline__ void rotate()
    {
        o[0] = o[1]; o[1] = o[2]; o[2] = o[3]; o[3] = o[4];
        o[4] += TILE::SLOT;
        if (o[4] == TILE::SLOTS * TILE::SLOT) o[4] = 0;
    }
};

// ---------------------------------------------------------------------------
// K2 (TMA form, radius <= 12): the same register sliding window as gauss_z_march_kernel, but the
// loads are bulk copies.  A WARP owns 128 adjacent x columns of one row and marches along z; lane 0
// keeps DEPTH planes of the warp's 512-byte row segment in flight with cp.async.bulk.tensor on a
// (w, h, planes) tensor map of Fxy (box 128 x 1 x 1, completion on one mbarrier per slot), every lane
// picks its four columns out of the landed segment with one 128-bit shared load and pushes them into
// its window (two independent FFMA2 chains).  Bytes in flight no longer depend on registers or on
// resident threads (DEPTH x 512 B per warp), so the kernel runs at the HBM rate at every radius.
// (With 64 columns per warp the per-plane bookkeeping -- barrier wait, copy issue, slot rotation --
// made the kernel issue-bound at 80 % issue utilisation, profiles/r1m.)
// Warps are independent (no __syncthreads): slot reuse is ordered by __syncwarp.
// Arithmetic and accumulation order are those of gauss_z_march_kernel (bit-identical results).
// ---------------------------------------------------------------------------
#ifndef ZT_DEPTH
#define ZT_DEPTH 16
#endif
#ifndef ZT_WARPS
#define ZT_WARPS 2
#endif
struct ZTile {
    static constexpr int COLS = 128, DEPTH = ZT_DEPTH, WARPS = ZT_WARPS;   // a lane owns 4 adjacent columns
    static constexpr int SMEM_BYTES = WARPS * (DEPTH * (COLS * 4 + 16) + 128);  // rings, one full and one empty mbarrier per slot, one scratch word per lane
};

template <int LZ, bool EXACT>
__global__ void __launch_bounds__(32 * ZTile::WARPS, 16 / ZTile::WARPS)      // up to 128 registers: no spills at radius 9, 12
gauss_z_tma_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ ZParams p, const __grid_constant__ GaussTaps taps)
{
    constexpr int Q = 2 * LZ + 1, D = ZTile::DEPTH;
    extern __shared__ __align__(128) float zring[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * ZTile::WARPS + warp;     // (x strip, row, z chunk)
    const int xs = (int)(item % p.nxs);
    const long long rest = item / p.nxs;
    const int y = (int)(rest % p.h);
    const int zc = (int)(rest / p.h);
    if (zc >= p.nzc) return;                                   // warp-uniform
    const int x0 = xs * ZTile::COLS, x = x0 + 4 * lane;
    const int t_begin = zc * p.zchunk;
    const int nout = min(p.zchunk, p.out_count - t_begin);
    const int zg0 = p.out_base + t_begin;
    const int nin = nout + 2 * LZ;                             // inputs of the chunk: global planes zg0 - LZ + j, clamped
    const float* ring = zring + warp * D * ZTile::COLS;
    const uint32_t ring_s = smem_u32(ring);
    const uint32_t bar_s = smem_u32(zring + ZTile::WARPS * D * ZTile::COLS) + warp * D * 8;
    // one "empty" mbarrier per slot next to the "full" one: the consumer-release half of the producer / consumer pair
    const uint32_t empty_s = smem_u32(zring + ZTile::WARPS * D * ZTile::COLS) + ZTile::WARPS * D * 8 + warp * D * 8;
    const uint32_t scratch_s = smem_u32(zring + ZTile::WARPS * D * ZTile::COLS) + 2 * ZTile::WARPS * D * 8 + warp * 128 + lane * 4;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) { mbar_init(bar_s + 8 * k, 1); mbar_init(empty_s + 8 * k, 32); }
        mbar_init_fence();
    }
    __syncwarp();
    // replicate clamp at the volume ends (frangi.cpp:758,776), then into the resident planes (only reached by window
    // entries that feed outputs beyond out_count, which are never produced); as tensor-map plane coordinates
    const int zlo = max(0, p.in_base) - p.tm_base, zhi = min(p.l - 1, p.in_base + p.in_count - 1) - p.tm_base;
    int zj = zg0 - LZ - p.tm_base;                             // lane 0: unclamped coordinate of the next copy
    uint32_t ibar = bar_s, idst = ring_s;                      // lane 0: barrier / destination of the next copy
    auto issue = [&]() {
        mbar_expect_tx(ibar, ZTile::COLS * 4);
        tma_load_3d(idst, &tm, ibar, x0, y, min(max(zj, zlo), zhi));
        ++zj; ibar += 8; idst += ZTile::COLS * 4;
        if (ibar == bar_s + 8 * D) { ibar = bar_s; idst = ring_s; }
    };
    if (lane == 0)
        for (int j = 0; j < min(D, nin); ++j) issue();
    int slot = 0;
    uint32_t par = 0;
    auto next = [&](int j) {                                   // input j of the chunk, this lane's four columns
        mbar_wait(bar_s + 8 * slot, par);
        const float4 v = *reinterpret_cast<const float4*>(ring + slot * ZTile::COLS + 4 * lane);
        // Slot reuse is ordered by a full / empty barrier pair: every lane ARRIVES on the slot's empty barrier once its
        // load has DELIVERED and lane 0 WAITS for all 32 arrivals before it arms the full barrier and issues the refill.
        // "Delivered" needs care: the value is first used many instructions later, and neither program order nor the
        // release semantics of the arrive make the hardware hold the arrive back until the load's data has returned
        // (measured twice: a refill issued right after the load INSTRUCTION -- by plain program order in round 1, behind
        // an arrive that directly follows the load in round 2 -- lands, served from L2, while a quarter-warp of the load
        // is still queued in the memory pipeline: rare 16-byte corruptions, tools/debug_streamed.py).  So the arrive is
        // made data-dependent on the load: a store of the loaded value cannot issue before the data is in the
        // register, and the arrive follows it in the same in-order memory pipe.
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(scratch_s), "f"(v.w) : "memory");
        mbar_arrive(empty_s + 8 * slot);
        if (lane == 0 && j + D < nin) {
            mbar_wait(empty_s + 8 * slot, par);                // use k of the slot completes phase k of both barriers
            issue();
        }
        if (++slot == D) { slot = 0; par ^= 1u; }
        return v;
    };
    float* __restrict__ dst = p.out + ((long long)t_begin * p.fplane + (long long)y * p.fpitch + x);
    const bool store = x < p.w;                                // the row pitch is a multiple of 32 floats: the quad fits
    float4 win[Q];
#pragma unroll
    for (int j = 0; j < Q - 1; ++j) win[j] = next(j);
    for (int t0 = 0; t0 < nout; t0 += Q) {
#pragma unroll
        for (int s = 0; s < Q; ++s) {
            const int t = t0 + s;
            if (t < nout) {
                win[(s + Q - 1) % Q] = next(t + Q - 1);
                float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int k = 0; k <= 2 * LZ; ++k) {
                    const float4 v = win[(s + k) % Q];
                    if (EXACT) {
                        a0.x = __fadd_rn(a0.x, __fmul_rn(v.x, taps.g[k])); a0.y = __fadd_rn(a0.y, __fmul_rn(v.y, taps.g[k]));
                        a1.x = __fadd_rn(a1.x, __fmul_rn(v.z, taps.g[k])); a1.y = __fadd_rn(a1.y, __fmul_rn(v.w, taps.g[k]));
                    } else {
                        const float2 g2 = make_float2(taps.g[k], taps.g[k]);
                        a0 = __ffma2_rn(make_float2(v.x, v.y), g2, a0);
                        a1 = __ffma2_rn(make_float2(v.z, v.w), g2, a1);
                    }
                }
                if (store) __stcs(reinterpret_cast<float4*>(dst), make_float4(a0.x, a0.y, a1.x, a1.y));
                dst += p.fplane;
            }
        }
    }
}


// ---------------------------------------------------------------------------
// Second differences with the reference's face rules (frangi.cpp:306-381):
// first difference along an axis = s * (f[hi] - f[lo]) with lo = max(c-1,0),
// hi = min(c+1,n-1), s = 1 on a face and 0.5 inside; the second difference
// applies the same rule to the first-difference field; then * sigma^2.
// Coordinates are GLOBAL (slab faces are not volume faces).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float face_scale(int c, int n) { return (c == 0 || c == n - 1) ? 1.0f : 0.5f; }

struct Hess { float xx, xy, xz, yy, yz, zz; };

// Generic form for voxels within two steps of a volume face.  `Field` provides
// at(x, y, z) for any coordinate inside the volume and within +-2 of the voxel.
template <class Field>
__device__ __forceinline__ float d_dx(const Field& f, int x, int y, int z)
{
    return __fmul_rn(face_scale(x, f.w), __fsub_rn(f.at(min(x + 1, f.w - 1), y, z), f.at(max(x - 1, 0), y, z)));
}
template <class Field>
__device__ __forceinline__ float d_dy(const Field& f, int x, int y, int z)
{
    return __fmul_rn(face_scale(y, f.h), __fsub_rn(f.at(x, min(y + 1, f.h - 1), z), f.at(x, max(y - 1, 0), z)));
}
template <class Field>
__device__ __forceinline__ float d_dz(const Field& f, int x, int y, int z)
{
    return __fmul_rn(face_scale(z, f.l), __fsub_rn(f.at(x, y, min(z + 1, f.l - 1)), f.at(x, y, max(z - 1, 0))));
}

// All 24 taps are loaded before the first difference is formed, so that a thread waits for memory once, not once per
// group of terms (the shell kernel is bound by the latency of these loads; same operations on the same operands as
// the nested d_dx / d_dy / d_dz form above).
template <class Field>
__device__ __noinline__ Hess hessian_at_face(const Field& f, int x, int y, int z, float sigma2)
{
    const int W = f.w - 1, Hh = f.h - 1, L = f.l - 1;
    const int xl = max(x - 1, 0), xh = min(x + 1, W);
    const int yl = max(y - 1, 0), yh = min(y + 1, Hh);
    const int zl = max(z - 1, 0), zh = min(z + 1, L);
    const int xm = max(x - 1, 0), xp = min(x + 1, W), ym = max(y - 1, 0), yp = min(y + 1, Hh), zm = max(z - 1, 0), zp = min(z + 1, L);
    // d/dx at (xh, y, z), (xl, y, z), (x, yh, z), (x, yl, z), (x, y, zh), (x, y, zl): hi and lo tap of each
    const float x0h = f.at(min(xh + 1, W), y, z), x0l = f.at(max(xh - 1, 0), y, z);
    const float x1h = f.at(min(xl + 1, W), y, z), x1l = f.at(max(xl - 1, 0), y, z);
    const float x2h = f.at(xp, yh, z), x2l = f.at(xm, yh, z);
    const float x3h = f.at(xp, yl, z), x3l = f.at(xm, yl, z);
    const float x4h = f.at(xp, y, zh), x4l = f.at(xm, y, zh);
    const float x5h = f.at(xp, y, zl), x5l = f.at(xm, y, zl);
    // d/dy at (x, yh, z), (x, yl, z), (x, y, zh), (x, y, zl)
    const float y0h = f.at(x, min(yh + 1, Hh), z), y0l = f.at(x, max(yh - 1, 0), z);
    const float y1h = f.at(x, min(yl + 1, Hh), z), y1l = f.at(x, max(yl - 1, 0), z);
    const float y2h = f.at(x, yp, zh), y2l = f.at(x, ym, zh);
    const float y3h = f.at(x, yp, zl), y3l = f.at(x, ym, zl);
    // d/dz at (x, y, zh), (x, y, zl)
    const float z0h = f.at(x, y, min(zh + 1, L)), z0l = f.at(x, y, max(zh - 1, 0));
    const float z1h = f.at(x, y, min(zl + 1, L)), z1l = f.at(x, y, max(zl - 1, 0));
    const float sx = face_scale(x, f.w), sy = face_scale(y, f.h), sz = face_scale(z, f.l);
    const float sxh = face_scale(xh, f.w), sxl = face_scale(xl, f.w);
    const float syh = face_scale(yh, f.h), syl = face_scale(yl, f.h);
    const float szh = face_scale(zh, f.l), szl = face_scale(zl, f.l);
    Hess H;
    H.xx = __fmul_rn(__fmul_rn(sx, __fsub_rn(__fmul_rn(sxh, __fsub_rn(x0h, x0l)), __fmul_rn(sxl, __fsub_rn(x1h, x1l)))), sigma2);
    H.xy = __fmul_rn(__fmul_rn(sy, __fsub_rn(__fmul_rn(sx, __fsub_rn(x2h, x2l)), __fmul_rn(sx, __fsub_rn(x3h, x3l)))), sigma2);
    H.xz = __fmul_rn(__fmul_rn(sz, __fsub_rn(__fmul_rn(sx, __fsub_rn(x4h, x4l)), __fmul_rn(sx, __fsub_rn(x5h, x5l)))), sigma2);
    H.yy = __fmul_rn(__fmul_rn(sy, __fsub_rn(__fmul_rn(syh, __fsub_rn(y0h, y0l)), __fmul_rn(syl, __fsub_rn(y1h, y1l)))), sigma2);
    H.yz = __fmul_rn(__fmul_rn(sz, __fsub_rn(__fmul_rn(sy, __fsub_rn(y2h, y2l)), __fmul_rn(sy, __fsub_rn(y3h, y3l)))), sigma2);
    H.zz = __fmul_rn(__fmul_rn(sz, __fsub_rn(__fmul_rn(szh, __fsub_rn(z0h, z0l)), __fmul_rn(szl, __fsub_rn(z1h, z1l)))), sigma2);
    return H;
}

// ---------------------------------------------------------------------------
// K3: Hessian -> eigen -> vesselness -> running max over scales.
//
// Two kernels per scale share the per-voxel stage:
//
// K3a hessian_eigen_kernel (interior, > 98 % of the voxels of a real volume):
// a CTA owns a 128 x 16 column of voxels and marches along z through its chunk.
// The five planes z-2 .. z+2 of the smoothed volume F that the twice-applied
// central difference touches live in a six-slot shared-memory ring of
// (16+4) x (128+4) tiles (TileRing: the plane being staged never aliases one
// being read and one __syncthreads per plane suffices).  The next plane's tile
// is copied global -> shared by TMA (one cp.async.bulk.tensor of thread 0, completion
// on the slot's mbarrier) while the current plane is processed.  A thread produces 4
// consecutive x voxels of 2 rows per plane: window rows are read with 128-bit /
// 64-bit shared loads (6.5 loads per voxel) and results leave as two 64-bit pairs
// (tiles start at x = 2, see the kernel).  Voxels at
// least two steps from every volume face take the closed interior form,
// bit-identical to the reference's two-pass form:
//   Dxx = ((F[x+2]-F[x]) - (F[x]-F[x-2])) * (sigma^2/4),
//   Dxy = ((F[+1,+1]-F[-1,+1]) - (F[+1,-1]-F[-1,-1])) * (sigma^2/4)
// (halving is exact, so the 0.5 factors commute with the roundings of
// frangi.cpp:308-381).  The kernel never touches a voxel within two steps of a
// volume face.
//
// K3b hessian_eigen_shell_kernel: the two-voxel-thick shell next to the volume
// faces (x, y < 2 or > n-3; z likewise), one thread per voxel, generic face
// rules straight from global memory.  Runs after K3a on the same stream.
//
// Outputs are dense over the slab's own planes [z_begin, z_begin + nz).
// MODE 0: first scale, store unconditionally (frangi.cpp:234-252);
// MODE 1: later scale, overwrite only on a strictly greater response (:254-271);
//         BRIGHT = true adds the shortcut ordering for bright ridges (frangi_voxel_math.cuh);
// MODE 2: stage dump of the six second differences (hessian3d parity).
// minmax[0] = bits of min J (taken on the first scale),
// minmax[1] = bits of max J (max over every value any scale leaves in J).  J >= 0, so the int
// order of the bit patterns is the float order.
// ---------------------------------------------------------------------------

struct FView {
    const float* F;     // plane 0 = global plane base
    int w, h, l;
    int fpitch;
    long long fplane;
    int base;           // first resident plane
    int count;          // resident planes
    __device__ __forceinline__ float at(int x, int y, int z) const
    {
        return __ldg(F + (long long)(z - base) * fplane + (long long)y * fpitch + x);
    }
};

#ifndef HESS_MIN_CTAS
#define HESS_MIN_CTAS 3
#endif
#ifndef HESS_RPT
#define HESS_RPT 2      // tile rows per thread (2: 256-thread CTAs, 1: 512-thread CTAs)
#endif
struct HessTile {
    static constexpr int TX = 128, TY = 16, RPT = HESS_RPT, NT = 32 * TY / RPT;
    static constexpr int PW = TX + 4;            // 132 floats per tile row (16-byte multiple)
    static constexpr int PH = TY + 4;
    static constexpr int PLANE = PW * PH;        // 2640 floats: the TMA box
    static constexpr int SLOT = (PLANE * 4 + 127) / 128 * 32;   // ring slot in floats (TMA destinations are 128-byte aligned)
#ifndef K3A_STREAM
#define K3A_STREAM 1        // 1: slot reuse by an empty mbarrier per slot (no __syncthreads in the plane loop, warps drift); 0: round-1 form
#endif
    static constexpr int SLOTS = K3A_STREAM ? 7 : 6;
    static constexpr int RING_BYTES = SLOTS * SLOT * 4;
    static constexpr int SMEM_BYTES = RING_BYTES + 2 * SLOTS * 8 + 16;   // + a full and an empty mbarrier per slot, issue counter
};

#ifndef K3C_SHELL_REM
#define K3C_SHELL_REM 1       // a last tile column of the compacting kernel with <= SHELL_EXTRA_X voxel columns goes to the shell
#endif
static constexpr int SHELL_EXTRA_X = 8;
struct VoxelParams {
    CUtensorMap tmap;     // F as a (w, h, planes) tensor, box = the launching kernel's tile (PW x PH x 1)
    FView f;
    float* J;
    uint8_t* Vx;
    uint8_t* Vy;
    uint8_t* Vz;
    uint8_t* scale_idx;   // nullable
    float* dir;           // nullable, 3 planar volumes of `voxels` floats
    float* D[6];          // MODE 2 only: Dzz, Dyy, Dyz, Dxx, Dxy, Dxz (reference argument order)
    long long voxels;     // own voxels
    int z_begin, nz;      // own planes
    int zchunk;           // planes per CTA along z (K3a)
    int ntx, nty;         // tiles along x and y (K3a)
    int scale;            // index of this scale
    int last_scale;
    int vec_ok;           // w % 2 == 0: the pairs of a quad (xq = 2 mod 4) are aligned for 64-bit / 16-bit vector accesses
    int* minmax;
    FrangiConsts k;
    // K3b: the face coordinates (deduplicated) and the three region sizes
    int xf[4 + SHELL_EXTRA_X], yf[4], zf[4];     // xf: the x faces, then up to SHELL_EXTRA_X columns a tile kernel leaves to the shell
    int nxf, nyf, nzf;
    long long n_zface, n_yface, n_xface;
};

// one voxel's update (scalar stores); returns the value J holds afterwards
template <int MODE>
__device__ __noinline__ float voxel_update(const VoxelParams& p, long long i, const Hess H)
{
    Eig3 e;
    eig_sym3<float>(H.xx, H.xy, H.xz, H.yy, H.yz, H.zz, e);
    const float v = vesselness<float>(e, p.k);
    float jold = 0.0f;
    bool write = MODE == 0;
    if (MODE == 1) { jold = p.J[i]; write = v > jold; }
    if (write) {
        p.J[i] = v;
        p.Vx[i] = (uint8_t)dir_code(e.vx); p.Vy[i] = (uint8_t)dir_code(e.vy); p.Vz[i] = (uint8_t)dir_code(e.vz);
        if (p.scale_idx) p.scale_idx[i] = (uint8_t)p.scale;
        if (p.dir) { p.dir[i] = e.vx; p.dir[p.voxels + i] = e.vy; p.dir[2 * p.voxels + i] = e.vz; }
        return v;
    }
    return jold;
}

// The planes and factors the z direction of the second differences needs at centre plane z,
// identical for every voxel of the plane (so they are set up once per plane, uniformly):
//   zl = max(z-1, 0), zh = min(z+1, l-1); first differences in x / y are taken on planes zl, zh;
//   dz(zh) = sh * (F[clamp(zh+1)] - F[clamp(zh-1)]), dz(zl) = sl * (F[clamp(zl+1)] - F[clamp(zl-1)]),
//   with the face scale 1 on a volume face and 0.5 inside (frangi.cpp:308-310); the second
//   difference applies sz = face scale of z itself (frangi.cpp:315-317), then * sigma^2.
// Two steps away from the z faces this reduces to the closed interior form (general == false).
struct ZPlanes {              // plane fields are float offsets into the shared-memory ring
    int P0;                   // plane z
    int Pzl, Pzh;             // planes zl, zh
    int Pa, Pb;               // planes clamp(zh+1), clamp(zh-1)
    int Pc, Pd;               // planes clamp(zl+1), clamp(zl-1)
    float sh, sl;
    float qs;                 // sigma^2 / 4          in-plane terms
    float qz;                 // 0.5 * sz * sigma^2   xz, yz
    float qzz;                // sz * sigma^2         zz
    bool general;
};

// interior planes (2 <= z <= l-3): only the five ring offsets and sigma^2/4 are needed.
// o[0..4] = ring offsets (floats) of planes z-2 .. z+2 (TileRing::o).
__device__ __forceinline__ ZPlanes z_planes_interior(const int* o, float sigma2)
{
    ZPlanes zp;
    zp.P0 = o[2];
    zp.Pzl = o[1]; zp.Pzh = o[3];
    zp.Pa = o[4]; zp.Pd = o[0];
    zp.Pb = zp.P0; zp.Pc = zp.P0;
    zp.sh = zp.sl = 0.5f;
    zp.qs = 0.25f * sigma2; zp.qz = zp.qs; zp.qzz = 0.5f * sigma2;
    zp.general = false;
    return zp;
}

__device__ __forceinline__ ZPlanes z_planes(const int* o, int z, int l, float sigma2)
{
    // a clamped plane is always one of the five resident ones
    auto plane = [&](int q) {
        const int k = clampi(q, 0, l - 1) - (z - 2);
        return k <= 0 ? o[0] : (k == 1 ? o[1] : (k == 2 ? o[2] : (k == 3 ? o[3] : o[4])));
    };
    const int zl = max(z - 1, 0), zh = min(z + 1, l - 1);
    const float sz = face_scale(z, l);
    ZPlanes zp;
    zp.P0 = plane(z); zp.Pzl = plane(zl); zp.Pzh = plane(zh);
    zp.Pa = plane(zh + 1); zp.Pb = plane(zh - 1); zp.Pc = plane(zl + 1); zp.Pd = plane(zl - 1);
    zp.sh = face_scale(zh, l); zp.sl = face_scale(zl, l);
    zp.qs = 0.25f * sigma2; zp.qz = 0.5f * sz * sigma2; zp.qzz = sz * sigma2;
    zp.general = z < 2 || z > l - 3;
    return zp;
}

// A quad starts at x = 2 mod 4, so it is stored as two 8-byte-aligned pairs (two 2-byte-aligned code pairs).
__device__ __forceinline__ void st_pair(float* dst, const float* v)
{
    *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
    *reinterpret_cast<float2*>(dst + 2) = make_float2(v[2], v[3]);
}
__device__ __forceinline__ void st_codes(uint8_t* dst, uint32_t c)
{
    *reinterpret_cast<uint16_t*>(dst) = (uint16_t)c;
    *reinterpret_cast<uint16_t*>(dst + 2) = (uint16_t)(c >> 16);
}

// Second differences of four consecutive x voxels of one tile row (x and y interior), as two
// packed pairs (voxels 0,1 and 2,3).  `o` = tile entry of (x - 2, y) of the first voxel.
// (hi - mid) - (mid - lo), then * sigma^2/4: FADD2 / FMUL2 round each lane exactly like the scalar
// __fsub_rn / __fmul_rn chain of the reference's two passes (halving is exact and commutes with the
// roundings; no multiply feeds an add, so nothing can fuse).
template <bool ZGEN>
__device__ __forceinline__ void quad_hessians(const float* ring, const ZPlanes& zp, int o, float2* Hxx, float2* Hxy,
                                              float2* Hxz, float2* Hyy, float2* Hyz, float2* Hzz)
{
    using T = HessTile;
    const float* P0 = ring + zp.P0;
    // plane z: rows y-1, y, y+1 over x-2 .. x+5; rows y-2, y+2 over x .. x+3
    float a[8], b[8], c[8];
    *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(P0 + o - T::PW);
    *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(P0 + o - T::PW + 4);
    *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(P0 + o);
    *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(P0 + o + 4);
    *reinterpret_cast<float4*>(c) = *reinterpret_cast<const float4*>(P0 + o + T::PW);
    *reinterpret_cast<float4*>(c + 4) = *reinterpret_cast<const float4*>(P0 + o + T::PW + 4);
    float t2[4], u2[4];
    *reinterpret_cast<float2*>(t2) = *reinterpret_cast<const float2*>(P0 + o - 2 * T::PW + 2);
    *reinterpret_cast<float2*>(t2 + 2) = *reinterpret_cast<const float2*>(P0 + o - 2 * T::PW + 4);
    *reinterpret_cast<float2*>(u2) = *reinterpret_cast<const float2*>(P0 + o + 2 * T::PW + 2);
    *reinterpret_cast<float2*>(u2 + 2) = *reinterpret_cast<const float2*>(P0 + o + 2 * T::PW + 4);
    // own column on the planes of the z second difference
    float pa[4], pd[4], pb[4], pc[4];
    *reinterpret_cast<float2*>(pa) = *reinterpret_cast<const float2*>(ring + zp.Pa + o + 2);
    *reinterpret_cast<float2*>(pa + 2) = *reinterpret_cast<const float2*>(ring + zp.Pa + o + 4);
    *reinterpret_cast<float2*>(pd) = *reinterpret_cast<const float2*>(ring + zp.Pd + o + 2);
    *reinterpret_cast<float2*>(pd + 2) = *reinterpret_cast<const float2*>(ring + zp.Pd + o + 4);
    if (ZGEN) {
        *reinterpret_cast<float2*>(pb) = *reinterpret_cast<const float2*>(ring + zp.Pb + o + 2);
        *reinterpret_cast<float2*>(pb + 2) = *reinterpret_cast<const float2*>(ring + zp.Pb + o + 4);
        *reinterpret_cast<float2*>(pc) = *reinterpret_cast<const float2*>(ring + zp.Pc + o + 2);
        *reinterpret_cast<float2*>(pc + 2) = *reinterpret_cast<const float2*>(ring + zp.Pc + o + 4);
    }
    // planes zl, zh: row y over x-1 .. x+4 (loaded as x-2 .. x+5), rows y-1, y+1 over x .. x+3
    float mm[8], nn[8], mu[4], md[4], nu[4], nd[4];
    *reinterpret_cast<float4*>(mm) = *reinterpret_cast<const float4*>(ring + zp.Pzl + o);
    *reinterpret_cast<float4*>(mm + 4) = *reinterpret_cast<const float4*>(ring + zp.Pzl + o + 4);
    *reinterpret_cast<float4*>(nn) = *reinterpret_cast<const float4*>(ring + zp.Pzh + o);
    *reinterpret_cast<float4*>(nn + 4) = *reinterpret_cast<const float4*>(ring + zp.Pzh + o + 4);
    *reinterpret_cast<float2*>(mu) = *reinterpret_cast<const float2*>(ring + zp.Pzl + o - T::PW + 2);
    *reinterpret_cast<float2*>(mu + 2) = *reinterpret_cast<const float2*>(ring + zp.Pzl + o - T::PW + 4);
    *reinterpret_cast<float2*>(md) = *reinterpret_cast<const float2*>(ring + zp.Pzl + o + T::PW + 2);
    *reinterpret_cast<float2*>(md + 2) = *reinterpret_cast<const float2*>(ring + zp.Pzl + o + T::PW + 4);
    *reinterpret_cast<float2*>(nu) = *reinterpret_cast<const float2*>(ring + zp.Pzh + o - T::PW + 2);
    *reinterpret_cast<float2*>(nu + 2) = *reinterpret_cast<const float2*>(ring + zp.Pzh + o - T::PW + 4);
    *reinterpret_cast<float2*>(nd) = *reinterpret_cast<const float2*>(ring + zp.Pzh + o + T::PW + 2);
    *reinterpret_cast<float2*>(nd + 2) = *reinterpret_cast<const float2*>(ring + zp.Pzh + o + T::PW + 4);
    const float qs = zp.qs, qz = ZGEN ? zp.qz : zp.qs;
    const float2 qs2 = make_float2(qs, qs), qz2 = make_float2(qz, qz);
#define PAIR(arr, i) make_float2((arr)[(i)], (arr)[(i) + 1])
#define DD(hi, mid, lo) vmul(vsub(vsub(hi, mid), vsub(mid, lo)), qs2)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int j = 2 * g;
        const float2 f0 = PAIR(b, j + 2);
        Hxx[g] = DD(PAIR(b, j + 4), f0, PAIR(b, j));
        Hyy[g] = DD(PAIR(u2, j), f0, PAIR(t2, j));
        if (ZGEN)
            Hzz[g] = vmul(vsub(vmul(make_float2(zp.sh, zp.sh), vsub(PAIR(pa, j), PAIR(pb, j))),
                               vmul(make_float2(zp.sl, zp.sl), vsub(PAIR(pc, j), PAIR(pd, j)))),
                          make_float2(zp.qzz, zp.qzz));
        else
            Hzz[g] = DD(PAIR(pa, j), f0, PAIR(pd, j));
        // the x+-1 taps sit at odd register offsets of the 128-bit loads: scalar form, no re-pairing moves
        Hxy[g].x = __fmul_rn(__fsub_rn(__fsub_rn(c[j + 3], c[j + 1]), __fsub_rn(a[j + 3], a[j + 1])), qs);
        Hxy[g].y = __fmul_rn(__fsub_rn(__fsub_rn(c[j + 4], c[j + 2]), __fsub_rn(a[j + 4], a[j + 2])), qs);
        Hxz[g].x = __fmul_rn(__fsub_rn(__fsub_rn(nn[j + 3], nn[j + 1]), __fsub_rn(mm[j + 3], mm[j + 1])), qz);
        Hxz[g].y = __fmul_rn(__fsub_rn(__fsub_rn(nn[j + 4], nn[j + 2]), __fsub_rn(mm[j + 4], mm[j + 2])), qz);
        Hyz[g] = vmul(vsub(vsub(PAIR(nd, j), PAIR(nu, j)), vsub(PAIR(md, j), PAIR(mu, j))), qz2);
    }
#undef PAIR
#undef DD
}

// Interior-plane second differences of the compacting kernel with the thread's z window kept in registers.  Of the
// 26 loads of quad_hessians, the rows of planes z-1 and z-2 and the centre row of plane z were all loaded one plane
// earlier (as planes z, z-1 and z+1): 10 loads and 40 shared-memory wavefronts per quad less (that kernel is bound
// by the shared-memory pipe).  Same values, same operations: bit-identical.  `have` is false on the first plane of a
// chunk and after a plane next to a z face.
struct ZWin {
    float b[8];      // row y of the next plane z   (this iteration's plane z+1 row y)
    float mm[8];     // row y of the next plane z-1 (this iteration's plane z row y)
    float mu[4], md[4];   // rows y-1, y+1 of the next plane z-1, columns x .. x+3
    float pd[4];     // row y of the next plane z-2, columns x .. x+3
};
__device__ __forceinline__ void quad_hessians_reuse(const float* ring, const ZPlanes& zp, int o, bool have, ZWin& w, float2* Hxx,
                                                    float2* Hxy, float2* Hxz, float2* Hyy, float2* Hyz, float2* Hzz)
{
    using T = HessTile;
    const float* P0 = ring + zp.P0;
    float a[8], b[8], c[8], t2[4], u2[4], pa[4], pd[4], mm[8], nn[8], mu[4], md[4], nu[4], nd[4];
    *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(P0 + o - T::PW);
    *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(P0 + o - T::PW + 4);
    *reinterpret_cast<float4*>(c) = *reinterpret_cast<const float4*>(P0 + o + T::PW);
    *reinterpret_cast<float4*>(c + 4) = *reinterpret_cast<const float4*>(P0 + o + T::PW + 4);
    *reinterpret_cast<float2*>(t2) = *reinterpret_cast<const float2*>(P0 + o - 2 * T::PW + 2);
    *reinterpret_cast<float2*>(t2 + 2) = *reinterpret_cast<const float2*>(P0 + o - 2 * T::PW + 4);
    *reinterpret_cast<float2*>(u2) = *reinterpret_cast<const float2*>(P0 + o + 2 * T::PW + 2);
    *reinterpret_cast<float2*>(u2 + 2) = *reinterpret_cast<const float2*>(P0 + o + 2 * T::PW + 4);
    *reinterpret_cast<float2*>(pa) = *reinterpret_cast<const float2*>(ring + zp.Pa + o + 2);
    *reinterpret_cast<float2*>(pa + 2) = *reinterpret_cast<const float2*>(ring + zp.Pa + o + 4);
    *reinterpret_cast<float4*>(nn) = *reinterpret_cast<const float4*>(ring + zp.Pzh + o);
    *reinterpret_cast<float4*>(nn + 4) = *reinterpret_cast<const float4*>(ring + zp.Pzh + o + 4);
    *reinterpret_cast<float2*>(nu) = *reinterpret_cast<const float2*>(ring + zp.Pzh + o - T::PW + 2);
    *reinterpret_cast<float2*>(nu + 2) = *reinterpret_cast<const float2*>(ring + zp.Pzh + o - T::PW + 4);
    *reinterpret_cast<float2*>(nd) = *reinterpret_cast<const float2*>(ring + zp.Pzh + o + T::PW + 2);
    *reinterpret_cast<float2*>(nd + 2) = *reinterpret_cast<const float2*>(ring + zp.Pzh + o + T::PW + 4);
    if (have) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { b[k] = w.b[k]; mm[k] = w.mm[k]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { mu[k] = w.mu[k]; md[k] = w.md[k]; pd[k] = w.pd[k]; }
    } else {
        *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(P0 + o);
        *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(P0 + o + 4);
        *reinterpret_cast<float2*>(pd) = *reinterpret_cast<const float2*>(ring + zp.Pd + o + 2);
        *reinterpret_cast<float2*>(pd + 2) = *reinterpret_cast<const float2*>(ring + zp.Pd + o + 4);
        *reinterpret_cast<float4*>(mm) = *reinterpret_cast<const float4*>(ring + zp.Pzl + o);
        *reinterpret_cast<float4*>(mm + 4) = *reinterpret_cast<const float4*>(ring + zp.Pzl + o + 4);
        *reinterpret_cast<float2*>(mu) = *reinterpret_cast<const float2*>(ring + zp.Pzl + o - T::PW + 2);
        *reinterpret_cast<float2*>(mu + 2) = *reinterpret_cast<const float2*>(ring + zp.Pzl + o - T::PW + 4);
        *reinterpret_cast<float2*>(md) = *reinterpret_cast<const float2*>(ring + zp.Pzl + o + T::PW + 2);
        *reinterpret_cast<float2*>(md + 2) = *reinterpret_cast<const float2*>(ring + zp.Pzl + o + T::PW + 4);
    }
    const float qs = zp.qs;
    const float2 qs2 = make_float2(qs, qs);
#define PAIR(arr, i) make_float2((arr)[(i)], (arr)[(i) + 1])
#define DD(hi, mid, lo) vmul(vsub(vsub(hi, mid), vsub(mid, lo)), qs2)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int j = 2 * g;
        const float2 f0 = PAIR(b, j + 2);
        Hxx[g] = DD(PAIR(b, j + 4), f0, PAIR(b, j));
        Hyy[g] = DD(PAIR(u2, j), f0, PAIR(t2, j));
        Hzz[g] = DD(PAIR(pa, j), f0, PAIR(pd, j));
        Hxy[g].x = __fmul_rn(__fsub_rn(__fsub_rn(c[j + 3], c[j + 1]), __fsub_rn(a[j + 3], a[j + 1])), qs);
        Hxy[g].y = __fmul_rn(__fsub_rn(__fsub_rn(c[j + 4], c[j + 2]), __fsub_rn(a[j + 4], a[j + 2])), qs);
        Hxz[g].x = __fmul_rn(__fsub_rn(__fsub_rn(nn[j + 3], nn[j + 1]), __fsub_rn(mm[j + 3], mm[j + 1])), qs);
        Hxz[g].y = __fmul_rn(__fsub_rn(__fsub_rn(nn[j + 4], nn[j + 2]), __fsub_rn(mm[j + 4], mm[j + 2])), qs);
        Hyz[g] = vmul(vsub(vsub(PAIR(nd, j), PAIR(nu, j)), vsub(PAIR(md, j), PAIR(mu, j))), qs2);
    }
#undef PAIR
#undef DD
    // the window of the next plane
#pragma unroll
    for (int k = 0; k < 8; ++k) { w.mm[k] = b[k]; w.b[k] = nn[k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) { w.pd[k] = mm[k + 2]; w.mu[k] = a[k + 2]; w.md[k] = c[k + 2]; }
}

template <int MODE, bool BRIGHT = false>
__global__ void __launch_bounds__(HessTile::NT, HESS_MIN_CTAS)
hessian_eigen_kernel(const __grid_constant__ VoxelParams p)
{
    using T = HessTile;
    extern __shared__ __align__(128) float ring[];
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    int bid = blockIdx.x;
    const int bx = bid % p.ntx; bid /= p.ntx;
    const int by = bid % p.nty;
    const int bz = bid / p.nty;
    // Tile bx covers the voxels x = 2 + TX * bx + [0, TX): the x faces 0, 1 belong to K3b anyway, and the box that
    // holds x-2 .. x+2 of them then starts at x0 = TX * bx, a multiple of 4 floats -- the copy engine rejects an
    // innermost coordinate that is not 16-byte aligned.  (y and z coordinates are free.)
    const int x0 = bx * T::TX, y0 = by * T::TY - 2;           // global coordinates of tile entry (0, 0)
    const int w = p.f.w, h = p.f.h, l = p.f.l;
    // centre planes of this CTA (z faces included: see ZPlanes; x and y faces belong to K3b)
    const int zs = p.z_begin + bz * p.zchunk;
    const int ze = min(zs + p.zchunk, p.z_begin + p.nz);
    if (zs >= ze) return;

    // plane tiles arrive by TMA (one instruction of one thread per plane), see TileRing
    TileRing<T> tr;
    tr.init(ring, reinterpret_cast<uint64_t*>(ring + T::SLOTS * T::SLOT), tid);
#if K3A_STREAM
    // Slot reuse without a CTA-wide barrier: an "empty" mbarrier per slot (NT arrivals: every thread releases plane z-2
    // after its iteration z) next to the "full" one; planes are issued in order by whichever warp first finds the next
    // one due and its slot released (compare-and-swap counter), as in the streaming kernel below.  Plane sequence
    // number n <-> plane zs - 2 + n, slot n % SLOTS, phase n / SLOTS.  Warps may drift SLOTS - 5 planes apart, which
    // takes them out of lock step: they no longer hit the FP32 and the ALU pipe all at the same time.
    const uint32_t empty_s = tr.mbar_s + 8 * T::SLOTS;
    int* s_next = reinterpret_cast<int*>(ring + T::SLOTS * T::SLOT) + 4 * T::SLOTS;
    const int nseq = (ze - zs) + 4;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < T::SLOTS; ++k) mbar_init(empty_s + 8 * k, T::NT);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0) {
        const int first = min(T::SLOTS, nseq);
        for (int n = 0; n < first; ++n) tr.issue(&p.tmap, x0, y0, zs - 2 + n, p.f.base, l);
        *s_next = first;
    }
    __syncthreads();
    auto issue_seq = [&](int n) {                 // one thread: plane sequence number n into slot n % SLOTS
        const int slot = n % T::SLOTS, plane = zs - 2 + n;
        const uint32_t bar = tr.mbar_s + 8 * slot;
        if (plane < 0 || plane > l - 1) mbar_arrive(bar);
        else {
            mbar_expect_tx(bar, T::PLANE * 4);
            tma_load_3d(tr.ring_s + slot * (T::SLOT * 4), &p.tmap, bar, x0, y0, plane - p.f.base);
        }
    };
#else
    __syncthreads();
    // prologue: planes zs-2 .. zs+2 on their way into the ring; the first four must have landed
    if (tid == 0)
        for (int q = zs - 2; q <= zs + 2; ++q) tr.issue(&p.tmap, x0, y0, q, p.f.base, l);
#endif
#pragma unroll 1
    for (int k = 0; k < 4; ++k) tr.wait_next();

    const int xq = bx * T::TX + 2 + 4 * tx;      // first of this thread's 4 x voxels (xq = 2 mod 4)
    bool m[4];                                   // voxel is interior in x
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = xq + j <= w - 3;
    const bool any_x = m[0] || m[1] || m[2] || m[3];
    const bool all_x = m[0] && m[1] && m[2] && m[3] && p.vec_ok;
    float vmin = 3.4e38f, vmax = 0.0f;

    const long long plane_vox = (long long)h * w;
    long long iz = (long long)(zs - p.z_begin) * plane_vox + xq;   // index of (xq, 0, z) in the own-plane outputs
    for (int z = zs; z < ze; ++z, iz += plane_vox) {
#if K3A_STREAM
        if (tx == 0) {                            // issue duty (lane 0 of every warp): planes up to z + SLOTS - 3
            const int c = z - zs;
            for (;;) {
                const int n = *reinterpret_cast<volatile int*>(s_next);
                if (n >= nseq || n > c + T::SLOTS - 1) break;
                if (!mbar_test(empty_s + 8 * (n % T::SLOTS), (unsigned)(n / T::SLOTS - 1) & 1u)) break;   // its slot is still in use
                if (atomicCAS(s_next, n, n + 1) == n) issue_seq(n);
            }
        }
        __syncwarp();
#else
        __syncthreads();                          // everyone is done with plane z-3's slot
        if (tid == 0 && z + 1 < ze) tr.issue(&p.tmap, x0, y0, z + 3, p.f.base, l);   // lands there while plane z is processed
#endif
        tr.wait_next();                           // plane z+2 has landed

        const bool z_general = z < 2 || z > l - 3;

#pragma unroll 1
        for (int half = 0; half < T::RPT; ++half) {
            const int yl = ty + (T::TY / T::RPT) * half;   // row inside the tile
            const int y = by * T::TY + yl;
            if (y < 2 || y > h - 3 || !any_x) continue;
            // second differences of the quad as two packed pairs (voxels 0,1 and 2,3)
            float2 Hxx[2], Hxy[2], Hxz[2], Hyy[2], Hyz[2], Hzz[2];
            if (z_general) quad_hessians<true>(ring, z_planes(tr.o, z, l, p.k.sigma2), (yl + 2) * T::PW + 4 * tx, Hxx, Hxy, Hxz, Hyy, Hyz, Hzz);
            else quad_hessians<false>(ring, z_planes_interior(tr.o, p.k.sigma2), (yl + 2) * T::PW + 4 * tx, Hxx, Hxy, Hxz, Hyy, Hyz, Hzz);

            const long long i0 = iz + (long long)y * w;
            if (MODE == 2 || !all_x) {            // stage dump; quads that straddle an x face; unaligned widths
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (m[j]) {
                        typedef Lanes<float2> L2;
                        Hess H;
                        H.xx = L2::get(Hxx[j >> 1], j & 1); H.xy = L2::get(Hxy[j >> 1], j & 1);
                        H.xz = L2::get(Hxz[j >> 1], j & 1); H.yy = L2::get(Hyy[j >> 1], j & 1);
                        H.yz = L2::get(Hyz[j >> 1], j & 1); H.zz = L2::get(Hzz[j >> 1], j & 1);
                        if (MODE == 2) {
                            p.D[0][i0 + j] = H.zz; p.D[1][i0 + j] = H.yy; p.D[2][i0 + j] = H.yz;
                            p.D[3][i0 + j] = H.xx; p.D[4][i0 + j] = H.xy; p.D[5][i0 + j] = H.xz;
                        } else {
                            const float jv = voxel_update<MODE == 2 ? 0 : MODE>(p, i0 + j, H);
                            vmin = fminf(vmin, jv); vmax = fmaxf(vmax, jv);
                        }
                    }
                continue;
            }
            float jold[4] = { 0.f, 0.f, 0.f, 0.f };
            if (MODE == 1) {
                *reinterpret_cast<float2*>(jold) = *reinterpret_cast<const float2*>(p.J + i0);
                *reinterpret_cast<float2*>(jold + 2) = *reinterpret_cast<const float2*>(p.J + i0 + 2);
            }
            float jn[4];
            uint32_t cx = 0, cy = 0, cz = 0;      // four direction codes each, byte j = voxel j
            float ex[4], ey[4], ez[4];
            bool wr[4];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int j = 2 * g;
                Eig3x2 e;
                eig_sym3<float2, BRIGHT>(Hxx[g], Hxy[g], Hxz[g], Hyy[g], Hyz[g], Hzz[g], e);
                const float2 v = vesselness<float2, BRIGHT>(e, p.k);
                wr[j] = MODE == 0 || v.x > jold[j];
                wr[j + 1] = MODE == 0 || v.y > jold[j + 1];
                jn[j] = wr[j] ? v.x : jold[j];
                jn[j + 1] = wr[j + 1] ? v.y : jold[j + 1];
                cx |= (dir_code(e.vx.x) << (8 * j)) | (dir_code(e.vx.y) << (8 * j + 8));
                cy |= (dir_code(e.vy.x) << (8 * j)) | (dir_code(e.vy.y) << (8 * j + 8));
                cz |= (dir_code(e.vz.x) << (8 * j)) | (dir_code(e.vz.y) << (8 * j + 8));
                ex[j] = e.vx.x; ey[j] = e.vy.x; ez[j] = e.vz.x;
                ex[j + 1] = e.vx.y; ey[j + 1] = e.vy.y; ez[j + 1] = e.vz.y;
                vmin = fminf(vmin, fminf(jn[j], jn[j + 1])); vmax = fmaxf(vmax, fmaxf(jn[j], jn[j + 1]));
            }
            if (MODE == 0) {
                st_pair(p.J + i0, jn);
                st_codes(p.Vx + i0, cx); st_codes(p.Vy + i0, cy); st_codes(p.Vz + i0, cz);
                if (p.scale_idx) st_codes(p.scale_idx + i0, 0u);
                if (p.dir) { st_pair(p.dir + i0, ex); st_pair(p.dir + p.voxels + i0, ey); st_pair(p.dir + 2 * p.voxels + i0, ez); }
            } else if (wr[0] || wr[1] || wr[2] || wr[3]) {
                st_pair(p.J + i0, jn);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (wr[j]) {
                        p.Vx[i0 + j] = (uint8_t)(cx >> (8 * j));
                        p.Vy[i0 + j] = (uint8_t)(cy >> (8 * j));
                        p.Vz[i0 + j] = (uint8_t)(cz >> (8 * j));
                        if (p.scale_idx) p.scale_idx[i0 + j] = (uint8_t)p.scale;
                        if (p.dir) {
                            p.dir[i0 + j] = ex[j];
                            p.dir[p.voxels + i0 + j] = ey[j];
                            p.dir[2 * p.voxels + i0 + j] = ez[j];
                        }
                    }
            }
        }
        tr.rotate();
#if K3A_STREAM
        // consumer release of plane z-2 (sequence number z - zs): every value this thread loaded from the slot has been
        // consumed by arithmetic that precedes the arrive in program order
        mbar_arrive(empty_s + 8 * ((z - zs) % T::SLOTS));
#endif
    }
    if (MODE == 2) return;
    // warp-shuffle reductions of min (first scale) and max (last scale), one atomic per warp
    if (MODE == 0) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, s));
        if (tx == 0 && vmin < 3.0e38f) atomicMin(p.minmax + 0, __float_as_int(vmin));
    }
    {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
        if (tx == 0) atomicMax(p.minmax + 1, __float_as_int(vmax));
    }
}

// K3a', later scales of a bright-ridge run (the common case: blackwhite == false,
// frangi.cpp:254-271): the same z-marching tile, but only voxels that can still win go through
// the eigen stage.  A voxel is overwritten only when its response is positive, which needs the
// two eigenvalues of largest magnitude to be <= 0; by Ky Fan's inequality the sum of any two
// diagonal entries is then <= e1 + e2 <= 0, so  Dxx+Dyy > 0 | Dxx+Dzz > 0 | Dyy+Dzz > 0  proves
// the response is 0 (three adds, before any eigen work; ~80 % of the voxels of a neuron volume).
// Phase A (all threads): second differences of the thread's quad, the test, survivors appended
// (warp-aggregated atomics) to a shared-memory queue of {6 second differences, packed voxel
// coordinate}.  Phase B: whenever the queue holds a full batch (2 entries per thread) every
// thread takes two entries through the packed eigen stage and updates J / V where the response
// beats the stored one, so the expensive stage always runs with full lanes.  The queue persists
// across planes and is flushed at the end of the chunk.
#ifndef K3C_SLOTS
#define K3C_SLOTS 6
#endif
#ifndef K3C_DRAIN_UNROLL
#define K3C_DRAIN_UNROLL 1
#endif
#ifndef K3C_V1
// ---- K3a' (current form): warp-autonomous streaming, mbarrier full / empty tile ring, per-warp survivor queues ----
// History (profiles/r2*): the first form (below, -DK3C_V1) read every quad's 26 row segments of the five planes from
// shared memory per plane and ran phase A / phase B in CTA-wide lock step (two __syncthreads per plane, one shared
// survivor queue): 13.4 ms per launch at 2048 x 2048 x 512, 48 % issue utilisation.  Testing the diagonal first
// and fetching the rest only for quads with a survivor did not help (survivors are noise voxels, spread evenly: 57 %
// of the quads hold one); halving the shared-memory wavefronts (aligned tiles + shuffles) did not help either: the
// kernel is bound by instruction count and by the lock step (stall reasons: fixed-latency wait, barrier), not by a
// single pipe.  This form attacks both:
//  * a WARP owns one tile row and marches along z on its own: the tile ring is filled by TMA and guarded by a
//    full (transaction) and an empty (consumer release, 256 arrivals) mbarrier per slot -- no __syncthreads in
//    the plane loop; slot s is refilled by lane 0 of warp s once every warp has released it, so warps may drift a
//    few planes apart and the eigen stage of one warp overlaps the second differences of another;
//  * each warp has its own survivor queue (no shared counter, no atomics): slots come from ballots, a drain runs
//    whenever the warp has 64 survivors (one packed pair per lane);
//  * tiles are 16-byte aligned with the quads: a row segment of a quad is ONE 128-bit load with a lane stride of
//    16 bytes, the columns next to a quad come from the neighbour lanes by shuffle; the outer half of the quads
//    of lanes 0 / 31 is the halo (a tile is 124 columns wide: box columns 2 .. 125), which a per-voxel mask covers,
//    so there is no edge-lane special case at all;
//  * a lane keeps its own row of the planes z-2 .. z+2 and the x first differences of z-1 .. z+1 in register
//    rings rotated by renaming (the plane loop dispatches on z mod 5 to five instantiations of phase A);
//  * survivors carry their 32-bit linear voxel offset; the stored response is gathered at the drain.
// Same operations on the same operands in the same order as quad_hessians<false>: bit-identical second differences.
struct HessTileC {                               // 124 x 8 voxels: one row per warp, one quad per lane
#ifndef K3C_TX
#define K3C_TX 124                               // 120: lanes 0 / 31 carry only halo (the first round-2 form), for A/B runs
#endif
    static constexpr int TX = K3C_TX, TY = 8, NT = 256;
    static_assert(TX == 124 || TX == 120, "tile width");
    static constexpr int XH = (128 - TX) / 2;    // halo columns each side: half a quad (of lanes 0 and 31), or a whole one
    static constexpr int PW = TX + 2 * XH;       // 128 floats per tile row
    static constexpr int PH = TY + 4;
    static constexpr int PLANE = PW * PH;        // 1536 floats: the TMA box
    static constexpr int SLOT = PLANE;           // 6144 bytes (a multiple of 128)
    static constexpr int SLOTS = 8;              // = warps: warp s refills slot s
    static constexpr int RING_BYTES = SLOTS * SLOT * 4;
    static constexpr int X_FIRST = XH == 2 ? 2 : 0;   // tile bx covers the voxels x = X_FIRST + TX * bx + [0, TX): TX 124 starts at the
                                                 // first interior column 2; TX 120 at 0 with x = 0, 1 masked (shell)
#ifndef K3C_AHEAD
#define K3C_AHEAD 6
#endif
#ifndef K3C_ROT5
#define K3C_ROT5 0                               // 1: the register rings rotate by renaming (five instantiations of phase A: no moves, but
                                                 // 5x the code -- measured 0.4 ms per step SLOWER, instruction-cache misses); 0: by 24 moves
#endif
    static constexpr int AHEAD = K3C_AHEAD;      // plane sequence number n is issued at iteration >= n - AHEAD (n = z + 2 at AHEAD 4)
    static_assert(AHEAD >= 4 && AHEAD <= 7, "a warp has always released sequence number c-1 at iteration c, never c: AHEAD 8 would wait for itself");
};
struct HessQueue {                               // per warp
    static constexpr int BATCH = 64;             // one packed pair per lane
    static constexpr int CAP = 256;              // power of two >= BATCH - 1 + the 124 entries one plane can add
// Issue duty: a warp looks for planes to issue at the top of every (K3C_DUTY + 1)-th iteration.  The SLOWEST warp is the
// one whose release frees a slot, and it may be the only warp not blocked on a full barrier, so its own checks must
// keep it supplied: a check at iteration c issues up to sequence number c + AHEAD, iteration c + k needs c + k + 4,
// hence the period must not exceed AHEAD - 3 (a period of 4 deadlocked on the GPU: every warp waiting for a plane
// nobody was going to issue).
#ifndef K3C_DUTY
#define K3C_DUTY 1
#endif
static_assert(K3C_DUTY == 0 || K3C_DUTY == 1, "issue duty period must be 1 or 2");
static_assert(K3C_DUTY + 1 <= K3C_AHEAD - 3, "issue duty period exceeds the look-ahead");
#ifndef K3C_AOS
#define K3C_AOS 0                                // 1: 32-byte entries written with two 128-bit stores (measured 0.6 ms per launch SLOWER); 0: structure of arrays, seven 32-bit stores
#endif
    // K3C_AOS: entry e = 8 floats at 8 * e: Dxx, Dxy, Dxz, Dyy | Dyz, Dzz, linear voxel offset, unused.  An append is two
    // predicated 128-bit stores (8 store instructions per quad instead of 28); a drain gives lane t the entries head + t and
    // head + 32 + t (a lane stride of 32 bytes: 2-way bank conflicts on the 0.4 drains per plane instead of 4-way).
    // else: field f of entry e at float f * CAP + e; fields = Dxx, Dxy, Dxz, Dyy, Dyz, Dzz, linear voxel offset
    static constexpr int FIELDS = K3C_AOS ? 8 : 7;
    static constexpr int WARP_FLOATS = FIELDS * CAP;
    static constexpr int BYTES = (HessTileC::NT / 32) * WARP_FLOATS * 4;
    static constexpr int SMEM_BYTES = HessTileC::RING_BYTES + BYTES + 2 * HessTileC::SLOTS * 8 + 16;   // + full and empty mbarriers, issue counter
};

template <int N> struct IntC { static constexpr int value = N; };

// the five resident planes of a z-marching tile as a field for hessian_at_face (planes next to a z face)
struct RingField {
    const float* ring;
    int o0, o1, o2, o3, o4;     // ring offsets of planes zc-2 .. zc+2
    int zc, xb, yb;             // centre plane; global coordinates of tile entry (0, 0)
    int w, h, l;
    __device__ __forceinline__ float at(int x, int y, int z) const
    {
        const int k = z - zc;
        const int o = k <= -2 ? o0 : (k == -1 ? o1 : (k == 0 ? o2 : (k == 1 ? o3 : o4)));
        return ring[o + (y - yb) * HessTileC::PW + (x - xb)];
    }
};

// seven predicated 32-bit shared stores of one queue entry (no branch)
__device__ __forceinline__ void queue_append7(uint32_t addr, bool pred, float d0, float d1, float d2, float d3, float d4,
                                              float d5, uint32_t off)
{
    constexpr int S = HessQueue::CAP * 4;          // byte stride between the field arrays
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "@p st.shared.f32 [%0], %2;\n\t"
        "@p st.shared.f32 [%0 + %9], %3;\n\t"
        "@p st.shared.f32 [%0 + %10], %4;\n\t"
        "@p st.shared.f32 [%0 + %11], %5;\n\t"
        "@p st.shared.f32 [%0 + %12], %6;\n\t"
        "@p st.shared.f32 [%0 + %13], %7;\n\t"
        "@p st.shared.b32 [%0 + %14], %8;\n\t}"
        ::"r"(addr), "r"((uint32_t)pred), "f"(d0), "f"(d1), "f"(d2), "f"(d3), "f"(d4), "f"(d5), "r"(off),
          "n"(S), "n"(2 * S), "n"(3 * S), "n"(4 * S), "n"(5 * S), "n"(6 * S) : "memory");
}

// two predicated 128-bit shared stores of one 32-byte queue entry (no branch)
__device__ __forceinline__ void queue_append_aos(uint32_t addr, bool pred, float d0, float d1, float d2, float d3, float d4,
                                                 float d5, uint32_t off)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "@p st.shared.v4.f32 [%0], {%2, %3, %4, %5};\n\t"
        "@p st.shared.v4.b32 [%0 + 16], {%6, %7, %8, %8};\n\t}"
        ::"r"(addr), "r"((uint32_t)pred), "f"(d0), "f"(d1), "f"(d2), "f"(d3), "r"(__float_as_uint(d4)), "r"(__float_as_uint(d5)),
          "r"(off) : "memory");
}

__global__ void __launch_bounds__(HessTileC::NT, 2)
hessian_eigen_compact_kernel(const __grid_constant__ VoxelParams p)
{
    using T = HessTileC;
    using Q = HessQueue;
    extern __shared__ __align__(128) float ring[];
    const int tid = threadIdx.x;
    const int tx = tid & 31, wid = tid >> 5;
    float* qw = ring + T::SLOTS * T::SLOT + wid * Q::WARP_FLOATS;          // this warp's queue
    const uint32_t full_s = smem_u32(ring + T::SLOTS * T::SLOT + (T::NT / 32) * Q::WARP_FLOATS);
    const uint32_t empty_s = full_s + 8 * T::SLOTS;
    int* s_next = reinterpret_cast<int*>(ring + T::SLOTS * T::SLOT + (T::NT / 32) * Q::WARP_FLOATS) + 4 * T::SLOTS;   // after the barriers
    const uint32_t ring_s = smem_u32(ring);
    int bid = blockIdx.x;
    const int bx = bid % p.ntx; bid /= p.ntx;
    const int by = bid % p.nty;
    const int bz = bid / p.nty;
    const int x0 = bx * T::TX + T::X_FIRST - T::XH, y0 = by * T::TY - 2;   // global coordinates of tile entry (0, 0); x0 is a multiple of 4 floats (TMA)
    static_assert((T::X_FIRST - T::XH) % 4 == 0 && T::TX % 4 == 0, "TMA: the innermost box coordinate must be a multiple of 16 bytes");
    const int w = p.f.w, h = p.f.h, l = p.f.l;
    const int zs = p.z_begin + bz * p.zchunk;
    const int ze = min(zs + p.zchunk, p.z_begin + p.nz);
    if (zs >= ze) return;
    const int nz = ze - zs;
    const int nseq = nz + 4;                      // plane sequence numbers: n <-> plane zs - 2 + n, slot n % 8, phase n / 8

    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < T::SLOTS; ++k) { mbar_init(full_s + 8 * k, 1); mbar_init(empty_s + 8 * k, T::NT); }
        mbar_init_fence();
    }
    __syncthreads();
    // start the copy of plane sequence number n into slot n % 8 (one lane)
    auto issue = [&](int n) {
        const int plane = zs - 2 + n;
        const uint32_t bar = full_s + 8 * (n & 7);
        if (plane < 0 || plane > l - 1) mbar_arrive(bar);          // beyond a z face: never read, but the phase completes
        else {
            mbar_expect_tx(bar, T::PLANE * 4);
            tma_load_3d(ring_s + (n & 7) * (T::SLOT * 4), &p.tmap, bar, x0, y0, plane - p.f.base);
        }
    };
    // Planes are issued in order by WHICHEVER warp first finds the next one due (within AHEAD of its own window) and
    // its slot released by every warp: s_next is the next sequence number to issue, claimed with a compare-and-swap.
    // No warp owns a slot, so a warp that is busy in its eigen stage never holds up the others' data.
    if (tid == 0) {
        const int first = min(T::SLOTS, nseq);
        for (int n = 0; n < first; ++n) issue(n);
        *s_next = first;
    }
    __syncthreads();
    const int xq = x0 + 4 * tx;                   // first voxel of the lane's quad (lanes 0 and 31: half or all of it is halo)
    const int yrow = by * T::TY + wid;            // the warp's row
    const bool row_ok = yrow >= 2 && yrow <= h - 3;
    bool m[4];                                    // voxel j of the quad is an interior voxel of this tile
#pragma unroll
    for (int j = 0; j < 4; ++j)
        m[j] = row_ok && 4 * tx + j >= T::XH && 4 * tx + j < T::XH + T::TX && xq + j >= 2 && xq + j <= w - 3;
    const int o_own = (wid + 2) * T::PW + 4 * tx; // tile entry of (xq, yrow)
    const float qs = 0.25f * p.k.sigma2;
    const float2 qs2 = make_float2(qs, qs);
    const unsigned below = (1u << tx) - 1u;
    const unsigned plane_vox = (unsigned)(h * w);                          // own voxels < 2^32 (checked by the host)
    unsigned off0 = (unsigned)(((long long)(zs - p.z_begin) * h + yrow) * w + xq);   // linear offset of (xq, yrow, z) in the own-plane outputs
    const uint32_t qw_s = smem_u32(qw);
    unsigned head = 0, tail = 0;                  // this warp's queue: entries [head, tail) pending (warp-uniform)
    float vmax = 0.0f;

    // Register rings (slot (rot + k) % 5 holds plane z-2+k): C = the quad's own row, G = its x first differences
    // F[x+1] - F[x-1] (planes z-1 .. z+1 live).  E0 = the columns x-2, x-1, x+4, x+5 of plane z.
    float C[5][4], G[5][4], E0[4];
    bool have = false;        // the rings hold planes z-2 .. z+1 (and G z-1, z; E0) of the plane about to be processed

    auto edges4 = [&](const float* c, float* e) {          // columns x-2, x-1, x+4, x+5 from the neighbour lanes
        e[0] = __shfl_up_sync(0xffffffffu, c[2], 1); e[1] = __shfl_up_sync(0xffffffffu, c[3], 1);
        e[2] = __shfl_down_sync(0xffffffffu, c[0], 1); e[3] = __shfl_down_sync(0xffffffffu, c[1], 1);
    };
    auto xdiff = [&](const float* c, float lo, float hi, float* g) {      // F[x+j+1] - F[x+j-1]
        g[0] = __fsub_rn(c[1], lo); g[1] = __fsub_rn(c[2], c[0]); g[2] = __fsub_rn(c[3], c[1]); g[3] = __fsub_rn(hi, c[2]);
    };
    auto xdiff_row = [&](const float* c, float* g) {        // the same for a row whose edge columns are not kept
        xdiff(c, __shfl_up_sync(0xffffffffu, c[3], 1), __shfl_down_sync(0xffffffffu, c[0], 1), g);
    };
    auto ld4a = [&](float* d, const float* s) { *reinterpret_cast<float4*>(d) = *reinterpret_cast<const float4*>(s); };

    // the diagonal-sum test (Ky Fan), queue slots from ballots, survivors appended
    auto test_append = [&](const float2* Hxx, const float2* Hxy, const float2* Hxz, const float2* Hyy, const float2* Hyz,
                           const float2* Hzz) {
        bool surv[4];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const float2 sxy = vadd(Hxx[g], Hyy[g]), sxz = vadd(Hxx[g], Hzz[g]), syz = vadd(Hyy[g], Hzz[g]);
            surv[2 * g] = m[2 * g] && fmaxf(fmaxf(sxy.x, sxz.x), syz.x) <= 0.0f;
            surv[2 * g + 1] = m[2 * g + 1] && fmaxf(fmaxf(sxy.y, sxz.y), syz.y) <= 0.0f;
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, surv[0]), b1 = __ballot_sync(0xffffffffu, surv[1]);
        const unsigned b2 = __ballot_sync(0xffffffffu, surv[2]), b3 = __ballot_sync(0xffffffffu, surv[3]);
        const unsigned s0 = tail + __popc(b0 & below);
        const unsigned t1 = tail + __popc(b0);
        const unsigned s1 = t1 + __popc(b1 & below);
        const unsigned t2 = t1 + __popc(b1);
        const unsigned s2 = t2 + __popc(b2 & below);
        const unsigned t3 = t2 + __popc(b2);
        const unsigned s3 = t3 + __popc(b3 & below);
        tail = t3 + __popc(b3);
        const unsigned slot[4] = { s0, s1, s2, s3 };
        typedef Lanes<float2> L2;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#if K3C_AOS
            queue_append_aos(qw_s + 32 * (slot[j] & (Q::CAP - 1)), surv[j],
                             L2::get(Hxx[j >> 1], j & 1), L2::get(Hxy[j >> 1], j & 1), L2::get(Hxz[j >> 1], j & 1),
                             L2::get(Hyy[j >> 1], j & 1), L2::get(Hyz[j >> 1], j & 1), L2::get(Hzz[j >> 1], j & 1), off0 + j);
#else
            queue_append7(qw_s + 4 * (slot[j] & (Q::CAP - 1)), surv[j],
                          L2::get(Hxx[j >> 1], j & 1), L2::get(Hxy[j >> 1], j & 1), L2::get(Hxz[j >> 1], j & 1),
                          L2::get(Hyy[j >> 1], j & 1), L2::get(Hyz[j >> 1], j & 1), L2::get(Hzz[j >> 1], j & 1), off0 + j);
#endif
    };
#define PAIR(arr, i) make_float2((arr)[(i)], (arr)[(i) + 1])
#define DD(hi, mid, lo) vmul(vsub(vsub(hi, mid), vsub(mid, lo)), qs2)
    // phase A of an interior plane (2 <= z <= l-3) with the register rings at rotation R; Pk = ring slot of plane z+k
    auto phase_a = [&](auto rc, const float* Pm2, const float* Pm1, const float* P0, const float* Pp1, const float* Pp2) {
        constexpr int R = decltype(rc)::value;
        float (&Cm2)[4] = C[R % 5];
        float (&Cm1)[4] = C[(R + 1) % 5];
        float (&C0)[4] = C[(R + 2) % 5];
        float (&Cp1)[4] = C[(R + 3) % 5];
        float (&Cp2)[4] = C[(R + 4) % 5];
        float (&Gm1)[4] = G[(R + 1) % 5];
        float (&G0)[4] = G[(R + 2) % 5];
        float (&Gp1)[4] = G[(R + 3) % 5];
        if (!have) {          // first plane of the chunk, or the plane after one next to a z face (warp-uniform)
            ld4a(Cm2, Pm2 + o_own); ld4a(Cm1, Pm1 + o_own); ld4a(C0, P0 + o_own); ld4a(Cp1, Pp1 + o_own);
            xdiff_row(Cm1, Gm1);
            edges4(C0, E0);
            xdiff(C0, E0[1], E0[2], G0);
        }
        // every lane loads and shuffles (rows beyond the volume are zero-filled tile rows)
        ld4a(Cp2, Pp2 + o_own);
        float E1[4];
        edges4(Cp1, E1);
        xdiff(Cp1, E1[1], E1[2], Gp1);
        float a[4], c[4], mu[4], md[4], nu[4], nd[4], t2[4], u2[4];
        ld4a(a, P0 + o_own - T::PW); ld4a(c, P0 + o_own + T::PW);
        ld4a(t2, P0 + o_own - 2 * T::PW); ld4a(u2, P0 + o_own + 2 * T::PW);
        ld4a(mu, Pm1 + o_own - T::PW); ld4a(md, Pm1 + o_own + T::PW);
        ld4a(nu, Pp1 + o_own - T::PW); ld4a(nd, Pp1 + o_own + T::PW);
        float ga[4], gc[4];
        xdiff_row(a, ga);
        xdiff_row(c, gc);
        float2 Hxx[2], Hxy[2], Hxz[2], Hyy[2], Hyz[2], Hzz[2];
        {
            const float2 hi0 = make_float2(C0[2], C0[3]), hi1 = make_float2(E0[2], E0[3]);
            const float2 lo0 = make_float2(E0[0], E0[1]), lo1 = make_float2(C0[0], C0[1]);
            Hxx[0] = DD(hi0, lo1, lo0);          // voxels 0, 1: centre = (C0[0], C0[1])
            Hxx[1] = DD(hi1, hi0, lo1);          // voxels 2, 3: centre = (C0[2], C0[3])
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int j = 2 * g;
            const float2 f0 = PAIR(C0, j);
            Hyy[g] = DD(PAIR(u2, j), f0, PAIR(t2, j));
            Hzz[g] = DD(PAIR(Cp2, j), f0, PAIR(Cm2, j));
            Hxy[g] = vmul(vsub(PAIR(gc, j), PAIR(ga, j)), qs2);
            Hxz[g] = vmul(vsub(PAIR(Gp1, j), PAIR(Gm1, j)), qs2);
            Hyz[g] = vmul(vsub(vsub(PAIR(nd, j), PAIR(nu, j)), vsub(PAIR(md, j), PAIR(mu, j))), qs2);
        }
        test_append(Hxx, Hxy, Hxz, Hyy, Hyz, Hzz);
#pragma unroll
        for (int k = 0; k < 4; ++k) E0[k] = E1[k];
        have = true;
    };
#undef PAIR
#undef DD
    // phase A of a plane next to a z face (z < 2 or z > l-3, at most four planes of a volume): the in-plane terms as in
    // phase_a, the z terms on the clamped planes with the face scales (ZPlanes; the arithmetic of quad_hessians<true>)
#ifndef K3C_ZFACE_FAST
#define K3C_ZFACE_FAST 1      // 0: hessian_at_face per voxel (7.5 x the time of an interior plane), for A/B runs
#endif
#if K3C_ZFACE_FAST
    auto phase_a_general = [&](int z, int o0, int o1, int o2, int o3, int o4) {
        const int o[5] = { o0, o1, o2, o3, o4 };
        const ZPlanes zp = z_planes(o, z, l, p.k.sigma2);
        const float* P0 = ring + zp.P0;
        float c0[4], E[4], a[4], c[4], t2[4], u2[4], ga[4], gc[4];
        ld4a(c0, P0 + o_own);
        edges4(c0, E);
        ld4a(a, P0 + o_own - T::PW); ld4a(c, P0 + o_own + T::PW);
        ld4a(t2, P0 + o_own - 2 * T::PW); ld4a(u2, P0 + o_own + 2 * T::PW);
        xdiff_row(a, ga);
        xdiff_row(c, gc);
        float pa[4], pb[4], pc[4], pd[4], mm[4], nn[4], gm[4], gn[4], mu[4], md[4], nu[4], nd[4];
        ld4a(pa, ring + zp.Pa + o_own); ld4a(pb, ring + zp.Pb + o_own);
        ld4a(pc, ring + zp.Pc + o_own); ld4a(pd, ring + zp.Pd + o_own);
        ld4a(mm, ring + zp.Pzl + o_own); ld4a(nn, ring + zp.Pzh + o_own);
        xdiff_row(mm, gm);
        xdiff_row(nn, gn);
        ld4a(mu, ring + zp.Pzl + o_own - T::PW); ld4a(md, ring + zp.Pzl + o_own + T::PW);
        ld4a(nu, ring + zp.Pzh + o_own - T::PW); ld4a(nd, ring + zp.Pzh + o_own + T::PW);
        const float2 qz2 = make_float2(zp.qz, zp.qz), qzz2 = make_float2(zp.qzz, zp.qzz);
        const float2 sh2 = make_float2(zp.sh, zp.sh), sl2 = make_float2(zp.sl, zp.sl);
        float2 Hxx[2], Hxy[2], Hxz[2], Hyy[2], Hyz[2], Hzz[2];
#define PAIR(arr, i) make_float2((arr)[(i)], (arr)[(i) + 1])
#define DD(hi, mid, lo) vmul(vsub(vsub(hi, mid), vsub(mid, lo)), qs2)
        {
            const float2 hi0 = make_float2(c0[2], c0[3]), hi1 = make_float2(E[2], E[3]);
            const float2 lo0 = make_float2(E[0], E[1]), lo1 = make_float2(c0[0], c0[1]);
            Hxx[0] = DD(hi0, lo1, lo0);
            Hxx[1] = DD(hi1, hi0, lo1);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int j = 2 * g;
            Hyy[g] = DD(PAIR(u2, j), PAIR(c0, j), PAIR(t2, j));
            Hzz[g] = vmul(vsub(vmul(sh2, vsub(PAIR(pa, j), PAIR(pb, j))), vmul(sl2, vsub(PAIR(pc, j), PAIR(pd, j)))), qzz2);
            Hxy[g] = vmul(vsub(PAIR(gc, j), PAIR(ga, j)), qs2);
            Hxz[g] = vmul(vsub(PAIR(gn, j), PAIR(gm, j)), qz2);
            Hyz[g] = vmul(vsub(vsub(PAIR(nd, j), PAIR(nu, j)), vsub(PAIR(md, j), PAIR(mu, j))), qz2);
        }
#undef PAIR
#undef DD
        test_append(Hxx, Hxy, Hxz, Hyy, Hyz, Hzz);
        have = false;
    };
#else
    auto phase_a_general = [&](int z, int o0, int o1, int o2, int o3, int o4) {
        RingField f;
        f.ring = ring; f.o0 = o0; f.o1 = o1; f.o2 = o2; f.o3 = o3; f.o4 = o4;
        f.zc = z; f.xb = x0; f.yb = y0; f.w = w; f.h = h; f.l = l;
        float2 Hxx[2], Hxy[2], Hxz[2], Hyy[2], Hyz[2], Hzz[2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            typedef Lanes<float2> L2;
            Hess H = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
            if (m[j]) H = hessian_at_face(f, xq + j, yrow, z, p.k.sigma2);
            L2::set(Hxx[j >> 1], j & 1, H.xx); L2::set(Hxy[j >> 1], j & 1, H.xy); L2::set(Hxz[j >> 1], j & 1, H.xz);
            L2::set(Hyy[j >> 1], j & 1, H.yy); L2::set(Hyz[j >> 1], j & 1, H.yz); L2::set(Hzz[j >> 1], j & 1, H.zz);
        }
        test_append(Hxx, Hxy, Hxz, Hyy, Hyz, Hzz);
        have = false;
    };
#endif

    // Phase B on this warp's entries [head, head + count), count <= 64
    auto drain = [&](int count) {
#if K3C_AOS
        // lane t takes entries head + t and head + 32 + t
        const bool has0 = tx < count, two = 32 + tx < count;
        if (has0) {
            const float4* q0 = reinterpret_cast<const float4*>(qw + 8 * ((head + tx) & (Q::CAP - 1)));
            const float4* q1 = reinterpret_cast<const float4*>(qw + 8 * ((head + 32 + tx) & (Q::CAP - 1)));
            const float4 a0 = q0[0], b0 = q0[1];
            const float4 a1 = two ? q1[0] : a0, b1 = two ? q1[1] : b0;
            const float2 f[6] = { make_float2(a0.x, a1.x), make_float2(a0.y, a1.y), make_float2(a0.z, a1.z),
                                  make_float2(a0.w, a1.w), make_float2(b0.x, b1.x), make_float2(b0.y, b1.y) };
            const unsigned i0 = __float_as_uint(b0.z), i1 = __float_as_uint(b1.z);
#else
        // lane t takes entries head + 2t, head + 2t + 1
        const int o = 2 * tx;
        if (o < count) {
            const bool two = o + 1 < count;
            const float* e0 = qw + ((head + o) & (Q::CAP - 1));                 // head is even: the pair does not wrap
            float2 f[Q::FIELDS];
#pragma unroll
            for (int k = 0; k < Q::FIELDS; ++k) f[k] = *reinterpret_cast<const float2*>(e0 + k * Q::CAP);
            const unsigned i0 = __float_as_uint(f[6].x), i1 = two ? __float_as_uint(f[6].y) : i0;
#endif
            // the stored responses: in flight during the eigen stage
            const float j0 = __ldcg(p.J + i0);
            const float j1 = two ? __ldcg(p.J + i1) : 3.4e38f;
            Eig3x2 e;
            eig_sym3<float2, true>(f[0], f[1], f[2], f[3], f[4], f[5], e);
            const float2 v = vesselness<float2, true>(e, p.k);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float vk = k ? v.y : v.x;
                if (vk > (k ? j1 : j0)) {
                    const size_t i = k ? i1 : i0;
                    p.J[i] = vk;
                    p.Vx[i] = (uint8_t)dir_code(k ? e.vx.y : e.vx.x);
                    p.Vy[i] = (uint8_t)dir_code(k ? e.vy.y : e.vy.x);
                    p.Vz[i] = (uint8_t)dir_code(k ? e.vz.y : e.vz.x);
                    if (p.scale_idx) p.scale_idx[i] = (uint8_t)p.scale;
                    if (p.dir) {
                        p.dir[i] = k ? e.vx.y : e.vx.x;
                        p.dir[p.voxels + i] = k ? e.vy.y : e.vy.x;
                        p.dir[2 * p.voxels + i] = k ? e.vz.y : e.vz.x;
                    }
                    vmax = fmaxf(vmax, vk);
                }
            }
        }
        head += count;
    };

    int rot = 0;
#pragma unroll 1
    for (int c = 0; c < nz; ++c) {                // centre plane z = zs + c; its window = sequence numbers c .. c+4
        const int z = zs + c;
        // issue duty (see s_next above); every plane is looked at by 8 / (K3C_DUTY + 1) of the warps
        if (tx == 0 && ((c + wid) & K3C_DUTY) == 0) {
            for (;;) {
                const int n = *reinterpret_cast<volatile int*>(s_next);
                if (n >= nseq || n > c + T::AHEAD) break;
                if (!mbar_test(empty_s + 8 * (n & 7), (unsigned)((n >> 3) - 1) & 1u)) break;   // its slot is still in use
                if (atomicCAS(s_next, n, n + 1) == n) issue(n);
            }
        }
        __syncwarp();
        if (c == 0) {
#pragma unroll 1
            for (int n = 0; n < 4; ++n) mbar_wait(full_s + 8 * n, 0u);
        }
        mbar_wait(full_s + 8 * ((c + 4) & 7), (unsigned)((c + 4) >> 3) & 1u);   // plane z+2 has landed
        const int o0 = (c & 7) * T::SLOT, o1 = ((c + 1) & 7) * T::SLOT, o2 = ((c + 2) & 7) * T::SLOT,
                  o3 = ((c + 3) & 7) * T::SLOT, o4 = ((c + 4) & 7) * T::SLOT;
        // ---- phase A: second differences, the diagonal-sum test, survivors appended to the warp's queue ----
        if (z < 2 || z > l - 3) phase_a_general(z, o0, o1, o2, o3, o4);
#if K3C_ROT5
        else switch (rot) {
            case 0: phase_a(IntC<0>(), ring + o0, ring + o1, ring + o2, ring + o3, ring + o4); break;
            case 1: phase_a(IntC<1>(), ring + o0, ring + o1, ring + o2, ring + o3, ring + o4); break;
            case 2: phase_a(IntC<2>(), ring + o0, ring + o1, ring + o2, ring + o3, ring + o4); break;
            case 3: phase_a(IntC<3>(), ring + o0, ring + o1, ring + o2, ring + o3, ring + o4); break;
            default: phase_a(IntC<4>(), ring + o0, ring + o1, ring + o2, ring + o3, ring + o4); break;
        }
        rot = rot == 4 ? 0 : rot + 1;
#else
        else {                                    // one instantiation, the rings shifted by register moves (24 per plane)
            phase_a(IntC<0>(), ring + o0, ring + o1, ring + o2, ring + o3, ring + o4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                C[0][k] = C[1][k]; C[1][k] = C[2][k]; C[2][k] = C[3][k]; C[3][k] = C[4][k];
                G[1][k] = G[2][k]; G[2][k] = G[3][k];
            }
        }
        (void)rot;
#endif
        off0 += plane_vox;
        // consumer release of the oldest plane of the window (sequence number c): the arrive has release semantics, so
        // this lane's loads of the slot are performed before the arrival is observed by the lane that refills it
        mbar_arrive(empty_s + 8 * (c & 7));
        // ---- phase B: full batches of this warp's queue ----
        while (tail - head >= (unsigned)Q::BATCH) drain(Q::BATCH);
    }
    if (tail != head) drain((int)(tail - head));  // the rest of the chunk (fewer than 64 entries)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
    if (tx == 0 && vmax > 0.0f) atomicMax(p.minmax + 1, __float_as_int(vmax));
}
#else   // K3C_V1: the first form (26 row segments per quad from shared memory, stored response streamed), kept for A/B timing
struct HessTileC {                               // tile of the compacting kernel: 128 x 8, one row per warp
    static constexpr int TX = 128, TY = 8, NT = 256;
    static constexpr int PW = HessTile::PW;      // same row pitch as HessTile (quad_hessians relies on it)
    static constexpr int PH = TY + 4;
    static constexpr int PLANE = PW * PH;        // 1584 floats: the TMA box
    static constexpr int SLOT = (PLANE * 4 + 127) / 128 * 32;   // 1600 floats
    static constexpr int SLOTS = K3C_SLOTS;
    static constexpr int RING_BYTES = SLOTS * SLOT * 4;    // 38400
    static constexpr int X_FIRST = 2;            // tile bx covers the voxels x = X_FIRST + TX * bx + [0, TX)
};
struct HessQueue {
    static constexpr int PAIRS = 2;                                  // packed pairs per thread per drain (ILP)
    static constexpr int BATCH = 2 * PAIRS * HessTileC::NT;          // 1024 entries per drain
    static constexpr int APPEND = HessTileC::TX * HessTileC::TY;     // most one plane can add (1024)
    // Appends only happen right after the plane barrier, when at most BATCH - 1 entries are pending
    // and no drain is in flight, so BATCH + APPEND ring entries can never collide; rounded up to a
    // power of two so that the ring index is a mask.
    static constexpr int CAP = 2048;
    static_assert(CAP >= BATCH + APPEND, "queue too small");
    // Structure of arrays: field f of entry e sits at float f * CAP + e, fields = Dxx, Dxy, Dxz, Dyy, Dyz, Dzz, stored J,
    // packed position.  Appends are eight predicated 32-bit stores; a drain reads entries (e, e + 1), e even, as one
    // 64-bit load per field, which IS the packed register pair of the eigen stage.
    static constexpr int FIELDS = 8;
    static constexpr int BYTES = CAP * FIELDS * 4;
    static constexpr int SMEM_BYTES = HessTileC::RING_BYTES + BYTES + HessTileC::SLOTS * 8 + 16;   // + mbarriers + tail counter
};

// eight predicated 32-bit shared stores of one queue entry (no branch, no register marshalling)
__device__ __forceinline__ void queue_append(uint32_t addr, bool pred, float d0, float d1, float d2, float d3, float d4,
                                             float d5, float jold, int pos)
{
    constexpr int S = HessQueue::CAP * 4;          // byte stride between the field arrays
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "@p st.shared.f32 [%0], %2;\n\t"
        "@p st.shared.f32 [%0 + %10], %3;\n\t"
        "@p st.shared.f32 [%0 + %11], %4;\n\t"
        "@p st.shared.f32 [%0 + %12], %5;\n\t"
        "@p st.shared.f32 [%0 + %13], %6;\n\t"
        "@p st.shared.f32 [%0 + %14], %7;\n\t"
        "@p st.shared.f32 [%0 + %15], %8;\n\t"
        "@p st.shared.b32 [%0 + %16], %9;\n\t}"
        ::"r"(addr), "r"((uint32_t)pred), "f"(d0), "f"(d1), "f"(d2), "f"(d3), "f"(d4), "f"(d5), "f"(jold), "r"(pos),
          "n"(S), "n"(2 * S), "n"(3 * S), "n"(4 * S), "n"(5 * S), "n"(6 * S), "n"(7 * S) : "memory");
}

__global__ void __launch_bounds__(HessTileC::NT, 2)
hessian_eigen_compact_kernel(const __grid_constant__ VoxelParams p)
{
    using T = HessTileC;
    using Q = HessQueue;
    extern __shared__ __align__(128) float ring[];
    float* qf = ring + T::SLOTS * T::SLOT;                                 // FIELDS arrays of CAP floats
    uint64_t* mbar = reinterpret_cast<uint64_t*>(qf + Q::FIELDS * Q::CAP);
    unsigned* s_tail = reinterpret_cast<unsigned*>(mbar + T::SLOTS);       // entries ever appended (ring index = tail % CAP)
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    int bid = blockIdx.x;
    const int bx = bid % p.ntx; bid /= p.ntx;
    const int by = bid % p.nty;
    const int bz = bid / p.nty;
    const int x0 = bx * T::TX, y0 = by * T::TY - 2;           // tile bx covers x = 2 + TX * bx + [0, TX), see hessian_eigen_kernel
    const int w = p.f.w, h = p.f.h, l = p.f.l;
    const int zs = p.z_begin + bz * p.zchunk;
    const int ze = min(zs + p.zchunk, p.z_begin + p.nz);
    if (zs >= ze) return;
    if (tid == 0) *s_tail = 0;

    TileRing<T> tr;
    tr.init(ring, mbar, tid);
    __syncthreads();
    if (tid == 0)
        for (int q = zs - 2; q <= zs + 2; ++q) tr.issue(&p.tmap, x0, y0, q, p.f.base, l);
#pragma unroll 1
    for (int k = 0; k < 4; ++k) tr.wait_next();

    const int xq = bx * T::TX + 2 + 4 * tx;
    bool m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = xq + j <= w - 3;
    const bool any_x = m[0] || m[1] || m[2] || m[3];
    const bool vec_j = p.vec_ok && xq + 3 < w;
    float vmax = 0.0f;
    unsigned head = 0;        // entries [head, tail) are pending; every thread carries the same value
    // The stored response of the thread's quad is fetched one plane ahead, so that the load is in
    // flight during a whole plane of work instead of stalling the append.
    const int yl = ty;                            // one tile row per warp
    const int yrow = by * T::TY + yl;
    const bool row_ok = yrow >= 2 && yrow <= h - 3 && any_x;
    const long long plane_vox = (long long)h * w;
    const float* jp = p.J + ((long long)(zs - p.z_begin) * h + yrow) * w + xq;   // the quad's stored J at plane z (dereferenced only if row_ok)
    float jnext[4] = { 0.f, 0.f, 0.f, 0.f };
    auto load_j = [&](const float* src) {
        if (vec_j) {                              // xq = 2 mod 4: the quad is two aligned pairs
            *reinterpret_cast<float2*>(jnext) = __ldcs(reinterpret_cast<const float2*>(src));
            *reinterpret_cast<float2*>(jnext + 2) = __ldcs(reinterpret_cast<const float2*>(src + 2));
        } else
#pragma unroll
            for (int j = 0; j < 4; ++j) if (m[j]) jnext[j] = src[j];
    };
    if (row_ok) load_j(jp);
    const uint32_t q_s = smem_u32(qf);
    const int o_quad = (yl + 2) * T::PW + 4 * tx;
    int pos0 = (yl << 7) | (4 * tx);              // + (z - zs) << 10
    ZWin zwin;
    bool win_ok = false;

    // Phase B on entries [first, first + count) (ring positions, first even): PAIRS independent packed pairs per thread
    auto drain = [&](unsigned first, int count) {
#pragma unroll
        for (int q = 0; q < Q::PAIRS; ++q) {
            const int o = 2 * (tid + q * T::NT);           // entries o, o + 1 of the batch
            if (o >= count) continue;
            const bool two = o + 1 < count;
            const float* e0 = qf + ((first + o) & (Q::CAP - 1));
            float2 f[Q::FIELDS];
#pragma unroll
            for (int k = 0; k < Q::FIELDS; ++k) f[k] = *reinterpret_cast<const float2*>(e0 + k * Q::CAP);
            Eig3x2 e;
            eig_sym3<float2, true>(f[0], f[1], f[2], f[3], f[4], f[5], e);
            const float2 v = vesselness<float2, true>(e, p.k);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k == 1 && !two) break;
                const float vk = k ? v.y : v.x;
                if (vk > (k ? f[6].y : f[6].x)) {
                    const int pos = __float_as_int(k ? f[7].y : f[7].x);   // (z - zs) << 10 | row << 7 | column
                    const int x = bx * T::TX + 2 + (pos & 127), y = by * T::TY + ((pos >> 7) & 7), z = zs + (pos >> 10);
                    const long long i = ((long long)(z - p.z_begin) * h + y) * w + x;
                    p.J[i] = vk;
                    p.Vx[i] = (uint8_t)dir_code(k ? e.vx.y : e.vx.x);
                    p.Vy[i] = (uint8_t)dir_code(k ? e.vy.y : e.vy.x);
                    p.Vz[i] = (uint8_t)dir_code(k ? e.vz.y : e.vz.x);
                    if (p.scale_idx) p.scale_idx[i] = (uint8_t)p.scale;
                    if (p.dir) {
                        p.dir[i] = k ? e.vx.y : e.vx.x;
                        p.dir[p.voxels + i] = k ? e.vy.y : e.vy.x;
                        p.dir[2 * p.voxels + i] = k ? e.vz.y : e.vz.x;
                    }
                    vmax = fmaxf(vmax, vk);
                }
            }
        }
    };

    for (int z = zs; z < ze; ++z) {
        __syncthreads();                          // nobody reads plane z-3's slot or drains any more
        if (tid == 0 && z + 1 < ze) tr.issue(&p.tmap, x0, y0, z + 3, p.f.base, l);
        tr.wait_next();                           // plane z+2 has landed
        const bool z_general = z < 2 || z > l - 3;
        // ---- phase A: second differences, the diagonal-sum test, append survivors ----
        bool surv[4] = { false, false, false, false };
        float2 Hxx[2], Hxy[2], Hxz[2], Hyy[2], Hyz[2], Hzz[2];
        const float jold[4] = { jnext[0], jnext[1], jnext[2], jnext[3] };
        jp += plane_vox;
        if (row_ok && z + 1 < ze) load_j(jp);
        if (row_ok) {
            if (z_general) { quad_hessians<true>(ring, z_planes(tr.o, z, l, p.k.sigma2), o_quad, Hxx, Hxy, Hxz, Hyy, Hyz, Hzz); win_ok = false; }
            else { quad_hessians_reuse(ring, z_planes_interior(tr.o, p.k.sigma2), o_quad, win_ok, zwin, Hxx, Hxy, Hxz, Hyy, Hyz, Hzz); win_ok = true; }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float2 sxy = vadd(Hxx[g], Hyy[g]), sxz = vadd(Hxx[g], Hzz[g]), syz = vadd(Hyy[g], Hzz[g]);
                surv[2 * g] = m[2 * g] && fmaxf(fmaxf(sxy.x, sxz.x), syz.x) <= 0.0f;
                surv[2 * g + 1] = m[2 * g + 1] && fmaxf(fmaxf(sxy.y, sxz.y), syz.y) <= 0.0f;
            }
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, surv[0]), b1 = __ballot_sync(0xffffffffu, surv[1]);
        const unsigned b2 = __ballot_sync(0xffffffffu, surv[2]), b3 = __ballot_sync(0xffffffffu, surv[3]);
        const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
        unsigned base = 0;
        if (tx == 0 && n0 + n1 + n2 + n3 > 0) base = atomicAdd(s_tail, (unsigned)(n0 + n1 + n2 + n3));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned below = (1u << tx) - 1u;
        const unsigned slot[4] = { base + __popc(b0 & below), base + n0 + __popc(b1 & below),
                                   base + n0 + n1 + __popc(b2 & below), base + n0 + n1 + n2 + __popc(b3 & below) };
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            typedef Lanes<float2> L2;
            queue_append(q_s + 4 * (slot[j] & (Q::CAP - 1)), surv[j],
                         L2::get(Hxx[j >> 1], j & 1), L2::get(Hxy[j >> 1], j & 1), L2::get(Hxz[j >> 1], j & 1),
                         L2::get(Hyy[j >> 1], j & 1), L2::get(Hyz[j >> 1], j & 1), L2::get(Hzz[j >> 1], j & 1),
                         jold[j], pos0 + j);
        }
        pos0 += 1 << 10;
        tr.rotate();
        __syncthreads();                          // appended entries and the tail are visible
        // ---- phase B: full batches from the head of the queue (everything after the last plane) ----
        const unsigned tail = *reinterpret_cast<volatile unsigned*>(s_tail);
        const bool flush = z + 1 == ze;
        while (tail - head >= (unsigned)Q::BATCH || (flush && tail != head)) {
            const int take = (int)min(tail - head, (unsigned)Q::BATCH);
            drain(head, take);
            head += take;
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
    if (tx == 0 && vmax > 0.0f) atomicMax(p.minmax + 1, __float_as_int(vmax));
}

#endif  // K3C_V1

// K3b: the shell.  Thread index -> region: z faces (whole planes), then y faces
// (rows of the own planes), then x faces (columns of the own planes).  Edges
// and corners are visited by more than one region, which is harmless: the
// update is idempotent (same value on the first scale, v > J false afterwards).
template <int MODE>
__global__ void __launch_bounds__(128)
hessian_eigen_shell_kernel(const __grid_constant__ VoxelParams p)
{
    long long t = (long long)blockIdx.x * 128 + threadIdx.x;
    const int w = p.f.w, h = p.f.h;
    int x, y, z;
    bool active = true;
    if (t < p.n_zface) {
        x = (int)(t % w); t /= w;
        y = (int)(t % h);
        z = p.zf[(int)(t / h)];
    } else if ((t -= p.n_zface) < p.n_yface) {
        x = (int)(t % w); t /= w;
        y = p.yf[(int)(t % p.nyf)];
        z = p.z_begin + (int)(t / p.nyf);
    } else if ((t -= p.n_yface) < p.n_xface) {
        x = p.xf[(int)(t % p.nxf)]; t /= p.nxf;
        y = (int)(t % h);
        z = p.z_begin + (int)(t / h);
    } else {
        active = false; x = y = 0; z = p.z_begin;
    }
    float jv = 0.0f;
    if (active) {
        const long long i = ((long long)(z - p.z_begin) * h + y) * w + x;
        if (MODE == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.J + i));    // the stored response: in flight beside the taps
        const Hess H = hessian_at_face(p.f, x, y, z, p.k.sigma2);
        if (MODE == 2) {
            p.D[0][i] = H.zz; p.D[1][i] = H.yy; p.D[2][i] = H.yz;
            p.D[3][i] = H.xx; p.D[4][i] = H.xy; p.D[5][i] = H.xz;
            return;
        }
        jv = voxel_update<MODE>(p, i, H);
    }
    if (MODE == 2) return;
    if (MODE == 0) {
        float mn = active ? jv : 3.4e38f;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
        // one atomic per warp on ONE address serialises in L2 (262 144 warps: the launch took as long as the atomics);
        // min / max only ever move one way, so a warp whose value cannot change the result (a plain read says so) skips it
        if ((threadIdx.x & 31) == 0 && mn < 3.0e38f && __float_as_int(mn) < *reinterpret_cast<volatile int*>(p.minmax + 0))
            atomicMin(p.minmax + 0, __float_as_int(mn));
    }
    {
        float mx = active ? jv : 0.0f;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        if ((threadIdx.x & 31) == 0 && __float_as_int(mx) > *reinterpret_cast<volatile int*>(p.minmax + 1))
            atomicMax(p.minmax + 1, __float_as_int(mx));
    }
}

// K9 reference_direction_kernel (FRANGI_GPU_FLAG_REFERENCE_DIRECTION), after the K3 launches of a scale.
// The direction the reference writes is column 0 of what its double-precision Householder / QL solver returns
// (frangi.cpp:198, 239-250, 1269-1495) -- an eigenvector whose SIGN is whatever that iteration ends on.  The closed form
// of K3 returns the same axis (within 0.5 degrees) with the other sign in half of the voxels, and downstream code is not
// indifferent to it (tests/test_plugin_e2e.py).  This pass re-derives, for every voxel whose running maximum scale `p.scale`
// has just taken (scale_idx == p.scale; on the first scale: every voxel), the second differences with the reference's
// face rules from F, runs the reference's solver on them (ref_eigen.h, rdouble: separately rounded double operations in
// the reference's order) and writes the reference's three direction bytes, round(((v + 1) / 2) * 255) in double.  With
// bit-exact smoothing the second differences are the reference's bit for bit, hence so are the bytes.  One thread per
// voxel from global memory, double precision: an opt-in parity mode, not the fast path.
__global__ void __launch_bounds__(128)
reference_direction_kernel(const __grid_constant__ VoxelParams p)
{
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= p.voxels) return;
    if (p.scale_idx[i] != (uint8_t)p.scale) return;
    const int w = p.f.w, h = p.f.h;
    const int x = (int)(i % w);
    const long long row = i / w;
    const int y = (int)(row % h);
    const int z = p.z_begin + (int)(row / h);
    const Hess H = hessian_at_face(p.f, x, y, z, p.k.sigma2);
    rdouble A[3][3], V[3][3], d[3];
    A[0][0] = rdouble((double)H.xx); A[0][1] = rdouble((double)H.xy); A[0][2] = rdouble((double)H.xz);
    A[1][0] = rdouble((double)H.xy); A[1][1] = rdouble((double)H.yy); A[1][2] = rdouble((double)H.yz);
    A[2][0] = rdouble((double)H.xz); A[2][1] = rdouble((double)H.yz); A[2][2] = rdouble((double)H.zz);
    ref_eigen_decomposition(A, V, d);
    uint8_t code[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const rdouble t = ((V[c][0] + rdouble(1.0)) / rdouble(2.0)) * rdouble(255.0);
        const double r = round(t.v);                               // half away from zero, as the C library's
        code[c] = r >= 255.0 ? (uint8_t)255 : (r > 0.0 ? (uint8_t)(int)r : (uint8_t)0);      // NaN -> 0 (x86: INT_MIN, clamped)
    }
    p.Vx[i] = code[0]; p.Vy[i] = code[1]; p.Vz[i] = code[2];
    if (p.dir) {
        p.dir[i] = (float)V[0][0].v; p.dir[p.voxels + i] = (float)V[1][0].v; p.dir[2 * p.voxels + i] = (float)V[2][0].v;
    }
}

// Stage kernel: eigen + vesselness on caller-supplied Hessians.
__global__ void __launch_bounds__(128)
vesselness_stage_kernel(const float* __restrict__ Dxx, const float* __restrict__ Dxy,
                        const float* __restrict__ Dxz, const float* __restrict__ Dyy,
                        const float* __restrict__ Dyz, const float* __restrict__ Dzz, long long n,
                        FrangiConsts k, float* __restrict__ v_out, float* __restrict__ dir_out,
                        float* __restrict__ lambda_out)
{
    // two matrices per thread through the packed path; an odd tail through the scalar one
    const long long i = 2 * ((long long)blockIdx.x * 128 + threadIdx.x);
    if (i >= n) return;
    float v[2], ev[2][3], el[2][3];
    if (i + 1 < n) {
        Eig3x2 e;
        eig_sym3<float2>(make_float2(Dxx[i], Dxx[i + 1]), make_float2(Dxy[i], Dxy[i + 1]), make_float2(Dxz[i], Dxz[i + 1]),
                         make_float2(Dyy[i], Dyy[i + 1]), make_float2(Dyz[i], Dyz[i + 1]), make_float2(Dzz[i], Dzz[i + 1]), e);
        const float2 vv = vesselness<float2>(e, k);
        v[0] = vv.x; v[1] = vv.y;
        ev[0][0] = e.vx.x; ev[0][1] = e.vy.x; ev[0][2] = e.vz.x; ev[1][0] = e.vx.y; ev[1][1] = e.vy.y; ev[1][2] = e.vz.y;
        el[0][0] = e.l1.x; el[0][1] = e.l2.x; el[0][2] = e.l3.x; el[1][0] = e.l1.y; el[1][1] = e.l2.y; el[1][2] = e.l3.y;
    } else {
        Eig3 e;
        eig_sym3<float>(Dxx[i], Dxy[i], Dxz[i], Dyy[i], Dyz[i], Dzz[i], e);
        v[0] = vesselness<float>(e, k);
        ev[0][0] = e.vx; ev[0][1] = e.vy; ev[0][2] = e.vz; el[0][0] = e.l1; el[0][1] = e.l2; el[0][2] = e.l3;
    }
    for (int q = 0; q < 2 && i + q < n; ++q) {
        v_out[i + q] = v[q];
        if (dir_out) { dir_out[i + q] = ev[q][0]; dir_out[n + i + q] = ev[q][1]; dir_out[2 * n + i + q] = ev[q][2]; }
        if (lambda_out) { lambda_out[3 * (i + q)] = el[q][0]; lambda_out[3 * (i + q) + 1] = el[q][1]; lambda_out[3 * (i + q) + 2] = el[q][2]; }
    }
}

// Scalar twin of the stage kernel (the path of the face shell and of ragged quads).
__global__ void __launch_bounds__(128)
vesselness_stage_scalar_kernel(const float* __restrict__ Dxx, const float* __restrict__ Dxy,
                               const float* __restrict__ Dxz, const float* __restrict__ Dyy,
                               const float* __restrict__ Dyz, const float* __restrict__ Dzz, long long n,
                               FrangiConsts k, float* __restrict__ v_out, float* __restrict__ dir_out,
                               float* __restrict__ lambda_out)
{
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    Eig3 e;
    eig_sym3<float>(Dxx[i], Dxy[i], Dxz[i], Dyy[i], Dyz[i], Dzz[i], e);
    v_out[i] = vesselness<float>(e, k);
    if (dir_out) { dir_out[i] = e.vx; dir_out[n + i] = e.vy; dir_out[2 * n + i] = e.vz; }
    if (lambda_out) { lambda_out[3 * i] = e.l1; lambda_out[3 * i + 1] = e.l2; lambda_out[3 * i + 2] = e.l3; }
}

// ---------------------------------------------------------------------------
// K4: J -> J8 (Advantra_plugin.cpp:2499-2512, round() from :120-123).
// minmax holds the float bit patterns of Jmin / Jmax (after the all-reduce).
// ---------------------------------------------------------------------------
// One voxel of K4.  The reference computes round(((J-Jmin)/(Jmax-Jmin))*255) with an IEEE float
// division, a float multiply and round-half-away (Advantra_plugin.cpp:120-123, 2508).
//  * division by the loop-invariant range: q0 = a*y, r = fma(-q0, b, a), q = fma(r, y, q0) with
//    y = RN(1/b) is the correctly rounded quotient (Markstein) unless b's significand is all ones
//    or the operands leave the comfortable exponent range; `fast` is false in those cases and the
//    plain division is used;
//  * rounding without conversions: t = qf + 2^23 rounds to nearest-even in the adder, the exact
//    difference tells a tie that went down, and the byte sits in the low mantissa bits of t.
__device__ __forceinline__ uint32_t j8_code(float j, float lo, float range, float y, bool fast, bool flat)
{
    const float a = __fsub_rn(j, lo);
    float q;
    if (fast) {
        const float q0 = __fmul_rn(a, y);
        const float r = __fmaf_rn(-q0, range, a);
        q = __fmaf_rn(r, y, q0);
    } else {
        q = __fdiv_rn(a, range);
    }
    const float qf = fminf(fmaxf(__fmul_rn(q, 255.0f), 0.0f), 255.0f);   // the reference clamps after rounding: same bytes
    const float t = __fadd_rn(qf, 8388608.0f);
    const float rn = __fsub_rn(t, 8388608.0f);
    const uint32_t v = (__float_as_uint(t) & 0x1ffu) + (__fsub_rn(qf, rn) == 0.5f ? 1u : 0u);
    return flat ? 0u : v;
}

__global__ void __launch_bounds__(256)
j_to_j8_kernel(const float* __restrict__ J, uint8_t* __restrict__ J8, long long n, const int* __restrict__ minmax)
{
    const float lo = __int_as_float(minmax[0]);
    const float hi = __int_as_float(minmax[1]);
    const float range = __fsub_rn(hi, lo);
    const bool flat = fabsf(range) <= 1.175494351e-38f;  // FLT_MIN
    const float y = __frcp_rn(range);
    const uint32_t rb = __float_as_uint(range);
    const int ex = (int)((rb >> 23) & 0xffu);
    const bool fast = (rb & 0x7fffffu) != 0x7fffffu && ex > 64 && ex < 190 && fabsf(lo) < 1e18f && fabsf(hi) < 1e18f;
    const long long n16 = n / 16;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n16; g += stride) {
        const float4* src = reinterpret_cast<const float4*>(J) + 4 * g;
        float4 q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = __ldcs(src + k);
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = j8_code(q[k].x, lo, range, y, fast, flat) | (j8_code(q[k].y, lo, range, y, fast, flat) << 8) |
                   (j8_code(q[k].z, lo, range, y, fast, flat) << 16) | (j8_code(q[k].w, lo, range, y, fast, flat) << 24);
        __stcs(reinterpret_cast<uint4*>(J8) + g, make_uint4(o[0], o[1], o[2], o[3]));
    }
    for (long long i = 16 * n16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        J8[i] = (uint8_t)j8_code(J[i], lo, range, y, fast, flat);
}

// Jmin = FLT_MAX, Jmax = -FLT_MAX as bit patterns (frangi.cpp:176-177) at the start of a run
__global__ void minmax_init_kernel(int* __restrict__ minmax)
{
    if (threadIdx.x == 0) { minmax[0] = 0x7f7fffff; minmax[1] = (int)0xff7fffffu; }
}

// min / max over the per-slab pairs gathered on one device (local-copy multi-slab mode)
__global__ void minmax_reduce_kernel(const int* __restrict__ pairs, int n, int* __restrict__ out)
{
    if (threadIdx.x != 0) return;
    int lo = pairs[0], hi = pairs[1];
    for (int k = 1; k < n; ++k) { lo = min(lo, pairs[2 * k]); hi = max(hi, pairs[2 * k + 1]); }
    out[0] = lo; out[1] = hi;
}

}  // namespace frangi
