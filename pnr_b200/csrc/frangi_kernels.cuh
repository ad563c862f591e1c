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
#include <cuda_runtime.h>
#include <stdint.h>

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
    const int z = bid / (p.nstrips * p.nsegs);
    const int x0 = strip * C::TW;
    const int ys = seg * p.seg_h;
    const int ye = min(ys + p.seg_h, p.h);
    const uint8_t* __restrict__ Iz = p.I + (long long)z * p.w * p.h;
    float* __restrict__ Oz = p.out + (long long)z * p.fplane;

    const int xr = tid & 15;   // x pass: row within the batch
    const int xc = tid >> 4;   // x pass: 16-column chunk
    const int nb = (ye - ys + C::RB - 1) / C::RB;
    int phase = 0;             // x-pass phases done; phase m covers rows ys - L + m*RB + [0, RB)

    for (int b = 0; b < nb; ++b) {
        const int need = b + C::NBLK;
        while (phase < need) {
            __syncthreads();  // s_in and the ring block about to be overwritten are no longer read
            // ---- stage RB input rows as float ----
            const int r0 = ys - L + phase * C::RB;
            constexpr int GROUPS = C::PIN0 / 4;
            for (int idx = tid; idx < C::RB * GROUPS; idx += C::NT) {
                const int r = idx / GROUPS;
                const int g = idx - r * GROUPS;
                const int y = clampi(r0 + r, 0, p.h - 1);
                const int xg = x0 - C::LAL + 4 * g;
                const uint8_t* row = Iz + (long long)y * p.w;
                float4 v;
                if (p.vec_ok && xg >= 0 && xg + 3 < p.w) {
                    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(row + xg));
                    v.x = (float)(u & 0xffu);
                    v.y = (float)((u >> 8) & 0xffu);
                    v.z = (float)((u >> 16) & 0xffu);
                    v.w = (float)(u >> 24);
                } else {
                    v.x = (float)__ldg(row + clampi(xg + 0, 0, p.w - 1));
                    v.y = (float)__ldg(row + clampi(xg + 1, 0, p.w - 1));
                    v.z = (float)__ldg(row + clampi(xg + 2, 0, p.w - 1));
                    v.w = (float)__ldg(row + clampi(xg + 3, 0, p.w - 1));
                }
                *reinterpret_cast<float4*>(s_in + r * C::PIN + 4 * g) = v;
            }
            __syncthreads();
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
            ++phase;
        }
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
// Per-voxel math of K3.
// ---------------------------------------------------------------------------
struct FrangiConsts {
    float inv_2a2;    // 1 / (2*alpha*alpha)   (float products as in frangi.cpp:215-217)
    float inv_2b2;
    float inv_2c2;
    float sigma2;     // sigma*sigma, float    (frangi.cpp:319)
    int blackwhite;
};

// 1 - exp(-x) for x >= 0 without cancellation (the reference evaluates
// 1 - exp(-x) in double; in float32 the subtraction would lose everything for
// the S term, where x ~ 1e-4 with C = 500).
__device__ __forceinline__ float one_minus_exp_neg(float x) { return -expm1f(-x); }

struct Eig3 {
    float l1, l2, l3;   // |l1| <= |l2| <= |l3| with the reference's tie rules
    float vx, vy, vz;   // unit eigenvector of l1
};

__device__ __forceinline__ void swapf(float& a, float& b) { float t = a; a = b; b = t; }

// Symmetric 3x3 eigen-decomposition, float32, non-iterative.
//  1. the eigenvalue at the isolated end of the spectrum from the
//     trigonometric (Cardano) formula, where it is well conditioned;
//  2. its eigenvector from the best-conditioned cross product of two rows of
//     (A - lambda I);
//  3. the other two eigenvalues (and, when needed, eigenvector) from the 2x2
//     projection of A on the orthogonal complement: their split is then a sum
//     of squares, not a cancelling difference.
// Replaces eigen_decomposition (frangi.cpp:1269-1306) semantically: returns
// the |lambda|-sorted eigenvalues and column 0 of V.
__device__ __forceinline__ void eig_sym3(float a00, float a01, float a02, float a11, float a12,
                                         float a22, Eig3& out)
{
    float e0, e1, e2;                    // ascending eigenvalues
    float v0x, v0y, v0z;                 // eigenvectors of e0 / e1 / e2 (built lazily below)
    float v1x, v1y, v1z, v2x, v2y, v2z;
    const float off = a01 * a01 + a02 * a02 + a12 * a12;
    if (off == 0.0f) {
        // Diagonal input: the reference's QL leaves the values and the identity
        // untouched, then selection-sorts ascending (first minimum wins ties).
        e0 = a00; e1 = a11; e2 = a22;
        v0x = 1.f; v0y = 0.f; v0z = 0.f;
        v1x = 0.f; v1y = 1.f; v1z = 0.f;
        v2x = 0.f; v2y = 0.f; v2z = 1.f;
        // i = 0: pick the first strict minimum of (e0,e1,e2)
        int k = 0; float pv = e0;
        if (e1 < pv) { k = 1; pv = e1; }
        if (e2 < pv) { k = 2; pv = e2; }
        if (k == 1) { swapf(e0, e1); swapf(v0x, v1x); swapf(v0y, v1y); swapf(v0z, v1z); }
        else if (k == 2) { swapf(e0, e2); swapf(v0x, v2x); swapf(v0y, v2y); swapf(v0z, v2z); }
        if (e2 < e1) { swapf(e1, e2); swapf(v1x, v2x); swapf(v1y, v2y); swapf(v1z, v2z); }
    } else {
        const float tr = a00 + a11 + a22;
        const float q = tr * (1.0f / 3.0f);
        const float b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
        const float p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0f * off;
        const float p = sqrtf(p2 * (1.0f / 6.0f));
        const float ip = 1.0f / p;
        const float c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip;
        const float c01 = a01 * ip, c02 = a02 * ip, c12 = a12 * ip;
        float hd = 0.5f * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) +
                           c02 * (c01 * c12 - c11 * c02));
        hd = fminf(fmaxf(hd, -1.0f), 1.0f);
        const float phi = acosf(hd) * (1.0f / 3.0f);
        const bool top = hd >= 0.0f;   // the largest eigenvalue is the isolated one
        const float beta = top ? 2.0f * cosf(phi) : 2.0f * cosf(phi + 2.0943951023931953f);
        const float lam = q + p * beta;
        // eigenvector of the isolated eigenvalue
        const float r00 = a00 - lam, r11 = a11 - lam, r22 = a22 - lam;
        float cx0 = a01 * a12 - a02 * r11, cy0 = a02 * a01 - r00 * a12, cz0 = r00 * r11 - a01 * a01;  // r0 x r1
        float cx1 = a01 * r22 - a02 * a12, cy1 = a02 * a02 - r00 * r22, cz1 = r00 * a12 - a01 * a02;  // r0 x r2
        float cx2 = r11 * r22 - a12 * a12, cy2 = a12 * a02 - a01 * r22, cz2 = a01 * a12 - r11 * a02;  // r1 x r2
        const float n0 = cx0 * cx0 + cy0 * cy0 + cz0 * cz0;
        const float n1 = cx1 * cx1 + cy1 * cy1 + cz1 * cz1;
        const float n2 = cx2 * cx2 + cy2 * cy2 + cz2 * cz2;
        float nx = cx0, ny = cy0, nz = cz0, nn = n0;
        if (n1 > nn) { nx = cx1; ny = cy1; nz = cz1; nn = n1; }
        if (n2 > nn) { nx = cx2; ny = cy2; nz = cz2; nn = n2; }
        const float inn = rsqrtf(nn);
        const float ix = nx * inn, iy = ny * inn, iz = nz * inn;
        // orthonormal complement (u, w) of i
        float ux, uy, uz;
        if (fabsf(ix) > fabsf(iy)) {
            const float s = rsqrtf(ix * ix + iz * iz);
            ux = -iz * s; uy = 0.0f; uz = ix * s;
        } else {
            const float s = rsqrtf(iy * iy + iz * iz);
            ux = 0.0f; uy = iz * s; uz = -iy * s;
        }
        const float wx = iy * uz - iz * uy, wy = iz * ux - ix * uz, wz = ix * uy - iy * ux;
        // 2x2 projection
        const float aux = a00 * ux + a01 * uy + a02 * uz;
        const float auy = a01 * ux + a11 * uy + a12 * uz;
        const float auz = a02 * ux + a12 * uy + a22 * uz;
        const float awx = a00 * wx + a01 * wy + a02 * wz;
        const float awy = a01 * wx + a11 * wy + a12 * wz;
        const float awz = a02 * wx + a12 * wy + a22 * wz;
        const float m00 = ux * aux + uy * auy + uz * auz;
        const float m01 = wx * aux + wy * auy + wz * auz;
        const float m11 = wx * awx + wy * awy + wz * awz;
        const float mean = 0.5f * (m00 + m11);
        const float hdiff = 0.5f * (m00 - m11);
        const float disc = sqrtf(hdiff * hdiff + m01 * m01);
        const float la = mean - disc, lb = mean + disc;
        const float li = tr - (m00 + m11);   // Rayleigh-consistent isolated eigenvalue
        // eigenvectors of la / lb inside span(u, w)
        // (M - l I) x = 0  ->  x orthogonal to the larger of the two rows
        float xa0, xa1;
        {
            const float d0 = m00 - la, d1 = m11 - la;
            if (fabsf(d0) >= fabsf(d1)) { xa0 = -m01; xa1 = d0; } else { xa0 = d1; xa1 = -m01; }
            const float nrm = xa0 * xa0 + xa1 * xa1;
            if (nrm > 0.0f) { const float s = rsqrtf(nrm); xa0 *= s; xa1 *= s; }
            else { xa0 = 1.0f; xa1 = 0.0f; }
        }
        const float ax = xa0 * ux + xa1 * wx, ay = xa0 * uy + xa1 * wy, az = xa0 * uz + xa1 * wz;
        // lb's eigenvector is orthogonal to la's inside the plane
        const float bx = -xa1 * ux + xa0 * wx, by = -xa1 * uy + xa0 * wy, bz = -xa1 * uz + xa0 * wz;
        if (top) {
            e0 = la; e1 = lb; e2 = li;
            v0x = ax; v0y = ay; v0z = az; v1x = bx; v1y = by; v1z = bz; v2x = ix; v2y = iy; v2z = iz;
        } else {
            e0 = li; e1 = la; e2 = lb;
            v0x = ix; v0y = iy; v0z = iz; v1x = ax; v1y = ay; v1z = az; v2x = bx; v2y = by; v2z = bz;
        }
        // rounding can misorder a nearly triple eigenvalue; restore ascending order
        if (e1 < e0) { swapf(e0, e1); swapf(v0x, v1x); swapf(v0y, v1y); swapf(v0z, v1z); }
        if (e2 < e1) { swapf(e1, e2); swapf(v1x, v2x); swapf(v1y, v2y); swapf(v1z, v2z); }
        if (e1 < e0) { swapf(e0, e1); swapf(v0x, v1x); swapf(v0y, v1y); swapf(v0z, v1z); }
    }
    // re-order by absolute value with the reference's rules (frangi.cpp:1284-1304)
    float d0 = e0, d1 = e1, d2 = e2;
    float m0 = fabsf(d0), m1 = fabsf(d1), m2 = fabsf(d2);
    if (m0 >= m1 && m0 > m2) {
        swapf(d0, d2); swapf(m0, m2); swapf(v0x, v2x); swapf(v0y, v2y); swapf(v0z, v2z);
    } else if (m1 >= m0 && m1 > m2) {
        swapf(d1, d2); swapf(m1, m2); swapf(v1x, v2x); swapf(v1y, v2y); swapf(v1z, v2z);
    }
    if (m0 > m1) {
        swapf(d0, d1); swapf(m0, m1); swapf(v0x, v1x); swapf(v0y, v1y); swapf(v0z, v1z);
    }
    out.l1 = d0; out.l2 = d1; out.l3 = d2;
    out.vx = v0x; out.vy = v0y; out.vz = v0z;
}

// Frangi vesselness from |lambda|-sorted eigenvalues (frangi.cpp:206-231).
__device__ __forceinline__ float vesselness(const Eig3& e, const FrangiConsts& k)
{
    const float a1 = fabsf(e.l1), a2 = fabsf(e.l2), a3 = fabsf(e.l3);
    const float Ra = a2 / a3;
    const float Rb = a1 / sqrtf(a2 * a3);
    const float S2 = a1 * a1 + a2 * a2 + a3 * a3;
    const float tRa = one_minus_exp_neg(Ra * Ra * k.inv_2a2);
    const float tRb = expf(-(Rb * Rb) * k.inv_2b2);
    const float tS = one_minus_exp_neg(S2 * k.inv_2c2);
    float v = tRa * tRb * tS;
    if (k.blackwhite) {
        if (e.l2 < 0.0f) v = 0.0f;
        if (e.l3 < 0.0f) v = 0.0f;
    } else {
        if (e.l2 > 0.0f) v = 0.0f;
        if (e.l3 > 0.0f) v = 0.0f;
    }
    if (!(v == v)) v = 0.0f;  // NaN (0/0 on a zero Hessian) -> 0, frangi.cpp:231
    return v;
}

// round((c+1)/2*255) clamped to a byte (frangi.cpp:240-250); arguments are in [-1,1]
__device__ __forceinline__ uint8_t dir_code(float c)
{
    const float t = (c + 1.0f) * 0.5f * 255.0f;
    int v = (int)floorf(t + 0.5f);
    v = min(max(v, 0), 255);
    return (uint8_t)v;
}

// ---------------------------------------------------------------------------
// Second differences with the reference's face rules (frangi.cpp:306-381):
// first difference along an axis = s * (f[hi] - f[lo]) with lo = max(c-1,0),
// hi = min(c+1,n-1), s = 1 on a face and 0.5 inside; the second difference
// applies the same rule to the first-difference field; then * sigma^2.
// Coordinates are GLOBAL (slab faces are not volume faces).
// ---------------------------------------------------------------------------
struct FView {
    const float* F;     // plane 0 = global plane base
    int w, h, l;
    int fpitch;
    long long fplane;
    int base;
    __device__ __forceinline__ float at(int x, int y, int z) const
    {
        return __ldg(F + (long long)(z - base) * fplane + (long long)y * fpitch + x);
    }
};

__device__ __forceinline__ float face_scale(int c, int n) { return (c == 0 || c == n - 1) ? 1.0f : 0.5f; }

__device__ __forceinline__ float d_dx(const FView& f, int x, int y, int z)
{
    return __fmul_rn(face_scale(x, f.w), __fsub_rn(f.at(min(x + 1, f.w - 1), y, z), f.at(max(x - 1, 0), y, z)));
}
__device__ __forceinline__ float d_dy(const FView& f, int x, int y, int z)
{
    return __fmul_rn(face_scale(y, f.h), __fsub_rn(f.at(x, min(y + 1, f.h - 1), z), f.at(x, max(y - 1, 0), z)));
}
__device__ __forceinline__ float d_dz(const FView& f, int x, int y, int z)
{
    return __fmul_rn(face_scale(z, f.l), __fsub_rn(f.at(x, y, min(z + 1, f.l - 1)), f.at(x, y, max(z - 1, 0))));
}

struct Hess { float xx, xy, xz, yy, yz, zz; };

__device__ __forceinline__ Hess hessian_at(const FView& f, int x, int y, int z, float sigma2)
{
    const int xl = max(x - 1, 0), xh = min(x + 1, f.w - 1);
    const int yl = max(y - 1, 0), yh = min(y + 1, f.h - 1);
    const int zl = max(z - 1, 0), zh = min(z + 1, f.l - 1);
    const float sx = face_scale(x, f.w), sy = face_scale(y, f.h), sz = face_scale(z, f.l);
    Hess H;
    H.xx = __fmul_rn(__fmul_rn(sx, __fsub_rn(d_dx(f, xh, y, z), d_dx(f, xl, y, z))), sigma2);
    H.xy = __fmul_rn(__fmul_rn(sy, __fsub_rn(d_dx(f, x, yh, z), d_dx(f, x, yl, z))), sigma2);
    H.xz = __fmul_rn(__fmul_rn(sz, __fsub_rn(d_dx(f, x, y, zh), d_dx(f, x, y, zl))), sigma2);
    H.yy = __fmul_rn(__fmul_rn(sy, __fsub_rn(d_dy(f, x, yh, z), d_dy(f, x, yl, z))), sigma2);
    H.yz = __fmul_rn(__fmul_rn(sz, __fsub_rn(d_dy(f, x, y, zh), d_dy(f, x, y, zl))), sigma2);
    H.zz = __fmul_rn(__fmul_rn(sz, __fsub_rn(d_dz(f, x, y, zh), d_dz(f, x, y, zl))), sigma2);
    return H;
}

// ---------------------------------------------------------------------------
// K3: Hessian -> eigen -> vesselness -> running max over scales.
// Outputs are dense over the slab's own planes [z_begin, z_begin + nz).
// first_scale: store unconditionally (frangi.cpp:234-252); otherwise overwrite
// only on a strictly greater response (frangi.cpp:254-271).
// minmax[0] = bits of min J (taken on the first scale),
// minmax[1] = bits of max J (taken on the last scale).  J >= 0, so the int
// order of the bit patterns is the float order.
// ---------------------------------------------------------------------------
struct VoxelParams {
    FView f;
    float* J;
    uint8_t* Vx;
    uint8_t* Vy;
    uint8_t* Vz;
    uint8_t* scale_idx;   // nullable
    float* dir;           // nullable, 3 planar volumes of `voxels` floats
    long long voxels;     // own voxels
    int z_begin, nz;
    int scale;            // index of this scale
    int first_scale, last_scale;
    int* minmax;
    FrangiConsts k;
};

__global__ void __launch_bounds__(128)
hessian_eigen_kernel(const __grid_constant__ VoxelParams p)
{
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int zl = blockIdx.z;
    float jval = 0.0f;
    const bool active = x < p.f.w;
    if (active) {
        const int z = p.z_begin + zl;
        const Hess H = hessian_at(p.f, x, y, z, p.k.sigma2);
        Eig3 e;
        eig_sym3(H.xx, H.xy, H.xz, H.yy, H.yz, H.zz, e);
        const float v = vesselness(e, p.k);
        const long long i = ((long long)zl * p.f.h + y) * p.f.w + x;
        bool write = p.first_scale;
        float jold = 0.0f;
        if (!write) { jold = p.J[i]; write = v > jold; }
        if (write) {
            p.J[i] = v;
            p.Vx[i] = dir_code(e.vx);
            p.Vy[i] = dir_code(e.vy);
            p.Vz[i] = dir_code(e.vz);
            if (p.scale_idx) p.scale_idx[i] = (uint8_t)p.scale;
            if (p.dir) {
                p.dir[i] = e.vx;
                p.dir[p.voxels + i] = e.vy;
                p.dir[2 * p.voxels + i] = e.vz;
            }
            jval = v;
        } else {
            jval = jold;
        }
    }
    // warp-shuffle reductions of min (first scale) and max (last scale)
    if (p.first_scale) {
        float m = active ? jval : 3.4e38f;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, s));
        if ((threadIdx.x & 31) == 0) atomicMin(p.minmax + 0, __float_as_int(m));
    }
    if (p.last_scale) {
        float m = active ? jval : 0.0f;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
        if ((threadIdx.x & 31) == 0) atomicMax(p.minmax + 1, __float_as_int(m));
    }
}

// Debug / stage kernel: dump the six second-difference volumes (hessian3d parity).
struct HessDumpParams {
    FView f;
    float* D[6];   // Dzz, Dyy, Dyz, Dxx, Dxy, Dxz (reference argument order), dense
    int z_begin;
    float sigma2;
};

__global__ void __launch_bounds__(128) hessian_dump_kernel(const __grid_constant__ HessDumpParams p)
{
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= p.f.w) return;
    const int y = blockIdx.y, zl = blockIdx.z;
    const Hess H = hessian_at(p.f, x, y, p.z_begin + zl, p.sigma2);
    const long long i = ((long long)zl * p.f.h + y) * p.f.w + x;
    p.D[0][i] = H.zz; p.D[1][i] = H.yy; p.D[2][i] = H.yz;
    p.D[3][i] = H.xx; p.D[4][i] = H.xy; p.D[5][i] = H.xz;
}

// Stage kernel: eigen + vesselness on caller-supplied Hessians.
__global__ void __launch_bounds__(128)
vesselness_stage_kernel(const float* __restrict__ Dxx, const float* __restrict__ Dxy,
                        const float* __restrict__ Dxz, const float* __restrict__ Dyy,
                        const float* __restrict__ Dyz, const float* __restrict__ Dzz, long long n,
                        FrangiConsts k, float* __restrict__ v_out, float* __restrict__ dir_out,
                        float* __restrict__ lambda_out)
{
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    Eig3 e;
    eig_sym3(Dxx[i], Dxy[i], Dxz[i], Dyy[i], Dyz[i], Dzz[i], e);
    v_out[i] = vesselness(e, k);
    if (dir_out) { dir_out[i] = e.vx; dir_out[n + i] = e.vy; dir_out[2 * n + i] = e.vz; }
    if (lambda_out) { lambda_out[3 * i] = e.l1; lambda_out[3 * i + 1] = e.l2; lambda_out[3 * i + 2] = e.l3; }
}

// ---------------------------------------------------------------------------
// K4: J -> J8 (Advantra_plugin.cpp:2499-2512, round() from :120-123).
// minmax holds the float bit patterns of Jmin / Jmax (after the all-reduce).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
j_to_j8_kernel(const float* __restrict__ J, uint8_t* __restrict__ J8, long long n, const int* __restrict__ minmax)
{
    const float lo = __int_as_float(minmax[0]);
    const float hi = __int_as_float(minmax[1]);
    const float range = __fsub_rn(hi, lo);
    const bool flat = fabsf(range) <= 1.175494351e-38f;  // FLT_MIN
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        int v = 0;
        if (!flat) {
            const float qf = __fmul_rn(__fdiv_rn(__fsub_rn(J[i], lo), range), 255.0f);
            const double r = (double)qf;
            v = (int)((r > 0.0) ? floor(r + 0.5) : ceil(r - 0.5));
            v = min(max(v, 0), 255);
        }
        J8[i] = (uint8_t)v;
    }
}

}  // namespace frangi
