// frangi_voxel_math.cuh -- the per-voxel stage of the Frangi filter:
// symmetric 3x3 eigen-decomposition, |lambda| ordering, vesselness, direction.
//
// Replaces, semantically, eigen_decomposition + tred2 + tql2 (frangi.cpp:1269-1493)
// and the per-voxel body of frangi3d (frangi.cpp:190-250).  The reference runs an
// iterative double-precision QL; this is a closed-form float32 stage written once
// over a lane-vector type V: V = float (one voxel: the face shell, ragged quads)
// or V = float2 (two voxels in packed float32x2 registers: the interior kernel).
// Blackwell's FMUL2 / FADD2 / FFMA2 take ONE issue slot for two lanes, negation
// and |.| being operand modifiers; the stage is issue-bound, so the packed form is
// what the hot path runs.  MUFU, compares and selects stay per lane.
//
// Algorithm (non-diagonal input):
//  1. the eigenvalue at the isolated end of the spectrum from the trigonometric
//     form, lam = q + sgn * p * g(|r|), with g(r) = 2 cos(acos(r)/3) as a degree-6
//     polynomial on [0,1] (|err| < 4e-7, float rounding level);
//  2. its eigenvector i = the largest column of adj(A - lam I) (every column is a
//     multiple of it; the largest diagonal cofactor picks the best conditioned);
//  3. a branch-free orthonormal complement (u, w) of i (Duff et al. 2017) and the
//     2x2 projection M of A on it; the other two eigenvalues are
//     mean -+ sqrt(hdiff^2 + m01^2) -- a sum of squares, not a cancelling
//     difference, so a close pair (the tube case l2 ~ l3) is split to float
//     accuracy; trace invariance gives m11 = tr - lam - m00, so A w is never formed;
//  4. |lambda| order with the reference's tie rules (frangi.cpp:1284-1304) carried
//     by predicates; the eigenvector of l1 is i, or the null vector of (M - l1 I)
//     mapped back through (u, w).
// Exactly diagonal inputs (flat background, axis-aligned synthetic data) take
// eig_diag, which reproduces the reference's conventions for them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef EIG_ORDER_C2
#define EIG_ORDER_C2 0      // 1: keep the (unreachable) middle-value case of the |lambda| ordering, for A/B timing
#endif
#ifndef EIG_PACK_ADDS
#define EIG_PACK_ADDS 1     // 0: the three per-lane adds of the stage as scalar instructions (round-1 form), for A/B timing
#endif

namespace frangi {

struct FrangiConsts {
    float inv_2a2;    // 1 / (2*alpha*alpha)   (float products as in frangi.cpp:215-217)
    float inv_2b2;
    float inv_2c2;
    float sigma2;     // sigma*sigma, float    (frangi.cpp:319)
    int blackwhite;
};

// ---- single-instruction MUFU forms (denormals flush; 1-2 ulp) -------------------
__device__ __forceinline__ float mufu_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- lane vectors --------------------------------------------------------------
template <class V> struct Lanes;
template <> struct Lanes<float> {
    static constexpr int N = 1;
    static __device__ __forceinline__ float get(const float& v, int) { return v; }
    static __device__ __forceinline__ void set(float& v, int, float x) { v = x; }
    static __device__ __forceinline__ float bc(float s) { return s; }
};
template <> struct Lanes<float2> {
    static constexpr int N = 2;
    static __device__ __forceinline__ float get(const float2& v, int k) { return k ? v.y : v.x; }
    static __device__ __forceinline__ void set(float2& v, int k, float x) { if (k) v.y = x; else v.x = x; }
    static __device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
};

__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float vabs(float a) { return fabsf(a); }
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 vabs(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return __fadd2_rn(a, vneg(b)); }
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// a*b - c*d
template <class V> __device__ __forceinline__ V vdiff(V a, V b, V c, V d) { return vfma(a, b, vneg(vmul(c, d))); }

template <class V> struct EigT {
    V l1, l2, l3;   // |l1| <= |l2| <= |l3| with the reference's tie rules
    V vx, vy, vz;   // unit eigenvector of l1
    bool zero[2];   // BRIGHT_LATER only: the sign gate of frangi.cpp:225-228 closes (response is 0)
};
typedef EigT<float> Eig3;
typedef EigT<float2> Eig3x2;

__device__ __forceinline__ void swapf(float& a, float& b) { float t = a; a = b; b = t; }

// Exactly diagonal input: the reference's QL leaves the values and the identity
// untouched, then selection-sorts ascending (first minimum wins ties), then
// orders by |lambda| (frangi.cpp:1473-1492, 1284-1304).
__device__ __noinline__ Eig3 eig_diag(float a00, float a11, float a22)
{
    Eig3 out;
    float e0 = a00, e1 = a11, e2 = a22;
    int i0 = 0, i1 = 1, i2 = 2;
    int k = 0; float pv = e0;
    if (e1 < pv) { k = 1; pv = e1; }
    if (e2 < pv) { k = 2; pv = e2; }
    if (k == 1) { swapf(e0, e1); i0 = 1; i1 = 0; }
    else if (k == 2) { swapf(e0, e2); i0 = 2; i2 = 0; }
    if (e2 < e1) { swapf(e1, e2); int t = i1; i1 = i2; i2 = t; }
    float m0 = fabsf(e0), m1 = fabsf(e1), m2 = fabsf(e2);
    if (m0 >= m1 && m0 > m2) { swapf(e0, e2); swapf(m0, m2); int t = i0; i0 = i2; i2 = t; }
    else if (m1 >= m0 && m1 > m2) { swapf(e1, e2); swapf(m1, m2); int t = i1; i1 = i2; i2 = t; }
    if (m0 > m1) { swapf(e0, e1); int t = i0; i0 = i1; i1 = t; }
    out.l1 = e0; out.l2 = e1; out.l3 = e2;
    out.vx = i0 == 0 ? 1.0f : 0.0f; out.vy = i0 == 1 ? 1.0f : 0.0f; out.vz = i0 == 2 ? 1.0f : 0.0f;
    return out;
}

// BRIGHT_LATER: a later scale of a bright-ridge run (blackwhite == false, frangi.cpp:254-271).
// There a voxel can only be overwritten when its response is positive, i.e. when the two
// eigenvalues of largest magnitude are <= 0; for the ascending triple e0 <= e1 <= e2 that is
//   e1 <= 0  and  (e2 <= 0  or  (|e2| <= |e1| and |e2| < |e0|))            (tie rules of :1289-1304)
// and then the |lambda| order is simply (l1, l2, l3) = (e2, e1, e0) with the direction the
// eigenvector of the LARGEST eigenvalue.  Voxels that fail the test are flagged `zero` (their
// response is 0 and never wins), so the general ordering logic is not needed.
template <class V, bool BRIGHT_LATER = false>
__device__ __forceinline__ void eig_sym3(V a00, V a01, V a02, V a11, V a12, V a22, EigT<V>& out)
{
    typedef Lanes<V> L;
    const V off = vfma(a01, a01, vfma(a02, a02, vmul(a12, a12)));
    const V tr = vadd(vadd(a00, a11), a22);
    const V q = vmul(tr, L::bc(1.0f / 3.0f));
    const V b00 = vsub(a00, q), b11 = vsub(a11, q), b22 = vsub(a22, q);
    const V p26 = vmul(vfma(b00, b00, vfma(b11, b11, vfma(b22, b22, vadd(off, off)))), L::bc(1.0f / 6.0f));
    V ip;
#pragma unroll
    for (int k = 0; k < L::N; ++k) L::set(ip, k, mufu_rsqrt(L::get(p26, k)));
    const V p = vmul(p26, ip);
    const V det = vfma(b00, vdiff(b11, b22, a12, a12),
                       vfma(vneg(a01), vdiff(a01, b22, a12, a02), vmul(a02, vdiff(a01, a12, b11, a02))));
    const V hd = vmul(det, vmul(vmul(ip, ip), vmul(ip, L::bc(0.5f))));
    V ar, sp;
    bool top[L::N];
#pragma unroll
    for (int k = 0; k < L::N; ++k) {
        const float h = L::get(hd, k), pk = L::get(p, k);
        L::set(ar, k, fminf(fabsf(h), 1.0f));
        top[k] = !signbit(h);                  // the largest eigenvalue is the isolated one
        L::set(sp, k, copysignf(pk, h));
    }
    V g = vfma(ar, L::bc(-0.00202751093f), L::bc(0.0100089306f));
    g = vfma(ar, g, L::bc(-0.0248566089f));
    g = vfma(ar, g, L::bc(0.0474178627f));
    g = vfma(ar, g, L::bc(-0.0959053785f));
    g = vfma(ar, g, L::bc(0.333311245f));
    g = vfma(ar, g, L::bc(1.73205118f));
    const V lam = vfma(sp, g, q);
    // adj(A - lam I): every column is a multiple of the eigenvector
    const V r00 = vsub(a00, lam), r11 = vsub(a11, lam), r22 = vsub(a22, lam);
    const V c00 = vdiff(r11, r22, a12, a12);
    const V c11 = vdiff(r00, r22, a02, a02);
    const V c22 = vdiff(r00, r11, a01, a01);
    const V c01 = vdiff(a02, a12, a01, r22);
    const V c02 = vdiff(a01, a12, a02, r11);
    const V c12 = vdiff(a01, a02, a12, r00);
    V nx, ny, nz;
#pragma unroll
    for (int k = 0; k < L::N; ++k) {
        const float d0 = fabsf(L::get(c00, k)), d1 = fabsf(L::get(c11, k)), d2 = fabsf(L::get(c22, k));
        const bool s1 = d1 > d0;
        const bool s2 = d2 > fmaxf(d0, d1);
        L::set(nx, k, s2 ? L::get(c02, k) : (s1 ? L::get(c01, k) : L::get(c00, k)));
        L::set(ny, k, s2 ? L::get(c12, k) : (s1 ? L::get(c11, k) : L::get(c01, k)));
        L::set(nz, k, s2 ? L::get(c22, k) : (s1 ? L::get(c12, k) : L::get(c02, k)));
    }
    const V nn = vfma(nx, nx, vfma(ny, ny, vmul(nz, nz)));
    V inn;
#pragma unroll
    for (int k = 0; k < L::N; ++k) L::set(inn, k, mufu_rsqrt(L::get(nn, k)));
    const V ix = vmul(nx, inn), iy = vmul(ny, inn), iz = vmul(nz, inn);
    // branch-free orthonormal complement of i (Duff, Burgess, Christensen, Hery, Kensler, Liani, Villemin 2017):
    //   s = sign(iz), a = -1/(s+iz), b = ix iy a, u = (1 + s ix^2 a, s b, -s ix), w = (b, s + iy^2 a, -iy)
    V s, a;
#if EIG_PACK_ADDS
#pragma unroll
    for (int k = 0; k < L::N; ++k) L::set(s, k, copysignf(1.0f, L::get(iz, k)));
    const V nsz = vsub(vneg(s), iz);                     // -(s + iz), one packed add
#pragma unroll
    for (int k = 0; k < L::N; ++k) L::set(a, k, mufu_rcp(L::get(nsz, k)));
#else
#pragma unroll
    for (int k = 0; k < L::N; ++k) {
        const float sk = copysignf(1.0f, L::get(iz, k));
        L::set(s, k, sk);
        L::set(a, k, -mufu_rcp(sk + L::get(iz, k)));
    }
#endif
    const V sx = vmul(s, ix), xa = vmul(ix, a), ya = vmul(iy, a);
    const V b = vmul(xa, iy);
    const V ux = vfma(sx, xa, L::bc(1.0f)), uy = vmul(s, b), uz = vneg(sx);
    const V wx = b, wy = vfma(ya, iy, s), wz = vneg(iy);
    // 2x2 projection; trace invariance supplies m11 (lam stands for i'Ai)
    const V aux = vfma(a00, ux, vfma(a01, uy, vmul(a02, uz)));
    const V auy = vfma(a01, ux, vfma(a11, uy, vmul(a12, uz)));
    const V auz = vfma(a02, ux, vfma(a12, uy, vmul(a22, uz)));
    const V m00 = vfma(ux, aux, vfma(uy, auy, vmul(uz, auz)));
    const V m01 = vfma(wx, aux, vfma(wy, auy, vmul(wz, auz)));
    const V sum = vsub(tr, lam);
    const V m11 = vsub(sum, m00);
    const V mean = vmul(sum, L::bc(0.5f));
    const V hdiff = vsub(m00, mean);
    const V d2 = vfma(hdiff, hdiff, vmul(m01, m01));
    V disc;
#pragma unroll
    for (int k = 0; k < L::N; ++k) L::set(disc, k, mufu_sqrt(L::get(d2, k)));
    const V la = vsub(mean, disc), lb = vadd(mean, disc);
    bool sel_i[L::N];
    if (BRIGHT_LATER) {
#pragma unroll
        for (int k = 0; k < L::N; ++k) {
            const float fa = L::get(la, k), fb = L::get(lb, k), fi = L::get(lam, k);
            const float e0 = top[k] ? fa : fi, e1 = top[k] ? fb : fa, e2 = top[k] ? fi : fb;
            const bool pass = e1 <= 0.0f && (e2 <= 0.0f || (fabsf(e2) <= fabsf(e1) && fabsf(e2) < fabsf(e0)));
            out.zero[k] = !pass;
            L::set(out.l1, k, e2); L::set(out.l2, k, e1); L::set(out.l3, k, e0);
            sel_i[k] = top[k];
        }
    } else {
    // |lambda| order with the reference's rules, applied to the ascending triple
    // (la, lb, lam) if top else (lam, la, lb); sel_i: column 0 is the isolated eigenvector
#pragma unroll
    for (int k = 0; k < L::N; ++k) {
        const float fa = L::get(la, k), fb = L::get(lb, k), fi = L::get(lam, k);
        const float e0 = top[k] ? fa : fi, e1 = top[k] ? fb : fa, e2 = top[k] ? fi : fb;
#if EIG_ORDER_C2
        const float m0 = fabsf(e0), m1 = fabsf(e1), m2 = fabsf(e2);
        const bool c1 = m0 >= m1 && m0 > m2;            // e0 has the largest magnitude
        const bool c2 = !c1 && m1 >= m0 && m1 > m2;     // else e1 has
        const float x = c1 ? e2 : e0;
        const float y = c2 ? e2 : e1;
        const bool sw = fabsf(x) > fabsf(y);
        L::set(out.l3, k, c1 ? e0 : (c2 ? e1 : e2));
        L::set(out.l1, k, sw ? y : x);
        L::set(out.l2, k, sw ? x : y);
        // where the isolated eigenvalue went: it is e2 if top (x when c1, y when c2), else e0 (x unless c1)
        sel_i[k] = top[k] ? ((c1 && !sw) || (c2 && sw)) : (!c1 && !sw);
#else
        // For an ascending triple the middle value never has the strictly largest magnitude (frangi.cpp:1289-1296 can
        // only pick e1 when |e1| >= |e0| and |e1| > |e2|, i.e. e0 = e1 < 0, and then the first rule has already taken
        // e0), and |e0| > |e2| implies |e0| >= |e1|: the largest is e0 if |e0| > |e2|, else e2.
        const bool c1 = fabsf(e0) > fabsf(e2);
        const float x = c1 ? e2 : e0;
        const bool sw = fabsf(x) > fabsf(e1);
        L::set(out.l3, k, c1 ? e0 : e2);
        L::set(out.l1, k, sw ? e1 : x);
        L::set(out.l2, k, sw ? x : e1);
        sel_i[k] = !sw && (c1 == top[k]);    // the isolated eigenvalue is e2 if top (it is x when c1), else e0 (x unless c1)
#endif
    }
    }
    // null vector of (M - l1 I) in (u, w) coordinates, from the larger row
    const V f0 = vsub(m00, out.l1), f1 = vsub(m11, out.l1);
    V y0, y1;
#pragma unroll
    for (int k = 0; k < L::N; ++k) {
        const float g0 = L::get(f0, k), g1 = L::get(f1, k), mk = L::get(m01, k);
        const bool r0 = fabsf(g0) >= fabsf(g1);
#if EIG_PACK_ADDS
        L::set(y0, k, r0 ? -mk : g1);
#else
        L::set(y0, k, (r0 ? -mk : g1) + 1e-18f);
#endif
        L::set(y1, k, r0 ? g0 : -mk);
    }
    if (EIG_PACK_ADDS) y0 = vadd(y0, L::bc(1e-18f));     // the epsilon turns the all-zero case (la == lb) into (1, 0)
    const V nrm = vfma(y0, y0, vmul(y1, y1));
    V sn;
#pragma unroll
    for (int k = 0; k < L::N; ++k) L::set(sn, k, mufu_rsqrt(L::get(nrm, k)));
    const V xa0 = vmul(y0, sn), xa1 = vmul(y1, sn);
    const V px = vfma(xa0, ux, vmul(xa1, wx)), py = vfma(xa0, uy, vmul(xa1, wy)), pz = vfma(xa0, uz, vmul(xa1, wz));
#pragma unroll
    for (int k = 0; k < L::N; ++k) {
        L::set(out.vx, k, sel_i[k] ? L::get(ix, k) : L::get(px, k));
        L::set(out.vy, k, sel_i[k] ? L::get(iy, k) : L::get(py, k));
        L::set(out.vz, k, sel_i[k] ? L::get(iz, k) : L::get(pz, k));
    }
    // exactly diagonal inputs follow the reference's conventions (rare outside flat background)
    float offmin = L::get(off, 0);
#pragma unroll
    for (int k = 1; k < L::N; ++k) offmin = fminf(offmin, L::get(off, k));
    if (offmin == 0.0f)
#pragma unroll
    for (int k = 0; k < L::N; ++k)
        if (L::get(off, k) == 0.0f) {
            const Eig3 e = eig_diag(L::get(a00, k), L::get(a11, k), L::get(a22, k));
            L::set(out.l1, k, e.l1); L::set(out.l2, k, e.l2); L::set(out.l3, k, e.l3);
            L::set(out.vx, k, e.vx); L::set(out.vy, k, e.vy); L::set(out.vz, k, e.vz);
            if (BRIGHT_LATER) out.zero[k] = e.l2 > 0.0f || e.l3 > 0.0f;
        }
}

// 1 - exp(-x) for x >= 0 without cancellation (the reference evaluates it in
// double; in float32 the subtraction would lose everything for the S term,
// where x ~ 1e-4 with C = 500): alternating series below 1/4 (truncation
// < 2e-6 relative), 1 - ex2 above.
// RARE_BIG: the argument is almost always below 1/4 (the S term), so the ex2 form sits
// behind a branch instead of costing a MUFU per voxel.
template <class V, bool RARE_BIG>
__device__ __forceinline__ V one_minus_exp_neg(V x)
{
    typedef Lanes<V> L;
    V s = vfma(x, L::bc(-1.0f / 720.0f), L::bc(1.0f / 120.0f));
    s = vfma(x, s, L::bc(-1.0f / 24.0f));
    s = vfma(x, s, L::bc(1.0f / 6.0f));
    s = vfma(x, s, L::bc(-0.5f));
    s = vfma(x, s, L::bc(1.0f));
    s = vmul(x, s);
    if (RARE_BIG) {
#pragma unroll
        for (int k = 0; k < L::N; ++k)
            if (L::get(x, k) >= 0.25f) L::set(s, k, 1.0f - mufu_ex2(-1.4426950408889634f * L::get(x, k)));
        return s;
    }
    const V xe = vmul(x, L::bc(-1.4426950408889634f));
    V r;
#pragma unroll
    for (int k = 0; k < L::N; ++k)
        L::set(r, k, L::get(x, k) < 0.25f ? L::get(s, k) : 1.0f - mufu_ex2(L::get(xe, k)));
    return r;
}

// Frangi vesselness from |lambda|-sorted eigenvalues (frangi.cpp:206-231).
template <class V, bool BRIGHT_LATER = false>
__device__ __forceinline__ V vesselness(const EigT<V>& e, const FrangiConsts& c)
{
    typedef Lanes<V> L;
    const V a1 = vabs(e.l1), a2 = vabs(e.l2), a3 = vabs(e.l3);
    const V a23 = vmul(a2, a3);
    V i23;
#pragma unroll
    for (int k = 0; k < L::N; ++k) L::set(i23, k, mufu_rcp(L::get(a23, k)));
    const V a11 = vmul(a1, a1);
    const V ra = vmul(vmul(a2, a2), i23);                                    // |l2| / |l3| = l2^2 / |l2 l3|
    const V Ra2 = vmul(ra, ra);
    const V Rb2 = vmul(a11, i23);                                            // (|l1| / sqrt(|l2 l3|))^2
    const V S2 = vfma(a2, a2, vfma(a3, a3, a11));
    // 1 - exp(-Ra^2/2a^2) straight from ex2: its absolute error (1e-7) is multiplied by tS, which is
    // orders of magnitude inside the 1e-6 absolute tolerance, and where J is strong this term is O(1)
    const V xa = vmul(Ra2, L::bc(-1.4426950408889634f * c.inv_2a2));
    V tRa;
#pragma unroll
    for (int k = 0; k < L::N; ++k) L::set(tRa, k, EIG_PACK_ADDS ? mufu_ex2(L::get(xa, k)) : 1.0f - mufu_ex2(L::get(xa, k)));
    if (EIG_PACK_ADDS) tRa = vsub(L::bc(1.0f), tRa);
    const V xb = vmul(Rb2, L::bc(-1.4426950408889634f * c.inv_2b2));
    V tRb;
#pragma unroll
    for (int k = 0; k < L::N; ++k) L::set(tRb, k, mufu_ex2(L::get(xb, k)));
    const V tS = one_minus_exp_neg<V, true>(vmul(S2, L::bc(c.inv_2c2)));
    V v = vmul(vmul(tRa, tRb), tS);
#pragma unroll
    for (int k = 0; k < L::N; ++k) {
        const float l2 = L::get(e.l2, k), l3 = L::get(e.l3, k);
        float vk = L::get(v, k);
        vk = fmaxf(vk, 0.0f);                            // NaN (0/0 on a zero Hessian) -> 0, frangi.cpp:231
        const bool gate = BRIGHT_LATER ? e.zero[k]
                                       : (c.blackwhite ? fminf(l2, l3) < 0.0f : fmaxf(l2, l3) > 0.0f);   // frangi.cpp:221-228
        L::set(v, k, gate ? 0.0f : vk);
    }
    return v;
}

// round((c+1)/2*255) clamped to a byte (frangi.cpp:240-250); |c| <= 1 up to rounding, so
// c*127.5 + 128 lies in (0.49, 255.51) and its floor is a byte; NaN converts to 0
__device__ __forceinline__ uint32_t dir_code(float c)
{
    return (uint32_t)__float2int_rd(fmaf(c, 127.5f, 128.0f)) & 0xffu;
}

}  // namespace frangi
