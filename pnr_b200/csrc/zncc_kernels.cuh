// zncc_kernels.cuh -- SURVEY.md 8f row f3, second half: the per-seed correlation score of the plugin's seed filter.
//
// After extractSeeds the plugin scores every seed with Tracker::znccBBB on the RAW image and drops those below
// znccth, then sorts by score (Advantra_plugin.cpp:2561-2586).  znccBBB (tracker.cpp:1891-1964) samples, for every
// sigma, a box of the image around the seed -- offsets (v, u, w) along the seed direction and two orthogonal
// directions, trilinear interpolation (tracker.cpp:2138-2215) -- and returns the largest zero-mean normalised
// cross-correlation of the samples with a Gaussian cross-section template (model2_*, tracker.cpp:170-232): 845 / 5625
// / 5625 samples per seed at sigma = 2, 4, 6.  Seeds are independent: one THREAD per seed, walking its samples in the
// reference's order with the reference's float / double mix operation by operation (no fused multiply-add, the
// squares of pow(x, 2) in double), so scores -- and therefore the filter decisions and the sort order downstream --
// are bit-identical.  The sample values are recomputed in the second pass instead of being stored (the reference keeps
// them in model2_img).  Cold code (one call per image); ~1e5 seeds x 1.2e4 samples take milliseconds.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace frangi {

struct ZnccParams {
    const uint8_t* img;       // [l][h][w]
    int w, h, l;
    const float4* samp;       // per sample: v, u, w offsets and (template weight - template mean)
    const int* first;         // first[s] .. first[s+1]: the samples of sigma s
    const float* corrc;       // per sigma: sum of squared centred template weights (tracker.cpp:1951)
    const float* sig;         // per sigma: the value reported for the best one
    int nsig;
    const float* seeds;       // n x 6: x, y, z, vx, vy, vz
    long long n;
    float* corr_out;
    float* sig_out;
    float xmax, ymax, zmax;   // (float)(dim - 1.001) (tracker.cpp:2140,2145,2178)
};

__device__ __forceinline__ float zn_clamp(float x, float lo, float hi) { const float c = x < lo ? lo : x; return c > hi ? hi : c; }

// Tracker::interp, 3-D branch (tracker.cpp:2178-2213): every product and sum separately rounded in float
__device__ __forceinline__ float zn_interp(const ZnccParams& p, float x, float y, float z)
{
    const float xc = zn_clamp(x, 0.0f, p.xmax), yc = zn_clamp(y, 0.0f, p.ymax), zc = zn_clamp(z, 0.0f, p.zmax);
    const int x1 = (int)xc, y1 = (int)yc, z1 = (int)zc;
    const float xf = __fsub_rn(xc, (float)x1), yf = __fsub_rn(yc, (float)y1), zf = __fsub_rn(zc, (float)z1);
    const float ax = __fsub_rn(1.0f, xf), ay = __fsub_rn(1.0f, yf), az = __fsub_rn(1.0f, zf);
    const uint8_t* q = p.img + ((long long)z1 * p.h + y1) * p.w + x1;
    const long long plane = (long long)p.w * p.h;
    const float i111 = (float)q[0], i112 = (float)q[1], i121 = (float)q[p.w], i122 = (float)q[p.w + 1];
    const float i211 = (float)q[plane], i212 = (float)q[plane + 1], i221 = (float)q[plane + p.w], i222 = (float)q[plane + p.w + 1];
    const float r11 = __fadd_rn(__fmul_rn(ax, i111), __fmul_rn(xf, i112));
    const float r12 = __fadd_rn(__fmul_rn(ax, i121), __fmul_rn(xf, i122));
    const float r21 = __fadd_rn(__fmul_rn(ax, i211), __fmul_rn(xf, i212));
    const float r22 = __fadd_rn(__fmul_rn(ax, i221), __fmul_rn(xf, i222));
    const float lo = __fadd_rn(__fmul_rn(ay, r11), __fmul_rn(yf, r12));
    const float hi = __fadd_rn(__fmul_rn(ay, r21), __fmul_rn(yf, r22));
    return __fadd_rn(__fmul_rn(az, lo), __fmul_rn(zf, hi));
}

__global__ void __launch_bounds__(128)
seed_zncc_kernel(const __grid_constant__ ZnccParams p)
{
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= p.n) return;
    const float* s = p.seeds + 6 * i;
    const float sx = s[0], sy = s[1], sz = s[2], vx = s[3], vy = s[4], vz = s[5];
    // the orthogonal frame (tracker.cpp:1893-1917): nrm = sqrt(pow(vx,2) + pow(vy,2)) in double, stored as float
    const float nrm = (float)sqrt((double)vx * (double)vx + (double)vy * (double)vy);
    float ux, uy, uz;
    if ((double)nrm > 0.0001) {
        const float sg = vy < 0 ? -1.0f : 1.0f;
        ux = __fmul_rn(sg, __fdiv_rn(vy, nrm));
        uy = __fmul_rn(-sg, __fdiv_rn(vx, nrm));
        uz = 0.0f;
    } else {
        ux = 1.0f; uy = 0.0f; uz = 0.0f;
    }
    const float wx = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy));
    const float wy = __fadd_rn(__fmul_rn(-ux, vz), __fmul_rn(uz, vx));
    const float wz = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
    const float nvx = -vx, nvy = -vy, nvz = -vz;
    auto sample = [&](const float4 o) {
        const float x = __fadd_rn(__fadd_rn(__fadd_rn(sx, __fmul_rn(o.x, nvx)), __fmul_rn(o.y, ux)), __fmul_rn(o.z, wx));
        const float y = __fadd_rn(__fadd_rn(__fadd_rn(sy, __fmul_rn(o.x, nvy)), __fmul_rn(o.y, uy)), __fmul_rn(o.z, wy));
        const float z = __fadd_rn(__fadd_rn(__fadd_rn(sz, __fmul_rn(o.x, nvz)), __fmul_rn(o.y, uz)), __fmul_rn(o.z, wz));
        return zn_interp(p, x, y, z);
    };
    float best = -3.402823466e+38f, best_sig = 0.0f;
    for (int k = 0; k < p.nsig; ++k) {
        const int a = p.first[k], b = p.first[k + 1];
        float ag = 0.0f;
        for (int j = a; j < b; ++j) ag = __fadd_rn(ag, sample(__ldg(p.samp + j)));
        ag = __fdiv_rn(ag, (float)(b - a));
        float corra = 0.0f, corrb = 0.0f;
        for (int j = a; j < b; ++j) {
            const float4 o = __ldg(p.samp + j);
            const float d = __fsub_rn(sample(o), ag);
            corra = __fadd_rn(corra, __fmul_rn(d, o.w));
            corrb = (float)((double)corrb + (double)d * (double)d);       // corrb += pow(d, 2): a double square
        }
        const float den = __fmul_rn(corrb, p.corrc[k]);
        const float c = den > 1.175494351e-38f ? __fdiv_rn(corra, __fsqrt_rn(den)) : 0.0f;
        if (c > best) { best = c; best_sig = p.sig[k]; }
    }
    p.corr_out[i] = best;
    if (p.sig_out) p.sig_out[i] = best_sig;
}

}  // namespace frangi
