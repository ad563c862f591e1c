// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" surface (ours) over the UNMODIFIED reference sources, which
// are compiled where they lie under /root/reference/pnr-vaa3d (never copied into
// this repository -- their licence forbids redistribution).  The resulting
// oracle/_ref/libpnr_ref.so is used by tests/ to pin oracle/frangi_oracle.c,
// to generate tests/golden/, as the downstream seed-extraction consumer, and by
// bench.py as the CPU baseline ("kind": "reference").
//
// Reference interfaces bound here:
//   Frangi::imgaussian(I,w,h,l,sig,zdist,F)      frangi.h:42,  frangi.cpp:647
//   Frangi::hessian3d(...)                       frangi.h:35,  frangi.cpp:291
//   Frangi::eigen_decomposition_static(A,V,d)    frangi.h:54,  frangi.cpp:1230
//   Frangi::frangi3d(...)                        frangi.h:33,  frangi.cpp:152
//   SeedExtractor::extractSeeds(...)             seed.h:101,   seed.cpp:556
#include <iostream>
#include <vector>

#include "frangi.h"
#include "seed.h"

namespace {
// The reference prints progress to std::cout from every stage; park the stream
// in a failed state for the duration of a call so timing is not I/O bound and
// stdout stays clean for the JSON line of bench.py.
struct QuietCout {
    std::ios_base::iostate saved;
    QuietCout() : saved(std::cout.rdstate()) { std::cout.setstate(std::ios_base::failbit); }
    ~QuietCout() { std::cout.clear(saved); }
};
}  // namespace

extern "C" {

__attribute__((visibility("default")))
void ref_imgaussian(unsigned char* I, int w, int h, int l, float sig, float zdist, float* F)
{
    QuietCout q;
    Frangi::imgaussian(I, w, h, l, sig, zdist, F);
}

__attribute__((visibility("default")))
void ref_hessian3d(unsigned char* I, int w, int h, int l, float sig, float zdist,
                   float* Dzz, float* Dyy, float* Dyz, float* Dxx, float* Dxy, float* Dxz)
{
    QuietCout q;
    std::vector<float> s(1, sig);
    Frangi f(s, zdist, .5f, .5f, 500.f, .5f, 15.f);
    f.hessian3d(I, w, h, l, sig, zdist, Dzz, Dyy, Dyz, Dxx, Dxy, Dxz);
}

__attribute__((visibility("default")))
void ref_eigen3(const double* A, double* V, double* d)
{
    double a[3][3], v[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) a[r][c] = A[3 * r + c];
    Frangi::eigen_decomposition_static(a, v, d);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) V[3 * r + c] = v[r][c];
}

__attribute__((visibility("default")))
void ref_frangi3d(unsigned char* I, int w, int h, int l, const float* sigmas, int nsig,
                  float zdist, float alpha, float beta, float C, int blackwhite,
                  float* J, float* Jmin, float* Jmax,
                  unsigned char* Vx, unsigned char* Vy, unsigned char* Vz)
{
    QuietCout q;
    std::vector<float> s(sigmas, sigmas + nsig);
    Frangi f(s, zdist, alpha, beta, C, .5f, 15.f);
    f.blackwhite = blackwhite != 0;
    float lo, hi;
    f.frangi3d(I, w, h, l, J, lo, hi, Vx, Vy, Vz);
    *Jmin = lo;
    *Jmax = hi;
}

// Seeds are returned as rows of 6 floats (x, y, z, vx, vy, vz).  Returns the
// number of seeds found; at most `cap` rows are written.
__attribute__((visibility("default")))
long ref_extract_seeds(double tolerance, unsigned char* J8, int w, int h, int l,
                       unsigned char* Vx, unsigned char* Vy, unsigned char* Vz,
                       float* out, long cap)
{
    QuietCout q;
    std::vector<seed> seeds;
    SeedExtractor::extractSeeds(tolerance, J8, w, h, l, Vx, Vy, Vz, seeds);
    long n = (long)seeds.size();
    for (long i = 0; i < n && i < cap; ++i) {
        out[6 * i + 0] = seeds[i].x;  out[6 * i + 1] = seeds[i].y;  out[6 * i + 2] = seeds[i].z;
        out[6 * i + 3] = seeds[i].vx; out[6 * i + 4] = seeds[i].vy; out[6 * i + 5] = seeds[i].vz;
    }
    return n;
}

}  // extern "C"
