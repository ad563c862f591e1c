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
//   Tracker::Tracker / znccBBB / trackPos / trackNeg   tracker.h:144,182,171,173  (ref_trace, config 5)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <iostream>
#include <vector>

#include "frangi.h"
#include "node.h"
#include "seed.h"
#include "tracker.h"

// The tracker reseeds the C generator with srand(time(NULL)) on every resampling step
// (tracker.cpp:655,808,1003,1098).  For a reproducible trace the library carries its own time():
// it is linked with -Bsymbolic-functions, so the reference objects inside this .so bind to it and
// every run (and both arms of the end-to-end comparison) draws the same numbers.
extern "C" time_t time(time_t* t) noexcept
{
    if (t) *t = (time_t)1;
    return (time_t)1;
}

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

// The 2-D path: Frangi::frangi2d (frangi.h:32, frangi.cpp:392) and Frangi::hessian2d (frangi.h:34, frangi.cpp:507)
__attribute__((visibility("default")))
void ref_frangi2d(unsigned char* I, int w, int h, const float* sigmas, int nsig, float beta_one, float beta_two, int blackwhite,
                  float* J, float* Jmin, float* Jmax, unsigned char* Vx, unsigned char* Vy, unsigned char* Vz)
{
    QuietCout q;
    std::vector<float> s(sigmas, sigmas + nsig);
    Frangi f(s, 1.0f, .5f, .5f, 500.f, beta_one, beta_two);
    f.blackwhite = blackwhite != 0;
    float lo, hi;
    f.frangi2d(I, w, h, 1, J, lo, hi, Vx, Vy, Vz);
    *Jmin = lo;
    *Jmax = hi;
}

__attribute__((visibility("default")))
void ref_hessian2d(unsigned char* I, int w, int h, float sig, float* Dyy, float* Dxy, float* Dxx)
{
    QuietCout q;
    std::vector<float> s(1, sig);
    Frangi f(s, 1.0f, .5f, .5f, 500.f, .5f, 15.f);
    f.hessian2d(I, w, h, sig, Dyy, Dxy, Dxx);
}

// The soma helpers: Frangi::imerode (frangi.h:47), Frangi::imdilate (frangi.h:49), in-place xy Frangi::imgaussian (frangi.h:43)
__attribute__((visibility("default")))
void ref_imerode(unsigned char* I, int w, int h, int l, float rad, unsigned char* E) { QuietCout q; Frangi::imerode(I, w, h, l, rad, E); }
__attribute__((visibility("default")))
void ref_imdilate(unsigned char* I, int w, int h, int l, float rad) { QuietCout q; Frangi::imdilate(I, w, h, l, rad); }
__attribute__((visibility("default")))
void ref_imgaussian_xy(unsigned char* I, int w, int h, int l, float sig) { QuietCout q; Frangi::imgaussian(I, w, h, l, sig); }

// the overloads no live code calls: z-scaled erosion (frangi.h:46), 2-D smoothing (frangi.h:44)
__attribute__((visibility("default")))
void ref_imerode_z(unsigned char* I, int w, int h, int l, float rad, float zdist, unsigned char* E) { QuietCout q; Frangi::imerode(I, w, h, l, rad, zdist, E); }
__attribute__((visibility("default")))
void ref_imgaussian2d(unsigned char* I, int w, int h, float sig, float* F) { QuietCout q; Frangi::imgaussian(I, w, h, sig, F); }

// host helpers of the class (frangi.h:28-31,51): direction tables, nearest table entry, z interpolation
__attribute__((visibility("default")))
void ref_unit_directions(int three_d, int ndir, float* out /* ndir x 3 */)
{
    QuietCout q;
    std::vector<float> s(1, 2.0f);
    Frangi f(s, 1.0f, .5f, .5f, 500.f, .5f, 15.f);
    std::vector<std::vector<float> > t;
    if (three_d) f.generate_3d_unit_directions((unsigned char)ndir, t);
    else f.generate_2d_unit_directions((unsigned char)ndir, t);
    for (size_t k = 0; k < t.size(); ++k)
        for (int c = 0; c < 3; ++c) out[3 * k + c] = t[k][c];
}
__attribute__((visibility("default")))
int ref_direction_idx(int three_d, float vx, float vy, float vz, const float* table, int ndir)
{
    QuietCout q;
    std::vector<float> s(1, 2.0f);
    Frangi f(s, 1.0f, .5f, .5f, 500.f, .5f, 15.f);
    std::vector<std::vector<float> > t(ndir, std::vector<float>(3));
    for (int k = 0; k < ndir; ++k)
        for (int c = 0; c < 3; ++c) t[k][c] = table[3 * k + c];
    return three_d ? f.get_direction_idx(vx, vy, vz, t) : f.get_direction_idx(vx, vy, t);
}
__attribute__((visibility("default")))
float ref_interpz(int x, int y, float z, float* img, int w, int h, int l)
{
    QuietCout q;
    std::vector<float> s(1, 2.0f);
    Frangi f(s, 1.0f, .5f, .5f, 500.f, .5f, 15.f);
    return f.interpz(x, y, z, img, w, h, l);
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


// Tracker::znccBBB for a list of seeds, with the tracker built as the plugin builds it (Advantra_plugin.cpp:2526,
// README parameters): rows of seeds6 = (x, y, z, vx, vy, vz); corr / sig = the score and the best-scoring sigma
__attribute__((visibility("default")))
void ref_seed_zncc(unsigned char* img, int w, int h, int l, const float* sigmas, int nsig, const float* seeds6, long n,
                   float* corr, float* sig)
{
    QuietCout q;
    FILE* saved_stdout = stdout;
    FILE* devnull = fopen("/dev/null", "w");
    if (devnull) stdout = devnull;
    std::vector<float> sigs(sigmas, sigmas + nsig);
    Tracker t(sigs, 2, 20, 200, 3.0f, l == 1, 0.3f, 20.0f, 0.8f, 2.0f, 4);
    for (long i = 0; i < n; ++i) {
        const float* s = seeds6 + 6 * i;
        float sg = 0;
        corr[i] = t.znccBBB(s[0], s[1], s[2], s[3], s[4], s[5], img, w, h, l, sg);
        sig[i] = sg;
    }
    stdout = saved_stdout;
    if (devnull) fclose(devnull);
}

// Everything the plugin does downstream of the Frangi filter up to the raw node list, restated from the
// one call site (Advantra_plugin.cpp) around the UNMODIFIED reference classes:
//   :2416-2419  node list with the dummy node 0            :2484  smap = 0 (somaradius == 0)
//   :2525-2526  SeedExtractor / Tracker construction        :2549  extractSeeds
//   :2560-2573  soma / correlation filter (Tracker::znccBBB, seeds below znccth dropped, walking backwards)
//   :2577-2586  sort by correlation, highest first (CompareSeedCorr :347-352)
//   :2602-2650  neighbour offsets: vol == 1 only here (ioff[i] = 0, the README usage `... 4 1`)
//   :2657-2719  trace loop: trackPos + trackNeg from every seed whose voxel holds < nodepervol nodes
// Plugin constants Kc = 20, neff_ratio = 0.8 (:63-64), MAX_TRACE_COUNT = 5000 (:72).
// Inputs are the image and the Frangi outputs (J8, Vx, Vy, Vz) of EITHER arm; the post-processing of
// reconstruct() is not run (it writes no final SWC with default settings), the comparison is
// on the node list n0.  Outputs: seeds_out rows (x, y, z, vx, vy, vz, score, corr) after filter + sort;
// nodes_out rows (x, y, z, vx, vy, vz, corr, sig, type, number of neighbours); nbr_out = the neighbour
// indices of all nodes, concatenated.  counts = {seeds extracted, seeds kept, nodes, neighbour entries, traces}.
__attribute__((visibility("default")))
int ref_trace(unsigned char* img, int w, int h, int l, unsigned char* J8, unsigned char* Vx, unsigned char* Vy,
              unsigned char* Vz, const float* sigmas, int nsig, float tolerance, float znccth, float kappa, int step,
              int ni, int np, float zdist, int nodepervol, int max_traces, float* seeds_out, long seeds_cap,
              float* nodes_out, long nodes_cap, int* nbr_out, long nbr_cap, long* counts)
{
    QuietCout q;
    FILE* saved_stdout = stdout;                 // the tracker also chats through printf
    FILE* devnull = fopen("/dev/null", "w");
    if (devnull) stdout = devnull;
    const float Kc = 20.0f, neff_ratio = 0.8f;
    const int vol = 1;
    const long size = (long)w * h * l;
    std::vector<float> sigs(sigmas, sigmas + nsig);
    std::vector<Node> n0;
    Node n00;
    n0.push_back(n00);
    std::vector<int> smap(size, 0);
    Tracker t(sigs, step, np, ni, kappa, l == 1, znccth, Kc, neff_ratio, zdist, nodepervol);
    std::vector<seed> seeds_init;
    SeedExtractor::extractSeeds(tolerance, J8, w, h, l, Vx, Vy, Vz, seeds_init);
    counts[0] = (long)seeds_init.size();
    float dummy_sig;
    for (long i = (long)seeds_init.size() - 1; i >= 0; --i) {
        const long j = (long)((int)round(seeds_init[i].z)) * w * h + (int)round(seeds_init[i].y) * w + (int)round(seeds_init[i].x);
        if (smap[j] > 0) seeds_init.erase(seeds_init.begin() + i);
        else {
            seeds_init[i].corr = t.znccBBB(seeds_init[i].x, seeds_init[i].y, seeds_init[i].z, seeds_init[i].vx,
                                           seeds_init[i].vy, seeds_init[i].vz, img, w, h, l, dummy_sig);
            if (seeds_init[i].corr < znccth) seeds_init.erase(seeds_init.begin() + i);
        }
    }
    std::vector<long> si(seeds_init.size());
    for (size_t i = 0; i < si.size(); ++i) si[i] = (long)i;
    std::sort(si.begin(), si.end(), [&](const int& a, const int& b) { return seeds_init[a].corr > seeds_init[b].corr; });
    std::vector<seed> seeds;
    for (size_t i = 0; i < si.size(); ++i) seeds.push_back(seeds_init[si[i]]);
    counts[1] = (long)seeds.size();
    for (long i = 0; i < (long)seeds.size() && i < seeds_cap; ++i) {
        float* r = seeds_out + 8 * i;
        r[0] = seeds[i].x; r[1] = seeds[i].y; r[2] = seeds[i].z; r[3] = seeds[i].vx; r[4] = seeds[i].vy; r[5] = seeds[i].vz;
        r[6] = seeds[i].score; r[7] = seeds[i].corr;
    }
    std::vector<long*> ioff(size, (long*)0);     // vol == 1: no neighbouring voxels
    std::vector<unsigned char> npervol_map(size, 0);
    std::vector<int> nidx_map(size, 0);
    long trace_count = 0;
    for (long i = 0; i < (long)seeds.size(); ++i) {
        const long sidx = (long)((int)round(seeds[i].z)) * w * h + (int)round(seeds[i].y) * w + (int)round(seeds[i].x);
        if ((int)npervol_map[sidx] < nodepervol) {
            trace_count++;
            t.trackPos(seeds[i], img, n0, w, h, l, smap.data(), npervol_map.data(), vol, ioff.data(), nidx_map.data());
            t.trackNeg(seeds[i], img, n0, w, h, l, smap.data(), npervol_map.data(), vol, ioff.data(), nidx_map.data());
            if (trace_count > max_traces) break;
        }
    }
    counts[2] = (long)n0.size();
    counts[4] = trace_count;
    long nn = 0;
    for (long i = 0; i < (long)n0.size(); ++i) {
        if (i < nodes_cap) {
            float* r = nodes_out + 10 * i;
            r[0] = n0[i].x; r[1] = n0[i].y; r[2] = n0[i].z; r[3] = n0[i].vx; r[4] = n0[i].vy; r[5] = n0[i].vz;
            r[6] = n0[i].corr; r[7] = n0[i].sig; r[8] = (float)n0[i].type; r[9] = (float)n0[i].nbr.size();
        }
        for (size_t k = 0; k < n0[i].nbr.size(); ++k, ++nn)
            if (nn < nbr_cap) nbr_out[nn] = n0[i].nbr[k];
    }
    counts[3] = nn;
    if (devnull) { fflush(devnull); stdout = saved_stdout; fclose(devnull); }
    return 0;
}

}  // extern "C"
