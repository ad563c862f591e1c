// frangi.h -- source-compatible stand-in for the hot-path part of the reference's
// `class Frangi` (pnr-vaa3d/frangi.h:5-59), so that the reference's only call
// site compiles unchanged against the GPU library:
//
//     Frangi frangiflt(sigs, zdist, alpha, beta, C, beta_one, beta_two);    // Advantra_plugin.cpp:2488
//     frangiflt.frangi3d(data1d, N, M, P, J, Jmin, Jmax, Vx, Vy, Vz);       // Advantra_plugin.cpp:2496
//
// Same public field names, constructor and member signatures as the reference for
// everything ON the path (frangi.h:8-24,33,35,42); every call forwards to the C-ABI
// of include/frangi_gpu.h.  The 2-D pair frangi2d / hessian2d (frangi.h:38,40) is
// provided too, and so are the soma helpers imerode (both overloads) / imdilate / in-place xy
// imgaussian (frangi.h:46,47,49,43) and the 2-D imgaussian (frangi.h:44).  The members that are
// plain host helpers in the reference -- the public eigen-solver entry points (frangi.h:53-58;
// Advantra_plugin.cpp:1727 calls eigen_decomposition_static), the direction tables
// (frangi.h:18-20,28-31) and interpz (frangi.h:51) -- are host code here too
// (frangi_shim_host.cpp), so every Frangi:: symbol of the reference header links.
//
// Error behaviour: the reference's members return void and fail only by uncaught
// std::bad_alloc; here a failed GPU call throws std::runtime_error carrying
// frangi_gpu_last_error().  There is no CPU fallback.
#ifndef PNR_B200_FRANGI_SHIM_H
#define PNR_B200_FRANGI_SHIM_H

#include <vector>

struct frangi_gpu;   // opaque handle of include/frangi_gpu.h

class Frangi {
public:
    // ---- the reference's public fields (frangi.h:8-22) ----
    std::vector<float> sig;
    float zdist;
    float alpha;
    float beta;
    float BetaOne;   // 2-D only (frangi2d)
    float BetaTwo;   // 2-D only
    float C;
    // direction tables (frangi.h:18-20): declared by the reference, filled by nobody on the live path
    std::vector<std::vector<float> > Vxyz;
    static unsigned char ndirs2d;
    static unsigned char ndirs3d;
    bool blackwhite; // true: dark ridges, false: bright ridges (default, frangi.cpp:54)

    // ---- additions (defaults keep the reference's behaviour) ----
    std::vector<int> devices;   // CUDA devices to shard z-slabs over; empty = device 0
    unsigned flags;             // FRANGI_GPU_FLAG_* of include/frangi_gpu.h; 0 = bit-exact smoothing

    Frangi(std::vector<float> sigs, float zdist_, float alpha_, float beta_, float C_, float beta_one, float beta_two);
    ~Frangi();
    Frangi(const Frangi&) = delete;
    Frangi& operator=(const Frangi&) = delete;

    // frangi.h:28-31 (host helpers, frangi_shim_host.cpp)
    void generate_3d_unit_directions(unsigned char Ndir, std::vector<std::vector<float> >& Vxyz);
    void generate_2d_unit_directions(unsigned char Ndir, std::vector<std::vector<float> >& Vxyz);
    unsigned char get_direction_idx(float vx, float vy, float vz, std::vector<std::vector<float> > Vxyz);
    unsigned char get_direction_idx(float vx, float vy, std::vector<std::vector<float> > Vxyz);

    // frangi.h:33 -- caller owns every buffer (w*h*l elements each), callee only fills
    void frangi3d(unsigned char* I, int w, int h, int l, float* J, float& Jmin, float& Jmax,
                  unsigned char* Vx, unsigned char* Vy, unsigned char* Vz);

    // As above plus the 8-bit normalisation the caller applies next (Advantra_plugin.cpp:2499-2512)
    // done on the device; J may be NULL (the caller frees it at once, :2514).
    void frangi3d_j8(unsigned char* I, int w, int h, int l, float* J, float& Jmin, float& Jmax,
                     unsigned char* Vx, unsigned char* Vy, unsigned char* Vz, unsigned char* J8);

    // frangi.h:35
    void hessian3d(unsigned char* I, int w, int h, int l, float sig_, float zdist_,
                   float* Dzz, float* Dyy, float* Dyz, float* Dxx, float* Dxy, float* Dxz);

    // frangi.h:38 -- single-plane images (the plugin's P == 1 branch, Advantra_plugin.cpp:2497); l is ignored as in
    // the reference's arithmetic (it only scales the loop bounds there and is 1 at the call site)
    void frangi2d(unsigned char* I, int w, int h, int l, float* J, float& Jmin, float& Jmax,
                  unsigned char* Vx, unsigned char* Vy, unsigned char* Vz);

    // frangi.h:40
    void hessian2d(unsigned char* I, int w, int h, float sig_, float* Dyy, float* Dxy, float* Dxx);

    // frangi.h:42
    static void imgaussian(unsigned char* I, int w, int h, int l, float sig_, float zdist_, float* F);

    // frangi.h:43,47,49 -- the soma branch (Advantra_plugin.cpp:2432,2438)
    static void imgaussian(unsigned char* I, int w, int h, int l, float sig_);
    static void imerode(unsigned char* I, int w, int h, int l, float rad, unsigned char* E);
    static void imdilate(unsigned char* I, int w, int h, int l, float rad);
    // frangi.h:44,46 -- the overloads no live code calls (2-D smoothing, z-scaled erosion); on the device as well
    static void imgaussian(unsigned char* I, int w, int h, float sig_, float* F);
    static void imerode(unsigned char* I, int w, int h, int l, float rad, float zdist_, unsigned char* E);

    // frangi.h:51-58 (host helpers, frangi_shim_host.cpp); eigen_decomposition_static is what
    // Advantra_plugin.cpp:1727 calls
    float interpz(int x, int y, float z, float* img, int w, int h, int l);
    void eigen_decomposition(double A[3][3], double V[3][3], double d[3]);
    static void eigen_decomposition_static(double A[3][3], double V[3][3], double d[3]);
    static void tred2(double V[3][3], double d[3], double e[3]);
    static void tql2(double V[3][3], double d[3], double e[3]);
    static double hypot2(double x, double y);
    static double absd(double val) { return val > 0 ? val : -val; }

private:
    frangi_gpu* handle_;
    int hw_, hh_, hl_;
    std::vector<float> hsig_;
    float hz_, ha_, hb_, hc_;
    bool hbw_;
    unsigned hflags_;
    std::vector<int> hdev_;
    void ensure_handle(int w, int h, int l);
    void release();
};

#endif
