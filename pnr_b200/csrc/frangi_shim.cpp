// frangi_shim.cpp -- the Frangi class of frangi.h on top of the C-ABI (include/frangi_gpu.h).
#include "frangi.h"

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/frangi_gpu.h"

namespace {
[[noreturn]] void raise(const char* what, int rc)
{
    throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + frangi_gpu_last_error());
}
// The reference's constructor has no place for flags, and a call site compiled unchanged cannot set the public field:
// the environment variable PNR_FRANGI_FLAGS (a sum of FRANGI_GPU_FLAG_* values) supplies the initial value of `flags`,
// e.g. 32 = the reference's eigenvector signs, 1 = fused multiply-add smoothing.  Unset: 0 (bit-exact smoothing).
unsigned initial_flags()
{
    const char* e = std::getenv("PNR_FRANGI_FLAGS");
    return e && *e ? (unsigned)std::strtoul(e, nullptr, 0) : 0u;
}
}  // namespace

Frangi::Frangi(std::vector<float> sigs, float zdist_, float alpha_, float beta_, float C_, float beta_one, float beta_two)
    : sig(sigs), zdist(zdist_), alpha(alpha_), beta(beta_), BetaOne(beta_one), BetaTwo(beta_two), C(C_),
      blackwhite(false), flags(initial_flags()), handle_(nullptr), hw_(0), hh_(0), hl_(0), hz_(0), ha_(0), hb_(0), hc_(0),
      hbw_(false), hflags_(0)
{
}

Frangi::~Frangi() { release(); }

void Frangi::release()
{
    if (handle_) frangi_gpu_destroy(handle_);
    handle_ = nullptr;
}

// The reference re-reads its public fields on every call, so the handle is rebuilt whenever
// a field or the volume shape changed since the last call.
void Frangi::ensure_handle(int w, int h, int l)
{
    const bool same = handle_ && w == hw_ && h == hh_ && l == hl_ && sig == hsig_ && zdist == hz_ && alpha == ha_ &&
                      beta == hb_ && C == hc_ && blackwhite == hbw_ && flags == hflags_ && devices == hdev_;
    if (same) return;
    release();
    const int ndev = devices.empty() ? 1 : (int)devices.size();
    const int rc = frangi_gpu_create(&handle_, sig.data(), (int)sig.size(), zdist, alpha, beta, C, blackwhite ? 1 : 0,
                                     w, h, l, devices.empty() ? nullptr : devices.data(), ndev, flags);
    if (rc) { handle_ = nullptr; raise("frangi_gpu_create", rc); }
    hw_ = w; hh_ = h; hl_ = l; hsig_ = sig; hz_ = zdist; ha_ = alpha; hb_ = beta; hc_ = C; hbw_ = blackwhite;
    hflags_ = flags; hdev_ = devices;
}

void Frangi::frangi3d(unsigned char* I, int w, int h, int l, float* J, float& Jmin, float& Jmax,
                      unsigned char* Vx, unsigned char* Vy, unsigned char* Vz)
{
    frangi3d_j8(I, w, h, l, J, Jmin, Jmax, Vx, Vy, Vz, nullptr);
}

void Frangi::frangi3d_j8(unsigned char* I, int w, int h, int l, float* J, float& Jmin, float& Jmax,
                         unsigned char* Vx, unsigned char* Vy, unsigned char* Vz, unsigned char* J8)
{
    ensure_handle(w, h, l);
    float lo = 0, hi = 0;
    const int rc = frangi_gpu_run(handle_, I, J, &lo, &hi, Vx, Vy, Vz, J8, nullptr, nullptr);
    if (rc) raise("frangi_gpu_run", rc);
    Jmin = lo;
    Jmax = hi;
}

void Frangi::hessian3d(unsigned char* I, int w, int h, int l, float sig_, float zdist_,
                       float* Dzz, float* Dyy, float* Dyz, float* Dxx, float* Dxy, float* Dxz)
{
    const int dev = devices.empty() ? 0 : devices[0];
    const int rc = frangi_gpu_hessian3d(I, w, h, l, sig_, zdist_, Dzz, Dyy, Dyz, Dxx, Dxy, Dxz, dev,
                                        flags & FRANGI_GPU_FLAG_FMA_SMOOTHING);
    if (rc) raise("frangi_gpu_hessian3d", rc);
}

void Frangi::frangi2d(unsigned char* I, int w, int h, int l, float* J, float& Jmin, float& Jmax,
                      unsigned char* Vx, unsigned char* Vy, unsigned char* Vz)
{
    if (l != 1) throw std::runtime_error("Frangi::frangi2d: the image must be a single plane (l == 1)");
    const int dev = devices.empty() ? 0 : devices[0];
    float lo = 0, hi = 0;
    const int rc = frangi_gpu_frangi2d(I, w, h, sig.data(), (int)sig.size(), BetaOne, BetaTwo, blackwhite ? 1 : 0, J, &lo, &hi,
                                       Vx, Vy, Vz, dev, flags & FRANGI_GPU_FLAG_FMA_SMOOTHING);
    if (rc) raise("frangi_gpu_frangi2d", rc);
    Jmin = lo;
    Jmax = hi;
}

void Frangi::hessian2d(unsigned char* I, int w, int h, float sig_, float* Dyy, float* Dxy, float* Dxx)
{
    const int dev = devices.empty() ? 0 : devices[0];
    const int rc = frangi_gpu_hessian2d(I, w, h, sig_, Dyy, Dxy, Dxx, dev, flags & FRANGI_GPU_FLAG_FMA_SMOOTHING);
    if (rc) raise("frangi_gpu_hessian2d", rc);
}

void Frangi::imgaussian(unsigned char* I, int w, int h, int l, float sig_, float zdist_, float* F)
{
    const int rc = frangi_gpu_imgaussian(I, w, h, l, sig_, zdist_, F, 0, 0);
    if (rc) raise("frangi_gpu_imgaussian", rc);
}

void Frangi::imgaussian(unsigned char* I, int w, int h, int l, float sig_)
{
    const int rc = frangi_gpu_imgaussian_xy(I, w, h, l, sig_, 0);
    if (rc) raise("frangi_gpu_imgaussian_xy", rc);
}

void Frangi::imerode(unsigned char* I, int w, int h, int l, float rad, unsigned char* E)
{
    const int rc = frangi_gpu_imerode(I, w, h, l, rad, E, 0);
    if (rc) raise("frangi_gpu_imerode", rc);
}

void Frangi::imdilate(unsigned char* I, int w, int h, int l, float rad)
{
    const int rc = frangi_gpu_imdilate(I, w, h, l, rad, 0);
    if (rc) raise("frangi_gpu_imdilate", rc);
}

void Frangi::imgaussian(unsigned char* I, int w, int h, float sig_, float* F)
{
    const int rc = frangi_gpu_imgaussian2d(I, w, h, sig_, F, 0);
    if (rc) raise("frangi_gpu_imgaussian2d", rc);
}

void Frangi::imerode(unsigned char* I, int w, int h, int l, float rad, float zdist_, unsigned char* E)
{
    const int rc = frangi_gpu_imerode_z(I, w, h, l, rad, zdist_, E, 0);
    if (rc) raise("frangi_gpu_imerode_z", rc);
}
