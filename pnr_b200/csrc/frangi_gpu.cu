// frangi_gpu.cu -- host side of the C-ABI declared in include/frangi_gpu.h.
//
// Owns device memory, streams, events and NCCL communicators; plans the tap
// tables exactly as the reference does on the host (frangi.cpp:651-680) and
// launches the kernels of frangi_kernels.cuh.  One `Slab` per device; a handle
// holds one slab (one-process-per-GPU jobs) or several (one process driving a
// whole box).  One slab: per scale xy pass (K1), z pass (K2), Hessian / eigen /
// vesselness / running max (K3), then the 8-bit map (K4); frangi_gpu_run pipelines
// upload, kernels and download over z chunks (run_streamed) and moves pageable host
// buffers through a pinned staging ring (host_stager.h).  Several slabs: the same
// per slab with the halo exchange of the xy-smoothed boundary planes (NCCL send /
// recv or peer copies on a side stream) pipelined two scales deep over two Fxy
// buffers (run_pipeline), then the Jmin / Jmax all-reduce.
#include "../../include/frangi_gpu.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_segmented_radix_sort.cuh>

#include "frangi_kernels.cuh"
#include "seed_kernels.cuh"
#include "frangi2d_kernels.cuh"
#include "soma_kernels.cuh"
#include "zncc_kernels.cuh"
#include "nccl_dyn.h"
#include "host_stager.h"

#define FRANGI_API extern "C" __attribute__((visibility("default")))

namespace {

using namespace frangi;

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(expr)                                                                         \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? FRANGI_GPU_ENOMEM : FRANGI_GPU_ECUDA, \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define NK(expr)                                                                         \
    do {                                                                                 \
        ncclx::ncclResult_t _r = (expr);                                                 \
        if (_r != ncclx::ncclSuccess)                                                    \
            return fail(FRANGI_GPU_ENCCL, "%s failed: %s (%s:%d)", #expr,                \
                        ncclx::api().GetErrorString ? ncclx::api().GetErrorString(_r) : "?", \
                        __FILE__, __LINE__);                                             \
    } while (0)

#define RC(expr)                      \
    do {                              \
        int _rc = (expr);             \
        if (_rc) return _rc;          \
    } while (0)

// An NCCL group that is closed on every path: NK() returns from the middle of a group on an error, and a group left
// open would swallow every later NCCL call of the thread (ncclCommDestroy included) -- one error must not become a hang.
struct NcclGroup {
    bool open = false;
    int begin()
    {
        if (ncclx::api().GroupStart() != ncclx::ncclSuccess) return fail(FRANGI_GPU_ENCCL, "ncclGroupStart failed");
        open = true;
        return 0;
    }
    int end()
    {
        open = false;
        const ncclx::ncclResult_t r = ncclx::api().GroupEnd();
        if (r != ncclx::ncclSuccess)
            return fail(FRANGI_GPU_ENCCL, "ncclGroupEnd failed: %s",
                        ncclx::api().GetErrorString ? ncclx::api().GetErrorString(r) : "?");
        return 0;
    }
    ~NcclGroup() { if (open) ncclx::api().GroupEnd(); }
};


// ---- tap planning (host; mirrors frangi.cpp:651-680 in float32) -------------
#ifndef Z_TMA
#define Z_TMA 1      // z pass by the TMA warp-stream kernel (0: the register-prefetch marching kernel)
#endif
constexpr int kEvPerScale = 5;   // timing events recorded per scale (see collect())
constexpr int kRadii[] = { 3, 6, 9, 12, 15, 18, 24, 30 };
constexpr int kNumRadii = sizeof(kRadii) / sizeof(kRadii[0]);

int reference_radius(float sigma) { return (int)std::ceil(3 * sigma); }

// Taps of true radius r are centred in a table of template radius R >= r and
// padded with zeros: acc + v*0 leaves acc untouched bit for bit, so a kernel
// instantiated for R reproduces the radius-r filter exactly.
int plan_taps(float sigma, int& r_true, int& r_tmpl, GaussTaps& t)
{
    r_true = reference_radius(sigma);
    r_tmpl = -1;
    for (int i = 0; i < kNumRadii; ++i)
        if (kRadii[i] >= r_true) { r_tmpl = kRadii[i]; break; }
    if (r_tmpl < 0 || r_true < 0)
        return fail(FRANGI_GPU_EINVAL, "sigma %g needs tap radius %d > supported %d", sigma, r_true, kMaxRadius);
    std::vector<float> g(2 * r_true + 1);
    float norm = 0;
    for (int i = -r_true; i <= r_true; ++i) {
        g[i + r_true] = std::exp(-(i * i) / (2 * sigma * sigma));  // float exp, as the reference
        norm += g[i + r_true];
    }
    for (auto& v : g) v /= norm;
    std::memset(&t, 0, sizeof t);
    for (int i = -r_true; i <= r_true; ++i) t.g[i + r_tmpl] = g[i + r_true];
    return 0;
}

struct ScalePlan {
    float sigma, sigma2;
    int rxy, rxy_t, rz, rz_t;
    GaussTaps txy, tz;
};

// ---- kernel dispatch ----------------------------------------------------------
template <int L, bool EXACT>
int launch_xy_t(const XYParams& p, const GaussTaps& t, int nblocks, cudaStream_t s)
{
    using C = XYCfg<L>;
    auto k = gauss_xy_kernel<L, EXACT>;
    static thread_local int configured_dev[64] = { 0 };
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 64 && !configured_dev[dev]) {
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured_dev[dev] = 1;
    }
    k<<<nblocks, C::NT, C::SMEM_BYTES, s>>>(p, t);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int L>
int launch_xy_fma_t(const XYParams& p, const GaussTaps& t, int nblocks, cudaStream_t s)
{
    using C = XYFmaCfg<L>;
    auto k = gauss_xy_fma_kernel<L>;
    static thread_local int configured_dev[64] = { 0 };
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 64 && !configured_dev[dev]) {
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured_dev[dev] = 1;
    }
    k<<<nblocks, C::NT, C::SMEM_BYTES, s>>>(p, t);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

#ifndef XY_WARP_KERNEL
#define XY_WARP_KERNEL 1           // FMA mode, radius <= 18: the warp-autonomous form (gauss_xy_warp_kernel)
#endif
template <int L>
int launch_xy_warp_t(const XYWarpParams& p, const GaussTaps& t, cudaStream_t s)
{
    using C = XYWarpCfg<L>;
    auto k = gauss_xy_warp_kernel<L>;
    static thread_local int configured_dev[64] = { 0 };
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 64 && !configured_dev[dev]) {
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured_dev[dev] = 1;
    }
    const long long nblocks = (p.items + C::NW - 1) / C::NW;
    if (nblocks > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "grid too large");
    k<<<(unsigned)nblocks, 32 * C::NW, C::SMEM_BYTES, s>>>(p, t);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

// Warps resident per SM of the warp-autonomous form, 0 where the CTA-wide form is used.  Measured at 2048 x 2048 x 512
// (profiles/r3c): radius 6: 3.26 ms against 3.74 (16 warps per SM, no barrier, FMA pipe 56 %); radius 12: 5.85 against
// 4.87, radius 18: 9.53 against 6.61 -- the private tiles leave only 11 / 8 warps per SM there, too few to cover the
// fixed-latency dependencies between the passes.  So: small radii only.
#ifndef XY_WARP_MAXL
#define XY_WARP_MAXL 6
#endif
int xy_warp_width(int L)
{
    if (L > XY_WARP_MAXL) return 0;
    switch (L) {
        case 3: return XYWarpCfg<3>::NW;
        case 6: return XYWarpCfg<6>::NW;
        case 9: return XYWarpCfg<9>::NW;
        case 12: return XYWarpCfg<12>::NW;
        case 15: return XYWarpCfg<15>::NW;
        case 18: return XYWarpCfg<18>::NW;
    }
    return 0;
}

int launch_xy_warp(int L, const XYWarpParams& p, const GaussTaps& t, cudaStream_t s)
{
    switch (L) {
        case 3: return launch_xy_warp_t<3>(p, t, s);
        case 6: return launch_xy_warp_t<6>(p, t, s);
        case 9: return launch_xy_warp_t<9>(p, t, s);
        case 12: return launch_xy_warp_t<12>(p, t, s);
        case 15: return launch_xy_warp_t<15>(p, t, s);
        case 18: return launch_xy_warp_t<18>(p, t, s);
    }
    return fail(FRANGI_GPU_EINVAL, "unsupported xy radius %d", L);
}

int launch_xy_fma(int L, const XYParams& p, const GaussTaps& t, int nblocks, cudaStream_t s)
{
    switch (L) {
        case 3: return launch_xy_fma_t<3>(p, t, nblocks, s);
        case 6: return launch_xy_fma_t<6>(p, t, nblocks, s);
        case 9: return launch_xy_fma_t<9>(p, t, nblocks, s);
        case 12: return launch_xy_fma_t<12>(p, t, nblocks, s);
        case 15: return launch_xy_fma_t<15>(p, t, nblocks, s);
        case 18: return launch_xy_fma_t<18>(p, t, nblocks, s);
        case 24: return launch_xy_fma_t<24>(p, t, nblocks, s);
        case 30: return launch_xy_fma_t<30>(p, t, nblocks, s);
    }
    return fail(FRANGI_GPU_EINVAL, "no gauss_xy instantiation for radius %d", L);
}

template <bool EXACT>
int launch_xy_e(int L, const XYParams& p, const GaussTaps& t, int nblocks, cudaStream_t s)
{
    switch (L) {
        case 3: return launch_xy_t<3, EXACT>(p, t, nblocks, s);
        case 6: return launch_xy_t<6, EXACT>(p, t, nblocks, s);
        case 9: return launch_xy_t<9, EXACT>(p, t, nblocks, s);
        case 12: return launch_xy_t<12, EXACT>(p, t, nblocks, s);
        case 15: return launch_xy_t<15, EXACT>(p, t, nblocks, s);
        case 18: return launch_xy_t<18, EXACT>(p, t, nblocks, s);
        case 24: return launch_xy_t<24, EXACT>(p, t, nblocks, s);
        case 30: return launch_xy_t<30, EXACT>(p, t, nblocks, s);
    }
    return fail(FRANGI_GPU_EINVAL, "no gauss_xy instantiation for radius %d", L);
}

template <int L, bool EXACT>
int launch_z_t(const ZParams& p, const GaussTaps& t, long long nblocks, cudaStream_t s)
{
    gauss_z_kernel<L, EXACT><<<(unsigned)nblocks, 128, 0, s>>>(p, t);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

template <bool EXACT>
int launch_z_e(int L, const ZParams& p, const GaussTaps& t, long long nblocks, cudaStream_t s)
{
    switch (L) {
        case 3: return launch_z_t<3, EXACT>(p, t, nblocks, s);
        case 6: return launch_z_t<6, EXACT>(p, t, nblocks, s);
        case 9: return launch_z_t<9, EXACT>(p, t, nblocks, s);
        case 12: return launch_z_t<12, EXACT>(p, t, nblocks, s);
        case 15: return launch_z_t<15, EXACT>(p, t, nblocks, s);
        case 18: return launch_z_t<18, EXACT>(p, t, nblocks, s);
        case 24: return launch_z_t<24, EXACT>(p, t, nblocks, s);
        case 30: return launch_z_t<30, EXACT>(p, t, nblocks, s);
    }
    return fail(FRANGI_GPU_EINVAL, "no gauss_z instantiation for radius %d", L);
}

// Device scratch of the pre-pass, kept between calls (allocation and release are synchronising and slow next to the
// kernels): per-layer ranges / counters / offsets, the two key arrays of the sort and its temporary storage.
struct SeedScratch {
    int *minmax = nullptr, *count = nullptr;
    long long *off = nullptr, *in = nullptr, *out = nullptr;
    void* tmp = nullptr;
    int layers = 0;
    long long keys = 0;
    size_t tmp_bytes = 0;
    void release()
    {
        cudaFree(minmax); cudaFree(count); cudaFree(off); cudaFree(in); cudaFree(out); cudaFree(tmp);
        *this = SeedScratch();
    }
};

// ---- one z-slab on one device -----------------------------------------------
struct Slab {
    int dev = 0;
    int index = 0;            // position of the slab in the volume (global rank)
    int zb = 0, ze = 0;       // own planes
    int fb = 0, fe = 0;       // planes of F held: own +-2, clipped to the volume
    int xb = 0, xe = 0;       // planes of Fxy held: F planes +- max z radius, clipped
    long long voxels = 0;     // own voxels
    uint8_t* dI = nullptr;
    float* dFxy = nullptr;
    float* dF = nullptr;
    float* dJ = nullptr;
    uint8_t *dVx = nullptr, *dVy = nullptr, *dVz = nullptr, *dScale = nullptr, *dJ8 = nullptr;
    float* dDir = nullptr;
    CUtensorMap tmF{}, tmFc{};   // TMA descriptors of dF: boxes of the full-eigen and of the compacting K3 kernel
    CUtensorMap tmFxy{};         // TMA descriptor of dFxy: 64 x 1 x 1 row segments for the z pass
    const float* dFxy0 = nullptr; // the pointer tmFxy was encoded with (views move dFxy and / or xb, the map stays)
    // overlapped schedule (one-slab handles, run_pipeline): a second Fxy / F buffer pair with its descriptors, the
    // low-priority stream of the z pass, and per-kernel timing events (6 per scale and timing set)
    float *dFxyB = nullptr, *dFB = nullptr;
    float* dFxyA = nullptr;        // the first Fxy buffer (dFxy points at the one of the scale in hand)
    CUtensorMap tmFxyA{};
    cudaEvent_t ev_halo2 = nullptr; // halo events alternate with the Fxy buffers (ev_halo: even scales)
    CUtensorMap tmFB{}, tmFcB{}, tmFxyB{};
    cudaStream_t s_aux = nullptr;
    std::vector<cudaEvent_t> ev_ov;
    SeedScratch seed;               // scratch of frangi_gpu_seed_candidates
    bool has_tm = false;
    int* dMinMax = nullptr;
    int* hMinMax = nullptr;   // pinned
    cudaStream_t s_main = nullptr, s_comm = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev_chunk;  // streamed run: input-ready / output-ready per chunk
    cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr, ev_done = nullptr;
    int* dGather = nullptr;   // local-copy mode, slab 0 only: min/max of every slab
    std::vector<cudaEvent_t> ev_all;    // timing_depth sets of (4 per scale + 2) events
    cudaEvent_t* ev_time = nullptr;     // the set of the run being recorded
    ncclx::ncclComm_t comm = nullptr;
};

}  // namespace

struct frangi_gpu {
    int w = 0, h = 0, l = 0;
    int fpitch = 0;
    long long fplane = 0;
    float zdist = 1, alpha = .5f, beta = .5f, C = 500;
    int blackwhite = 0;
    unsigned flags = 0;
    int nslabs_total = 1;        // slabs in the whole job (all processes)
    std::string warnings;        // conditions that do not fail a call but that the caller should know about (frangi_gpu_warnings)
    int rz_max = 0;
    std::vector<ScalePlan> scales;
    std::vector<Slab> slabs;     // slabs driven by this process
    bool ran = false;
    bool local_halo = false;     // halos move by peer copies inside this process instead of NCCL
    int stream_chunk = -1;       // frangi_gpu_run: planes per pipelined chunk; 0 = off, -1 = automatic
    HostStager stager;           // pinned slots + host threads for calls with pageable host buffers (frangi_gpu_run)
    bool last_streamed = false;  // the last run recorded no per-class events
    bool overlap = false;        // one-slab handle on the overlapped schedule (see run_pipeline)
    int timing_depth = 1;        // event sets kept per slab
    long long runs_recorded = 0; // runs since the last frangi_gpu_timing_depth call
    float last_ms[8] = { 0 };
};

namespace {

void free_slab(Slab& s)
{
    cudaSetDevice(s.dev);
    if (s.comm && ncclx::api().ok) ncclx::api().CommDestroy(s.comm);
    cudaFree(s.dI); cudaFree(s.dFxyA ? s.dFxyA : s.dFxy); cudaFree(s.dF); cudaFree(s.dJ);
    cudaFree(s.dVx); cudaFree(s.dVy); cudaFree(s.dVz); cudaFree(s.dScale); cudaFree(s.dJ8);
    cudaFree(s.dDir); cudaFree(s.dMinMax);
    s.seed.release();
    if (s.hMinMax) cudaFreeHost(s.hMinMax);
    for (auto e : s.ev_all) cudaEventDestroy(e);
    if (s.ev_boundary) cudaEventDestroy(s.ev_boundary);
    if (s.ev_halo) cudaEventDestroy(s.ev_halo);
    if (s.ev_halo2) cudaEventDestroy(s.ev_halo2);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    cudaFree(s.dGather);
    cudaFree(s.dFxyB); cudaFree(s.dFB);
    for (auto e : s.ev_ov) cudaEventDestroy(e);
    if (s.s_aux) cudaStreamDestroy(s.s_aux);
    if (s.s_main) cudaStreamDestroy(s.s_main);
    if (s.s_comm) cudaStreamDestroy(s.s_comm);
    if (s.s_h2d) cudaStreamDestroy(s.s_h2d);
    if (s.s_d2h) cudaStreamDestroy(s.s_d2h);
    for (auto e : s.ev_chunk) cudaEventDestroy(e);
    s = Slab();
}

int check_device(int dev)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(FRANGI_GPU_ECUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (dev < 0 || dev >= n) return fail(FRANGI_GPU_EINVAL, "device %d out of range (0..%d)", dev, n - 1);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(FRANGI_GPU_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", dev,
                    prop.major, prop.minor);
    return 0;
}

int plan_common(frangi_gpu* H, const float* sigmas, int nsig, float zdist, float alpha, float beta,
                float C, int blackwhite, int w, int h, int l, unsigned flags)
{
    if (!sigmas || nsig < 1 || nsig > 64) return fail(FRANGI_GPU_EINVAL, "nsig must be 1..64");
    if (w < 2 || h < 2 || l < 2)
        return fail(FRANGI_GPU_EINVAL, "frangi3d needs w,h,l >= 2 (got %d,%d,%d); 2-D images are out of scope", w, h, l);
    if ((long long)((w + 31) / 32 * 32) * h > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "plane of %d x %d is too large", w, h);
    if (!(zdist > 0)) return fail(FRANGI_GPU_EINVAL, "zdist must be > 0");
    H->w = w; H->h = h; H->l = l;
    H->fpitch = (w + 31) / 32 * 32;
    H->fplane = (long long)H->fpitch * h;
    H->zdist = zdist; H->alpha = alpha; H->beta = beta; H->C = C;
    H->blackwhite = blackwhite; H->flags = flags;
    H->scales.resize(nsig);
    H->rz_max = 0;
    for (int i = 0; i < nsig; ++i) {
        ScalePlan& sp = H->scales[i];
        if (!(sigmas[i] > 0)) return fail(FRANGI_GPU_EINVAL, "sigma[%d] must be > 0", i);
        sp.sigma = sigmas[i];
        sp.sigma2 = sigmas[i] * sigmas[i];
        RC(plan_taps(sp.sigma, sp.rxy, sp.rxy_t, sp.txy));
        const float sigma_z = sp.sigma / zdist;   // frangi.cpp:649
        RC(plan_taps(sigma_z, sp.rz, sp.rz_t, sp.tz));
        if (sp.rz > H->rz_max) H->rz_max = sp.rz;
    }
    return 0;
}

// TMA descriptor of a slab's F buffer seen as a (w, h, planes) float tensor with rows `fpitch` floats apart;
// a box is one (PW x PH x 1) plane tile.  Columns >= w and rows >= h are outside the tensor, so the copy
// engine zero-fills them (as it does for negative coordinates).  cuTensorMapEncodeTiled is a driver entry
// point: it is fetched through the runtime, so the library has no link-time dependency on libcuda.
int make_tile_map(CUtensorMap* tm, float* F, int w, int h, int planes, int fpitch, long long fplane, int box_w, int box_h)
{
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    // fetched once, also when handles are created from several threads (a function-local static is initialised under a lock)
    static const encode_fn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (encode_fn)fn;
    }();
    if (!encode) return fail(FRANGI_GPU_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t dims[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes };
    const cuuint64_t strides[2] = { (cuuint64_t)fpitch * 4, (cuuint64_t)fplane * 4 };
    const cuuint32_t box[3] = { (cuuint32_t)box_w, (cuuint32_t)box_h, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, F, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FRANGI_GPU_ECUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d volume", (int)r, w, h, planes);
    return 0;
}

int alloc_slab(frangi_gpu* H, Slab& s, int dev, int index, int zb, int ze)
{
    s.dev = dev; s.index = index; s.zb = zb; s.ze = ze;
    s.fb = std::max(zb - 2, 0);
    s.fe = std::min(ze + 2, H->l);
    s.xb = std::max(s.fb - H->rz_max, 0);
    s.xe = std::min(s.fe + H->rz_max, H->l);
    s.voxels = (long long)H->w * H->h * (ze - zb);
    CK(cudaSetDevice(dev));
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));     // numerically lowest = highest priority
    CK(cudaStreamCreateWithPriority(&s.s_main, cudaStreamNonBlocking, prio_hi));
    CK(cudaStreamCreateWithPriority(&s.s_aux, cudaStreamNonBlocking, prio_lo));
    CK(cudaStreamCreateWithFlags(&s.s_comm, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.s_d2h, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&s.ev_boundary, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_halo, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    s.ev_all.resize(kEvPerScale * H->scales.size() + 2);
    for (auto& e : s.ev_all) CK(cudaEventCreate(&e));
    s.ev_time = s.ev_all.data();
    CK(cudaMalloc(&s.dI, (size_t)s.voxels));
    CK(cudaMalloc(&s.dFxy, sizeof(float) * (size_t)H->fplane * (s.xe - s.xb)));
    CK(cudaMalloc(&s.dF, sizeof(float) * (size_t)H->fplane * (s.fe - s.fb)));
    RC(make_tile_map(&s.tmFxy, s.dFxy, H->w, H->h, s.xe - s.xb, H->fpitch, H->fplane, ZTile::COLS, 1));
    s.dFxy0 = s.dFxy;
    // finite from the start: a z pass whose template radius exceeds the true one reads a few planes with zero taps,
    // possibly halo planes that have not arrived yet (0 x finite = 0; never 0 x garbage)
    CK(cudaMemset(s.dFxy, 0, sizeof(float) * (size_t)H->fplane * (s.xe - s.xb)));
    s.dFxyA = s.dFxy; s.tmFxyA = s.tmFxy;
    if (H->nslabs_total > 1 && H->scales.size() > 1) {
        // multi-slab: two Fxy buffers alternate by scale, so that the xy pass and the halo exchange of scale s+1 (s+2)
        // are issued before the z pass of scale s has read its own buffer (run_pipeline)
        CK(cudaMalloc(&s.dFxyB, sizeof(float) * (size_t)H->fplane * (s.xe - s.xb)));
        CK(cudaMemset(s.dFxyB, 0, sizeof(float) * (size_t)H->fplane * (s.xe - s.xb)));
        RC(make_tile_map(&s.tmFxyB, s.dFxyB, H->w, H->h, s.xe - s.xb, H->fpitch, H->fplane, ZTile::COLS, 1));
        CK(cudaEventCreateWithFlags(&s.ev_halo2, cudaEventDisableTiming));
    }
    if (H->overlap) {
        CK(cudaMalloc(&s.dFxyB, sizeof(float) * (size_t)H->fplane * (s.xe - s.xb)));
        CK(cudaMalloc(&s.dFB, sizeof(float) * (size_t)H->fplane * (s.fe - s.fb)));
        CK(cudaMemset(s.dFxyB, 0, sizeof(float) * (size_t)H->fplane * (s.xe - s.xb)));
        RC(make_tile_map(&s.tmFxyB, s.dFxyB, H->w, H->h, s.xe - s.xb, H->fpitch, H->fplane, ZTile::COLS, 1));
        RC(make_tile_map(&s.tmFB, s.dFB, H->w, H->h, s.fe - s.fb, H->fpitch, H->fplane, HessTile::PW, HessTile::PH));
        RC(make_tile_map(&s.tmFcB, s.dFB, H->w, H->h, s.fe - s.fb, H->fpitch, H->fplane, HessTileC::PW, HessTileC::PH));
        s.ev_ov.resize(6 * H->scales.size());
        for (auto& e : s.ev_ov) CK(cudaEventCreate(&e));
    }
    if (H->w >= 5 && H->h >= 5 && H->l >= 5) {   // thinner volumes go to the shell kernel entirely (launch_voxel)
        RC(make_tile_map(&s.tmF, s.dF, H->w, H->h, s.fe - s.fb, H->fpitch, H->fplane, HessTile::PW, HessTile::PH));
        RC(make_tile_map(&s.tmFc, s.dF, H->w, H->h, s.fe - s.fb, H->fpitch, H->fplane, HessTileC::PW, HessTileC::PH));
        s.has_tm = true;
    }
    CK(cudaMalloc(&s.dJ, sizeof(float) * (size_t)s.voxels));
    CK(cudaMalloc(&s.dVx, (size_t)s.voxels));
    CK(cudaMalloc(&s.dVy, (size_t)s.voxels));
    CK(cudaMalloc(&s.dVz, (size_t)s.voxels));
    CK(cudaMalloc(&s.dJ8, (size_t)s.voxels));
    if (H->flags & (FRANGI_GPU_FLAG_SCALE_IDX | FRANGI_GPU_FLAG_REFERENCE_DIRECTION)) CK(cudaMalloc(&s.dScale, (size_t)s.voxels));
    if (H->flags & FRANGI_GPU_FLAG_DIR_F32) CK(cudaMalloc(&s.dDir, sizeof(float) * 3 * (size_t)s.voxels));
    CK(cudaMalloc(&s.dMinMax, 2 * sizeof(int)));
    CK(cudaHostAlloc(&s.hMinMax, 2 * sizeof(int), cudaHostAllocDefault));
    return 0;
}

// xy smoothing (K1) of nz dense planes I -> out
#ifndef XY_FMA_PACKED
#define XY_FMA_PACKED 1
#endif
int launch_xy_planes(const uint8_t* I, float* out, int w, int h, int nz, int fpitch, long long fplane, const ScalePlan& sp,
                     unsigned flags, cudaStream_t st, int zsplit = -1, int zgap = 0)
{
    if (nz <= 0) return 0;
    XYParams p;
    p.I = I; p.out = out;
    p.w = w; p.h = h; p.nz = nz;
    p.zsplit = zsplit < 0 ? nz : zsplit; p.zgap = zgap;
    p.fpitch = fpitch; p.fplane = fplane;
    if (XY_WARP_KERNEL && (flags & FRANGI_GPU_FLAG_FMA_SMOOTHING) && XY_FMA_PACKED && xy_warp_width(sp.rxy_t) > 0) {
        // warp-autonomous form: items = 64-column strips x y segments x planes, one item per warp
        XYWarpParams q;
        q.b = p;
        q.b.nstrips = (w + 63) / 64;
        const long long slots = 148LL * xy_warp_width(sp.rxy_t);
        const long long per_seg = (long long)q.b.nstrips * nz;
        const int cand[] = { 1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32 };
        double best = 0;
        int best_h = ((h + 15) / 16) * 16;
        for (int c : cand) {
            const int sh = ((h + c - 1) / c + 15) / 16 * 16;
            if (c > 1 && sh < 32) break;
            const int ns = (h + sh - 1) / sh;
            const long long waves = (per_seg * ns + slots - 1) / slots;
            const double cost = (double)waves * (sh + 2 * sp.rxy + 8);
            if (best == 0 || cost < best * 0.999) { best = cost; best_h = sh; }
        }
        q.b.seg_h = best_h;
        q.b.nsegs = (h + best_h - 1) / best_h;
        q.b.vec_ok = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.I) & 3) == 0);
        q.items = (long long)q.b.nstrips * q.b.nsegs * nz;
        return launch_xy_warp(sp.rxy_t, q, sp.txy, st);
    }
    p.nstrips = (w + 255) / 256;
    // y is split into segments so that the launch fills the GPU.  Each segment redoes the x pass of its 2L halo rows
    // and a partly filled last wave costs a whole wave (a slab's boundary launch is one or two waves), so the split is
    // chosen by a small cost model: waves x (rows per segment + halo rows), over a handful of candidates.
    {
        const bool fma = (flags & FRANGI_GPU_FLAG_FMA_SMOOTHING) && XY_FMA_PACKED;
        const long long slots = 148LL * ((fma && sp.rxy_t <= XY_FMA_CTAS3_MAXL) ? 3 : 2);
        const long long per_seg = (long long)p.nstrips * p.nz;
        const int cand[] = { 1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32 };
        double best = 0;
        int best_h = ((h + 15) / 16) * 16;
        for (int c : cand) {
            const int sh = ((h + c - 1) / c + 15) / 16 * 16;
            if (c > 1 && sh < 32) break;
            const int ns = (h + sh - 1) / sh;
            const long long waves = (per_seg * ns + slots - 1) / slots;
            const double cost = (double)waves * (sh + 2 * sp.rxy + 8);
            if (best == 0 || cost < best * 0.999) { best = cost; best_h = sh; }
        }
        p.seg_h = best_h;
        p.nsegs = (h + best_h - 1) / best_h;
    }
    p.vec_ok = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.I) & 3) == 0);
    const long long nblocks = (long long)p.nstrips * p.nsegs * p.nz;
    if (nblocks > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "grid too large");
    if (flags & FRANGI_GPU_FLAG_FMA_SMOOTHING)
        return XY_FMA_PACKED ? launch_xy_fma(sp.rxy_t, p, sp.txy, (int)nblocks, st)
                             : launch_xy_e<false>(sp.rxy_t, p, sp.txy, (int)nblocks, st);
    return launch_xy_e<true>(sp.rxy_t, p, sp.txy, (int)nblocks, st);
}

// xy smoothing of own planes [z0, z1) of slab s for one scale
// the planes [z0, z0 + n) and [z1 - n, z1) of a slab (its two boundary groups, n < (z1 - z0) / 2) in ONE launch
int launch_xy_ends(frangi_gpu* H, Slab& s, const ScalePlan& sp, const uint8_t* I_own, int z0, int z1, int n)
{
    return launch_xy_planes(I_own + (long long)(z0 - s.zb) * H->w * H->h, s.dFxy + (long long)(z0 - s.xb) * H->fplane, H->w, H->h,
                            2 * n, H->fpitch, H->fplane, sp, H->flags, s.s_main, n, (z1 - z0) - 2 * n);
}

int launch_xy(frangi_gpu* H, Slab& s, const ScalePlan& sp, const uint8_t* I_own, int z0, int z1)
{
    if (z1 <= z0) return 0;
    return launch_xy_planes(I_own + (long long)(z0 - s.zb) * H->w * H->h, s.dFxy + (long long)(z0 - s.xb) * H->fplane, H->w, H->h,
                            z1 - z0, H->fpitch, H->fplane, sp, H->flags, s.s_main);
}

template <int L, bool EXACT>
int launch_zm_t(const ZParams& p, const GaussTaps& t, long long nblocks, cudaStream_t s)
{
    gauss_z_march_kernel<L, EXACT><<<(unsigned)nblocks, ZM_THREADS, 0, s>>>(p, t);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int L, bool EXACT>
int launch_zt_t(const CUtensorMap& tm, const ZParams& p, const GaussTaps& t, long long nwarps, cudaStream_t s)
{
    const long long nblocks = (nwarps + ZTile::WARPS - 1) / ZTile::WARPS;
    if (nblocks > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "grid too large");
    gauss_z_tma_kernel<L, EXACT><<<(unsigned)nblocks, 32 * ZTile::WARPS, ZTile::SMEM_BYTES, s>>>(tm, p, t);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

template <bool EXACT>
int launch_zt_e(int L, const CUtensorMap& tm, const ZParams& p, const GaussTaps& t, long long nwarps, cudaStream_t s)
{
    switch (L) {
        case 3: return launch_zt_t<3, EXACT>(tm, p, t, nwarps, s);
        case 6: return launch_zt_t<6, EXACT>(tm, p, t, nwarps, s);
        case 9: return launch_zt_t<9, EXACT>(tm, p, t, nwarps, s);
        case 12: return launch_zt_t<12, EXACT>(tm, p, t, nwarps, s);
    }
    return fail(FRANGI_GPU_EINVAL, "no TMA gauss_z instantiation for radius %d", L);
}

template <bool EXACT>
int launch_zm_e(int L, const ZParams& p, const GaussTaps& t, long long nblocks, cudaStream_t s)
{
    switch (L) {
        case 3: return launch_zm_t<3, EXACT>(p, t, nblocks, s);
        case 6: return launch_zm_t<6, EXACT>(p, t, nblocks, s);
        case 9: return launch_zm_t<9, EXACT>(p, t, nblocks, s);
        case 12: return launch_zm_t<12, EXACT>(p, t, nblocks, s);
    }
    return fail(FRANGI_GPU_EINVAL, "no marching gauss_z instantiation for radius %d", L);
}

// z smoothing of the F planes [f0, f1) of slab s (default: all of [s.fb, s.fe)); plane p lands at plane p - s.fb of s.dF
int launch_z(frangi_gpu* H, Slab& s, const ScalePlan& sp, cudaStream_t stream = nullptr, int f0 = -1, int f1 = -1)
{
    if (f0 < 0) { f0 = s.fb; f1 = s.fe; }
    if (f1 <= f0) return 0;
    ZParams p;
    p.in = s.dFxy; p.out = s.dF + (size_t)(f0 - s.fb) * H->fplane;
    p.w = H->w; p.h = H->h; p.l = H->l;
    p.fpitch = H->fpitch; p.fplane = H->fplane;
    p.in_base = s.xb; p.in_count = s.xe - s.xb;
    p.out_base = f0; p.out_count = f1 - f0;
    const bool fma = (H->flags & FRANGI_GPU_FLAG_FMA_SMOOTHING) != 0;
    cudaStream_t st = stream ? stream : s.s_main;
    if (Z_TMA && sp.rz_t <= 12) {
        // TMA form: one warp per 64-column row segment, one chunk per column unless that cannot fill the GPU
        p.nxs = (H->w + ZTile::COLS - 1) / ZTile::COLS;
        p.tm_base = s.xb - (int)((s.dFxy - s.dFxy0) / H->fplane);   // global plane of plane 0 of the tensor map
        const long long cols = (long long)p.nxs * H->h;
        long long nzc = std::max<long long>(1, std::min<long long>((148 * 32 + cols - 1) / cols, (p.out_count + 15) / 16));
        p.zchunk = (int)((p.out_count + nzc - 1) / nzc);
        nzc = (p.out_count + p.zchunk - 1) / p.zchunk;
        p.nzc = (int)nzc;
        if (fma) return launch_zt_e<false>(sp.rz_t, s.tmFxy, p, sp.tz, cols * nzc, st);
        return launch_zt_e<true>(sp.rz_t, s.tmFxy, p, sp.tz, cols * nzc, st);
    }
    if (sp.rz_t <= 12) {
        // marching form: one chunk per column unless the plane alone cannot fill the GPU
        p.nxs = (H->w + 2 * ZM_THREADS - 1) / (2 * ZM_THREADS);
        const long long cols = (long long)p.nxs * H->h;
        long long nzc = std::max<long long>(1, std::min<long long>((1184 + cols - 1) / cols, (p.out_count + 15) / 16));
        p.zchunk = (int)((p.out_count + nzc - 1) / nzc);
        nzc = (p.out_count + p.zchunk - 1) / p.zchunk;
        p.nzc = (int)nzc;
        const long long nblocks = cols * nzc;
        if (nblocks > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "grid too large");
        if (fma) return launch_zm_e<false>(sp.rz_t, p, sp.tz, nblocks, s.s_main);
        return launch_zm_e<true>(sp.rz_t, p, sp.tz, nblocks, s.s_main);
    }
    p.zchunk = 8;
    p.nzc = (p.out_count + 7) / 8;
    p.nxs = (H->w + 127) / 128;
    const long long nblocks = (long long)p.nzc * p.nxs * H->h;
    if (nblocks > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "grid too large");
    if (fma) return launch_z_e<false>(sp.rz_t, p, sp.tz, nblocks, s.s_main);
    return launch_z_e<true>(sp.rz_t, p, sp.tz, nblocks, s.s_main);
}

FView make_fview(frangi_gpu* H, Slab& s)
{
    FView f;
    f.F = s.dF; f.w = H->w; f.h = H->h; f.l = H->l;
    f.fpitch = H->fpitch; f.fplane = H->fplane; f.base = s.fb; f.count = s.fe - s.fb;
    return f;
}

FrangiConsts make_consts(frangi_gpu* H, float sigma2)
{
    FrangiConsts k;
    const float a2 = 2 * H->alpha * H->alpha, b2 = 2 * H->beta * H->beta, c2 = 2 * H->C * H->C;
    k.inv_2a2 = 1.0f / a2; k.inv_2b2 = 1.0f / b2; k.inv_2c2 = 1.0f / c2;
    k.sigma2 = sigma2; k.blackwhite = H->blackwhite;
    return k;
}

template <int MODE, bool BRIGHT = false>
int launch_voxel_t(const VoxelParams& p, long long nblocks, long long nshell, cudaStream_t st)
{
    if (nblocks > 0) {
        auto k = hessian_eigen_kernel<MODE, BRIGHT>;
        static thread_local int configured_dev[64] = { 0 };
        int dev = 0;
        CK(cudaGetDevice(&dev));
        if (dev < 64 && !configured_dev[dev]) {
            CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, HessTile::SMEM_BYTES));
            configured_dev[dev] = 1;
        }
        k<<<(unsigned)nblocks, HessTile::NT, HessTile::SMEM_BYTES, st>>>(p);
        g_launches++;
        CK(cudaGetLastError());
    }
    if (nshell > 0) {
        hessian_eigen_shell_kernel<MODE><<<(unsigned)((nshell + 127) / 128), 128, 0, st>>>(p);
        g_launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

// the coordinates 0, 1, n-2, n-1 clipped to [lo, hi) without duplicates
int face_list(int n, int lo, int hi, int* out)
{
    const int cand[4] = { 0, 1, n - 2, n - 1 };
    int k = 0;
    for (int c : cand) {
        if (c < lo || c >= hi) continue;
        bool dup = false;
        for (int q = 0; q < k; ++q) dup |= out[q] == c;
        if (!dup) out[k++] = c;
    }
    for (int q = k; q < 4; ++q) out[q] = 0;
    return k;
}

// mode 0 / 1: vesselness update of scale si; mode 2: dump the six second differences into D
int launch_voxel_core(frangi_gpu* H, Slab& s, const ScalePlan& sp, int si, float* const* D, VoxelParams& p)
{
    p.tmap = s.tmF;
    p.f = make_fview(H, s);
    p.J = s.dJ; p.Vx = s.dVx; p.Vy = s.dVy; p.Vz = s.dVz;
    p.scale_idx = s.dScale; p.dir = s.dDir; p.voxels = s.voxels;
    for (int k = 0; k < 6; ++k) p.D[k] = D ? D[k] : nullptr;
    p.z_begin = s.zb; p.nz = s.ze - s.zb;
    p.ntx = (H->w - 4 + HessTile::TX - 1) / HessTile::TX;    // tiles cover x in [2, w-2): the x faces belong to the shell
    p.nty = (H->h + HessTile::TY - 1) / HessTile::TY;
    // enough CTAs for a few waves of 148 SMs x 2 resident CTAs; each z chunk re-stages 4 planes
    const long long tiles = std::max<long long>(1, (long long)p.ntx * p.nty);   // (w < 5: no tile at all, see below)
    long long nzc = std::max<long long>(1, std::min<long long>((1184 + tiles - 1) / tiles, (p.nz + 7) / 8));
    // a tall slab with plenty of tiles: a few chunks of >= 128 planes shorten the last, partly filled wave of CTAs
    // (2048 x 2048 x 512: 1 -> 3 chunks, K3 35.1 -> 34.7 ms; shorter chunks cost more in ring fills than they balance)
    nzc = std::max(nzc, std::min<long long>((6000 + tiles - 1) / tiles, p.nz / 128));
    p.zchunk = (int)((p.nz + nzc - 1) / nzc);
    nzc = (p.nz + p.zchunk - 1) / p.zchunk;
    p.scale = si; p.last_scale = si == (int)H->scales.size() - 1;
    p.vec_ok = (H->w % 2 == 0);
    p.minmax = s.dMinMax;
    p.k = make_consts(H, sp.sigma2);
    // the two-voxel shell next to the volume faces goes to the shell kernel
    p.nxf = face_list(H->w, 0, H->w, p.xf);
    p.nyf = face_list(H->h, 0, H->h, p.yf);
    p.nzf = face_list(H->l, s.zb, s.ze, p.zf);
    p.n_zface = (long long)H->w * H->h * p.nzf;
    p.n_yface = (long long)H->w * p.nyf * p.nz;
    p.n_xface = (long long)p.nxf * H->h * p.nz;
    long long nblocks = tiles * nzc;
    if (H->w < 5 || H->h < 5 || H->l < 5) nblocks = 0;   // volumes this thin go to the shell kernel entirely
    if (nblocks > 0) { p.nzf = 0; p.n_zface = 0; }       // the tile kernels handle the z faces themselves
    long long nshell = p.n_zface + p.n_yface + p.n_xface;
    if (nblocks > 0x7fffffffLL || (nshell + 127) / 128 > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "grid too large");
    if (D) return launch_voxel_t<2>(p, nblocks, nshell, s.s_main);
    if (si == 0) return launch_voxel_t<0>(p, nblocks, nshell, s.s_main);
    if (H->blackwhite) return launch_voxel_t<1>(p, nblocks, nshell, s.s_main);
#ifdef K3C_V1
    if (p.zchunk >= (1 << 21)) return launch_voxel_t<1, true>(p, nblocks, nshell, s.s_main);   // packed z offset has 21 bits
#else
    if (s.voxels >= (1LL << 32)) return launch_voxel_t<1, true>(p, nblocks, nshell, s.s_main);  // survivors carry 32-bit voxel offsets
#endif
    // later scale of a bright-ridge run: the compacting kernel (its own 124 x 8 tiling), then the shell
    if (nblocks > 0) {
        p.ntx = (H->w - 2 - HessTileC::X_FIRST + HessTileC::TX - 1) / HessTileC::TX;   // tiles cover x up to w-3
        // a last tile column that would hold only a few voxel columns costs a full tile's work per warp (with 120-column
        // tiles a 2048-wide volume had 6 columns in its 18th tile column: 1 / 18 of the launch): those columns go to the
        // shell launch below instead
        const int rem = H->w - 2 - HessTileC::X_FIRST - (p.ntx - 1) * HessTileC::TX;
        if (K3C_SHELL_REM && p.ntx > 1 && rem <= SHELL_EXTRA_X) {
            p.ntx -= 1;
            for (int k = 0; k < rem; ++k) p.xf[p.nxf++] = H->w - 2 - rem + k;
            p.n_xface = (long long)p.nxf * H->h * p.nz;
            nshell = p.n_zface + p.n_yface + p.n_xface;
            if ((nshell + 127) / 128 > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "grid too large");
        }
        p.nty = (H->h + HessTileC::TY - 1) / HessTileC::TY;
        nblocks = (long long)p.ntx * p.nty * nzc;
        if (nblocks > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "grid too large");
        constexpr int smem = HessQueue::SMEM_BYTES;
        p.tmap = s.tmFc;
        static thread_local int configured_dev[64] = { 0 };
        int dev = 0;
        CK(cudaGetDevice(&dev));
        if (dev < 64 && !configured_dev[dev]) {
            CK(cudaFuncSetAttribute(hessian_eigen_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured_dev[dev] = 1;
        }
        hessian_eigen_compact_kernel<<<(unsigned)nblocks, HessTileC::NT, smem, s.s_main>>>(p);
        g_launches++;
        CK(cudaGetLastError());
    }
    return launch_voxel_t<1>(p, 0, nshell, s.s_main);
}

// The Hessian / eigen stage of scale si on slab s; with FRANGI_GPU_FLAG_REFERENCE_DIRECTION followed by the pass that
// rewrites the direction of every voxel this scale has just won with the reference's own solver (K9).
int launch_voxel(frangi_gpu* H, Slab& s, const ScalePlan& sp, int si, float* const* D = nullptr)
{
    VoxelParams p;
    RC(launch_voxel_core(H, s, sp, si, D, p));
    if (!D && (H->flags & FRANGI_GPU_FLAG_REFERENCE_DIRECTION) && s.voxels > 0) {
        const long long nb = (s.voxels + 127) / 128;
        if (nb > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "grid too large");
        reference_direction_kernel<<<(unsigned)nb, 128, 0, s.s_main>>>(p);
        g_launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

// Exchange the xy-smoothed boundary planes of every local slab with its z
// neighbours.  Slab k sends its lowest `halo` own planes down and its highest
// `halo` own planes up, and receives the matching planes into the halo regions
// of its Fxy buffer.  All calls of all local slabs sit in one NCCL group.
// Local-copy form (every slab lives in this process; used when device ids repeat or
// FRANGI_GPU_FLAG_LOCAL_HALO is set): each slab PULLS its halo planes from its neighbours'
// boundary planes with peer copies on its comm stream, after both its own and the
// neighbour's boundary smoothing.  The neighbour may not overwrite those planes for the next
// scale before the pull is done: run_pipeline makes each slab wait for its neighbours'
// ev_halo of the previous scale before it smooths again.
int exchange_halos_local(frangi_gpu* H, int halo)
{
    const size_t plane = (size_t)H->fplane;
    for (size_t k = 0; k < H->slabs.size(); ++k) {
        Slab& s = H->slabs[k];
        CK(cudaSetDevice(s.dev));
        if (k > 0) {
            Slab& nb = H->slabs[k - 1];
            const int n_recv = s.zb - std::max(s.zb - halo, 0);
            CK(cudaStreamWaitEvent(s.s_comm, nb.ev_boundary, 0));
            CK(cudaMemcpyPeerAsync(s.dFxy + (size_t)(s.zb - n_recv - s.xb) * plane, s.dev,
                                   nb.dFxy + (size_t)(nb.ze - n_recv - nb.xb) * plane, nb.dev,
                                   (size_t)n_recv * plane * sizeof(float), s.s_comm));
        }
        if (k + 1 < H->slabs.size()) {
            Slab& nb = H->slabs[k + 1];
            const int n_recv = std::min(s.ze + halo, H->l) - s.ze;
            CK(cudaStreamWaitEvent(s.s_comm, nb.ev_boundary, 0));
            CK(cudaMemcpyPeerAsync(s.dFxy + (size_t)(s.ze - s.xb) * plane, s.dev,
                                   nb.dFxy + (size_t)(nb.zb - nb.xb) * plane, nb.dev,
                                   (size_t)n_recv * plane * sizeof(float), s.s_comm));
        }
    }
    return 0;
}

int exchange_halos(frangi_gpu* H, int halo)
{
    if (H->nslabs_total == 1) return 0;
    if (H->local_halo) return exchange_halos_local(H, halo);
    auto& N = ncclx::api();
    const size_t plane = (size_t)H->fplane;
    NcclGroup group;
    RC(group.begin());
    for (auto& s : H->slabs) {
        const bool has_lo = s.index > 0, has_hi = s.index < H->nslabs_total - 1;
        if (has_lo) {
            const int n_recv = s.zb - std::max(s.zb - halo, 0);
            NK(N.Send(s.dFxy + (size_t)(s.zb - s.xb) * plane, (size_t)halo * plane, ncclx::ncclFloat32,
                      s.index - 1, s.comm, s.s_comm));
            NK(N.Recv(s.dFxy + (size_t)(s.zb - n_recv - s.xb) * plane, (size_t)n_recv * plane,
                      ncclx::ncclFloat32, s.index - 1, s.comm, s.s_comm));
        }
        if (has_hi) {
            const int n_recv = std::min(s.ze + halo, H->l) - s.ze;
            NK(N.Send(s.dFxy + (size_t)(s.ze - halo - s.xb) * plane, (size_t)halo * plane, ncclx::ncclFloat32,
                      s.index + 1, s.comm, s.s_comm));
            NK(N.Recv(s.dFxy + (size_t)(s.ze - s.xb) * plane, (size_t)n_recv * plane, ncclx::ncclFloat32,
                      s.index + 1, s.comm, s.s_comm));
        }
    }
    RC(group.end());
    return 0;
}

// The whole multi-scale pipeline on inputs already resident on the devices.
// I_own[k] = dense u8 planes of local slab k.
int run_pipeline(frangi_gpu* H, const std::vector<const uint8_t*>& I_own)
{
    const int S = (int)H->scales.size();
    const bool multi = H->nslabs_total > 1;
    const int ev_set = (int)(H->runs_recorded % H->timing_depth);
    H->runs_recorded++;
    H->last_streamed = false;
    for (size_t k = 0; k < H->slabs.size(); ++k) {
        Slab& s = H->slabs[k];
        CK(cudaSetDevice(s.dev));
        s.ev_time = s.ev_all.data() + (size_t)ev_set * (kEvPerScale * S + 2);
        // Jmin = FLT_MAX, Jmax = -FLT_MAX (frangi.cpp:176-177), set ON the device: the pinned pair hMinMax is only ever
        // the target of the result copy, so back-to-back asynchronous runs cannot pick up the previous run's result
        minmax_init_kernel<<<1, 32, 0, s.s_main>>>(s.dMinMax);
        g_launches++;
        CK(cudaEventRecord(s.ev_time[0], s.s_main));
    }
    // Multi-slab.  Two Fxy buffers alternate by scale (buffer si & 1, halo event si & 1), so a scale's xy pass and halo
    // exchange never wait for an earlier scale's z pass to have read "the" Fxy buffer.
    auto use_fxy = [&](Slab& s, int si) {
        const bool b = (si & 1) && s.dFxyB;
        s.dFxy = b ? s.dFxyB : s.dFxyA;
        s.dFxy0 = s.dFxy;
        s.tmFxy = b ? s.tmFxyB : s.tmFxyA;
    };
    auto halo_event = [&](Slab& s, int si) { return ((si & 1) && s.ev_halo2) ? s.ev_halo2 : s.ev_halo; };
    // The xy pass of scale si on every local slab -- both boundary plane groups in ONE launch, then the halo
    // exchange of those planes on the comm streams, then the interior planes while the exchange is in flight.
    auto xy_and_exchange = [&](int si) -> int {
        const ScalePlan& sp = H->scales[si];
        const int halo = sp.rz + 2;
        for (size_t k = 0; k < H->slabs.size(); ++k) {
            Slab& s = H->slabs[k];
            CK(cudaSetDevice(s.dev));
            use_fxy(s, si);
            if (H->local_halo && si >= 2) {  // neighbours have pulled the boundary planes this buffer held two scales ago
                if (k > 0) CK(cudaStreamWaitEvent(s.s_main, halo_event(H->slabs[k - 1], si), 0));
                if (k + 1 < H->slabs.size()) CK(cudaStreamWaitEvent(s.s_main, halo_event(H->slabs[k + 1], si), 0));
            }
            const int nz = s.ze - s.zb;
            if (nz <= 2 * halo) RC(launch_xy(H, s, sp, I_own[k], s.zb, s.ze));
            else RC(launch_xy_ends(H, s, sp, I_own[k], s.zb, s.ze, halo));
            CK(cudaEventRecord(s.ev_boundary, s.s_main));
            CK(cudaStreamWaitEvent(s.s_comm, s.ev_boundary, 0));
        }
        RC(exchange_halos(H, halo));
        for (size_t k = 0; k < H->slabs.size(); ++k) {
            Slab& s = H->slabs[k];
            CK(cudaSetDevice(s.dev));
            CK(cudaEventRecord(halo_event(s, si), s.s_comm));
            const int nz = s.ze - s.zb;
            if (nz > 2 * halo) RC(launch_xy(H, s, sp, I_own[k], s.zb + halo, s.ze - halo));
        }
        return 0;
    };
    for (int si = 0; si < S; ++si) {
        const ScalePlan& sp = H->scales[si];
        if (multi) {
            // Scales are software-pipelined two deep so that a halo exchange never waits in the open: the run starts
            // with the xy passes and exchanges of scales 0 AND 1 (the first exchange hides behind the second xy pass),
            // and the xy pass + exchange of scale si+2 are issued right after the z pass of scale si -- the last
            // reader of that Fxy buffer -- and BEFORE its Hessian / eigen stage (which reads neither Fxy nor the
            // halo).  Per scale and slab: wait for the halo, ONE z pass over all F planes, [xy + exchange of scale
            // si+2], ONE Hessian / eigen launch: no plane is smoothed or staged twice.  The single F buffer is safe:
            // the next z pass follows this scale's Hessian / eigen stage on the same stream.  A receive into a
            // buffer's halo planes cannot start before the z pass that last read them is done: the comm stream waits
            // for the boundary event, which follows that z pass on the main stream.
            if (si == 0) {
                RC(xy_and_exchange(0));
                if (S > 1) RC(xy_and_exchange(1));
            }
            for (auto& s : H->slabs) {
                CK(cudaSetDevice(s.dev));
                cudaEvent_t* ev = s.ev_time + kEvPerScale * si;
                CK(cudaEventRecord(ev[1], s.s_main));
                CK(cudaStreamWaitEvent(s.s_main, halo_event(s, si), 0));
                CK(cudaEventRecord(ev[2], s.s_main));
                use_fxy(s, si);
                RC(launch_z(H, s, sp));
                CK(cudaEventRecord(ev[3], s.s_main));
            }
            if (si + 2 < S) RC(xy_and_exchange(si + 2));
            for (auto& s : H->slabs) {
                CK(cudaSetDevice(s.dev));
                cudaEvent_t* ev = s.ev_time + kEvPerScale * si;
                CK(cudaEventRecord(ev[4], s.s_main));
                RC(launch_voxel(H, s, sp, si));
                CK(cudaEventRecord(ev[5], s.s_main));
            }
        } else if (H->overlap) {
            // One slab, overlapped schedule.  The z pass is HBM-bound and the stages around it are issue-bound, so
            // it runs on a low-priority stream NEXT to them: K2(si) beside K1(si+1), K2(si+1) beside K3(si).
            // Two Fxy / F buffer pairs alternate by scale.  Main stream: K1(0) | K1(1) K3(0) | K1(2) K3(1) | ... ;
            // aux stream: K2(si) after K1(si).  Buffer reuse is ordered by the main stream itself: K1(si+1) follows
            // K3(si-1), which waited for K2(si-1) (last reader of that Fxy), and K2(si) follows K1(si), which
            // follows K3(si-2) (last reader of that F).
            Slab& s = H->slabs[0];
            CK(cudaSetDevice(s.dev));
            cudaEvent_t* eo = s.ev_ov.data() + 6 * si;          // K1 start/end, K2 start/end, K3 start/end
            auto buf = [&](int q) {
                Slab v = s;
                if (q & 1) { v.dFxy = s.dFxyB; v.dFxy0 = s.dFxyB; v.dF = s.dFB; v.tmFxy = s.tmFxyB; v.tmF = s.tmFB; v.tmFc = s.tmFcB; }
                return v;
            };
            if (si == 0) {
                Slab v = buf(0);
                CK(cudaEventRecord(eo[0], s.s_main));
                RC(launch_xy(H, v, sp, I_own[0], s.zb, s.ze));
                CK(cudaEventRecord(eo[1], s.s_main));
            }
            {
                Slab v = buf(si);
                CK(cudaStreamWaitEvent(s.s_aux, eo[1], 0));
                CK(cudaEventRecord(eo[2], s.s_aux));
                RC(launch_z(H, v, sp, s.s_aux));
                CK(cudaEventRecord(eo[3], s.s_aux));
            }
            if (si + 1 < S) {
                Slab v = buf(si + 1);
                CK(cudaEventRecord(eo[6], s.s_main));
                RC(launch_xy(H, v, H->scales[si + 1], I_own[0], s.zb, s.ze));
                CK(cudaEventRecord(eo[7], s.s_main));
            }
            {
                Slab v = buf(si);
                CK(cudaStreamWaitEvent(s.s_main, eo[3], 0));
                CK(cudaEventRecord(eo[4], s.s_main));
                RC(launch_voxel(H, v, sp, si));
                CK(cudaEventRecord(eo[5], s.s_main));
            }
        } else {
            Slab& s = H->slabs[0];
            CK(cudaSetDevice(s.dev));
            cudaEvent_t* ev = s.ev_time + kEvPerScale * si;
            RC(launch_xy(H, s, sp, I_own[0], s.zb, s.ze));
            CK(cudaEventRecord(ev[1], s.s_main));
            CK(cudaEventRecord(ev[2], s.s_main));
            RC(launch_z(H, s, sp));
            CK(cudaEventRecord(ev[3], s.s_main));
            CK(cudaEventRecord(ev[4], s.s_main));
            RC(launch_voxel(H, s, sp, si));
            CK(cudaEventRecord(ev[5], s.s_main));
        }
    }
    if (!multi && H->overlap) CK(cudaEventRecord(H->slabs[0].ev_time[kEvPerScale * S], H->slabs[0].s_main));
    // global Jmin / Jmax across slabs, then the 8-bit normalisation
    if (multi && H->local_halo) {
        // gather every slab's pair on slab 0, reduce there, hand the result back
        Slab& s0 = H->slabs[0];
        const int n = (int)H->slabs.size();
        for (int k = 0; k < n; ++k) {
            Slab& s = H->slabs[k];
            CK(cudaSetDevice(s.dev));
            CK(cudaEventRecord(s.ev_done, s.s_main));
        }
        CK(cudaSetDevice(s0.dev));
        for (int k = 0; k < n; ++k) {
            Slab& s = H->slabs[k];
            CK(cudaStreamWaitEvent(s0.s_main, s.ev_done, 0));
            CK(cudaMemcpyPeerAsync(s0.dGather + 2 * k, s0.dev, s.dMinMax, s.dev, 2 * sizeof(int), s0.s_main));
        }
        minmax_reduce_kernel<<<1, 32, 0, s0.s_main>>>(s0.dGather, n, s0.dMinMax);
        g_launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(s0.ev_done, s0.s_main));
        for (int k = 1; k < n; ++k) {
            Slab& s = H->slabs[k];
            CK(cudaSetDevice(s.dev));
            CK(cudaStreamWaitEvent(s.s_main, s0.ev_done, 0));
            CK(cudaMemcpyPeerAsync(s.dMinMax, s.dev, s0.dMinMax, s0.dev, 2 * sizeof(int), s.s_main));
        }
    } else if (multi) {
        auto& N = ncclx::api();
        NcclGroup group;
        RC(group.begin());
        for (auto& s : H->slabs) {
            NK(N.AllReduce(s.dMinMax, s.dMinMax, 1, ncclx::ncclInt32, ncclx::ncclMin, s.comm, s.s_main));
            NK(N.AllReduce(s.dMinMax + 1, s.dMinMax + 1, 1, ncclx::ncclInt32, ncclx::ncclMax, s.comm, s.s_main));
        }
        RC(group.end());
    }
    for (auto& s : H->slabs) {
        CK(cudaSetDevice(s.dev));
        const int nb = (int)std::min<long long>((s.voxels + 255) / 256, 148 * 16);
        j_to_j8_kernel<<<nb, 256, 0, s.s_main>>>(s.dJ, s.dJ8, s.voxels, s.dMinMax);
        g_launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(s.hMinMax, s.dMinMax, 2 * sizeof(int), cudaMemcpyDeviceToHost, s.s_main));
        CK(cudaEventRecord(s.ev_time[kEvPerScale * S + 1], s.s_main));
    }
    H->ran = true;
    return 0;
}

int sync_all(frangi_gpu* H)
{
    for (auto& s : H->slabs) {
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.s_main));
        CK(cudaStreamSynchronize(s.s_comm));
        CK(cudaStreamSynchronize(s.s_h2d));
        CK(cudaStreamSynchronize(s.s_d2h));
    }
    return 0;
}

int collect(frangi_gpu* H, float* Jmin, float* Jmax)
{
    RC(sync_all(H));
    Slab& s0 = H->slabs[0];
    float lo, hi;
    std::memcpy(&lo, &s0.hMinMax[0], 4);
    std::memcpy(&hi, &s0.hMinMax[1], 4);
    if (Jmin) *Jmin = lo;
    if (Jmax) *Jmax = hi;
    // per-class device times of slab 0: mean over the recorded runs (at most timing_depth)
    const int S = (int)H->scales.size();
    std::memset(H->last_ms, 0, sizeof H->last_ms);
    CK(cudaSetDevice(s0.dev));
    int nsets = (int)std::min<long long>(H->runs_recorded, H->timing_depth);
    if (H->last_streamed) {
        float t;
        CK(cudaEventElapsedTime(&t, s0.ev_all[0], s0.ev_all[kEvPerScale * S + 1]));
        H->last_ms[5] = t;
        nsets = 0;
    }
    for (int r = 0; r < nsets; ++r) {
        const cudaEvent_t* ev = s0.ev_all.data() + (size_t)r * (kEvPerScale * S + 2);
        float t;
        for (int si = 0; si < S && H->overlap; ++si) {
            // overlapped schedule: every kernel class has its own start / end pair (the classes run side by side, their
            // sum exceeds the step); the per-kernel events are those of the LAST run
            const cudaEvent_t* e = s0.ev_ov.data() + 6 * si;
            CK(cudaEventElapsedTime(&t, e[0], e[1])); H->last_ms[0] += t;
            CK(cudaEventElapsedTime(&t, e[2], e[3])); H->last_ms[1] += t;
            CK(cudaEventElapsedTime(&t, e[4], e[5])); H->last_ms[2] += t;
        }
        for (int si = 0; si < S && !H->overlap; ++si) {
            const cudaEvent_t* e = ev + kEvPerScale * si;     // e[0] = end of the previous scale (or run start)
            CK(cudaEventElapsedTime(&t, e[0], e[1])); H->last_ms[0] += t;   // xy smoothing (multi-slab: of the first scale)
            CK(cudaEventElapsedTime(&t, e[1], e[2])); H->last_ms[4] += t;   // exposed halo wait
            CK(cudaEventElapsedTime(&t, e[2], e[3])); H->last_ms[1] += t;   // z smoothing
            CK(cudaEventElapsedTime(&t, e[3], e[4])); H->last_ms[0] += t;   // multi-slab: xy smoothing of the NEXT scale, issued early
            CK(cudaEventElapsedTime(&t, e[4], e[5])); H->last_ms[2] += t;   // Hessian / eigen
        }
        CK(cudaEventElapsedTime(&t, ev[kEvPerScale * S], ev[kEvPerScale * S + 1])); H->last_ms[3] += t;
        CK(cudaEventElapsedTime(&t, ev[0], ev[kEvPerScale * S + 1])); H->last_ms[5] += t;
    }
    if (nsets > 0)
        for (int i = 0; i < 6; ++i) H->last_ms[i] /= (float)nsets;
    return 0;
}

int download_slab(frangi_gpu* H, Slab& s, long long off, float* J, uint8_t* Vx, uint8_t* Vy, uint8_t* Vz,
                  uint8_t* J8, uint8_t* sc, float* dir, long long total_voxels)
{
    CK(cudaSetDevice(s.dev));
    const size_t n = (size_t)s.voxels;
    if (J) CK(cudaMemcpyAsync(J + off, s.dJ, n * 4, cudaMemcpyDeviceToHost, s.s_main));
    if (Vx) CK(cudaMemcpyAsync(Vx + off, s.dVx, n, cudaMemcpyDeviceToHost, s.s_main));
    if (Vy) CK(cudaMemcpyAsync(Vy + off, s.dVy, n, cudaMemcpyDeviceToHost, s.s_main));
    if (Vz) CK(cudaMemcpyAsync(Vz + off, s.dVz, n, cudaMemcpyDeviceToHost, s.s_main));
    if (J8) CK(cudaMemcpyAsync(J8 + off, s.dJ8, n, cudaMemcpyDeviceToHost, s.s_main));
    if (sc) {
        if (!s.dScale) return fail(FRANGI_GPU_ESTATE, "scale index not kept: create with FRANGI_GPU_FLAG_SCALE_IDX");
        CK(cudaMemcpyAsync(sc + off, s.dScale, n, cudaMemcpyDeviceToHost, s.s_main));
    }
    if (dir) {
        if (!s.dDir) return fail(FRANGI_GPU_ESTATE, "float direction not kept: create with FRANGI_GPU_FLAG_DIR_F32");
        for (int c = 0; c < 3; ++c)
            CK(cudaMemcpyAsync(dir + (size_t)c * total_voxels + off, s.dDir + (size_t)c * n, n * 4,
                               cudaMemcpyDeviceToHost, s.s_main));
    }
    return 0;
}

long long local_voxels(frangi_gpu* H)
{
    long long n = 0;
    for (auto& s : H->slabs) n += s.voxels;
    return n;
}

}  // namespace

// frangi_gpu_run on one device, pipelined over z chunks: the chunk's input planes go up on
// one stream, its three scales run on the main stream, its finished rows of J / V come down
// on a third stream while the next chunk is being computed.  A chunk is a view of the slab
// (own planes [cz0, cz1), smoothed planes computed locally from the resident input, so no
// exchange); results are bit-identical to the one-piece run.  Host<->device copies overlap
// the kernels only when the host buffers are pinned (frangi_gpu_host_alloc).
int run_streamed(frangi_gpu* H, const uint8_t* I_host, float* J, uint8_t* Vx, uint8_t* Vy, uint8_t* Vz, uint8_t* J8,
                 uint8_t* sc, float* dir, int ch)
{
    Slab& s = H->slabs[0];
    const int S = (int)H->scales.size();
    const int nz = s.ze - s.zb;
    const int nch = (nz + ch - 1) / ch;
    const size_t wh = (size_t)H->w * H->h;
    if (sc && !s.dScale) return fail(FRANGI_GPU_ESTATE, "scale index not kept: create with FRANGI_GPU_FLAG_SCALE_IDX");
    if (dir && !s.dDir) return fail(FRANGI_GPU_ESTATE, "float direction not kept: create with FRANGI_GPU_FLAG_DIR_F32");
    CK(cudaSetDevice(s.dev));
    // Ordinary (pageable) host buffers -- what the unmodified call site passes, Advantra_plugin.cpp:2490-2494 -- go
    // through the pinned staging ring (host_stager.h); pinned buffers are copied directly.
    const void* outs[] = { J, Vx, Vy, Vz, J8, sc, dir };
    bool staged_out = false;
    for (const void* q : outs) staged_out = staged_out || (q && H->stager.pageable(q));
    const bool staged_in = H->stager.pageable(I_host);
    if (staged_in || staged_out) {
        // the host-side memcpy must keep up with the DMA (~54 GB/s on PCIe 5): one core moves 4-5 GB/s
        const unsigned hw = std::thread::hardware_concurrency();
        const int nthreads = (int)std::max(2u, std::min(16u, hw > 2 ? hw - 2 : 1u));
        if (!H->stager.start(s.dev, 24, nthreads)) return fail(FRANGI_GPU_ENOMEM, "pinned staging ring: allocation failed");
    }
    auto down = [&](void* dst, const void* src, size_t bytes, cudaStream_t st) -> cudaError_t {
        if (H->stager.pageable(dst)) return H->stager.d2h(dst, src, bytes, st);
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
    };
    while ((int)s.ev_chunk.size() < 2 * nch) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s.ev_chunk.push_back(e);
    }
    // a streamed run records only the first / last event of set 0; the per-class sets start afresh with the next
    // resident run, so collect() never reads an event that was not recorded
    H->runs_recorded = 0;
    H->last_streamed = true;
    s.ev_time = s.ev_all.data();
    minmax_init_kernel<<<1, 32, 0, s.s_main>>>(s.dMinMax);    // see run_pipeline
    g_launches++;
    CK(cudaEventRecord(s.ev_time[0], s.s_main));
    int uploaded = s.zb;                      // planes [zb, uploaded) of the input are on their way
    for (int c = 0; c < nch; ++c) {
        const int cz0 = s.zb + c * ch, cz1 = std::min(cz0 + ch, s.ze);
        Slab v = s;                           // shallow view: same buffers, this chunk's planes
        v.zb = cz0; v.ze = cz1;
        v.fb = std::max(cz0 - 2, s.zb); v.fe = std::min(cz1 + 2, s.ze);
        const size_t off = (size_t)(cz0 - s.zb) * wh;
        v.dJ = s.dJ + off; v.dVx = s.dVx + off; v.dVy = s.dVy + off; v.dVz = s.dVz + off;
        if (s.dScale) v.dScale = s.dScale + off;
        if (s.dDir) v.dDir = s.dDir + off;
        const uint8_t* I_view = s.dI + off;
        const int need_hi = std::min(v.fe + H->rz_max, s.ze);
        if (need_hi > uploaded) {
            uint8_t* dst = s.dI + (size_t)(uploaded - s.zb) * wh;
            const uint8_t* src = I_host + (size_t)(uploaded - s.zb) * wh;
            const size_t nb = (size_t)(need_hi - uploaded) * wh;
            if (staged_in) CK(H->stager.h2d(dst, src, nb, s.s_h2d));
            else CK(cudaMemcpyAsync(dst, src, nb, cudaMemcpyHostToDevice, s.s_h2d));
            uploaded = need_hi;
        }
        CK(cudaEventRecord(s.ev_chunk[2 * c], s.s_h2d));
        CK(cudaStreamWaitEvent(s.s_main, s.ev_chunk[2 * c], 0));
        for (int si = 0; si < S; ++si) {
            const ScalePlan& sp = H->scales[si];
            v.xb = std::max(v.fb - sp.rz, s.zb); v.xe = std::min(v.fe + sp.rz, s.ze);
            RC(launch_xy(H, v, sp, I_view, v.xb, v.xe));
            RC(launch_z(H, v, sp));
            RC(launch_voxel(H, v, sp, si));
        }
        CK(cudaEventRecord(s.ev_chunk[2 * c + 1], s.s_main));
        CK(cudaStreamWaitEvent(s.s_d2h, s.ev_chunk[2 * c + 1], 0));
        const size_t n = (size_t)(cz1 - cz0) * wh;
        if (J) CK(down(J + off, v.dJ, n * 4, s.s_d2h));
        if (Vx) CK(down(Vx + off, v.dVx, n, s.s_d2h));
        if (Vy) CK(down(Vy + off, v.dVy, n, s.s_d2h));
        if (Vz) CK(down(Vz + off, v.dVz, n, s.s_d2h));
        if (sc) CK(down(sc + off, v.dScale, n, s.s_d2h));
        if (dir)
            for (int k = 0; k < 3; ++k)
                CK(down(dir + (size_t)k * s.voxels + off, v.dDir + (size_t)k * s.voxels, n * 4, s.s_d2h));
    }
    // the 8-bit map needs the global min / max: after the last chunk
    const int nb = (int)std::min<long long>((s.voxels + 255) / 256, 148 * 16);
    j_to_j8_kernel<<<nb, 256, 0, s.s_main>>>(s.dJ, s.dJ8, s.voxels, s.dMinMax);
    g_launches++;
    CK(cudaGetLastError());
    if (J8) CK(down(J8, s.dJ8, (size_t)s.voxels, s.s_main));
    CK(cudaMemcpyAsync(s.hMinMax, s.dMinMax, 2 * sizeof(int), cudaMemcpyDeviceToHost, s.s_main));
    CK(cudaEventRecord(s.ev_time[kEvPerScale * S + 1], s.s_main));
    if (staged_out) H->stager.drain();        // the workers have moved every piece into the caller's buffers
    H->ran = true;
    return 0;
}

// =============================== C-ABI ========================================

FRANGI_API const char* frangi_gpu_last_error(void) { return g_err.c_str(); }
FRANGI_API const char* frangi_gpu_version(void) { return "frangi-b200 0.1 (sm_100a)"; }
FRANGI_API uint64_t frangi_gpu_launch_count(void) { return g_launches.load(); }

FRANGI_API int frangi_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

FRANGI_API void* frangi_gpu_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        fail(FRANGI_GPU_ENOMEM, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

FRANGI_API void frangi_gpu_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

FRANGI_API int frangi_gpu_nccl_unique_id(void* out128)
{
    auto& N = ncclx::api();
    if (!N.ok) return fail(FRANGI_GPU_ENCCL, "libnccl.so.2 not found or incomplete");
    ncclx::ncclUniqueId id;
    NK(N.GetUniqueId(&id));
    std::memcpy(out128, &id, 128);
    return 0;
}

FRANGI_API void frangi_gpu_destroy(frangi_gpu_t* H)
{
    if (!H) return;
    for (auto& s : H->slabs) free_slab(s);
    delete H;
}

FRANGI_API int frangi_gpu_create(frangi_gpu_t** out, const float* sigmas, int nsig, float zdist, float alpha,
                                 float beta, float C, int blackwhite, int w, int h, int l,
                                 const int* device_ids, int ndev, unsigned flags)
{
    if (!out) return fail(FRANGI_GPU_EINVAL, "out is NULL");
    *out = nullptr;
    if (ndev < 1) return fail(FRANGI_GPU_EINVAL, "ndev must be >= 1");
    frangi_gpu* H = new (std::nothrow) frangi_gpu();
    if (!H) return fail(FRANGI_GPU_ENOMEM, "out of host memory");
    int rc = plan_common(H, sigmas, nsig, zdist, alpha, beta, C, blackwhite, w, h, l, flags);
    if (rc) { delete H; return rc; }
    // a neighbour must be able to supply a whole halo: slab thickness >= rz_max + 2
    const int min_thick = H->rz_max + 2;
    int nuse = std::min(ndev, std::max(1, l / min_thick));
    if (nuse < ndev) {       // not an error (the result is the same), but the caller asked for more devices than are used
        char buf[160];
        snprintf(buf, sizeof buf, "%d of the %d devices are used: a slab must hold at least %d planes and the volume has %d; ",
                 nuse, ndev, min_thick, l);
        H->warnings += buf;
    }
    H->nslabs_total = nuse;
    H->slabs.resize(nuse);
    // one slab, several scales, TMA z pass available for every radius: the overlapped schedule (run_pipeline)
    H->overlap = (flags & FRANGI_GPU_FLAG_OVERLAP_Z) && nuse == 1 && nsig >= 2 && Z_TMA;
    for (const auto& sp : H->scales) H->overlap = H->overlap && sp.rz_t <= 12;
    std::vector<int> devs(nuse);
    for (int k = 0; k < nuse; ++k) devs[k] = device_ids ? device_ids[k] : k;
    for (int k = 0; k < nuse && !rc; ++k) {
        rc = check_device(devs[k]);
        if (!rc) {
            const int zb = (int)((long long)l * k / nuse), ze = (int)((long long)l * (k + 1) / nuse);
            rc = alloc_slab(H, H->slabs[k], devs[k], k, zb, ze);
        }
    }
    bool repeated = false;
    for (int a = 0; a < nuse; ++a)
        for (int b = a + 1; b < nuse; ++b) repeated |= devs[a] == devs[b];
    H->local_halo = nuse > 1 && (repeated || (flags & FRANGI_GPU_FLAG_LOCAL_HALO));
    if (!rc && H->local_halo) {
        cudaError_t e = cudaSetDevice(devs[0]);
        if (e == cudaSuccess) e = cudaMalloc(&H->slabs[0].dGather, 2 * sizeof(int) * nuse);
        for (int a = 0; a < nuse && e == cudaSuccess; ++a)       // peer access where the devices differ
            for (int b = 0; b < nuse && e == cudaSuccess; ++b)
                if (devs[a] != devs[b]) {
                    e = cudaSetDevice(devs[a]);
                    if (e == cudaSuccess) {
                        int can = 0;
                        e = cudaDeviceCanAccessPeer(&can, devs[a], devs[b]);
                        if (e != cudaSuccess) break;
                        cudaError_t pe = can ? cudaDeviceEnablePeerAccess(devs[b], 0) : cudaErrorPeerAccessUnsupported;
                        if (pe == cudaErrorPeerAccessAlreadyEnabled) pe = cudaSuccess;
                        if (pe != cudaSuccess) {
                            // the halo pulls still work (staged through the host) but are far slower: say so
                            char buf[160];
                            snprintf(buf, sizeof buf, "no peer access from device %d to device %d: halo copies are staged through the host; ",
                                     devs[a], devs[b]);
                            H->warnings += buf;
                        }
                        cudaGetLastError();
                    }
                }
        if (e != cudaSuccess) rc = fail(FRANGI_GPU_ECUDA, "local halo setup: %s", cudaGetErrorString(e));
    } else if (!rc && nuse > 1) {
        auto& N = ncclx::api();
        if (!N.ok) rc = fail(FRANGI_GPU_ENCCL, "libnccl.so.2 not found or incomplete");
        else {
            std::vector<ncclx::ncclComm_t> comms(nuse);
            ncclx::ncclResult_t r = N.CommInitAll(comms.data(), nuse, devs.data());
            if (r != ncclx::ncclSuccess) rc = fail(FRANGI_GPU_ENCCL, "ncclCommInitAll failed (%d)", (int)r);
            else for (int k = 0; k < nuse; ++k) H->slabs[k].comm = comms[k];
        }
    }
    if (rc) { frangi_gpu_destroy(H); return rc; }
    *out = H;
    return 0;
}

FRANGI_API int frangi_gpu_create_slab(frangi_gpu_t** out, const float* sigmas, int nsig, float zdist, float alpha,
                                      float beta, float C, int blackwhite, int w, int h, int l, int z_begin,
                                      int z_end, int rank, int nranks, const void* nccl_unique_id, int device,
                                      unsigned flags)
{
    if (!out) return fail(FRANGI_GPU_EINVAL, "out is NULL");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(FRANGI_GPU_EINVAL, "bad rank/nranks");
    if (z_begin < 0 || z_end > l || z_end <= z_begin) return fail(FRANGI_GPU_EINVAL, "bad z range");
    if ((rank == 0) != (z_begin == 0) || (rank == nranks - 1) != (z_end == l))
        return fail(FRANGI_GPU_EINVAL, "slabs must tile [0,l) in rank order");
    frangi_gpu* H = new (std::nothrow) frangi_gpu();
    if (!H) return fail(FRANGI_GPU_ENOMEM, "out of host memory");
    int rc = plan_common(H, sigmas, nsig, zdist, alpha, beta, C, blackwhite, w, h, l, flags);
    if (!rc && nranks > 1 && (z_end - z_begin) < H->rz_max + 2)
        rc = fail(FRANGI_GPU_EINVAL, "slab of %d planes is thinner than the halo %d", z_end - z_begin, H->rz_max + 2);
    if (!rc) rc = check_device(device);
    if (!rc) {
        H->nslabs_total = nranks;
        H->slabs.resize(1);
        rc = alloc_slab(H, H->slabs[0], device, rank, z_begin, z_end);
    }
    if (!rc && nranks > 1) {
        auto& N = ncclx::api();
        if (!N.ok) rc = fail(FRANGI_GPU_ENCCL, "libnccl.so.2 not found or incomplete");
        else if (!nccl_unique_id) rc = fail(FRANGI_GPU_EINVAL, "nccl_unique_id is NULL");
        else {
            ncclx::ncclUniqueId id;
            std::memcpy(&id, nccl_unique_id, 128);
            ncclx::ncclResult_t r = N.CommInitRank(&H->slabs[0].comm, nranks, id, rank);
            if (r != ncclx::ncclSuccess) rc = fail(FRANGI_GPU_ENCCL, "ncclCommInitRank failed (%d)", (int)r);
        }
    }
    if (rc) { frangi_gpu_destroy(H); return rc; }
    *out = H;
    return 0;
}

FRANGI_API int frangi_gpu_upload(frangi_gpu_t* H, const uint8_t* I_host)
{
    if (!H || !I_host) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    long long off = 0;
    for (auto& s : H->slabs) {
        CK(cudaSetDevice(s.dev));
        CK(cudaMemcpyAsync(s.dI, I_host + off, (size_t)s.voxels, cudaMemcpyHostToDevice, s.s_main));
        off += s.voxels;
    }
    return 0;
}

FRANGI_API int frangi_gpu_run_resident(frangi_gpu_t* H, float* Jmin, float* Jmax)
{
    if (!H) return fail(FRANGI_GPU_EINVAL, "NULL handle");
    std::vector<const uint8_t*> in;
    for (auto& s : H->slabs) in.push_back(s.dI);
    RC(run_pipeline(H, in));
    if (Jmin || Jmax) RC(collect(H, Jmin, Jmax));
    return 0;
}

FRANGI_API int frangi_gpu_run_device(frangi_gpu_t* H, const uint8_t* I_dev, float* Jmin, float* Jmax)
{
    if (!H || !I_dev) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    if (H->slabs.size() != 1) return fail(FRANGI_GPU_ESTATE, "run_device needs a single-device handle");
    std::vector<const uint8_t*> in(1, I_dev);
    RC(run_pipeline(H, in));
    if (Jmin || Jmax) RC(collect(H, Jmin, Jmax));
    return 0;
}

FRANGI_API int frangi_gpu_sync(frangi_gpu_t* H)
{
    if (!H) return fail(FRANGI_GPU_EINVAL, "NULL handle");
    return collect(H, nullptr, nullptr);
}

FRANGI_API int frangi_gpu_download(frangi_gpu_t* H, float* J, uint8_t* Vx, uint8_t* Vy, uint8_t* Vz, uint8_t* J8,
                                   uint8_t* sc, float* dir)
{
    if (!H) return fail(FRANGI_GPU_EINVAL, "NULL handle");
    if (!H->ran) return fail(FRANGI_GPU_ESTATE, "nothing has been run on this handle");
    const long long total = local_voxels(H);
    long long off = 0;
    for (auto& s : H->slabs) {
        RC(download_slab(H, s, off, J, Vx, Vy, Vz, J8, sc, dir, total));
        off += s.voxels;
    }
    return sync_all(H);
}

FRANGI_API int frangi_gpu_run(frangi_gpu_t* H, const uint8_t* I_host, float* J, float* Jmin, float* Jmax,
                              uint8_t* Vx, uint8_t* Vy, uint8_t* Vz, uint8_t* J8, uint8_t* sc, float* dir)
{
    if (!H || !I_host) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    if (H->nslabs_total == 1) {
        // one device: pipeline upload / kernels / download over z chunks when the volume is deep enough
        const Slab& s0 = H->slabs[0];
        const int halo = H->rz_max + 2;
        int ch = H->stream_chunk < 0 ? std::max(32, 2 * halo) : H->stream_chunk;
        if (ch > 0 && ch < halo) ch = halo;
        if (ch > 0 && (s0.ze - s0.zb) >= 2 * ch) {
            RC(run_streamed(H, I_host, J, Vx, Vy, Vz, J8, sc, dir, ch));
            return collect(H, Jmin, Jmax);
        }
    }
    RC(frangi_gpu_upload(H, I_host));
    std::vector<const uint8_t*> in;
    for (auto& s : H->slabs) in.push_back(s.dI);
    RC(run_pipeline(H, in));
    const long long total = local_voxels(H);
    long long off = 0;
    for (auto& s : H->slabs) {
        RC(download_slab(H, s, off, J, Vx, Vy, Vz, J8, sc, dir, total));
        off += s.voxels;
    }
    return collect(H, Jmin, Jmax);
}

FRANGI_API int frangi_gpu_set_stream_chunk(frangi_gpu_t* H, int planes)
{
    if (!H || planes < -1) return fail(FRANGI_GPU_EINVAL, "bad argument");
    H->stream_chunk = planes;
    return 0;
}

FRANGI_API int frangi_gpu_device_outputs(frangi_gpu_t* H, int slab, frangi_gpu_outputs_t* out)
{
    if (!H || !out || slab < 0 || slab >= (int)H->slabs.size()) return fail(FRANGI_GPU_EINVAL, "bad argument");
    Slab& s = H->slabs[slab];
    out->J = s.dJ; out->Vx = s.dVx; out->Vy = s.dVy; out->Vz = s.dVz;
    out->scale_idx = s.dScale; out->dir_xyz = s.dDir; out->voxels = s.voxels;
    return 0;
}

FRANGI_API int frangi_gpu_last_timings(frangi_gpu_t* H, float* ms, int n)
{
    if (!H || !ms) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    for (int i = 0; i < n && i < 8; ++i) ms[i] = H->last_ms[i];
    return 0;
}

FRANGI_API int frangi_gpu_timing_depth(frangi_gpu_t* H, int depth)
{
    if (!H || depth < 1 || depth > 4096) return fail(FRANGI_GPU_EINVAL, "timing depth must be 1..4096");
    RC(sync_all(H));
    const size_t per = kEvPerScale * H->scales.size() + 2;
    for (auto& s : H->slabs) {
        CK(cudaSetDevice(s.dev));
        while (s.ev_all.size() < per * (size_t)depth) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            s.ev_all.push_back(e);
        }
        s.ev_time = s.ev_all.data();
    }
    H->timing_depth = depth;
    H->runs_recorded = 0;
    return 0;
}

FRANGI_API int frangi_gpu_slab_count(frangi_gpu_t* H) { return H ? (int)H->slabs.size() : 0; }

FRANGI_API const char* frangi_gpu_warnings(frangi_gpu_t* H) { return H ? H->warnings.c_str() : ""; }

FRANGI_API void* frangi_gpu_stream(frangi_gpu_t* H, int slab)
{
    if (!H || slab < 0 || slab >= (int)H->slabs.size()) return nullptr;
    return (void*)H->slabs[slab].s_main;
}

// ---- stage entry points --------------------------------------------------------

namespace {
int stage_smooth(frangi_gpu_t** Hout, const uint8_t* I_host, int w, int h, int l, float sigma, float zdist,
                 int device, unsigned flags)
{
    int dev = device;
    RC(frangi_gpu_create(Hout, &sigma, 1, zdist, .5f, .5f, 500.f, 0, w, h, l, &dev, 1, flags));
    frangi_gpu* H = *Hout;
    Slab& s = H->slabs[0];
    RC(frangi_gpu_upload(H, I_host));
    RC(launch_xy(H, s, H->scales[0], s.dI, s.zb, s.ze));
    RC(launch_z(H, s, H->scales[0]));
    return 0;
}
}  // namespace

FRANGI_API int frangi_gpu_imgaussian(const uint8_t* I_host, int w, int h, int l, float sigma, float zdist,
                                     float* F_host, int device, unsigned flags)
{
    if (!I_host || !F_host) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    frangi_gpu* H = nullptr;
    int rc = stage_smooth(&H, I_host, w, h, l, sigma, zdist, device, flags);
    if (!rc) {
        Slab& s = H->slabs[0];
        cudaError_t e = cudaMemcpy2DAsync(F_host, (size_t)w * 4, s.dF, (size_t)H->fpitch * 4, (size_t)w * 4,
                                          (size_t)h * l, cudaMemcpyDeviceToHost, s.s_main);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.s_main);
        if (e != cudaSuccess) rc = fail(FRANGI_GPU_ECUDA, "imgaussian copy-back: %s", cudaGetErrorString(e));
    }
    frangi_gpu_destroy(H);
    return rc;
}

FRANGI_API int frangi_gpu_hessian3d(const uint8_t* I_host, int w, int h, int l, float sigma, float zdist,
                                    float* Dzz, float* Dyy, float* Dyz, float* Dxx, float* Dxy, float* Dxz,
                                    int device, unsigned flags)
{
    if (!I_host || !Dzz || !Dyy || !Dyz || !Dxx || !Dxy || !Dxz) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    frangi_gpu* H = nullptr;
    int rc = stage_smooth(&H, I_host, w, h, l, sigma, zdist, device, flags);
    float* host[6] = { Dzz, Dyy, Dyz, Dxx, Dxy, Dxz };
    float* dD[6] = { nullptr };
    if (!rc) {
        Slab& s = H->slabs[0];
        const size_t n = (size_t)s.voxels;
        cudaError_t e = cudaSuccess;
        for (int k = 0; k < 6 && e == cudaSuccess; ++k) e = cudaMalloc(&dD[k], n * 4);
        if (e == cudaSuccess) {
            rc = launch_voxel(H, s, H->scales[0], 0, dD);
            if (rc) e = cudaErrorUnknown;
        }
        for (int k = 0; k < 6 && e == cudaSuccess; ++k)
            e = cudaMemcpyAsync(host[k], dD[k], n * 4, cudaMemcpyDeviceToHost, s.s_main);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.s_main);
        if (e != cudaSuccess && !rc) rc = fail(FRANGI_GPU_ECUDA, "hessian3d stage: %s", cudaGetErrorString(e));
    }
    for (int k = 0; k < 6; ++k) cudaFree(dD[k]);
    frangi_gpu_destroy(H);
    return rc;
}

FRANGI_API int frangi_gpu_vesselness_stage(const float* Dxx, const float* Dxy, const float* Dxz, const float* Dyy,
                                           const float* Dyz, const float* Dzz, int64_t n, float alpha, float beta,
                                           float C, int blackwhite, float* v_out, float* dir_out,
                                           float* lambda_out, int device, unsigned stage_flags)
{
    if (!Dxx || !Dxy || !Dxz || !Dyy || !Dyz || !Dzz || !v_out || n < 1) return fail(FRANGI_GPU_EINVAL, "bad argument");
    RC(check_device(device));
    CK(cudaSetDevice(device));
    const float* host[6] = { Dxx, Dxy, Dxz, Dyy, Dyz, Dzz };
    float* d[6] = { nullptr };
    float *dv = nullptr, *ddir = nullptr, *dlam = nullptr;
    int rc = 0;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 6 && e == cudaSuccess; ++k) {
        e = cudaMalloc(&d[k], (size_t)n * 4);
        if (e == cudaSuccess) e = cudaMemcpy(d[k], host[k], (size_t)n * 4, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc(&dv, (size_t)n * 4);
    if (e == cudaSuccess && dir_out) e = cudaMalloc(&ddir, (size_t)n * 12);
    if (e == cudaSuccess && lambda_out) e = cudaMalloc(&dlam, (size_t)n * 12);
    if (e == cudaSuccess) {
        frangi_gpu tmp;
        tmp.alpha = alpha; tmp.beta = beta; tmp.C = C; tmp.blackwhite = blackwhite;
        FrangiConsts k = make_consts(&tmp, 1.0f);
        if (stage_flags & FRANGI_GPU_STAGE_SCALAR)
            vesselness_stage_scalar_kernel<<<(unsigned)((n + 127) / 128), 128>>>(d[0], d[1], d[2], d[3], d[4], d[5], n,
                                                                                 k, dv, ddir, dlam);
        else
            vesselness_stage_kernel<<<(unsigned)((n + 255) / 256), 128>>>(d[0], d[1], d[2], d[3], d[4], d[5], n, k, dv,
                                                                          ddir, dlam);
        g_launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(v_out, dv, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && dir_out) e = cudaMemcpy(dir_out, ddir, (size_t)n * 12, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && lambda_out) e = cudaMemcpy(lambda_out, dlam, (size_t)n * 12, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(FRANGI_GPU_ECUDA, "vesselness stage: %s", cudaGetErrorString(e));
    for (int k = 0; k < 6; ++k) cudaFree(d[k]);
    cudaFree(dv); cudaFree(ddir); cudaFree(dlam);
    return rc;
}

// ---- f3: pre-pass of SeedExtractor::extractSeeds on the device (seed_kernels.cuh) ------------------
namespace {

// The pre-pass on `nl` layers of a dense device J8 volume; host outputs (layer_min/max/n_max per layer,
// keys appended at keys[*n_keys ...]).  keys == NULL or a too small keys_cap only counts.
int seed_candidates_device(SeedScratch& sc, const uint8_t* dJ8, int w, int h, int nl, cudaStream_t st, uint8_t* layer_min,
                           uint8_t* layer_max, int* n_max, int64_t* keys, int64_t keys_cap, int64_t* n_keys)
{
    if (nl < 1) return 0;
    if (nl > 65535) return fail(FRANGI_GPU_EINVAL, "seed_candidates: more than 65535 layers on one device");
    const long long plane = (long long)w * h;
    if (sc.layers < nl) {
        cudaFree(sc.minmax); cudaFree(sc.count); cudaFree(sc.off);
        sc.minmax = sc.count = nullptr; sc.off = nullptr; sc.layers = 0;
        CK(cudaMalloc(&sc.minmax, sizeof(int) * 2 * nl));
        CK(cudaMalloc(&sc.count, sizeof(int) * nl));
        CK(cudaMalloc(&sc.off, sizeof(long long) * (nl + 1)));
        sc.layers = nl;
    }
    std::vector<int> h_minmax(2 * (size_t)nl), h_count(nl);
    for (int z = 0; z < nl; ++z) { h_minmax[2 * z] = 255; h_minmax[2 * z + 1] = 0; }
    CK(cudaMemcpyAsync(sc.minmax, h_minmax.data(), sizeof(int) * 2 * nl, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(sc.count, 0, sizeof(int) * nl, st));
    const int nb = (int)std::min<long long>(std::max<long long>(1, plane / (4 * 256 * 8)), 148 * 8);
    j8_layer_minmax_kernel<<<dim3(nb, nl), 256, 0, st>>>(dJ8, plane, sc.minmax);
    g_launches++;
    const dim3 grid((w + 31) / 32, (h + 7) / 8, nl), block(32, 8);
    if (grid.y > 65535) return fail(FRANGI_GPU_EINVAL, "seed_candidates: layer too tall");
    j8_local_maxima_kernel<false><<<grid, block, 0, st>>>(dJ8, w, h, sc.minmax, sc.count, nullptr, nullptr);
    g_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_count.data(), sc.count, sizeof(int) * nl, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_minmax.data(), sc.minmax, sizeof(int) * 2 * nl, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<long long> h_off(nl + 1, 0);
    for (int z = 0; z < nl; ++z) {
        h_off[z + 1] = h_off[z] + h_count[z];
        layer_min[z] = (uint8_t)h_minmax[2 * z]; layer_max[z] = (uint8_t)h_minmax[2 * z + 1];
        n_max[z] = h_count[z];
    }
    const long long total = h_off[nl];
    const int64_t at = *n_keys;
    *n_keys = at + total;
    if (!keys || at + total > keys_cap || total == 0) return 0;   // counted only (the caller checks n_keys against its capacity)
    if (sc.keys < total) {
        cudaFree(sc.in); cudaFree(sc.out);
        sc.in = sc.out = nullptr; sc.keys = 0;
        CK(cudaMalloc(&sc.in, sizeof(long long) * total));
        CK(cudaMalloc(&sc.out, sizeof(long long) * total));
        sc.keys = total;
    }
    CK(cudaMemcpyAsync(sc.off, h_off.data(), sizeof(long long) * (nl + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(sc.count, 0, sizeof(int) * nl, st));
    j8_local_maxima_kernel<true><<<grid, block, 0, st>>>(dJ8, w, h, sc.minmax, sc.count, sc.off, sc.in);
    g_launches++;
    CK(cudaGetLastError());
    size_t tmp_bytes = 0;
    CK(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, tmp_bytes, sc.in, sc.out, total, nl, sc.off, sc.off + 1, 0, 63, st));
    if (sc.tmp_bytes < tmp_bytes) {
        cudaFree(sc.tmp); sc.tmp = nullptr; sc.tmp_bytes = 0;
        CK(cudaMalloc(&sc.tmp, std::max<size_t>(tmp_bytes, 16)));
        sc.tmp_bytes = std::max<size_t>(tmp_bytes, 16);
    }
    CK(cub::DeviceSegmentedRadixSort::SortKeys(sc.tmp, tmp_bytes, sc.in, sc.out, total, nl, sc.off, sc.off + 1, 0, 63, st));
    g_launches++;
    CK(cudaMemcpyAsync(keys + at, sc.out, sizeof(long long) * total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace

FRANGI_API int frangi_gpu_seed_candidates(frangi_gpu_t* H, uint8_t* layer_min, uint8_t* layer_max, int* n_max, int64_t* keys,
                                          int64_t keys_cap, int64_t* n_keys)
{
    if (!H || !layer_min || !layer_max || !n_max || !n_keys) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    if (!H->ran) return fail(FRANGI_GPU_ESTATE, "nothing has been run on this handle");
    *n_keys = 0;
    int zoff = 0;
    for (auto& s : H->slabs) {       // layers are independent: every slab handles its own, in z order
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.s_main));
        RC(seed_candidates_device(s.seed, s.dJ8, H->w, H->h, s.ze - s.zb, s.s_main, layer_min + zoff,
                                  layer_max + zoff, n_max + zoff, keys, keys_cap, n_keys));
        zoff += s.ze - s.zb;
    }
    if (keys && *n_keys > keys_cap) return fail(FRANGI_GPU_EINVAL, "seed_candidates: %lld keys, capacity %lld", (long long)*n_keys, (long long)keys_cap);
    return 0;
}

FRANGI_API int frangi_gpu_seed_candidates_host(const uint8_t* J8_host, int w, int h, int l, uint8_t* layer_min, uint8_t* layer_max,
                                               int* n_max, int64_t* keys, int64_t keys_cap, int64_t* n_keys, int device)
{
    if (!J8_host || !layer_min || !layer_max || !n_max || !n_keys) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    if (w < 1 || h < 1 || l < 1 || (long long)w * h > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "bad volume %d x %d x %d", w, h, l);
    RC(check_device(device));
    CK(cudaSetDevice(device));
    const size_t n = (size_t)w * h * l;
    uint8_t* d = nullptr;
    CK(cudaMalloc(&d, n));
    cudaError_t e = cudaMemcpy(d, J8_host, n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return fail(FRANGI_GPU_ECUDA, "upload failed: %s", cudaGetErrorString(e)); }
    *n_keys = 0;
    int rc = 0;
    SeedScratch sc;
    for (int z0 = 0; z0 < l && rc == 0; z0 += 65535)
        rc = seed_candidates_device(sc, d + (size_t)z0 * w * h, w, h, std::min(65535, l - z0), 0, layer_min + z0, layer_max + z0,
                                    n_max + z0, keys, keys_cap, n_keys);
    sc.release();
    cudaFree(d);
    if (rc) return rc;
    if (keys && *n_keys > keys_cap) return fail(FRANGI_GPU_EINVAL, "seed_candidates: %lld keys, capacity %lld", (long long)*n_keys, (long long)keys_cap);
    return 0;
}

// ---- f4: the 2-D path, Frangi::frangi2d / hessian2d (frangi2d_kernels.cuh) -----------------------------
namespace {
struct Dev2D {     // device buffers of one 2-D call, freed on every path
    uint8_t *I = nullptr, *V[3] = { nullptr, nullptr, nullptr };
    float *F = nullptr, *J = nullptr, *D[3] = { nullptr, nullptr, nullptr };
    int* mm = nullptr;
    ~Dev2D() { cudaFree(I); cudaFree(F); cudaFree(J); cudaFree(mm); for (auto v : V) cudaFree(v); for (auto d : D) cudaFree(d); }
};

int run_2d(const uint8_t* I_host, int w, int h, const float* sigmas, int nsig, float beta_one, float beta_two, int blackwhite,
           float* J_host, float* Jmin, float* Jmax, uint8_t* Vx, uint8_t* Vy, uint8_t* Vz, float* const* D_host, int device,
           unsigned flags)
{
    if (!I_host || !sigmas || nsig < 1) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    if (w < 2 || h < 2 || (long long)w * h > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "frangi2d needs 2 <= w, h (got %d, %d)", w, h);
    RC(check_device(device));
    CK(cudaSetDevice(device));
    const int fpitch = (w + 31) / 32 * 32;
    const size_t n = (size_t)w * h;
    Dev2D d;
    CK(cudaMalloc(&d.I, n));
    CK(cudaMalloc(&d.F, sizeof(float) * (size_t)fpitch * h));
    CK(cudaMalloc(&d.J, sizeof(float) * n));
    CK(cudaMalloc(&d.mm, 2 * sizeof(int)));
    for (auto& v : d.V) CK(cudaMalloc(&v, n));
    if (D_host) for (auto& q : d.D) CK(cudaMalloc(&q, sizeof(float) * n));
    CK(cudaMemcpy(d.I, I_host, n, cudaMemcpyHostToDevice));
    const int init[2] = { 0x7f7fffff, (int)0xff7fffffu };      // FLT_MAX, -FLT_MAX (frangi.cpp:414-415)
    CK(cudaMemcpy(d.mm, init, sizeof init, cudaMemcpyHostToDevice));
    for (int si = 0; si < nsig; ++si) {
        if (!(sigmas[si] > 0)) return fail(FRANGI_GPU_EINVAL, "sigma[%d] must be > 0", si);
        ScalePlan sp;
        sp.sigma = sigmas[si]; sp.sigma2 = sigmas[si] * sigmas[si];
        RC(plan_taps(sp.sigma, sp.rxy, sp.rxy_t, sp.txy));
        // always the separately rounded smoothing: the 2-D measure (values up to 1, ratios of eigenvalues squared) does
        // not keep the 1e-4 tolerance under fused multiply-add, and a single image has no speed to gain from it
        RC(launch_xy_planes(d.I, d.F, w, h, 1, fpitch, (long long)fpitch * h, sp, flags & ~(unsigned)FRANGI_GPU_FLAG_FMA_SMOOTHING, 0));
        F2DParams p;
        p.F = d.F; p.w = w; p.h = h; p.fpitch = fpitch;
        p.sigma2 = sp.sigma2;
        p.beta = (float)(2 * ((double)beta_one * (double)beta_one));
        p.c = (float)(2 * ((double)beta_two * (double)beta_two));
        p.blackwhite = blackwhite; p.first = si == 0;
        p.J = d.J; p.Vx = d.V[0]; p.Vy = d.V[1]; p.Vz = d.V[2];
        for (int k = 0; k < 3; ++k) p.D[k] = D_host ? d.D[k] : nullptr;
        p.minmax = d.mm;
        frangi2d_pixel_kernel<<<dim3((w + 127) / 128, h), 128>>>(p);
        g_launches++;
        CK(cudaGetLastError());
    }
    CK(cudaDeviceSynchronize());
    if (D_host) {
        for (int k = 0; k < 3; ++k) CK(cudaMemcpy(D_host[k], d.D[k], sizeof(float) * n, cudaMemcpyDeviceToHost));
        return 0;
    }
    int mm[2];
    CK(cudaMemcpy(mm, d.mm, sizeof mm, cudaMemcpyDeviceToHost));
    if (Jmin) std::memcpy(Jmin, &mm[0], 4);
    if (Jmax) std::memcpy(Jmax, &mm[1], 4);
    if (J_host) CK(cudaMemcpy(J_host, d.J, sizeof(float) * n, cudaMemcpyDeviceToHost));
    uint8_t* const out[3] = { Vx, Vy, Vz };
    for (int k = 0; k < 3; ++k) if (out[k]) CK(cudaMemcpy(out[k], d.V[k], n, cudaMemcpyDeviceToHost));
    return 0;
}
}  // namespace

FRANGI_API int frangi_gpu_frangi2d(const uint8_t* I_host, int w, int h, const float* sigmas, int nsig, float beta_one, float beta_two,
                                   int blackwhite, float* J_host, float* Jmin, float* Jmax, uint8_t* Vx_host, uint8_t* Vy_host,
                                   uint8_t* Vz_host, int device, unsigned flags)
{
    return run_2d(I_host, w, h, sigmas, nsig, beta_one, beta_two, blackwhite, J_host, Jmin, Jmax, Vx_host, Vy_host, Vz_host, nullptr,
                  device, flags);
}

FRANGI_API int frangi_gpu_hessian2d(const uint8_t* I_host, int w, int h, float sigma, float* Dyy, float* Dxy, float* Dxx, int device,
                                    unsigned flags)
{
    if (!Dyy || !Dxy || !Dxx) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    float* const D[3] = { Dyy, Dxy, Dxx };
    return run_2d(I_host, w, h, &sigma, 1, .5f, 15.f, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, D, device, flags);
}

// Frangi::imgaussian(I, w, h, sig, F), the 2-D overload (frangi.h:44, frangi.cpp:563-645): K1 on one plane in the
// separately rounded mode (same taps, clamp and accumulation order as the 3-D x and y passes)
FRANGI_API int frangi_gpu_imgaussian2d(const uint8_t* I_host, int w, int h, float sigma, float* F_host, int device)
{
    if (!I_host || !F_host) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    if (w < 1 || h < 1 || (long long)w * h > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "bad image %d x %d", w, h);
    if (!(sigma > 0)) return fail(FRANGI_GPU_EINVAL, "sigma must be > 0");
    RC(check_device(device));
    CK(cudaSetDevice(device));
    const int fpitch = (w + 31) / 32 * 32;
    Dev2D d;
    CK(cudaMalloc(&d.I, (size_t)w * h));
    CK(cudaMalloc(&d.F, sizeof(float) * (size_t)fpitch * h));
    CK(cudaMemcpy(d.I, I_host, (size_t)w * h, cudaMemcpyHostToDevice));
    ScalePlan sp;
    sp.sigma = sigma; sp.sigma2 = sigma * sigma;
    RC(plan_taps(sp.sigma, sp.rxy, sp.rxy_t, sp.txy));
    RC(launch_xy_planes(d.I, d.F, w, h, 1, fpitch, (long long)fpitch * h, sp, 0u, 0));
    CK(cudaMemcpy2D(F_host, sizeof(float) * (size_t)w, d.F, sizeof(float) * (size_t)fpitch, sizeof(float) * (size_t)w, h,
                    cudaMemcpyDeviceToHost));
    return 0;
}

// ---- f4, second part: the soma helpers (soma_kernels.cuh) -------------------------------------------------
namespace {
struct DevSoma {
    uint8_t *a = nullptr, *b = nullptr;
    float* K = nullptr;
    ~DevSoma() { cudaFree(a); cudaFree(b); cudaFree(K); }
};

int soma_args(const void* I, int w, int h, int l, int device)
{
    if (!I) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    if (w < 1 || h < 1 || l < 1 || (long long)w * h > 0x7fffffffLL || h > 65535)
        return fail(FRANGI_GPU_EINVAL, "bad volume %d x %d x %d", w, h, l);
    RC(check_device(device));
    CK(cudaSetDevice(device));
    return 0;
}

template <bool IS_MIN>
int morph(const uint8_t* I_host, int w, int h, int l, float rad, uint8_t* out_host, int device)
{
    RC(soma_args(I_host, w, h, l, device));
    if (!out_host || !(rad >= 0)) return fail(FRANGI_GPU_EINVAL, "bad argument");
    const int L = (int)std::ceil(rad);                         // frangi.cpp:885
    const size_t n = (size_t)w * h * l;
    DevSoma d;
    CK(cudaMalloc(&d.a, n));
    CK(cudaMalloc(&d.b, n));
    CK(cudaMemcpy(d.a, I_host, n, cudaMemcpyHostToDevice));
    const dim3 grid((w + 255) / 256, h, (unsigned)std::min(l, 4096));
    morph_pass_kernel<IS_MIN, false><<<grid, 256>>>(d.a, d.b, w, h, l, L);
    morph_pass_kernel<IS_MIN, true><<<grid, 256>>>(d.b, d.a, w, h, l, L);
    g_launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpy(out_host, d.a, n, cudaMemcpyDeviceToHost));
    return 0;
}
}  // namespace

FRANGI_API int frangi_gpu_imerode(const uint8_t* I_host, int w, int h, int l, float rad, uint8_t* E_host, int device)
{
    return morph<true>(I_host, w, h, l, rad, E_host, device);
}

// the z-scaled overload (frangi.h:46, frangi.cpp:971-1108): x, y as above, then the minimum along z over
// ceil(rad / zdist) planes each side; a single plane skips the z pass (:1062-1064)
FRANGI_API int frangi_gpu_imerode_z(const uint8_t* I_host, int w, int h, int l, float rad, float zdist, uint8_t* E_host, int device)
{
    RC(soma_args(I_host, w, h, l, device));
    if (!E_host || !(rad >= 0) || !(zdist > 0)) return fail(FRANGI_GPU_EINVAL, "bad argument");
    const int L = (int)std::ceil(rad), Lz = (int)std::ceil(rad / zdist);
    const size_t n = (size_t)w * h * l;
    DevSoma d;
    CK(cudaMalloc(&d.a, n));
    CK(cudaMalloc(&d.b, n));
    CK(cudaMemcpy(d.a, I_host, n, cudaMemcpyHostToDevice));
    const dim3 grid((w + 255) / 256, h, (unsigned)std::min(l, 4096));
    morph_pass_kernel<true, false><<<grid, 256>>>(d.a, d.b, w, h, l, L);
    morph_pass_kernel<true, true><<<grid, 256>>>(d.b, d.a, w, h, l, L);
    g_launches += 2;
    const uint8_t* res = d.a;
    if (l > 1) {
        morph_min_z_kernel<<<grid, 256>>>(d.a, d.b, w, h, l, Lz);
        g_launches++;
        res = d.b;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(E_host, res, n, cudaMemcpyDeviceToHost));
    return 0;
}

FRANGI_API int frangi_gpu_imdilate(uint8_t* I_host, int w, int h, int l, float rad, int device)
{
    return morph<false>(I_host, w, h, l, rad, I_host, device);
}

FRANGI_API int frangi_gpu_imgaussian_xy(uint8_t* I_host, int w, int h, int l, float sig, int device)
{
    RC(soma_args(I_host, w, h, l, device));
    if (!(sig > 0)) return fail(FRANGI_GPU_EINVAL, "sigma must be > 0");
    int r_true = 0, r_tmpl = 0;
    GaussTaps taps;
    RC(plan_taps(sig, r_true, r_tmpl, taps));                  // frangi.cpp:792-804: the same taps as every other imgaussian
    const size_t n = (size_t)w * h * l;
    DevSoma d;
    CK(cudaMalloc(&d.a, n));
    CK(cudaMalloc(&d.K, sizeof(float) * n));
    CK(cudaMemcpy(d.a, I_host, n, cudaMemcpyHostToDevice));
    const dim3 grid((w + 255) / 256, h, (unsigned)std::min(l, 4096));
    gauss_x_u8_kernel<<<grid, 256>>>(d.a, d.K, w, h, l, r_true, r_tmpl, taps);
    gauss_y_trunc_kernel<<<grid, 256>>>(d.K, d.a, w, h, l, r_true, r_tmpl, taps);
    g_launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpy(I_host, d.a, n, cudaMemcpyDeviceToHost));
    return 0;
}

// ---- f3, second half: the per-seed correlation score of the plugin's seed filter (zncc_kernels.cuh) -------------
namespace {
// The template tables of Tracker (tracker.cpp:170-232, 3-D branch), built on the host exactly as the reference's
// constructor builds them: float loop counters stepping by Vs = max(3 sigma / 12, 1), weights
// exp(-(u^2 + w^2) / (2 sigma^2)) evaluated in double and stored as float, their mean accumulated in float.
struct ZnccModel {
    std::vector<float4> samp;
    std::vector<int> first;
    std::vector<float> corrc, sig;
};

void build_zncc_model(const float* sigmas, int nsig, ZnccModel& m)
{
    m.samp.clear(); m.first.assign(1, 0); m.corrc.clear(); m.sig.assign(sigmas, sigmas + nsig);
    for (int i = 0; i < nsig; ++i) {
        const float sg = sigmas[i];
        const int V2 = (int)std::round(1 * sg), U2 = (int)std::round(3 * sg), W2 = (int)std::round(3 * sg);
        float Vs = (float)((3.0 * sg) / 12);
        Vs = (Vs < 1.0) ? 1.0f : Vs;
        std::vector<float> wgt;
        std::vector<float4> off;
        float avg = 0.0f;
        for (float vv = -V2; vv <= V2 + FLT_MIN; vv += Vs)
            for (float uu = -U2; uu <= U2 + FLT_MIN; uu += Vs)
                for (float ww = -W2; ww <= W2 + FLT_MIN; ww += Vs) {
                    const float value = (float)std::exp(-((uu * uu) + (ww * ww)) / (2 * std::pow((double)sg, 2)));
                    wgt.push_back(value);
                    off.push_back(make_float4(vv, uu, ww, 0.0f));
                    avg += value;
                }
        avg /= wgt.size();
        float corrc = 0.0f;
        for (size_t k = 0; k < wgt.size(); ++k) {
            const float d = wgt[k] - avg;                          // model2_wgt - model2_avg, a float difference
            off[k].w = d;
            corrc = (float)((double)corrc + (double)d * (double)d);  // corrc += pow(d, 2)
        }
        m.samp.insert(m.samp.end(), off.begin(), off.end());
        m.first.push_back((int)m.samp.size());
        m.corrc.push_back(corrc);
    }
}

struct DevZncc {
    float4* samp = nullptr; int* first = nullptr; float *corrc = nullptr, *sig = nullptr, *seeds = nullptr, *corr = nullptr, *sg = nullptr;
    uint8_t* img = nullptr;
    ~DevZncc() { cudaFree(samp); cudaFree(first); cudaFree(corrc); cudaFree(sig); cudaFree(seeds); cudaFree(corr); cudaFree(sg); cudaFree(img); }
};

int run_zncc(const uint8_t* img_dev, int w, int h, int l, const float* sigmas, int nsig, const float* seeds6, long long n,
             float* corr_out, float* sig_out, DevZncc& d)
{
    if (!seeds6 || !corr_out || n < 0 || nsig < 1) return fail(FRANGI_GPU_EINVAL, "bad argument");
    if (l < 2) return fail(FRANGI_GPU_EINVAL, "the 2-D template (tracker.cpp:191-207) is not provided: l must be >= 2");
    if (n == 0) return 0;
    ZnccModel m;
    build_zncc_model(sigmas, nsig, m);
    CK(cudaMalloc(&d.samp, sizeof(float4) * m.samp.size()));
    CK(cudaMalloc(&d.first, sizeof(int) * m.first.size()));
    CK(cudaMalloc(&d.corrc, sizeof(float) * nsig));
    CK(cudaMalloc(&d.sig, sizeof(float) * nsig));
    CK(cudaMalloc(&d.seeds, sizeof(float) * 6 * (size_t)n));
    CK(cudaMalloc(&d.corr, sizeof(float) * (size_t)n));
    CK(cudaMalloc(&d.sg, sizeof(float) * (size_t)n));
    CK(cudaMemcpy(d.samp, m.samp.data(), sizeof(float4) * m.samp.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.first, m.first.data(), sizeof(int) * m.first.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.corrc, m.corrc.data(), sizeof(float) * nsig, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.sig, m.sig.data(), sizeof(float) * nsig, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.seeds, seeds6, sizeof(float) * 6 * (size_t)n, cudaMemcpyHostToDevice));
    ZnccParams p;
    p.img = img_dev; p.w = w; p.h = h; p.l = l;
    p.samp = d.samp; p.first = d.first; p.corrc = d.corrc; p.sig = d.sig; p.nsig = nsig;
    p.seeds = d.seeds; p.n = n; p.corr_out = d.corr; p.sig_out = d.sg;
    p.xmax = (float)(w - 1.001); p.ymax = (float)(h - 1.001); p.zmax = (float)(l - 1.001);
    const long long nb = (n + 127) / 128;
    if (nb > 0x7fffffffLL) return fail(FRANGI_GPU_EINVAL, "too many seeds");
    seed_zncc_kernel<<<(unsigned)nb, 128>>>(p);
    g_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpy(corr_out, d.corr, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    if (sig_out) CK(cudaMemcpy(sig_out, d.sg, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    return 0;
}
}  // namespace

// the image is the input the handle holds on its device (after frangi_gpu_run / frangi_gpu_upload); one-slab handles
FRANGI_API int frangi_gpu_seed_zncc(frangi_gpu_t* H, const float* seeds6, int64_t n, float* corr_out, float* sig_out)
{
    if (!H) return fail(FRANGI_GPU_EINVAL, "NULL handle");
    if (H->nslabs_total != 1) return fail(FRANGI_GPU_ESTATE, "seed scoring needs the whole image on one device (one-slab handle)");
    Slab& s = H->slabs[0];
    CK(cudaSetDevice(s.dev));
    RC(sync_all(H));
    std::vector<float> sig;
    for (const auto& sp : H->scales) sig.push_back(sp.sigma);
    DevZncc d;
    return run_zncc(s.dI, H->w, H->h, H->l, sig.data(), (int)sig.size(), seeds6, n, corr_out, sig_out, d);
}

FRANGI_API int frangi_gpu_seed_zncc_host(const uint8_t* I_host, int w, int h, int l, const float* sigmas, int nsig,
                                         const float* seeds6, int64_t n, float* corr_out, float* sig_out, int device)
{
    RC(soma_args(I_host, w, h, l, device));
    if (!sigmas) return fail(FRANGI_GPU_EINVAL, "NULL argument");
    DevZncc d;
    const size_t nb = (size_t)w * h * l;
    CK(cudaMalloc(&d.img, nb));
    CK(cudaMemcpy(d.img, I_host, nb, cudaMemcpyHostToDevice));
    return run_zncc(d.img, w, h, l, sigmas, nsig, seeds6, n, corr_out, sig_out, d);
}
