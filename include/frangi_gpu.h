/*
 * frangi_gpu.h -- C-ABI of the B200-native multi-scale 3-D Frangi filter.
 *
 * This is the drop-in boundary for the one hot path of miroslavradojevic/pnr
 * that this library replaces: Frangi::frangi3d and the stages under it
 * (pnr-vaa3d/frangi.h:33, frangi.cpp:152-289 with hessian3d :291-390,
 * imgaussian :647-784, eigen_decomposition :1269-1493).  The reference has no
 * FFI of its own (it is a single C++ plugin); the binding a maintainer adds is
 * the Frangi class shim in pnr_b200/csrc/frangi.h, whose frangi3d forwards to
 * frangi_gpu_create / frangi_gpu_run / frangi_gpu_destroy.
 *
 * Conventions: extern "C", plain pointers and sizes, return 0 on success or a
 * FRANGI_GPU_E* code (message via frangi_gpu_last_error), no C++ exceptions
 * cross the boundary, the caller owns every host buffer, the handle owns all
 * device memory, streams and NCCL communicators, one handle is used by one
 * thread at a time.  Volumes are x-fastest: index = z*w*h + y*w + x
 * (frangi.cpp:307).  There is no CPU fallback: every entry point that computes
 * fails with FRANGI_GPU_ECUDA when no sm_100 device is usable.
 */
#ifndef FRANGI_GPU_H
#define FRANGI_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct frangi_gpu frangi_gpu_t;

enum {
    FRANGI_GPU_OK = 0,
    FRANGI_GPU_EINVAL = 1, /* bad argument (dims < 2, nsig < 1, sigma out of range, ...) */
    FRANGI_GPU_ECUDA = 2,  /* CUDA runtime / driver failure, or no usable device */
    FRANGI_GPU_ENOMEM = 3, /* device or host allocation failed */
    FRANGI_GPU_ENCCL = 4,  /* NCCL unavailable or failed */
    FRANGI_GPU_ESTATE = 5  /* call not valid for this handle (e.g. outputs not requested) */
};

/* flags for frangi_gpu_create* */
enum {
    /* Gaussian passes use fused multiply-add instead of the reference's separate
     * float32 multiply and add (frangi.cpp:694,729,762).  Faster; the smoothed
     * volume is then no longer bit-identical to the reference's, the final
     * outputs stay within the tolerances of BASELINE.json. */
    FRANGI_GPU_FLAG_FMA_SMOOTHING = 1,
    /* keep an unquantised float32 direction (3 planar volumes x,y,z) */
    FRANGI_GPU_FLAG_DIR_F32 = 2,
    /* keep the arg-max scale index per voxel (uint8) */
    FRANGI_GPU_FLAG_SCALE_IDX = 4,
    /* frangi_gpu_create with ndev > 1: move halos with peer copies inside the process
     * instead of NCCL send/recv (implied when device_ids repeats a device, which is how
     * the slab decomposition is exercised on a single GPU) */
    FRANGI_GPU_FLAG_LOCAL_HALO = 8,
    /* one-device handles with >= 2 scales: run the (HBM-bound) z pass of a scale on a second,
     * low-priority stream beside the (issue-bound) xy pass of the next scale and the Hessian /
     * eigen stage of the previous one; needs a second pair of intermediate buffers
     * (+8 B/voxel).  Results are bit-identical.  Off by default: measured gain 1.7 %
     * (DESIGN.md section 6), and multi-slab handles do not have it. */
    FRANGI_GPU_FLAG_OVERLAP_Z = 16,
    /* Direction bytes with the reference's SIGN.  The reference writes column 0 of what its double-precision
     * Householder / QL solver returns (frangi.cpp:198,239-250,1269-1495): the sign of that eigenvector is an
     * accident of the iteration, and the closed form of the fast path returns the other one in half of the voxels
     * (same axis within 0.5 degrees).  With this flag every voxel a scale wins is re-solved on the device with the
     * reference's own algorithm in double precision, operation for operation, and Vx/Vy/Vz (and the float direction)
     * are the reference's: byte-identical with bit-exact smoothing, except where the arg-max scale itself is a tie.
     * J, Jmin, Jmax, J8 are unaffected.  Costs a double-precision pass per scale (an opt-in parity mode). */
    FRANGI_GPU_FLAG_REFERENCE_DIRECTION = 32
};

/* ---- whole-volume handle, one process driving ndev devices ---------------
 * Replaces: Frangi::Frangi(sigs, zdist, alpha, beta, C, beta_one, beta_two)
 * (frangi.h:24, frangi.cpp:35-56) plus the scratch allocation at the top of
 * frangi3d (frangi.cpp:158-163).  beta_one/beta_two are 2-D only and unused.
 * The volume is split into ndev contiguous z-slabs, one per device; halos of
 * the xy-smoothed volume are exchanged between neighbouring slabs per scale.
 * device_ids == NULL means devices 0..ndev-1.  `sigmas` must be ascending as
 * the caller guarantees (Advantra_plugin.cpp:1895). */
int frangi_gpu_create(frangi_gpu_t** out, const float* sigmas, int nsig, float zdist,
                      float alpha, float beta, float C, int blackwhite,
                      int w, int h, int l, const int* device_ids, int ndev, unsigned flags);

/* ---- one-slab handle for one-process-per-GPU jobs (torchrun) --------------
 * This rank owns planes [z_begin, z_end) of the w*h*l volume and talks to ranks
 * rank-1 / rank+1 over NCCL send/recv.  nccl_unique_id is the 128-byte id from
 * frangi_gpu_nccl_unique_id on rank 0, distributed by the caller (ignored when
 * nranks == 1).  All host/device buffers passed to run calls on this handle
 * cover only the rank's own planes. */
int frangi_gpu_create_slab(frangi_gpu_t** out, const float* sigmas, int nsig, float zdist,
                           float alpha, float beta, float C, int blackwhite,
                           int w, int h, int l, int z_begin, int z_end,
                           int rank, int nranks, const void* nccl_unique_id,
                           int device, unsigned flags);

int frangi_gpu_nccl_unique_id(void* out128);

/* Replaces: Frangi::~Frangi and the delete[]s at frangi.cpp:279-284. */
void frangi_gpu_destroy(frangi_gpu_t* h);

/* ---- the hot call ----------------------------------------------------------
 * Replaces: void Frangi::frangi3d(unsigned char* I, int w, int h, int l,
 *     float* J, float& Jmin, float& Jmax, unsigned char* Vx, Vy, Vz)
 * (frangi.h:33, frangi.cpp:152-289).  Host buffers; copies are inside the call.
 * J_host may be NULL when only J8_host is wanted (the caller at
 * Advantra_plugin.cpp:2499-2514 converts J to 8 bits and frees J at once);
 * J8_host, scale_idx_host, dir_xyz_host are optional (NULL) extras:
 *   J8        = clamp(round((J-Jmin)/(Jmax-Jmin)*255)), 0 when |Jmax-Jmin|<=FLT_MIN
 *   scale_idx = index into sigmas[] of the scale that produced J
 *   dir_xyz   = 3 planar float32 volumes (all x, then all y, then all z). */
int frangi_gpu_run(frangi_gpu_t* h, const uint8_t* I_host,
                   float* J_host, float* Jmin, float* Jmax,
                   uint8_t* Vx_host, uint8_t* Vy_host, uint8_t* Vz_host,
                   uint8_t* J8_host, uint8_t* scale_idx_host, float* dir_xyz_host);

/* On a single-device handle frangi_gpu_run pipelines host->device copies, kernels and
 * device->host copies over z chunks of `planes` planes (results are bit-identical to the
 * one-piece run; the overlap needs pinned host buffers).  planes = 0 switches the
 * pipelining off, -1 (default) picks max(32, 2*halo) and applies it to volumes of at least
 * two chunks. */
int frangi_gpu_set_stream_chunk(frangi_gpu_t* h, int planes);

/* Device-resident variant (single-device handles only): I_dev is a dense uint8
 * device buffer of the handle's planes; results stay on the device and are read
 * through frangi_gpu_device_outputs.  Asynchronous on the handle's stream;
 * frangi_gpu_sync waits.  Used for kernel-only timing. */
int frangi_gpu_run_device(frangi_gpu_t* h, const uint8_t* I_dev, float* Jmin, float* Jmax);
int frangi_gpu_upload(frangi_gpu_t* h, const uint8_t* I_host); /* fills the handle's own input buffer */
int frangi_gpu_run_resident(frangi_gpu_t* h, float* Jmin, float* Jmax); /* runs on that buffer */
int frangi_gpu_sync(frangi_gpu_t* h);

typedef struct frangi_gpu_outputs {
    const float* J;          /* device pointers, dense, own planes */
    const uint8_t* Vx;
    const uint8_t* Vy;
    const uint8_t* Vz;
    const uint8_t* scale_idx; /* NULL unless FRANGI_GPU_FLAG_SCALE_IDX */
    const float* dir_xyz;     /* NULL unless FRANGI_GPU_FLAG_DIR_F32 */
    int64_t voxels;           /* own voxels = w*h*(z_end-z_begin) */
} frangi_gpu_outputs_t;
int frangi_gpu_device_outputs(frangi_gpu_t* h, int slab, frangi_gpu_outputs_t* out);

/* Copies results of the last run to host buffers (any may be NULL). */
int frangi_gpu_download(frangi_gpu_t* h, float* J_host, uint8_t* Vx_host, uint8_t* Vy_host,
                        uint8_t* Vz_host, uint8_t* J8_host, uint8_t* scale_idx_host,
                        float* dir_xyz_host);

/* ---- stage entry points (parity tests, and the other public Frangi members) --
 * Replaces: static Frangi::imgaussian(I,w,h,l,sig,zdist,F) (frangi.h:42,
 * frangi.cpp:647-784) and Frangi::hessian3d (frangi.h:35, frangi.cpp:291-390).
 * Host buffers, device 0 of the handle-less call. */
int frangi_gpu_imgaussian(const uint8_t* I_host, int w, int h, int l, float sigma, float zdist,
                          float* F_host, int device, unsigned flags);
int frangi_gpu_hessian3d(const uint8_t* I_host, int w, int h, int l, float sigma, float zdist,
                         float* Dzz, float* Dyy, float* Dyz, float* Dxx, float* Dxy, float* Dxz,
                         int device, unsigned flags);
/* Per-voxel stage alone: eigen-decomposition + vesselness + direction of n
 * symmetric 3x3 matrices given as six float32 arrays (host).  v_out[n],
 * dir_out[3*n] planar, lambda_out[3*n] interleaved |l1|<=|l2|<=|l3| (nullable).
 * stage_flags: 0 = the packed two-voxel path of the interior kernel,
 * FRANGI_GPU_STAGE_SCALAR = the scalar path of the face shell. */
enum { FRANGI_GPU_STAGE_SCALAR = 1 };
int frangi_gpu_vesselness_stage(const float* Dxx, const float* Dxy, const float* Dxz,
                                const float* Dyy, const float* Dyz, const float* Dzz, int64_t n,
                                float alpha, float beta, float C, int blackwhite,
                                float* v_out, float* dir_out, float* lambda_out, int device,
                                unsigned stage_flags);

/* ---- f3: the data-parallel pre-pass of the first consumer -----------------------
 * Replaces the per-layer scans at the head of SeedExtractor::extractSeeds (seed.h:101,
 * seed.cpp:574-632): for every z layer of the 8-bit vesselness the layer range
 * (globalMin / globalMax, :578-586), the 8-neighbour candidate maxima (:590-614) and their
 * ranking keys `(int)((v-globalMin)*vFactor) << 32 | y*w+x` sorted ascending (:616-630) --
 * bit for bit the reference's maxPoints arrays, all layers concatenated (layer z starts at
 * sum(n_max[0..z-1])).  The sequential flood-fill analysis (:643-782) stays with the caller.
 * frangi_gpu_seed_candidates works on the J8 volume the last run left on the device(s)
 * (every slab handles its own layers; layer arrays hold the handle's local planes);
 * the _host variant takes any uint8 volume.  keys may be NULL (count only); on return
 * *n_keys is the number of candidates, an error is raised if it exceeds keys_cap. */
int frangi_gpu_seed_candidates(frangi_gpu_t* h, uint8_t* layer_min, uint8_t* layer_max, int* n_max,
                               int64_t* keys, int64_t keys_cap, int64_t* n_keys);
int frangi_gpu_seed_candidates_host(const uint8_t* J8_host, int w, int h, int l, uint8_t* layer_min,
                                    uint8_t* layer_max, int* n_max, int64_t* keys, int64_t keys_cap,
                                    int64_t* n_keys, int device);

/* ---- f4: the 2-D path ---------------------------------------------------------
 * Replaces Frangi::frangi2d(I,w,h,l,J,Jmin,Jmax,Vx,Vy,Vz) (frangi.h:38, frangi.cpp:392-505;
 * the plugin calls it for single-plane images, Advantra_plugin.cpp:2497) and
 * Frangi::hessian2d(I,w,h,sig,Dyy,Dxy,Dxx) (frangi.h:40, frangi.cpp:507-560) over the 2-D
 * imgaussian (frangi.cpp:562-645).  Host buffers of w*h elements, handle-less; beta_one /
 * beta_two are the constructor's 2-D constants (frangi.h:13-14).  Vz is filled with 0 as
 * the reference does.  flags is accepted for symmetry; the 2-D path always smooths with the
 * reference's separately rounded multiply and add (FRANGI_GPU_FLAG_FMA_SMOOTHING is ignored). */
int frangi_gpu_frangi2d(const uint8_t* I_host, int w, int h, const float* sigmas, int nsig,
                        float beta_one, float beta_two, int blackwhite, float* J_host, float* Jmin,
                        float* Jmax, uint8_t* Vx_host, uint8_t* Vy_host, uint8_t* Vz_host, int device,
                        unsigned flags);
int frangi_gpu_hessian2d(const uint8_t* I_host, int w, int h, float sigma, float* Dyy, float* Dxy,
                         float* Dxx, int device, unsigned flags);
/* Frangi::imgaussian(I,w,h,sig,F), the 2-D overload (frangi.h:44, frangi.cpp:563-645): F_host = w*h
 * floats, always the separately rounded smoothing (bit-identical to the reference). */
int frangi_gpu_imgaussian2d(const uint8_t* I_host, int w, int h, float sigma, float* F_host, int device);

/* ---- f4, second part: the helpers of the plugin's soma branch -----------------------
 * (Advantra_plugin.cpp:2426-2440, only reached with somaradius > 0).  Host buffers of
 * w*h*l uint8, handle-less.  Replace, byte for byte:
 *   Frangi::imerode(I,w,h,l,rad,E)   frangi.h:47, frangi.cpp:879-969   xy minimum, radius ceil(rad)
 *   Frangi::imdilate(I,w,h,l,rad)    frangi.h:49, frangi.cpp:1110-1199 xy maximum, in place
  *   Frangi::imgaussian(I,w,h,l,sig)  frangi.h:43, frangi.cpp:786-877   xy Gaussian in place; the y
 *                                    pass accumulates into the unsigned char, truncating per tap */
int frangi_gpu_imerode(const uint8_t* I_host, int w, int h, int l, float rad, uint8_t* E_host, int device);
/*   Frangi::imerode(I,w,h,l,rad,zdist,E)  frangi.h:46, frangi.cpp:971-1108  the same followed by the minimum
 *                                    along z over ceil(rad/zdist) planes each side (skipped when l == 1) */
int frangi_gpu_imerode_z(const uint8_t* I_host, int w, int h, int l, float rad, float zdist, uint8_t* E_host,
                         int device);
int frangi_gpu_imdilate(uint8_t* I_host, int w, int h, int l, float rad, int device);
int frangi_gpu_imgaussian_xy(uint8_t* I_host, int w, int h, int l, float sig, int device);

/* ---- f3, second half: the per-seed correlation score of the plugin's seed filter -------------
 * Tracker::znccBBB (tracker.cpp:1891-1964) as the plugin calls it for every extracted seed
 * (Advantra_plugin.cpp:2561-2573, 3-D images): seeds6 = n rows of (x, y, z, vx, vy, vz) as
 * extractSeeds produced them; corr_out[n] = the score, sig_out[n] (nullable) = the sigma that
 * scored best.  Bit-identical to the reference (one thread per seed in the reference's operation
 * order), so the znccth filter and the sort by score that follow decide exactly as the plugin does.
 * frangi_gpu_seed_zncc scores against the input image the handle holds on its device after
 * frangi_gpu_run / frangi_gpu_upload (one-slab handles), with the handle's sigmas -- the plugin
 * passes the same list to the tracker (Advantra_plugin.cpp:2488,2526). */
int frangi_gpu_seed_zncc(frangi_gpu_t* h, const float* seeds6, int64_t n, float* corr_out, float* sig_out);
int frangi_gpu_seed_zncc_host(const uint8_t* I_host, int w, int h, int l, const float* sigmas, int nsig,
                              const float* seeds6, int64_t n, float* corr_out, float* sig_out, int device);

/* ---- utilities --------------------------------------------------------------*/
void* frangi_gpu_host_alloc(size_t bytes); /* pinned host memory (NULL on failure) */
void frangi_gpu_host_free(void* p);
int frangi_gpu_device_count(void);
/* kernels launched by this library in this process so far */
uint64_t frangi_gpu_launch_count(void);
/* per-kernel-class device time of the last run on slab 0, milliseconds:
 * [0]=gauss_xy [1]=gauss_z [2]=hessian_eigen [3]=minmax/j8 [4]=halo wait; n<=8 */
int frangi_gpu_last_timings(frangi_gpu_t* h, float* ms, int n);
/* Keep the per-kernel event sets of the last `depth` runs (default 1) so that a
 * benchmark can launch K runs back to back without a host sync and still read
 * every launch's device time; frangi_gpu_last_timings then reports the MEAN over
 * the runs recorded since this call (at most `depth`). */
int frangi_gpu_timing_depth(frangi_gpu_t* h, int depth);
/* Slabs the handle really uses: frangi_gpu_create gives every slab at least ceil(3 sigma_max / zdist) + 2
 * planes, so a thin volume uses fewer devices than were passed (the results are the same). */
int frangi_gpu_slab_count(frangi_gpu_t* h);
/* Conditions that did not fail a call but that the caller should know about ("" when there are none):
 * devices left unused, peer access unavailable between two devices of a local-halo handle (the halo
 * copies are then staged through the host). */
const char* frangi_gpu_warnings(frangi_gpu_t* h);
/* The cudaStream_t (as void*) on which slab `slab` of the handle launches its
 * kernels, so that a caller can bracket runs with its own CUDA events. */
void* frangi_gpu_stream(frangi_gpu_t* h, int slab);
const char* frangi_gpu_last_error(void);
const char* frangi_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FRANGI_GPU_H */
