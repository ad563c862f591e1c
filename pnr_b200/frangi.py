"""Python host-side mirror of the reference's ``Frangi`` class over the C-ABI.

The reference interface (pnr-vaa3d/frangi.h:5-59) is a C++ class; the real
drop-in is the C++ shim in ``pnr_b200/csrc/frangi.h``.  This module binds the
same C-ABI (``include/frangi_gpu.h``) with ctypes for the test and benchmark
harness and keeps the reference's names and argument meaning:

    f = Frangi(sigs, zdist, alpha, beta, C, beta_one, beta_two)   # frangi.h:24
    f.blackwhite = False                                          # frangi.h:22
    J, Jmin, Jmax, Vx, Vy, Vz = f.frangi3d(I)                     # frangi.h:33
    F = Frangi.imgaussian(I, sig, zdist)                          # frangi.h:42
    D = f.hessian3d(I, sig, zdist)                                # frangi.h:35

Volumes are numpy uint8 arrays shaped [l][h][w] (x fastest, frangi.cpp:307).
There is no CPU fallback: if the CUDA library is missing or no B200 is visible
every compute call raises ``FrangiGpuError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FRANGI_GPU_LIB selects another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("FRANGI_GPU_LIB") or os.path.join(_HERE, "_lib", "libfrangi_gpu.so")

FLAG_FMA_SMOOTHING = 1
FLAG_DIR_F32 = 2
FLAG_SCALE_IDX = 4
FLAG_LOCAL_HALO = 8
FLAG_OVERLAP_Z = 16
FLAG_REFERENCE_DIRECTION = 32      # Vx / Vy / Vz with the reference's eigenvector sign (include/frangi_gpu.h)

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)


class FrangiGpuError(RuntimeError):
    pass


class _Outputs(C.Structure):
    _fields_ = [("J", C.c_void_p), ("Vx", C.c_void_p), ("Vy", C.c_void_p), ("Vz", C.c_void_p),
                ("scale_idx", C.c_void_p), ("dir_xyz", C.c_void_p), ("voxels", C.c_int64)]


# every symbol include/frangi_gpu.h declares: (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "frangi_gpu_create": (C.c_int, [C.POINTER(_VP), _f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_uint]),
    "frangi_gpu_create_slab": (C.c_int, [C.POINTER(_VP), _f32p, C.c_int, C.c_float, C.c_float, C.c_float,
                                         C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, _VP, C.c_int, C.c_uint]),
    "frangi_gpu_nccl_unique_id": (C.c_int, [_VP]),
    "frangi_gpu_destroy": (None, [_VP]),
    "frangi_gpu_run": (C.c_int, [_VP, _VP, _VP, _f32p, _f32p, _VP, _VP, _VP, _VP, _VP, _VP]),
    "frangi_gpu_set_stream_chunk": (C.c_int, [_VP, C.c_int]),
    "frangi_gpu_run_device": (C.c_int, [_VP, _VP, _f32p, _f32p]),
    "frangi_gpu_upload": (C.c_int, [_VP, _VP]),
    "frangi_gpu_run_resident": (C.c_int, [_VP, _f32p, _f32p]),
    "frangi_gpu_sync": (C.c_int, [_VP]),
    "frangi_gpu_device_outputs": (C.c_int, [_VP, C.c_int, C.POINTER(_Outputs)]),
    "frangi_gpu_download": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "frangi_gpu_imgaussian": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _VP, C.c_int, C.c_uint]),
    "frangi_gpu_hessian3d": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                       _VP, _VP, _VP, _VP, _VP, _VP, C.c_int, C.c_uint]),
    "frangi_gpu_vesselness_stage": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_int64, C.c_float, C.c_float,
                                              C.c_float, C.c_int, _VP, _VP, _VP, C.c_int, C.c_uint]),
    "frangi_gpu_seed_candidates": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_int64, C.POINTER(C.c_int64)]),
    "frangi_gpu_seed_candidates_host": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP, C.c_int64,
                                                  C.POINTER(C.c_int64), C.c_int]),
    "frangi_gpu_frangi2d": (C.c_int, [_VP, C.c_int, C.c_int, _f32p, C.c_int, C.c_float, C.c_float, C.c_int, _VP, _f32p, _f32p,
                                      _VP, _VP, _VP, C.c_int, C.c_uint]),
    "frangi_gpu_hessian2d": (C.c_int, [_VP, C.c_int, C.c_int, C.c_float, _VP, _VP, _VP, C.c_int, C.c_uint]),
    "frangi_gpu_imgaussian2d": (C.c_int, [_VP, C.c_int, C.c_int, C.c_float, _VP, C.c_int]),
    "frangi_gpu_imerode": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_float, _VP, C.c_int]),
    "frangi_gpu_imerode_z": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _VP, C.c_int]),
    "frangi_gpu_imdilate": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]),
    "frangi_gpu_imgaussian_xy": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]),
    "frangi_gpu_seed_zncc": (C.c_int, [_VP, _VP, C.c_int64, _VP, _VP]),
    "frangi_gpu_seed_zncc_host": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, _VP, C.c_int64, _VP, _VP, C.c_int]),
    "frangi_gpu_host_alloc": (_VP, [C.c_size_t]),
    "frangi_gpu_host_free": (None, [_VP]),
    "frangi_gpu_device_count": (C.c_int, []),
    "frangi_gpu_launch_count": (C.c_uint64, []),
    "frangi_gpu_last_timings": (C.c_int, [_VP, _f32p, C.c_int]),
    "frangi_gpu_timing_depth": (C.c_int, [_VP, C.c_int]),
    "frangi_gpu_slab_count": (C.c_int, [_VP]),
    "frangi_gpu_warnings": (C.c_char_p, [_VP]),
    "frangi_gpu_stream": (_VP, [_VP, C.c_int]),
    "frangi_gpu_last_error": (C.c_char_p, []),
    "frangi_gpu_version": (C.c_char_p, []),
}

_lib = None


def load_library(path: str = LIB_PATH):
    """Loads libfrangi_gpu.so and binds every symbol of include/frangi_gpu.h.
    Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FrangiGpuError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(make -C pnr_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        msg = load_library().frangi_gpu_last_error().decode(errors="replace")
        raise FrangiGpuError(f"frangi_gpu error {rc}: {msg}")


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _vol(I):
    I = np.ascontiguousarray(I, dtype=np.uint8)
    if I.ndim != 3:
        raise ValueError("volume must be uint8 [l][h][w]")
    l, h, w = I.shape
    return I, w, h, l


def launch_count() -> int:
    return int(load_library().frangi_gpu_launch_count())


class PinnedBuffer:
    """Pinned host memory from frangi_gpu_host_alloc exposed as a numpy array."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = self.lib.frangi_gpu_host_alloc(max(nbytes, 1))
        if not self.ptr:
            raise FrangiGpuError("pinned allocation failed: " + self.lib.frangi_gpu_last_error().decode())
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.frangi_gpu_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FrangiPlan:
    """A created handle (frangi_gpu_create / frangi_gpu_create_slab) for one volume shape."""

    def __init__(self, sigmas, zdist, alpha, beta, Cc, blackwhite, w, h, l, devices=(0,), flags=0,
                 slab=None):
        self.lib = load_library()
        self.w, self.h, self.l = w, h, l
        self.flags = flags
        s = np.ascontiguousarray(sigmas, np.float32)
        self.handle = C.c_void_p()
        if slab is None:
            devs = (C.c_int * len(devices))(*devices)
            _check(self.lib.frangi_gpu_create(C.byref(self.handle), s.ctypes.data_as(_f32p), len(s), zdist,
                                              alpha, beta, Cc, int(blackwhite), w, h, l, devs, len(devices), flags))
            self.z_begin, self.z_end = 0, l
        else:
            z0, z1, rank, nranks, uid, device = slab
            uid_buf = C.create_string_buffer(bytes(uid), 128) if uid is not None else None
            _check(self.lib.frangi_gpu_create_slab(C.byref(self.handle), s.ctypes.data_as(_f32p), len(s), zdist,
                                                   alpha, beta, Cc, int(blackwhite), w, h, l, z0, z1, rank,
                                                   nranks, C.cast(uid_buf, C.c_void_p) if uid_buf else None,
                                                   device, flags))
            self.z_begin, self.z_end = z0, z1
        self.nz = self.z_end - self.z_begin

    def close(self):
        if self.handle:
            self.lib.frangi_gpu_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-buffer call: the reference-facing path ----
    def run(self, I, J=None, Vx=None, Vy=None, Vz=None, J8=None, scale=None, direction=None,
            want_J=True, want_J8=False):
        I = np.ascontiguousarray(I, np.uint8)
        shp = (self.nz, self.h, self.w)
        if I.shape != shp:
            raise ValueError(f"input shape {I.shape} != {shp}")
        if J is None and want_J:
            J = np.empty(shp, np.float32)
        Vx = np.empty(shp, np.uint8) if Vx is None else Vx
        Vy = np.empty(shp, np.uint8) if Vy is None else Vy
        Vz = np.empty(shp, np.uint8) if Vz is None else Vz
        if J8 is None and want_J8:
            J8 = np.empty(shp, np.uint8)
        if scale is None and (self.flags & FLAG_SCALE_IDX):
            scale = np.empty(shp, np.uint8)
        if direction is None and (self.flags & FLAG_DIR_F32):
            direction = np.empty((3,) + shp, np.float32)
        lo, hi = C.c_float(), C.c_float()
        _check(self.lib.frangi_gpu_run(self.handle, _ptr(I), _ptr(J), C.byref(lo), C.byref(hi), _ptr(Vx),
                                       _ptr(Vy), _ptr(Vz), _ptr(J8), _ptr(scale), _ptr(direction)))
        return dict(J=J, Jmin=lo.value, Jmax=hi.value, Vx=Vx, Vy=Vy, Vz=Vz, J8=J8, scale=scale, dir=direction)

    # ---- device-resident path (kernel-only timing) ----
    def upload(self, I):
        I = np.ascontiguousarray(I, np.uint8)
        self._uploaded = I   # keep the (possibly pageable) source alive until the copy has run
        _check(self.lib.frangi_gpu_upload(self.handle, _ptr(I)))

    def run_resident(self, sync=True):
        if sync:
            lo, hi = C.c_float(), C.c_float()
            _check(self.lib.frangi_gpu_run_resident(self.handle, C.byref(lo), C.byref(hi)))
            return lo.value, hi.value
        _check(self.lib.frangi_gpu_run_resident(self.handle, None, None))
        return None

    def run_device(self, dev_ptr: int, sync=True):
        if sync:
            lo, hi = C.c_float(), C.c_float()
            _check(self.lib.frangi_gpu_run_device(self.handle, C.c_void_p(dev_ptr), C.byref(lo), C.byref(hi)))
            return lo.value, hi.value
        _check(self.lib.frangi_gpu_run_device(self.handle, C.c_void_p(dev_ptr), None, None))
        return None

    def sync(self):
        _check(self.lib.frangi_gpu_sync(self.handle))

    def download(self, want_J8=False, want_J=True, want_V=True):
        shp = (self.nz, self.h, self.w)
        J = np.empty(shp, np.float32) if want_J else None
        V = [np.empty(shp, np.uint8) if want_V else None for _ in range(3)]
        J8 = np.empty(shp, np.uint8) if want_J8 else None
        sc = np.empty(shp, np.uint8) if (self.flags & FLAG_SCALE_IDX) else None
        dr = np.empty((3,) + shp, np.float32) if (self.flags & FLAG_DIR_F32) else None
        _check(self.lib.frangi_gpu_download(self.handle, _ptr(J), _ptr(V[0]), _ptr(V[1]), _ptr(V[2]), _ptr(J8),
                                            _ptr(sc), _ptr(dr)))
        return dict(J=J, Vx=V[0], Vy=V[1], Vz=V[2], J8=J8, scale=sc, dir=dr)

    def timings(self):
        ms = (C.c_float * 8)()
        _check(self.lib.frangi_gpu_last_timings(self.handle, ms, 8))
        return dict(gauss_xy=ms[0], gauss_z=ms[1], hessian_eigen=ms[2], j8=ms[3], halo_wait=ms[4], total=ms[5])

    def set_stream_chunk(self, planes: int):
        """Planes per pipelined chunk of run(): 0 = one piece, -1 = automatic."""
        _check(self.lib.frangi_gpu_set_stream_chunk(self.handle, planes))

    def timing_depth(self, depth: int):
        """Keep the event sets of the last `depth` runs; timings() then returns their mean."""
        _check(self.lib.frangi_gpu_timing_depth(self.handle, depth))

    def stream(self, slab=0) -> int:
        """cudaStream_t of the slab's kernels as an integer (for torch.cuda.ExternalStream)."""
        return int(self.lib.frangi_gpu_stream(self.handle, slab) or 0)

    def device_outputs(self, slab=0):
        o = _Outputs()
        _check(self.lib.frangi_gpu_device_outputs(self.handle, slab, C.byref(o)))
        return o


    def seed_zncc(self, seeds):
        """Scores seeds (rows x, y, z, vx, vy, vz) against the input image the handle holds on its device."""
        seeds = np.ascontiguousarray(np.asarray(seeds, np.float32)[:, :6])
        corr = np.empty(len(seeds), np.float32)
        sig = np.empty(len(seeds), np.float32)
        _check(self.lib.frangi_gpu_seed_zncc(self.handle, _ptr(seeds), len(seeds), _ptr(corr), _ptr(sig)))
        return corr, sig

    def seed_candidates(self, cap=None):
        """Pre-pass of SeedExtractor::extractSeeds (seed.cpp:574-632) on the J8 volume the last run left on the
        device: dict(layer_min u8[l], layer_max u8[l], n_max i32[l], keys i64[sum n_max]) (sorted per layer)."""
        l = self.l
        lo = np.empty(l, np.uint8); hi = np.empty(l, np.uint8); n = np.empty(l, np.int32)
        total = C.c_int64(0)
        if cap is None:   # count first
            _check(self.lib.frangi_gpu_seed_candidates(self.handle, _ptr(lo), _ptr(hi), _ptr(n), None, 0, C.byref(total)))
            cap = max(1, total.value)
        keys = np.empty(cap, np.int64)
        _check(self.lib.frangi_gpu_seed_candidates(self.handle, _ptr(lo), _ptr(hi), _ptr(n), _ptr(keys), cap, C.byref(total)))
        return dict(layer_min=lo, layer_max=hi, n_max=n, keys=keys[:total.value].copy())


class Frangi:
    """Same constructor, public fields and hot method as the reference class
    (frangi.h:5-59); frangi3d runs on the GPU(s)."""

    def __init__(self, sigs, zdist, alpha, beta, C_, beta_one=0.5, beta_two=15.0, devices=(0,), flags=0):
        self.sig = [float(s) for s in sigs]
        self.zdist = float(zdist)
        self.alpha = float(alpha)
        self.beta = float(beta)
        self.C = float(C_)
        self.BetaOne = float(beta_one)   # 2-D only, unused on this path
        self.BetaTwo = float(beta_two)
        self.blackwhite = False          # frangi.cpp:54
        self.devices = tuple(devices)
        self.flags = flags
        self._plan = None
        self._plan_key = None

    def _get_plan(self, w, h, l):
        key = (tuple(self.sig), self.zdist, self.alpha, self.beta, self.C, bool(self.blackwhite), w, h, l,
               self.devices, self.flags)
        if self._plan is None or self._plan_key != key:
            if self._plan is not None:
                self._plan.close()
            self._plan = FrangiPlan(self.sig, self.zdist, self.alpha, self.beta, self.C, self.blackwhite,
                                    w, h, l, self.devices, self.flags)
            self._plan_key = key
        return self._plan

    def frangi3d(self, I):
        """Returns (J, Jmin, Jmax, Vx, Vy, Vz) like the out-parameters of frangi.h:33."""
        I, w, h, l = _vol(I)
        r = self._get_plan(w, h, l).run(I)
        self.last = r
        return r["J"], r["Jmin"], r["Jmax"], r["Vx"], r["Vy"], r["Vz"]

    def frangi3d_full(self, I, want_J8=True):
        I, w, h, l = _vol(I)
        return self._get_plan(w, h, l).run(I, want_J8=want_J8)

    def frangi2d(self, I):
        """Frangi::frangi2d (frangi.h:38) on a uint8 image [h][w]: dict(J, Jmin, Jmax, Vx, Vy, Vz)."""
        I = np.ascontiguousarray(I, np.uint8)
        if I.ndim != 2:
            raise ValueError("image must be uint8 [h][w]")
        h, w = I.shape
        s = np.ascontiguousarray(self.sig, np.float32)
        J = np.empty(I.shape, np.float32)
        V = [np.empty(I.shape, np.uint8) for _ in range(3)]
        lo, hi = C.c_float(), C.c_float()
        _check(load_library().frangi_gpu_frangi2d(_ptr(I), w, h, s.ctypes.data_as(_f32p), len(s), self.BetaOne, self.BetaTwo,
                                                  int(self.blackwhite), _ptr(J), C.byref(lo), C.byref(hi), _ptr(V[0]),
                                                  _ptr(V[1]), _ptr(V[2]), self.devices[0], self.flags & FLAG_FMA_SMOOTHING))
        return dict(J=J, Jmin=lo.value, Jmax=hi.value, Vx=V[0], Vy=V[1], Vz=V[2])

    def hessian2d(self, I, sig):
        I = np.ascontiguousarray(I, np.uint8)
        h, w = I.shape
        D = {k: np.empty(I.shape, np.float32) for k in ("Dyy", "Dxy", "Dxx")}
        _check(load_library().frangi_gpu_hessian2d(_ptr(I), w, h, sig, _ptr(D["Dyy"]), _ptr(D["Dxy"]), _ptr(D["Dxx"]),
                                                   self.devices[0], self.flags & FLAG_FMA_SMOOTHING))
        return D

    @staticmethod
    def imgaussian(I, sig, zdist, device=0, flags=0):
        I, w, h, l = _vol(I)
        F = np.empty(I.shape, np.float32)
        _check(load_library().frangi_gpu_imgaussian(_ptr(I), w, h, l, sig, zdist, _ptr(F), device, flags))
        return F

    def hessian3d(self, I, sig, zdist=None, device=0):
        """dict Dzz,Dyy,Dyz,Dxx,Dxy,Dxz (the reference's argument order, frangi.h:35)."""
        I, w, h, l = _vol(I)
        zdist = self.zdist if zdist is None else zdist
        names = ["Dzz", "Dyy", "Dyz", "Dxx", "Dxy", "Dxz"]
        D = [np.empty(I.shape, np.float32) for _ in names]
        _check(load_library().frangi_gpu_hessian3d(_ptr(I), w, h, l, sig, zdist, *[_ptr(d) for d in D],
                                                   device, self.flags & FLAG_FMA_SMOOTHING))
        return dict(zip(names, D))

    def vesselness_stage(self, D, device=0, want_lambda=True, scalar=False):
        arrs = [np.ascontiguousarray(D[k], np.float32) for k in ("Dxx", "Dxy", "Dxz", "Dyy", "Dyz", "Dzz")]
        n = arrs[0].size
        v = np.empty(arrs[0].shape, np.float32)
        dr = np.empty((3,) + arrs[0].shape, np.float32)
        lam = np.empty(arrs[0].shape + (3,), np.float32) if want_lambda else None
        _check(load_library().frangi_gpu_vesselness_stage(*[_ptr(a) for a in arrs], n, self.alpha, self.beta,
                                                          self.C, int(self.blackwhite), _ptr(v), _ptr(dr),
                                                          _ptr(lam), device, 1 if scalar else 0))
        return v, dr, lam

    def close(self):
        if self._plan is not None:
            self._plan.close()
            self._plan = None


def seed_candidates(J8, device=0):
    """The same pre-pass on any uint8 volume [l][h][w] held by the host."""
    J8, w, h, l = _vol(J8)
    lib = load_library()
    lo = np.empty(l, np.uint8); hi = np.empty(l, np.uint8); n = np.empty(l, np.int32)
    total = C.c_int64(0)
    _check(lib.frangi_gpu_seed_candidates_host(_ptr(J8), w, h, l, _ptr(lo), _ptr(hi), _ptr(n), None, 0, C.byref(total), device))
    keys = np.empty(max(1, total.value), np.int64)
    _check(lib.frangi_gpu_seed_candidates_host(_ptr(J8), w, h, l, _ptr(lo), _ptr(hi), _ptr(n), _ptr(keys), keys.size,
                                               C.byref(total), device))
    return dict(layer_min=lo, layer_max=hi, n_max=n, keys=keys[:total.value].copy())


def seed_zncc(I, sigmas, seeds, device=0):
    """Tracker::znccBBB for every row (x, y, z, vx, vy, vz) of `seeds` on the raw image I [l][h][w]
    (Advantra_plugin.cpp:2561-2573): returns (corr[n], sigma[n])."""
    I, w, h, l = _vol(I)
    seeds = np.ascontiguousarray(np.asarray(seeds, np.float32)[:, :6])
    s = np.ascontiguousarray(sigmas, np.float32)
    corr = np.empty(len(seeds), np.float32)
    sig = np.empty(len(seeds), np.float32)
    _check(load_library().frangi_gpu_seed_zncc_host(_ptr(I), w, h, l, s.ctypes.data_as(_f32p), len(s), _ptr(seeds), len(seeds),
                                                    _ptr(corr), _ptr(sig), device))
    return corr, sig


def imerode(I, rad, device=0):
    """Frangi::imerode(I,w,h,l,rad,E) (frangi.h:47): xy minimum filter of a uint8 volume [l][h][w]."""
    I, w, h, l = _vol(I)
    out = np.empty_like(I)
    _check(load_library().frangi_gpu_imerode(_ptr(I), w, h, l, rad, _ptr(out), device))
    return out


def imerode_z(I, rad, zdist, device=0):
    """Frangi::imerode(I,w,h,l,rad,zdist,E) (frangi.h:46): the xy minimum followed by the minimum along z over
    ceil(rad/zdist) planes each side."""
    I, w, h, l = _vol(I)
    out = np.empty_like(I)
    _check(load_library().frangi_gpu_imerode_z(_ptr(I), w, h, l, rad, zdist, _ptr(out), device))
    return out


def imgaussian2d(I, sig, device=0):
    """Frangi::imgaussian(I,w,h,sig,F) (frangi.h:44), the 2-D overload on an image [h][w]."""
    I = np.ascontiguousarray(I, np.uint8)
    h, w = I.shape
    F = np.empty(I.shape, np.float32)
    _check(load_library().frangi_gpu_imgaussian2d(_ptr(I), w, h, sig, _ptr(F), device))
    return F


def imdilate(I, rad, device=0):
    """Frangi::imdilate(I,w,h,l,rad) (frangi.h:49); returns the dilated copy."""
    I, w, h, l = _vol(I)
    out = I.copy()
    _check(load_library().frangi_gpu_imdilate(_ptr(out), w, h, l, rad, device))
    return out


def imgaussian_xy(I, sig, device=0):
    """Frangi::imgaussian(I,w,h,l,sig) (frangi.h:43), the in-place xy Gaussian of the soma branch; returns the copy."""
    I, w, h, l = _vol(I)
    out = I.copy()
    _check(load_library().frangi_gpu_imgaussian_xy(_ptr(out), w, h, l, sig, device))
    return out
