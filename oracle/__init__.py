"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers of the Frangi hot path:

* ``Oracle``     -> oracle/liboracle.so, our C restatement (frangi_oracle.c,
                    seeds_oracle.c) of pnr-vaa3d/frangi.cpp and seed.cpp:556-791.
* ``Reference``  -> oracle/_ref/libpnr_ref.so, the UNMODIFIED reference sources
                    compiled in place from /root/reference (oracle/Makefile,
                    ref_wrap.cpp).  Present only where it was built (the build
                    container) or where the prebuilt file travelled (gpurun).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  pnr_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libpnr_ref.so")

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and, when /root/reference is present, _ref/libpnr_ref.so."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _check_vol(I):
    I = np.ascontiguousarray(I, dtype=np.uint8)
    if I.ndim != 3:
        raise ValueError("volume must be [l][h][w] uint8")
    l, h, w = I.shape
    return I, w, h, l


class Oracle:
    """Our C restatement.  Volumes are numpy arrays shaped [l][h][w] (x fastest)."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.oracle_gauss_radius.restype = C.c_int
        L.oracle_gauss_radius.argtypes = [C.c_float]
        L.oracle_gauss_taps.restype = None
        L.oracle_gauss_taps.argtypes = [C.c_float, C.c_int, _f32p]
        L.oracle_imgaussian.restype = C.c_int
        L.oracle_imgaussian.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _f32p]
        L.oracle_hessian3d.restype = C.c_int
        L.oracle_hessian3d.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float] + [_f32p] * 6
        L.oracle_eigen3.restype = None
        L.oracle_eigen3.argtypes = [_f64p, _f64p, _f64p]
        L.oracle_frangi3d.restype = C.c_int
        L.oracle_frangi3d.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_float,
                                      C.c_float, C.c_float, C.c_float, C.c_int,
                                      _f32p, _f32p, _f32p, _u8p, _u8p, _u8p, _u8p, _f32p]
        L.oracle_vesselness_stage.restype = None
        L.oracle_vesselness_stage.argtypes = [_f32p] * 6 + [C.c_int64, C.c_float, C.c_float, C.c_float,
                                                            C.c_int, _f32p, _f32p, _f64p]
        L.oracle_j_to_j8.restype = None
        L.oracle_j_to_j8.argtypes = [_f32p, C.c_int64, C.c_float, C.c_float, _u8p]
        L.oracle_extract_seeds.restype = C.c_long
        L.oracle_extract_seeds.argtypes = [C.c_double, _u8p, C.c_int, C.c_int, C.c_int,
                                           _u8p, _u8p, _u8p, _f32p, C.c_long]
        L.oracle_frangi2d.restype = C.c_int
        L.oracle_frangi2d.argtypes = [_u8p, C.c_int, C.c_int, _f32p, C.c_int, C.c_float, C.c_float, C.c_int,
                                      _f32p, _f32p, _f32p, _u8p, _u8p, _u8p]
        L.oracle_hessian2d.restype = C.c_int
        L.oracle_hessian2d.argtypes = [_u8p, C.c_int, C.c_int, C.c_float, _f32p, _f32p, _f32p, _f32p]
        L.oracle_morph_xy.restype = C.c_int
        L.oracle_morph_xy.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _u8p]
        L.oracle_imgaussian_xy_u8.restype = C.c_int
        L.oracle_imgaussian_xy_u8.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.oracle_seed_candidates.restype = C.c_long
        L.oracle_seed_candidates.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, _u8p, C.POINTER(C.c_int),
                                             C.POINTER(C.c_int64), C.c_long]

    def gauss_taps(self, sigma: float):
        r = self.lib.oracle_gauss_radius(sigma)
        t = np.zeros(2 * r + 1, np.float32)
        self.lib.oracle_gauss_taps(sigma, r, _p(t, _f32p))
        return r, t

    def imgaussian(self, I, sigma, zdist):
        I, w, h, l = _check_vol(I)
        F = np.empty(I.shape, np.float32)
        rc = self.lib.oracle_imgaussian(_p(I, _u8p), w, h, l, sigma, zdist, _p(F, _f32p))
        if rc:
            raise RuntimeError(f"oracle_imgaussian rc={rc}")
        return F

    def hessian3d(self, I, sigma, zdist):
        """Returns dict with keys Dzz,Dyy,Dyz,Dxx,Dxy,Dxz (reference argument order)."""
        I, w, h, l = _check_vol(I)
        names = ["Dzz", "Dyy", "Dyz", "Dxx", "Dxy", "Dxz"]
        D = [np.empty(I.shape, np.float32) for _ in names]
        rc = self.lib.oracle_hessian3d(_p(I, _u8p), w, h, l, sigma, zdist, *[_p(d, _f32p) for d in D])
        if rc:
            raise RuntimeError(f"oracle_hessian3d rc={rc}")
        return dict(zip(names, D))

    def eigen3(self, A):
        A = np.ascontiguousarray(A, np.float64).reshape(9)
        V = np.empty(9, np.float64)
        d = np.empty(3, np.float64)
        self.lib.oracle_eigen3(_p(A, _f64p), _p(V, _f64p), _p(d, _f64p))
        return V.reshape(3, 3), d

    def frangi3d(self, I, sigmas, zdist=2.0, alpha=0.5, beta=0.5, Cc=500.0, blackwhite=False,
                 want_scale=True, want_dir=True):
        I, w, h, l = _check_vol(I)
        s = np.ascontiguousarray(sigmas, np.float32)
        J = np.empty(I.shape, np.float32)
        V = [np.empty(I.shape, np.uint8) for _ in range(3)]
        sc = np.zeros(I.shape, np.uint8) if want_scale else None
        dr = np.zeros((3,) + I.shape, np.float32) if want_dir else None
        lo, hi = C.c_float(), C.c_float()
        rc = self.lib.oracle_frangi3d(_p(I, _u8p), w, h, l, _p(s, _f32p), len(s), zdist, alpha, beta, Cc,
                                      int(blackwhite), _p(J, _f32p), C.byref(lo), C.byref(hi),
                                      _p(V[0], _u8p), _p(V[1], _u8p), _p(V[2], _u8p),
                                      _p(sc, _u8p), _p(dr, _f32p))
        if rc:
            raise RuntimeError(f"oracle_frangi3d rc={rc}")
        return dict(J=J, Jmin=lo.value, Jmax=hi.value, Vx=V[0], Vy=V[1], Vz=V[2], scale=sc, dir=dr)

    def vesselness_stage(self, D, alpha=0.5, beta=0.5, Cc=500.0, blackwhite=False, want_lambda=False):
        """D: dict Dxx,Dxy,Dxz,Dyy,Dyz,Dzz of equal-shape float32 arrays."""
        arrs = [np.ascontiguousarray(D[k], np.float32) for k in ("Dxx", "Dxy", "Dxz", "Dyy", "Dyz", "Dzz")]
        n = arrs[0].size
        v = np.empty(arrs[0].shape, np.float32)
        dr = np.empty((3,) + arrs[0].shape, np.float32)
        lam = np.empty(arrs[0].shape + (3,), np.float64) if want_lambda else None
        self.lib.oracle_vesselness_stage(*[_p(a, _f32p) for a in arrs], n, alpha, beta, Cc, int(blackwhite),
                                         _p(v, _f32p), _p(dr, _f32p), _p(lam, _f64p))
        return v, dr, lam

    def j_to_j8(self, J, Jmin, Jmax):
        J = np.ascontiguousarray(J, np.float32)
        out = np.empty(J.shape, np.uint8)
        self.lib.oracle_j_to_j8(_p(J, _f32p), J.size, Jmin, Jmax, _p(out, _u8p))
        return out

    def extract_seeds(self, tolerance, J8, Vx, Vy, Vz):
        J8, w, h, l = _check_vol(J8)
        Vx, Vy, Vz = (np.ascontiguousarray(v, np.uint8) for v in (Vx, Vy, Vz))
        cap = max(1024, J8.size // 8)
        out = np.empty((cap, 6), np.float32)
        n = self.lib.oracle_extract_seeds(tolerance, _p(J8, _u8p), w, h, l, _p(Vx, _u8p), _p(Vy, _u8p),
                                          _p(Vz, _u8p), _p(out, _f32p), cap)
        if n < 0 or n > cap:
            raise RuntimeError(f"oracle_extract_seeds returned {n} (cap {cap})")
        return out[:n].copy()


    def frangi2d(self, I, sigmas, beta_one=0.5, beta_two=15.0, blackwhite=False):
        """Frangi::frangi2d (frangi.cpp:392-505) on a uint8 image [h][w]."""
        I = np.ascontiguousarray(I, np.uint8)
        h, w = I.shape
        s = np.ascontiguousarray(sigmas, np.float32)
        J = np.empty(I.shape, np.float32)
        V = [np.empty(I.shape, np.uint8) for _ in range(3)]
        lo = np.zeros(1, np.float32); hi = np.zeros(1, np.float32)
        rc = self.lib.oracle_frangi2d(_p(I, _u8p), w, h, _p(s, _f32p), len(s), beta_one, beta_two, int(blackwhite),
                                      _p(J, _f32p), _p(lo, _f32p), _p(hi, _f32p), _p(V[0], _u8p), _p(V[1], _u8p),
                                      _p(V[2], _u8p))
        if rc:
            raise RuntimeError("oracle_frangi2d failed")
        return dict(J=J, Jmin=float(lo[0]), Jmax=float(hi[0]), Vx=V[0], Vy=V[1], Vz=V[2])

    def hessian2d(self, I, sigma):
        I = np.ascontiguousarray(I, np.uint8)
        h, w = I.shape
        D = {k: np.empty(I.shape, np.float32) for k in ("Dyy", "Dxy", "Dxx", "F")}
        self.lib.oracle_hessian2d(_p(I, _u8p), w, h, sigma, _p(D["Dyy"], _f32p), _p(D["Dxy"], _f32p),
                                  _p(D["Dxx"], _f32p), _p(D["F"], _f32p))
        return D

    def imerode(self, I, rad):
        I, w, h, l = _check_vol(I)
        out = np.empty_like(I)
        self.lib.oracle_morph_xy(_p(I, _u8p), w, h, l, rad, 1, _p(out, _u8p))
        return out

    def imdilate(self, I, rad):
        I, w, h, l = _check_vol(I)
        out = np.empty_like(I)
        self.lib.oracle_morph_xy(_p(I, _u8p), w, h, l, rad, 0, _p(out, _u8p))
        return out

    def imerode_z(self, I, rad, zdist):
        """Frangi::imerode(I,w,h,l,rad,zdist,E) (frangi.h:46)."""
        I, w, h, l = _check_vol(I)
        out = np.empty_like(I)
        self.lib.oracle_imerode_z.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _u8p]
        self.lib.oracle_imerode_z(_p(I, _u8p), w, h, l, rad, zdist, _p(out, _u8p))
        return out

    def imgaussian2d(self, I, sigma):
        """Frangi::imgaussian(I,w,h,sig,F) (frangi.h:44): the smoothed image hessian2d starts from."""
        return self.hessian2d(I, sigma)["F"]

    def imgaussian_xy(self, I, sigma):
        I, w, h, l = _check_vol(I)
        out = I.copy()
        self.lib.oracle_imgaussian_xy_u8(_p(out, _u8p), w, h, l, sigma)
        return out

    def seed_candidates(self, J8):
        """The pre-pass of extractSeeds (seed.cpp:574-632): per-layer range, candidate counts, ranked keys."""
        J8, w, h, l = _check_vol(J8)
        lo = np.empty(l, np.uint8); hi = np.empty(l, np.uint8); n = np.empty(l, np.int32)
        cap = J8.size
        keys = np.empty(cap, np.int64)
        tot = self.lib.oracle_seed_candidates(_p(J8, _u8p), w, h, l, _p(lo, _u8p), _p(hi, _u8p),
                                              n.ctypes.data_as(C.POINTER(C.c_int)),
                                              keys.ctypes.data_as(C.POINTER(C.c_int64)), cap)
        if tot < 0 or tot > cap:
            raise RuntimeError(f"oracle_seed_candidates returned {tot}")
        return dict(layer_min=lo, layer_max=hi, n_max=n, keys=keys[:tot].copy())


class Reference:
    """The unmodified reference (compiled in place).  Same conventions as Oracle."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not built: run `make -C oracle ref` where /root/reference exists")
        self.lib = L = C.CDLL(path)
        L.ref_imgaussian.restype = None
        L.ref_imgaussian.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _f32p]
        L.ref_hessian3d.restype = None
        L.ref_hessian3d.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float] + [_f32p] * 6
        L.ref_eigen3.restype = None
        L.ref_eigen3.argtypes = [_f64p, _f64p, _f64p]
        L.ref_frangi3d.restype = None
        L.ref_frangi3d.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_float,
                                   C.c_float, C.c_float, C.c_float, C.c_int,
                                   _f32p, _f32p, _f32p, _u8p, _u8p, _u8p]
        L.ref_extract_seeds.restype = C.c_long
        L.ref_extract_seeds.argtypes = [C.c_double, _u8p, C.c_int, C.c_int, C.c_int,
                                        _u8p, _u8p, _u8p, _f32p, C.c_long]
        self.has_2d = hasattr(L, "ref_frangi2d")
        if self.has_2d:
            L.ref_frangi2d.restype = None
            L.ref_frangi2d.argtypes = [_u8p, C.c_int, C.c_int, _f32p, C.c_int, C.c_float, C.c_float, C.c_int,
                                       _f32p, _f32p, _f32p, _u8p, _u8p, _u8p]
            L.ref_hessian2d.restype = None
            L.ref_hessian2d.argtypes = [_u8p, C.c_int, C.c_int, C.c_float, _f32p, _f32p, _f32p]
        self.has_soma = hasattr(L, "ref_imerode")
        if self.has_soma:
            L.ref_imerode.restype = None
            L.ref_imerode.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, _u8p]
            L.ref_imdilate.restype = None
            L.ref_imdilate.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float]
            L.ref_imgaussian_xy.restype = None
            L.ref_imgaussian_xy.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float]
        self.has_trace = hasattr(L, "ref_trace")
        if self.has_trace:
            L.ref_trace.restype = C.c_int
            L.ref_trace.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _u8p, _u8p, _f32p, C.c_int,
                                    C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                    C.c_int, _f32p, C.c_long, _f32p, C.c_long, C.POINTER(C.c_int), C.c_long,
                                    C.POINTER(C.c_long)]

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def imgaussian(self, I, sigma, zdist):
        I, w, h, l = _check_vol(I)
        F = np.empty(I.shape, np.float32)
        self.lib.ref_imgaussian(_p(I, _u8p), w, h, l, sigma, zdist, _p(F, _f32p))
        return F

    def hessian3d(self, I, sigma, zdist):
        I, w, h, l = _check_vol(I)
        names = ["Dzz", "Dyy", "Dyz", "Dxx", "Dxy", "Dxz"]
        D = [np.empty(I.shape, np.float32) for _ in names]
        self.lib.ref_hessian3d(_p(I, _u8p), w, h, l, sigma, zdist, *[_p(d, _f32p) for d in D])
        return dict(zip(names, D))

    def eigen3(self, A):
        A = np.ascontiguousarray(A, np.float64).reshape(9)
        V = np.empty(9, np.float64)
        d = np.empty(3, np.float64)
        self.lib.ref_eigen3(_p(A, _f64p), _p(V, _f64p), _p(d, _f64p))
        return V.reshape(3, 3), d

    def frangi3d(self, I, sigmas, zdist=2.0, alpha=0.5, beta=0.5, Cc=500.0, blackwhite=False):
        I, w, h, l = _check_vol(I)
        s = np.ascontiguousarray(sigmas, np.float32)
        J = np.empty(I.shape, np.float32)
        V = [np.empty(I.shape, np.uint8) for _ in range(3)]
        lo, hi = C.c_float(), C.c_float()
        self.lib.ref_frangi3d(_p(I, _u8p), w, h, l, _p(s, _f32p), len(s), zdist, alpha, beta, Cc,
                              int(blackwhite), _p(J, _f32p), C.byref(lo), C.byref(hi),
                              _p(V[0], _u8p), _p(V[1], _u8p), _p(V[2], _u8p))
        return dict(J=J, Jmin=lo.value, Jmax=hi.value, Vx=V[0], Vy=V[1], Vz=V[2])

    def extract_seeds(self, tolerance, J8, Vx, Vy, Vz):
        J8, w, h, l = _check_vol(J8)
        Vx, Vy, Vz = (np.ascontiguousarray(v, np.uint8) for v in (Vx, Vy, Vz))
        cap = max(1024, J8.size // 8)
        out = np.empty((cap, 6), np.float32)
        n = self.lib.ref_extract_seeds(tolerance, _p(J8, _u8p), w, h, l, _p(Vx, _u8p), _p(Vy, _u8p),
                                       _p(Vz, _u8p), _p(out, _f32p), cap)
        if n < 0 or n > cap:
            raise RuntimeError(f"ref_extract_seeds returned {n} (cap {cap})")
        return out[:n].copy()

    def frangi2d(self, I, sigmas, beta_one=0.5, beta_two=15.0, blackwhite=False):
        I = np.ascontiguousarray(I, np.uint8)
        h, w = I.shape
        s = np.ascontiguousarray(sigmas, np.float32)
        J = np.empty(I.shape, np.float32)
        V = [np.empty(I.shape, np.uint8) for _ in range(3)]
        lo, hi = C.c_float(), C.c_float()
        self.lib.ref_frangi2d(_p(I, _u8p), w, h, _p(s, _f32p), len(s), beta_one, beta_two, int(blackwhite),
                              _p(J, _f32p), C.byref(lo), C.byref(hi), _p(V[0], _u8p), _p(V[1], _u8p), _p(V[2], _u8p))
        return dict(J=J, Jmin=lo.value, Jmax=hi.value, Vx=V[0], Vy=V[1], Vz=V[2])

    def hessian2d(self, I, sigma):
        I = np.ascontiguousarray(I, np.uint8)
        h, w = I.shape
        D = {k: np.empty(I.shape, np.float32) for k in ("Dyy", "Dxy", "Dxx")}
        self.lib.ref_hessian2d(_p(I, _u8p), w, h, sigma, _p(D["Dyy"], _f32p), _p(D["Dxy"], _f32p), _p(D["Dxx"], _f32p))
        return D

    def imerode(self, I, rad):
        I, w, h, l = _check_vol(I)
        out = np.empty_like(I)
        self.lib.ref_imerode(_p(I, _u8p), w, h, l, rad, _p(out, _u8p))
        return out

    def imdilate(self, I, rad):
        I, w, h, l = _check_vol(I)
        out = I.copy()
        self.lib.ref_imdilate(_p(out, _u8p), w, h, l, rad)
        return out

    def imgaussian_xy(self, I, sigma):
        I, w, h, l = _check_vol(I)
        out = I.copy()
        self.lib.ref_imgaussian_xy(_p(out, _u8p), w, h, l, sigma)
        return out

    # ---- the members no live code calls (frangi.h:28-31,44,46,51); present when oracle/_ref was built from this tree
    @property
    def has_cold(self):
        return hasattr(self.lib, "ref_imerode_z")

    def imerode_z(self, I, rad, zdist):
        I, w, h, l = _check_vol(I)
        out = np.empty_like(I)
        f = self.lib.ref_imerode_z
        f.restype = None
        f.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _u8p]
        f(_p(I, _u8p), w, h, l, rad, zdist, _p(out, _u8p))
        return out

    def imgaussian2d(self, I, sigma):
        I = np.ascontiguousarray(I, np.uint8)
        h, w = I.shape
        F = np.empty(I.shape, np.float32)
        f = self.lib.ref_imgaussian2d
        f.restype = None
        f.argtypes = [_u8p, C.c_int, C.c_int, C.c_float, _f32p]
        f(_p(I, _u8p), w, h, sigma, _p(F, _f32p))
        return F

    @property
    def has_zncc(self):
        return hasattr(self.lib, "ref_seed_zncc")

    def seed_zncc(self, I, sigmas, seeds):
        """Tracker::znccBBB (tracker.cpp:1891-1964) for rows (x, y, z, vx, vy, vz): (corr[n], sigma[n])."""
        I, w, h, l = _check_vol(I)
        seeds = np.ascontiguousarray(np.asarray(seeds, np.float32)[:, :6])
        s = np.ascontiguousarray(sigmas, np.float32)
        corr = np.empty(len(seeds), np.float32)
        sig = np.empty(len(seeds), np.float32)
        f = self.lib.ref_seed_zncc
        f.restype = None
        f.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, _f32p, C.c_long, _f32p, _f32p]
        f(_p(I, _u8p), w, h, l, _p(s, _f32p), len(s), _p(seeds, _f32p), len(seeds), _p(corr, _f32p), _p(sig, _f32p))
        return corr, sig

    def unit_directions(self, three_d, ndir):
        out = np.empty((ndir, 3), np.float32)
        f = self.lib.ref_unit_directions
        f.restype = None
        f.argtypes = [C.c_int, C.c_int, _f32p]
        f(int(three_d), ndir, _p(out, _f32p))
        return out

    def direction_idx(self, v, table):
        table = np.ascontiguousarray(table, np.float32)
        f = self.lib.ref_direction_idx
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, _f32p, C.c_int]
        three_d = len(v) == 3
        return f(int(three_d), v[0], v[1], v[2] if three_d else 0.0, _p(table, _f32p), len(table))

    def interpz(self, x, y, z, img):
        img = np.ascontiguousarray(img, np.float32)
        l, h, w = img.shape
        f = self.lib.ref_interpz
        f.restype = C.c_float
        f.argtypes = [C.c_int, C.c_int, C.c_float, _f32p, C.c_int, C.c_int, C.c_int]
        return f(x, y, z, _p(img, _f32p), w, h, l)

    def trace(self, I, J8, Vx, Vy, Vz, sigmas, tolerance=5.0, znccth=0.3, kappa=3.0, step=2, ni=200, np_=20,
              zdist=2.0, nodepervol=4, max_traces=5000):
        """The plugin's pipeline downstream of Frangi (seeds -> correlation filter -> SMC traces), README usage
        `-p 2,4,6 0 5 0.3 3 2 200 20 2 4 1`.  Returns dict(seeds [n,8], nodes [m,10], nbr int[], counts)."""
        I, w, h, l = _check_vol(I)
        J8, Vx, Vy, Vz = (np.ascontiguousarray(v, np.uint8) for v in (J8, Vx, Vy, Vz))
        s = np.ascontiguousarray(sigmas, np.float32)
        scap = max(1024, I.size // 8)
        ncap = max(1 << 16, I.size // 4)
        seeds = np.empty((scap, 8), np.float32)
        nodes = np.empty((ncap, 10), np.float32)
        nbr = np.empty(4 * ncap, np.int32)
        counts = (C.c_long * 5)()
        rc = self.lib.ref_trace(_p(I, _u8p), w, h, l, _p(J8, _u8p), _p(Vx, _u8p), _p(Vy, _u8p), _p(Vz, _u8p),
                                _p(s, _f32p), len(s), tolerance, znccth, kappa, step, ni, np_, zdist, nodepervol,
                                max_traces, _p(seeds, _f32p), scap, _p(nodes, _f32p), ncap,
                                nbr.ctypes.data_as(C.POINTER(C.c_int)), nbr.size, counts)
        c = [int(v) for v in counts]
        if rc != 0 or c[1] > scap or c[2] > ncap or c[3] > nbr.size:
            raise RuntimeError(f"ref_trace rc={rc} counts={c}")
        return dict(seeds=seeds[:c[1]].copy(), nodes=nodes[:c[2]].copy(), nbr=nbr[:c[3]].copy(),
                    n_extracted=c[0], n_traces=c[4])


PLUGIN_REF_SO = os.path.join(_HERE, "_ref", "libpnr_plugin_ref.so")
PLUGIN_GPU_SO = os.path.join(_HERE, "_ref", "libpnr_plugin_gpu.so")
PLUGIN_REPLAY_SO = os.path.join(_HERE, "_ref", "libpnr_plugin_replay.so")
_PLUGIN_SO = dict(ref=PLUGIN_REF_SO, gpu=PLUGIN_GPU_SO, replay=PLUGIN_REPLAY_SO)


class Plugin:
    """The reference's WHOLE plugin translation unit (Advantra_plugin.cpp, unmodified, against the Qt / Vaa3D stand-ins of
    oracle/stubs/; oracle/plugin_wrap.cpp), driven through its batch entry point Advantra::dofunc("advantra_func").
    arm = "ref": the reference's own Frangi; arm = "gpu": the same unchanged call site with the drop-in class Frangi of
    pnr_b200/csrc/frangi.h (needs a GPU at run time); arm = "replay": the call site is handed filter outputs the caller
    supplies (run(..., replay=(J8, Vx, Vy, Vz))) -- outputs captured on a GPU box, or the reference's own with some
    eigenvector signs turned -- every other Frangi member being the reference's.  Everything the plugin writes lands
    under `workdir`."""

    PARAMS = ("2,4,6", "0", "5", "0.3", "3", "2", "200", "20", "2", "4", "1")     # the README's usage line

    def __init__(self, arm: str = "ref"):
        path = _PLUGIN_SO[arm]
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.arm = arm
        self.lib = C.CDLL(path)
        self.lib.plugin_run.restype = C.c_int
        self.lib.plugin_run.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_char_p), C.c_int,
                                        C.c_int, C.c_int, C.c_int]

    @staticmethod
    def available(arm: str = "ref") -> bool:
        return os.path.exists(_PLUGIN_SO[arm])

    def run(self, I, workdir, params=None, save_midres=True, single_tree=True, max_traces=0, name="img", replay=None):
        """Runs the plugin on I[l][h][w]; returns {file name without the prefix: text} for every file it wrote."""
        I, w, h, l = _check_vol(I)
        if (self.arm == "replay") != (replay is not None):
            raise ValueError("replay arrays go with the replay arm, and only with it")
        if replay is not None:
            keep = [np.ascontiguousarray(v, np.uint8) for v in replay]
            if len(keep) != 4 or any(v.shape != I.shape for v in keep):
                raise ValueError("replay = (J8, Vx, Vy, Vz), each of the volume's shape")
            self.lib.plugin_set_replay.restype = None
            self.lib.plugin_set_replay.argtypes = [_u8p] * 4
            self.lib.plugin_set_replay(*[_p(v, _u8p) for v in keep])
        p = [str(v) for v in (params or self.PARAMS)]
        arr = (C.c_char_p * len(p))(*[s.encode() for s in p])
        prefix = os.path.join(str(workdir), name)
        n = self.lib.plugin_run(_p(I, _u8p), w, h, l, prefix.encode(), arr, len(p), int(save_midres), int(single_tree),
                                int(max_traces))
        if n < 0:
            raise RuntimeError("Advantra::dofunc refused the arguments")
        out = {}
        for f in sorted(os.listdir(str(workdir))):
            if f.startswith(name):
                with open(os.path.join(str(workdir), f)) as fh:
                    out[f[len(name):]] = fh.read()
        return out
