"""z-slab decomposition of the Frangi volume across GPUs (host-side planning).

Every stage of frangi3d is local in z: the x and y smoothing passes work per
plane (frangi.cpp:683-748), the z pass has radius Lz = ceil(3 sigma / zdist)
(frangi.cpp:751-782, radius at :670) and the twice-applied central difference
adds 2 planes (frangi.cpp:306-381), so a slab needs Lz + 2 planes of its z
neighbours per scale; only Jmin / Jmax are global (frangi.cpp:237-238,257-258).
The same arithmetic lives in pnr_b200/csrc/frangi_gpu.cu (alloc_slab,
exchange_halos); this module is its Python mirror for bench.py, the tests and
multi-process launchers.
"""
from __future__ import annotations

import math


def gauss_radius(sigma: float) -> int:
    """Tap radius of the reference's truncated Gaussian: ceil(3 sigma) (frangi.cpp:651,654)."""
    return int(math.ceil(3 * sigma))


def z_radius(sigma: float, zdist: float) -> int:
    """Radius of the z pass: the kernel of sigma / zdist (frangi.cpp:649,670)."""
    return gauss_radius(sigma / zdist)


def halo_planes(sigma: float, zdist: float) -> int:
    """Planes a slab needs from each z neighbour for one scale: z radius + 2."""
    return z_radius(sigma, zdist) + 2


def slab_ranges(l: int, n: int) -> list[tuple[int, int]]:
    """Contiguous z ranges [z0, z1) of n slabs tiling [0, l) (rank k owns planes l*k//n .. l*(k+1)//n)."""
    if n < 1 or l < n:
        raise ValueError(f"cannot cut {l} planes into {n} slabs")
    return [(l * k // n, l * (k + 1) // n) for k in range(n)]


def max_slabs(l: int, sigmas, zdist: float) -> int:
    """Largest slab count for which every slab can supply a whole halo to its neighbours."""
    need = max(halo_planes(s, zdist) for s in sigmas)
    return max(1, l // need)


def exchange_plan(l: int, n: int, rank: int, sigma: float, zdist: float):
    """Planes of the xy-smoothed volume that `rank` sends and receives for one scale.

    Returns dict(send_down=(a,b), recv_down=(a,b), send_up=(a,b), recv_up=(a,b)) of global plane
    ranges, None where there is no neighbour.  Received ranges are clipped to the volume."""
    z0, z1 = slab_ranges(l, n)[rank]
    H = halo_planes(sigma, zdist)
    if n > 1 and (z1 - z0) < H:
        raise ValueError(f"slab of {z1 - z0} planes is thinner than the halo {H}")
    plan = dict(send_down=None, recv_down=None, send_up=None, recv_up=None)
    if rank > 0:
        plan["send_down"] = (z0, z0 + H)
        plan["recv_down"] = (max(z0 - H, 0), z0)
    if rank < n - 1:
        plan["send_up"] = (z1 - H, z1)
        plan["recv_up"] = (z1, min(z1 + H, l))
    return plan
