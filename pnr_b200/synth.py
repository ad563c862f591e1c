"""Seeded synthetic tubular-neuron volumes (uint8, [l][h][w], x fastest).

This is the input generator for tests/ and bench.py: the
reference's sample images are only linked by URL (README.md:11-13) and there is
no network.  K random-walk "neurites" with smoothly varying radius (1..6 voxels,
so every sigma in 1..6 wins somewhere) and peak intensity (80..220, which breaks
plateaus so that seed extraction finds maxima), Gaussian cross-section squeezed
along z by `zdist` (the stack's z spacing in pixels, Advantra_plugin.cpp:56),
plus a background offset and additive uniform noise, saturated to 255.

A z-range can be generated on its own (multi-GPU ranks build only their slab):
paths are drawn globally from the seed, noise is seeded per z-plane, so any
slab equals the same planes of the whole volume.
"""
from __future__ import annotations

import hashlib

import numpy as np

DEFAULT_SEED = 20181009


def _paths(w, h, l, seed, n_neurites, zdist):
    rng = np.random.default_rng([seed, 0])
    k = n_neurites if n_neurites is not None else max(2, w // 16)
    paths = []
    for _ in range(k):
        n_steps = int(rng.integers(max(8, (w + h) // 4), max(16, w + h)))
        step = 2.0
        pos = np.array([rng.uniform(0, w), rng.uniform(0, h), rng.uniform(0, l * zdist)])
        d = rng.normal(size=3)
        d[2] *= 0.5
        d /= np.linalg.norm(d)
        radius = rng.uniform(1.0, 6.0)
        amp = rng.uniform(80.0, 220.0)
        pts = np.empty((n_steps, 5))
        for s in range(n_steps):
            pts[s] = (pos[0], pos[1], pos[2] / zdist, radius, amp)
            d = d + rng.normal(scale=0.18, size=3)
            d /= np.linalg.norm(d)
            pos = pos + step * d
            # reflect at the (physical) volume faces so paths stay inside
            for ax, hi in ((0, w - 1.0), (1, h - 1.0), (2, (l - 1.0) * zdist)):
                if pos[ax] < 0:
                    pos[ax] = -pos[ax]; d[ax] = -d[ax]
                elif pos[ax] > hi:
                    pos[ax] = 2 * hi - pos[ax]; d[ax] = -d[ax]
            radius = float(np.clip(radius + rng.normal(scale=0.08), 1.0, 6.0))
            amp = float(np.clip(amp + rng.normal(scale=4.0), 80.0, 220.0))
        paths.append(pts)
    return paths


def make_volume(w: int, h: int, l: int, seed: int = DEFAULT_SEED, zdist: float = 2.0,
                z_range: tuple[int, int] | None = None, n_neurites: int | None = None,
                noise: int = 15, background: int = 5) -> np.ndarray:
    """Returns planes [z0, z1) of the w x h x l volume as uint8 [z1-z0][h][w]."""
    z0, z1 = (0, l) if z_range is None else z_range
    vol = np.zeros((z1 - z0, h, w), np.float32)
    for pts in _paths(w, h, l, seed, n_neurites, zdist):
        for (px, py, pz, r, a) in pts:
            rx = int(np.ceil(3 * r))
            rz = int(np.ceil(3 * r / zdist))
            cz, cy, cx = int(round(pz)), int(round(py)), int(round(px))
            za, zb = max(cz - rz, z0), min(cz + rz + 1, z1)
            if za >= zb:
                continue
            ya, yb = max(cy - rx, 0), min(cy + rx + 1, h)
            xa, xb = max(cx - rx, 0), min(cx + rx + 1, w)
            if ya >= yb or xa >= xb:
                continue
            dz = ((np.arange(za, zb, dtype=np.float32) - np.float32(pz)) * np.float32(zdist)) ** 2
            dy = (np.arange(ya, yb, dtype=np.float32) - np.float32(py)) ** 2
            dx = (np.arange(xa, xb, dtype=np.float32) - np.float32(px)) ** 2
            d2 = dz[:, None, None] + dy[None, :, None] + dx[None, None, :]
            blob = np.float32(a) * np.exp(-d2 / np.float32(2 * r * r))
            sub = vol[za - z0:zb - z0, ya:yb, xa:xb]
            np.maximum(sub, blob, out=sub)
    out = np.empty(vol.shape, np.uint8)
    for z in range(z0, z1):
        rng = np.random.default_rng([seed, 1, z])
        nz = rng.integers(0, noise + 1, size=(h, w), dtype=np.int32) if noise > 0 else 0
        plane = vol[z - z0] + np.float32(background) + nz
        np.clip(plane, 0, 255, out=plane)
        out[z - z0] = plane.astype(np.uint8)
    return out


def straight_tube(w=64, h=64, l=32, cy=32, cz=16, zdist=2.0, amp=200.0, sigma=3.0) -> np.ndarray:
    """The x-aligned Gaussian tube of SURVEY.md section 8c (known-answer vectors):
    I[z,y,x] = (uint8)(amp * expf(-((y-cy)^2 + ((z-cz)*zdist)^2) / (2*sigma^2))), float32 math, truncation."""
    y = np.arange(h, dtype=np.float32)[None, :]
    z = np.arange(l, dtype=np.float32)[:, None]
    num = (y - np.float32(cy)) ** 2 + ((z - np.float32(cz)) * np.float32(zdist)) ** 2
    den = np.float32(2.0) * np.float32(sigma) * np.float32(sigma)
    plane = (np.float32(amp) * np.exp(-(num / den)).astype(np.float32)).astype(np.float32)
    vol = np.repeat(plane.astype(np.uint8)[:, :, None], w, axis=2)
    return np.ascontiguousarray(vol)


def volume_hash(vol: np.ndarray) -> str:
    """Short content hash recorded with every result."""
    return hashlib.blake2b(np.ascontiguousarray(vol).tobytes(), digest_size=8).hexdigest()
