"""pnr_b200 -- B200-native multi-scale 3-D Frangi filter (the hot path of
miroslavradojevic/pnr's Advantra plugin) behind the reference's Frangi class
interface.  CUDA kernels + C-ABI live in csrc/ (built into _lib/); this package
is the thin Python host mirror used by tests and benchmarks."""
from .frangi import (FLAG_DIR_F32, FLAG_FMA_SMOOTHING, FLAG_LOCAL_HALO, FLAG_OVERLAP_Z, FLAG_REFERENCE_DIRECTION, FLAG_SCALE_IDX, Frangi, FrangiGpuError,
                     FrangiPlan, PinnedBuffer, imdilate, imerode, imgaussian_xy,
                     launch_count, load_library, seed_candidates, seed_zncc, imerode_z, imgaussian2d)

__all__ = ["Frangi", "FrangiPlan", "FrangiGpuError", "PinnedBuffer", "load_library", "launch_count", "seed_candidates", "seed_zncc", "imerode_z", "imgaussian2d", "imerode", "imdilate", "imgaussian_xy",
           "FLAG_FMA_SMOOTHING", "FLAG_DIR_F32", "FLAG_SCALE_IDX", "FLAG_LOCAL_HALO", "FLAG_OVERLAP_Z", "FLAG_REFERENCE_DIRECTION"]
