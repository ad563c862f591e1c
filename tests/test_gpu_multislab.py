"""z-slab decomposition on the GPU: N slabs must reproduce the one-slab result bit for bit
(same kernels, same per-voxel arithmetic; only the tiling changes -- SURVEY.md section 8e).
On a one-GPU box the slabs share device 0 and halos move by device-to-device copies; with
two or more GPUs the same test also runs one slab per device over NCCL send/recv."""
import numpy as np
import pytest

import pnr_b200
from pnr_b200 import FLAG_DIR_F32, FLAG_FMA_SMOOTHING, FLAG_LOCAL_HALO, FLAG_SCALE_IDX, FrangiPlan
from pnr_b200.synth import make_volume

pytestmark = pytest.mark.gpu

KEYS = ("J", "Vx", "Vy", "Vz", "J8", "scale", "dir")


def _run(I, sigs, devices, flags):
    l, h, w = I.shape
    p = FrangiPlan(sigs, 2.0, .5, .5, 500., False, w, h, l, devices=devices, flags=flags)
    r = p.run(I, want_J8=True)
    p.close()
    return r


@pytest.mark.parametrize("nslabs", [2, 3, 4])
@pytest.mark.parametrize("fma", [False, True])
def test_slabs_on_one_device_equal_single_slab(nslabs, fma):
    I = make_volume(150, 70, 48, seed=33, n_neurites=6)      # ragged width, 48 planes: up to 4 slabs of >= 11
    sigs = [2.0, 4.0, 6.0]
    flags = FLAG_DIR_F32 | FLAG_SCALE_IDX | (FLAG_FMA_SMOOTHING if fma else 0)
    one = _run(I, sigs, (0,), flags)
    many = _run(I, sigs, (0,) * nslabs, flags)
    assert one["Jmax"] > 0
    assert many["Jmin"] == one["Jmin"] and many["Jmax"] == one["Jmax"]
    for k in KEYS:
        assert np.array_equal(one[k], many[k]), k


@pytest.mark.parametrize("nslabs", [2, 3])
def test_six_scales_template_radius_above_true_radius(nslabs):
    """sigma = 1, 3, 5 have z radii 2, 5, 8 but run in the radius-3, 6, 9 kernels: the z pass then reads planes with
    zero taps, possibly halo planes still in flight; two runs on one handle, so stale halos of the last scale are
    there when the first scale of the second run starts."""
    I = make_volume(96, 64, 60, seed=35, n_neurites=6)
    sigs = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    flags = FLAG_DIR_F32 | FLAG_SCALE_IDX
    one = _run(I, sigs, (0,), flags)
    l, h, w = I.shape
    p = FrangiPlan(sigs, 2.0, .5, .5, 500., False, w, h, l, devices=(0,) * nslabs, flags=flags)
    p.run(I, want_J8=True)
    many = p.run(I, want_J8=True)
    p.close()
    assert many["Jmin"] == one["Jmin"] and many["Jmax"] == one["Jmax"]
    for k in KEYS:
        assert np.array_equal(one[k], many[k]), k


def test_too_many_slabs_are_reduced_to_what_the_halo_allows():
    I = make_volume(40, 36, 24, seed=2, n_neurites=3)         # 24 planes, halo 11 -> at most 2 slabs
    one = _run(I, [6.0], (0,), 0)
    many = _run(I, [6.0], (0,) * 8, 0)
    assert np.array_equal(one["J"], many["J"]) and np.array_equal(one["Vz"], many["Vz"])


def test_slabs_over_real_devices_nccl_and_peer_copies():
    n = pnr_b200.load_library().frangi_gpu_device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    I = make_volume(160, 96, 64, seed=34, n_neurites=8)
    sigs = [2.0, 4.0, 6.0]
    flags = FLAG_DIR_F32 | FLAG_SCALE_IDX
    one = _run(I, sigs, (0,), flags)
    for devs, fl in (((0, 1), flags), ((0, 1), flags | FLAG_LOCAL_HALO), (tuple(range(min(n, 4))), flags)):
        many = _run(I, sigs, devs, fl)
        assert many["Jmax"] == one["Jmax"]
        for k in KEYS:
            assert np.array_equal(one[k], many[k]), (devs, fl, k)


@pytest.mark.parametrize("chunk", [11, 16, 24])
def test_pipelined_run_equals_one_piece_run(chunk):
    """frangi_gpu_run pipelines copies and kernels over z chunks; the chunks are views of the slab
    and must give the one-piece result bit for bit (J, V, J8, Jmin, Jmax, extras)."""
    I = make_volume(150, 70, 48, seed=35, n_neurites=6)
    sigs = [2.0, 4.0, 6.0]
    flags = FLAG_DIR_F32 | FLAG_SCALE_IDX
    l, h, w = I.shape
    p = FrangiPlan(sigs, 2.0, .5, .5, 500., False, w, h, l, flags=flags)
    p.set_stream_chunk(0)
    one = p.run(I, want_J8=True)
    p.set_stream_chunk(chunk)
    before = pnr_b200.launch_count()
    many = p.run(I, want_J8=True)
    assert pnr_b200.launch_count() - before > 3 * 4 + 1, "the chunked path should launch per chunk"
    p.close()
    assert many["Jmin"] == one["Jmin"] and many["Jmax"] == one["Jmax"]
    for k in KEYS:
        assert np.array_equal(one[k], many[k]), k
