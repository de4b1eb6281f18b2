"""The render kernels themselves on the CPU.  tests/host_emul/emul_warp.h is a SIMT shim (every lane a fiber, the lanes
of a warp switch at each warp collective) behind which g++ compiles render_kernels.cuh unchanged; emul_binned.cpp
launches one block of four warps.  This checks the WARP-LEVEL logic that test_host_emul.py cannot see -- ticket queue,
bins, path pool, sample staging, lane regeneration, chunk epilogues -- against the oracle, without a GPU.
Test harness only: the shipped library has no CPU execution path."""
import os

import numpy as np
import pytest

import oracle_util
from miniraytracer_b200 import accfile

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")


def _same(acc, ref, rays, ref_rays):
    assert rays == ref_rays
    np.testing.assert_array_equal(acc[..., 3], ref[..., 3])
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=1e-5)
    assert res["n_bad"] == 0, res


@needs_ref
@pytest.mark.parametrize("scene,w,h,spp,bins", [(5, 24, 13, 36, 1), (5, 24, 13, 36, 2), (5, 24, 13, 36, 3), (6, 20, 11, 36, 2),
                                                 (7, 20, 11, 16, 2), (8, 16, 9, 16, 2), (0, 16, 16, 9, 2), (5, 7, 3, 1, 2),
                                                 (5, 6, 4, 169, 2), (5, 3, 2, 1089, 2)])   # > 128 samples per pixel: the lane-strided sums
def test_mode_b_kernel_matches_oracle(emul_kernel_bin, scene, w, h, spp, bins):
    ref, rmeta = oracle_util.ref_render(scene, w, h, spp)
    acc, meta = oracle_util.emul_binned(emul_kernel_bin, scene, w, h, spp, bins=bins)
    assert meta["info"]["tasks"] > 4 or w * h < 32          # several tickets per warp
    _same(acc, ref, meta["rays"], rmeta["rays"])


@pytest.mark.parametrize("scene,w,h,spp", [(5, 9, 5, 36), (8, 6, 4, 16)])
def test_mode_b_sums_few_samples_in_the_reference_order(emul_kernel_bin, scene, w, h, spp):
    """Up to 128 samples per pixel mode B (the SEQ instantiation) adds a pixel's samples one after the other in sample order,
    the order of the reference's own loop (main.cpp:154-166): the accumulator EQUALS the float32 running sum of the samples
    rendered one by one."""
    acc, _ = oracle_util.emul_binned(emul_kernel_bin, scene, w, h, spp, bins=2)
    run = np.zeros_like(acc)
    for s in range(spp):
        one, _ = oracle_util.emul_binned(emul_kernel_bin, scene, w, h, spp, bins=2, s0=s, s1=s + 1)
        run = (run + one).astype(np.float32)
    np.testing.assert_array_equal(acc, run)


@needs_ref
def test_mode_b_result_is_schedule_independent(emul_kernel_bin):
    """Finished samples are summed per pixel in item order from the staging array: chunk size and bins do not change a bit."""
    base, _ = oracle_util.emul_binned(emul_kernel_bin, 6, 20, 11, 64, bins=1, chunk=1)
    for bins, chunk in ((2, 1), (3, 2), (2, 5), (2, 128)):
        acc, meta = oracle_util.emul_binned(emul_kernel_bin, 6, 20, 11, 64, bins=bins, chunk=chunk)
        np.testing.assert_array_equal(acc, base)


@needs_ref
def test_mode_b_with_the_shipped_planner(emul_kernel_bin):
    """-plan W: the ticket queue laid out by schedule.h (the planner mrt_gpu_render_async calls) for W resident warps -- big chunks,
    then quarter and sixteenth chunks -- gives the same accumulator as one-pixel chunks."""
    base, _ = oracle_util.emul_binned(emul_kernel_bin, 6, 40, 22, 16, bins=2, chunk=1)
    for warps, extra in ((1, ()), (2, ("-chunkpaths", "512")), (4, ())):
        acc, meta = oracle_util.emul_binned(emul_kernel_bin, 6, 40, 22, 16, bins=2, extra=("-plan", str(warps)) + extra)
        np.testing.assert_array_equal(acc, base)


@needs_ref
def test_mode_b_sample_slice(emul_kernel_bin):
    ref, rmeta = oracle_util.ref_render(5, 20, 11, 49, s0=10, s1=37)
    acc, meta = oracle_util.emul_binned(emul_kernel_bin, 5, 20, 11, 49, s0=10, s1=37)
    _same(acc, ref, meta["rays"], rmeta["rays"])


@needs_ref
@pytest.mark.parametrize("mode,scene,w,h,spp", [("W", 5, 20, 11, 36), ("W", 7, 16, 9, 36), ("P", 5, 40, 11, 9), ("P", 0, 33, 7, 4)])
def test_modes_w_and_p_kernels_match_oracle(emul_kernel_bin, mode, scene, w, h, spp):
    ref, rmeta = oracle_util.ref_render(scene, w, h, spp)
    acc, meta = oracle_util.emul_binned(emul_kernel_bin, scene, w, h, spp, mode=mode)
    _same(acc, ref, meta["rays"], rmeta["rays"])


def test_mode_b_kernel_matches_golden_without_the_reference(emul_kernel_bin):
    """Same check against a committed golden fixture, so it also runs where oracle/_ref is absent."""
    g = np.load(os.path.join(GOLDEN, "golden_scene5.npz"))
    w, h, spp, depth = int(g["width"]), int(g["height"]), int(g["spp"]), int(g["depth"])
    if w * h * spp > 400000:
        pytest.skip("golden frame too large for the fiber emulation")
    acc, meta = oracle_util.emul_binned(emul_kernel_bin, 5, w, h, spp, depth=depth)
    _same(acc, g["acc"], meta["rays"], int(g["rays"]))


@needs_ref
@pytest.mark.parametrize("scene,w,h,spp,crop,s0,s1", [(8, 3840, 2160, 4096, (1900, 2040, 1906, 2042), 4000, 4024),
                                                        (7, 1920, 1080, 4096, (1130, 660, 1136, 662), 100, 124)])
def test_mode_b_crop_window_at_true_config_size(emul_kernel_bin, scene, w, h, spp, crop, s0, s1):
    """The kernels' crop window (MrtRenderParams.crop_*): window pixel -> frame (x, y) -> stream id above 2^32 / 2^34."""
    ref, rmeta = oracle_util.ref_render(scene, w, h, spp, s0=s0, s1=s1, crop=crop)
    acc, meta = oracle_util.emul_binned(emul_kernel_bin, scene, w, h, spp, s0=s0, s1=s1, crop=crop)
    _same(acc, ref, meta["rays"], rmeta["rays"])


@needs_ref
@pytest.mark.parametrize("scene,w,h,spp", [(8, 16, 9, 16), (0, 16, 16, 9), (1, 12, 12, 9), (7, 20, 11, 16)])
def test_cooperative_tree_traversal_matches_oracle_and_per_lane(emul_kernel_bin, scene, w, h, spp):
    """Warp-cooperative BVH traversal (coop_tree.cuh) in the CPU SIMT emulation: same trace() count as the oracle, accumulator
    bit-identical to the per-lane depth-first traversal (every box / primitive test is the same function on the same operands;
    the leaf with the lowest depth-first rank wins)."""
    ref, rmeta = oracle_util.ref_render(scene, w, h, spp)
    coop, meta = oracle_util.emul_binned(emul_kernel_bin, scene, w, h, spp, extra=["-coop", "1"])
    lane, meta0 = oracle_util.emul_binned(emul_kernel_bin, scene, w, h, spp)
    assert meta["info"]["coop_node_steps"] > 0 and meta["info"]["coop_leaf_steps"] > 0 and meta0["info"]["coop_node_steps"] == 0
    np.testing.assert_array_equal(coop, lane)
    _same(coop, ref, meta["rays"], rmeta["rays"])


@needs_ref
def test_mode_b_guided_schedule_does_not_change_a_bit(emul_kernel_bin):
    """Guided self-scheduling (big chunks first, the last pixels in chunks of a quarter / a sixteenth of the size) only changes
    which warp task a pixel belongs to; samples are summed per pixel in item order, so the accumulator is bit-identical."""
    base, m0 = oracle_util.emul_binned(emul_kernel_bin, 6, 20, 11, 16, chunk=16)
    for tail in (37, 220):
        acc, m = oracle_util.emul_binned(emul_kernel_bin, 6, 20, 11, 16, chunk=16, extra=["-tail", str(tail)])
        assert m["info"]["tasks"] > m0["info"]["tasks"]
        np.testing.assert_array_equal(acc, base)
