"""Logic check of the device tracer core without a GPU: trace_core.h (the code the CUDA kernels are
built from) is compiled with g++ by the test-suite and rendered against the oracle.  This is a test
harness only -- the shipped library has no CPU execution path."""
import os

import numpy as np
import pytest

import oracle_util
from miniraytracer_b200 import accfile

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")


@pytest.mark.parametrize("scene", range(9))
def test_core_matches_golden(emul_bin, scene):
    g = np.load(os.path.join(GOLDEN, f"golden_scene{scene}.npz"))
    acc, meta = oracle_util.emul_render(emul_bin, scene, int(g["width"]), int(g["height"]), int(g["spp"]), int(g["depth"]))
    assert meta["rays"] == int(g["rays"])            # same number of trace() calls
    np.testing.assert_array_equal(acc[..., 3], g["acc"][..., 3])   # same samples dropped as non-finite
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(g["acc"]), rel=1e-5)
    assert res["n_bad"] == 0, res


@needs_ref
@pytest.mark.parametrize("scene,w,h,spp", [(0, 160, 160, 16), (5, 192, 108, 16), (6, 192, 108, 16), (7, 192, 108, 16), (8, 192, 108, 16)])
def test_core_matches_oracle_configs(emul_bin, scene, w, h, spp):
    ref, rmeta = oracle_util.ref_render(scene, w, h, spp)
    acc, meta = oracle_util.emul_render(emul_bin, scene, w, h, spp)
    assert meta["rays"] == rmeta["rays"]
    np.testing.assert_array_equal(acc[..., 3], ref[..., 3])
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=1e-5)
    assert res["n_bad"] == 0, res


@needs_ref
def test_sample_slices_compose(emul_bin):
    # spp sharding: slices [0,8) + [8,16) of the streams == the full render
    full, _ = oracle_util.ref_render(5, 96, 54, 16)
    a, _ = oracle_util.emul_render(emul_bin, 5, 96, 54, 16, s0=0, s1=8)
    b, _ = oracle_util.emul_render(emul_bin, 5, 96, 54, 16, s0=8, s1=16)
    res = accfile.compare(accfile.finalize(a + b), accfile.finalize(full), rel=1e-5)
    assert res["n_bad"] == 0, res


@needs_ref
def test_depth_limit_and_seed(emul_bin):
    for depth, seed in ((0, oracle_util.DEFAULT_SEED), (3, 12345)):
        ref, rmeta = oracle_util.ref_render(0, 80, 80, 4, depth=depth, seed=seed)
        acc, meta = oracle_util.emul_render(emul_bin, 0, 80, 80, 4, depth=depth, seed=seed)
        assert meta["rays"] == rmeta["rays"]
        res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=1e-5)
        assert res["n_bad"] == 0, res


@pytest.mark.parametrize("scene,w,h,spp", [(5, 160, 90, 16), (6, 160, 90, 16), (7, 128, 72, 9)])
def test_translate_cull_box_is_conservative(emul_bin, scene, w, h, spp):
    """The flattener gives every translate node an inflated parent-frame box and the traversal skips the node for
    rays that miss it (trace_core.h: cull_miss).  The reference tests nothing there, so the cull must never change a
    result: with and without it the accumulators are bit-identical, and fewer transforms are entered."""
    a, ma = oracle_util.emul_render(emul_bin, scene, w, h, spp)
    b, mb = oracle_util.emul_render(emul_bin, scene, w, h, spp, extra=["-nocull"])
    assert ma["rays"] == mb["rays"]
    np.testing.assert_array_equal(a, b)
    assert ma["counters"]["xform"] < mb["counters"]["xform"]


# crop windows of the BASELINE configurations at their true sizes: PCG32 stream ids (y*W+x)*N+s above 2^32 (C4) and
# 2^34 (C5) -- pcg.cpp:28-35, main.cpp:156-157 -- on a handful of pixels and a slice of the samples
BIG_WINDOWS = [(8, 3840, 2160, 4096, (1900, 2040, 1906, 2042), 4000, 4024), (7, 1920, 1080, 4096, (1130, 660, 1136, 662), 100, 124),
               (5, 1920, 1080, 1024, (950, 540, 956, 542), 1000, 1024)]


@needs_ref
@pytest.mark.parametrize("scene,w,h,spp,crop,s0,s1", BIG_WINDOWS)
def test_core_crop_window_at_true_config_size(emul_bin, scene, w, h, spp, crop, s0, s1):
    ref, rmeta = oracle_util.ref_render(scene, w, h, spp, s0=s0, s1=s1, crop=crop)
    acc, meta = oracle_util.emul_render(emul_bin, scene, w, h, spp, s0=s0, s1=s1, crop=crop)
    assert acc.shape == (crop[3] - crop[1], crop[2] - crop[0], 4)
    assert meta["rays"] == rmeta["rays"] and rmeta["rays"] > acc.shape[0] * acc.shape[1] * (s1 - s0)   # paths do bounce there
    np.testing.assert_array_equal(acc[..., 3], ref[..., 3])
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=1e-5)
    assert res["n_bad"] == 0, res


@needs_ref
@pytest.mark.parametrize("scene,w,h,spp", [(5, 96, 54, 16), (7, 96, 54, 16)])
def test_sphere_light_importance_sampling(emul_bin, scene, w, h, spp):
    """sphere::pdf_value / pdf_generate (sphere.cpp:63-79) and random_towards_sphere (pcg.cpp:125-133) are dead in the nine stock
    scenes: the Cornell box and the final scene allocate a light list {ceiling light, glass sphere} but pass count 1
    (scene.cpp:326-329, 456-459).  With both entries (MRT_SCENE_ALL_LIGHTS here, `-lights all` in the oracle harness, which only
    sets the list's count to the allocated 2) the sphere-light path runs on both sides and must agree like any other scene."""
    ref, rmeta = oracle_util.ref_render(scene, w, h, spp, all_lights=True)
    stock, smeta = oracle_util.ref_render(scene, w, h, spp)
    assert rmeta["rays"] != smeta["rays"]            # the second light changes the paths
    acc, meta = oracle_util.emul_render(emul_bin, scene | 0x100, w, h, spp)
    assert meta["rays"] == rmeta["rays"]
    np.testing.assert_array_equal(acc[..., 3], ref[..., 3])
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=1e-5)
    assert res["n_bad"] == 0, res


@needs_ref
def test_triangle_scene_object(emul_bin):
    """triangle_scene_object (triangle.cpp:5-175): a lone triangle as a scene object -- a class of the reference that none of its
    scenes uses.  MRT_SCENE_EXTRA_TRIANGLES / the harness's `-extra triangles` put two of them (face normal, vertex normals) into
    the Cornell box on both sides."""
    ref, rmeta = oracle_util.ref_render(5, 96, 54, 16, extra_triangles=True)
    stock, smeta = oracle_util.ref_render(5, 96, 54, 16)
    assert rmeta["rays"] != smeta["rays"]
    acc, meta = oracle_util.emul_render(emul_bin, 5 | 0x200, 96, 54, 16)
    assert meta["rays"] == rmeta["rays"] and meta["counters"]["tri"] > 0
    np.testing.assert_array_equal(acc[..., 3], ref[..., 3])
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=1e-5)
    assert res["n_bad"] == 0, res
